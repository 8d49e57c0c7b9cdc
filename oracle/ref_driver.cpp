// TEST INFRASTRUCTURE (oracle) — headless driver around the UNMODIFIED reference sources.
//
// This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs load the library built from it
// (oracle/_ref/libref_det.so, oracle/_ref/libref_stock.so; recipe: oracle/Makefile).
//
// What it restates (the reference's own main.cpp is bound to GLUT and cannot run headless):
//   * trace_line()            main.cpp:183-198   pixel x sample loop, jitter, /nsamples, gamma
//   * ray_trace_from_camera() main.cpp:200-249   camera.apply(), matrix refresh, row threads
//   * main()                  main.cpp:418-432   camera.move(0,0,-3.1), scene table
// Everything below those calls (Scene::rayTrace, intersections, KD-tree, materials, camera
// matrices, loaders, scene builders) is the reference's code, compiled from
// /root/reference/src where it lies, with the reference's flags (-O3, no -march, no
// -ffast-math  =>  no FMA; Makefile:26-27).
//
// Built with -fno-access-control so the driver can read Scene's containers (private by
// class default, Scene.h:57-65) for the canonical scene dump and for composing config 5.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <random>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

// NOTE on include order: <math.h> (libstdc++'s wrapper, which injects the float overloads of
// cos/sin/... into the global namespace) must NOT be included before the reference headers:
// Mesh::rotate_* (Mesh.h:198-224) call cos()/sin() unqualified and, in the reference's own
// translation units, bind to the double versions.
#include "Vec3.h"
#include "Camera.h"
#include "Scene.h"
#include "matrixUtilities.h"
#include "Functions.h"
#include "Constants.h"
// KDTree::Node is an incomplete type outside KDTree.cpp (KDTree.hpp:18); the dump walks the
// tree, so the reference file is compiled as part of this translation unit.
#include "KDTree.cpp"

#include "det_rng.h"

// ---- ray counting (libref_count.so only; recipe in oracle/Makefile) ------------------------------------------------
// The COUNTING build compiles this translation unit (which holds the reference's header-only Scene, Scene.h) with
// -finstrument-functions: gcc then calls __cyg_profile_func_enter at the entry of every function of the listed files,
// inlined or not. The hook compares the function address with Scene::computeIntersection (Scene.h:202) and
// Scene::computeShadow (Scene.h:235) and counts: the reference's own calls, no source line of the reference touched.
// The counting library is never timed (the hook slows it down); with the deterministic stream it traces exactly the
// rays the timed deterministic library traces, so its counts are that run's numerator.
#ifdef ORACLE_COUNT_RAYS
#pragma GCC diagnostic ignored "-Wpmf-conversions"
namespace oracle {
static thread_local uint64_t t_closest = 0, t_shadow = 0;
static void *fn_closest = nullptr, *fn_shadow = nullptr;
static std::atomic<uint64_t> g_closest(0), g_shadow(0);
}
extern "C" {
__attribute__((no_instrument_function)) void __cyg_profile_func_enter(void *fn, void *) {
    if (fn == oracle::fn_closest) ++oracle::t_closest;
    else if (fn == oracle::fn_shadow) ++oracle::t_shadow;
}
__attribute__((no_instrument_function)) void __cyg_profile_func_exit(void *, void *) {}
}
#define ORACLE_NOINSTR __attribute__((no_instrument_function))
#else
#define ORACLE_NOINSTR
#endif

namespace oracle {
static thread_local DetCtx g_ctx = {0u, 0u, 0ull};
DetCtx &ctx() { return g_ctx; }
float det_next() {
    DetCtx &c = g_ctx;
    c.total++;
    return draw(c.key, c.ctr++);
}
}  // namespace oracle

namespace {

struct RefScene {
    Scene *scene;       // placement-new'ed into zeroed storage (skybox.w/h etc.)
    void *storage;
    int id;
};

// Overwrite the stack region the next call will use with zeros, so that the reference's
// `Material white = Material();` (Scene.h:477,481 — members the constructor never sets)
// reads zeros instead of stale stack bytes. Identical policy to the product's host API,
// which zero-initialises Material.
__attribute__((noinline)) void scrub_stack() {
    volatile char pad[512 * 1024];
    for (size_t i = 0; i < sizeof pad; ++i) pad[i] = 0;
    __asm__ volatile("" ::: "memory");
}

void add_mesh(Scene &s, const char *file, bool unit, Vec3 scale, float ry, Vec3 t, Vec3 kd) {
    s.meshes.resize(s.meshes.size() + 1);
    Mesh &m = s.meshes.back();
    m.loadOFF(file);
    if (unit) m.centerAndScaleToUnit();
    m.scale(scale);
    m.rotate_y(ry);
    m.translate(t);
    m.build_arrays();
    m.material.diffuse_material = kd;
    m.material.specular_material = Vec3(0.9f, 0.9f, 0.9f);
    m.material.shininess = 6.;
}

// BASELINE.json config 5 has no builder in the reference: it is composed
// through the reference's own classes — setup_random_spheres (motion blur on the 79 random
// spheres, Scene.h:922) plus triceratops.off and gorilla.off with KD-trees.
void setup_config5(Scene &s) {
    s.setup_random_spheres();
    add_mesh(s, "mesh/triceratops.off", true, Vec3(2.0f), 200.f, Vec3(-3.2f, -2.6f, -3.5f),
             Vec3(0.35f, 0.55f, 0.25f));
    add_mesh(s, "mesh/gorilla.off", true, Vec3(1.8f), 160.f, Vec3(3.0f, -2.3f, -4.5f),
             Vec3(0.45f, 0.35f, 0.3f));
    s.computeKDTrees();
}

uint64_t fnv1a(const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

struct Dump {
    std::vector<uint32_t> w;
    void u(uint32_t v) { w.push_back(v); }
    void f(float v) { uint32_t b; std::memcpy(&b, &v, 4); w.push_back(b); }
    void v3(const Vec3 &v) { f(v[0]); f(v[1]); f(v[2]); }
    void img(const ppmLoader::ImageRGB &im) {
        if (im.w < 1 || im.h < 1 || im.data.size() < (size_t)im.w * im.h) { u(0); u(0); u(0); u(0); return; }
        uint64_t h = fnv1a(im.data.data(), (size_t)im.w * im.h * 3);
        u(im.w); u(im.h); u((uint32_t)h); u((uint32_t)(h >> 32));
    }
};

int index_of(const std::vector<ppmLoader::ImageRGB> &v, const ppmLoader::ImageRGB *p) {
    if (!p || v.empty()) return -1;
    ptrdiff_t d = p - v.data();
    return (d >= 0 && (size_t)d < v.size()) ? (int)d : -1;
}

void dump_material(Dump &d, const Scene &s, const Material &m) {
    d.u((uint32_t)m.type);
    d.u((uint32_t)m.texture_type);
    d.v3(m.diffuse_material);
    d.f(m.transparency);
    d.f(m.index_medium);
    d.v3(m.checkerboard_color1);
    d.v3(m.checkerboard_color2);
    d.f(m.texture_scale_x);
    d.f(m.texture_scale_y);
    unsigned char eb; std::memcpy(&eb, &m.emissive, 1);
    const bool em = eb != 0;
    d.u(em ? 1u : 0u);
    if (em) { d.v3(m.light_color); d.f(m.light_intensity); } else { d.v3(Vec3(0.f)); d.f(0.f); }
    d.u((uint32_t)(m.texture_type == Texture_Image ? index_of(s.textures, m.image) : -1));
    d.u((uint32_t)(m.has_normal_map ? index_of(s.normals, m.normals) : -1));
    d.u(m.has_normal_map ? 1u : 0u);
    d.v3(m.motion_blur_translation);
}

void dump_node(Dump &d, const KDTree::Node *n, uint32_t &count) {
    ++count;
    d.v3(n->aabb.p0);
    d.v3(n->aabb.p1);
    d.u((n->leaf() ? 1u : 0u) | (n->left ? 2u : 0u) | (n->right ? 4u : 0u));
    d.u((uint32_t)n->triangles.size());
    for (const MeshTriangle &t : n->triangles) { d.u(t.v[0]); d.u(t.v[1]); d.u(t.v[2]); d.u(t.v[3]); }
    if (n->left) dump_node(d, n->left, count);
    if (n->right) dump_node(d, n->right, count);
}

void dump_scene(Dump &d, const Scene &s) {
    d.u(0x44533748u);  // 'H7SD'
    d.u(1u);
    d.u(s.dark_sky ? 1u : 0u);
    d.img(s.skybox);
    d.u((uint32_t)s.textures.size());
    for (auto &t : s.textures) d.img(t);
    d.u((uint32_t)s.normals.size());
    for (auto &t : s.normals) d.img(t);
    d.u((uint32_t)s.lights.size());
    for (auto &l : s.lights) { d.v3(l.pos); d.f(l.radius); d.v3(l.material); }
    d.u((uint32_t)s.spheres.size());
    for (auto &sp : s.spheres) { d.v3(sp.m_center); d.f(sp.m_radius); dump_material(d, s, sp.material); }
    d.u((uint32_t)s.squares.size());
    for (auto &sq : s.squares) {
        for (int k = 0; k < 4; ++k) d.v3(sq.vertices[k].position);
        d.v3(sq.m_right_vector);
        d.v3(sq.m_up_vector);
        dump_material(d, s, sq.material);
    }
    d.u((uint32_t)s.meshes.size());
    for (auto &m : s.meshes) {
        d.u((uint32_t)m.vertices.size());
        d.u((uint32_t)m.triangles.size());
        d.u((uint32_t)m.colorType);
        d.u(m.kdtree ? 1u : 0u);
        d.v3(m.aabb.p0);
        d.v3(m.aabb.p1);
        dump_material(d, s, m.material);
        for (auto &v : m.vertices) d.v3(v.position);
        for (auto &t : m.triangles) { d.u(t.v[0]); d.u(t.v[1]); d.u(t.v[2]); }
        if (m.colorType == ColorType_Vertex) for (auto &c : m.vertColors) d.v3(c);
        if (m.colorType == ColorType_Face) for (auto &c : m.faceColors) d.v3(c);
        if (m.kdtree) {
            d.v3(m.kdtree->aabb.p0);
            d.v3(m.kdtree->aabb.p1);
            size_t at = d.w.size();
            d.u(0u);
            uint32_t count = 0;
            if (m.kdtree->root) dump_node(d, m.kdtree->root, count);
            d.w[at] = count;
        }
    }
}

struct CameraState {
    MatrixUtilities mu;
};

// main.cpp:418 + Camera.cpp:46-56,125-132 + main.cpp:212-214
void make_camera(int w, int h, MatrixUtilities &mu) {
    Camera camera;
    camera.resize(w, h);
    camera.move(0., 0., -3.1);
    glMatrixMode(GL_MODELVIEW);
    camera.apply();
    mu.updated();
    mu.updateMatrices();
}

inline uint32_t fbits(float f) { uint32_t b; std::memcpy(&b, &f, 4); return b; }

}  // namespace

extern "C" {

// Returns an opaque handle, or null. asset_root must contain img/ and mesh/ laid out as in the
// reference checkout (the builders use relative paths, Scene.h:360,424,...). A missing .off
// makes the reference call exit(1) (Mesh.cpp:11-13): callers check files first.
void *ref_scene_create(int scene_id, float aspect, uint32_t seed, const char *asset_root) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return nullptr;
    if (asset_root && chdir(asset_root) != 0) return nullptr;
    RefScene *rs = new RefScene;
    rs->storage = std::calloc(1, sizeof(Scene));
    rs->scene = new (rs->storage) Scene();
    rs->id = scene_id;
    Scene &s = *rs->scene;
    // scene-construction stream (Scene.h:895-922 calls random_float() and rand())
    oracle::ctx().key = oracle::path_key(seed, 0xFFFFFFFFu, 0u);
    oracle::ctx().ctr = 0;
    srand(seed);
    scrub_stack();
    switch (scene_id) {
        case 0: s.setup_single_sphere(); break;
        case 1: s.setup_single_square(); break;
        case 2: s.setup_cornell_box(aspect); break;
        case 3: s.setup_mesh(); break;
        case 4: s.setup_rt_in_a_weekend(); break;
        case 5: s.setup_random_spheres(); break;
        case 6: s.setup_debug_refraction(); break;
        case 7: s.setup_flamingo(); break;
        case 8: s.setup_raccoon(); break;
        case 9: s.setup_flamingo_pond(); break;
        case 10: s.setup_backrooms_pool(); break;
        case 11: s.setup_flamingo_lake(); break;
        case 100: setup_config5(s); break;
        default:
            if (chdir(cwd) != 0) {}
            std::free(rs->storage);
            delete rs;
            return nullptr;
    }
    if (chdir(cwd) != 0) {}
    return rs;
}

void ref_scene_destroy(void *h) {
    if (!h) return;
    RefScene *rs = (RefScene *)h;
    // KD-trees are leaked by the reference (no Mesh destructor frees them); so do we.
    rs->scene->~Scene();
    std::free(rs->storage);
    delete rs;
}

// Canonical scene dump (format: tests/scene_dump.py). Call with out == null to get the size.
size_t ref_scene_dump(void *h, uint32_t *out, size_t cap_words) {
    Dump d;
    dump_scene(d, *((RefScene *)h)->scene);
    if (out) std::memcpy(out, d.w.data(), std::min(cap_words, d.w.size()) * 4);
    return d.w.size();
}

// The inverse matrices the reference's ray generation uses for a w x h window.
void ref_camera(int w, int h, double *mv_inv16, double *proj_inv16, double *depth_range2) {
    MatrixUtilities mu;
    make_camera(w, h, mu);
    std::memcpy(mv_inv16, mu.modelviewInverse, 16 * sizeof(double));
    std::memcpy(proj_inv16, mu.projectionInverse, 16 * sizeof(double));
    depth_range2[0] = mu.nearAndFarPlanes[0];
    depth_range2[1] = mu.nearAndFarPlanes[1];
}

// Headless restatement of ray_trace_from_camera()/trace_line() for the pixel rectangle
// [x0,x1) x [y0,y1) of a w x h image. Outputs are (x1-x0)*(y1-y0) pixels, row-major:
//   linear_rgb : sum of samples / nsamples                      (main.cpp:193-195)
//   gamma_rgb  : after gamma_correct                            (main.cpp:196)
//   prim_ids   : per pixel {type, objectIndex, tIndex, bits(t)} of computeIntersection() on
//                sample 0's camera ray (type 0 = miss; tIndex only meaningful for type 3)
// Jitter (u, v, time) are draws 0,1,2 of the path's stream; random_float() continues from 3.
// threads <= 0 means hardware_concurrency. Returns wall seconds spent tracing.
double ref_render(void *h, int w, int ht, int spp, uint32_t seed, int threads, int x0, int y0, int x1,
                  int y1, float *linear_rgb, float *gamma_rgb, uint32_t *prim_ids, uint64_t *n_random) {
    Scene &scene = *((RefScene *)h)->scene;
    MatrixUtilities mu;
    make_camera(w, ht, mu);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    const int cw = x1 - x0, ch = y1 - y0;
    std::atomic<int> next_row(0);
    std::atomic<uint64_t> total_random(0);
#ifdef ORACLE_COUNT_RAYS
    oracle::fn_closest = (void *)(&Scene::computeIntersection);
    oracle::fn_shadow = (void *)(&Scene::computeShadow);
    oracle::g_closest = 0; oracle::g_shadow = 0;
#endif
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        MatrixUtilities lmu = mu;  // flags are clear: no GL reads on worker threads
        oracle::ctx().total = 0;
#ifdef ORACLE_COUNT_RAYS
        oracle::t_closest = 0; oracle::t_shadow = 0;
#endif
        for (;;) {
            int r = next_row.fetch_add(1);
            if (r >= ch) break;
            const int y = y0 + r;
            for (int x = x0; x < x1; ++x) {
                Vec3 acc(0, 0, 0);
                const uint32_t pixel = (uint32_t)x + (uint32_t)y * (uint32_t)w;
                const size_t o = (size_t)(x - x0) + (size_t)r * cw;
                for (int s = 0; s < spp; ++s) {
                    oracle::DetCtx &c = oracle::ctx();
                    c.key = oracle::path_key(seed, pixel, (uint32_t)s);
                    c.ctr = 0;
                    Vec3 pos, dir;
                    float u = ((float)(x) + oracle::det_next()) / w;
                    float v = ((float)(y) + oracle::det_next()) / ht;
                    lmu.screen_space_to_world_space_ray(u, v, pos, dir);
                    Ray ray(pos, dir, oracle::det_next());
                    if (s == 0 && prim_ids) {
#ifdef ORACLE_COUNT_RAYS
                        const uint64_t keep = oracle::t_closest;   // the id probe is the driver's call, not the render's
#endif
                        RaySceneIntersection hit = scene.computeIntersection(ray);
#ifdef ORACLE_COUNT_RAYS
                        oracle::t_closest = keep;
#endif
                        uint32_t *p = prim_ids + 4 * o;
                        p[0] = hit.typeOfIntersectedObject;
                        p[1] = hit.typeOfIntersectedObject ? hit.objectIndex : 0u;
                        p[2] = hit.typeOfIntersectedObject == 3 ? hit.rayMeshIntersection.tIndex : 0u;
                        p[3] = fbits(hit.t);
                    }
                    Vec3 color = scene.rayTrace(ray);
                    acc += color;
                }
                acc /= (unsigned int)spp;
                if (linear_rgb) { linear_rgb[3 * o] = acc[0]; linear_rgb[3 * o + 1] = acc[1]; linear_rgb[3 * o + 2] = acc[2]; }
                gamma_correct(acc);
                if (gamma_rgb) { gamma_rgb[3 * o] = acc[0]; gamma_rgb[3 * o + 1] = acc[1]; gamma_rgb[3 * o + 2] = acc[2]; }
            }
        }
        total_random += oracle::ctx().total;
#ifdef ORACLE_COUNT_RAYS
        oracle::g_closest += oracle::t_closest; oracle::g_shadow += oracle::t_shadow;
#endif
    };
    if (threads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (n_random) *n_random = total_random.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Closest hit of arbitrary rays through Scene::computeIntersection (Scene.h:202-230).
// dirs are normalised by the Ray constructor exactly as the reference does (Line.h:13-16).
// out4 per ray: {type, objectIndex, tIndex, bits(t)}; aux per ray (8 floats, may be null):
//   sphere: theta, phi, n.xyz | square: u, v, n.xyz | mesh: w0, w1, w2, n.xyz (first 6..7 slots)
void ref_trace_rays(void *h, size_t n, const float *org, const float *dir, const float *time, uint32_t *out4,
                    float *aux8) {
    Scene &scene = *((RefScene *)h)->scene;
    for (size_t i = 0; i < n; ++i) {
        Ray ray(Vec3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), Vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]),
                time ? time[i] : 0.f);
        RaySceneIntersection hit = scene.computeIntersection(ray);
        uint32_t *p = out4 + 4 * i;
        p[0] = hit.typeOfIntersectedObject;
        p[1] = hit.typeOfIntersectedObject ? hit.objectIndex : 0u;
        p[2] = hit.typeOfIntersectedObject == 3 ? hit.rayMeshIntersection.tIndex : 0u;
        p[3] = fbits(hit.t);
        if (aux8) {
            float *a = aux8 + 8 * i;
            for (int k = 0; k < 8; ++k) a[k] = 0.f;
            if (p[0] == 1) {
                const RaySphereIntersection &s = hit.raySphereIntersection;
                a[0] = s.theta; a[1] = s.phi; a[2] = s.normal[0]; a[3] = s.normal[1]; a[4] = s.normal[2];
            } else if (p[0] == 2) {
                const RaySquareIntersection &s = hit.raySquareIntersection;
                a[0] = s.u; a[1] = s.v; a[2] = s.normal[0]; a[3] = s.normal[1]; a[4] = s.normal[2];
            } else if (p[0] == 3) {
                const RayTriangleIntersection &s = hit.rayMeshIntersection;
                a[0] = s.w0; a[1] = s.w1; a[2] = s.w2; a[3] = s.normal[0]; a[4] = s.normal[1]; a[5] = s.normal[2];
            }
        }
    }
}

// Scene::rayTrace (Scene.h:345-350) of arbitrary rays with an explicit stream per ray:
// ray i uses key(seed, pixel = i, sample = 0) and starts drawing at counter 3.
void ref_shade_rays(void *h, size_t n, const float *org, const float *dir, const float *time, uint32_t seed,
                    float *rgb) {
    Scene &scene = *((RefScene *)h)->scene;
    for (size_t i = 0; i < n; ++i) {
        oracle::ctx().key = oracle::path_key(seed, (uint32_t)i, 0u);
        oracle::ctx().ctr = 3;
        Ray ray(Vec3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), Vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]),
                time ? time[i] : 0.f);
        Vec3 c = scene.rayTrace(ray);
        rgb[3 * i] = c[0]; rgb[3 * i + 1] = c[1]; rgb[3 * i + 2] = c[2];
    }
}

// Ray counts of the last ref_render call: {computeIntersection calls, computeShadow calls}. Only the counting build
// (libref_count.so) counts; the others return 0 and leave `out2` zeroed.
int ref_ray_counts(uint64_t *out2) {
    out2[0] = 0; out2[1] = 0;
#ifdef ORACLE_COUNT_RAYS
    out2[0] = oracle::g_closest.load(); out2[1] = oracle::g_shadow.load();
    return 1;
#else
    return 0;
#endif
}

// ray_trace_from_camera() AS THE REFERENCE THREADS IT (main.cpp:229-238): one std::thread per scanline, all created at
// once and joined at the end; jitter and time from a thread_local mt19937 seeded by random_device (main.cpp:181,189-192).
// In libref_stock.so random_float() is the reference's own: ONE function-static mt19937 shared by all those threads
// without a lock (Functions.cpp:4-8) — the data race and its cache-line ping-pong are part of what is timed.
// Rows y0..y1-1, columns x0..x1-1 of a w x ht image; gamma_rgb may be null. Returns wall seconds.
double ref_render_rows(void *h, int w, int ht, int spp, int x0, int y0, int x1, int y1, float *gamma_rgb) {
    Scene &scene = *((RefScene *)h)->scene;
    MatrixUtilities mu;
    make_camera(w, ht, mu);
    const int cw = x1 - x0;
    auto t0 = std::chrono::steady_clock::now();
    auto trace_line = [&](int y) {
        static thread_local std::mt19937 rng{std::random_device{}()};
        std::uniform_real_distribution<float> dist(0.f, 1.f);
        MatrixUtilities lmu = mu;
        for (int x = x0; x < x1; ++x) {
            Vec3 acc(0, 0, 0);
            for (int s = 0; s < spp; ++s) {
                float u = ((float)(x) + dist(rng)) / w;
                float v = ((float)(y) + dist(rng)) / ht;
                Vec3 pos, dir;
                lmu.screen_space_to_world_space_ray(u, v, pos, dir);
                acc += scene.rayTrace(Ray(pos, dir, dist(rng)));
            }
            acc /= (unsigned int)spp;
            gamma_correct(acc);
            if (gamma_rgb) {
                const size_t o = (size_t)(x - x0) + (size_t)(y - y0) * cw;
                gamma_rgb[3 * o] = acc[0]; gamma_rgb[3 * o + 1] = acc[1]; gamma_rgb[3 * o + 2] = acc[2];
            }
        }
    };
    std::vector<std::thread> threads;
    for (int y = y0; y < y1; ++y) threads.emplace_back(trace_line, y);
    for (auto &t : threads) t.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_is_deterministic(void) {
#ifdef ORACLE_STOCK_RNG
    return 0;
#else
    return 1;
#endif
}

}  // extern "C"
