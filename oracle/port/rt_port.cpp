// TEST INFRASTRUCTURE (oracle) — plain scalar C++ restatement of the reference's render path.
//
// Purpose: an oracle that exists wherever this repository is checked out, including machines
// without /root/reference (where oracle/_ref, the reference itself, cannot be built). It is written
// independently of the product's device code: recursive like the reference, no hoisted constants,
// no culling, one function per reference function, each citing the lines it follows. It consumes
// the same POD scene description as the CUDA library (include/hai719_rt.h), so it also checks
// flatten(). PINNED: tests/test_oracle_port.py compares it bit for bit with oracle/_ref on every
// scene (and with the committed goldens when the reference is absent).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built
// from this file (oracle/port/_build/librt_port.so). The product never does.
//
// Build flags matter: -O2, no -march, no -ffast-math  =>  no FMA, IEEE fp32/fp64, like the reference.
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "hai719_rt.h"

namespace {

// ---- Vec3.h ------------------------------------------------------------------------------------
struct V { float x, y, z; };
inline V mk(float x, float y, float z) { V r = {x, y, z}; return r; }
inline V mk(const float *p) { return mk(p[0], p[1], p[2]); }
inline V operator+(V a, V b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }          // Vec3.h:92
inline V operator-(V a, V b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }          // Vec3.h:95
inline V operator*(float a, V b) { return mk(a * b.x, a * b.y, a * b.z); }            // Vec3.h:98
inline V operator*(V b, float a) { return mk(a * b.x, a * b.y, a * b.z); }            // Vec3.h:101
inline V operator/(V a, float b) { return mk(a.x / b, a.y / b, a.z / b); }            // Vec3.h:104
inline float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }              // Vec3.h:36
inline V cross(V a, V b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }  // Vec3.h:39
inline float sqlen(V a) { return a.x * a.x + a.y * a.y + a.z * a.z; }                 // Vec3.h:29
inline float len(V a) { return (float)std::sqrt((double)sqlen(a)); }                  // Vec3.h:32 (sqrt(double) of a float)
inline V normalize(V a) { const float L = len(a); return mk(a.x / L, a.y / L, a.z / L); }   // Vec3.h:35
inline V cmul(V a, V b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }               // Vec3.h:66

const double EPS = 0.00001;   // Constants.h:17 — a double

struct Ray { V o, d; float time; };
inline Ray make_ray(V o, V d, float time) { Ray r; r.o = o; r.d = normalize(d); r.time = time; return r; }  // Line.h:13-16, Ray.h:8

// ---- deterministic stream (include/hai719_rt.h "Random numbers") -------------------------------
inline uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
struct Rng {
    uint32_t key, ctr;
    float next() { const uint32_t r = fmix32(key + ctr * 0x9E3779B9u); ++ctr; return (float)(r >> 8) * (1.0f / 16777216.0f); }
};
inline Rng make_rng(uint32_t seed, uint32_t pixel, uint32_t sample) {
    Rng r; r.key = fmix32(fmix32(seed ^ ((pixel + 1u) * 0x9E3779B9u)) + (sample + 1u) * 0x85EBCA6Bu); r.ctr = 0; return r;
}
inline float rnd_range(Rng &g, float lo, float hi) { return lo + (hi - lo) * g.next(); }        // Functions.cpp:10-12
inline V random_unit_vector(Rng &g) {                                                           // Functions.cpp:14-18
    const float z = rnd_range(g, -1, 1), y = rnd_range(g, -1, 1), x = rnd_range(g, -1, 1);      // g++: arguments right to left
    return normalize(mk(x, y, z));
}
inline float fmin_(float a, float b) { return a < b ? a : b; }   // Functions.cpp:20
inline float fmax_(float a, float b) { return a > b ? a : b; }   // Functions.cpp:24

struct Scene {
    RtSceneDesc d;
    // deep copies so that the description may go away
    std::vector<RtSphere> spheres; std::vector<RtSquare> squares; std::vector<RtLight> lights;
    std::vector<RtSceneMesh> meshes; std::vector<RtImage> textures, normals;
    std::vector<std::vector<float>> pos, vcol, fcol; std::vector<std::vector<uint32_t>> tri;
    std::vector<std::vector<RtKdNode>> nodes; std::vector<std::vector<RtTriRef>> refs;
    std::vector<std::vector<uint8_t>> tex_px, nrm_px; std::vector<uint8_t> sky_px;
    RtImage sky;
};

// ---- Sphere::intersect (Sphere.h:91-132) -------------------------------------------------------
struct SphereHit { bool exists; float t, theta, phi; V p, n; };
SphereHit sphere_intersect(const RtSphere &s, const Ray &ray) {
    SphereHit r; r.exists = false; r.t = FLT_MAX;
    const V c = mk(s.center) + ray.time * mk(s.material.motion);            // :94
    const V o = ray.o, d = ray.d;
    const float a = dot(d, d);                                              // :100
    const float b = (float)(2. * (double)dot(d, o - c));                    // :101
    const float cc = dot(o - c, o - c) - s.radius * s.radius;               // :102
    const float delta = b * b - 4 * a * cc;                                 // :103
    if (delta < 0) return r;                                                // :105-109
    const float sq = (float)std::sqrt((double)delta);
    float t = (-b - sq) / (2 * a);                                          // :112
    const float t1 = (-b + sq) / (2 * a);
    if ((double)t1 > EPS && t1 < t) t = t1;                                 // :115 (never true)
    if ((double)t < -EPS) return r;                                         // :119
    r.p = o + t * d;
    r.n = normalize(r.p - c);
    r.exists = true; r.t = t;
    r.theta = (float)std::acos((double)r.n.y * -1.);                        // :129
    r.phi = (float)(std::atan2((double)r.n.z * -1., (double)r.n.x) + M_PI); // :130
    return r;
}

// ---- Square::intersect (Square.h:65-126) -------------------------------------------------------
struct SquareHit { bool exists; float t, u, v; V p, n; };
SquareHit square_intersect(const RtSquare &q, const Ray &ray) {
    SquareHit r; r.exists = false; r.t = FLT_MAX;
    const V bl = mk(q.v0) + ray.time * mk(q.material.motion);               // :68
    const V right = mk(q.v1) - mk(q.v0), up = mk(q.v3) - mk(q.v0);          // :69-70
    const V n = normalize(cross(right, up));                                // :71-72
    const float dotRN = dot(ray.d, n);
    if (dotRN == 0) return r;                                               // :77
    if (dotRN > 0 && q.material.type != RT_MAT_GLASS) return r;             // :84
    const float D = dot(bl, n);
    const float t = (D - dot(ray.o, n)) / dotRN;                            // :91
    if ((double)t < -EPS) return r;                                         // :94
    if ((double)t >= EPS) {                                                 // :100
        const V p = ray.o + t * ray.d;
        const V w = p - bl;
        const float proj1 = dot(w, right) / len(right);                     // :106
        const float proj2 = dot(w, up) / len(up);                           // :110
        if ((proj1 <= len(right) && proj1 >= 0) && (proj2 <= len(up) && proj2 >= 0)) {   // :112
            r.exists = true; r.t = t; r.u = proj1 / len(right); r.v = proj2 / len(up); r.p = p; r.n = n;
        }
    }
    return r;
}

// ---- AABB::intersects (AABB.h:48-65) -----------------------------------------------------------
bool aabb_intersects(const float *p0, const float *p1, const Ray &ray) {
    float tmin = (float)EPS, tmax = FLT_MAX;
    const float o[3] = {ray.o.x, ray.o.y, ray.o.z}, d[3] = {ray.d.x, ray.d.y, ray.d.z};
    for (int axis = 0; axis < 3; axis++) {
        const double adinv = 1.0 / d[axis];
        const float t0 = (p0[axis] - o[axis]) * adinv;
        const float t1 = (p1[axis] - o[axis]) * adinv;
        if (t0 < t1) { if (t0 > tmin) tmin = t0; if (t1 < tmax) tmax = t1; }
        else         { if (t1 > tmin) tmin = t1; if (t0 < tmax) tmax = t0; }
        if (tmax <= tmin) return false;
    }
    return true;
}

// ---- Triangle (Triangle.h:26-37, 62-126), built per ray per triangle like KDTree.cpp:38-40 -------
struct TriHit { bool exists; float t, w0, w1, w2; uint32_t tIndex; V p, n; };
TriHit triangle_intersect(V c0, V c1, V c2, const Ray &ray) {
    TriHit r; r.exists = false; r.t = FLT_MAX; r.tIndex = 0;
    const V nn = cross(c1 - c0, c2 - c0);                                   // :32
    const float norm = len(nn);
    const V n = nn / norm;                                                  // :34 (0/0 = NaN for zero-area triangles)
    const float dotRN = dot(ray.d, n);
    if (dotRN == 0) return r;                                               // :81
    if (dotRN > 0) return r;                                                // :88
    const float D = dot(c0, n);
    const float t = (D - dot(ray.o, n)) / dotRN;                            // :96
    if (t < 0) return r;                                                    // :97
    const V p = ray.o + t * ray.d;
    const V v0 = c1 - c0, v1 = c2 - c0, v2 = p - c0;                        // :63-65
    const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    const float denom = d00 * d11 - d01 * d01;
    const float u1 = (d11 * d20 - d01 * d21) / denom, u2 = (d00 * d21 - d01 * d20) / denom, u0 = 1 - u1 - u2;   // :72-74
    if (u0 >= 0 && u0 <= 1 && u1 >= 0 && u1 <= 1 && u2 >= 0 && u2 <= 1) {   // :112
        r.exists = true; r.t = t; r.w0 = u0; r.w1 = u1; r.w2 = u2; r.p = p; r.n = n;
    }
    return r;
}

// ---- KDTree::Node::intersect (KDTree.cpp:31-69) over the flattened pre-order nodes --------------
// children of node i: left = i+1 if it lies inside the subtree, right = where the left subtree ends
TriHit node_intersect(const Scene &sc, int mi, uint32_t i, const Ray &ray) {
    const RtSceneMesh &m = sc.meshes[mi];
    const RtKdNode &n = sc.nodes[mi][i];
    TriHit none; none.exists = false; none.t = FLT_MAX; none.tIndex = 0;
    if (!aabb_intersects(n.bmin, n.bmax, ray)) return none;                 // :32
    if (n.is_leaf || n.skip == i + 1) {                                     // leaf() — also an inner node that lost both children
        TriHit best = none;
        if (!n.is_leaf) return best;
        for (uint32_t k = n.first_ref; k < n.first_ref + n.n_refs; k++) {   // :37-46
            const RtTriRef &t = sc.refs[mi][k];
            const float s = 1.000001f;                                      // TRIANGLE_SCALING, Mesh.h:23
            TriHit h = triangle_intersect(mk(m.positions + 3 * t.v[0]) * s, mk(m.positions + 3 * t.v[1]) * s, mk(m.positions + 3 * t.v[2]) * s, ray);
            if (h.t < best.t) { best = h; best.tIndex = t.tri_index; }
        }
        return best;
    }
    // the flattened array cannot tell an only-left child from an only-right child; both orders give the
    // same result because the missing side contributes t = FLT_MAX and "left.t < right.t ? left : right"
    // then returns the present side either way (ties at FLT_MAX carry no hit)
    TriHit left = none, right = none;
    const uint32_t l = i + 1;
    left = node_intersect(sc, mi, l, ray);                                  // :52
    const uint32_t r = sc.nodes[mi][l].skip;
    if (r < n.skip) right = node_intersect(sc, mi, r, ray);                 // :58
    return left.t < right.t ? left : right;                                 // :63-67
}
TriHit mesh_intersect(const Scene &sc, int mi, const Ray &ray) {            // Mesh.cpp:112-117, KDTree.cpp:80-85
    const RtSceneMesh &m = sc.meshes[mi];
    TriHit none; none.exists = false; none.t = FLT_MAX; none.tIndex = 0;
    if (m.n_nodes == 0) return none;
    if (!aabb_intersects(m.root_bmin, m.root_bmax, ray)) return none;
    return node_intersect(sc, mi, 0, ray);
}

// ---- Scene::computeIntersection (Scene.h:202-230) ----------------------------------------------
struct SceneHit { unsigned type; int obj; float t; SphereHit sp; SquareHit sq; TriHit tr; };
SceneHit compute_intersection(const Scene &sc, const Ray &ray) {
    SceneHit res; res.type = 0; res.obj = -1; res.t = FLT_MAX;
    for (size_t i = 0; i < sc.spheres.size(); i++) {
        SphereHit h = sphere_intersect(sc.spheres[i], ray);
        if (h.exists && h.t < res.t && (double)h.t >= EPS) { res.type = 1; res.obj = (int)i; res.t = h.t; res.sp = h; }
    }
    for (size_t i = 0; i < sc.squares.size(); i++) {
        SquareHit h = square_intersect(sc.squares[i], ray);
        if (h.exists && h.t < res.t && (double)h.t >= EPS) { res.type = 2; res.obj = (int)i; res.t = h.t; res.sq = h; }
    }
    for (size_t i = 0; i < sc.meshes.size(); i++) {
        TriHit h = mesh_intersect(sc, (int)i, ray);
        if (h.exists && h.t < res.t && (double)h.t >= EPS) { res.type = 3; res.obj = (int)i; res.t = h.t; res.tr = h; }
    }
    return res;
}

// ---- Scene::computeShadow (Scene.h:235-255) ----------------------------------------------------
bool compute_shadow(const Scene &sc, const Ray &ray, float t, Rng &g) {
    for (size_t i = 0; i < sc.spheres.size(); i++) {
        SphereHit h = sphere_intersect(sc.spheres[i], ray);
        if (h.exists && h.t < t && (double)h.t >= EPS) if (g.next() > sc.spheres[i].material.transparency) return true;
    }
    for (size_t i = 0; i < sc.squares.size(); i++) {
        SquareHit h = square_intersect(sc.squares[i], ray);
        if (h.exists && h.t < t && (double)h.t >= EPS) if (g.next() > sc.squares[i].material.transparency) return true;
    }
    for (size_t i = 0; i < sc.meshes.size(); i++) {
        TriHit h = mesh_intersect(sc, (int)i, ray);
        if (h.exists && h.t < t && (double)h.t >= EPS) if (g.next() > sc.meshes[i].material.transparency) return true;
    }
    return false;
}

// ---- Material (Material.cpp:13-130) ------------------------------------------------------------
struct Mat { RtMaterial m; V kd; };
int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
void mat_texture(const Scene &sc, const RtMaterial &m, V &color, float u, float v) {            // :63-92
    if (m.texture_type == RT_TEX_CHECKER) {
        color = ((int)(u * m.texture_scale_x) % 2 == (int)(v * m.texture_scale_y) % 2) ? mk(m.checker1) : mk(m.checker2);
    } else if (m.texture_type == RT_TEX_IMAGE) {
        RtImage im = {0, 0, nullptr};
        if (m.image >= 0) im = sc.textures[m.image];
        if (im.w < 1 || im.h < 1) { color = ((int)(u * 8.) % 2 == (int)(v * 8.) % 2) ? mk(0, 0, 0) : mk(1, 0, 1); return; }
        u = std::fmod(u * m.texture_scale_x, 1.);
        v = 1 - std::fmod(v * m.texture_scale_y, 1.);
        const int x = int(u * (im.w - 1)), y = int(v * (im.h - 1));
        const int index = clampi(y * im.w + x, 0, im.w * im.h - 1);     // the reference does not clamp (never out of range for u,v in [0,1])
        const uint8_t *p = im.rgb + 3 * (size_t)index;
        color = mk(p[0] / 255., p[1] / 255., p[2] / 255.);
    }
}
V mat_emit(const Scene &sc, const RtMaterial &m, float u, float v) {                             // :13-24
    if (!m.emissive) return mk(0, 0, 0);
    V c = mk(0, 0, 0);
    if (m.texture_type == RT_TEX_NONE) c = mk(m.light_color); else mat_texture(sc, m, c, u, v);
    return c * m.light_intensity;
}
void mat_get_normal(const Scene &sc, const RtMaterial &m, V &normal, float u, float v, V T, V B) {   // :114-130
    if (m.normal_map < 0) return;
    const RtImage im = sc.normals[m.normal_map];
    if (im.w < 1 || im.h < 1) return;
    u = std::fmod(u * m.texture_scale_x, 1.);
    v = 1 - std::fmod(v * m.texture_scale_y, 1.);
    const int x = int(u * (im.w - 1)), y = int(v * (im.h - 1));
    const uint8_t *p = im.rgb + 3 * (size_t)clampi(y * im.w + x, 0, im.w * im.h - 1);
    const V nm = mk(p[0] / 127.5 - 1., p[1] / 127.5 - 1., p[2] / 127.5 - 1.);
    normal = normalize(nm.x * T + nm.y * B + nm.z * normal);
}
V reflect_(V d, V n) { return d - 2 * dot(d, n) * n; }                                           // Functions.cpp:38-40
V refract_(V d, V n, float eta) {                                                                // Functions.cpp:42-47
    const float cos_theta = fmin_(dot(d, n), 1.0);
    const V perp = eta * (d + cos_theta * n);
    const V par = (float)(-std::sqrt(std::fabs(1.0 - sqlen(perp)))) * n;
    return perp + par;
}
float reflectance_(float cosine, float ref_idx) {                                                // Functions.cpp:49-54
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1 - r0) * std::pow((1 - cosine), 5);
}
Ray mat_scatter(const RtMaterial &m, const Ray &in, V normal, V P, Rng &g) {                     // Material.cpp:26-60
    V dir = mk(0, 0, 0);
    if (m.type == RT_MAT_GLASS) {
        float ri;
        if (dot(in.d, normal) > 0) ri = 1. / m.index_medium; else ri = m.index_medium;
        const float cos_theta = fmin_(dot(in.d * -1., normal), 1.0);
        const float sin_theta = std::sqrt(1. - cos_theta * cos_theta);
        const bool cannot_refract = (ri * sin_theta) - 0.6 > 1.0;
        if (cannot_refract || reflectance_(cos_theta, ri) > g.next()) dir = reflect_(in.d, normal);
        else dir = refract_(in.d, normal, ri);
    } else if (m.type == RT_MAT_DIFFUSE) {
        dir = normal + random_unit_vector(g);
        if ((double)len(dir) <= EPS) dir = normal;
    } else if (m.type == RT_MAT_MIRROR) {
        dir = reflect_(in.d, normal);
    }
    dir = normalize(dir);
    return make_ray(P + (float)EPS * dir, dir, in.time);
}

// ---- Scene::skyboxTexture (Scene.h:149-161) ----------------------------------------------------
V skybox(const Scene &sc, V d, int N) {
    if (sc.sky.w < 1 || sc.sky.h < 1) {
        if (sc.d.dark_sky) return mk(0, 0, 0);
        const float a = 0.5 * (d.y + 1.0);
        return (float)(1.0 - a) * mk(1, 1, 1) + (a * mk(0.5f, 0.7f, 1.0f)) * (float)(N + 1);
    }
    // atan2f / asinf pinned to the correctly rounded value, like oracle/libm_pin.cpp
    const float at = (float)std::atan2((double)d.z, (double)d.x), as = (float)std::asin((double)d.y);
    const float u = 0.5 + at / (2 * M_PI);
    const float v = 0.5 - as / M_PI;
    const int x = u * sc.sky.w, y = v * sc.sky.h;
    const uint8_t *p = sc.sky.rgb + 3 * (size_t)clampi(y * sc.sky.w + x, 0, sc.sky.w * sc.sky.h - 1);
    return mk(p[0] / 255., p[1] / 255., p[2] / 255.) * (float)N;
}

// ---- Scene::rayTraceRecursive (Scene.h:258-342) ------------------------------------------------
V ray_trace_recursive(const Scene &sc, Ray ray, int N, Rng &g, int nb_ech) {
    V color = mk(0, 0, 0);
    if (N == 0) return color;
    const SceneHit hit = compute_intersection(sc, ray);
    RtMaterial mat;
    V normal, P, emission = mk(0, 0, 0), kd;
    switch (hit.type) {
        case 1: {
            mat = sc.spheres[hit.obj].material; kd = mk(mat.diffuse);
            P = hit.sp.p; normal = hit.sp.n;
            if (mat.texture_type != RT_TEX_NONE) mat_texture(sc, mat, kd, hit.sp.phi / (2 * M_PI), hit.sp.theta / M_PI);   // sphere_texture
            emission = mat_emit(sc, mat, hit.sp.phi / (2 * M_PI), hit.sp.theta / M_PI);
            break;
        }
        case 2: {
            mat = sc.squares[hit.obj].material; kd = mk(mat.diffuse);
            P = hit.sq.p; normal = hit.sq.n;
            mat_texture(sc, mat, kd, hit.sq.u, hit.sq.v);
            mat_get_normal(sc, mat, normal, hit.sq.u, hit.sq.v, mk(sc.squares[hit.obj].right), mk(sc.squares[hit.obj].up));
            emission = mat_emit(sc, mat, hit.sq.u, hit.sq.v);
            break;
        }
        case 3: {
            const RtSceneMesh &m = sc.meshes[hit.obj];
            mat = m.material; kd = mk(mat.diffuse);
            P = hit.tr.p; normal = hit.tr.n;
            if (m.color_type == RT_COLOR_VERTEX) {
                const uint32_t *t = m.triangles + 3 * hit.tr.tIndex;
                kd = hit.tr.w0 * mk(m.vert_colors + 3 * t[0]) + hit.tr.w1 * mk(m.vert_colors + 3 * t[1]) + hit.tr.w2 * mk(m.vert_colors + 3 * t[2]);
            } else if (m.color_type == RT_COLOR_FACE) {
                kd = mk(m.face_colors + 3 * hit.tr.tIndex);
            }
            break;
        }
        default:
            return skybox(sc, ray.d, N);
    }
    for (size_t i = 0; i < sc.lights.size(); i++) {                                               // :305-334
        const V lp = mk(sc.lights[i].pos);
        V L = normalize(lp - P);
        const float dotLN = dot(L, normal);
        color = color + cmul(mk(sc.lights[0].color), kd) * fmax_(0.0, dotLN) * (float)(1. - mat.transparency);
        int blocked = 0;
        const float delta = sc.lights[i].radius / 2.;
        for (int j = 0; j < nb_ech; j++) {
            const V lj = lp + random_unit_vector(g) * delta;
            L = normalize(lj - P);
            const float tLight = len(lj - P);
            if (compute_shadow(sc, make_ray(P + L * (float)EPS, L, ray.time), tLight, g)) blocked++;
        }
        const float shadow = 1. - float(blocked) / float(nb_ech);
        color = color * shadow;
    }
    Ray next = mat_scatter(mat, ray, normal, P, g);
    next.time = ray.time;
    const V deeper = cmul(ray_trace_recursive(sc, next, N - 1, g, nb_ech), kd);
    return color + deeper + emission;
}

// ---- MatrixUtilities (matrixUtilities.h:53-74, 210-216) ----------------------------------------
void mult4(const double *m, double x, double y, double z, double w, double *r) {
    r[0] = m[0] * x + m[4] * y + m[8] * z + m[12] * w;
    r[1] = m[1] * x + m[5] * y + m[9] * z + m[13] * w;
    r[2] = m[2] * x + m[6] * y + m[10] * z + m[14] * w;
    r[3] = m[3] * x + m[7] * y + m[11] * z + m[15] * w;
}
Ray camera_ray(const RtCamera &c, float u, float v, float time) {
    double a[4], b[4], p0[4];
    mult4(c.modelview_inverse, 0.0, 0.0, 0.0, 1.0, p0);
    const V pos = mk((float)(p0[0] / p0[3]), (float)(p0[1] / p0[3]), (float)(p0[2] / p0[3]));
    mult4(c.projection_inverse, (double)2.f * u - 1.f, -((double)2.f * v - 1.f), c.depth_near, 1.0, a);
    mult4(c.modelview_inverse, a[0], a[1], a[2], a[3], b);
    const V p = mk((float)(b[0] / b[3]), (float)(b[1] / b[3]), (float)(b[2] / b[3]));
    return make_ray(pos, normalize(p - pos), time);
}

template <class T> void own(std::vector<T> &dst, const T *src, size_t n) { dst.assign(src, src + n); }

}  // namespace

extern "C" {

void *port_scene_create(const RtSceneDesc *d) {
    Scene *s = new Scene;
    s->d = *d;
    own(s->spheres, d->spheres, d->n_spheres); own(s->squares, d->squares, d->n_squares); own(s->lights, d->lights, d->n_lights);
    own(s->meshes, d->meshes, d->n_meshes);
    const size_t nm = d->n_meshes;
    s->pos.resize(nm); s->vcol.resize(nm); s->fcol.resize(nm); s->tri.resize(nm); s->nodes.resize(nm); s->refs.resize(nm);
    for (size_t i = 0; i < nm; i++) {
        RtSceneMesh &m = s->meshes[i];
        own(s->pos[i], m.positions, 3 * (size_t)m.n_vertices); m.positions = s->pos[i].data();
        own(s->tri[i], m.triangles, 3 * (size_t)m.n_triangles); m.triangles = s->tri[i].data();
        if (m.vert_colors) { own(s->vcol[i], m.vert_colors, 3 * (size_t)m.n_vertices); m.vert_colors = s->vcol[i].data(); }
        if (m.face_colors) { own(s->fcol[i], m.face_colors, 3 * (size_t)m.n_triangles); m.face_colors = s->fcol[i].data(); }
        own(s->nodes[i], m.nodes, m.n_nodes); own(s->refs[i], m.leaf_refs, m.n_leaf_refs);
    }
    auto img = [](const RtImage &im, std::vector<uint8_t> &store) {
        RtImage r = {0, 0, nullptr};
        if (im.w >= 1 && im.h >= 1 && im.rgb) { store.assign(im.rgb, im.rgb + (size_t)im.w * im.h * 3); r.w = im.w; r.h = im.h; r.rgb = store.data(); }
        return r;
    };
    s->tex_px.resize(d->n_textures); s->nrm_px.resize(d->n_normal_maps);
    for (uint32_t i = 0; i < d->n_textures; i++) s->textures.push_back(img(d->textures[i], s->tex_px[i]));
    for (uint32_t i = 0; i < d->n_normal_maps; i++) s->normals.push_back(img(d->normal_maps[i], s->nrm_px[i]));
    s->sky = img(d->skybox, s->sky_px);
    return s;
}
void port_scene_destroy(void *h) { delete (Scene *)h; }

// trace_line / ray_trace_from_camera (main.cpp:183-249) for the rectangle of `p`; outputs row-major
// (linear = before gamma, gamma = after, ids = {type, obj, tIndex, bits(t)} of sample 0's camera ray)
void port_render(void *h, const RtCamera *cam, const RtRenderParams *p, float *linear, float *gamma, uint32_t *ids, int threads) {
    const Scene &sc = *(const Scene *)h;
    const bool full = (p->x0 | p->y0 | p->x1 | p->y1) == 0;
    const int x0 = full ? 0 : p->x0, y0 = full ? 0 : p->y0, x1 = full ? p->width : p->x1, y1 = full ? p->height : p->y1;
    const int rw = x1 - x0, rh = y1 - y0;
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int r = next.fetch_add(1);
            if (r >= rh) break;
            const int y = y0 + r;
            for (int x = x0; x < x1; x++) {
                const size_t o = (size_t)r * rw + (x - x0);
                V acc = mk(0, 0, 0);
                for (int s = 0; s < p->spp; s++) {
                    Rng g = make_rng(p->seed, (uint32_t)x + (uint32_t)y * (uint32_t)p->width, (uint32_t)s);
                    const float u = ((float)(x) + g.next()) / p->width;                        // main.cpp:189
                    const float v = ((float)(y) + g.next()) / p->height;                       // main.cpp:190
                    const float time = g.next();
                    const Ray ray = camera_ray(*cam, u, v, time);                               // main.cpp:191-192
                    if (s == 0 && ids) {
                        const SceneHit hit = compute_intersection(sc, ray);
                        uint32_t *q = ids + 4 * o, tb; std::memcpy(&tb, &hit.t, 4);
                        q[0] = hit.type; q[1] = hit.type ? (uint32_t)hit.obj : 0u; q[2] = hit.type == 3 ? hit.tr.tIndex : 0u; q[3] = tb;
                    }
                    V c = mk(0, 0, 0) + ray_trace_recursive(sc, ray, p->max_bounces, g, p->nb_ech);   // Scene.h:345-350
                    c = c / (float)p->max_bounces;
                    acc = acc + c;                                                              // main.cpp:193
                }
                acc = acc / (float)(unsigned int)p->spp;                                        // main.cpp:195
                if (linear) { linear[3 * o] = acc.x; linear[3 * o + 1] = acc.y; linear[3 * o + 2] = acc.z; }
                if (gamma) {                                                                    // Functions.cpp:56-60
                    gamma[3 * o] = std::pow(acc.x, 1.0 / 2.2); gamma[3 * o + 1] = std::pow(acc.y, 1.0 / 2.2); gamma[3 * o + 2] = std::pow(acc.z, 1.0 / 2.2);
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
}

}  // extern "C"
