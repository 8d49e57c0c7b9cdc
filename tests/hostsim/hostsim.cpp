// TEST-ONLY DEBUG AID — csrc/rt_core.cuh compiled as plain C++ and run on the CPU.
//
// This is NOT a fallback and is never linked into, loaded by, or reachable from the product
// libraries (hai719-raytracing_b200/lib/*.so). It exists because the development container has no
// GPU: running the very same device functions on the host against the oracle catches logic errors
// in rt_core.cuh before GPU time is spent. Only tests/test_hostsim.py builds and loads it.
// It consumes the same RtSceneDesc the CUDA library does and mirrors rt_scene_create's
// precomputation (k_precompute_squares / k_precompute_tris) and k_render_paths / k_resolve.
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

#include "hai719_rt.h"
#include "rt_core.cuh"
#include "rt_pack.hpp"
#include "rt_bvh.hpp"

using namespace rt;

struct SimScene {
    DScene d{};
    std::vector<float4> sph_a, sph_b;
    std::vector<DMaterial> sph_mat, sq_mat, mesh_mat;
    std::vector<DSquare> squares;
    std::vector<float> sq_tr, mesh_tr;
    std::vector<DLight> lights;
    std::vector<DImage> tex, nrm;
    std::vector<DMesh> meshes;
    PackedMeshes pk;
    Accel ac;
    AnalyticAccel aa;
    std::vector<float4> plane, edge, always_bound;
    std::vector<float2> den;
};

static DMaterial to_dmat(const RtMaterial &m) {
    DMaterial d{};
    d.type = m.type; d.texture_type = m.texture_type;
    for (int k = 0; k < 3; ++k) { d.kd[k] = m.diffuse[k]; d.checker1[k] = m.checker1[k]; d.checker2[k] = m.checker2[k]; d.light_color[k] = m.light_color[k]; d.motion[k] = m.motion[k]; }
    d.transparency = m.transparency; d.index_medium = m.index_medium; d.tsx = m.texture_scale_x; d.tsy = m.texture_scale_y;
    d.emissive = m.emissive; d.light_intensity = m.light_intensity; d.image = m.image; d.normal_map = m.normal_map;
    return d;
}
static DImage to_dimg(const RtImage &im) {
    DImage d; d.w = 0; d.h = 0; d.rgb = nullptr;
    if (im.w >= 1 && im.h >= 1 && im.rgb) { d.w = im.w; d.h = im.h; d.rgb = im.rgb; }
    return d;
}

static std::atomic<unsigned long long> g_dbg_lights(0), g_dbg_unoccluded(0), g_dbg_cands(0), g_dbg_overflow(0), g_dbg_tris(0);

extern "C" {

// debug: lights lit in wavefront mode (variant 6), how many needed no shadow sample, candidates of the others
// debug: per mesh {bvh_root, always_count}
int sim_mesh_info(void *h, int *out, int cap) { const SimScene *s = (const SimScene *)h; int n = 0; for (int i = 0; i < s->d.n_meshes && 2 * i + 1 < cap; ++i) { out[2 * i] = s->d.meshes[i].bvh_root; out[2 * i + 1] = (int)s->d.meshes[i].always_count; ++n; } return n; }
void sim_debug_counts(unsigned long long *out5) { out5[0] = g_dbg_lights; out5[1] = g_dbg_unoccluded; out5[2] = g_dbg_cands; out5[3] = g_dbg_overflow; out5[4] = g_dbg_tris; g_dbg_lights = 0; g_dbg_unoccluded = 0; g_dbg_cands = 0; g_dbg_overflow = 0; g_dbg_tris = 0; }


void *sim_scene_create(const RtSceneDesc *desc) {
    SimScene *s = new SimScene;
    DScene &d = s->d;
    d.n_spheres = desc->n_spheres; d.n_squares = desc->n_squares; d.n_meshes = desc->n_meshes; d.n_lights = desc->n_lights;
    d.dark_sky = desc->dark_sky;
    for (uint32_t i = 0; i < desc->n_textures; ++i) s->tex.push_back(to_dimg(desc->textures[i]));
    for (uint32_t i = 0; i < desc->n_normal_maps; ++i) s->nrm.push_back(to_dimg(desc->normal_maps[i]));
    d.textures = s->tex.data(); d.normal_maps = s->nrm.data(); d.sky = to_dimg(desc->skybox);
    for (uint32_t i = 0; i < desc->n_spheres; ++i) {
        const RtSphere &sp = desc->spheres[i];
        s->sph_a.push_back(make_float4(sp.center[0], sp.center[1], sp.center[2], sp.radius));
        s->sph_b.push_back(make_float4(sp.material.motion[0], sp.material.motion[1], sp.material.motion[2], sp.material.transparency));
        s->sph_mat.push_back(to_dmat(sp.material));
    }
    d.sph_a = s->sph_a.data(); d.sph_b = s->sph_b.data(); d.sph_mat = s->sph_mat.data();
    for (uint32_t i = 0; i < desc->n_squares; ++i) {
        const RtSquare &q = desc->squares[i];
        const V3 v0 = ld3(q.v0), right = ld3(q.v1) - v0, up = ld3(q.v3) - v0;
        const V3 n = normalized(cross(right, up));
        DSquare o;
        o.v0[0] = v0.x; o.v0[1] = v0.y; o.v0[2] = v0.z; o.n[0] = n.x; o.n[1] = n.y; o.n[2] = n.z;
        o.right[0] = right.x; o.right[1] = right.y; o.right[2] = right.z; o.up[0] = up.x; o.up[1] = up.y; o.up[2] = up.z;
        o.len_r = length(right); o.len_u = length(up);
        for (int k = 0; k < 3; ++k) { o.motion[k] = q.material.motion[k]; o.tan_r[k] = q.right[k]; o.tan_u[k] = q.up[k]; }
        o.glass = q.material.type == RT_MAT_GLASS;
        s->squares.push_back(o);
        s->sq_mat.push_back(to_dmat(q.material));
        s->sq_tr.push_back(q.material.transparency);
    }
    d.squares = s->squares.data(); d.sq_mat = s->sq_mat.data(); d.sq_transparency = s->sq_tr.data();
    for (uint32_t i = 0; i < desc->n_lights; ++i) {
        DLight l;
        for (int k = 0; k < 3; ++k) { l.pos[k] = desc->lights[i].pos[k]; l.color[k] = desc->lights[i].color[k]; }
        l.radius = desc->lights[i].radius;
        s->lights.push_back(l);
    }
    d.lights = s->lights.data();
    build_analytic_accel(*desc, s->aa);
    d.abvh_root = s->aa.root; d.abvh_n_nodes = (int)(s->aa.nodes.size() / 4); d.abvh_nodes = s->aa.nodes.data(); d.abvh_prims = s->aa.tris.data();
    for (int k = 0; k < 3; ++k) d.abvh_c[k] = s->aa.center[k];
    d.abvh_r = s->aa.radius;
    for (int k = 0; k < 3; ++k) d.abvh_cs[k] = s->aa.center_s[k];
    d.abvh_rs = s->aa.radius_s;
    d.abvh_flat = s->aa.flat.empty() ? nullptr : s->aa.flat.data();
    const uint32_t nm = desc->n_meshes;
    if (!pack_meshes(*desc, s->pk).empty()) { delete s; return nullptr; }
    build_accel(*desc, s->pk, s->ac);
    d.bvh_nodes = s->ac.nodes.data(); d.bvh4_nodes = s->ac.nodes4.data(); d.bvh_tris = s->ac.tris.data();
    d.ref_next = s->ac.ref_next.data(); d.ref_leaf = s->ac.ref_leaf.data(); d.node_parent = s->ac.node_parent.data();
    for (uint32_t i = 0; i < nm; ++i) {
        const RtSceneMesh &src = desc->meshes[i];
        DMesh o;
        memset(&o, 0, sizeof o);
        o.node_begin = s->pk.node_begin[i]; o.node_end = s->pk.node_end[i]; o.color_type = src.color_type;
        o.bvh_root = s->ac.mesh_root[i]; o.bvh4_root = s->ac.mesh_root4[i]; o.always_first = s->ac.always_first[i]; o.always_count = s->ac.always_count[i];
        for (uint32_t k = 0; k < src.n_leaf_refs; ++k) {
            const RtTriRef &r = src.leaf_refs[k];
            const TriConst c = precompute_triangle(ld3(src.positions + 3 * r.v[0]), ld3(src.positions + 3 * r.v[1]), ld3(src.positions + 3 * r.v[2]), r.tri_index);
            s->plane.push_back(c.plane);
            s->edge.push_back(c.c0); s->edge.push_back(c.e0); s->edge.push_back(c.e1);
            s->den.push_back(c.den);
        }
        o.triangles = src.triangles; o.vert_colors = src.vert_colors; o.face_colors = src.face_colors;
        s->meshes.push_back(o);
        s->mesh_mat.push_back(to_dmat(src.material));
        s->mesh_tr.push_back(src.material.transparency);
    }
    d.node_lo = s->pk.lo.data(); d.node_hi = s->pk.hi.data();
    d.tri_plane = s->plane.data(); d.tri_edge = s->edge.data(); d.tri_den = s->den.data();
    s->always_bound.assign(3 * s->ac.tris.size(), make_float4(0.f, 0.f, 0.f, 0.f));
    for (const DMesh &o : s->meshes)
        for (uint32_t k = o.always_first; k < o.always_first + o.always_count; ++k) {
            const uint32_t r0 = s->ac.tris[k];
            const AlwaysBound ab = always_bound_of(s->edge[3 * r0], s->edge[3 * r0 + 1], s->edge[3 * r0 + 2], s->den[r0].x);
            s->always_bound[3 * k] = ab.g1; s->always_bound[3 * k + 1] = ab.g2; s->always_bound[3 * k + 2] = ab.b;
        }
    d.always_bound = s->always_bound.empty() ? nullptr : s->always_bound.data();
    d.meshes = s->meshes.data(); d.mesh_mat = s->mesh_mat.data(); d.mesh_transparency = s->mesh_tr.data();
    return s;
}

void sim_scene_destroy(void *h) { delete (SimScene *)h; }

// rt_bvh.hpp's host replay of the stored barycentric denominator against precompute_triangle (the function the device runs, which this file
// compiles for the host): every leaf reference of every mesh of the scene; returns the number of references whose bits differ.
int sim_check_den_replay(void *h, const RtSceneDesc *desc, unsigned long long *n_checked) {
    const SimScene *s = (const SimScene *)h;
    int bad = 0;
    size_t ref = 0;
    *n_checked = 0;
    for (uint32_t i = 0; i < desc->n_meshes; ++i) {
        const RtSceneMesh &src = desc->meshes[i];
        for (uint32_t k = 0; k < src.n_leaf_refs; ++k, ++ref) {
            const RtTriRef &r = src.leaf_refs[k];
            float c[3][3];
            for (int v = 0; v < 3; ++v) for (int a = 0; a < 3; ++a) c[v][a] = 1.000001f * src.positions[3 * r.v[v] + a];
            const float want = s->den[ref].x, got = bvh_detail::stored_denominator(c[0], c[1], c[2]);
            if (f2u(want) != f2u(got)) ++bad;
            ++*n_checked;
        }
    }
    return bad;
}

// The ball the quadratic padding term takes its distance from (AnalyticAccel::center_s / radius_s) must hold every sphere centre at every time in
// [0, 1] (the device evaluates c = centre + time * motion in float): returns the number of (sphere, time) pairs outside it; *slack = the smallest margin seen.
int sim_check_centre_ball(void *h, const RtSceneDesc *desc, float *slack) {
    const SimScene *s = (const SimScene *)h;
    int bad = 0;
    *slack = FLT_MAX;
    if (s->d.abvh_root < 0) { *slack = 0.f; return 0; }   // no analytic hierarchy (too few primitives): nothing is padded
    for (uint32_t i = 0; i < desc->n_spheres; ++i)
        for (int k = 0; k <= 16; ++k) {
            const float time = (float)k / 16.f;
            const float4 a = s->sph_a[i], b = s->sph_b[i];
            const V3 c = v3(a.x, a.y, a.z) + time * v3(b.x, b.y, b.z);
            const float d = length(c - ld3(s->d.abvh_cs));
            if (!(d == d)) continue;   // a non-finite centre: its box is everything, the padding does not matter
            if (d > s->d.abvh_rs) ++bad;
            if (s->d.abvh_rs - d < *slack) *slack = s->d.abvh_rs - d;
        }
    return bad;
}

// ---- property checks of the conservative candidate filters (tests/test_hostsim.py) ---------------------------------
static float urand(uint32_t &st) { st = st * 1664525u + 1013904223u; return (float)(st >> 8) * (1.0f / 16777216.0f); }
static V3 vrand(uint32_t &st, float lo, float hi) { const float x = lo + (hi - lo) * urand(st), y = lo + (hi - lo) * urand(st), z = lo + (hi - lo) * urand(st); return v3(x, y, z); }

// lc_cannot_occlude's hull test (lc_hull_misses) against the reference's sphere test: whenever the filter drops a sphere, no shadow
// sample of that light — generated exactly as path_shadow_sample_body generates them — may get EPSILON < t < t_light from sphere_t.
// Spheres are placed AT the boundary of the hull +- a relative offset between 1e-6 and 1, so both verdicts occur and the margin is
// probed. out[0] = cases dropped, out[1] = cases kept, out[2] = hits among the kept ones (the test is not vacuous); returns violations.
int sim_check_hull(unsigned int seed, int n_cases, int n_rays, unsigned long long *out) {
    uint32_t st = seed * 2654435761u + 12345u;
    int bad = 0;
    out[0] = out[1] = out[2] = 0;
    for (int k = 0; k < n_cases; ++k) {
        const float scale = expf(logf(0.1f) + urand(st) * logf(1000.f));   // scene sizes 0.1 .. 100
        const V3 P = vrand(st, -scale, scale), lp = vrand(st, -scale, scale);
        const float lrad = scale * (0.01f + 0.5f * urand(st));
        const float r = scale * expf(logf(0.002f) + urand(st) * logf(500.f));
        const V3 D = lp - P;
        const float delta = (lrad / 2.f) * 1.0001f + 1e-6f, Dlen = length(D);
        if (!(Dlen > 0.f)) continue;
        // a point on the hull's surface, then the centre at distance r * (1 + e) from it, outwards
        const float s_ax = urand(st) * 1.2f - 0.1f;
        const V3 axis = D / Dlen;
        V3 perp = cross(axis, vrand(st, -1.f, 1.f));
        if (!(length(perp) > 1e-3f)) continue;
        perp = perp / length(perp);
        const float sgn = urand(st) < 0.5f ? -1.f : 1.f;
        const float e = sgn * expf(logf(1e-6f) + urand(st) * logf(1e6f));
        const float sc = s_ax < 0.f ? 0.f : (s_ax > 1.f ? 1.f : s_ax);
        const V3 c = P + s_ax * D + perp * (sc * delta + r * (1.f + e));
        const float4 a = make_float4(c.x, c.y, c.z, r), b = make_float4(0.f, 0.f, 0.f, 0.f);
        const V3 w = P - c;
        const float wl = RT_FAST_SQRT(dot(w, w));
        const bool dropped = r > 0.f && lc_hull_misses(c - P, D, delta, Dlen, r + 64.f * 5.96e-8f * (wl * wl * RT_FAST_RCP(r) + r));
        out[dropped ? 0 : 1]++;
        Rng rng; rng.init(seed, (uint32_t)k, 7u);
        for (int j = 0; j < n_rays; ++j) {
            const V3 lj = lp + random_unit_vector(rng) * (lrad / 2.f);
            const V3 Lj = normalized(lj - P);
            const float t_light = length(lj - P);
            const Ray ray = make_ray(P + Lj * RT_EPSF, Lj, 0.f);
            const float t = sphere_t(ray, make_sphere_ray(ray), a, b);
            const bool hit = t < t_light && t > RT_EPSF;
            if (hit && dropped) ++bad;
            if (hit && !dropped) out[2]++;
        }
    }
    return bad;
}

// The box padding of build_accel (slop = 128 eps kappa lmax) against the reference's triangle test, for conditioning numbers up to the limit
// RT_BVH_KAPPA_MAX: every point triangle_t accepts must lie within slop (+ the static pad's 1e-4 of the coordinates) of the TRUE triangle
// (distance in double). out[0] = triangles, out[1] = accepted hits checked, out[2] = largest accepted distance / slop seen, in 1e-6 units.
static double pt_seg_d2(const double *p, const double *a, const double *b) {
    double ab[3], ap[3], t = 0, l = 0;
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ap[k] = p[k] - a[k]; t += ab[k] * ap[k]; l += ab[k] * ab[k]; }
    t = l > 0 ? t / l : 0; t = t < 0 ? 0 : (t > 1 ? 1 : t);
    double d = 0;
    for (int k = 0; k < 3; ++k) { const double q = a[k] + t * ab[k] - p[k]; d += q * q; }
    return d;
}
int sim_check_slop(unsigned int seed, int n_tris, int n_rays, double kappa_lo, double kappa_hi, unsigned long long *out) {
    uint32_t st = seed * 3266489917u + 5u;
    int bad = 0;
    out[0] = out[1] = out[2] = 0;
    for (int k = 0; k < n_tris; ++k) {
        const float scale = expf(logf(0.01f) + urand(st) * logf(1e4f));
        const V3 p0 = vrand(st, -scale, scale);
        V3 dir = vrand(st, -1.f, 1.f);
        if (!(length(dir) > 1e-3f)) continue;
        dir = dir / length(dir);
        V3 perp = cross(dir, vrand(st, -1.f, 1.f));
        if (!(length(perp) > 1e-3f)) continue;
        perp = perp / length(perp);
        const float len = scale * expf(logf(1e-3f) + urand(st) * logf(1e3f));
        const float fa = 0.2f + 0.8f * urand(st), fb = 0.2f + 1.8f * urand(st);
        // kappa = d00 d11 / denom = 1 / sin^2(angle at p0): pick the angle for a kappa in [kappa_lo, kappa_hi]
        const double kap = exp(log(kappa_lo) + (double)urand(st) * log(kappa_hi / kappa_lo));
        const float hgt = (float)((double)(fb * len) * tan(asin(1.0 / sqrt(kap))));
        const V3 p1 = p0 + dir * (fa * len), p2 = p0 + dir * (fb * len) + perp * hgt;
        const TriConst tc = precompute_triangle(p0, p1, p2, 0u);
        // the builder's numbers, in double, from the float corners it sees (1.000001f * vertex)
        double c[3][3];
        const V3 pv[3] = {p0, p1, p2};
        for (int v = 0; v < 3; ++v) { c[v][0] = 1.000001f * pv[v].x; c[v][1] = 1.000001f * pv[v].y; c[v][2] = 1.000001f * pv[v].z; }
        double e0[3], e1[3], d00 = 0, d01 = 0, d11 = 0;
        for (int a = 0; a < 3; ++a) { e0[a] = c[1][a] - c[0][a]; e1[a] = c[2][a] - c[0][a]; d00 += e0[a] * e0[a]; d01 += e0[a] * e1[a]; d11 += e1[a] * e1[a]; }
        const double denom = d00 * d11 - d01 * d01;
        if (!(denom > 0.0)) continue;
        const double kappa = d00 * d11 / denom, lmax = sqrt(d00 > d11 ? d00 : d11);
        if (!(kappa <= kappa_hi * 1.5)) continue;
        const double slop = 128.0 * 5.96e-8 * kappa * lmax;
        float4 plane[1] = {tc.plane}, edge[3] = {tc.c0, tc.e0, tc.e1};
        float2 den[1] = {tc.den};
        DScene d;
        memset(&d, 0, sizeof d);
        d.tri_plane = plane; d.tri_edge = edge; d.tri_den = den;
        out[0]++;
        const V3 n = v3(tc.plane.x, tc.plane.y, tc.plane.z);
        if (!(n.x == n.x)) continue;
        const V3 c0 = v3(tc.c0.x, tc.c0.y, tc.c0.z);
        for (int j = 0; j < n_rays; ++j) {
            // a point of the triangle's neighbourhood: barycentric coordinates a little outside [0, 1], by up to a few slops
            const float over = (float)(4.0 * slop / lmax) * urand(st);
            const float u1 = -over + (1.f + 2.f * over) * urand(st), u2 = -over + (1.f + 2.f * over) * urand(st);
            const V3 x = c0 + v3(tc.e0.x, tc.e0.y, tc.e0.z) * u1 + v3(tc.e1.x, tc.e1.y, tc.e1.z) * u2;
            V3 dd = vrand(st, -1.f, 1.f);
            if (dot(dd, n) > 0.f) dd = v3(0.f) - dd;
            if (!(length(dd) > 1e-3f)) continue;
            const float t0 = scale * (0.01f + 10.f * urand(st));
            const Ray ray = make_ray(x - (dd / length(dd)) * t0, dd, 0.f);
            float w0, w1, w2;
            const float t = triangle_t<false>(ray, d, 0u, w0, w1, w2, nullptr);
            if (t == FLT_MAX) continue;
            out[1]++;
            const V3 pf = ray.o + t * ray.d;
            const double p[3] = {pf.x, pf.y, pf.z};
            // distance to the true triangle: inside its prism -> distance to the plane, else to the nearest edge
            double nn[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
            double q[3] = {p[0] - c[0][0], p[1] - c[0][1], p[2] - c[0][2]};
            const double d20 = q[0] * e0[0] + q[1] * e0[1] + q[2] * e0[2], d21 = q[0] * e1[0] + q[1] * e1[1] + q[2] * e1[2];
            const double b1 = (d11 * d20 - d01 * d21) / denom, b2 = (d00 * d21 - d01 * d20) / denom;
            double dist;
            if (b1 >= 0 && b2 >= 0 && b1 + b2 <= 1) {
                const double nl = sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
                dist = fabs(q[0] * nn[0] + q[1] * nn[1] + q[2] * nn[2]) / nl;
            } else {
                const double da = pt_seg_d2(p, c[0], c[1]), db = pt_seg_d2(p, c[1], c[2]), dc = pt_seg_d2(p, c[2], c[0]);
                dist = sqrt(da < db ? (da < dc ? da : dc) : (db < dc ? db : dc));
            }
            const double coords = fabs(p[0]) + fabs(p[1]) + fabs(p[2]) + (double)length(ray.o) + (double)t0;
            const double allow = slop + 1e-4 * coords + 1e-5;
            const unsigned long long ratio = (unsigned long long)(1e6 * dist / (slop + 1e-30));
            if (dist <= slop * 100 && ratio > out[2]) out[2] = ratio;
            if (dist > allow) ++bad;
        }
    }
    return bad;
}

// always_bound_of against the reference's triangle test (triangle_t on the constants precompute_triangle stores): for nearly and exactly
// collinear triangles, every point the test ACCEPTS must lie inside both slabs (|q . g| <= alpha + beta |q| + the rounding of q), and a
// triangle flagged "never" must accept nothing. Rays are aimed at points of the triangle's plane scattered around the sliver's line up
// to `reach` longest edges away. out[0] = triangles, out[1] = accepted hits checked, out[2] = triangles flagged never; returns violations.
int sim_check_always(unsigned int seed, int n_tris, int n_rays, float reach, unsigned long long *out) {
    uint32_t st = seed * 2246822519u + 99u;
    int bad = 0;
    out[0] = out[1] = out[2] = 0;
    for (int k = 0; k < n_tris; ++k) {
        const float scale = expf(logf(0.01f) + urand(st) * logf(1e4f));
        const V3 p0 = vrand(st, -scale, scale);
        V3 dir = vrand(st, -1.f, 1.f);
        if (!(length(dir) > 1e-3f)) continue;
        dir = dir / length(dir);
        V3 perp = cross(dir, vrand(st, -1.f, 1.f));
        if (!(length(perp) > 1e-3f)) continue;
        perp = perp / length(perp);
        const float len = scale * expf(logf(1e-3f) + urand(st) * logf(1e3f));
        const float fa = 0.2f + 0.8f * urand(st), fb = (urand(st) < 0.3f) ? 2.f * fa : 0.2f + 1.8f * urand(st);   // 30 %: equally spaced corners
        const float hgt = urand(st) < 0.25f ? 0.f : len * expf(logf(1e-9f) + urand(st) * logf(1e6f));            // 25 %: exactly on the line (before rounding)
        const V3 p1 = p0 + dir * (fa * len), p2 = p0 + dir * (fb * len) + perp * hgt;
        const TriConst tc = precompute_triangle(p0, p1, p2, 0u);
        float4 plane[1] = {tc.plane}, edge[3] = {tc.c0, tc.e0, tc.e1};
        float2 den[1] = {tc.den};
        DScene d;
        memset(&d, 0, sizeof d);
        d.tri_plane = plane; d.tri_edge = edge; d.tri_den = den;
        const AlwaysBound ab = always_bound_of(tc.c0, tc.e0, tc.e1, tc.den.x);
        out[0]++;
        if (ab.b.z != 0.f) out[2]++;
        const V3 n = v3(tc.plane.x, tc.plane.y, tc.plane.z);
        if (!(n.x == n.x)) continue;   // NaN normal: the test rejects everything (dotRN < 0 is never true)
        const V3 c0 = v3(tc.c0.x, tc.c0.y, tc.c0.z);
        const float lmax = sqrtf(fmaxf(tc.c0.w, tc.e1.w));
        // in-plane frame of the STORED triangle
        V3 ex = v3(tc.e1.x, tc.e1.y, tc.e1.z);
        if (!(length(ex) > 0.f)) continue;
        ex = ex / length(ex);
        V3 ey = cross(n, ex);
        if (!(length(ey) > 0.f)) continue;
        ey = ey / length(ey);
        for (int j = 0; j < n_rays; ++j) {
            const float along = (urand(st) * 2.f - 1.f) * reach * lmax;
            const float wid = fabsf(along) * expf(logf(1e-6f) + urand(st) * logf(1e6f)) + lmax * expf(logf(1e-9f) + urand(st) * logf(1e9f)) * (urand(st) < 0.5f ? 1.f : 0.f);
            const float side = (urand(st) < 0.5f ? -1.f : 1.f) * wid * (urand(st) < 0.3f ? 0.f : 1.f);
            const V3 x = c0 + ex * along + ey * side;
            V3 dd = vrand(st, -1.f, 1.f);
            if (dot(dd, n) > 0.f) dd = v3(0.f) - dd;
            if (!(length(dd) > 1e-3f)) continue;
            const float t0 = scale * (0.01f + 10.f * urand(st));
            const Ray ray = make_ray(x - (dd / length(dd)) * t0, dd, 0.f);
            float w0, w1, w2;
            const float t = triangle_t<false>(ray, d, 0u, w0, w1, w2, nullptr);
            if (t == FLT_MAX) continue;
            out[1]++;
            if (ab.b.z != 0.f) { ++bad; continue; }
            const V3 p = ray.o + t * ray.d;
            const V3 q = p - c0;
            const float ql = length(q);
            const float absm = 1e-4f * (ql + length(c0) + length(ray.o)) + 1e-6f;
            if (ab.g1.w < 1e30f && fabsf(dot(q, v3(ab.g1.x, ab.g1.y, ab.g1.z))) > ab.g1.w + ab.b.x * ql * 1.001f + absm) ++bad;
            if (ab.g2.w < 1e30f && fabsf(dot(q, v3(ab.g2.x, ab.g2.y, ab.g2.z))) > ab.g2.w + ab.b.y * ql * 1.001f + absm) ++bad;
        }
    }
    return bad;
}

static DCamera make_cam(const RtCamera *c) {
    DCamera d;
    for (int i = 0; i < 16; ++i) { d.mvi[i] = c->modelview_inverse[i]; d.pi[i] = c->projection_inverse[i]; }
    d.depth_near = c->depth_near;
    const double *m = c->modelview_inverse;
    double r[4];
    for (int k = 0; k < 4; ++k) r[k] = m[k] * 0.0 + m[4 + k] * 0.0 + m[8 + k] * 0.0 + m[12 + k] * 1.0;
    for (int k = 0; k < 3; ++k) d.pos[k] = (float)(r[k] / r[3]);
    return d;
}

// row-major rectangle, no tiling; linear, gamma: rect_h*rect_w*3; ids: 4 per pixel (sample 0)
void sim_render(void *h, const RtCamera *camera, const RtRenderParams *p, float *linear, float *gamma, uint32_t *ids, int threads) {
    const SimScene *s = (const SimScene *)h;
    const DCamera cam = make_cam(camera);
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
            for (int x = x0; x < x1; ++x) {
                const size_t o = (size_t)r * rw + (x - x0);
                V3 acc = v3(0.f);
                for (int sm = 0; sm < p->spp; ++sm) {
                    Rng rng;
                    rng.init(p->seed, (uint32_t)x + (uint32_t)y * (uint32_t)p->width, (uint32_t)sm);
                    const Ray ray = primary_ray(cam, x, y, p->width, p->height, rng);
                    if (sm == 0 && ids) {
                        float u, v;
                        const Hit hit = closest_hit<false>(s->d, ray, u, v, nullptr);
                        uint32_t *q = ids + 4 * o;
                        q[0] = hit.type; q[1] = hit.type ? hit.obj : 0; q[2] = hit.type == 3 ? f2u(s->d.tri_den[hit.ref].y) : 0; q[3] = f2u(hit.t);
                    }
                    if ((p->variant & 0xFF) == 6 && p->max_bounces > 0) {
                        // the wavefront kernels (k_wf_trace / k_wf_light), one lane: records in "global" memory, stride 1
                        const bool lc = s->d.abvh_root >= 0;
                        float4 rec[3 * RT_MAX_BOUNCES];
                        PathState st;
                        CandList cands;
                        st.cl = cands.v;
                        path_begin(st, ray, rng, 0u, p->max_bounces);
                        st.recs = nullptr; st.wf_rec = rec; st.wf_stride = 1;
                        V3 c = v3(0.f);
                        bool fin = false;
                        for (int level = 0; level < p->max_bounces && !fin; ++level) {
                            Hit hit; float hu = 0.f, hv = 0.f; bool blocked;
                            if (lc) intersect_lc<false, true>(s->d, st, true, true, hit, hu, hv, blocked, nullptr, true, level >= 1);   // as k_wf_trace: slab test, flat test from bounce 1 on
                            else intersect_ray<false, true>(s->d, st.ray, st.mode, st.t_light, st.rng, hit, hu, hv, blocked, nullptr);
                            if (path_shade<false, true>(s->d, st, hit, hu, hv, c, nullptr)) { fin = true; break; }
                            if (lc) fin = path_next_light_or_bounce<false, true, true>(s->d, st, p->nb_ech, c, nullptr);
                            else fin = path_next_light_or_bounce<false, false, true>(s->d, st, p->nb_ech, c, nullptr);
                            while (!fin && st.mode != 0) {
                                if (lc) {
                                    const bool was3 = st.mode == 3;
                                    // as the wavefront's kernels instantiate it: classify = the cone walk only, samples per queue (list / overflow with
                                    // the any-hit mesh walk)
                                    if (was3) intersect_lc<false, false, true>(s->d, st, true, true, hit, hu, hv, blocked, nullptr);
                                    else if (st.cl_n >= 0) intersect_lc<false, false, false, 2>(s->d, st, false, true, hit, hu, hv, blocked, nullptr);
                                    else intersect_lc<false, false, false, 3>(s->d, st, false, true, hit, hu, hv, blocked, nullptr);
                                    if (was3) {
                                        g_dbg_lights++;
                                        if (st.cl_n < 0) g_dbg_overflow++; else g_dbg_tris += st.cl_n;
                                        if (lc_light_unoccluded(st)) g_dbg_unoccluded++;
                                        else g_dbg_cands += __builtin_popcount(st.cm0) + __builtin_popcount(st.cm1) + __builtin_popcount(st.cm2) + __builtin_popcount(st.cm3);
                                    }
                                    fin = path_advance<false, true, true>(s->d, st, hit, hu, hv, blocked, p->nb_ech, c, nullptr);
                                } else {
                                    intersect_ray<false, true>(s->d, st.ray, st.mode, st.t_light, st.rng, hit, hu, hv, blocked, nullptr);
                                    fin = path_advance<false, false, true>(s->d, st, hit, hu, hv, blocked, p->nb_ech, c, nullptr);
                                }
                            }
                        }
                        acc = acc + c;
                    } else
                    if ((p->variant & 0xFF) >= 2) {   // the ray-level state machine of k_render_regen, one lane
                        const bool voted = (p->variant & 0xFF) == 4;
                        const bool lc = (p->variant & 0xFF) == 5 && s->d.abvh_root >= 0 && s->d.n_lights > 0;   // as rt_render_device selects
                        const bool accel = (p->variant & 0xFF) == 3 || ((p->variant & 0xFF) == 5 && !lc);
                        PathState st;
                        CandList cands;
                        st.cl = cands.v;
                        PathRecs recs;
                        path_begin(st, ray, rng, 0u, p->max_bounces);
                        st.recs = &recs;
                        V3 c = v3(0.f);
                        bool fin = false;
                        if (p->max_bounces == 0) { c = path_fold(st, v3(0.f)); fin = true; }
                        while (!fin) {
                            Hit hit; float hu = 0.f, hv = 0.f; bool blocked;
                            if (lc) {   // one lane: it always gets the step kind it wants
                                intersect_lc<false>(s->d, st, st.mode != 1, true, hit, hu, hv, blocked, nullptr);
                                fin = path_advance<false, true>(s->d, st, hit, hu, hv, blocked, p->nb_ech, c, nullptr);
                                continue;
                            }
                            if (voted) intersect_ray_voted<false>(s->d, st.ray, st.mode, st.t_light, st.rng, hit, hu, hv, blocked, nullptr);
                            else if (accel) intersect_ray<false, true>(s->d, st.ray, st.mode, st.t_light, st.rng, hit, hu, hv, blocked, nullptr);
                            else intersect_ray<false, false>(s->d, st.ray, st.mode, st.t_light, st.rng, hit, hu, hv, blocked, nullptr);
                            fin = path_advance<false>(s->d, st, hit, hu, hv, blocked, p->nb_ech, c, nullptr);
                        }
                        acc = acc + c;
                    } else
                    acc = acc + trace_path<false>(s->d, ray, rng, p->max_bounces, p->nb_ech, nullptr);
                }
                acc = acc / (float)(unsigned int)p->spp;
                if (linear) { linear[3 * o] = acc.x; linear[3 * o + 1] = acc.y; linear[3 * o + 2] = acc.z; }
                if (gamma) { gamma[3 * o] = gamma_channel(acc.x); gamma[3 * o + 1] = gamma_channel(acc.y); gamma[3 * o + 2] = gamma_channel(acc.z); }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto &t : pool) t.join();
}

void sim_stats(void *h, const RtCamera *camera, const RtRenderParams *p, unsigned long long *out10) {
    const SimScene *s = (const SimScene *)h;
    const DCamera cam = make_cam(camera);
    Counters c;
    memset(&c, 0, sizeof c);
    for (int y = 0; y < p->height; ++y)
        for (int x = 0; x < p->width; ++x)
            for (int sm = 0; sm < p->spp; ++sm) {
                Rng rng;
                rng.init(p->seed, (uint32_t)x + (uint32_t)y * (uint32_t)p->width, (uint32_t)sm);
                c.rnd += 3;
                const Ray ray = primary_ray(cam, x, y, p->width, p->height, rng);
                trace_path<true>(s->d, ray, rng, p->max_bounces, p->nb_ech, &c);
            }
    const unsigned long long v[10] = {c.closest, c.shadow, c.sphere, c.square, c.mesh, c.node, c.tri, c.tri_full, c.tex, c.rnd};
    memcpy(out10, v, sizeof v);
}

}  // extern "C"
