// rt_core.cuh — the render path as device functions (sm_100a).
//
// Everything the reference does per ray lives here, restated for the flattened scene:
//   closest hit      Scene::computeIntersection  Scene.h:202-230
//   occlusion        Scene::computeShadow        Scene.h:235-255
//   sphere           Sphere::intersect           Sphere.h:91-132
//   square           Square::intersect           Square.h:65-126
//   box slab         AABB::intersects            AABB.h:48-65
//   triangle         Triangle ctor/getIntersection/computeBarycentricCoordinates  Triangle.h:26-37,62-126
//   KD traversal     KDTree::intersect / Node::intersect  KDTree.cpp:31-85
//   shading          Scene::rayTraceRecursive    Scene.h:258-342
//   materials        Material::{scatter,emit,texture,sphere_texture,get_normal}  Material.cpp:13-130
//   helpers          reflect/refract/reflectance/random_unit_vector  Functions.cpp:14-54
//   sky              Scene::skyboxTexture        Scene.h:149-161
//
// PARITY RULES. The reference is plain IEEE fp32 evaluated left to right with NO
// fused multiply-add (g++ -O3, x86-64 baseline), promoting to fp64 at specific places. This file
// is compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, every expression keeps
// the reference's association, and the fp64 promotions are written out. Comparisons keep the
// reference's polarity so NaNs (zero-area triangles, rays parallel to a slab) fall the same way.
// Comparisons against the double constant EPSILON = 0.00001 are folded to float compares:
//   x >= EPSILON  <=>  x > 1e-5f ;  x < -EPSILON  <=>  x < -1e-5f ;  x <= EPSILON  <=>  x <= 1e-5f
// because 1e-5f = 9.99999975e-6 is the largest float below the double 1.0000000000000001e-5.
//
// The file also compiles as plain C++ (RT_HD expands to `inline`) so that tests can run these
// exact functions on the CPU for debugging; that build is test-only (tests/hostsim/) and is never
// linked into the product library.
#ifndef HAI719_RT_CORE_CUH
#define HAI719_RT_CORE_CUH

#include <cfloat>
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define RT_HD __device__ __forceinline__
#define RT_HHD __host__ __device__ __forceinline__
// Per-hit code that drags in software fp64 routines (acos, atan2, asin, pow, fmod, fp64 division) is kept
// OUT of line: inlined everywhere, k_render_regen was ~10 000 SASS instructions (160 KB) and its warps spent
// most of their stalled cycles waiting for instruction fetch (ncu: stall_no_instruction 6.5 per issue,
// profiles/r01_notes.md). The intersection loops stay inline.
#define RT_COLD __device__ __noinline__
// Small helpers that are used from a dozen places (normalisation = sqrt + three IEEE divisions, ~45 SASS
// instructions per copy) are SHARED, not inlined: after variant 5 the kernel waited on instruction fetch more
// than on anything else (ncu: stall_no_instruction 9.3 per issue, profiles/r01_notes.md), and one hot copy
// stays in the instruction cache where twenty cold ones do not. RT_OPT_NOINL: 0 = inline everything (old).
#ifndef RT_OPT_NOINL
#define RT_OPT_NOINL 1
#endif
#if RT_OPT_NOINL >= 1
#define RT_SHARED1 __device__ __noinline__
#else
#define RT_SHARED1 __device__ __forceinline__
#endif
#if RT_OPT_NOINL >= 2
#define RT_SHARED2 __device__ __noinline__
#else
#define RT_SHARED2 __device__ __forceinline__
#endif
#ifndef RT_OPT_NOINL_BOUNCE
#define RT_OPT_NOINL_BOUNCE 1
#endif
#if RT_OPT_NOINL_BOUNCE
#define RT_SHARED_BOUNCE __device__ __noinline__
#else
#define RT_SHARED_BOUNCE __device__ __forceinline__
#endif
#define RT_LDG(p) __ldg(p)
#define RT_LD_STREAM(p) __ldcs(p)          /* written once, read once, far larger than L2: evict-first */
#define RT_ST_STREAM(p, v) __stcs((p), (v))
#else
#define RT_LD_STREAM(p) (*(p))
#define RT_ST_STREAM(p, v) (*(p) = (v))
#define RT_SHARED_BOUNCE inline
#define RT_HD inline
#define RT_HHD inline
#define RT_COLD inline
#define RT_SHARED1 inline
#define RT_SHARED2 inline
#define RT_LDG(p) (*(p))
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#endif

// warp-level helpers (the plain-C++ build of this header is a single lane)
#ifdef __CUDA_ARCH__
#define RT_WARP_ALL(pred) __all_sync(0xFFFFFFFFu, (pred))
#define RT_BALLOT(pred) __ballot_sync(0xFFFFFFFFu, (pred))
#define RT_POPC(x) __popc(x)
#define RT_FFS(x) __ffs(x)
#define RT_FAST_RCP(x) __fdividef(1.f, (x))   /* 2-ulp reciprocal: only used where a 1e-5 relative margin follows */
#define RT_FMA(a, b, c) __fmaf_rn((a), (b), (c))   /* explicit: the build runs with -fmad=false; only in the conservative filters */
#define RT_FAST_SQRT(x) ((x) * __frsqrt_rn(fmaxf((x), 1e-37f)))   /* ~2-ulp square root (0 for x = 0): only where a 1e-4 relative margin follows */
#else   /* one lane */
#define RT_FAST_SQRT(x) sqrtf(x)
#define RT_FFS(x) __builtin_ffs(x)
#define RT_FAST_RCP(x) (1.f / (x))
#define RT_FMA(a, b, c) fmaf((a), (b), (c))
#define RT_WARP_ALL(pred) (pred)
#define RT_BALLOT(pred) ((pred) ? 1u : 0u)
#define RT_POPC(x) ((int)((x) != 0u))
#endif

// A/B switches of control-flow variants (all give identical results; defaults = what measured fastest on B200,
// profiles/r01_notes.md). Alternate builds: make -C hai719-raytracing_b200 alt NAME=x DEFS="-DRT_OPT_WW=0".
#ifndef RT_OPT_MASKLOOP
#define RT_OPT_MASKLOOP 0   /* shadow candidates: one loop taking the k-th candidate of every lane together (15 % slower) */
#endif
#ifndef RT_OPT_REACH2
#define RT_OPT_REACH2 1     /* leaf reachability: cheap test on every leaf of the triangle first, exact chains only if none passes */
#endif
#ifndef RT_OPT_WW
#define RT_OPT_WW 1         /* mesh culling hierarchy: while-while traversal */
#endif
#ifndef RT_OPT_WW_LC
#define RT_OPT_WW_LC 1      /* analytic culling hierarchy (variants 5, 6): while-while traversal */
#endif
#ifndef RT_OPT_MESH_MERGED
#define RT_OPT_MESH_MERGED 1 /* all meshes in one per-lane while-while loop instead of one loop per mesh: config 4 -10 %, config 5 -3 %, config 3 +1 % */
#endif

namespace rt {

#define RT_EPSF 1e-5f          /* (float)EPSILON */
#define RT_MAX_BOUNCES 16      /* capacity of the per-path radiance records */
#ifndef RT_LC_MAXC
#define RT_LC_MAXC 16          /* capacity of the per-light list of candidate triangles (variants 5, 6) */
#endif
#define RT_PI 3.14159265358979323846 /* M_PI */

// ---- vectors -----------------------------------------------------------------------------------
struct V3 { float x, y, z; };
RT_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD V3 v3(float f) { return v3(f, f, f); }
RT_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
RT_HD V3 operator*(V3 a, float s) { return v3(s * a.x, s * a.y, s * a.z); }
RT_HD V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
RT_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // (x+y)+z, Vec3.h:36
RT_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RT_HD V3 comp_product(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD float length(V3 a) { return sqrtf(dot(a, a)); }                        // Vec3.h:32
RT_SHARED1 V3 normalized(V3 a) { const float L = length(a); return v3(a.x / L, a.y / L, a.z / L); }  // Vec3.h:35
// the same arithmetic, always inlined: for the few kernels small enough that the call costs more than the copies (the sample kernel
// of a scene without meshes: three normalisations per shadow sample; config 2 30.4 -> 29.7 ms with every copy inlined, but config 5
// 53.5 -> 54.0, so only that kernel asks for it; profiles/r02_notes.md, r03v)
RT_HD V3 normalized_inl(V3 a) { const float L = length(a); return v3(a.x / L, a.y / L, a.z / L); }
RT_HD V3 ld3(const float *p) { return v3(p[0], p[1], p[2]); }
RT_HD float fminr(float a, float b) { return a < b ? a : b; }                // ::min, Functions.cpp:20
RT_HD float fmaxr(float a, float b) { return a > b ? a : b; }                // ::max, Functions.cpp:24
RT_HHD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
RT_HHD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// ---- random stream (definition in include/hai719_rt.h) ----------------------------------------
RT_HD uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
struct Rng {
    uint32_t key, ctr;
    RT_HD void init(uint32_t seed, uint32_t pixel, uint32_t sample) {
        key = fmix32(fmix32(seed ^ ((pixel + 1u) * 0x9E3779B9u)) + (sample + 1u) * 0x85EBCA6Bu);
        ctr = 0;
    }
    RT_HD float next() {   // random_float(), Functions.cpp:4-8
        const uint32_t r = fmix32(key + ctr * 0x9E3779B9u);
        ++ctr;
        return (float)(r >> 8) * (1.0f / 16777216.0f);
    }
};
// random_unit_vector (Functions.cpp:14-18): a point of the cube [-1,1]^3, normalised. The three
// draws are constructor arguments, evaluated right to left by g++: z, then y, then x.
RT_SHARED2 V3 random_unit_vector(Rng &rng) {
    const float z = -1.f + 2.f * rng.next();
    const float y = -1.f + 2.f * rng.next();
    const float x = -1.f + 2.f * rng.next();
    return normalized(v3(x, y, z));
}
RT_HD V3 random_unit_vector_inl(Rng &rng) {
    const float z = -1.f + 2.f * rng.next();
    const float y = -1.f + 2.f * rng.next();
    const float x = -1.f + 2.f * rng.next();
    return normalized_inl(v3(x, y, z));
}

// ---- device scene ------------------------------------------------------------------------------
struct DMaterial {          // mirrors RtMaterial (include/hai719_rt.h)
    int type, texture_type;
    float kd[3];
    float transparency, index_medium;
    float checker1[3], checker2[3];
    float tsx, tsy;
    int emissive;
    float light_color[3];
    float light_intensity;
    int image, normal_map;
    float motion[3];
};
struct DImage { int w, h; const unsigned char *rgb; };
struct DSquare {            // per-square constants of Square::intersect hoisted out of the ray loop
    float v0[3];            // vertices[0].position
    float n[3];             // normalize(cross(v1-v0, v3-v0))            Square.h:69-72
    float right[3], up[3];  // v1-v0, v3-v0
    float len_r, len_u;     // |right|, |up|                             Square.h:106,110
    float motion[3];
    int glass;              // material.type == Material_Glass           Square.h:84
    float tan_r[3], tan_u[3];  // m_right_vector / m_up_vector members (normal-map frame)
};
// All meshes share one node array and one set of triangle arrays (rt_pack.hpp). Mesh m owns the node
// range [node_begin, node_end): first a synthetic root whose box is KDTree::aabb — the gate
// KDTree::intersect tests before descending (KDTree.cpp:80-83) — then the tree in pre-order, skip
// links and leaf ranges already offset into the shared arrays. An empty tree owns an empty range.
struct DMesh {
    uint32_t node_begin, node_end;
    int bvh_root;              // variant 3 (rt_bvh.hpp): root of this mesh's culling BVH, -1 = none
    int bvh4_root;             // the same hierarchy with 4-wide nodes (DScene::bvh4_nodes), -1 = none: walk the binary one
    uint32_t always_first, always_count;   // range of bvh_tris tested for every ray (ill-conditioned triangles)
    int color_type;
    const float *vert_colors, *face_colors;
    const uint32_t *triangles;
};
struct DLight { float pos[3]; float radius; float color[3]; };
struct DScene {
    int n_spheres, n_squares, n_meshes, n_lights;
    const float4 *sph_a;       // {c.xyz, r}
    const float4 *sph_b;       // {motion.xyz, transparency}
    const DSquare *squares;
    const float *sq_transparency;
    const DMesh *meshes;
    const float *mesh_transparency;
    // node i: lo = {bmin.xyz, bits(inner: skip | leaf: first_ref)}, hi = {bmax.xyz, bits(leaf: 0x80000000|n_refs, inner: 0)}
    const float4 *node_lo, *node_hi;
    const float4 *tri_plane;   // per leaf ref: {n.xyz, D = c0.n}
    const float4 *tri_edge;    // per leaf ref, 3 entries: {c0.xyz, d00}, {e0.xyz, d01}, {e1.xyz, d11}
    const float2 *tri_den;     // per leaf ref: {denom, bits(tri_index within its mesh)}
    // exact culling structure of variant 3 (rt_bvh.hpp)
    const float4 *bvh_nodes;   // 4 per node
    const float4 *bvh4_nodes;  // 8 per node: lo.x lo.y lo.z hi.x hi.y hi.z of the four children, their codes, spare (rt_bvh.hpp : collapse4)
    const uint32_t *bvh_tris;  // representative leaf ref per distinct triangle, BVH leaf order
    const float4 *always_bound; // 3 per entry of bvh_tris, filled for the always-tested ranges only (always_bound_of)
    const uint32_t *ref_next, *ref_leaf, *node_parent;
    // culling hierarchy over spheres + squares (rt_bvh.hpp : build_analytic_accel); abvh_root < 0 = linear loops
    const float4 *abvh_nodes;
    const uint32_t *abvh_prims;
    int abvh_root;
    int abvh_n_nodes;          // inner nodes of that hierarchy (<= 127: it is only built for <= 128 primitives)
    float abvh_c[3], abvh_r;
    float abvh_cs[3], abvh_rs;   // ball of the sphere centres (quadratic term of the per-ray padding), see AnalyticAccel
    const float4 *abvh_flat;     // <= 32 analytic primitives: their padded boxes in sequence order, {lo.xyz, coef} {hi.xyz, -} each; else null
    const DMaterial *sph_mat, *sq_mat, *mesh_mat;
    const DLight *lights;
    const DImage *textures, *normal_maps;
    DImage sky;
    int dark_sky;
};

// The scene of the render in flight lives in the constant bank (one upload per render call, rt_capi.cu : SceneBinding):
// kernels and — what matters — the OUT-OF-LINE helpers read its fields as c[bank][offset] operands. Passed as a kernel
// parameter and handed on by reference, the 264-byte struct was copied to every thread's local-memory frame at kernel
// entry (19 STL.128) and read back through LDL by each helper. RT_SCENE_CONST=0 restores the by-value parameter (A/B).
#ifndef RT_SCENE_CONST
#define RT_SCENE_CONST 1
#endif
#if defined(__CUDACC__) && RT_SCENE_CONST
__constant__ DScene c_scene;
#define RT_S(s_) c_scene
#else
#define RT_S(s_) (s_)
#endif

// The analytic culling hierarchy staged in shared memory. It is built for at most 128 primitives, i.e. at most 127 nodes
// of 64 B: the WHOLE hierarchy fits in 8 KB. Each CTA of the kernels that walk it pulls it in once, at kernel start, with
// one bulk asynchronous copy (cp.async.bulk global -> shared, completion on an mbarrier: SASS UBLKCP + SYNCS), and the
// walks then read nodes with LDS instead of LDG (north_star: "top levels staged into shared memory via TMA"; here the top
// levels are all levels). RT_STAGE_ABVH=0: nodes through L1 (A/B, profiles/r02_notes.md).
#ifndef RT_STAGE_ABVH
#define RT_STAGE_ABVH 0
#endif
#define RT_ABVH_MAX_NODES 127
#if defined(__CUDACC__) && RT_STAGE_ABVH
__device__ __forceinline__ float4 *abvh_sm() { __shared__ alignas(128) float4 a[4 * RT_ABVH_MAX_NODES]; return a; }
__device__ __forceinline__ void stage_abvh(const DScene &s) {
    __shared__ alignas(8) unsigned long long mbar;
    if (s.abvh_root < 0 || s.abvh_n_nodes <= 0 || s.abvh_n_nodes > RT_ABVH_MAX_NODES) return;   // kernel-uniform
    const unsigned int bar = (unsigned int)__cvta_generic_to_shared(&mbar);
    const unsigned int dst = (unsigned int)__cvta_generic_to_shared(abvh_sm());
    const unsigned int bytes = (unsigned int)s.abvh_n_nodes * 64u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(s.abvh_nodes), "r"(bytes), "r"(bar) : "memory");
    }
    unsigned int done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar) : "memory");
    }
}
#define RT_ABVH_NODES(s) ((const float4 *)abvh_sm())   /* rt_scene_create builds no hierarchy beyond RT_ABVH_MAX_NODES nodes */
#define RT_ABVH_LD(p) (*(p))
#else
#if defined(__CUDACC__)
__device__ __forceinline__ void stage_abvh(const DScene &) {}
#endif
#define RT_ABVH_NODES(s) ((s).abvh_nodes)
#define RT_ABVH_LD(p) RT_LDG(p)
#endif

struct Counters {   // per-thread work counters (only touched when STATS)
    unsigned long long closest, shadow, sphere, square, mesh, node, tri, tri_full, tex, rnd;
};

struct Ray {
    V3 o, d;
    float time;
};
// Ray(o, d, time): the Line constructor normalises d (Line.h:13-16)
RT_SHARED2 Ray make_ray(V3 o, V3 d, float time) { Ray r; r.o = o; r.d = normalized(d); r.time = time; return r; }
RT_HD Ray make_ray_inl(V3 o, V3 d, float time) { Ray r; r.o = o; r.d = normalized_inl(d); r.time = time; return r; }

struct Hit {
    int type;       // 0 miss, 1 sphere, 2 square, 3 mesh
    int obj;
    float t;
    uint32_t ref;   // mesh: leaf-ref index of the winning triangle test
};

// ---- sphere ------------------------------------------------------------------------------------
// Sphere::intersect up to t (Sphere.h:94-123):  t = (-b - sqrt(delta)) / (2a), near root only (the
// far-root branch at Sphere.h:112-117 can never be taken). Every caller keeps a hit
// only if EPSILON <= t < something, so the two cheap rejections below are exact:
//   * b >= 0 (or -0)  =>  -b - sqrt(delta) <= 0  =>  t <= 0 < EPSILON           (a = d.d > 0)
//   * delta < 0       =>  intersectionExists == false
// NaNs fail both "proceed" tests and are dropped, as they are by the reference's comparisons.
// Returns FLT_MAX for "no usable hit". A4 = 4*a and A2 = 2*a are per-ray constants (same products
// the reference forms: 4*a*c associates as (4*a)*c).
struct SphereRay { float A4, A2; };
RT_HD SphereRay make_sphere_ray(const Ray &ray) {
    const float A = dot(ray.d, ray.d);
    SphereRay r; r.A4 = 4.f * A; r.A2 = 2.f * A; return r;
}
RT_HD float sphere_t(const Ray &ray, const SphereRay &sr, float4 a, float4 b) {
    const V3 c = v3(a.x, a.y, a.z) + ray.time * v3(b.x, b.y, b.z);
    const V3 oc = ray.o - c;
    const float B = 2.f * dot(ray.d, oc);
    if (!(B < 0.f)) return FLT_MAX;
    const float C = dot(oc, oc) - a.w * a.w;
    const float delta = B * B - sr.A4 * C;
    if (!(delta >= 0.f)) return FLT_MAX;
    return (-B - sqrtf(delta)) / sr.A2;
}

// ---- square ------------------------------------------------------------------------------------
// Square::intersect (Square.h:65-126). Returns t and (u, v), or FLT_MAX when there is no hit.
RT_HD float square_t(const Ray &ray, const DSquare &q, float &u, float &v) {
    const V3 n = ld3(q.n);
    const float dotRN = dot(ray.d, n);
    // parallel (== 0) or back face of a non-glass quad (> 0) miss; a NaN goes on and fails below
    if (dotRN == 0.f) return FLT_MAX;
    if (dotRN > 0.f && !q.glass) return FLT_MAX;
    const V3 bl = ld3(q.v0) + ray.time * ld3(q.motion);
    const float D = dot(bl, n);
    const float t = (D - dot(ray.o, n)) / dotRN;
    if (t < -RT_EPSF) return FLT_MAX;
    if (t > RT_EPSF) {
        const V3 p = ray.o + t * ray.d;
        const V3 w = p - bl;
        const float proj1 = dot(w, ld3(q.right)) / q.len_r;
        const float proj2 = dot(w, ld3(q.up)) / q.len_u;
        if ((proj1 <= q.len_r && proj1 >= 0.f) && (proj2 <= q.len_u && proj2 >= 0.f)) {
            u = proj1 / q.len_r;
            v = proj2 / q.len_u;
            return t;
        }
    }
    return FLT_MAX;
}

// ---- box ---------------------------------------------------------------------------------------
// AABB::intersects (AABB.h:48-65): tmin = (float)EPSILON, tmax = FLT_MAX; 1/d and the two products
// in fp64, rounded to float when stored in t0/t1. inv = 1.0 / (double)d per axis, hoisted per ray.
struct RayInv { double x, y, z; };
RT_HD RayInv make_inv(const Ray &r) { RayInv i; i.x = 1.0 / (double)r.d.x; i.y = 1.0 / (double)r.d.y; i.z = 1.0 / (double)r.d.z; return i; }

#define RT_SLAB_AXIS(LO, HI, O, INV)                                  \
    {                                                                 \
        const float t0 = (float)((double)((LO) - (O)) * (INV));       \
        const float t1 = (float)((double)((HI) - (O)) * (INV));       \
        if (t0 < t1) {                                                \
            if (t0 > tmin) tmin = t0;                                 \
            if (t1 < tmax) tmax = t1;                                 \
        } else {                                                      \
            if (t1 > tmin) tmin = t1;                                 \
            if (t0 < tmax) tmax = t0;                                 \
        }                                                             \
        if (tmax <= tmin) return false;                               \
    }
RT_HD bool slab_hit(const Ray &r, const RayInv &inv, float lx, float ly, float lz, float hx, float hy, float hz) {
    float tmin = RT_EPSF, tmax = FLT_MAX;
    RT_SLAB_AXIS(lx, hx, r.o.x, inv.x)
    RT_SLAB_AXIS(ly, hy, r.o.y, inv.y)
    RT_SLAB_AXIS(lz, hz, r.o.z, inv.z)
    return true;
}

// ---- triangle ----------------------------------------------------------------------------------
// Per-reference constants (computed once by precompute_triangle, same arithmetic as the Triangle
// constructor that the reference runs per ray per triangle, KDTree.cpp:38-40):
//   c_i = 1.000001f * vertex_i ; e0 = c1-c0 ; e1 = c2-c0 ; n = cross(e0,e1)/|cross| ; D = c0.n
//   d00 = e0.e0 ; d01 = e0.e1 ; d11 = e1.e1 ; denom = d00*d11 - d01*d01
struct TriConst { float4 plane, c0, e0, e1; float2 den; };
RT_HD TriConst precompute_triangle(V3 p0, V3 p1, V3 p2, uint32_t tri_index) {
    const float s = 1.000001f;
    const V3 c0 = p0 * s, c1 = p1 * s, c2 = p2 * s;
    const V3 e0 = c1 - c0, e1 = c2 - c0;
    const V3 nn = cross(e0, e1);
    const float norm = length(nn);
    const V3 n = nn / norm;                       // 0/0 = NaN for zero-area triangles, kept (A.1-18)
    TriConst k;
    k.plane = make_float4(n.x, n.y, n.z, dot(c0, n));
    const float d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1);
    k.c0 = make_float4(c0.x, c0.y, c0.z, d00);
    k.e0 = make_float4(e0.x, e0.y, e0.z, d01);
    k.e1 = make_float4(e1.x, e1.y, e1.z, d11);
    k.den = make_float2(d00 * d11 - d01 * d01, u2f(tri_index));
    return k;
}

// Triangle::getIntersection (Triangle.h:77-126). Returns t, or FLT_MAX for "no intersection".
template <bool STATS>
RT_HD float triangle_t(const Ray &ray, const DScene &m, uint32_t ref, float &w0, float &w1, float &w2, Counters *cnt) {
    const float4 pl = RT_LDG(m.tri_plane + ref);
    const V3 n = v3(pl.x, pl.y, pl.z);
    const float dotRN = dot(ray.d, n);
    if (!(dotRN < 0.f)) return FLT_MAX;           // == 0, > 0; a NaN normal can never pass the final test
    const float t = (pl.w - dot(ray.o, n)) / dotRN;
    if (!(t >= 0.f)) return FLT_MAX;              // t < 0; NaN likewise ends in "no intersection"
    if (STATS) cnt->tri_full++;
    const V3 p = ray.o + t * ray.d;
    const float4 a = RT_LDG(m.tri_edge + 3 * ref), b = RT_LDG(m.tri_edge + 3 * ref + 1), c = RT_LDG(m.tri_edge + 3 * ref + 2);
    const V3 q = p - v3(a.x, a.y, a.z);
    const float d20 = dot(q, v3(b.x, b.y, b.z));
    const float d21 = dot(q, v3(c.x, c.y, c.z));
    const float denom = RT_LDG(m.tri_den + ref).x;
    const float u1 = (c.w * d20 - b.w * d21) / denom;
    const float u2 = (a.w * d21 - b.w * d20) / denom;
    const float u0 = 1.f - u1 - u2;
    if (u0 >= 0.f && u0 <= 1.f && u1 >= 0.f && u1 <= 1.f && u2 >= 0.f && u2 <= 1.f) {
        w0 = u0; w1 = u1; w2 = u2;
        return t;
    }
    return FLT_MAX;
}

// Where can the reference's barycentric test accept a point, for a triangle too ill-conditioned to be boxed (the "always-tested" list of
// rt_bvh.hpp: nearly or exactly collinear corners)? It accepts q = p - c0 only if 0 <= u1, u2 <= 1 with u1 = N1 / den,
// N1 = fl(fl(d11 * d20) - fl(d01 * d21)), d20 = fl(q . e0), d21 = fl(q . e1) (triangle_t), hence only if |N1| <= |den| (1 + eps).
// On the STORED constants, d11 (q . e0) - d01 (q . e1) = q . g1 exactly, g1 = d11 e0 - d01 e1, and N1 differs from that by at most
// (4 eps + O(eps^2)) |q| (|d11| |e0| + |d01| |e1|) (three roundings per dot product, one per product, one for the difference). So
//     |q . g1 / |g1||  <=  alpha1 + beta1 |q|,   alpha1 = 1.001 |den| / |g1|,   beta1 = 8 eps (|d11| |e0| + |d01| |e1|) / |g1|
// (twice the first-order constant), and the same with g2 = d00 e1 - d01 e0 for u2: the accepted points lie in a slab around the plane
// through c0 perpendicular to g1 whose half-width grows with the distance from c0 — for collinear corners a thin double wedge around
// the plane that holds the sliver's line. A light cone whose hull lies on one side of that slab cannot be occluded by the triangle
// (intersect_lc). beta >= 0.25, g = 0 or a non-finite constant: no statement (alpha = FLT_MAX). Evaluated in double from the float
// constants, once per always-tested triangle (k_always_bounds / the CPU simulation).
struct AlwaysBound { float4 g1, g2, b; };   // {g1 / |g1|, alpha1} {g2 / |g2|, alpha2} {beta1, beta2, -, -}
RT_HD AlwaysBound always_bound_of(float4 ta, float4 tb, float4 tc, float den) {
    AlwaysBound o;
    o.g1 = make_float4(0.f, 0.f, 0.f, FLT_MAX); o.g2 = make_float4(0.f, 0.f, 0.f, FLT_MAX); o.b = make_float4(0.f, 0.f, 0.f, 0.f);
    const double e0[3] = {tb.x, tb.y, tb.z}, e1[3] = {tc.x, tc.y, tc.z};
    const double d00 = ta.w, d01 = tb.w, d11 = tc.w, dn = fabs((double)den);
    const double le0 = sqrt(e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2]), le1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    double g1[3], g2[3], n1 = 0.0, n2 = 0.0;
    for (int a = 0; a < 3; ++a) { g1[a] = d11 * e0[a] - d01 * e1[a]; g2[a] = d00 * e1[a] - d01 * e0[a]; n1 += g1[a] * g1[a]; n2 += g2[a] * g2[a]; }
    n1 = sqrt(n1); n2 = sqrt(n2);
    const double eps = 5.96e-8;
    if (!(dn > 0.0) || !(dn < 3e38)) { o.b.z = 1.f; return o; }   // den = 0, inf or NaN: N / den is never in [0, 1] — the triangle reports no hit at all
    if (n1 > 0.0 && n1 < 1e300 && dn < 1e300) {
        const double alpha = 1.001 * dn / n1, beta = 8.0 * eps * (fabs(d11) * le0 + fabs(d01) * le1) / n1;
        if (beta < 0.25 && alpha < 1e30) { o.g1 = make_float4((float)(g1[0] / n1), (float)(g1[1] / n1), (float)(g1[2] / n1), (float)alpha * 1.000001f + 1e-37f); o.b.x = (float)beta * 1.000001f; }
    }
    if (n2 > 0.0 && n2 < 1e300 && dn < 1e300) {
        const double alpha = 1.001 * dn / n2, beta = 8.0 * eps * (fabs(d00) * le1 + fabs(d01) * le0) / n2;
        if (beta < 0.25 && alpha < 1e30) { o.g2 = make_float4((float)(g2[0] / n2), (float)(g2[1] / n2), (float)(g2[2] / n2), (float)alpha * 1.000001f + 1e-37f); o.b.y = (float)beta * 1.000001f; }
    }
    return o;
}

// ---- mesh: stackless pre-order KD traversal ----------------------------------------------------
// KDTree::intersect + Node::intersect. The reference recurses into BOTH children of every node
// whose box the ray touches, without ordering or early exit, and combines with
// "left.t < right.t ? left : right"  =>  over the leaves in depth-first order the result is the
// minimum t, the LAST leaf winning ties; inside a leaf the FIRST triangle wins ties (strict <).
// The pre-order array reproduces that order with no stack: a missed box jumps to `skip`.
template <bool STATS>
RT_HD bool mesh_closest(const Ray &ray, const RayInv &inv, const DScene &s, const DMesh &m, float &t_out, uint32_t &ref_out, Counters *cnt) {
    float best_t = FLT_MAX;
    uint32_t best_ref = 0xFFFFFFFFu;
    uint32_t i = m.node_begin;
    const uint32_t end = m.node_end;
    while (i < end) {
        const float4 lo = RT_LDG(s.node_lo + i), hi = RT_LDG(s.node_hi + i);
        const uint32_t hw = f2u(hi.w);
        const bool leaf = (hw & 0x80000000u) != 0u;
        if (STATS) cnt->node++;
        if (!slab_hit(ray, inv, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) {
            i = leaf ? i + 1 : f2u(lo.w);
            continue;
        }
        if (leaf) {
            const uint32_t first = f2u(lo.w), count = hw & 0x7FFFFFFFu;
            float leaf_t = FLT_MAX;
            uint32_t leaf_ref = 0xFFFFFFFFu;
            for (uint32_t k = first; k < first + count; ++k) {
                float a, b, c;
                if (STATS) cnt->tri++;
                const float t = triangle_t<STATS>(ray, s, k, a, b, c, cnt);
                if (t < leaf_t) { leaf_t = t; leaf_ref = k; }
            }
            if (leaf_ref != 0xFFFFFFFFu && leaf_t <= best_t) { best_t = leaf_t; best_ref = leaf_ref; }
        }
        ++i;
    }
    if (best_ref == 0xFFFFFFFFu) return false;
    t_out = best_t;
    ref_out = best_ref;
    return true;
}

// NOTE on the out-of-line (RT_COLD) functions below: they take the ray BY VALUE. By reference, the address of the caller's
// ray — a member of the kernel's path state — escaped into a call, and the compiler then kept the WHOLE path state in the
// thread's local-memory frame instead of in registers: every field assignment became an STL, and local stores write through
// to L2 (ncu, config 2, level-0 trace kernel: 226 M sectors = 7.2 GB of local stores against 1.7 GB of record stores, 5.3 TB/s
// of L2 write traffic in a 1.35 ms kernel; profiles/r02_notes.md).
// ---- mesh: exact culling traversal (variant 3, construction and proof sketch in rt_bvh.hpp) ------
// Is the leaf that holds reference r reachable by the reference's traversal for this ray, given that
// the triangle was hit at parameter t?  Fast accept: t lies strictly inside the slab interval of
// the leaf's box on every axis, with a margin (1e-5 relative to the slab parameters + 1e-6) far above
// the rounding of either arithmetic, and t > 1e-3 >> EPSILON: then for the leaf and for every
// ancestor (whose boxes contain it) AABB::intersects keeps tmin < t < tmax at every step and
// returns true. Anything else — grazing, axis-parallel, tiny t — re-runs the reference's own fp64
// slab test (slab_hit) up the parent chain, synthetic root (= KDTree::aabb gate) included.
RT_COLD bool leaf_reachable_exact(const DScene &s_, const Ray ray, uint32_t leaf) {
    const DScene &s = RT_S(s_);
    const RayInv inv = make_inv(ray);
    uint32_t n = leaf;
    while (n != 0xFFFFFFFFu) {
        const float4 l = RT_LDG(s.node_lo + n), h = RT_LDG(s.node_hi + n);
        if (!slab_hit(ray, inv, l.x, l.y, l.z, h.x, h.y, h.z)) return false;
        n = RT_LDG(s.node_parent + n);
    }
    return true;
}
// the cheap sufficient condition alone (see above); i* = approximate reciprocals of the ray direction
RT_HD bool leaf_contains_hit(const DScene &s, const Ray &ray, float ix, float iy, float iz, float t, uint32_t leaf) {
    const float4 lo = RT_LDG(s.node_lo + leaf), hi = RT_LDG(s.node_hi + leaf);
    const float a0 = (lo.x - ray.o.x) * ix, a1 = (hi.x - ray.o.x) * ix;
    const float b0 = (lo.y - ray.o.y) * iy, b1 = (hi.y - ray.o.y) * iy;
    const float c0 = (lo.z - ray.o.z) * iz, c1 = (hi.z - ray.o.z) * iz;
    const float ma = 1e-5f * (fabsf(a0) + fabsf(a1)) + 1e-6f, mb = 1e-5f * (fabsf(b0) + fabsf(b1)) + 1e-6f,
                mc = 1e-5f * (fabsf(c0) + fabsf(c1)) + 1e-6f;
    bool ok = t > 1e-3f;
    ok = ok && (fminr(a0, a1) + ma < t) && (t < fmaxr(a0, a1) - ma);
    ok = ok && (fminr(b0, b1) + mb < t) && (t < fmaxr(b0, b1) - mb);
    ok = ok && (fminr(c0, c1) + mc < t) && (t < fmaxr(c0, c1) - mc);
    return ok;
}
RT_HD bool leaf_reachable(const DScene &s, const Ray &ray, float t, uint32_t leaf) {
    if (leaf_contains_hit(s, ray, RT_FAST_RCP(ray.d.x), RT_FAST_RCP(ray.d.y), RT_FAST_RCP(ray.d.z), t, leaf)) return true;
    return leaf_reachable_exact(s, ray, leaf);
}
RT_COLD bool tri_reachable_exact(const DScene &s_, const Ray ray, uint32_t r0) {
    const DScene &s = RT_S(s_);
    for (uint32_t r = r0; r != 0xFFFFFFFFu; r = RT_LDG(s.ref_next + r))
        if (leaf_reachable_exact(s, ray, RT_LDG(s.ref_leaf + r))) return true;
    return false;
}
// Is ANY leaf holding the triangle (references chained from r0) reachable? Two passes: the cheap
// sufficient test on every such leaf first - the hit point lies inside one or two of the 3-5 cells
// that reference a triangle, and walking the parent chain in fp64 for each cell tried before that
// one was 20 % of the pond scene's instructions at 1.7 active lanes (profiles/r01_notes.md) - and
// only if no cell contains the hit with margin, the reference's own arithmetic up the chains.
RT_HD bool tri_reachable(const DScene &s, const Ray &ray, float t, uint32_t r0) {
#if RT_OPT_REACH2
    const float ix = RT_FAST_RCP(ray.d.x), iy = RT_FAST_RCP(ray.d.y), iz = RT_FAST_RCP(ray.d.z);
    for (uint32_t r = r0; r != 0xFFFFFFFFu; r = RT_LDG(s.ref_next + r))
        if (leaf_contains_hit(s, ray, ix, iy, iz, t, RT_LDG(s.ref_leaf + r))) return true;
    return tri_reachable_exact(s, ray, r0);
#else
    for (uint32_t r = r0; r != 0xFFFFFFFFu; r = RT_LDG(s.ref_next + r))
        if (leaf_reachable(s, ray, t, RT_LDG(s.ref_leaf + r))) return true;
    return false;
#endif
}
// last reachable reference (decides ties between different triangles at the same t)
RT_COLD uint32_t tri_last_reachable(const DScene &s_, const Ray ray, float t, uint32_t r0) {
    const DScene &s = RT_S(s_);
    uint32_t last = 0xFFFFFFFFu;
    for (uint32_t r = r0; r != 0xFFFFFFFFu; r = RT_LDG(s.ref_next + r))
        if (leaf_reachable(s, ray, t, RT_LDG(s.ref_leaf + r))) last = r;
    return last;
}

// Traversal stack of the culling-hierarchy walks. The first RT_SM_STACK entries of every thread live in SHARED memory
// (one column per thread of a [RT_SM_STACK][128] slab: entry k of thread t at slab[k * 128 + t], bank = t mod 32 whatever k,
// so the lanes of a warp never conflict although each is at its own depth); deeper entries — the builders cap the depth at
// 56, walks rarely go beyond a dozen pending nodes — spill to a local-memory tail that is normally never touched. Until
// round 2 the whole stack was `int stack[64]` in local memory: an LDL/STL on the critical path of every push and pop,
// and 256 B of every thread's frame competing with the scene for L1. All kernels that walk run 128-thread CTAs.
// One walk is live per thread at any time, so every walk shares the same slab. RT_SM_STACK=0: the old local array (A/B).
// Measured on B200 (profiles/r02_notes.md, r02c; kernel ms of a 16 / 2 / 4 / 2-spp frame of configs 2 / 3 / 4 / 5): local
// array 11.4 / 66.0 / 75.5 / 67.0, 12 shared entries 11.7 / 70.5 / 79.9 / 68.5, 16: 11.8 / 70.5 / 80.0 / 68.9, 24: 12.4 / 70.5 /
// 83.2 / 71.2. SLOWER, more so the larger the slab: 8-12 KB of shared memory per CTA x 8 CTAs come out of the SM's L1, which
// is what serves the hierarchy nodes, and the local stack's few hot words were L1 hits anyway. So the default is 0.
#ifndef RT_SM_STACK
#define RT_SM_STACK 0
#endif
#define RT_CTA_THREADS 128
struct TStack {
#if defined(__CUDA_ARCH__) && RT_SM_STACK > 0
    int *sm;
    int ovf[64 - RT_SM_STACK];
    __device__ __forceinline__ TStack() { __shared__ int slab[RT_SM_STACK * RT_CTA_THREADS]; sm = slab + threadIdx.x; }
    __device__ __forceinline__ void put(int sp, int v) { if (sp < RT_SM_STACK) sm[sp * RT_CTA_THREADS] = v; else ovf[sp - RT_SM_STACK] = v; }
    __device__ __forceinline__ int get(int sp) const { return sp < RT_SM_STACK ? sm[sp * RT_CTA_THREADS] : ovf[sp - RT_SM_STACK]; }
#else
    int a[64];
    RT_HD void put(int sp, int v) { a[sp] = v; }
    RT_HD int get(int sp) const { return a[sp]; }
#endif
};

// Box tests of the culling hierarchies as one FMA per plane: (l - o) * iv  ==  l * iv - o * iv with o * iv rounded once per
// ray (RT_OPT_BOXFMA). The planes then carry an ABSOLUTE error of up to 2^-24 |o * iv| (rounding of the product) next to
// the relative one of the FMA itself, so the per-ray slack constant grows by 2^-22 max_axis |o * iv| (4x that bound). A ray
// almost parallel to an axis (|d| < 1e-6) thereby loses its culling - it visits more boxes, never fewer: the hierarchies
// only filter. 6 instructions fewer per box of ~32. Measured on B200 (profiles/r01_notes.md, r01t): config 3 -2 %, configs 2,
// 4 and 5 +0.7 % (a few more spill bytes; the walks wait on loads, not on issue slots) - off THEN. Round 2, on the kernels compiled
// for one mode each (smaller frames, issue utilisation ~80 %; profiles/r02_notes.md, r03n): config 2 32.6 -> 31.6 ms (64 spp), config 3
// 66.2 -> 62.9, config 4 60.1 -> 58.5, config 5 56.4 -> 55.3 with both this and RT_OPT_CONEFMA - ON.
#ifndef RT_OPT_BOXFMA
#define RT_OPT_BOXFMA 1
#endif
struct Inv32 {
    float x, y, z;
#if RT_OPT_BOXFMA
    float ox, oy, oz;   // o * iv
    float e;            // additive slack: 1e-6 + 2^-22 max |o * iv|
#endif
};
RT_HD float safe_inv(float d) { return fabsf(d) > 1e-30f ? RT_FAST_RCP(d) : (d < 0.f ? -1e30f : 1e30f); }
RT_HD Inv32 make_inv32(const Ray &r) {
    Inv32 iv; iv.x = safe_inv(r.d.x); iv.y = safe_inv(r.d.y); iv.z = safe_inv(r.d.z);
#if RT_OPT_BOXFMA
    iv.ox = r.o.x * iv.x; iv.oy = r.o.y * iv.y; iv.oz = r.o.z * iv.z;
    iv.e = 2.4e-7f * fmaxf(fmaxf(fabsf(iv.ox), fabsf(iv.oy)), fabsf(iv.oz)) + 1e-6f;
#endif
    return iv;
}
// conservative ray/box overlap on [0, limit]; returns the entry parameter through `near`
RT_HD bool bvh_box(const Ray &r, const Inv32 &iv, float lx, float ly, float lz, float hx, float hy, float hz, float limit, float &near) {
#if RT_OPT_BOXFMA
    const float a0 = RT_FMA(lx, iv.x, -iv.ox), a1 = RT_FMA(hx, iv.x, -iv.ox);
    const float b0 = RT_FMA(ly, iv.y, -iv.oy), b1 = RT_FMA(hy, iv.y, -iv.oy);
    const float c0 = RT_FMA(lz, iv.z, -iv.oz), c1 = RT_FMA(hz, iv.z, -iv.oz);
#else
    const float a0 = (lx - r.o.x) * iv.x, a1 = (hx - r.o.x) * iv.x;
    const float b0 = (ly - r.o.y) * iv.y, b1 = (hy - r.o.y) * iv.y;
    const float c0 = (lz - r.o.z) * iv.z, c1 = (hz - r.o.z) * iv.z;
#endif
    const float tn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fminf(c0, c1));
    const float tf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fmaxf(c0, c1));
#if RT_OPT_BOXFMA
    const float slack = RT_FMA(1e-5f, fabsf(tn) + fabsf(tf), iv.e);
#else
    const float slack = 1e-5f * (fabsf(tn) + fabsf(tf)) + 1e-6f;
#endif
    near = tn;
    return (tn - slack <= tf + slack) && (tf + slack >= 0.f) && (tn - slack <= limit);
}

// Closest reachable hit of mesh m with t < limit (limit = the scene's current best / the light
// distance, exactly the bound Scene::computeIntersection / computeShadow apply to the mesh's
// answer). Same (t, triangle) as mesh_closest() whenever that answer would be accepted.
// One candidate triangle (first reference r0) against the running minimum.
template <bool STATS>
RT_HD void bvh_consider(const Ray &ray, const DScene &s, uint32_t r0, float &best_t, uint32_t &best_ref, Counters *cnt) {
    const uint32_t NONE = 0xFFFFFFFFu;
    float a, b, c;
    if (STATS) cnt->tri++;
    const float t = triangle_t<STATS>(ray, s, r0, a, b, c, cnt);
    if (t < best_t) {
        if (tri_reachable(s, ray, t, r0)) { best_t = t; best_ref = r0; }
    } else if (t == best_t && best_ref != NONE && r0 != best_ref) {
        // two different triangles at the same parameter: the reference keeps the one in the LAST
        // reachable leaf (KDTree.cpp:63-67), the FIRST reference inside a shared leaf (:42)
        const uint32_t ln = tri_last_reachable(s, ray, t, r0);
        if (ln != NONE) {
            const uint32_t lb = tri_last_reachable(s, ray, t, best_ref);
            const bool same_leaf = RT_LDG(s.ref_leaf + ln) == RT_LDG(s.ref_leaf + lb);
            if (same_leaf ? (ln < lb) : (ln > lb)) best_ref = r0;
        }
    }
}

// Closest reachable hit of mesh m with t < limit (limit = the scene's current best / the light
// distance, exactly the bound Scene::computeIntersection / computeShadow apply to the mesh's
// answer). Same (t, triangle) as mesh_closest() whenever that answer would be accepted.
//
// any_hit (shadow samples): Scene::computeShadow only asks whether the mesh's closest hit lies in (EPSILON, limit)
// (Scene.h:248-253), i.e. whether SOME reachable hit does and NONE lies in [0, EPSILON]. So the first accepted hit above
// EPSILON answers the first half, and the search goes on with the bound pulled in to EPSILON: every box further away
// is culled at once, and what is left can only find a hit at t <= EPSILON, which is then the answer (closest hit too
// near: the reference does not draw). Same boolean and same draw as the full closest-hit search; t_out is then a
// witness, not the minimum. Measured (profiles/r01_notes.md): config 3 -1.2 %, config 4 +-0, config 5 +3 % (slower) on
// the automatic kernels - the shadow rays that cost time are the unoccluded ones grazing the terrain - so it is OFF.
#ifndef RT_OPT_ANYHIT
#define RT_OPT_ANYHIT 0
#endif
#define RT_ANYHIT_STEP() \
    if (RT_OPT_ANYHIT && any_hit && best_ref != NONE) { \
        if (!(best_t > RT_EPSF)) { t_out = best_t; ref_out = best_ref; return true; } \
        found_t = best_t; found_ref = best_ref; best_t = eps_next; best_ref = NONE; \
    }
template <bool STATS>
RT_HD bool mesh_closest_bvh(const Ray &ray, const DScene &s, const DMesh &m, float limit, float &t_out, uint32_t &ref_out, Counters *cnt,
                            bool any_hit = false) {
    const uint32_t NONE = 0xFFFFFFFFu;
    float best_t = limit;
    uint32_t best_ref = NONE;
    const float eps_next = u2f(f2u(RT_EPSF) + 1u);   // t < eps_next  <=>  t <= EPSILON
    float found_t = 0.f;
    uint32_t found_ref = NONE;
    for (uint32_t k = m.always_first; k < m.always_first + m.always_count; ++k) {
        bvh_consider<STATS>(ray, s, RT_LDG(s.bvh_tris + k), best_t, best_ref, cnt);
        RT_ANYHIT_STEP()
    }
    if (m.bvh_root >= 0) {
        const Inv32 iv = make_inv32(ray);
        TStack stack;
        int sp = 0;
        int node = m.bvh_root;
#if RT_OPT_WW
        // "while-while": every lane first descends to its next leaf (box tests only), then the lanes that have one
        // test its triangles together; with one loop doing either per iteration the triangle tests ran at 2.8 of
        // 32 lanes on the pond scene (profiles/r01_notes.md)
        const int DONE = 0x7FFFFFFF;
        for (;;) {
            while (node >= 0 && node != DONE) {
                const float4 n0 = RT_LDG(s.bvh_nodes + 4 * node), n1 = RT_LDG(s.bvh_nodes + 4 * node + 1),
                             n2 = RT_LDG(s.bvh_nodes + 4 * node + 2), n3 = RT_LDG(s.bvh_nodes + 4 * node + 3);
                if (STATS) cnt->node++;
                float d0, d1;
                const bool h0 = bvh_box(ray, iv, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, best_t, d0);
                const bool h1 = bvh_box(ray, iv, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, best_t, d1);
                const int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y);
                if (h0 && h1) {
                    const bool swap = d1 < d0;
                    node = swap ? c1 : c0;
                    stack.put(sp++, swap ? c0 : c1);    // depth <= 56 by construction (rt_bvh.hpp)
                } else if (h0) node = c0;
                else if (h1) node = c1;
                else node = sp > 0 ? stack.get(--sp) : DONE;
            }
            if (node == DONE) break;
            const uint32_t code = (uint32_t)(-(node + 1));
            const uint32_t first = code >> 3, count = code & 7u;
            for (uint32_t k = first; k < first + count; ++k) {
                bvh_consider<STATS>(ray, s, RT_LDG(s.bvh_tris + k), best_t, best_ref, cnt);
                RT_ANYHIT_STEP()
            }
            if (sp == 0) break;
            node = stack.get(--sp);
        }
#else
        for (;;) {
            if (node >= 0) {
                const float4 n0 = RT_LDG(s.bvh_nodes + 4 * node), n1 = RT_LDG(s.bvh_nodes + 4 * node + 1),
                             n2 = RT_LDG(s.bvh_nodes + 4 * node + 2), n3 = RT_LDG(s.bvh_nodes + 4 * node + 3);
                if (STATS) cnt->node++;
                float d0, d1;
                const bool h0 = bvh_box(ray, iv, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, best_t, d0);
                const bool h1 = bvh_box(ray, iv, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, best_t, d1);
                const int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y);
                if (h0 && h1) {
                    const bool swap = d1 < d0;
                    node = swap ? c1 : c0;
                    stack.put(sp++, swap ? c0 : c1);    // depth <= 56 by construction (rt_bvh.hpp)
                    continue;
                }
                if (h0) { node = c0; continue; }
                if (h1) { node = c1; continue; }
            } else {
                const uint32_t code = (uint32_t)(-(node + 1));
                const uint32_t first = code >> 3, count = code & 7u;
                for (uint32_t k = first; k < first + count; ++k) {
                    bvh_consider<STATS>(ray, s, RT_LDG(s.bvh_tris + k), best_t, best_ref, cnt);
                    RT_ANYHIT_STEP()
                }
            }
            if (sp == 0) break;
            node = stack.get(--sp);
        }
#endif
    }
    if (RT_OPT_ANYHIT && any_hit && found_ref != NONE) { t_out = found_t; ref_out = found_ref; return true; }   // best_ref is NONE here
    if (best_ref == NONE) return false;
    t_out = best_t;
    ref_out = best_ref;
    return true;
}
#undef RT_ANYHIT_STEP

// All meshes of the scene in ONE per-lane loop. With one loop per mesh the lanes of a warp meet again after every mesh
// and the warp pays the slowest lane each time: the box tests of the pond scene ran at 8 of 32 lanes
// (profiles/r01_notes.md). Here a lane that has finished mesh i goes straight on to mesh i+1 while its neighbours are
// still walking, and all of them share the same three pieces of code per round: pending triangles, descent to the next
// leaf (box tests only, the tight loop), then either "open the leaf" or "mesh finished". Same candidates, same
// per-triangle routine, same acceptance rules and draw order as mesh_closest_bvh called mesh after mesh
// (Scene.h:221-228 closest, 248-253 shadow).
#define RT_WALK_DESCEND() \
        while ((unsigned int)node < (unsigned int)MESH_END) { \
            const float4 n0 = RT_LDG(s.bvh_nodes + 4 * node), n1 = RT_LDG(s.bvh_nodes + 4 * node + 1), \
                         n2 = RT_LDG(s.bvh_nodes + 4 * node + 2), n3 = RT_LDG(s.bvh_nodes + 4 * node + 3); \
            if (STATS) cnt->node++; \
            float d0, d1; \
            const bool h0 = bvh_box(ray, iv, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, best_t, d0); \
            const bool h1 = bvh_box(ray, iv, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, best_t, d1); \
            const int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y); \
            if (h0 && h1) { \
                const bool swap = d1 < d0; \
                node = swap ? c1 : c0; \
                stack.put(sp++, swap ? c0 : c1);    /* depth <= 56 by construction (rt_bvh.hpp) */ \
            } else if (h0) node = c0; \
            else if (h1) node = c1; \
            else node = sp > 0 ? stack.get(--sp) : MESH_END; \
        }
// The same descent over 4-wide nodes (RT_OPT_BVH4): four box tests per step, the children that are hit sorted by entry
// parameter with a five-comparator network, the nearest taken, the others pushed far to near. Half the dependent steps of
// the binary walk for about the same number of box tests.
#ifndef RT_OPT_BVH4
#define RT_OPT_BVH4 0
#endif
#define RT_CSWAP4(KA, CA, KB, CB) { const bool sw_ = KB < KA; const float tk_ = sw_ ? KB : KA; const int tc_ = sw_ ? CB : CA; KB = sw_ ? KA : KB; CB = sw_ ? CA : CB; KA = tk_; CA = tc_; }
#define RT_WALK_DESCEND4() \
        while ((unsigned int)node < (unsigned int)MESH_END) { \
            const float4 *nb = s.bvh4_nodes + 8 * node; \
            const float4 lx = RT_LDG(nb), ly = RT_LDG(nb + 1), lz = RT_LDG(nb + 2), hx = RT_LDG(nb + 3), hy = RT_LDG(nb + 4), hz = RT_LDG(nb + 5), cc = RT_LDG(nb + 6); \
            if (STATS) cnt->node++; \
            float k0, k1, k2, k3; \
            const bool h0 = bvh_box(ray, iv, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, best_t, k0); \
            const bool h1 = bvh_box(ray, iv, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, best_t, k1); \
            const bool h2 = bvh_box(ray, iv, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, best_t, k2); \
            const bool h3 = bvh_box(ray, iv, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, best_t, k3); \
            int c0 = (int)f2u(cc.x), c1 = (int)f2u(cc.y), c2 = (int)f2u(cc.z), c3 = (int)f2u(cc.w); \
            k0 = h0 ? k0 : FLT_MAX; k1 = h1 ? k1 : FLT_MAX; k2 = h2 ? k2 : FLT_MAX; k3 = h3 ? k3 : FLT_MAX; \
            const int nh = (int)h0 + (int)h1 + (int)h2 + (int)h3; \
            RT_CSWAP4(k0, c0, k1, c1) RT_CSWAP4(k2, c2, k3, c3) RT_CSWAP4(k0, c0, k2, c2) RT_CSWAP4(k1, c1, k3, c3) RT_CSWAP4(k1, c1, k2, c2) \
            if (nh == 0) { node = sp > 0 ? stack.get(--sp) : MESH_END; } \
            else { \
                if (nh > 3) stack.put(sp++, c3); \
                if (nh > 2) stack.put(sp++, c2); \
                if (nh > 1) stack.put(sp++, c1); \
                node = c0; \
            } \
        }
// ANYHIT (shadow samples only, mode != 0): Scene::computeShadow asks of a mesh only whether its closest hit lies in (EPSILON, limit)
// (Scene.h:248-253): the first reachable hit above EPSILON answers "some hit does", and the search goes on with the bound pulled in to
// EPSILON — every box farther away is culled at once — for a hit at t <= EPSILON, which would make the closest hit too near to count.
// Same boolean, same draw as the full closest-hit search (the reasoning of RT_OPT_ANYHIT in mesh_closest_bvh).
template <bool STATS, bool ANYHIT = false>
RT_HD void meshes_walk_merged(const DScene &s, const Ray &ray, int mode, Rng &rng, Hit &h, bool &blocked, bool &done, Counters *cnt) {
    if (done || s.n_meshes <= 0) return;
    const uint32_t NONE = 0xFFFFFFFFu;
    const int MESH_END = 0x7FFFFFFD;
    const Inv32 iv = make_inv32(ray);
    TStack stack;
    int sp = 0, mi = 0;
    // the always-tested triangles of a mesh come first, like a leaf
    uint32_t k = s.meshes[0].always_first, kend = k + s.meshes[0].always_count;
#if RT_OPT_BVH4
    bool wide = s.meshes[0].bvh4_root >= 0;
    int node = wide ? s.meshes[0].bvh4_root : (s.meshes[0].bvh_root >= 0 ? s.meshes[0].bvh_root : MESH_END);
#else
    int node = s.meshes[0].bvh_root >= 0 ? s.meshes[0].bvh_root : MESH_END;
#endif
    float best_t = h.t;
    uint32_t best_ref = NONE;
    bool found = false;   // ANYHIT: this mesh has a reachable hit in (EPSILON, limit)
    const float eps_next = u2f(f2u(RT_EPSF) + 1u);   // t < eps_next  <=>  t <= EPSILON
    if (STATS) cnt->mesh++;
    for (;;) {
        for (; k < kend; ++k) {
            bvh_consider<STATS>(ray, s, RT_LDG(s.bvh_tris + k), best_t, best_ref, cnt);
            if (ANYHIT && best_ref != NONE && best_t > RT_EPSF) { found = true; best_t = eps_next; best_ref = NONE; }
        }
#if RT_OPT_BVH4
        if (wide) { RT_WALK_DESCEND4() } else { RT_WALK_DESCEND() }
#else
        RT_WALK_DESCEND()
#endif
        if (node < 0) {   // a leaf: its triangles are tested at the top of the next round
            const uint32_t code = (uint32_t)(-(node + 1));
            k = code >> 3; kend = k + (code & 7u);
            node = sp > 0 ? stack.get(--sp) : MESH_END;
            continue;
        }
        // mesh mi is finished (ANYHIT: a hit found above EPSILON counts unless the search below EPSILON found one as well: best_ref set)
        if (ANYHIT ? (found && best_ref == NONE) : (best_ref != NONE && best_t < h.t && best_t > RT_EPSF)) {
            if (mode == 0) { h.type = 3; h.obj = mi; h.t = best_t; h.ref = best_ref; }
            else { if (STATS) cnt->rnd++; if (rng.next() > RT_LDG(s.mesh_transparency + mi)) { blocked = true; done = true; break; } }
        }
        if (++mi >= s.n_meshes) break;
        const DMesh &m = s.meshes[mi];
        if (STATS) cnt->mesh++;
        k = m.always_first; kend = k + m.always_count;
#if RT_OPT_BVH4
        wide = m.bvh4_root >= 0;
        node = wide ? m.bvh4_root : (m.bvh_root >= 0 ? m.bvh_root : MESH_END);
#else
        node = m.bvh_root >= 0 ? m.bvh_root : MESH_END;
#endif
        sp = 0;
        best_t = h.t; best_ref = NONE; found = false;
    }
}

#undef RT_WALK_DESCEND
#undef RT_WALK_DESCEND4
#undef RT_CSWAP4

// Can meshes_walk_merged find anything for this ray below `limit`? Only if some mesh has always-tested triangles or the ray
// passes the FIRST step of a mesh's walk (one of the two boxes of its root node, same conservative test, same limit):
// otherwise every walk ends at its root with no candidate. The wavefront uses this to send only such rays to the mesh-walk
// kernel; the others keep their analytic hit, which is what the walk would have left them with.
RT_HD bool ray_touches_meshes(const DScene &s, const Ray &ray, float limit) {
    const Inv32 iv = make_inv32(ray);
    for (int mi = 0; mi < s.n_meshes; ++mi) {
        const DMesh &m = s.meshes[mi];
        if (m.always_count > 0u) return true;
        if (m.bvh_root < 0) continue;
        const float4 n0 = RT_LDG(s.bvh_nodes + 4 * m.bvh_root), n1 = RT_LDG(s.bvh_nodes + 4 * m.bvh_root + 1), n2 = RT_LDG(s.bvh_nodes + 4 * m.bvh_root + 2);
        float d;
        if (bvh_box(ray, iv, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, limit, d)) return true;
        if (bvh_box(ray, iv, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, limit, d)) return true;
    }
    return false;
}

// ---- analytic primitives through their culling hierarchy (variant 3) ----------------------------
// Visits every sphere/square whose padded, motion-swept box the ray can touch within [0, limit] and
// calls f(seq) for it (seq: sphere i -> i, square j -> n_spheres + j). `limit` is re-read after every
// call so a closest-hit caller can shrink it.
template <bool STATS, class F>
RT_HD void analytic_candidates(const DScene &s, const Ray &ray, const float &limit, F &&f, Counters *cnt) {
    const Inv32 iv = make_inv32(ray);
    // per-ray enlargement (see build_analytic_accel): quadratic term x 1/r of the spheres below, plus a linear term
    const float dist = length(ray.o - ld3(s.abvh_c)) + s.abvh_r, dist_s = length(ray.o - ld3(s.abvh_cs)) + s.abvh_rs;
    const float kq = 32.f * 5.96e-8f * dist_s * dist_s, kl = 64.f * 5.96e-8f * dist;
    TStack stack;
    int sp = 0;
    int node = s.abvh_root;
    for (;;) {
        if (node >= 0) {
            const float4 *const an = RT_ABVH_NODES(s);
            const float4 n0 = RT_ABVH_LD(an + 4 * node), n1 = RT_ABVH_LD(an + 4 * node + 1),
                         n2 = RT_ABVH_LD(an + 4 * node + 2), n3 = RT_ABVH_LD(an + 4 * node + 3);
            if (STATS) cnt->node++;
            const float p0 = kq * n3.z + kl, p1 = kq * n3.w + kl;
            float d0, d1;
            const bool h0 = bvh_box(ray, iv, n0.x - p0, n0.y - p0, n0.z - p0, n0.w + p0, n1.x + p0, n1.y + p0, limit, d0);
            const bool h1 = bvh_box(ray, iv, n1.z - p1, n1.w - p1, n2.x - p1, n2.y + p1, n2.z + p1, n2.w + p1, limit, d1);
            const int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y);
            if (h0 && h1) {
                const bool swap = d1 < d0;
                node = swap ? c1 : c0;
                stack.put(sp++, swap ? c0 : c1);
                continue;
            }
            if (h0) { node = c0; continue; }
            if (h1) { node = c1; continue; }
        } else {
            const uint32_t code = (uint32_t)(-(node + 1));
            const uint32_t first = code >> 3, count = code & 7u;
            for (uint32_t k = first; k < first + count; ++k) f(RT_LDG(s.abvh_prims + k));
        }
        if (sp == 0) break;
        node = stack.get(--sp);
    }
}

// ---- scene: closest hit ------------------------------------------------------------------------
// Scene::computeIntersection: spheres, then squares, then meshes; a candidate replaces the current
// best only if t < best.t && t >= EPSILON (strict <: first object wins ties within a type, and
// spheres beat squares beat meshes).
template <bool STATS>
RT_HD Hit closest_hit(const DScene &s, const Ray &ray, float &aux_u, float &aux_v, Counters *cnt) {
    Hit h; h.type = 0; h.obj = -1; h.t = FLT_MAX; h.ref = 0;
    if (STATS) cnt->closest++;
    const SphereRay sr = make_sphere_ray(ray);
    for (int i = 0; i < s.n_spheres; ++i) {
        if (STATS) cnt->sphere++;
        const float t = sphere_t(ray, sr, RT_LDG(s.sph_a + i), RT_LDG(s.sph_b + i));
        if (t < h.t && t > RT_EPSF) { h.type = 1; h.obj = i; h.t = t; }
    }
    for (int i = 0; i < s.n_squares; ++i) {
        if (STATS) cnt->square++;
        float u, v;
        const float t = square_t(ray, s.squares[i], u, v);
        if (t < h.t && t > RT_EPSF) { h.type = 2; h.obj = i; h.t = t; aux_u = u; aux_v = v; }
    }
    if (s.n_meshes > 0) {
        const RayInv inv = make_inv(ray);
        for (int i = 0; i < s.n_meshes; ++i) {
            if (STATS) cnt->mesh++;
            float t; uint32_t ref;
            if (mesh_closest<STATS>(ray, inv, s, s.meshes[i], t, ref, cnt) && t < h.t && t > RT_EPSF) {
                h.type = 3; h.obj = i; h.t = t; h.ref = ref;
            }
        }
    }
    return h;
}

// ---- scene: occlusion --------------------------------------------------------------------------
// Scene::computeShadow: every candidate blocker with EPSILON <= t < tLight draws one
// random_float(); it blocks if the draw exceeds its transparency. For a mesh the candidate is the
// mesh's CLOSEST hit (t >= 0), so a hit below EPSILON hides farther triangles.
template <bool STATS>
RT_HD bool shadow_hit(const DScene &s, const Ray &ray, float t_light, Rng &rng, Counters *cnt) {
    if (STATS) cnt->shadow++;
    const SphereRay sr = make_sphere_ray(ray);
    for (int i = 0; i < s.n_spheres; ++i) {
        if (STATS) cnt->sphere++;
        const float4 b = RT_LDG(s.sph_b + i);
        const float t = sphere_t(ray, sr, RT_LDG(s.sph_a + i), b);
        if (t < t_light && t > RT_EPSF) {
            if (STATS) cnt->rnd++;
            if (rng.next() > b.w) return true;
        }
    }
    for (int i = 0; i < s.n_squares; ++i) {
        if (STATS) cnt->square++;
        float u, v;
        const float t = square_t(ray, s.squares[i], u, v);
        if (t < t_light && t > RT_EPSF) {
            if (STATS) cnt->rnd++;
            if (rng.next() > RT_LDG(s.sq_transparency + i)) return true;
        }
    }
    if (s.n_meshes > 0) {
        const RayInv inv = make_inv(ray);
        for (int i = 0; i < s.n_meshes; ++i) {
            if (STATS) cnt->mesh++;
            float t; uint32_t ref;
            if (mesh_closest<STATS>(ray, inv, s, s.meshes[i], t, ref, cnt) && t < t_light && t > RT_EPSF) {
                if (STATS) cnt->rnd++;
                if (rng.next() > RT_LDG(s.mesh_transparency + i)) return true;
            }
        }
    }
    return false;
}

// theta, phi of a sphere normal (Sphere.h:129-130) mapped to texture coordinates (Material.cpp:98, Scene.h:277)
RT_COLD void sphere_uv(V3 n, float &tu, float &tv) {
    const float theta = (float)acos((double)n.y * -1.);
    const float phi = (float)(atan2((double)n.z * -1., (double)n.x) + RT_PI);
    tu = (float)((double)phi / (2 * RT_PI));
    tv = (float)((double)theta / RT_PI);
}

// ---- textures ----------------------------------------------------------------------------------
RT_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// texel address of Material::texture / get_normal (Material.cpp:82-86,119-123): fmod wrap in fp64,
// v flipped, truncation to int — nearest texel, no filtering. The index is clamped to the image
// (the reference would read out of bounds for negative u; squares and spheres never produce one).
RT_HD int texel_index(const DImage &im, float u, float v, float sx, float sy) {
    u = (float)fmod((double)(u * sx), 1.);
    v = (float)(1 - fmod((double)(v * sy), 1.));
    const int x = (int)(u * (im.w - 1));
    const int y = (int)(v * (im.h - 1));
    return clampi(y * im.w + x, 0, im.w * im.h - 1);
}

// Material::texture (Material.cpp:63-92). `color` is left untouched for Texture_None.
template <bool STATS>
RT_COLD void material_texture(const DScene &s_, const DMaterial &m, V3 &color, float u, float v, Counters *cnt) {
    const DScene &s = RT_S(s_);
    if (m.texture_type == 1) {
        color = ((int)(u * m.tsx) % 2 == (int)(v * m.tsy) % 2) ? ld3(m.checker1) : ld3(m.checker2);
    } else if (m.texture_type == 2) {
        DImage im; im.w = 0; im.h = 0; im.rgb = nullptr;
        if (m.image >= 0) im = s.textures[m.image];
        if (im.w < 1 || im.h < 1) {   // Material.cpp:74-80: magenta/black 8x8 fallback
            color = ((int)(u * 8.) % 2 == (int)(v * 8.) % 2) ? v3(0.f) : v3(1.f, 0.f, 1.f);
            return;
        }
        if (STATS) cnt->tex++;
        const unsigned char *p = im.rgb + 3 * (size_t)texel_index(im, u, v, m.tsx, m.tsy);
        color = v3((float)(p[0] / 255.), (float)(p[1] / 255.), (float)(p[2] / 255.));
    }
}

// Material::emit (Material.cpp:13-24)
template <bool STATS>
RT_COLD V3 material_emit(const DScene &s_, const DMaterial &m, float u, float v, Counters *cnt) {
    const DScene &s = RT_S(s_);
    if (!m.emissive) return v3(0.f);
    V3 c = v3(0.f);
    if (m.texture_type == 0) c = ld3(m.light_color);
    else material_texture<STATS>(s, m, c, u, v, cnt);
    return c * m.light_intensity;
}

// Material::get_normal (Material.cpp:114-130): tangent-space map, T/B = the square's stale
// m_right_vector / m_up_vector members.
template <bool STATS>
RT_COLD V3 material_normal(const DScene &s_, const DMaterial &m, V3 n, float u, float v, V3 T, V3 B, Counters *cnt) {
    const DScene &s = RT_S(s_);
    if (m.normal_map < 0) return n;
    const DImage im = s.normal_maps[m.normal_map];
    if (im.w < 1 || im.h < 1) return n;   // reference would dereference an empty image; defined as "no map"
    if (STATS) cnt->tex++;
    const unsigned char *p = im.rgb + 3 * (size_t)texel_index(im, u, v, m.tsx, m.tsy);
    const float a = (float)(p[0] / 127.5 - 1.), b = (float)(p[1] / 127.5 - 1.), c = (float)(p[2] / 127.5 - 1.);
    return normalized(a * T + b * B + c * n);
}

// Scene::skyboxTexture (Scene.h:149-161)
template <bool STATS>
RT_COLD V3 sky_color(const DScene &s_, V3 d, int n_remaining, Counters *cnt) {
    const DScene &s = RT_S(s_);
    if (s.sky.w < 1 || s.sky.h < 1) {
        if (s.dark_sky) return v3(0.f);
        const float a = (float)(0.5 * ((double)d.y + 1.0));
        // (1-a)*white + a*blue*(N+1): only the second term is scaled (operator precedence)
        return (float)(1.0 - (double)a) * v3(1.f, 1.f, 1.f) + (a * v3(0.5f, 0.7f, 1.0f)) * (float)(n_remaining + 1);
    }
    if (STATS) cnt->tex++;
    // atan2f / asinf, correctly rounded (fp64 function rounded once) — see oracle/libm_pin.cpp
    const float at = (float)atan2((double)d.z, (double)d.x);
    const float as = (float)asin((double)d.y);
    const float u = (float)(0.5 + (double)at / (2 * RT_PI));
    const float v = (float)(0.5 - (double)as / RT_PI);
    const int x = (int)(u * s.sky.w);
    const int y = (int)(v * s.sky.h);
    // the reference does not clamp (u == 1 or v == 1 read past the row / the image)
    const int idx = clampi(y * s.sky.w + x, 0, s.sky.w * s.sky.h - 1);
    const unsigned char *p = s.sky.rgb + 3 * (size_t)idx;
    return v3((float)(p[0] / 255.), (float)(p[1] / 255.), (float)(p[2] / 255.)) * (float)n_remaining;
}

// ---- scattering --------------------------------------------------------------------------------
RT_HD V3 reflect_dir(V3 d, V3 n) { return d - (2.f * dot(d, n)) * n; }            // Functions.cpp:38-40
RT_HD V3 refract_dir(V3 d, V3 n, float eta) {                                      // Functions.cpp:42-47
    const float cos_theta = fminr(dot(d, n), 1.0f);
    const V3 perp = eta * (d + cos_theta * n);
    const V3 par = (float)(-sqrt(fabs(1.0 - (double)dot(perp, perp)))) * n;
    return perp + par;
}
RT_HD float schlick(float cosine, float ref_idx) {                                 // Functions.cpp:49-54
    float r0 = (1.f - ref_idx) / (1.f + ref_idx);
    r0 = r0 * r0;
    return (float)((double)r0 + (double)(1.f - r0) * pow((double)(1.f - cosine), 5.0));
}
// Material::scatter (Material.cpp:26-60). Returns the new ray (origin P + EPSILON*dir).
template <bool STATS>
RT_COLD Ray material_scatter(const DMaterial &m, const Ray &in, V3 n, V3 P, Rng &rng, Counters *cnt) {
    V3 dir = v3(0.f);
    if (m.type == 1) {          // glass (inverted convention and the -0.6 test are the reference's, A.1-8)
        const float ri = dot(in.d, n) > 0.f ? (float)(1. / (double)m.index_medium) : m.index_medium;
        const float cos_theta = fminr(dot(in.d * -1.f, n), 1.0f);
        const float sin_theta = (float)sqrt(1. - (double)(cos_theta * cos_theta));
        const bool cannot_refract = (double)(ri * sin_theta) - 0.6 > 1.0;
        bool do_reflect = cannot_refract;
        if (!do_reflect) {
            if (STATS) cnt->rnd++;
            do_reflect = schlick(cos_theta, ri) > rng.next();
        }
        dir = do_reflect ? reflect_dir(in.d, n) : refract_dir(in.d, n, ri);
    } else if (m.type == 0) {   // diffuse: normal + normalised cube point
        if (STATS) cnt->rnd += 3;
        dir = n + random_unit_vector(rng);
        if (length(dir) <= RT_EPSF) dir = n;
    } else if (m.type == 2) {   // mirror
        dir = reflect_dir(in.d, n);
    }
    dir = normalized(dir);
    return make_ray(P + RT_EPSF * dir, dir, in.time);
}

// ---- one path ----------------------------------------------------------------------------------
// Scene::rayTrace = rayTraceRecursive(ray, MAXBOUNCES) / MAXBOUNCES, unrolled into a loop. The
// recursion returns (color_k + result_{k+1} (*) kd_k) + e_k; to reproduce its rounding the loop
// records (color, kd, e) per depth and folds them back to front once the path ends.
template <bool STATS>
RT_HD V3 trace_path(const DScene &s, Ray ray, Rng &rng, int max_bounces, int nb_ech, Counters *cnt) {
    V3 rec_c[RT_MAX_BOUNCES], rec_kd[RT_MAX_BOUNCES], rec_e[RT_MAX_BOUNCES];
    int depth = 0;
    V3 tail = v3(0.f);
    for (int N = max_bounces; N > 0; --N) {
        float hu = 0.f, hv = 0.f;
        const Hit h = closest_hit<STATS>(s, ray, hu, hv, cnt);
        if (h.type == 0) { tail = sky_color<STATS>(s, ray.d, N, cnt); break; }

        V3 P, n, kd, e;
        float transparency;
        const DMaterial *mat;
        if (h.type == 1) {
            // Sphere.h:124-130 for the winning sphere only
            mat = s.sph_mat + h.obj;
            const float4 a = RT_LDG(s.sph_a + h.obj), b = RT_LDG(s.sph_b + h.obj);
            const V3 c = v3(a.x, a.y, a.z) + ray.time * v3(b.x, b.y, b.z);
            P = ray.o + h.t * ray.d;
            n = normalized(P - c);
            kd = ld3(mat->kd);
            float tu = 0.f, tv = 0.f;
            if (mat->texture_type != 0) {
                // theta/phi (Sphere.h:129-130) feed only sphere_texture() and a textured emit(); the
                // reference evaluates them for every candidate, the result is the same without
                sphere_uv(n, tu, tv);
                material_texture<STATS>(s, *mat, kd, tu, tv, cnt);   // sphere_texture
            }
            e = material_emit<STATS>(s, *mat, tu, tv, cnt);
        } else if (h.type == 2) {
            mat = s.sq_mat + h.obj;
            const DSquare &q = s.squares[h.obj];
            P = ray.o + h.t * ray.d;
            n = ld3(q.n);
            kd = ld3(mat->kd);
            material_texture<STATS>(s, *mat, kd, hu, hv, cnt);
            n = material_normal<STATS>(s, *mat, n, hu, hv, ld3(q.tan_r), ld3(q.tan_u), cnt);
            e = material_emit<STATS>(s, *mat, hu, hv, cnt);
        } else {
            mat = s.mesh_mat + h.obj;
            const DMesh &m = s.meshes[h.obj];
            float w0, w1, w2;
            triangle_t<false>(ray, s, h.ref, w0, w1, w2, nullptr);   // recompute the barycentrics of the winner
            P = ray.o + h.t * ray.d;
            const float4 pl = RT_LDG(s.tri_plane + h.ref);
            n = v3(pl.x, pl.y, pl.z);                                // flat face normal (Triangle.h:119)
            kd = ld3(mat->kd);
            const uint32_t ti = f2u(RT_LDG(s.tri_den + h.ref).y);
            if (m.color_type == 0) {
                const uint32_t i0 = m.triangles[3 * ti], i1 = m.triangles[3 * ti + 1], i2 = m.triangles[3 * ti + 2];
                kd = w0 * ld3(m.vert_colors + 3 * i0) + w1 * ld3(m.vert_colors + 3 * i1) + w2 * ld3(m.vert_colors + 3 * i2);
            } else if (m.color_type == 1) {
                kd = ld3(m.face_colors + 3 * ti);
            }
            e = v3(0.f);
        }
        transparency = mat->transparency;

        // direct lighting with soft shadows (Scene.h:305-334)
        V3 color = v3(0.f);
        for (int i = 0; i < s.n_lights; ++i) {
            const V3 lp = ld3(s.lights[i].pos);
            const V3 L = normalized(lp - P);
            const float dotLN = dot(L, n);
            // lights[0].material for every light; (1 - transparency) in fp64 (Scene.h:311)
            color = color + (comp_product(ld3(s.lights[0].color), kd) * fmaxr(0.0f, dotLN)) * (float)(1. - (double)transparency);
            int blocked = 0;
            const float delta = s.lights[i].radius / 2.f;
            for (int j = 0; j < nb_ech; ++j) {
                if (STATS) cnt->rnd += 3;
                const V3 lj = lp + random_unit_vector(rng) * delta;
                const V3 Lj = normalized(lj - P);
                const float t_light = length(lj - P);
                const Ray sh = make_ray(P + Lj * RT_EPSF, Lj, ray.time);
                if (shadow_hit<STATS>(s, sh, t_light, rng, cnt)) ++blocked;
            }
            const float shadow = (float)(1. - (double)((float)blocked / (float)nb_ech));
            color = color * shadow;   // scales the light accumulated so far, earlier lights included
        }

        ray = material_scatter<STATS>(*mat, ray, n, P, rng, cnt);
        rec_c[depth] = color; rec_kd[depth] = kd; rec_e[depth] = e;
        ++depth;
    }
    V3 r = tail;
    for (int k = depth - 1; k >= 0; --k) r = (rec_c[k] + comp_product(r, rec_kd[k])) + rec_e[k];
    r = v3(0.f) + r;
    return r / (float)max_bounces;
}

// ---- one path, as a ray-level state machine ----------------------------------------------------
// Same arithmetic and the same random_float() order as trace_path above, cut at every ray: a lane
// holds one path and, per step, ONE ray to intersect — a closest-hit ray (Scene::computeIntersection)
// or one soft-shadow sample (Scene::computeShadow). The kernel (k_render_regen) runs the steps of
// 32 paths in lockstep, so lanes that are at different bounces, or one shading and one shadowing,
// still execute the intersection loops together; a lane whose path ends takes a new path at once.


struct PathRecs { V3 rec_c[RT_MAX_BOUNCES], rec_kd[RT_MAX_BOUNCES], rec_e[RT_MAX_BOUNCES]; };
struct CandList { uint32_t v[RT_LC_MAXC]; };
struct PathState {
    Rng rng;
    Ray ray;             // the ray to intersect next
    int mode;            // 0 closest hit, 1 shadow sample, 2 idle (no path), 3 collect the occluder candidates of light `light` (variant 5)
    int N;               // remaining bounces (NRemainingBounces of rayTraceRecursive)
    int depth;           // records written so far
    uint32_t path;       // index of the path in the chunk (output slot)
    // context of the hit being lit (valid while mode == 1)
    V3 P, n, kd, e, color, in_d;
    const DMaterial *mat;
    int light, j, blocked;
    float t_light;
    int max_bounces;
    uint32_t cm0, cm1, cm2, cm3;   // variant 5: analytic primitives (sequence index = bit) that can occlude light `light` from P
    // ... and the triangles the cone towards that light can touch: (mesh << 27 | first leaf ref of the triangle), in mesh
    // order; cl_n < 0 = too many for the list, shadow samples walk the mesh hierarchies themselves
    // The list is an array of its own (CandList, declared by the kernel next to the path state), reached through this
    // pointer: as a MEMBER array it was indexed dynamically, and that alone kept the whole path state in the thread's
    // local-memory frame instead of in registers.
    uint32_t *cl;
    int cl_n;
    // per-depth radiance records: the state-machine kernels keep them in the thread's frame (PathRecs, 576 B, declared
    // by the kernel); the wavefront kernels (variant 6) keep them in global memory instead: the record of depth d of path
    // slot `path` is the three float4 at wf_rec[3 * (d * wf_stride + path)] ({kd, flags} {colour} {e}: 48 contiguous bytes,
    // one or two sectors, where three separate planes cost three half-used ones), and `recs` is null — their frame
    // does not carry the arrays
    PathRecs *recs;
    float4 *wf_rec;
    unsigned long long wf_stride;
};

// Scene::computeIntersection and Scene::computeShadow over the same loops. mode 0: closest hit into
// (h, hu, hv). mode 1: `blocked` is computeShadow's return value; candidates draw from rng in the
// reference's order (spheres, squares, meshes) until one blocks. Lanes that have their answer
// (blocked, or idle) skip the tests; the loops end early only when the whole warp is done.
// The analytic primitives (spheres, then squares) of intersect_ray; `done` in: this lane has nothing to test (it still takes part
// in the warp votes); out: blocked.
template <bool STATS, bool ACCEL>
RT_HD void intersect_ray_analytic(const DScene &s, const Ray &ray, int mode, float t_light, Rng &rng, Hit &h, float &hu, float &hv,
                                  bool &blocked, bool &done, Counters *cnt) {
    h.type = 0; h.obj = -1; h.t = (mode == 0) ? FLT_MAX : t_light; h.ref = 0;
    blocked = false;
    if (STATS && !done) { if (mode == 0) cnt->closest++; else if (mode == 1) cnt->shadow++; }
    const SphereRay sr = make_sphere_ray(ray);
    if (ACCEL && s.abvh_root >= 0 && s.n_spheres + s.n_squares >= 24) {   // below 24 the linear loops are as fast (the small hierarchy exists for variants 5/6)
        // variant 3: candidates from the culling hierarchy, exact tests, order semantics restored
        if (!done) {
            const int ns = s.n_spheres;
            int best_seq = 0x7FFFFFFF;
            uint32_t mask[4] = {0u, 0u, 0u, 0u};
            // one traversal for both ray kinds (h.t is the limit either way: current best, or the light distance)
            analytic_candidates<STATS>(s, ray, h.t, [&](uint32_t seq) {
                float t, u = 0.f, v = 0.f;
                if ((int)seq < ns) { if (STATS) cnt->sphere++; t = sphere_t(ray, sr, RT_LDG(s.sph_a + seq), RT_LDG(s.sph_b + seq)); }
                else { if (STATS) cnt->square++; t = square_t(ray, s.squares[seq - ns], u, v); }
                if (!(t > RT_EPSF)) return;
                if (mode == 0) {
                    // sequential strict '<' of the reference == smallest t, earliest in sequence on ties
                    if (t < h.t || (t == h.t && best_seq != 0x7FFFFFFF && (int)seq < best_seq)) {
                        best_seq = (int)seq; h.t = t;
                        if ((int)seq < ns) { h.type = 1; h.obj = (int)seq; } else { h.type = 2; h.obj = (int)seq - ns; hu = u; hv = v; }
                    }
                } else if (t < h.t) {
                    mask[seq >> 5] |= 1u << (seq & 31u);
                }
            }, cnt);
            if (mode != 0) {
                // replay the candidate blockers in the reference's order: one draw each until one blocks
                for (int w = 0; w < 4 && !done; ++w) {
                    uint32_t m = mask[w];
                    while (m) {
                        const int bit = RT_FFS((int)m) - 1;
                        m &= m - 1u;
                        const int seq = w * 32 + bit;
                        const float tr = seq < ns ? RT_LDG(s.sph_b + seq).w : RT_LDG(s.sq_transparency + (seq - ns));
                        if (STATS) cnt->rnd++;
                        if (rng.next() > tr) { blocked = true; done = true; break; }
                    }
                }
            }
        }
    } else {
        for (int i = 0; i < s.n_spheres; ++i) {
            if ((i & 3) == 0 && RT_WARP_ALL(done)) break;
            if (done) continue;
            if (STATS) cnt->sphere++;
            const float4 b = RT_LDG(s.sph_b + i);
            const float t = sphere_t(ray, sr, RT_LDG(s.sph_a + i), b);
            if (t < h.t && t > RT_EPSF) {
                if (mode == 0) { h.type = 1; h.obj = i; h.t = t; }
                else { if (STATS) cnt->rnd++; if (rng.next() > b.w) { blocked = true; done = true; } }
            }
        }
        for (int i = 0; i < s.n_squares; ++i) {
            if ((i & 3) == 0 && RT_WARP_ALL(done)) break;
            if (done) continue;
            if (STATS) cnt->square++;
            float u, v;
            const float t = square_t(ray, s.squares[i], u, v);
            if (t < h.t && t > RT_EPSF) {
                if (mode == 0) { h.type = 2; h.obj = i; h.t = t; hu = u; hv = v; }
                else { if (STATS) cnt->rnd++; if (rng.next() > RT_LDG(s.sq_transparency + i)) { blocked = true; done = true; } }
            }
        }
    }
}
template <bool STATS, bool ACCEL>
RT_HD void intersect_ray(const DScene &s, const Ray &ray, int mode, float t_light, Rng &rng, Hit &h, float &hu, float &hv,
                         bool &blocked, Counters *cnt) {
    bool done = (mode == 2);
    intersect_ray_analytic<STATS, ACCEL>(s, ray, mode, t_light, rng, h, hu, hv, blocked, done, cnt);
    if (ACCEL && RT_OPT_MESH_MERGED) {
        meshes_walk_merged<STATS>(s, ray, mode, rng, h, blocked, done, cnt);
    } else if (ACCEL) {
        // variant 3: per-mesh exact culling traversal (mesh_closest_bvh), same acceptance rules
        for (int i = 0; i < s.n_meshes; ++i) {
            if (done) break;
            if (STATS) cnt->mesh++;
            float t; uint32_t ref;
            if (mesh_closest_bvh<STATS>(ray, s, s.meshes[i], h.t, t, ref, cnt, mode != 0) && t < h.t && t > RT_EPSF) {
                if (mode == 0) { h.type = 3; h.obj = i; h.t = t; h.ref = ref; }
                else { if (STATS) cnt->rnd++; if (rng.next() > RT_LDG(s.mesh_transparency + i)) { blocked = true; done = true; } }
            }
        }
    } else if (s.n_meshes > 0) {
        // Every lane walks the shared pre-order node array with its own ray: mesh after mesh, node
        // after node, and inside a leaf triangle after triangle. A lane is always in one of two
        // states — "next step is a triangle test" or "next step is a node (box test / leaf open /
        // mesh hand-over)". Each round the warp votes (__ballot_sync/__popc) and executes only the
        // step kind the MAJORITY of its lanes needs; the others wait one round. That keeps the two
        // expensive bodies (triangle test, fp64 slab test) convergent although the 32 rays are in
        // different nodes: a plain per-lane nested loop ran at 3.5 of 32 lanes (profiles/r01).
        const RayInv inv = make_inv(ray);
        const uint32_t NONE = 0xFFFFFFFFu;
        bool active = !done;
        int mi = 0;
        uint32_t i = 0, mesh_end = 0, k = 0, kend = 0;
        if (active) { i = s.meshes[0].node_begin; mesh_end = s.meshes[0].node_end; if (STATS) cnt->mesh++; }
        float leaf_t = FLT_MAX, best_t = FLT_MAX;
        uint32_t leaf_ref = NONE, best_ref = NONE;
        for (;;) {
            const bool want_tri = active && k < kend;
            const bool want_node = active && !want_tri;
            const unsigned int bt = RT_BALLOT(want_tri), bn = RT_BALLOT(want_node);
            if ((bt | bn) == 0u) break;
            if (RT_POPC(bt) >= RT_POPC(bn)) {
                if (want_tri) {
                    float a, b, c;
                    if (STATS) cnt->tri++;
                    const float t = triangle_t<STATS>(ray, s, k, a, b, c, cnt);
                    if (t < leaf_t) { leaf_t = t; leaf_ref = k; }        // first triangle wins ties inside a leaf
                    ++k;
                    if (k == kend && leaf_ref != NONE && leaf_t <= best_t) { best_t = leaf_t; best_ref = leaf_ref; }  // last leaf wins ties
                }
            } else if (want_node) {
                if (i >= mesh_end) {
                    // mesh mi is done: Scene-level acceptance (Scene.h:221-228 closest, 248-253 shadow)
                    if (best_ref != NONE && best_t < h.t && best_t > RT_EPSF) {
                        if (mode == 0) { h.type = 3; h.obj = mi; h.t = best_t; h.ref = best_ref; }
                        else { if (STATS) cnt->rnd++; if (rng.next() > RT_LDG(s.mesh_transparency + mi)) { blocked = true; active = false; } }
                    }
                    if (++mi >= s.n_meshes) active = false;
                    if (active) {
                        i = s.meshes[mi].node_begin; mesh_end = s.meshes[mi].node_end;
                        best_t = FLT_MAX; best_ref = NONE;
                        if (STATS) cnt->mesh++;
                    }
                } else {
                    const float4 lo = RT_LDG(s.node_lo + i), hi = RT_LDG(s.node_hi + i);
                    const uint32_t hw = f2u(hi.w);
                    const bool leaf = (hw & 0x80000000u) != 0u;
                    if (STATS) cnt->node++;
                    if (!slab_hit(ray, inv, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) {
                        i = leaf ? i + 1 : f2u(lo.w);
                    } else {
                        if (leaf) { k = f2u(lo.w); kend = k + (hw & 0x7FFFFFFFu); leaf_t = FLT_MAX; leaf_ref = NONE; }
                        ++i;
                    }
                }
            }
        }
    }
}

// ---- variant 4: ONE warp-voted walk over the analytic hierarchy and every mesh hierarchy ---------
// Same candidates, same exact tests and same order rules as the ACCEL path of intersect_ray; what
// changes is the control flow. A lane is always either about to TEST a primitive (sphere, square or
// triangle) or about to handle a NODE (pop, box tests, open a leaf, finish a phase); each round the
// warp executes only the kind most of its lanes want (__ballot_sync/__popc), so the two bodies stay
// convergent although every lane walks its own path through its own hierarchy — per-lane loops ran
// at 7 of 32 lanes on the pond scene (profiles/r01_notes.md).
template <bool STATS>
RT_HD void intersect_ray_voted(const DScene &s, const Ray &ray, int mode, float t_light, Rng &rng, Hit &h, float &hu, float &hv,
                               bool &blocked, Counters *cnt) {
    const uint32_t NONE = 0xFFFFFFFFu;
    const int EMPTY = 0x7FFFFFFF;
    h.type = 0; h.obj = -1; h.t = (mode == 0) ? FLT_MAX : t_light; h.ref = 0;
    blocked = false;
    bool active = (mode != 2);
    if (STATS) { if (mode == 0) cnt->closest++; else if (mode == 1) cnt->shadow++; }
    const SphereRay sr = make_sphere_ray(ray);
    const int ns = s.n_spheres;
    if (s.abvh_root < 0) {   // few analytic primitives: the reference's linear loops
        for (int i = 0; i < s.n_spheres; ++i) {
            if (!active) break;
            if (STATS) cnt->sphere++;
            const float4 b = RT_LDG(s.sph_b + i);
            const float t = sphere_t(ray, sr, RT_LDG(s.sph_a + i), b);
            if (t < h.t && t > RT_EPSF) {
                if (mode == 0) { h.type = 1; h.obj = i; h.t = t; }
                else { if (STATS) cnt->rnd++; if (rng.next() > b.w) { blocked = true; active = false; } }
            }
        }
        for (int i = 0; i < s.n_squares; ++i) {
            if (!active) break;
            if (STATS) cnt->square++;
            float u, v;
            const float t = square_t(ray, s.squares[i], u, v);
            if (t < h.t && t > RT_EPSF) {
                if (mode == 0) { h.type = 2; h.obj = i; h.t = t; hu = u; hv = v; }
                else { if (STATS) cnt->rnd++; if (rng.next() > RT_LDG(s.sq_transparency + i)) { blocked = true; active = false; } }
            }
        }
    }
    const Inv32 iv = make_inv32(ray);
    float kq = 0.f, kl = 0.f;          // per-ray box enlargement, analytic phase only (build_analytic_accel)
    int phase = -1;                    // -1: analytic hierarchy; m >= 0: mesh m
    int node = EMPTY, sp = 0;
    TStack stack;
    uint32_t k = 0, kend = 0;
    float best_t = h.t;
    uint32_t best_ref = NONE;
    int best_seq = 0x7FFFFFFF;
    uint32_t m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
    if (s.abvh_root >= 0) {
        const float dist = length(ray.o - ld3(s.abvh_c)) + s.abvh_r, dist_s = length(ray.o - ld3(s.abvh_cs)) + s.abvh_rs;
        kq = 32.f * 5.96e-8f * dist_s * dist_s; kl = 64.f * 5.96e-8f * dist;
        node = s.abvh_root;
    }
    for (;;) {
        const bool want_test = active && k < kend;
        const bool want_node = active && !want_test;
        const unsigned int bt = RT_BALLOT(want_test), bn = RT_BALLOT(want_node);
        if ((bt | bn) == 0u) break;
        if (RT_POPC(bt) >= RT_POPC(bn)) {
            if (want_test) {
                if (phase < 0) {
                    const uint32_t seq = RT_LDG(s.abvh_prims + k);
                    float t, u = 0.f, v = 0.f;
                    if ((int)seq < ns) { if (STATS) cnt->sphere++; t = sphere_t(ray, sr, RT_LDG(s.sph_a + seq), RT_LDG(s.sph_b + seq)); }
                    else { if (STATS) cnt->square++; t = square_t(ray, s.squares[seq - ns], u, v); }
                    if (t > RT_EPSF) {
                        if (mode == 0) {
                            if (t < h.t || (t == h.t && best_seq != 0x7FFFFFFF && (int)seq < best_seq)) {
                                best_seq = (int)seq; h.t = t;
                                if ((int)seq < ns) { h.type = 1; h.obj = (int)seq; } else { h.type = 2; h.obj = (int)seq - ns; hu = u; hv = v; }
                            }
                        } else if (t < h.t) {
                            const uint32_t bit = 1u << (seq & 31u);
                            const uint32_t w = seq >> 5;
                            if (w == 0u) m0 |= bit; else if (w == 1u) m1 |= bit; else if (w == 2u) m2 |= bit; else m3 |= bit;
                        }
                    }
                } else {
                    bvh_consider<STATS>(ray, s, RT_LDG(s.bvh_tris + k), best_t, best_ref, cnt);
                }
                ++k;
            }
        } else if (want_node) {
            if (node == EMPTY) {
                if (sp > 0) {
                    node = stack.get(--sp);
                } else {
                    // the current phase is exhausted
                    if (phase < 0) {
                        if (s.abvh_root >= 0 && mode != 0) {
                            // replay the candidate blockers in the reference's order: one draw each until one blocks
                            for (int w = 0; w < 4 && active; ++w) {
                                uint32_t m = w == 0 ? m0 : (w == 1 ? m1 : (w == 2 ? m2 : m3));
                                while (m) {
                                    const int bit = RT_FFS((int)m) - 1;
                                    m &= m - 1u;
                                    const int seq = w * 32 + bit;
                                    const float tr = seq < ns ? RT_LDG(s.sph_b + seq).w : RT_LDG(s.sq_transparency + (seq - ns));
                                    if (STATS) cnt->rnd++;
                                    if (rng.next() > tr) { blocked = true; active = false; break; }
                                }
                            }
                        }
                    } else {
                        // mesh `phase` is done: Scene-level acceptance (Scene.h:221-228 closest, 248-253 shadow)
                        if (best_ref != NONE && best_t < h.t && best_t > RT_EPSF) {
                            if (mode == 0) { h.type = 3; h.obj = phase; h.t = best_t; h.ref = best_ref; }
                            else { if (STATS) cnt->rnd++; if (rng.next() > RT_LDG(s.mesh_transparency + phase)) { blocked = true; active = false; } }
                        }
                    }
                    ++phase;
                    if (phase >= s.n_meshes) active = false;
                    if (active) {
                        const DMesh &m = s.meshes[phase];
                        if (STATS) cnt->mesh++;
                        k = m.always_first; kend = k + m.always_count;
                        node = m.bvh_root >= 0 ? m.bvh_root : EMPTY;
                        best_t = h.t; best_ref = NONE;
                        kq = 0.f; kl = 0.f;
                    }
                }
            }
            if (active && node != EMPTY && k >= kend) {
                if (node >= 0) {
                    const float4 *nodes = phase < 0 ? s.abvh_nodes : s.bvh_nodes;
                    const float4 n0 = RT_LDG(nodes + 4 * node), n1 = RT_LDG(nodes + 4 * node + 1), n2 = RT_LDG(nodes + 4 * node + 2),
                                 n3 = RT_LDG(nodes + 4 * node + 3);
                    if (STATS) cnt->node++;
                    const float p0 = kq * n3.z + kl, p1 = kq * n3.w + kl;
                    const float limit = phase < 0 ? h.t : best_t;
                    float d0, d1;
                    const bool h0 = bvh_box(ray, iv, n0.x - p0, n0.y - p0, n0.z - p0, n0.w + p0, n1.x + p0, n1.y + p0, limit, d0);
                    const bool h1 = bvh_box(ray, iv, n1.z - p1, n1.w - p1, n2.x - p1, n2.y + p1, n2.z + p1, n2.w + p1, limit, d1);
                    const int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y);
                    if (h0 && h1) {
                        const bool swap = d1 < d0;
                        node = swap ? c1 : c0;
                        stack.put(sp++, swap ? c0 : c1);
                    } else if (h0) node = c0;
                    else if (h1) node = c1;
                    else node = EMPTY;
                } else {
                    const uint32_t code = (uint32_t)(-(node + 1));
                    k = code >> 3; kend = k + (code & 7u);
                    node = EMPTY;
                }
            }
        }
    }
}

// ---- variant 5: occluder candidates per (hit point, light) instead of per shadow ray ---------------
// Scene::rayTraceRecursive fires NB_ECH shadow rays per light from the same point P towards points
// lj = light.pos + delta * u, |u| = 1 (Scene.h:325-330): 81 % of all rays on config 2, each of which
// walked the culling hierarchy on its own in variant 3. Every point such a ray can report a hit at,
// P + t*Lj with t < |lj - P|, lies on the segment P..lj, hence within s*delta of the point
// P + s*(light.pos - P) of the central segment (s in [0,1]): inside a CONE with apex P. Variant 5
// walks the hierarchy ONCE with that cone (mode 3 of the path state machine), collecting every
// primitive whose padded box the cone can touch into a 128-bit mask; the NB_ECH shadow samples then
// run the reference's own sphere/square arithmetic on the candidates only, in sequence order, with
// the reference's random_float() draws. The filter can only add candidates (it is conservative by
// construction: max-norm cone, padded boxes, slack on every comparison), never drop an occluder, so
// the bits are those of variants 1-4 (tests: test_kernel_variants_agree_bit_for_bit, test_exact_culling_at_scale).
//
// The box test is the slab test generalised to a cone of max-norm radius s*delta around O + s*D:
//   lo - s*delta <= O + s*D <= hi + s*delta   per axis
//   <=>  s*(D + delta) >= lo - O   and   s*(D - delta) <= hi - O
// Each inequality bounds s from below or above depending on the sign of its coefficient; with
// delta = 0 it is the ordinary slab test, so closest-hit rays (mode 0) and cones (mode 3) run the
// same instructions side by side.
#ifndef RT_OPT_CONEFMA
#define RT_OPT_CONEFMA 1   /* the same FMA form as bvh_box (RT_OPT_BOXFMA) for the cone: 7 more live registers per lane */
#endif
struct Cone {
    V3 o;
#if RT_OPT_CONEFMA
    float oax, oay, oaz, obx, oby, obz;   // o / (D + delta), o / (D - delta)
    float e;                              // additive slack: 1e-6 + 2^-22 max |o * i|
#endif
    float iax, iay, iaz;    // 1 / (D + delta)
    float ibx, iby, ibz;    // 1 / (D - delta)
    bool nnx, nny, nnz;     // both coefficients negative: the upper face bounds s from below
    bool pnx, pny, pnz;     // D + delta >= 0 > D - delta: both faces bound s from below, none from above
};
RT_HD void cone_axis(float d, float delta, float &ia, float &ib, bool &nn, bool &pn) {
    const float A = d + delta, B = d - delta;
    ia = safe_inv(A); ib = safe_inv(B);
    nn = A < 0.f;
    pn = !(A < 0.f) && (B < 0.f);
}
RT_HD Cone make_cone(V3 o, V3 D, float delta) {
    Cone c; c.o = o;
    cone_axis(D.x, delta, c.iax, c.ibx, c.nnx, c.pnx);
    cone_axis(D.y, delta, c.iay, c.iby, c.nny, c.pny);
    cone_axis(D.z, delta, c.iaz, c.ibz, c.nnz, c.pnz);
#if RT_OPT_CONEFMA
    c.oax = o.x * c.iax; c.oay = o.y * c.iay; c.oaz = o.z * c.iaz;
    c.obx = o.x * c.ibx; c.oby = o.y * c.iby; c.obz = o.z * c.ibz;
    c.e = 2.4e-7f * fmaxf(fmaxf(fmaxf(fabsf(c.oax), fabsf(c.oay)), fmaxf(fabsf(c.oaz), fabsf(c.obx))), fmaxf(fabsf(c.oby), fabsf(c.obz))) + 1e-6f;
#endif
    return c;
}
#if RT_OPT_CONEFMA
#define RT_CONE_PLANES(LO, HI, O, IA, IB, OA, OB) const float x0 = RT_FMA((LO), (IA), -(OA)), x1 = RT_FMA((HI), (IB), -(OB));
#else
#define RT_CONE_PLANES(LO, HI, O, IA, IB, OA, OB) const float x0 = ((LO) - (O)) * (IA), x1 = ((HI) - (O)) * (IB);
#endif
#define RT_CONE_AXIS(LO, HI, O, IA, IB, OA, OB, NN, PN, TN, TF)            \
    {                                                                      \
        RT_CONE_PLANES(LO, HI, O, IA, IB, OA, OB)                          \
        const float un = (NN) ? x1 : x0, uf = (NN) ? x0 : x1;              \
        TN = (PN) ? fmaxf(un, uf) : un;                                    \
        TF = (PN) ? FLT_MAX : uf;                                          \
    }
// conservative cone/box overlap for s in [0, limit]; `near` = entry parameter (child ordering only)
RT_HD bool cone_box(const Cone &c, float lx, float ly, float lz, float hx, float hy, float hz, float limit, float &near) {
    float tnx, tfx, tny, tfy, tnz, tfz;
#if RT_OPT_CONEFMA
    RT_CONE_AXIS(lx, hx, c.o.x, c.iax, c.ibx, c.oax, c.obx, c.nnx, c.pnx, tnx, tfx)
    RT_CONE_AXIS(ly, hy, c.o.y, c.iay, c.iby, c.oay, c.oby, c.nny, c.pny, tny, tfy)
    RT_CONE_AXIS(lz, hz, c.o.z, c.iaz, c.ibz, c.oaz, c.obz, c.nnz, c.pnz, tnz, tfz)
#else
    RT_CONE_AXIS(lx, hx, c.o.x, c.iax, c.ibx, 0.f, 0.f, c.nnx, c.pnx, tnx, tfx)
    RT_CONE_AXIS(ly, hy, c.o.y, c.iay, c.iby, 0.f, 0.f, c.nny, c.pny, tny, tfy)
    RT_CONE_AXIS(lz, hz, c.o.z, c.iaz, c.ibz, 0.f, 0.f, c.nnz, c.pnz, tnz, tfz)
#endif
    const float tn = fmaxf(fmaxf(tnx, tny), tnz);
    const float tf = fminf(fminf(fminf(tfx, tfy), tfz), limit);
#if RT_OPT_CONEFMA
    const float slack = RT_FMA(1e-5f, fabsf(tn) + fabsf(tf), c.e);
#else
    const float slack = 1e-5f * (fabsf(tn) + fabsf(tf)) + 1e-6f;
#endif
    near = tn;
    return (tn - slack <= tf + slack) && (tf + slack >= 0.f);
}

// Candidates that provably cannot occlude ANY shadow sample of the cone are dropped when the mask is collected.
// Every sample ray leaves P + EPSILON*Lj along Lj = normalize(lj - P), |lj - light.pos| <= delta (1.0001 delta + 1e-6
// here), so for a fixed vector w:   Lj . w  >=  (D . w - delta*|w|) / |lj - P|,   D = light.pos - P.
//   * non-glass square: Square::intersect returns "no hit" whenever d . n >= 0 (Square.h:78-86). With w = n
//     (unit): D . n - delta > margin  =>  every sample sees the back face or the edge-on plane.
//   * sphere: the near root -b - sqrt(delta) is <= 0 < EPSILON whenever b = 2 d . (o - c) >= 0 (see sphere_t). With
//     w = P - c (and o - c = w + EPSILON*Lj only helps): D . w - delta*|w| > margin  =>  the sphere lies behind every
//     sample ray. This is the lit side of the very sphere that was hit, and everything on the far side of P.
// margin = 1e-4 * |w| * (|D| + delta): three orders of magnitude above the rounding of the fp32 dot products the
// reference-exact tests evaluate (a few eps * |w|), so "cannot occlude" here implies FLT_MAX there.
#ifndef RT_OPT_LC_ANYHIT
#define RT_OPT_LC_ANYHIT 1   /* any-hit mesh walk for the shadow samples of the overflow queue (meshes_walk_merged<.., ANYHIT>) */
#endif
#ifndef RT_OPT_LC_HULL
#define RT_OPT_LC_HULL 1
#endif
// Does a ball (centre P + v, radius r) stay clear of the hull H of P and the ball of sample points, H = union over s in [0, 1] of
// ball(P + s D, s delta)? Every shadow sample of the light runs inside H (from P + EPSILON L to a point within delta of the light's
// centre). The distance of the ball's centre from H is min_s f(s), f(s) = |v - s D| - s delta: f is convex (a norm of an affine
// function minus a linear one), its stationary point solves (a s - b)^2 = delta^2 |v - s D|^2 with a s >= b, i.e.
// s* = (b + delta sqrt((a cc - b^2) / (a - delta^2))) / a with a = D.D, b = v.D, cc = v.v, and the minimum over the segment is at s*
// clamped to [0, 1.0001] (the ray ends EPSILON beyond its sample point). An error ds in s* changes f by f'' ds^2 / 2 only. "Clear" asks
// for f > r + 1e-4 of every length involved; the caller adds the slop of the reference's own test to r.
RT_HD bool lc_hull_misses(V3 v, V3 D, float delta, float Dlen, float r) {
    const float A = dot(D, D), d2 = delta * delta;
    if (!(A > 4.f * d2)) return false;   // P inside or near the ball of sample points: no statement
    const float B = dot(v, D), CC = dot(v, v);
    const float cr = fmaxf(A * CC - B * B, 0.f);
    // approximate reciprocals / square roots (2 ulp): three orders of magnitude inside the 1e-4 margin below
    float sx = (B + delta * RT_FAST_SQRT(cr * RT_FAST_RCP(A - d2))) * RT_FAST_RCP(A);
    sx = sx < 0.f ? 0.f : (sx > 1.0001f ? 1.0001f : sx);
    const V3 q = v - sx * D;
    const float f = RT_FAST_SQRT(dot(q, q)) - sx * delta;
    return f > r + 1e-4f * (RT_FAST_SQRT(CC) + Dlen + delta + r) + 1e-6f;
}
// Is the sphere (centre P + v, radius r > 0) hit by EVERY shadow sample of the cone, with EPSILON < t < t_light? Sufficient: P clearly outside
// it, the whole sphere nearer than every sample point (|v| + r < |D| - delta), and every sample direction inside the cone the sphere
// subtends shrunk by the slop of the reference's sphere test: a sample direction deviates from D by at most asin(delta / |D|), D from v by
// theta, and sin(theta + beta) <= sin theta + sin beta, so sin theta + delta / |D| < (r - pad) / |v| (with cos theta > 0) suffices; pad as in
// lc_cannot_occlude, 1e-3 relative on top. Then the discriminant of the reference's test is positive by ~25x its rounding error, the near root
// is >= |v| - r > 1e-3 >> EPSILON and <= |v| < t_light.
RT_HD bool lc_sphere_covers(V3 v, V3 D, float delta, float Dlen, float r) {
    const float vl = length(v);
    const float pad = 64.f * 5.96e-8f * (vl * vl / r + r) + 1e-3f * r;
    if (!(r > pad) || !(vl - r > 1e-3f + 1e-3f * vl) || !(vl + r < (Dlen - delta) * 0.999f) || !(Dlen > 4.f * delta)) return false;
    const float ct = dot(v, D) / (vl * Dlen);
    if (!(ct > 0.f)) return false;
    const float st2 = fmaxf(1.f - ct * ct, 0.f);
    return sqrtf(st2) + delta / Dlen < ((r - pad) / vl) * 0.999f - 1e-5f;
}
RT_HD bool lc_cannot_occlude(const DScene &s, uint32_t seq, int ns, V3 P, float time, V3 D, float delta, float Dlen) {
    if ((int)seq < ns) {
        const float4 a = RT_LDG(s.sph_a + seq), b = RT_LDG(s.sph_b + seq);
        const V3 c = v3(a.x, a.y, a.z) + time * v3(b.x, b.y, b.z);
        const V3 w = P - c;
        const float wl = RT_FAST_SQRT(dot(w, w));   // only compared with margins of 1e-4
        if (dot(D, w) - delta * wl > 1e-4f * wl * (Dlen + delta) + 1e-12f) return true;
#if RT_OPT_LC_HULL
        // * sphere, part 2: the sphere stays clear of the hull of P and the ball of sample points (lc_hull_misses). The box test that found
        //   this candidate is a max-norm cone against a swept, padded box: in config 2, 62 % of the (hit, light) pairs that went on to be
        //   sampled never hit any of their candidates; with this test 37 % fewer pairs are sampled (profiles/r02_notes.md, r03y). The radius
        //   carries the slop of the reference's sphere test: a ray that misses the sphere by up to ~2.5 eps (|o - c|^2 / r + r) can still be
        //   reported as a hit; taken 25x.
        const float r = fabsf(a.w);
        if (r > 0.f && lc_hull_misses(c - P, D, delta, Dlen, r + 64.f * 5.96e-8f * (wl * wl * RT_FAST_RCP(r) + r))) return true;
#endif
        return false;
    }
    const DSquare &q = s.squares[seq - ns];
    if (q.glass) return false;
    return dot(D, ld3(q.n)) - delta > 1e-4f * (Dlen + delta) + 1e-12f;
}
// UMBRA: every candidate is opaque (transparency 0), there are no candidate triangles, and one candidate sphere is hit by every sample of
// the cone (lc_sphere_covers). Then each of the NB_ECH samples draws its three direction numbers and exactly ONE more — for the first
// candidate it hits in sequence order, whichever that is — and is blocked iff that draw exceeds the transparency 0 (Scene.h:236-247): no
// direction, no test needed. A draw of exactly 0 (1 in 2^24) would pass the occluder and go on to the next candidate: such a pair is left
// to the sample kernel untouched (the stream is only committed when all NB_ECH draws block). In config 2, 40 % of the (hit, light) pairs that
// used to be sampled are such pairs (profiles/r02_notes.md, r03z). MEASURED: exact (107 GPU tests, CPU simulation), the sample kernels of config 2
// lose 25-35 % of their time (1.27 -> 0.97 ms per 16-spp frame) — and the classify kernels, which run the check for every pair with a candidate,
// gain more than that (2.47 -> 3.23 ms: issue utilisation 68-74 % -> 55-60 % around the call): frame 28.6 -> 30.1 ms at 64 spp. OFF.
#ifndef RT_OPT_LC_UMBRA
#define RT_OPT_LC_UMBRA 0
#endif
// Out of line and by value: inlined, its registers cost the classify kernel's cone walk more than the samples it saves (config 2 28.6 ->
// 29.9 ms), and a reference into the path state would pin the whole state in the thread's frame. Returns the stream position after the
// NB_ECH samples, or 0xFFFFFFFF when the pair is not an umbra pair.
RT_COLD uint32_t lc_umbra_cold(const DScene &s_, V3 P, float time, int light, uint32_t cm0, uint32_t cm1, uint32_t cm2, uint32_t cm3, Rng r, int nb_ech) {
    const DScene &s = RT_S(s_);
    const DLight &L = s.lights[light];
    const V3 D = ld3(L.pos) - P;
    const float delta = (L.radius / 2.f) * 1.0001f + 1e-6f, Dlen = length(D);
    const int ns = s.n_spheres;
    bool cover = false;
    for (int w = 0; w < 4; ++w) {
        uint32_t m = w == 0 ? cm0 : (w == 1 ? cm1 : (w == 2 ? cm2 : cm3));
        while (m) {
            const int seq = w * 32 + RT_FFS((int)m) - 1;
            m &= m - 1u;
            if (seq < ns) {
                const float4 a = RT_LDG(s.sph_a + seq), b = RT_LDG(s.sph_b + seq);
                if (b.w != 0.f) return 0xFFFFFFFFu;
                if (!cover) cover = lc_sphere_covers((v3(a.x, a.y, a.z) + time * v3(b.x, b.y, b.z)) - P, D, delta, Dlen, fabsf(a.w));
            } else if (RT_LDG(s.sq_transparency + (seq - ns)) != 0.f) return 0xFFFFFFFFu;
        }
    }
    if (!cover) return 0xFFFFFFFFu;
    for (int j = 0; j < nb_ech; ++j) { r.ctr += 3u; if (!(r.next() > 0.f)) return 0xFFFFFFFFu; }
    return r.ctr;
}
RT_HD bool lc_umbra(const DScene &s, PathState &st, int nb_ech) {
    if (!RT_OPT_LC_UMBRA || st.cl_n != 0) return false;
    const uint32_t c = lc_umbra_cold(s, st.P, st.ray.time, st.light, st.cm0, st.cm1, st.cm2, st.cm3, st.rng, nb_ech);
    if (c == 0xFFFFFFFFu) return false;   // (a stream never gets this far: 2^32 - 1 draws)
    st.rng.ctr = c;
    return true;
}
// No candidate at all (and no mesh in the cone): the NB_ECH samples of this light are all unoccluded. Each would
// draw three numbers for its direction and nothing else (Scene.h:325-330, computeShadow draws only per candidate hit),
// so the stream advances by 3*NB_ECH and shadow = 1 - 0/NB_ECH = 1 leaves the colour unchanged (x * 1.0f == x).
RT_HD bool lc_light_unoccluded(const PathState &st) { return (st.cm0 | st.cm1 | st.cm2 | st.cm3) == 0u && st.cl_n == 0; }

// Scene::computeShadow (Scene.h:235-247) over the analytic occluder candidates of the current light, for the shadow
// sample in st.ray: ascending sequence index = the reference's order (spheres by index, then squares by index), one
// draw per candidate hit until one blocks. Returns computeShadow's answer so far (true: blocked, the meshes are skipped).
template <bool STATS>
RT_HD bool lc_shadow_analytic(const DScene &s, PathState &st, Counters *cnt) {
    const Ray &ray = st.ray;
    const int ns = s.n_spheres;
    if (STATS) cnt->shadow++;
    const SphereRay sr = make_sphere_ray(ray);
    for (int w = 0; w < 4; ++w) {
        uint32_t m = w == 0 ? st.cm0 : (w == 1 ? st.cm1 : (w == 2 ? st.cm2 : st.cm3));
        while (m) {
            const int seq = w * 32 + RT_FFS((int)m) - 1;
            m &= m - 1u;
            float t, tr, u, v;
            if (seq < ns) {
                if (STATS) cnt->sphere++;
                const float4 b = RT_LDG(s.sph_b + seq);
                t = sphere_t(ray, sr, RT_LDG(s.sph_a + seq), b);
                tr = b.w;
            } else {
                if (STATS) cnt->square++;
                t = square_t(ray, s.squares[seq - ns], u, v);
                tr = RT_LDG(s.sq_transparency + (seq - ns));
            }
            if (t < st.t_light && t > RT_EPSF) {
                if (STATS) cnt->rnd++;
                if (st.rng.next() > tr) return true;
            }
        }
    }
    return false;
}

// One step of the variant-5 state machine for the lanes selected by `mine`.
//   run_t (warp-uniform) true : lanes in mode 0 (closest hit) and mode 3 (collect candidates) walk the
//                               analytic hierarchy together; mode 0 goes on to the meshes.
//   run_t false               : lanes in mode 1 test their candidate mask, then the meshes.
// Requires s.abvh_root >= 0 (the kernel is only selected for such scenes).
// with_meshes = false: closest-hit rays stop after the analytic primitives (the wavefront walks the meshes in a kernel of its
// own, for the rays that touch a mesh at all: ray_touches_meshes below).
// CLOSEST = true (the wavefront's trace kernels: every ray is a closest-hit ray): the cone degenerates to the ray (delta = 0), so
// the walk runs the plain slab test (bvh_box, three reciprocals per ray instead of six and no sign selects per plane) on the
// same padded boxes with the same slack. Both tests are conservative filters in front of the same exact primitive tests with
// the same order rule (smallest t, earliest in sequence on ties), so the hit is the same; only the number of boxes visited may
// differ by a rounding. RT_OPT_LC_SLAB=0 keeps the cone test there (A/B).
#ifndef RT_OPT_LC_SLAB
#define RT_OPT_LC_SLAB 1
#endif
// COLLECT = true (the wavefront's classify kernel, k_wf_light phase 1: every lane that walks is in mode 3): only the cone walk and
// the candidate collection are compiled in — no exact sphere / square tests in the leaves, no shadow-sample branch, no closest-hit
// mesh walk. RT_OPT_LC_COLLECT=0 keeps the general function there (A/B).
#ifndef RT_OPT_LC_COLLECT
#define RT_OPT_LC_COLLECT 1
#endif
// SAMPLE = true (the sample kernel of a one-light scene, k_wf_light phase 3: every lane that works is in mode 1): only the
// candidate tests of a shadow sample (masks, triangle list or mesh walk) are compiled in.
// SAMPLE_ = 2 / 3: moreover every lane is known to hold a candidate-triangle list (cl_n >= 0) / to have overflowed it (cl_n < 0): the
// park queue and the overflow queue are sampled by separate launches, each compiled for its half.
template <bool STATS, bool CLOSEST = false, bool COLLECT_ = false, int SAMPLE_ = 0, bool FLAT = true>
RT_HD void intersect_lc(const DScene &s, PathState &st, bool run_t_, bool mine, Hit &h, float &hu, float &hv, bool &blocked,
                        Counters *cnt, bool with_meshes = true, bool flat = false) {
    const Ray &ray = st.ray;
    const bool COLLECT = COLLECT_ && RT_OPT_LC_COLLECT, SAMPLE = SAMPLE_ != 0 && RT_OPT_LC_COLLECT;
    const bool run_t = SAMPLE ? false : ((CLOSEST && RT_OPT_LC_SLAB) || COLLECT ? true : run_t_);
    const int mode = SAMPLE ? 1 : (CLOSEST && RT_OPT_LC_SLAB ? 0 : (COLLECT ? 3 : st.mode));
    h.type = 0; h.obj = -1; h.t = (mode == 0) ? FLT_MAX : st.t_light; h.ref = 0;
    blocked = false;
    bool done = !mine;
    const int ns = s.n_spheres;
    if (FLAT && CLOSEST && RT_OPT_LC_SLAB && flat && s.abvh_flat) {   // FLAT = false: not compiled in (kernels that never see such a scene in practice)
        // FLAT (a scene with <= 32 analytic primitives, from bounce 1 on): every lane tests ALL the primitives' boxes — one convergent loop,
        // ~25 instructions per box — and keeps the ones it touches in a 32-bit mask; then each lane runs the reference's test on its own
        // candidates in ascending sequence order (strict '<': the reference's tie rule as it stands). The hierarchy walk over the pool
        // scene's 30 primitives ran 16.8 of 32 lanes from bounce 1 on, ~2150 warp instructions per batch; this is ~750 + the longest
        // candidate list of the warp x 45 (profiles/r02_notes.md, r04e).
        if (mine) {
            if (STATS) cnt->closest++;
            const Inv32 iv32 = make_inv32(ray);
            const SphereRay sr = make_sphere_ray(ray);
            const float dist = length(ray.o - ld3(s.abvh_c)) + s.abvh_r, dist_s = length(ray.o - ld3(s.abvh_cs)) + s.abvh_rs;
            const float kq = 32.f * 5.96e-8f * dist_s * dist_s, kl = 64.f * 5.96e-8f * dist;
            const int n = ns + s.n_squares;
            uint32_t cand = 0u;
            for (int i = 0; i < n; ++i) {
                const float4 lo = RT_LDG(s.abvh_flat + 2 * i), hi = RT_LDG(s.abvh_flat + 2 * i + 1);
                const float p = kq * lo.w + kl;
                float tn;
                if (STATS) cnt->node++;
                if (bvh_box(ray, iv32, lo.x - p, lo.y - p, lo.z - p, hi.x + p, hi.y + p, hi.z + p, FLT_MAX, tn)) cand |= 1u << i;
            }
            while (cand) {
                const int seq = RT_FFS((int)cand) - 1;
                cand &= cand - 1u;
                float t, u = 0.f, v = 0.f;
                if (seq < ns) { if (STATS) cnt->sphere++; t = sphere_t(ray, sr, RT_LDG(s.sph_a + seq), RT_LDG(s.sph_b + seq)); }
                else { if (STATS) cnt->square++; t = square_t(ray, s.squares[seq - ns], u, v); }
                if (t < h.t && t > RT_EPSF) {
                    h.t = t;
                    if (seq < ns) { h.type = 1; h.obj = seq; } else { h.type = 2; h.obj = seq - ns; hu = u; hv = v; }
                }
            }
        }
    } else
    if (run_t) {
        if (mine) {
            const bool SLAB = CLOSEST && RT_OPT_LC_SLAB;
            const bool collect = COLLECT || (!SLAB && (mode == 3));
            if (STATS && !collect) cnt->closest++;
            Cone cone;
            Inv32 iv32;
            float limit_c = 1.0001f;
            V3 coneD = v3(0.f);
            float cone_delta = 0.f, cone_len = 0.f;
            if (SLAB) {
                iv32 = make_inv32(ray);
                cone.o = ray.o;
            } else if (collect) {
                const DLight &L = s.lights[st.light];
                coneD = ld3(L.pos) - st.P;
                cone_delta = (L.radius / 2.f) * 1.0001f + 1e-6f;
                cone_len = length(coneD);
                cone = make_cone(st.P, coneD, cone_delta);
            } else {
                cone = make_cone(ray.o, ray.d, 0.f);
            }
            const SphereRay sr = make_sphere_ray(ray);
            const float dist = length(cone.o - ld3(s.abvh_c)) + s.abvh_r, dist_s = length(cone.o - ld3(s.abvh_cs)) + s.abvh_rs;
            const float kq = 32.f * 5.96e-8f * dist_s * dist_s, kl = 64.f * 5.96e-8f * dist;
            uint32_t m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
            int best_seq = 0x7FFFFFFF;
            TStack stack;
            int sp = 0;
            int node = s.abvh_root;
#if RT_OPT_WW_LC
            // while-while: descend to the next leaf with box tests only, then test its primitives — the lanes of a warp
            // reach their leaves at different iterations, and with one loop doing either the primitive tests ran at 7 of
            // 32 lanes on the secondary rays of config 2 (profiles/r01_notes.md)
            const int DONE = 0x7FFFFFFF;
            for (;;) {
                while (node >= 0 && node != DONE) {
                    const float4 *const an = RT_ABVH_NODES(s);
                    const float4 n0 = RT_ABVH_LD(an + 4 * node), n1 = RT_ABVH_LD(an + 4 * node + 1),
                                 n2 = RT_ABVH_LD(an + 4 * node + 2), n3 = RT_ABVH_LD(an + 4 * node + 3);
                    if (STATS) cnt->node++;
                    const float p0 = kq * n3.z + kl, p1 = kq * n3.w + kl;
                    const float limit = collect ? limit_c : h.t;
                    float d0, d1;
                    bool h0, h1;
                    if (SLAB) {
                        h0 = bvh_box(ray, iv32, n0.x - p0, n0.y - p0, n0.z - p0, n0.w + p0, n1.x + p0, n1.y + p0, limit, d0);
                        h1 = bvh_box(ray, iv32, n1.z - p1, n1.w - p1, n2.x - p1, n2.y + p1, n2.z + p1, n2.w + p1, limit, d1);
                    } else {
                        h0 = cone_box(cone, n0.x - p0, n0.y - p0, n0.z - p0, n0.w + p0, n1.x + p0, n1.y + p0, limit, d0);
                        h1 = cone_box(cone, n1.z - p1, n1.w - p1, n2.x - p1, n2.y + p1, n2.z + p1, n2.w + p1, limit, d1);
                    }
                    const int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y);
                    if (h0 && h1) {
                        const bool swap = d1 < d0;
                        node = swap ? c1 : c0;
                        stack.put(sp++, swap ? c0 : c1);
                    } else if (h0) node = c0;
                    else if (h1) node = c1;
                    else node = sp > 0 ? stack.get(--sp) : DONE;
                }
                if (node == DONE) break;
                {
                    const uint32_t code = (uint32_t)(-(node + 1));
                    const uint32_t first = code >> 3, count = code & 7u;
                    for (uint32_t k = first; k < first + count; ++k) {
                        const uint32_t seq = RT_LDG(s.abvh_prims + k);
                        if (collect) {
                            if (lc_cannot_occlude(s, seq, ns, st.P, ray.time, coneD, cone_delta, cone_len)) continue;
                            const uint32_t bit = 1u << (seq & 31u), w = seq >> 5;
                            if (w == 0u) m0 |= bit; else if (w == 1u) m1 |= bit; else if (w == 2u) m2 |= bit; else m3 |= bit;
                            continue;
                        }
                        float t, u = 0.f, v = 0.f;
                        if ((int)seq < ns) { if (STATS) cnt->sphere++; t = sphere_t(ray, sr, RT_LDG(s.sph_a + seq), RT_LDG(s.sph_b + seq)); }
                        else { if (STATS) cnt->square++; t = square_t(ray, s.squares[seq - ns], u, v); }
                        if (!(t > RT_EPSF)) continue;
                        // sequential strict '<' of the reference == smallest t, earliest in sequence on ties
                        if (t < h.t || (t == h.t && best_seq != 0x7FFFFFFF && (int)seq < best_seq)) {
                            best_seq = (int)seq; h.t = t;
                            if ((int)seq < ns) { h.type = 1; h.obj = (int)seq; } else { h.type = 2; h.obj = (int)seq - ns; hu = u; hv = v; }
                        }
                    }
                }
                if (sp == 0) break;
                node = stack.get(--sp);
            }
#else
            for (;;) {
                if (node >= 0) {
                    const float4 *const an = RT_ABVH_NODES(s);
                    const float4 n0 = RT_ABVH_LD(an + 4 * node), n1 = RT_ABVH_LD(an + 4 * node + 1),
                                 n2 = RT_ABVH_LD(an + 4 * node + 2), n3 = RT_ABVH_LD(an + 4 * node + 3);
                    if (STATS) cnt->node++;
                    const float p0 = kq * n3.z + kl, p1 = kq * n3.w + kl;
                    const float limit = collect ? limit_c : h.t;
                    float d0, d1;
                    const bool h0 = cone_box(cone, n0.x - p0, n0.y - p0, n0.z - p0, n0.w + p0, n1.x + p0, n1.y + p0, limit, d0);
                    const bool h1 = cone_box(cone, n1.z - p1, n1.w - p1, n2.x - p1, n2.y + p1, n2.z + p1, n2.w + p1, limit, d1);
                    const int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y);
                    if (h0 && h1) {
                        const bool swap = d1 < d0;
                        node = swap ? c1 : c0;
                        stack.put(sp++, swap ? c0 : c1);
                        continue;
                    }
                    if (h0) { node = c0; continue; }
                    if (h1) { node = c1; continue; }
                } else {
                    const uint32_t code = (uint32_t)(-(node + 1));
                    const uint32_t first = code >> 3, count = code & 7u;
                    for (uint32_t k = first; k < first + count; ++k) {
                        const uint32_t seq = RT_LDG(s.abvh_prims + k);
                        if (collect) {
                            if (lc_cannot_occlude(s, seq, ns, st.P, ray.time, coneD, cone_delta, cone_len)) continue;
                            const uint32_t bit = 1u << (seq & 31u), w = seq >> 5;
                            if (w == 0u) m0 |= bit; else if (w == 1u) m1 |= bit; else if (w == 2u) m2 |= bit; else m3 |= bit;
                            continue;
                        }
                        float t, u = 0.f, v = 0.f;
                        if ((int)seq < ns) { if (STATS) cnt->sphere++; t = sphere_t(ray, sr, RT_LDG(s.sph_a + seq), RT_LDG(s.sph_b + seq)); }
                        else { if (STATS) cnt->square++; t = square_t(ray, s.squares[seq - ns], u, v); }
                        if (!(t > RT_EPSF)) continue;
                        // sequential strict '<' of the reference == smallest t, earliest in sequence on ties
                        if (t < h.t || (t == h.t && best_seq != 0x7FFFFFFF && (int)seq < best_seq)) {
                            best_seq = (int)seq; h.t = t;
                            if ((int)seq < ns) { h.type = 1; h.obj = (int)seq; } else { h.type = 2; h.obj = (int)seq - ns; hu = u; hv = v; }
                        }
                    }
                }
                if (sp == 0) break;
                node = stack.get(--sp);
            }
#endif
            if (collect) {
                st.cm0 = m0; st.cm1 = m1; st.cm2 = m2; st.cm3 = m3; done = true;
                // Triangles a shadow sample of this light can hit: ONE walk of each mesh's culling hierarchy with the cone
                // (the boxes bound every well-conditioned triangle and are already padded for the reference's test), plus
                // the few ill-conditioned triangles that are tested for every ray. A candidate is dropped when its plane
                // shows it cannot be hit: Triangle::getIntersection needs d.n < 0 and t = (D - o.n)/(d.n) >= 0, i.e. the
                // origin on the front side; a plane with P behind it, or seen from behind by every cone direction, is out
                // (same margins as lc_cannot_occlude; a NaN plane fails both tests and stays).
                int cn = 0;
                const float plen = length(st.P);
                for (int i = 0; with_meshes && i < s.n_meshes && cn >= 0; ++i) {   // with_meshes = false: the caller knows the scene has none
                    const DMesh &m = s.meshes[i];
                    if (STATS) cnt->mesh++;
                    int msp = 0;
                    int mnode = m.bvh_root >= 0 ? m.bvh_root : 0x7FFFFFFF;
                    uint32_t k = m.always_first, kend = m.always_first + m.always_count;
                    for (;;) {
                        for (; k < kend && cn >= 0; ++k) {
                            const uint32_t r0 = RT_LDG(s.bvh_tris + k);
                            const float4 pl = RT_LDG(s.tri_plane + r0);
                            const V3 n = v3(pl.x, pl.y, pl.z);
                            const float nl = length(n);
                            const bool faces_away = dot(coneD, n) - cone_delta * nl > 1e-4f * nl * (cone_len + cone_delta) + 1e-12f;
                            const bool behind = dot(st.P, n) - pl.w < -(1e-4f * (plen * nl + fabsf(pl.w)) + 2e-5f * nl + 1e-12f);
                            if (faces_away || behind) continue;
#if RT_OPT_LC_HULL
                            // An always-tested triangle accepts points only inside two slabs around c0 (always_bound_of), and only ON ITS PLANE:
                            // a sample ray (from P towards P + D + delta u, |u| <= 1) crosses the plane at s = h / -(D + delta u) . n of the way,
                            // h = P . n - w >= 0, so between s_lo = h / (-D.n + delta |n|) and s_hi = h / (-D.n - delta |n|) when the whole cone
                            // faces the plane (else anywhere in [0, 1]); the reference's own t carries a relative error of a few eps and an
                            // absolute one of <= 4 eps (|w| + |o| |n|) / |d . n| (s units: / (-D.n - delta |n|)), taken 4x, and 1e-3 relative. The
                            // crossing points therefore lie in the capsule around the segment P + [s_lo, s_hi] D with radius s_hi delta. The
                            // functional q . g is linear: over the capsule it ranges between its values at the two ends -+ the radius, and
                            // |q| <= Rmax there. The triangle is dropped when the capsule lies on one side of either slab, when the plane lies
                            // beyond the sample points, or when its stored denominator is 0 / not finite (u1, u2 are then never in [0, 1]).
                            // These few triangles were the ONLY candidates of 53 % of config 5's (hit, light) pairs — all sampled for nothing
                            // (profiles/r02_notes.md, r03y).
                            if (k < m.always_first + m.always_count && s.always_bound) {
                                const float4 b0 = RT_LDG(s.always_bound + 3 * k), b1 = RT_LDG(s.always_bound + 3 * k + 1), b2 = RT_LDG(s.always_bound + 3 * k + 2);
                                if (b2.z != 0.f) continue;
                                const float4 tc0 = RT_LDG(s.tri_edge + 3 * r0);
                                const V3 c0v = v3(tc0.x, tc0.y, tc0.z);
                                const float hPn = fmaxf(dot(st.P, n) - pl.w, 0.f), an = -dot(coneD, n), dn_ = cone_delta * nl;
                                float s_lo = 0.f, s_hi = 1.0001f;
                                if (an > 2.f * dn_ + 1e-4f * nl * (cone_len + cone_delta) + 1e-12f) {
                                    const float ds = 16.f * 5.96e-8f * (fabsf(pl.w) + plen * nl) / (an - dn_);
                                    s_lo = fmaxf(hPn / (an + dn_) * 0.999f - ds, 0.f);
                                    s_hi = fminf(hPn / (an - dn_) * 1.001f + ds, 1.0001f);
                                    if (s_lo > 1.0001f) continue;   // the plane lies beyond every sample point
                                }
                                const V3 qP = st.P - c0v;
                                const V3 qA = qP + s_lo * coneD, qB = qP + s_hi * coneD;
                                const float rad = s_hi * cone_delta;
                                const float Rmax = fmaxf(length(qA), length(qB)) + rad;
                                const float absm = 1e-4f * (Rmax + length(c0v) + plen) + 1e-6f;   // rounding of the hit point and of q
                                const V3 g1v = v3(b0.x, b0.y, b0.z), g2v = v3(b1.x, b1.y, b1.z);
                                const float hA1 = dot(qA, g1v), hB1 = dot(qB, g1v), lim1 = b0.w + b2.x * Rmax * 1.001f + absm + rad;
                                const float hA2 = dot(qA, g2v), hB2 = dot(qB, g2v), lim2 = b1.w + b2.y * Rmax * 1.001f + absm + rad;
                                const bool out1 = b0.w < 1e30f && ((hA1 > lim1 && hB1 > lim1) || (hA1 < -lim1 && hB1 < -lim1));
                                const bool out2 = b1.w < 1e30f && ((hA2 > lim2 && hB2 > lim2) || (hA2 < -lim2 && hB2 < -lim2));
                                if (out1 || out2) continue;
                            }
#endif
                            if (cn >= RT_LC_MAXC || i >= 32 || r0 >= (1u << 27)) { cn = -1; break; }
                            st.cl[cn++] = ((uint32_t)i << 27) | r0;
                        }
                        if (cn < 0) break;
                        // next leaf of this mesh's hierarchy the cone touches
                        while (mnode >= 0 && mnode != 0x7FFFFFFF) {
                            const float4 n0 = RT_LDG(s.bvh_nodes + 4 * mnode), n1 = RT_LDG(s.bvh_nodes + 4 * mnode + 1),
                                         n2 = RT_LDG(s.bvh_nodes + 4 * mnode + 2), n3 = RT_LDG(s.bvh_nodes + 4 * mnode + 3);
                            if (STATS) cnt->node++;
                            float d0, d1;
                            const bool h0 = cone_box(cone, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, limit_c, d0);
                            const bool h1 = cone_box(cone, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, limit_c, d1);
                            const int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y);
                            if (h0 && h1) { mnode = c0; stack.put(msp++, c1); }
                            else if (h0) mnode = c0;
                            else if (h1) mnode = c1;
                            else mnode = msp > 0 ? stack.get(--msp) : 0x7FFFFFFF;
                        }
                        if (mnode == 0x7FFFFFFF) break;
                        const uint32_t code = (uint32_t)(-(mnode + 1));
                        k = code >> 3; kend = k + (code & 7u);
                        mnode = msp > 0 ? stack.get(--msp) : 0x7FFFFFFF;
                    }
                }
                st.cl_n = cn;
            }
        }
    } else if (mine) {
#if RT_OPT_MASKLOOP
        // Scene::computeShadow over the candidates, ascending sequence index = the reference's order
        // (spheres by index, then squares by index): one draw per candidate hit until one blocks.
        // All lanes take their k-th candidate together, spheres in a first loop and squares in a second
        // (squares follow every sphere in the sequence), so both bodies stay convergent.
        if (STATS) cnt->shadow++;
        const SphereRay sr = make_sphere_ray(ray);
        uint32_t m0 = st.cm0, m1 = st.cm1, m2 = st.cm2, m3 = st.cm3;
        for (;;) {
            if (done) break;
            const int base = m0 ? 0 : (m1 ? 32 : (m2 ? 64 : 96));
            const uint32_t m = m0 ? m0 : (m1 ? m1 : (m2 ? m2 : m3));
            if (m == 0u) break;
            const int seq = base + RT_FFS((int)m) - 1;
            if (seq >= ns) break;
            const uint32_t rest = m & (m - 1u);
            if (base == 0) m0 = rest; else if (base == 32) m1 = rest; else if (base == 64) m2 = rest; else m3 = rest;
            if (STATS) cnt->sphere++;
            const float4 b = RT_LDG(s.sph_b + seq);
            const float t = sphere_t(ray, sr, RT_LDG(s.sph_a + seq), b);
            if (t < h.t && t > RT_EPSF) {
                if (STATS) cnt->rnd++;
                if (st.rng.next() > b.w) { blocked = true; done = true; }
            }
        }
        for (;;) {
            if (done) break;
            const int base = m0 ? 0 : (m1 ? 32 : (m2 ? 64 : 96));
            const uint32_t m = m0 ? m0 : (m1 ? m1 : (m2 ? m2 : m3));
            if (m == 0u) break;
            const int seq = base + RT_FFS((int)m) - 1;
            const uint32_t rest = m & (m - 1u);
            if (base == 0) m0 = rest; else if (base == 32) m1 = rest; else if (base == 64) m2 = rest; else m3 = rest;
            if (STATS) cnt->square++;
            float u, v;
            const float t = square_t(ray, s.squares[seq - ns], u, v);
            if (t < h.t && t > RT_EPSF) {
                if (STATS) cnt->rnd++;
                if (st.rng.next() > RT_LDG(s.sq_transparency + (seq - ns))) { blocked = true; done = true; }
            }
        }
    }
#else
        if (lc_shadow_analytic<STATS>(s, st, cnt)) { blocked = true; done = true; }
    }
#endif
    if (!with_meshes || COLLECT) return;   // mode 3 never walks the meshes here (done is set above)
    // meshes. Closest-hit rays, and shadow samples whose light has too many candidate triangles for the list, walk the
    // exact culling hierarchies (variant 3); other shadow samples test the listed candidates, mesh after mesh in the
    // reference's order, with the same per-triangle routine (bvh_consider) the walk uses.
    if (mode == 1 && !run_t && (SAMPLE && SAMPLE_ == 2 ? true : (SAMPLE && SAMPLE_ == 3 ? false : st.cl_n >= 0))) {
        int k = 0;
        while (k < st.cl_n && !done) {
            const uint32_t mi = st.cl[k] >> 27;
            if (STATS) cnt->mesh++;
            float best_t = h.t;
            uint32_t best_ref = 0xFFFFFFFFu;
            for (; k < st.cl_n && (st.cl[k] >> 27) == mi; ++k)
                bvh_consider<STATS>(ray, s, st.cl[k] & 0x07FFFFFFu, best_t, best_ref, cnt);
            if (best_ref != 0xFFFFFFFFu && best_t < h.t && best_t > RT_EPSF) {
                if (STATS) cnt->rnd++;
                if (st.rng.next() > RT_LDG(s.mesh_transparency + mi)) { blocked = true; done = true; }
            }
        }
        return;
    }
#if RT_OPT_MESH_MERGED
    meshes_walk_merged<STATS, SAMPLE_ == 3 && RT_OPT_LC_ANYHIT>(s, ray, mode, st.rng, h, blocked, done, cnt);   // the overflow queue's samples: any-hit
#else
    for (int i = 0; i < s.n_meshes; ++i) {
        if (done) break;
        if (STATS) cnt->mesh++;
        float t; uint32_t ref;
        if (mesh_closest_bvh<STATS>(ray, s, s.meshes[i], h.t, t, ref, cnt, mode != 0) && t < h.t && t > RT_EPSF) {
            if (mode == 0) { h.type = 3; h.obj = i; h.t = t; h.ref = ref; }
            else { if (STATS) cnt->rnd++; if (st.rng.next() > RT_LDG(s.mesh_transparency + i)) { blocked = true; done = true; } }
        }
    }
#endif
}

RT_HD void path_begin(PathState &st, const Ray &primary, const Rng &rng, uint32_t path, int max_bounces) {
    st.rng = rng; st.ray = primary; st.mode = 0; st.N = max_bounces; st.depth = 0; st.path = path; st.max_bounces = max_bounces;
}

// fold the records back to front: result_k = (color_k + result_{k+1} (*) kd_k) + e_k  (Scene.h:339-341)
// The wavefront's radiance records: the record of depth d of path slot `path` is three adjacent float4 {kd, flags} {colour}
// {e}; colour and e are stored (and read) only when they are not +0 (flags bit 0 / 1), which they mostly are: no light
// reached / not an emitter. "+0" is tested on the bits: a stored -0 must come back as -0 (x + -0 and x + +0 differ for x = -0).
RT_HD void wf_rec_store(float4 *wf_rec, unsigned long long stride, uint32_t path, int depth, V3 color, V3 kd, V3 e) {
    float4 *rec = wf_rec + 3ull * ((unsigned long long)depth * stride + path);
    const uint32_t fl = ((f2u(color.x) | f2u(color.y) | f2u(color.z)) ? 1u : 0u) | ((f2u(e.x) | f2u(e.y) | f2u(e.z)) ? 2u : 0u);
    RT_ST_STREAM(rec, make_float4(kd.x, kd.y, kd.z, u2f(fl)));
    if (fl & 1u) RT_ST_STREAM(rec + 1, make_float4(color.x, color.y, color.z, 0.f));
    if (fl & 2u) RT_ST_STREAM(rec + 2, make_float4(e.x, e.y, e.z, 0.f));
}
// fold the records back to front: result_k = (color_k + result_{k+1} (*) kd_k) + e_k  (Scene.h:339-341), then / MAXBOUNCES (:345-350)
RT_HD V3 wf_fold(const float4 *wf_rec, unsigned long long stride, uint32_t path, int depth, int max_bounces, V3 tail) {
    V3 r = tail;
    for (int k = depth - 1; k >= 0; --k) {
        const float4 *rec = wf_rec + 3ull * ((unsigned long long)k * stride + path);
        const float4 kd = RT_LD_STREAM(rec);
        const uint32_t fl = f2u(kd.w);
        V3 c = v3(0.f), e = v3(0.f);
        if (fl & 1u) { const float4 t = RT_LD_STREAM(rec + 1); c = v3(t.x, t.y, t.z); }
        if (fl & 2u) { const float4 t = RT_LD_STREAM(rec + 2); e = v3(t.x, t.y, t.z); }
        r = (c + comp_product(r, v3(kd.x, kd.y, kd.z))) + e;
    }
    r = v3(0.f) + r;
    return r / (float)max_bounces;
}
template <bool WF = false>
RT_HD V3 path_fold(const PathState &st, V3 tail) {
    if (WF) return wf_fold(st.wf_rec, st.wf_stride, st.path, st.depth, st.max_bounces, tail);
    V3 r = tail;
    for (int k = st.depth - 1; k >= 0; --k) r = (st.recs->rec_c[k] + comp_product(r, st.recs->rec_kd[k])) + st.recs->rec_e[k];
    r = v3(0.f) + r;
    return r / (float)st.max_bounces;
}

// next soft-shadow sample of light st.light (Scene.h:325-330)
// The wavefront's light kernels (WF) INLINE this and the bounce step below; the state-machine kernels call shared
// out-of-line copies. Measured (profiles/r02_notes.md, r02n): once nothing else pins the path state in local memory (ray by
// value, candidate list out of the struct), inlining lets the wavefront keep it in registers / spill slots: config 2
// 10.8 -> 10.5 ms, config 5 63.3 -> 62.4; the state-machine kernel of config 3, whose hot code is twice as long, loses 1.4 %
// the same way and keeps the calls. (Round 1 measured the inlined sample neutral to slower: the state was pinned then.)
template <bool STATS, bool INL = false>
RT_HD void path_shadow_sample_body(const DScene &s, PathState &st, Counters *cnt) {
    if (STATS) cnt->rnd += 3;
    const V3 lp = ld3(s.lights[st.light].pos);
    const float delta = s.lights[st.light].radius / 2.f;
    const V3 lj = lp + (INL ? random_unit_vector_inl(st.rng) : random_unit_vector(st.rng)) * delta;
    const V3 Lj = INL ? normalized_inl(lj - st.P) : normalized(lj - st.P);
    st.t_light = length(lj - st.P);
    const float time = st.ray.time;
    st.ray = INL ? make_ray_inl(st.P + Lj * RT_EPSF, Lj, time) : make_ray(st.P + Lj * RT_EPSF, Lj, time);
    st.mode = 1;
}
template <bool STATS>
RT_COLD void path_shadow_sample_shared(const DScene &s_, PathState &st, Counters *cnt) { path_shadow_sample_body<STATS>(RT_S(s_), st, cnt); }
template <bool STATS, bool WF = false>
RT_HD void path_shadow_sample(const DScene &s, PathState &st, Counters *cnt) {
    if (WF) path_shadow_sample_body<STATS>(s, st, cnt);
    else path_shadow_sample_shared<STATS>(s, st, cnt);
}

// Start lighting with light st.light, or — when the lights are exhausted — scatter and continue.
// Returns true when the path has ended (result in `out`).
template <bool STATS, bool LC, bool WF>
RT_HD bool path_next_light_or_bounce_body(const DScene &s, PathState &st, int nb_ech, V3 &out, Counters *cnt) {
    if (st.light < s.n_lights) {
        const V3 L = normalized(ld3(s.lights[st.light].pos) - st.P);
        const float dotLN = dot(L, st.n);
        st.color = st.color + (comp_product(ld3(s.lights[0].color), st.kd) * fmaxr(0.0f, dotLN)) * (float)(1. - (double)st.mat->transparency);
        st.j = 0; st.blocked = 0;
        if (LC) { st.mode = 3; return false; }   // variant 5: first collect the occluder candidates of this light
        path_shadow_sample<STATS, WF>(s, st, cnt);
        return false;
    }
    Ray in; in.o = st.P; in.d = st.in_d; in.time = st.ray.time;
    Rng rng = st.rng;   // a copy: material_scatter is out of line, and a reference into `st` would pin the whole state in local memory
    st.ray = material_scatter<STATS>(*st.mat, in, st.n, st.P, rng, cnt);
    st.rng = rng;
    if (WF) {
        wf_rec_store(st.wf_rec, st.wf_stride, st.path, st.depth, st.color, st.kd, st.e);
    } else {
        st.recs->rec_c[st.depth] = st.color; st.recs->rec_kd[st.depth] = st.kd; st.recs->rec_e[st.depth] = st.e;
    }
    ++st.depth;
    --st.N;
    if (st.N == 0) { out = path_fold<WF>(st, v3(0.f)); st.mode = 2; return true; }
    st.mode = 0;
    return false;
}
template <bool STATS, bool LC>
RT_SHARED_BOUNCE bool path_next_light_or_bounce_shared(const DScene &s_, PathState &st, int nb_ech, V3 &out, Counters *cnt) {
    return path_next_light_or_bounce_body<STATS, LC, false>(RT_S(s_), st, nb_ech, out, cnt);
}
template <bool STATS, bool LC = false, bool WF = false>
RT_HD bool path_next_light_or_bounce(const DScene &s, PathState &st, int nb_ech, V3 &out, Counters *cnt) {
    if (WF) return path_next_light_or_bounce_body<STATS, LC, true>(s, st, nb_ech, out, cnt);
    return path_next_light_or_bounce_shared<STATS, LC>(s, st, nb_ech, out, cnt);
}

// Shade the closest hit of st.ray (Scene.h:270-304): on a miss the path ends (true, result in `out`); otherwise the
// hit context (P, n, kd, e, mat, in_d) is set up for lighting.
// NOMESH: the scene has no meshes (the triangle branch is not compiled in).
template <bool STATS, bool WF = false, bool NOMESH = false>
RT_HD bool path_shade(const DScene &s, PathState &st, const Hit &h, float hu, float hv, V3 &out, Counters *cnt) {
    const Ray &ray = st.ray;
    if (h.type == 0) { out = path_fold<WF>(st, sky_color<STATS>(s, ray.d, st.N, cnt)); st.mode = 2; return true; }
    V3 P, n, kd, e;
    const DMaterial *mat;
    if (h.type == 1) {
        mat = s.sph_mat + h.obj;
        const float4 a = RT_LDG(s.sph_a + h.obj), b = RT_LDG(s.sph_b + h.obj);
        const V3 c = v3(a.x, a.y, a.z) + ray.time * v3(b.x, b.y, b.z);
        P = ray.o + h.t * ray.d;
        n = NOMESH ? normalized_inl(P - c) : normalized(P - c);
        kd = ld3(mat->kd);
        float tu = 0.f, tv = 0.f;
        if (mat->texture_type != 0) {
            sphere_uv(n, tu, tv);
            material_texture<STATS>(s, *mat, kd, tu, tv, cnt);
        }
        e = material_emit<STATS>(s, *mat, tu, tv, cnt);
    } else if (NOMESH || h.type == 2) {
        mat = s.sq_mat + h.obj;
        const DSquare &q = s.squares[h.obj];
        P = ray.o + h.t * ray.d;
        n = ld3(q.n);
        kd = ld3(mat->kd);
        material_texture<STATS>(s, *mat, kd, hu, hv, cnt);
        n = material_normal<STATS>(s, *mat, n, hu, hv, ld3(q.tan_r), ld3(q.tan_u), cnt);
        e = material_emit<STATS>(s, *mat, hu, hv, cnt);
    } else {
        mat = s.mesh_mat + h.obj;
        const DMesh &m = s.meshes[h.obj];
        float w0, w1, w2;
        triangle_t<false>(ray, s, h.ref, w0, w1, w2, nullptr);
        P = ray.o + h.t * ray.d;
        const float4 pl = RT_LDG(s.tri_plane + h.ref);
        n = v3(pl.x, pl.y, pl.z);
        kd = ld3(mat->kd);
        const uint32_t ti = f2u(RT_LDG(s.tri_den + h.ref).y);
        if (m.color_type == 0) {
            const uint32_t i0 = m.triangles[3 * ti], i1 = m.triangles[3 * ti + 1], i2 = m.triangles[3 * ti + 2];
            kd = w0 * ld3(m.vert_colors + 3 * i0) + w1 * ld3(m.vert_colors + 3 * i1) + w2 * ld3(m.vert_colors + 3 * i2);
        } else if (m.color_type == 1) {
            kd = ld3(m.face_colors + 3 * ti);
        }
        e = v3(0.f);
    }
    st.P = P; st.n = n; st.kd = kd; st.e = e; st.mat = mat; st.in_d = ray.d;
    st.color = v3(0.f);
    st.light = 0;
    return false;
}

// The NB_ECH samples of light st.light are in st.blocked: scale the colour (Scene.h:331-333) and go on.
template <bool STATS, bool LC = false, bool WF = false>
RT_HD bool path_finish_light(const DScene &s, PathState &st, int nb_ech, V3 &out, Counters *cnt) {
    const float shadow = (float)(1. - (double)((float)st.blocked / (float)nb_ech));
    st.color = st.color * shadow;
    ++st.light;
    return path_next_light_or_bounce<STATS, LC, WF>(s, st, nb_ech, out, cnt);
}

// Consume the result of intersect_ray for this lane's ray and set up the next ray.
template <bool STATS, bool LC = false, bool WF = false>
RT_HD bool path_advance(const DScene &s, PathState &st, const Hit &h, float hu, float hv, bool blocked, int nb_ech, V3 &out,
                        Counters *cnt) {
    if (LC && st.mode == 3) {   // candidates collected
        if (lc_light_unoccluded(st)) {
            if (STATS) { cnt->shadow += nb_ech; cnt->rnd += 3 * nb_ech; }
            st.rng.ctr += 3u * (uint32_t)nb_ech;
            ++st.light;
            return path_next_light_or_bounce<STATS, LC, WF>(s, st, nb_ech, out, cnt);
        }
        if (lc_umbra(s, st, nb_ech)) {   // every sample blocked by construction
            if (STATS) { cnt->shadow += nb_ech; cnt->rnd += 4 * nb_ech; }
            st.blocked = nb_ech;
            return path_finish_light<STATS, LC, WF>(s, st, nb_ech, out, cnt);
        }
        path_shadow_sample<STATS, WF>(s, st, cnt);   // first sample
        return false;
    }
    if (st.mode == 1) {
        if (blocked) ++st.blocked;
        if (++st.j < nb_ech) { path_shadow_sample<STATS, WF>(s, st, cnt); return false; }
        return path_finish_light<STATS, LC, WF>(s, st, nb_ech, out, cnt);
    }
    // mode 0
    if (path_shade<STATS, WF>(s, st, h, hu, hv, out, cnt)) return true;
    return path_next_light_or_bounce<STATS, LC, WF>(s, st, nb_ech, out, cnt);
}

// ---- camera ------------------------------------------------------------------------------------
// MatrixUtilities::screen_space_to_world_space_ray (matrixUtilities.h:53-74): unprojection in fp64
// (sums left to right), dehomogenised in fp64, cast to float; direction normalised there and
// again by the Ray constructor. cam_pos is cameraSpaceToWorldSpace(0,0,0), constant per frame.
struct DCamera { double mvi[16], pi[16], depth_near; float pos[3]; };
RT_HD void mat4_mul(const double *m, double x, double y, double z, double w, double *r) {
    r[0] = m[0] * x + m[4] * y + m[8] * z + m[12] * w;
    r[1] = m[1] * x + m[5] * y + m[9] * z + m[13] * w;
    r[2] = m[2] * x + m[6] * y + m[10] * z + m[14] * w;
    r[3] = m[3] * x + m[7] * y + m[11] * z + m[15] * w;
}
RT_HD Ray camera_ray(const DCamera &c, float u, float v, float time) {
    double a[4], b[4];
    mat4_mul(c.pi, 2.0 * (double)u - 1.0, -(2.0 * (double)v - 1.0), c.depth_near, 1.0, a);
    mat4_mul(c.mvi, a[0], a[1], a[2], a[3], b);
    const V3 p = v3((float)(b[0] / b[3]), (float)(b[1] / b[3]), (float)(b[2] / b[3]));
    const V3 pos = ld3(c.pos);
    const V3 dir = normalized(p - pos);
    return make_ray(pos, dir, time);
}
// trace_line's per-sample prologue (main.cpp:189-192): u, v, time are draws 0, 1, 2
RT_HD Ray primary_ray_inline(const DCamera &c, int x, int y, int w, int h, Rng &rng) {
    const float u = ((float)x + rng.next()) / (float)w;
    const float v = ((float)y + rng.next()) / (float)h;
    const float time = rng.next();
    return camera_ray(c, u, v, time);
}
// out of line for the kernels that generate a camera ray now and then (refill of k_render_regen): keeps ~400 fp64-heavy
// instructions out of their hot code. k_camera_rays inlines it: called by reference, the 272-byte camera struct would be
// copied to local memory by every thread (37 STL + 39 LDL per path, measured 10 % of config 2's GPU time).
RT_COLD Ray primary_ray(const DCamera &c, int x, int y, int w, int h, Rng &rng) { return primary_ray_inline(c, x, y, w, h, rng); }

// gamma_correct (Functions.cpp:56-60)
RT_HD float gamma_channel(float c) { return (float)pow((double)c, 1.0 / 2.2); }

}  // namespace rt
#endif
