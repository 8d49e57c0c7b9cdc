// rt_capi.cu — kernels and the extern "C" layer of include/hai719_rt.h (sm_100a only).
//
// Kernel map (reference function each replaces; device functions in rt_core.cuh):
//   k_precompute_squares / k_precompute_tris   per-primitive constants that Square::intersect
//                          (Square.h:68-72) and the Triangle constructor (Triangle.h:26-37, called
//                          per ray per triangle from KDTree.cpp:38-40) recompute for every ray
//   k_render_paths         trace_line's pixel x sample loop (main.cpp:183-193) + Scene::rayTrace:
//                          persistent CTAs, warps fetch batches of 32 paths from an atomic counter
//   k_resolve              image[x+y*w] += color in sample order, /= nsamples, gamma_correct
//                          (main.cpp:193-196)
//   k_primary_ids / k_trace_rays / k_shade_rays   Scene::computeIntersection / rayTrace on given rays
//   k_untile               packed tiles -> row-major image (no reference equivalent)
//
// No CPU fallback: every entry point that needs the GPU fails with RT_ERR_NO_DEVICE without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "hai719_rt.h"
#include "rt_core.cuh"
#include "rt_pack.hpp"
#include "rt_bvh.hpp"

using namespace rt;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
namespace {
thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define RT_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA,                     \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                           \
        }                                                                                              \
    } while (0)
}  // namespace

// ------------------------------------------------------------------------------------------------
// device-side records used by the kernels
// ------------------------------------------------------------------------------------------------
struct TileRec { int x0, y0, w, h; };   // one tile of this rank, in full-image pixel coordinates

struct RenderArgs {
    const TileRec *tiles;
    const unsigned int *tile_off;       // n_tiles + 1 prefix sums of tile pixel counts (packed order)
    int n_tiles;
    int width, height, spp, max_bounces, nb_ech;
    unsigned int seed;
    unsigned long long pixel_begin;     // first packed pixel of this chunk
    unsigned long long n_paths;         // paths in this chunk = chunk pixels * spp
    float *samples;                     // n_paths * 3, path-major
    unsigned long long *work_counter;
    unsigned long long *stats;          // 10 counters, or null
    int regen_min;                      // variant 1: refill idle lanes once at least this many are idle
    int t_min;                          // variant 5: run a traversal step once at least this many lanes wait for one
    const float4 *cam_rays;             // k_camera_rays output per path {dir.xyz, time}, or null: generate in the render kernel
    const unsigned int *cam_keys;       // ... and the key of the path's random stream (3 draws already taken)
    unsigned int sample_base;           // index of this call's first sample of every pixel (progressive accumulation, else 0)
};

// packed pixel index -> (x, y) via the tile prefix array
__device__ __forceinline__ void packed_to_xy(const RenderArgs &a, unsigned int lp, int &x, int &y) {
    int lo = 0, hi = a.n_tiles - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(a.tile_off + mid) <= lp) lo = mid; else hi = mid - 1;
    }
    const TileRec t = a.tiles[lo];
    const unsigned int r = lp - __ldg(a.tile_off + lo);
    x = t.x0 + (int)(r % (unsigned int)t.w);
    y = t.y0 + (int)(r / (unsigned int)t.w);
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
struct SquareIn { float v0[3], v1[3], v3[3], right[3], up[3], motion[3]; int glass; };

__global__ void k_precompute_squares(const SquareIn *in, DSquare *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SquareIn s = in[i];
    const V3 v0 = ld3(s.v0);
    const V3 right = ld3(s.v1) - v0, up = ld3(s.v3) - v0;
    const V3 nrm = normalized(cross(right, up));
    DSquare q;
    q.v0[0] = v0.x; q.v0[1] = v0.y; q.v0[2] = v0.z;
    q.n[0] = nrm.x; q.n[1] = nrm.y; q.n[2] = nrm.z;
    q.right[0] = right.x; q.right[1] = right.y; q.right[2] = right.z;
    q.up[0] = up.x; q.up[1] = up.y; q.up[2] = up.z;
    q.len_r = length(right);
    q.len_u = length(up);
    for (int k = 0; k < 3; ++k) { q.motion[k] = s.motion[k]; q.tan_r[k] = s.right[k]; q.tan_u[k] = s.up[k]; }
    q.glass = s.glass;
    out[i] = q;
}

// always_bound_of for the always-tested entries [first, first + count) of bvh_tris (after k_precompute_tris of that mesh, same stream)
__global__ void k_always_bounds(const uint32_t *tris, uint32_t first, uint32_t count, const float4 *edge, const float2 *den, float4 *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t k = first + i, r0 = tris[k];
    const AlwaysBound ab = always_bound_of(edge[3 * r0], edge[3 * r0 + 1], edge[3 * r0 + 2], den[r0].x);
    out[3 * k] = ab.g1; out[3 * k + 1] = ab.g2; out[3 * k + 2] = ab.b;
}
__global__ void k_precompute_tris(const float *positions, const RtTriRef *refs, unsigned int n_refs, float4 *plane,
                                  float4 *edge, float2 *den) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_refs) return;
    const RtTriRef r = refs[i];
    const TriConst k = precompute_triangle(ld3(positions + 3 * r.v[0]), ld3(positions + 3 * r.v[1]),
                                           ld3(positions + 3 * r.v[2]), r.tri_index);
    plane[i] = k.plane;
    edge[3 * i] = k.c0;
    edge[3 * i + 1] = k.e0;
    edge[3 * i + 2] = k.e1;
    den[i] = k.den;
}

__device__ __forceinline__ void flush_counters(const Counters &c, unsigned long long *g) {
    const unsigned long long v[10] = {c.closest, c.shadow, c.sphere, c.square, c.mesh, c.node, c.tri, c.tri_full, c.tex, c.rnd};
    for (int k = 0; k < 10; ++k) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(g + k, x);
    }
}

// One path per lane; warps pull batches of 32 consecutive paths (same pixel for spp >= 32, so a
// warp's primary rays are coherent) from a global counter until the chunk is exhausted.
#ifndef RT_PATHS_MINB
#define RT_PATHS_MINB 4
#endif
template <bool STATS>
__global__ void __launch_bounds__(128, RT_PATHS_MINB) k_render_paths(const DScene scene_, const DCamera cam, const RenderArgs a) {
    const DScene &scene = RT_S(scene_);
    Counters cnt;
    if (STATS) memset(&cnt, 0, sizeof cnt);
    const unsigned int lane = threadIdx.x & 31u;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.work_counter, 32ull);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= a.n_paths) break;
        const unsigned long long p = base + lane;
        if (p < a.n_paths) {
            const unsigned int lp = (unsigned int)(a.pixel_begin + p / (unsigned int)a.spp);
            const unsigned int smp = (unsigned int)(p % (unsigned int)a.spp);
            int x, y;
            packed_to_xy(a, lp, x, y);
            Rng rng;
            rng.init(a.seed, (unsigned int)x + (unsigned int)y * (unsigned int)a.width, a.sample_base + smp);
            if (STATS) cnt.rnd += 3;
            const Ray ray = primary_ray(cam, x, y, a.width, a.height, rng);
            const V3 c = trace_path<STATS>(scene, ray, rng, a.max_bounces, a.nb_ech, &cnt);
            float *o = a.samples + 3ull * p;
            o[0] = c.x; o[1] = c.y; o[2] = c.z;
        }
    }
    if (STATS) flush_counters(cnt, a.stats);
}

// Variant 1: ray-level state machine with per-lane path regeneration. Every loop iteration each
// lane intersects ONE ray (closest-hit or shadow sample) and advances its path; a lane whose path
// has ended takes the next path index from the global counter (one warp-aggregated atomicAdd,
// __ballot_sync/__popc ranks). Lanes stay busy whatever the depth at which their paths end.
// Wavefront stage 1 — camera-ray generation (trace_line's per-sample prologue, main.cpp:189-192, and
// screen_space_to_world_space_ray, matrixUtilities.h:53-74) for every path of a chunk: one thread per path,
// fully convergent, instead of a few refilling lanes of the render kernel running ~600 instructions of fp64
// unprojection each (and keeping them in its instruction-cache footprint). 20 B per path.
// (x, y) of every packed pixel of the chunk, once per pixel: the tile lookup is a binary search of ~11 dependent loads,
// and k_camera_rays used to repeat it for each of the pixel's spp paths (camera rays were 10 % of config 2's GPU time)
__global__ void __launch_bounds__(256) k_pixel_xy(const RenderArgs a, unsigned int n_pixels, unsigned int *xy) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    int x, y;
    packed_to_xy(a, (unsigned int)a.pixel_begin + i, x, y);
    xy[i] = (unsigned int)x | ((unsigned int)y << 16);
}

__global__ void __launch_bounds__(256) k_camera_rays(const DCamera cam, const RenderArgs a, const unsigned int *xy, float4 *rays, unsigned int *keys) {
    const unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_paths) return;
    const unsigned int smp = (unsigned int)(p % (unsigned int)a.spp);
    const unsigned int packed = __ldg(xy + p / (unsigned int)a.spp);
    const int x = (int)(packed & 0xFFFFu), y = (int)(packed >> 16);
    Rng rng;
    rng.init(a.seed, (unsigned int)x + (unsigned int)y * (unsigned int)a.width, a.sample_base + smp);
    const Ray ray = primary_ray_inline(cam, x, y, a.width, a.height, rng);
    rays[p] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.time);
    keys[p] = rng.key;
}

#ifndef RT_OPT_REGEN_MERGE
#define RT_OPT_REGEN_MERGE 1
#endif
#ifndef RT_OPT_GRIDCONST
#define RT_OPT_GRIDCONST 0   /* measured slower on every config (profiles/r01_notes.md) */
#endif
#if RT_OPT_GRIDCONST
// out-of-line device functions take the scene by reference: __grid_constant__ lets them read it from the
// parameter bank instead of a per-thread local-memory copy (264 STL.128 at kernel entry, LDL on every use)
#define RT_PARAM __grid_constant__
#else
#define RT_PARAM
#endif
template <bool STATS, int ACCEL, int MINB>
__global__ void __launch_bounds__(128, MINB) k_render_regen(const RT_PARAM DScene scene_, const RT_PARAM DCamera cam, const RenderArgs a) {
    const DScene &scene = RT_S(scene_);
    stage_abvh(scene);
    Counters cnt;
    if (STATS) memset(&cnt, 0, sizeof cnt);
    const unsigned int lane = threadIdx.x & 31u;
    PathState st;
    PathRecs recs;
    CandList cands;
    st.recs = &recs; st.cl = cands.v; st.wf_rec = nullptr; st.wf_stride = 0;
    st.mode = 2;
    st.ray.o = v3(0.f); st.ray.d = v3(0.f, 0.f, 1.f); st.ray.time = 0.f; st.t_light = 0.f;
    st.rng.key = 0; st.rng.ctr = 0;
    bool exhausted = false;
    // idle lanes (mask `need`, all of them) take the next paths of the chunk: one warp-aggregated atomicAdd
    auto refill = [&](unsigned int need) {
        const int leader = __ffs(need) - 1;
        const unsigned long long n = (unsigned long long)__popc(need);
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(a.work_counter, n);
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (base + n >= a.n_paths) exhausted = true;
        if (st.mode == 2) {
            const unsigned long long p = base + (unsigned long long)__popc(need & ((1u << lane) - 1u));
            if (p < a.n_paths) {
                Rng rng;
                Ray ray;
                if (STATS) cnt.rnd += 3;
                if (a.cam_rays) {
                    const float4 r = __ldg(a.cam_rays + p);
                    ray.o = ld3(cam.pos); ray.d = v3(r.x, r.y, r.z); ray.time = r.w;
                    rng.key = __ldg(a.cam_keys + p); rng.ctr = 3u;
                } else {
                    const unsigned int lp = (unsigned int)(a.pixel_begin + p / (unsigned int)a.spp);
                    const unsigned int smp = (unsigned int)(p % (unsigned int)a.spp);
                    int x, y;
                    packed_to_xy(a, lp, x, y);
                    rng.init(a.seed, (unsigned int)x + (unsigned int)y * (unsigned int)a.width, a.sample_base + smp);
                    ray = primary_ray(cam, x, y, a.width, a.height, rng);
                }
                path_begin(st, ray, rng, (uint32_t)p, a.max_bounces);
                if (a.max_bounces == 0) {   // rayTraceRecursive(ray, 0) / 0
                    const V3 c = path_fold(st, v3(0.f));
                    float *o = a.samples + 3ull * p;
                    o[0] = c.x; o[1] = c.y; o[2] = c.z;
                    st.mode = 2;
                }
            }
        }
    };
    for (;;) {
        if (ACCEL == 3) {
            // variant 5: a lane wants either a TRAVERSAL step (mode 0 closest hit, mode 3 occluder candidates of a
            // light; an idle lane that can still get a path counts as wanting one) or a SHADOW-SAMPLE step (mode 1:
            // exact tests on its candidate mask). One kind per iteration for the whole warp: traversal once enough
            // lanes wait for it (or nobody wants a sample), samples otherwise. Idle lanes refill right before a
            // traversal step, so new paths join the cohort that is about to traverse.
#if RT_OPT_REGEN_MERGE
            const unsigned int bi = exhausted ? 0u : __ballot_sync(0xFFFFFFFFu, st.mode == 2);
#else
            if (!exhausted) {
                const unsigned int need = __ballot_sync(0xFFFFFFFFu, st.mode == 2);
                if (__popc(need) >= a.regen_min || need == 0xFFFFFFFFu) refill(need);
            }
            const unsigned int bi = 0u;
#endif
            const unsigned int bt = __ballot_sync(0xFFFFFFFFu, st.mode == 0 || st.mode == 3), bs = __ballot_sync(0xFFFFFFFFu, st.mode == 1);
            if ((bi | bt | bs) == 0u) break;
            const bool run_t = bs == 0u || __popc(bt | bi) >= a.t_min;
            if (run_t && bi) refill(bi);
            const bool mine = run_t ? (st.mode == 0 || st.mode == 3) : (st.mode == 1);
            Hit h;
            float hu = 0.f, hv = 0.f;
            bool blocked;
            intersect_lc<STATS>(scene, st, run_t, mine, h, hu, hv, blocked, &cnt);
            if (mine) {
                V3 c;
                if (path_advance<STATS, true>(scene, st, h, hu, hv, blocked, a.nb_ech, c, &cnt)) {
                    float *o = a.samples + 3ull * st.path;
                    o[0] = c.x; o[1] = c.y; o[2] = c.z;
                }
            }
            continue;
        }
        if (!exhausted) {
            const unsigned int need = __ballot_sync(0xFFFFFFFFu, st.mode == 2);
            if (__popc(need) >= a.regen_min || need == 0xFFFFFFFFu) refill(need);
        }
        if (__all_sync(0xFFFFFFFFu, st.mode == 2)) break;
        Hit h;
        float hu = 0.f, hv = 0.f;
        bool blocked;
        if (ACCEL == 2) intersect_ray_voted<STATS>(scene, st.ray, st.mode, st.t_light, st.rng, h, hu, hv, blocked, &cnt);
        else intersect_ray<STATS, ACCEL == 1>(scene, st.ray, st.mode, st.t_light, st.rng, h, hu, hv, blocked, &cnt);
        if (st.mode != 2) {
            V3 c;
            if (path_advance<STATS>(scene, st, h, hu, hv, blocked, a.nb_ech, c, &cnt)) {
                float *o = a.samples + 3ull * st.path;
                o[0] = c.x; o[1] = c.y; o[2] = c.z;
            }
        }
    }
    if (STATS) flush_counters(cnt, a.stats);
}

// ------------------------------------------------------------------------------------------------
// Variant 6 — wavefront. The same per-ray device functions as the state-machine kernel, cut into
// three kernels per bounce level so that every warp resident on an SM runs the same small piece of
// code (the single kernel waited on instruction fetch more than on anything else: its hot code is
// ~2 000 instructions against a 32 KB L1.5 instruction cache, profiles/r01_notes.md):
//   k_camera_rays  primary rays of the chunk (above)
//   k_wf_trace     closest hit (Scene::computeIntersection) + shading of the hit (Scene.h:270-304) for
//                  every live path; paths that miss fold their sky colour and finish. In a scene
//                  without lights the light stage is only Material::scatter, so it runs right here
//                  (NOLIGHT) and the hit never travels through HBM.
//   k_wf_light     direct lighting of the hit — per light one cone walk for the occluder candidates,
//                  then the NB_ECH shadow samples in registers — then Material::scatter; paths
//                  that used their last bounce fold and finish, the others queue for the next level
// Both are PERSISTENT: warps fetch batches of 32 queue entries with one atomicAdd, and survivors are
// appended to the next queue by warp-aggregated block reservation (__ballot_sync/__popc ranks), so the
// host never needs to know how many paths are alive.
//
// Path state between kernels lives in HBM as float4 SoA records (one 128-bit access per field and lane),
// STORED AT THE QUEUE POSITION of the entry, not at the path's slot: entry i of the live queue has its
// ray in ray0[i] / ray1[i] / rng_ray[i], entry i of the hit queue its hit in hit0..4[i] / rng_hit[i]; the
// queue word itself is the path slot (where the radiance records and the final sample go). A kernel
// therefore reads its input as contiguous 128-byte lines and writes its output the same way. Until round 2
// the records sat at the slot and were gathered through the queue: half-used sectors from bounce 1 on,
// 877 B of DRAM traffic per path against 599 B of records (config 2; the light kernel of the pool scene
// moved 330 B per path and level for 176 B of records, profiles/r02_notes.md).
#define WF_HIT_E 0x08000000u   /* hit record flags in the code word of hit1.w: emission stored / incoming direction stored */
#define WF_HIT_D 0x04000000u
#define WF_NCTR 12  /* counters per bounce level: [0] trace head, [1] live-queue length, [2] light head, [3] hit-queue length, [4] parked
                       head, [5] parked length, [6] overflow head, [7] overflow length, [8] rays traced, [9] hits lit, [10] mesh-walk
                       head, [11] mesh-walk length (lengths count reserved positions, padding included; [8] / [9] count real entries) */
#define WF_INVALID 0xFFFFFFFFu /* queue word of a padding position */
struct WfArgs {
    unsigned int n_paths;              // paths of this chunk (slots 0 .. n_paths-1)
    unsigned int max_grab;             // batches a warp may fetch with one atomic (see wf_next_batch)
    unsigned int block;                // queue positions a warp reserves with one atomic (see wf_push)
    int flat;                          // trace kernels: flat box test over <= 32 analytic primitives instead of the hierarchy walk from bounce flat - 1 on (0: never)
    int sort;                          // trace kernels from bounce 1 on: order every fetched window of the live queue by ray direction (wf_next_batch_sorted)
    int max_bounces, nb_ech, level;
    const float4 *cam_rays; const unsigned int *cam_keys;   // level 0 input, by slot
    // live queue of this level (trace input) ...
    const unsigned int *q_live; const float4 *ray0, *ray1; const uint2 *rng_ray;   // {o.xyz, time}, {d.xyz, bits(N | depth << 8)}, {key, ctr}
    // ... and of the next level (light output; NOLIGHT: trace output)
    unsigned int *q_next; float4 *nray0, *nray1; uint2 *nrng;
    // hit queue of this level (trace output, light input): {P, time} {n, bits(kind << 28 | obj)} {kd, bits(N | depth << 8)} {e, -} {in_d, -}
    unsigned int *q_hit; float4 *hit0, *hit1, *hit2, *hit3, *hit4; uint2 *rng_hit;
    float4 *rec; unsigned long long rec_stride;   // radiance records, by slot
    unsigned int *q_mesh; float4 *mesh_hit;   // trace phase A -> phase B: live-queue positions of the rays that touch a mesh, and their analytic
                                       // hit so far {t, bits(type << 28 | obj), u, v} (by position in q_mesh)
    unsigned int *q_park;              // light phase A -> phase B: hit-queue positions whose light needs its shadow samples traced
    unsigned int *q_over;              // ... those whose candidate-triangle list overflowed: their samples walk the mesh hierarchies,
    int which_park;                    //     so they get warps of their own (phase B runs once per queue: 0 = q_park, 1 = q_over)
    float4 *park0, *park1;             // ... and what phase B needs besides the hit record, by hit-queue position:
    float4 *park2;                     //     {colour so far, light | (cl_n + 1) << 8}, {cm0..cm3}, RT_LC_MAXC / 4 planes of candidate triangles
    unsigned long long park_stride;
    unsigned int *ctr;                 // WF_NCTR counters per level
    float *samples;
    unsigned long long *stats;
};

// Queue traffic. Every batch of 32 entries used to cost two atomicAdds on two counters of one 32-byte sector (work
// fetch + append of the survivors): ~1.1 M same-address atomics per chunk, which L2 retires at roughly one per 1.3 ns —
// the level-0 trace kernel ran exactly that long (0.68 ms for 520 k atomics, profiles/r01_notes.md). Now
//   * a warp FETCHES several batches per atomic, guided self-scheduling: 8 batches while plenty of work is left,
//     shrinking to 1 near the end so that the tail stays balanced (scenes with meshes always fetch 1: measured, their
//     batches are too uneven);
//   * a warp RESERVES output positions a block at a time (256 for large chunks) with one atomicAdd and hands them to its
//     survivors in rank order (__ballot_sync/__popc), so every survivor knows its queue position at once and stores its
//     records there — coalesced, no staging in shared memory. When the kernel ends, the unused tail of a warp's last
//     block is padded with WF_INVALID words: the consumer skips those positions (they come in runs, so whole batches
//     are skipped at the price of one 128-byte read).
// Path state, queues and radiance records are written once by one kernel and read once by the next, hundreds of MB per
// level: streaming (evict-first) accesses keep them from pushing the per-thread stacks and the scene out of L2.
#ifndef RT_WF_STREAM
#define RT_WF_STREAM 1
#endif
#if RT_WF_STREAM
#define WF_LD(p) __ldcs(p)
#define WF_ST(p, v) __stcs((p), (v))
#else
#define WF_LD(p) (*(p))
#define WF_ST(p, v) (*(p) = (v))
#endif
// Per-warp bookkeeping of the persistent loops lives in SHARED memory (a few words per warp, read as broadcasts, written
// by lane 0): work-fetch window, the current block of each output queue, the count of entries processed. In registers
// these eight values stayed live across the whole loop body, and at the 64-register cap of these kernels each of them
// pushed something else of the walk into a spill slot (ptxas: 24 -> 272 bytes of spill stores in k_wf_trace when the
// queue positions were introduced).
#define WQ_CUR 0    /* work fetch: next entry of the window, end of the window */
#define WQ_END 1
#define WQ_N 2      /* entries really processed (padding excluded) */
#define WQ_OUT 3    /* output queue k: WQ_OUT + 2k = next free position, + 1 = end of the block */
#define WQ_BASE 9   /* sorted fetch: first position of the current window, and whether it was sorted */
#define WQ_SORTED 10
#define WQ_WORDS 12
__device__ __forceinline__ unsigned int *wq_init(unsigned int (*all)[WQ_WORDS]) {
    unsigned int *ws = all[threadIdx.x >> 5];
    if ((threadIdx.x & 31u) < WQ_WORDS) ws[threadIdx.x & 31u] = 0u;
    __syncwarp();
    return ws;
}
__device__ __forceinline__ bool wf_next_batch(unsigned int *head, unsigned int count, unsigned int max_grab, unsigned int *ws, unsigned int &base) {
    unsigned int cur = ws[WQ_CUR];
    const unsigned int end = ws[WQ_END];
    __syncwarp();
    if (cur >= end) {   // warp-uniform
        const unsigned int lane = threadIdx.x & 31u;
        const unsigned int left = count > end ? count - end : 0u;                     // end ~ global progress at the last fetch
        const unsigned int per_warp = left / (gridDim.x * (blockDim.x >> 5) * 64u);   // half of an even share, in batches
        const unsigned int grab = 32u * (per_warp < 1u ? 1u : (per_warp > max_grab ? max_grab : per_warp));
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(head, grab);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if (b >= count) return false;
        cur = b;
        if (lane == 0) ws[WQ_END] = b + grab < count ? b + grab : count;
    }
    base = cur;
    if ((threadIdx.x & 31u) == 0u) ws[WQ_CUR] = cur + 32u;
    __syncwarp();
    return true;
}
// Coherence from bounce 1 on. The entries of a live queue are in the order their paths were processed: neighbours are samples of
// the same few pixels, so their rays start close to each other but leave in unrelated directions (a diffuse scatter covers the
// hemisphere), and the 32 walks of a batch part ways at the first box: the trace kernels of config 2 ran 31 of 32 lanes at bounce 0
// and 18 / 15 / 13 at bounces 1 / 2 / 3 (profiles/r02_notes.md). The producer therefore tags every live entry with a 6-bit
// DIRECTION CODE (bits 26..31 of the queue word; path slots are below 2^26): the cell of its ray on an 8 x 8 octahedral map of the
// sphere, cells numbered along a Morton curve, so that close codes are close directions. The consumer orders every window it
// fetches (up to 8 batches = 256 entries with one atomic, wf_next_batch) by that code — a counting sort inside the warp:
// __match_any_sync ranks, 64 counters and a 256-byte permutation per warp in shared memory — and takes its batches in sorted
// order: 32 rays from nearby origins into nearby directions. Records are read at the permuted positions (a gather inside a
// 4 KB window per array: every sector is still used by one of the window's batches). Results do not depend on the order in which
// paths are processed (each path has its own slot, stream and records); only the order of queue entries changes.
// MEASURED (profiles/r02_notes.md, r03d/e; config 2, 16 spp): lanes of the trace kernels 18.0 -> 20.4 / 14.6 -> 15.6 / 12.6 -> 13.8 at bounces
// 1 / 2 / 3, warp instructions -6.5 % at bounce 1, but issue utilisation 82.5 -> 73.6 % behind the gathered record loads: frame 10.4 ->
// 10.6 ms. The walks of a batch differ for reasons a 64-cell direction code over 256 neighbours does not capture. OFF (RT_WF_SORT=0:
// plain slots in the queue words, no shared memory for the sort); make alt DEFS=-DRT_WF_SORT=1 builds it, HAI719_WF_SORT=0/1 then
// switches it per call.
#ifndef RT_WF_SORT
#define RT_WF_SORT 0
#endif
#define WF_SLOT_MASK 0x03FFFFFFu
__device__ __forceinline__ unsigned int wf_dir_code(V3 d) {
    const float s = __fdividef(1.f, fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + 1e-30f);
    float u = d.x * s, v = d.y * s;
    if (d.z < 0.f) { const float u2 = copysignf(1.f - fabsf(v), u), v2 = copysignf(1.f - fabsf(u), v); u = u2; v = v2; }
    int iu = (int)(u * 4.f + 4.f), iv = (int)(v * 4.f + 4.f);
    iu = iu < 0 ? 0 : (iu > 7 ? 7 : iu); iv = iv < 0 ? 0 : (iv > 7 ? 7 : iv);
    const unsigned int a = (unsigned int)iu, b = (unsigned int)iv;
    const unsigned int m = (a & 1u) | ((b & 1u) << 1) | ((a & 2u) << 1) | ((b & 2u) << 2) | ((a & 4u) << 2) | ((b & 4u) << 3);
    return m > 62u ? 62u : m;   // 63 is what a padding word (WF_INVALID) decodes to: sorted to the end of its window
}
__device__ __forceinline__ unsigned int wf_live_word(unsigned int slot, V3 d) { return RT_WF_SORT ? slot | (wf_dir_code(d) << 26) : slot; }
struct WfSortSm { unsigned char perm[256], rank[256]; unsigned short cnt[64]; };   // per warp
// wf_next_batch with sorted windows: `pos` = the queue position this lane takes in this batch.
__device__ __forceinline__ bool wf_next_batch_sorted(unsigned int *head, unsigned int count, unsigned int max_grab, unsigned int *ws, const unsigned int *q,
                                                     WfSortSm *sm, unsigned int &pos) {
    unsigned int cur = ws[WQ_CUR];
    const unsigned int end = ws[WQ_END];
    __syncwarp();
    const unsigned int lane = threadIdx.x & 31u;
    if (cur >= end) {   // warp-uniform
        const unsigned int left = count > end ? count - end : 0u;
        const unsigned int per_warp = left / (gridDim.x * (blockDim.x >> 5) * 64u);
        const unsigned int grab = 32u * (per_warp < 1u ? 1u : (per_warp > max_grab ? max_grab : per_warp));
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(head, grab);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if (b >= count) return false;
        cur = b;
        const unsigned int e = b + grab < count ? b + grab : count;
        const unsigned int n = e - b;
        const bool sorted = n > 32u && n <= 256u;
        if (lane == 0) { ws[WQ_END] = e; ws[WQ_BASE] = b; ws[WQ_SORTED] = sorted ? 1u : 0u; }
        if (sorted) {
            sm->cnt[lane] = 0; sm->cnt[lane + 32u] = 0;
            __syncwarp();
            unsigned long long codes = 0ull;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (32u * j < n) {   // warp-uniform
                    const unsigned int p = b + 32u * j + lane;
                    unsigned int code = 63u;
                    if (p < e) code = WF_LD(q + p) >> 26;
                    const unsigned int peers = __match_any_sync(0xFFFFFFFFu, code);
                    const unsigned int before = peers & ((1u << lane) - 1u);
                    const unsigned int r = sm->cnt[code] + __popc(before);
                    __syncwarp();
                    if (before == 0u) sm->cnt[code] = (unsigned short)(sm->cnt[code] + __popc(peers));
                    __syncwarp();
                    sm->rank[32 * j + lane] = (unsigned char)r;
                    codes |= (unsigned long long)code << (6 * j);
                }
            }
            // exclusive prefix sum over the 64 counters: lane L owns codes 2L and 2L + 1
            const unsigned int c0 = sm->cnt[2u * lane], c1 = sm->cnt[2u * lane + 1u];
            unsigned int incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += t; }
            const unsigned int excl = incl - (c0 + c1);
            __syncwarp();
            sm->cnt[2u * lane] = (unsigned short)excl; sm->cnt[2u * lane + 1u] = (unsigned short)(excl + c0);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (32u * j < n) {
                    const unsigned int code = (unsigned int)(codes >> (6 * j)) & 63u;
                    sm->perm[sm->cnt[code] + sm->rank[32 * j + lane]] = (unsigned char)(32 * j + lane);
                }
            }
        }
        __syncwarp();
    }
    const unsigned int wb = ws[WQ_BASE];
    unsigned int off = cur - wb + lane;
    if (ws[WQ_SORTED]) off = sm->perm[off];
    pos = wb + off;
    __syncwarp();
    if (lane == 0u) ws[WQ_CUR] = cur + 32u;
    __syncwarp();
    return true;
}
// Appends the slots of the lanes with `alive` to output queue k of this warp and returns each such lane's position (all
// lanes of the warp call it). *length is the queue's length in reserved positions.
__device__ __forceinline__ unsigned int wf_push(unsigned int *queue, unsigned int *length, unsigned int block, unsigned int *ws, int k, bool alive, unsigned int slot) {
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int m = __ballot_sync(0xFFFFFFFFu, alive);
    if (m == 0u) return 0u;
    const unsigned int n = __popc(m), rank = __popc(m & ((1u << lane) - 1u));
    const unsigned int next = ws[WQ_OUT + 2 * k], room = ws[WQ_OUT + 2 * k + 1] - next;
    __syncwarp();
    unsigned int pos = next + rank;
    if (n > room) {   // warp-uniform: the survivors beyond the room left go to a new block
        unsigned int nb = 0u;
        if (lane == 0) nb = atomicAdd(length, block);
        nb = __shfl_sync(0xFFFFFFFFu, nb, 0);
        if (rank >= room) pos = nb + (rank - room);
        if (lane == 0) { ws[WQ_OUT + 2 * k] = nb + (n - room); ws[WQ_OUT + 2 * k + 1] = nb + block; }
    } else if (lane == 0) {
        ws[WQ_OUT + 2 * k] = next + n;
    }
    __syncwarp();
    if (alive) WF_ST(queue + pos, slot);
    return pos;
}
// End of the kernel: pad the rest of the warp's last block of output queue k.
__device__ __forceinline__ void wf_out_finish(unsigned int *queue, const unsigned int *ws, int k) {
    const unsigned int end = ws[WQ_OUT + 2 * k + 1];
    for (unsigned int p = ws[WQ_OUT + 2 * k] + (threadIdx.x & 31u); p < end; p += 32u) WF_ST(queue + p, WF_INVALID);
}
// Entries a kernel really processed (its input queue's padding excluded): per-warp count, one atomic at the end.
__device__ __forceinline__ void wf_note_entries(unsigned int *ws, unsigned int ballot) {
    if ((threadIdx.x & 31u) == 0u) ws[WQ_N] += __popc(ballot);
}
__device__ __forceinline__ void wf_count_entries(unsigned int *entries, const unsigned int *ws) {
    __syncwarp();
    if ((threadIdx.x & 31u) == 0u && ws[WQ_N]) atomicAdd(entries, ws[WQ_N]);
}

#ifndef RT_WF_MINB
#define RT_WF_MINB 8
#endif
__device__ __forceinline__ void wf_state_init(PathState &st, CandList &cands, const WfArgs &w) {
    st.cl = cands.v;
    st.mode = 2; st.t_light = 0.f; st.light = 0;
    st.ray.o = v3(0.f); st.ray.d = v3(0.f, 0.f, 1.f); st.ray.time = 0.f;
    st.rng.key = 0; st.rng.ctr = 0;
    st.cm0 = st.cm1 = st.cm2 = st.cm3 = 0u; st.cl_n = 0;
    st.recs = nullptr; st.wf_rec = w.rec; st.wf_stride = w.rec_stride; st.max_bounces = w.max_bounces;
}
__device__ __forceinline__ void wf_store_ray(const WfArgs &w, unsigned int pos, const PathState &st) {
    WF_ST(w.nray0 + pos, make_float4(st.ray.o.x, st.ray.o.y, st.ray.o.z, st.ray.time));
    WF_ST(w.nray1 + pos, make_float4(st.ray.d.x, st.ray.d.y, st.ray.d.z, u2f((uint32_t)st.N | ((uint32_t)st.depth << 8))));
    WF_ST(w.nrng + pos, make_uint2(st.rng.key, st.rng.ctr));
}

// MESH = 3: the scene has no meshes (neither the walk nor the triangle shading is compiled in).
// MESH = 0: one kernel per level, analytic primitives and meshes. MESH = 1 / 2: two kernels. Phase 1 intersects the analytic
// primitives and asks ray_touches_meshes: rays that cannot hit a mesh are shaded at once; the others go to a queue with their
// analytic hit so far. Phase 2 walks the meshes for THOSE rays only and shades them. Two smaller kernels (frames of 320 and
// 344 B against 592) instead of one; whether that pays depends on the scene (rt_render_device: wf_split).
// CTAs per SM of the instantiations for scenes without meshes (A/B knobs; 8 = 64 registers, 7 = 72, 6 = 80)
#ifndef RT_MINB_TRACE_NOMESH
#define RT_MINB_TRACE_NOMESH RT_WF_MINB
#endif
#ifndef RT_MINB_CLASSIFY_NOMESH
#define RT_MINB_CLASSIFY_NOMESH RT_WF_MINB
#endif
#ifndef RT_MINB_SAMPLE_NOMESH
#define RT_MINB_SAMPLE_NOMESH RT_WF_MINB
#endif
template <bool STATS, bool LC, bool NOLIGHT, int MESH>
__global__ void __launch_bounds__(128, MESH == 3 ? RT_MINB_TRACE_NOMESH : RT_WF_MINB) k_wf_trace(const DScene scene_, const DCamera cam, const WfArgs w) {
    const DScene &scene = RT_S(scene_);
    stage_abvh(scene);
    Counters cnt;
    if (STATS) memset(&cnt, 0, sizeof cnt);
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int count = MESH == 2 ? w.ctr[WF_NCTR * w.level + 11] : (w.level == 0 ? w.n_paths : w.ctr[WF_NCTR * w.level + 1]);
    __shared__ unsigned int wq_all[4][WQ_WORDS];
    unsigned int *const ws = wq_init(wq_all);
    __shared__ WfSortSm sort_all[RT_WF_SORT && MESH != 2 ? 4 : 1];
    const bool sorting = RT_WF_SORT && MESH != 2 && w.sort != 0 && w.level > 0;
    for (;;) {
        unsigned int base, i;
        if (sorting) {
            if (!wf_next_batch_sorted(w.ctr + WF_NCTR * w.level, count, w.max_grab, ws, w.q_live, sort_all + (RT_WF_SORT && MESH != 2 ? (threadIdx.x >> 5) : 0), i)) break;
        } else {
            if (!wf_next_batch(w.ctr + WF_NCTR * w.level + (MESH == 2 ? 10 : 0), count, w.max_grab, ws, base)) break;
            i = base + lane;
        }
        bool valid = i < count;
        const unsigned int mi = i;   // MESH == 2: position in the mesh-walk queue
        if (MESH == 2 && valid) {
            i = WF_LD(w.q_mesh + mi);   // position of the ray in the live queue (level 0: the path slot)
            valid = i != WF_INVALID;
        }
        PathState st;
        CandList cands;
        wf_state_init(st, cands, w);
        unsigned int slot = 0;
        if (valid) {
            if (w.level == 0) {
                slot = i;
                const float4 r = __ldg(w.cam_rays + slot);
                st.ray.o = ld3(cam.pos); st.ray.d = v3(r.x, r.y, r.z); st.ray.time = r.w;
                st.rng.key = __ldg(w.cam_keys + slot); st.rng.ctr = 3u;
                st.N = w.max_bounces; st.depth = 0;
                if (STATS && MESH != 2) cnt.rnd += 3;
            } else {
                slot = WF_LD(w.q_live + i);
                valid = slot != WF_INVALID;
                if (RT_WF_SORT) slot &= WF_SLOT_MASK;   // bits 26..31: direction code of the ray (wf_dir_code)
                if (valid) {
                    const float4 a = WF_LD(w.ray0 + i), b = WF_LD(w.ray1 + i);
                    const uint2 g = WF_LD(w.rng_ray + i);
                    st.ray.o = v3(a.x, a.y, a.z); st.ray.time = a.w; st.ray.d = v3(b.x, b.y, b.z);
                    st.N = (int)(f2u(b.w) & 0xFFu); st.depth = (int)(f2u(b.w) >> 8);
                    st.rng.key = g.x; st.rng.ctr = g.y;
                }
            }
            if (valid) { st.path = slot; st.mode = 0; }
        }
        {
            const unsigned int bv = __ballot_sync(0xFFFFFFFFu, valid);
            if (bv == 0u) continue;   // a run of padding
            if (MESH != 2) wf_note_entries(ws, bv);
        }
        Hit h;
        float hu = 0.f, hv = 0.f;
        bool blocked;
        if (MESH == 2) {
            // the analytic hit so far, then the meshes (Scene.h:221-228), as intersect_lc would have gone on
            h.type = 0; h.obj = -1; h.t = FLT_MAX; h.ref = 0;
            bool done = !valid;
            blocked = false;
            if (valid) {
                const float4 m = WF_LD(w.mesh_hit + mi);
                h.t = m.x; h.type = (int)(f2u(m.y) >> 28); h.obj = h.type ? (int)(f2u(m.y) & 0x0FFFFFFFu) : -1; hu = m.z; hv = m.w;
            }
            meshes_walk_merged<STATS>(scene, st.ray, 0, st.rng, h, blocked, done, &cnt);
        } else if (LC) {
            // the flat box test (<= 32 analytic primitives) is compiled into the split trace kernel only: that is where small analytic sets end up
            // (scenes with meshes and no lights); in the other instantiations its code alone cost config 2 and 5 time
            intersect_lc<STATS, true, false, 0, MESH == 1>(scene, st, true, valid, h, hu, hv, blocked, &cnt, MESH == 0, MESH == 1 && w.flat != 0 && w.level >= w.flat - 1);   // MESH == 3: a scene without meshes
        } else {
            intersect_ray<STATS, true>(scene, st.ray, st.mode, st.t_light, st.rng, h, hu, hv, blocked, &cnt);
        }
        if (MESH == 1) {
            // rays that touch a mesh leave for the walk kernel with what they have; the others are final
            const bool touch = valid && ray_touches_meshes(scene, st.ray, h.t);
            const unsigned int mp = wf_push(w.q_mesh, w.ctr + WF_NCTR * w.level + 11, w.block, ws, 1, touch, i);
            if (touch) {
                WF_ST(w.mesh_hit + mp, make_float4(h.t, u2f(((uint32_t)h.type << 28) | ((uint32_t)h.obj & 0x0FFFFFFFu)), hu, hv));
                valid = false;
            }
        }
        // a path is lit iff its ray hit something (path_shade ends the path on a miss, Scene.h:270-272): the position of its hit
        // record is reserved BEFORE the hit is shaded, and the record is written in the very branch that shaded it, so that the
        // shaded values go from registers straight to the record (with the reservation or a merge point in between they made a
        // round trip through the thread's local-memory frame)
        unsigned int pos = 0u;
        if (!NOLIGHT) pos = wf_push(w.q_hit, w.ctr + WF_NCTR * w.level + 3, w.block, ws, 0, valid && h.type != 0, slot);
        bool alive = false;
        if (valid) {
            V3 c;
            if (path_shade<STATS, true, MESH == 3 || MESH == 1>(scene, st, h, hu, hv, c, &cnt)) {   // MESH == 1: rays that may hit a mesh have left for the walk kernel
                float *o = w.samples + 3ull * slot;
                o[0] = c.x; o[1] = c.y; o[2] = c.z;
            } else if (NOLIGHT) {
                // no lights: the light stage would only scatter (Scene.h:305-342 with an empty light list)
                if (path_next_light_or_bounce<STATS, LC, true>(scene, st, w.nb_ech, c, &cnt)) {
                    float *o = w.samples + 3ull * slot;
                    o[0] = c.x; o[1] = c.y; o[2] = c.z;
                } else {
                    alive = true;
                }
            } else {
                const uint32_t kind = (uint32_t)h.type, obj = (uint32_t)h.obj;   // st.mat = {sph,sq,mesh}_mat[obj]
                WF_ST(w.hit0 + pos, make_float4(st.P.x, st.P.y, st.P.z, st.ray.time));
                // emission and incoming direction travel only when they matter (an emitter; a glass or mirror scatter):
                // the hit record is the largest item of the wavefront's HBM traffic
                const uint32_t has_e = (st.e.x != 0.f || st.e.y != 0.f || st.e.z != 0.f) ? WF_HIT_E : 0u;
                const uint32_t has_d = st.mat->type != 0 ? WF_HIT_D : 0u;
                WF_ST(w.hit1 + pos, make_float4(st.n.x, st.n.y, st.n.z, u2f((kind << 28) | has_e | has_d | obj)));
                WF_ST(w.hit2 + pos, make_float4(st.kd.x, st.kd.y, st.kd.z, u2f((uint32_t)st.N | ((uint32_t)st.depth << 8))));
                if (has_e) WF_ST(w.hit3 + pos, make_float4(st.e.x, st.e.y, st.e.z, 0.f));
                if (has_d) WF_ST(w.hit4 + pos, make_float4(st.in_d.x, st.in_d.y, st.in_d.z, 0.f));
                WF_ST(w.rng_hit + pos, make_uint2(st.rng.key, st.rng.ctr));
            }
        }
        if (NOLIGHT) {
            pos = wf_push(w.q_next, w.ctr + WF_NCTR * (w.level + 1) + 1, w.block, ws, 0, alive, alive ? wf_live_word(slot, st.ray.d) : 0u);
            if (alive) wf_store_ray(w, pos, st);
        }
    }
    wf_out_finish(NOLIGHT ? w.q_next : w.q_hit, ws, 0);
    if (MESH == 1) wf_out_finish(w.q_mesh, ws, 1);
    if (MESH != 2) wf_count_entries(w.ctr + WF_NCTR * w.level + 8, ws);   // closest-hit rays of this level
    if (STATS) flush_counters(cnt, w.stats);
}

// PHASE 0: everything in one kernel (scenes without the analytic hierarchy: no candidate masks to classify by).
// PHASE 1: per light, ONE cone walk; a light whose cone holds no possible occluder is finished on the spot (all its
//          samples are unoccluded: lc_light_unoccluded) and the path goes on to the next light / the scatter. A path
//          that reaches a light with candidates is PARKED: colour so far, light index and masks go to HBM and its
//          hit-queue position to the shadow queue. Most hits of config 2 never leave this phase.
// PHASE 2: the parked paths, now packed into full warps: the NB_ECH samples of the parked light, any further light
//          (walk + samples inline), scatter.
// PHASE 3: phase 2 for a scene with exactly ONE light (every BASELINE scene with lights): compiled for the sample tests only.
// Split because a warp is as slow as its slowest lane: with both phases in one loop, one lane that has to sample keeps
// 31 finished lanes waiting for ten rounds.
// NOMESH: the scene has no meshes (phases 1 and 3): no candidate-triangle collection, list or mesh walk is compiled in — the sample kernel
// then needs no traversal stack at all.
// OVER (phase 3): 0 = the park queue (every entry has its candidate-triangle list), 1 = the overflow queue (every entry walks the meshes).
template <bool STATS, bool LC, int PHASE, bool NOMESH = false, int OVER = 0>
__global__ void __launch_bounds__(128, NOMESH ? (PHASE == 3 ? RT_MINB_SAMPLE_NOMESH : RT_MINB_CLASSIFY_NOMESH) : RT_WF_MINB) k_wf_light(const DScene scene_, const WfArgs w) {
    const DScene &scene = RT_S(scene_);
    stage_abvh(scene);
    Counters cnt;
    if (STATS) memset(&cnt, 0, sizeof cnt);
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int count = w.ctr[WF_NCTR * w.level + (PHASE >= 2 ? 5 + 2 * w.which_park : 3)];
    unsigned int *const head = w.ctr + WF_NCTR * w.level + (PHASE >= 2 ? 4 + 2 * w.which_park : 2);
    const unsigned int *const q_in = PHASE >= 2 ? (w.which_park ? w.q_over : w.q_park) : w.q_hit;
    __shared__ unsigned int wq_all[4][WQ_WORDS];
    unsigned int *const ws = wq_init(wq_all);
    for (;;) {
        unsigned int base;
        if (!wf_next_batch(head, count, w.max_grab, ws, base)) break;
        const unsigned int i = base + lane;
        bool valid = i < count;
        PathState st;
        CandList cands;
        wf_state_init(st, cands, w);
        unsigned int slot = 0, hp = i;   // hp: position of the hit in the hit queue (where its records are)
        bool fin = true, parked = false;
        V3 c = v3(0.f);
        if (valid) {
            if (PHASE >= 2) {
                hp = WF_LD(q_in + i);
                valid = hp != WF_INVALID;
                if (valid) slot = WF_LD(w.q_hit + hp);
            } else {
                slot = WF_LD(q_in + i);
                valid = slot != WF_INVALID;
            }
        }
        {
            const unsigned int bv = __ballot_sync(0xFFFFFFFFu, valid);
            if (bv == 0u) continue;   // a run of padding
            wf_note_entries(ws, bv);
        }
        if (valid) {
            const float4 h0 = WF_LD(w.hit0 + hp), h1 = WF_LD(w.hit1 + hp), h2 = WF_LD(w.hit2 + hp);
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 h3 = (f2u(h1.w) & WF_HIT_E) ? WF_LD(w.hit3 + hp) : zero4, h4 = (f2u(h1.w) & WF_HIT_D) ? WF_LD(w.hit4 + hp) : zero4;
            const uint2 g = WF_LD(w.rng_hit + hp);
            st.P = v3(h0.x, h0.y, h0.z); st.ray.time = h0.w;
            st.n = v3(h1.x, h1.y, h1.z);
            const uint32_t code = f2u(h1.w), kind = code >> 28, obj = code & 0x03FFFFFFu;
            st.mat = kind == 3u ? scene.mesh_mat + obj : (kind == 2u ? scene.sq_mat + obj : scene.sph_mat + obj);
            st.kd = v3(h2.x, h2.y, h2.z);
            st.N = (int)(f2u(h2.w) & 0xFFu); st.depth = (int)(f2u(h2.w) >> 8);
            st.e = v3(h3.x, h3.y, h3.z);
            st.in_d = v3(h4.x, h4.y, h4.z);
            st.rng.key = g.x; st.rng.ctr = g.y;
            st.path = slot;
            if (PHASE >= 2) {
                const float4 p0 = WF_LD(w.park0 + hp), p1 = WF_LD(w.park1 + hp);
                st.color = v3(p0.x, p0.y, p0.z);
                st.light = (int)(f2u(p0.w) & 0xFFu); st.cl_n = (int)((f2u(p0.w) >> 8) & 0xFFu) - 1;
                for (int q4 = 0; !NOMESH && !(PHASE == 3 && OVER) && q4 * 4 < st.cl_n; ++q4) {
                    const float4 v = WF_LD(w.park2 + (size_t)q4 * w.park_stride + hp);
                    st.cl[4 * q4] = f2u(v.x); st.cl[4 * q4 + 1] = f2u(v.y); st.cl[4 * q4 + 2] = f2u(v.z); st.cl[4 * q4 + 3] = f2u(v.w);
                }
                st.cm0 = f2u(p1.x); st.cm1 = f2u(p1.y); st.cm2 = f2u(p1.z); st.cm3 = f2u(p1.w);
                st.j = 0; st.blocked = 0; st.mode = 3;
                if (PHASE == 3) {
                    if (NOMESH) path_shadow_sample_body<STATS, true>(scene, st, &cnt);   // candidates known: first sample (normalisations inlined: see normalized_inl)
                    else path_shadow_sample<STATS, true>(scene, st, &cnt);
                    fin = false;
                } else {
                    Hit h; h.type = 0; h.obj = -1; h.t = 0.f; h.ref = 0;
                    fin = path_advance<STATS, LC, true>(scene, st, h, 0.f, 0.f, false, w.nb_ech, c, &cnt);   // candidates known: first sample
                }
            } else {
                st.color = v3(0.f); st.light = 0; st.mode = 0;
                fin = path_next_light_or_bounce<STATS, LC, true>(scene, st, w.nb_ech, c, &cnt);
            }
        }
        if (PHASE == 1 && LC && RT_OPT_LC_COLLECT) {
            // classify: a live lane is always in mode 3 here (a light with candidates parks the path, one without goes on to the next
            // light or scatters), so only the cone walk is compiled in (intersect_lc<.., COLLECT>) and the unoccluded light is
            // finished right here (path_advance's mode-3 branch without its first-sample half)
            for (;;) {
                const bool mine = valid && !fin && !parked && st.mode == 3;
                if (__ballot_sync(0xFFFFFFFFu, mine) == 0u) break;
                Hit h;
                float hu = 0.f, hv = 0.f;
                bool blocked;
                intersect_lc<STATS, false, true>(scene, st, true, mine, h, hu, hv, blocked, &cnt, !NOMESH);
                if (mine) {
                    bool go_on = false;
                    if (lc_light_unoccluded(st)) {
                        if (STATS) { cnt.shadow += w.nb_ech; cnt.rnd += 3 * w.nb_ech; }
                        st.rng.ctr += 3u * (uint32_t)w.nb_ech;
                        go_on = true;
                    } else if (lc_umbra(scene, st, w.nb_ech)) {
                        // every sample blocked by construction (the stream already advanced): path_finish_light with blocked = NB_ECH,
                        // i.e. shadow = (float)(1. - (double)((float)n / (float)n)) = 0.f (Scene.h:331-333)
                        if (STATS) { cnt.shadow += w.nb_ech; cnt.rnd += 4 * w.nb_ech; }
                        st.color = st.color * 0.f;
                        go_on = true;
                    } else parked = true;
                    if (go_on) {
                        ++st.light;
                        fin = path_next_light_or_bounce<STATS, LC, true>(scene, st, w.nb_ech, c, &cnt);
                    }
                }
            }
        } else if (PHASE == 3) {
            // samples of a scene with ONE light: a live lane is always in mode 1 here (after the last sample the path scatters: there
            // is no further light to collect candidates for), so only the sample tests are compiled in (intersect_lc<.., SAMPLE>)
            for (;;) {
                const bool mine = valid && !fin && st.mode == 1;
                if (__ballot_sync(0xFFFFFFFFu, mine) == 0u) break;
                Hit h;
                float hu = 0.f, hv = 0.f;
                bool blocked;
                intersect_lc<STATS, false, false, NOMESH ? 1 : (OVER ? 3 : 2)>(scene, st, false, mine, h, hu, hv, blocked, &cnt, !NOMESH);
                if (mine) {
                    if (blocked) ++st.blocked;
                    if (++st.j < w.nb_ech) { if (NOMESH) path_shadow_sample_body<STATS, true>(scene, st, &cnt); else path_shadow_sample<STATS, true>(scene, st, &cnt); }
                    else fin = path_finish_light<STATS, LC, true>(scene, st, w.nb_ech, c, &cnt);   // light 1 of 1: scatters
                }
            }
        } else
        for (;;) {
            const bool live = valid && !fin && !parked && st.mode != 0;
            const unsigned int bt = __ballot_sync(0xFFFFFFFFu, live && st.mode == 3), bs = __ballot_sync(0xFFFFFFFFu, live && st.mode == 1);
            if ((bt | bs) == 0u) break;
            const bool run_t = bt != 0u;
            const bool mine = live && (run_t ? st.mode == 3 : st.mode == 1);
            Hit h;
            float hu = 0.f, hv = 0.f;
            bool blocked;
            if (LC) intersect_lc<STATS>(scene, st, run_t, mine, h, hu, hv, blocked, &cnt);
            else intersect_ray<STATS, true>(scene, st.ray, mine ? st.mode : 2, st.t_light, st.rng, h, hu, hv, blocked, &cnt);
            if (mine) {
                if (PHASE == 1 && !lc_light_unoccluded(st)) parked = true;   // mode 3 only in this phase
                else fin = path_advance<STATS, LC, true>(scene, st, h, hu, hv, blocked, w.nb_ech, c, &cnt);
            }
        }
        // reserve the queue positions first (who is alive / parked is known by now), then write each path's records in the
        // branch that holds them in registers
        const bool alive = valid && !parked && !fin;
        const unsigned int pos = wf_push(w.q_next, w.ctr + WF_NCTR * (w.level + 1) + 1, w.block, ws, 0, alive, alive ? wf_live_word(st.path, st.ray.d) : 0u);   // st.path = the slot
        if (PHASE == 1) {
            wf_push(w.q_park, w.ctr + WF_NCTR * w.level + 5, w.block, ws, 1, parked && st.cl_n >= 0, hp);
            wf_push(w.q_over, w.ctr + WF_NCTR * w.level + 7, w.block, ws, 2, parked && st.cl_n < 0, hp);
        }
        if (valid) {
            if (parked) {
                WF_ST(w.park0 + hp, make_float4(st.color.x, st.color.y, st.color.z, u2f((uint32_t)st.light | ((uint32_t)(st.cl_n + 1) << 8))));
                for (int q4 = 0; !NOMESH && q4 * 4 < st.cl_n; ++q4)
                    WF_ST(w.park2 + (size_t)q4 * w.park_stride + hp, make_float4(u2f(st.cl[4 * q4]), u2f(st.cl[4 * q4 + 1]), u2f(st.cl[4 * q4 + 2]), u2f(st.cl[4 * q4 + 3])));
                WF_ST(w.park1 + hp, make_float4(u2f(st.cm0), u2f(st.cm1), u2f(st.cm2), u2f(st.cm3)));
                WF_ST(w.rng_hit + hp, make_uint2(st.rng.key, st.rng.ctr));
            } else if (fin) {
                float *o = w.samples + 3ull * st.path;
                o[0] = c.x; o[1] = c.y; o[2] = c.z;
            } else {
                wf_store_ray(w, pos, st);
            }
        }
    }
    wf_out_finish(w.q_next, ws, 0);
    if (PHASE == 1) {
        wf_out_finish(w.q_park, ws, 1);
        wf_out_finish(w.q_over, ws, 2);
    }
    if (PHASE < 2) wf_count_entries(w.ctr + WF_NCTR * w.level + 9, ws);   // hits of this level
    if (STATS) flush_counters(cnt, w.stats);
}

// The light stage of a scene WITHOUT lights (Scene.h:305-342 with an empty light list): colour stays 0, the path scatters
// (Material::scatter), records (0, kd, e) and either ends (last bounce: fold) or queues its next ray. Same arithmetic as
// k_wf_light's phase 1 on such a scene, without the machinery that phase carries for the cone walks (path state in the
// thread's frame, traversal stack, park queues): the pool scene's light kernel ran at 30 % issue utilisation behind 3 TB/s of
// DRAM traffic, 490 B of L2 writes per hit of which 92 B were records — the rest was the frame (profiles/r02_notes.md).
template <bool STATS>
__global__ void __launch_bounds__(128, RT_WF_MINB) k_wf_scatter(const DScene scene_, const WfArgs w) {
    const DScene &scene = RT_S(scene_);
    Counters cnt;
    if (STATS) memset(&cnt, 0, sizeof cnt);
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int count = w.ctr[WF_NCTR * w.level + 3];
    __shared__ unsigned int wq_all[4][WQ_WORDS];
    unsigned int *const ws = wq_init(wq_all);
    for (;;) {
        unsigned int base;
        if (!wf_next_batch(w.ctr + WF_NCTR * w.level + 2, count, w.max_grab, ws, base)) break;
        const unsigned int i = base + lane;
        unsigned int slot = WF_INVALID;
        if (i < count) slot = WF_LD(w.q_hit + i);
        const bool valid = slot != WF_INVALID;
        {
            const unsigned int bv = __ballot_sync(0xFFFFFFFFu, valid);
            if (bv == 0u) continue;   // a run of padding
            wf_note_entries(ws, bv);
        }
        bool alive = false;
        Ray next; next.o = v3(0.f); next.d = v3(0.f, 0.f, 1.f); next.time = 0.f;
        Rng rng; rng.key = 0u; rng.ctr = 0u;
        int N = 0, depth = 0;
        if (valid) {
            const float4 h0 = WF_LD(w.hit0 + i), h1 = WF_LD(w.hit1 + i), h2 = WF_LD(w.hit2 + i);
            const uint32_t code = f2u(h1.w), kind = code >> 28, obj = code & 0x03FFFFFFu;
            const DMaterial *mat = kind == 3u ? scene.mesh_mat + obj : (kind == 2u ? scene.sq_mat + obj : scene.sph_mat + obj);
            V3 e = v3(0.f);
            if (code & WF_HIT_E) { const float4 h3 = WF_LD(w.hit3 + i); e = v3(h3.x, h3.y, h3.z); }
            Ray in; in.o = v3(h0.x, h0.y, h0.z); in.d = v3(0.f); in.time = h0.w;
            if (code & WF_HIT_D) { const float4 h4 = WF_LD(w.hit4 + i); in.d = v3(h4.x, h4.y, h4.z); }
            const uint2 g = WF_LD(w.rng_hit + i);
            rng.key = g.x; rng.ctr = g.y;
            N = (int)(f2u(h2.w) & 0xFFu); depth = (int)(f2u(h2.w) >> 8);
            next = material_scatter<STATS>(*mat, in, v3(h1.x, h1.y, h1.z), in.o, rng, &cnt);
            wf_rec_store(w.rec, w.rec_stride, slot, depth, v3(0.f), v3(h2.x, h2.y, h2.z), e);
            ++depth; --N;
            if (N == 0) {
                const V3 c = wf_fold(w.rec, w.rec_stride, slot, depth, w.max_bounces, v3(0.f));
                float *o = w.samples + 3ull * slot;
                o[0] = c.x; o[1] = c.y; o[2] = c.z;
            } else {
                alive = true;
            }
        }
        const unsigned int pos = wf_push(w.q_next, w.ctr + WF_NCTR * (w.level + 1) + 1, w.block, ws, 0, alive, alive ? wf_live_word(slot, next.d) : 0u);
        if (alive) {
            WF_ST(w.nray0 + pos, make_float4(next.o.x, next.o.y, next.o.z, next.time));
            WF_ST(w.nray1 + pos, make_float4(next.d.x, next.d.y, next.d.z, u2f((uint32_t)N | ((uint32_t)depth << 8))));
            WF_ST(w.nrng + pos, make_uint2(rng.key, rng.ctr));
        }
    }
    wf_out_finish(w.q_next, ws, 0);
    wf_count_entries(w.ctr + WF_NCTR * w.level + 9, ws);   // hits of this level
    if (STATS) flush_counters(cnt, w.stats);
}

// image[pixel] = (sum of its samples, in sample order) / nsamples ; then gamma
// Progressive accumulation (rt_accum_add): `sum` holds each pixel's running sum over the `prior` samples of earlier
// passes; this pass continues THAT sum with its own samples in order, stores it back and divides by prior + spp, so
// the float additions are the very sequence a single render of prior + spp samples performs (main.cpp:188-195).
// Output layout: packed (this rank's pixels in tile order) or — image_mode — the pixel's place in the row-major image
// of the rendered rectangle. In image mode the buffers may live on ANOTHER GPU (peer-mapped over NVLink): every rank's
// resolve then stores its tiles straight into rank 0's framebuffer — the multi-GPU "gather" is these stores, there is
// no pack / collective / untile pass behind them.
__global__ void k_resolve(const RenderArgs a, const unsigned int *xy, unsigned int n_pixels, float *linear_out, float *gamma_out,
                          float *sum, unsigned int prior, int image_mode, int rx0, int ry0, int rw) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const int spp = a.spp;
    const float *s = a.samples + 3ull * i * (unsigned int)spp;
    const unsigned long long op = 3ull * (a.pixel_begin + i);   // packed position (the running sums are always packed)
    unsigned long long o = op;
    if (image_mode) {
        int x, y;
        if (xy) { const unsigned int v = __ldg(xy + i); x = (int)(v & 0xFFFFu); y = (int)(v >> 16); }
        else packed_to_xy(a, (unsigned int)a.pixel_begin + i, x, y);
        o = 3ull * ((unsigned long long)(y - ry0) * (unsigned long long)rw + (unsigned long long)(x - rx0));
    }
    V3 acc = v3(0.f);
    if (sum && prior) acc = v3(sum[op], sum[op + 1], sum[op + 2]);
    for (int k = 0; k < spp; ++k) acc = acc + v3(s[3 * k], s[3 * k + 1], s[3 * k + 2]);
    if (sum) { sum[op] = acc.x; sum[op + 1] = acc.y; sum[op + 2] = acc.z; }
    acc = acc / (float)(prior + (unsigned int)spp);
    if (linear_out) { linear_out[o] = acc.x; linear_out[o + 1] = acc.y; linear_out[o + 2] = acc.z; }
    if (gamma_out) { gamma_out[o] = gamma_channel(acc.x); gamma_out[o + 1] = gamma_channel(acc.y); gamma_out[o + 2] = gamma_channel(acc.z); }
}

// Ray counts of a wavefront chunk from its queue counters, for free: level L traced ctr[L][8] closest-hit rays (level 0:
// every path) and lit ctr[L][9] hits; every hit fires nb_ech shadow rays at each light (Scene.h:305-334), traced or
// proven unoccluded. tally[0] += closest rays, tally[1] += hits. (Without lights the hits are not queued, and not needed.)
__global__ void k_wf_tally(const unsigned int *ctr, unsigned int n_paths, int levels, unsigned long long *tally) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long closest = 0, hits = 0;
    for (int l = 0; l < levels; ++l) {
        closest += ctr[WF_NCTR * l + 8];
        hits += ctr[WF_NCTR * l + 9];
    }
    (void)n_paths;
    atomicAdd(tally, closest);       // chunks of one frame run on two streams
    atomicAdd(tally + 1, hits);
}

// main.cpp:258: (int)(255.f * std::min<float>(1.f, c)); std::min(a, b) = (b < a) ? b : a, so a NaN pixel becomes 1.f and
// is written as 255 (the float P3 writer of host/Renderer.cpp does the same); negative values, whose int the reference
// would print with a minus sign that no PPM reader accepts, are written as 0
__global__ void k_quantize(const float *v, size_t n, uint8_t *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float c = v[i];
    const float m = 255.f * (c < 1.f ? c : 1.f);
    out[i] = m > 0.f ? (uint8_t)(int)m : (uint8_t)0;
}

__global__ void k_untile(const TileRec *tiles, const unsigned int *tile_off, int n_tiles, const float *packed,
                         float *image, int rect_x0, int rect_y0, int rect_w, unsigned int n_pixels) {
    const unsigned int lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= n_pixels) return;
    int lo = 0, hi = n_tiles - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_off[mid] <= lp) lo = mid; else hi = mid - 1;
    }
    const TileRec t = tiles[lo];
    const unsigned int r = lp - tile_off[lo];
    const int x = t.x0 + (int)(r % (unsigned int)t.w) - rect_x0, y = t.y0 + (int)(r / (unsigned int)t.w) - rect_y0;
    const size_t o = 3 * ((size_t)y * rect_w + x);
    image[o] = packed[3ull * lp]; image[o + 1] = packed[3ull * lp + 1]; image[o + 2] = packed[3ull * lp + 2];
}

__global__ void k_primary_ids(const DScene scene_, const DCamera cam, int width, int height, unsigned int seed, int x0,
                              int y0, int rw, int rh, uint32_t *ids) {
    const DScene &scene = RT_S(scene_);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rw * rh) return;
    const int x = x0 + i % rw, y = y0 + i / rw;
    Rng rng;
    rng.init(seed, (unsigned int)x + (unsigned int)y * (unsigned int)width, 0u);
    const Ray ray = primary_ray(cam, x, y, width, height, rng);
    float u, v;
    const Hit h = closest_hit<false>(scene, ray, u, v, nullptr);
    uint32_t *o = ids + 4 * (size_t)i;
    o[0] = (uint32_t)h.type;
    o[1] = h.type ? (uint32_t)h.obj : 0u;
    o[2] = h.type == 3 ? f2u(__ldg(scene.tri_den + h.ref).y) : 0u;
    o[3] = f2u(h.t);
}

__global__ void k_trace_rays(const DScene scene_, size_t n, const float *org, const float *dir, const float *time,
                             uint32_t *ids, float *aux) {
    const DScene &scene = RT_S(scene_);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Ray ray = make_ray(ld3(org + 3 * i), ld3(dir + 3 * i), time ? time[i] : 0.f);
    float u = 0.f, v = 0.f;
    const Hit h = closest_hit<false>(scene, ray, u, v, nullptr);
    uint32_t *o = ids + 4 * i;
    o[0] = (uint32_t)h.type;
    o[1] = h.type ? (uint32_t)h.obj : 0u;
    o[2] = h.type == 3 ? f2u(__ldg(scene.tri_den + h.ref).y) : 0u;
    o[3] = f2u(h.t);
    if (!aux) return;
    float *q = aux + 8 * i;
    for (int k = 0; k < 8; ++k) q[k] = 0.f;
    if (h.type == 1) {
        const float4 a = scene.sph_a[h.obj], b = scene.sph_b[h.obj];
        const V3 c = v3(a.x, a.y, a.z) + ray.time * v3(b.x, b.y, b.z);
        const V3 P = ray.o + h.t * ray.d;
        const V3 nn = normalized(P - c);
        q[0] = (float)acos((double)nn.y * -1.);
        q[1] = (float)(atan2((double)nn.z * -1., (double)nn.x) + RT_PI);
        q[2] = nn.x; q[3] = nn.y; q[4] = nn.z;
    } else if (h.type == 2) {
        q[0] = u; q[1] = v;
        q[2] = scene.squares[h.obj].n[0]; q[3] = scene.squares[h.obj].n[1]; q[4] = scene.squares[h.obj].n[2];
    } else if (h.type == 3) {
        float w0 = 0, w1 = 0, w2 = 0;
        triangle_t<false>(ray, scene, h.ref, w0, w1, w2, nullptr);
        const float4 pl = scene.tri_plane[h.ref];
        q[0] = w0; q[1] = w1; q[2] = w2; q[3] = pl.x; q[4] = pl.y; q[5] = pl.z;
    }
}

__global__ void k_shade_rays(const DScene scene_, size_t n, const float *org, const float *dir, const float *time,
                             unsigned int seed, int max_bounces, int nb_ech, float *rgb) {
    const DScene &scene = RT_S(scene_);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Ray ray = make_ray(ld3(org + 3 * i), ld3(dir + 3 * i), time ? time[i] : 0.f);
    Rng rng;
    rng.init(seed, (unsigned int)i, 0u);
    rng.ctr = 3;
    const V3 c = trace_path<false>(scene, ray, rng, max_bounces, nb_ech, nullptr);
    rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
}

// FP32 peak probes: 8 independent register chains per thread, 4096 iterations, no memory traffic.
template <bool FUSED>
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, float a, float b, int iters) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
        if (FUSED) {
            x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
            x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
        } else {
            x0 = __fadd_rn(__fmul_rn(x0, a), b); x1 = __fadd_rn(__fmul_rn(x1, a), b); x2 = __fadd_rn(__fmul_rn(x2, a), b); x3 = __fadd_rn(__fmul_rn(x3, a), b);
            x4 = __fadd_rn(__fmul_rn(x4, a), b); x5 = __fadd_rn(__fmul_rn(x5, a), b); x6 = __fadd_rn(__fmul_rn(x6, a), b); x7 = __fadd_rn(__fmul_rn(x7, a), b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// ------------------------------------------------------------------------------------------------
// host side of the C ABI
// ------------------------------------------------------------------------------------------------
// Per-render scratch (grow-only). It outlives the scene that used it: rt_scene_destroy hands it to a per-device pool
// and the next rt_scene_create on that device takes it back, so re-uploading a scene every frame (the host API's
// default flow, and bench.py's e2e leg) does not pay cudaMalloc/cudaFree of gigabytes of path state per frame.
struct Scratch {
    float *samples = nullptr; size_t samples_cap = 0;
    float4 *cam_rays = nullptr; unsigned int *cam_keys = nullptr; size_t cam_cap = 0;   // per path of a chunk (k_camera_rays)
    unsigned int *pix_xy = nullptr; size_t pix_cap = 0;                                 // per pixel of a chunk (k_pixel_xy)
    // wavefront state (variant 6): 13 float4 planes + 2 rng planes + 4 queues per queue position, and 3 float4 per bounce and path slot
    float4 *wf_f4 = nullptr; uint2 *wf_rng = nullptr; unsigned int *wf_q = nullptr; float4 *wf_rec = nullptr; unsigned int *wf_ctr = nullptr;
    size_t wf_cap = 0, wf_pos_cap = 0; int wf_bounces = 0;   // wf_cap: path slots; wf_pos_cap: queue positions (slots + padding of the block reservation)
    unsigned long long *counters = nullptr;   // [0] work counter, [1..10] stats
    TileRec *d_tiles = nullptr; unsigned int *d_tile_off = nullptr; size_t tiles_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float *out_img = nullptr; size_t out_cap = 0;       // framebuffer(s) of rt_render / rt_render_rgb8 / rt_render_multi: no cudaMalloc per frame
    uint8_t *out_bytes = nullptr; size_t out_bytes_cap = 0;
    cudaStream_t stream = nullptr;                      // rt_render_multi: this device's stream (one host thread per device)
    // pinned staging for the scene's small arrays: rt_scene_create copies each array here and issues an ASYNCHRONOUS
    // H2D copy on the null stream (2-3 us each) instead of a synchronous pageable cudaMemcpy (8-10 us each, ~30 per scene);
    // one stream synchronise at the end of the upload. Arrays that do not fit (meshes, images) are copied synchronously.
    char *stage = nullptr; size_t stage_cap = 0, stage_used = 0;
    // The wavefront renders a frame's chunks on TWO streams at once (render_device_impl): the second stream's chunks use this
    // second set of per-chunk buffers (samples, camera rays, pixel table, wavefront state); everything else is shared.
    Scratch *second = nullptr;
    cudaStream_t aux_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // arena for the scene's own arrays (textures, primitives, hierarchies): bump-allocated blocks, reset when the scene
    // is destroyed and reused by the next scene on this device — a re-upload performs no cudaMalloc/cudaFree at all
    struct Block { char *base; size_t cap, used; };
    std::vector<Block> arena;
    void *arena_alloc(size_t bytes, cudaError_t *err) {
        *err = cudaSuccess;
        bytes = (bytes + 255) & ~(size_t)255;
        for (Block &b : arena)
            if (b.cap - b.used >= bytes) { void *p = b.base + b.used; b.used += bytes; return p; }
        Block nb;
        nb.cap = std::max<size_t>(bytes, (size_t)32 << 20);
        nb.used = bytes;
        nb.base = nullptr;
        *err = cudaMalloc((void **)&nb.base, nb.cap);
        if (*err != cudaSuccess) return nullptr;
        arena.push_back(nb);
        return nb.base;
    }
    void arena_reset() { for (Block &b : arena) b.used = 0; }
    void release() {
        for (Block &b : arena) cudaFree(b.base);
        arena.clear();
        if (samples) cudaFree(samples);
        if (cam_rays) cudaFree(cam_rays);
        if (cam_keys) cudaFree(cam_keys);
        if (pix_xy) cudaFree(pix_xy);
        if (wf_f4) cudaFree(wf_f4);
        if (wf_rng) cudaFree(wf_rng);
        if (wf_q) cudaFree(wf_q);
        if (wf_rec) cudaFree(wf_rec);
        if (wf_ctr) cudaFree(wf_ctr);
        if (counters) cudaFree(counters);
        if (d_tiles) cudaFree(d_tiles);
        if (d_tile_off) cudaFree(d_tile_off);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (out_img) cudaFree(out_img);
        if (out_bytes) cudaFree(out_bytes);
        if (stream) cudaStreamDestroy(stream);
        if (stage) cudaFreeHost(stage);
        if (aux_stream) cudaStreamDestroy(aux_stream);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (second) { second->release(); delete second; }
        *this = Scratch();
    }
};
namespace {
std::mutex g_scratch_mu;
std::map<int, std::vector<Scratch>> g_scratch_pool;   // device -> idle scratch sets
}  // namespace

namespace {
// device copies of per-rank tile tables for rt_untile_device, kept per (device, geometry) so that the
// per-frame call allocates nothing and never synchronises
struct UntileKey {
    int device, w, h, x0, y0, x1, y1, tw, th, nr, rk;
    bool operator<(const UntileKey &o) const { return memcmp(this, &o, sizeof *this) < 0; }
};
struct UntileTab { TileRec *tiles = nullptr; unsigned int *off = nullptr; int n = 0; unsigned int pixels = 0; };
std::map<UntileKey, UntileTab> g_untile;
std::mutex g_untile_mu;
}  // namespace

struct RtScene : Scratch {
    std::atomic<int> refs{1};   // the creator's handle + one per accumulator / rt_scene_retain: rt_scene_destroy frees at 0
    int device = 0;
    int sm_count = 0;
    DScene d{};
    size_t bytes = 0;
    std::vector<TileRec> h_tiles; std::vector<unsigned int> h_tile_off;
    // device staging of the squares and capacity of the analytic hierarchy arrays: rt_scene_update_analytic rewrites the
    // analytic primitives IN PLACE (same counts), so an animated scene never re-uploads its meshes and textures
    SquareIn *d_square_in = nullptr;
    size_t abvh_node_cap = 0, abvh_prim_cap = 0;
    uint32_t n_textures = 0, n_normal_maps = 0;
    size_t h2d_bytes = 0;                      // bytes the upload really copied (cached images are not copied again)
    std::vector<uint64_t> cached_images;       // image-cache entries this scene holds a reference on
};

namespace {

// The kernels read the scene from the constant bank (rt_core.cuh : c_scene), one copy per device. A render binds its
// scene for the duration of its launches: the upload is ordered on the render's stream, and a render of ANOTHER scene
// (or on another stream) on the same device first waits for the event that ends the previous binding, so kernels
// still in flight keep the constants they were launched with. Host threads serialise per device while they ENQUEUE
// (launches are asynchronous: a few hundred microseconds), not while the GPU works.
struct ConstSlot {
    std::mutex mu;
    const void *owner = nullptr;
    cudaStream_t last_stream = nullptr;
    cudaEvent_t ev = nullptr;
    bool used = false;
};
std::mutex g_const_mu;
std::map<int, ConstSlot *> g_const_slots;
ConstSlot *const_slot(int device) {
    std::lock_guard<std::mutex> lock(g_const_mu);
    ConstSlot *&p = g_const_slots[device];
    if (!p) p = new ConstSlot();
    return p;
}
class SceneBinding {
public:
    SceneBinding() {}
    ~SceneBinding() { done(); }
    int bind(const RtScene *s, cudaStream_t st);
    void done() {
        if (!slot_) return;
        if (slot_->ev && cudaEventRecord(slot_->ev, st_) == cudaSuccess) slot_->used = true;
        slot_->mu.unlock();
        slot_ = nullptr;
    }
private:
    ConstSlot *slot_ = nullptr;
    cudaStream_t st_ = nullptr;
};
int SceneBinding::bind(const RtScene *s, cudaStream_t st) {
#if RT_SCENE_CONST
    ConstSlot *c = const_slot(s->device);
    c->mu.lock();
    slot_ = c; st_ = st;
    if (!c->ev) RT_CUDA(cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming));
    if (c->used && (c->owner != (const void *)s || c->last_stream != st)) RT_CUDA(cudaStreamWaitEvent(st, c->ev, 0));
    RT_CUDA(cudaMemcpyToSymbolAsync(c_scene, &s->d, sizeof(DScene), 0, cudaMemcpyHostToDevice, st));
    c->owner = s; c->last_stream = st;
#else
    (void)s; (void)st;
#endif
    return RT_OK;
}

struct Rect { int x0, y0, x1, y1, tw, th; };

int resolve_rect(const RtRenderParams &p, Rect &r) {
    if (p.width < 1 || p.height < 1) return fail(RT_ERR_INVALID, "width/height must be >= 1");
    if (p.width > 65535 || p.height > 65535) return fail(RT_ERR_INVALID, "width/height must be <= 65535");
    if (p.spp < 1) return fail(RT_ERR_INVALID, "spp must be >= 1");
    if (p.max_bounces < 0 || p.max_bounces > RT_MAX_BOUNCES) return fail(RT_ERR_INVALID, "max_bounces must be in [0, 16]");
    if (p.nb_ech < 1) return fail(RT_ERR_INVALID, "nb_ech must be >= 1");
    const bool full = (p.x0 | p.y0 | p.x1 | p.y1) == 0;
    r.x0 = full ? 0 : p.x0; r.y0 = full ? 0 : p.y0; r.x1 = full ? p.width : p.x1; r.y1 = full ? p.height : p.y1;
    if (r.x0 < 0 || r.y0 < 0 || r.x1 > p.width || r.y1 > p.height || r.x0 >= r.x1 || r.y0 >= r.y1)
        return fail(RT_ERR_INVALID, "render rectangle out of range");
    r.tw = p.tile_w > 0 ? p.tile_w : 32;
    r.th = p.tile_h > 0 ? p.tile_h : 32;
    if (p.n_ranks > 1 && (p.rank < 0 || p.rank >= p.n_ranks)) return fail(RT_ERR_INVALID, "rank out of range");
    return RT_OK;
}

// Which rank renders tile (tx, ty) of the rectangle's tile grid: (tx + K * ty) mod n, K odd and coprime to n (3, else 5, else 1).
// Round-robin along a row, and every row shifted by K against the one above, so that a rank's tiles lie on diagonals. Plain
// t mod n over the row-major tile number puts a rank's tiles in COLUMNS whenever n divides the tiles per row (3840 / 32 = 120
// tiles, 8 ranks): on the pond scene, whose cost per pixel is a function of x as much as of y, the slowest of 8 ranks then
// needed 17 % longer than the mean (72.7 ms against 62.1, profiles/r02_notes.md).
int tile_owner(int tx, int ty, int n) {
    const int k = n % 3 ? 3 : (n % 5 ? 5 : 1);
    return (int)(((long long)tx + (long long)k * ty) % n);
}
// tiles of the rectangle, row-major numbering; keep the tiles tile_owner() gives to `rank`
void build_tiles(const RtRenderParams &p, const Rect &r, std::vector<TileRec> &tiles, std::vector<unsigned int> &off) {
    tiles.clear(); off.clear();
    const int ntx = (r.x1 - r.x0 + r.tw - 1) / r.tw, nty = (r.y1 - r.y0 + r.th - 1) / r.th;
    const int nr = p.n_ranks > 1 ? p.n_ranks : 1, rk = p.n_ranks > 1 ? p.rank : 0;
    unsigned int acc = 0;
    off.push_back(0);
    for (int ty = 0; ty < nty; ++ty)
        for (int tx = 0; tx < ntx; ++tx) {
            if (tile_owner(tx, ty, nr) != rk) continue;
            TileRec tr;
            tr.x0 = r.x0 + tx * r.tw; tr.y0 = r.y0 + ty * r.th;
            tr.w = std::min(r.tw, r.x1 - tr.x0); tr.h = std::min(r.th, r.y1 - tr.y0);
            tiles.push_back(tr);
            acc += (unsigned int)(tr.w * tr.h);
            off.push_back(acc);
        }
}

// host -> device for the scene's arrays: through the pinned staging buffer and asynchronous when the piece is small,
// synchronous otherwise. Everything is ordered on the null stream; rt_scene_create synchronises it once at the end.
#define RT_STAGE_BYTES ((size_t)1 << 20)
#define RT_STAGE_PIECE ((size_t)128 << 10)
cudaError_t h2d(RtScene *s, void *dst, const void *src, size_t bytes) {
    s->h2d_bytes += bytes;
    if (bytes <= RT_STAGE_PIECE) {
        if (!s->stage && cudaMallocHost((void **)&s->stage, RT_STAGE_BYTES) == cudaSuccess) s->stage_cap = RT_STAGE_BYTES;
        cudaGetLastError();
        const size_t need = (bytes + 63) & ~(size_t)63;
        if (s->stage && s->stage_used + need <= s->stage_cap) {
            char *st = s->stage + s->stage_used;
            s->stage_used += need;
            memcpy(st, src, bytes);
            return cudaMemcpyAsync(dst, st, bytes, cudaMemcpyHostToDevice, 0);
        }
    }
    return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
}
template <class T> int dev_upload(RtScene *s, const T *host, size_t n, const T **out) {
    *out = nullptr;
    if (n == 0) return RT_OK;
    cudaError_t e;
    void *p = s->arena_alloc(n * sizeof(T), &e);
    RT_CUDA(e);
    s->bytes += n * sizeof(T);
    RT_CUDA(h2d(s, p, host, n * sizeof(T)));
    *out = (const T *)p;
    return RT_OK;
}
template <class T> int dev_alloc(RtScene *s, size_t n, T **out) {
    *out = nullptr;
    if (n == 0) return RT_OK;
    cudaError_t e;
    void *p = s->arena_alloc(n * sizeof(T), &e);
    RT_CUDA(e);
    s->bytes += n * sizeof(T);
    *out = (T *)p;
    return RT_OK;
}

struct DevTmp {   // staging buffer for the precompute kernels: lives in the scene's arena (reclaimed with the scene)
    void *p = nullptr;
    cudaError_t put(RtScene *s, const void *h, size_t bytes);
};

cudaError_t DevTmp::put(RtScene *s, const void *h, size_t bytes) {
    cudaError_t e;
    p = s->arena_alloc(bytes ? bytes : 1, &e);
    if (e == cudaSuccess && h && bytes) e = h2d(s, p, h, bytes);
    return e;
}

DMaterial to_dmat(const RtMaterial &m) {
    DMaterial d{};
    d.type = m.type; d.texture_type = m.texture_type;
    for (int k = 0; k < 3; ++k) { d.kd[k] = m.diffuse[k]; d.checker1[k] = m.checker1[k]; d.checker2[k] = m.checker2[k]; d.light_color[k] = m.light_color[k]; d.motion[k] = m.motion[k]; }
    d.transparency = m.transparency; d.index_medium = m.index_medium;
    d.tsx = m.texture_scale_x; d.tsy = m.texture_scale_y;
    d.emissive = m.emissive; d.light_intensity = m.light_intensity;
    d.image = m.image; d.normal_map = m.normal_map;
    return d;
}

// Device-side image cache (RtImage::content_id != 0): textures, normal maps and the sky image stay resident across scene
// re-uploads. Entries are reference counted by the scenes that use them; unreferenced entries are kept (that is the
// point) until the cache exceeds its budget, least recently used first, or rt_release_cached_memory is called.
struct ImgEntry { unsigned char *d = nullptr; size_t bytes = 0; int w = 0, h = 0, refs = 0; unsigned long long stamp = 0; };
struct ImgCache { std::map<uint64_t, ImgEntry> map; size_t total = 0; unsigned long long clock = 0; };
std::mutex g_img_mu;
std::map<int, ImgCache> g_img_cache;
const size_t RT_IMG_CACHE_BUDGET = (size_t)2 << 30;
void img_cache_trim(ImgCache &c, size_t budget) {
    while (c.total > budget) {
        auto victim = c.map.end();
        for (auto it = c.map.begin(); it != c.map.end(); ++it)
            if (it->second.refs == 0 && (victim == c.map.end() || it->second.stamp < victim->second.stamp)) victim = it;
        if (victim == c.map.end()) break;
        cudaFree(victim->second.d);
        c.total -= victim->second.bytes;
        c.map.erase(victim);
    }
}
void img_cache_release(RtScene *s) {
    if (s->cached_images.empty()) return;
    std::lock_guard<std::mutex> lock(g_img_mu);
    ImgCache &c = g_img_cache[s->device];
    for (uint64_t id : s->cached_images) {
        auto it = c.map.find(id);
        if (it != c.map.end() && it->second.refs > 0) --it->second.refs;
    }
    s->cached_images.clear();
    img_cache_trim(c, RT_IMG_CACHE_BUDGET);
}
int upload_image(RtScene *s, const RtImage &im, DImage &out) {
    out.w = 0; out.h = 0; out.rgb = nullptr;
    if (im.w < 1 || im.h < 1 || !im.rgb) return RT_OK;
    if (im.content_id != 0) {
        const size_t bytes = (size_t)im.w * im.h * 3;
        std::lock_guard<std::mutex> lock(g_img_mu);
        ImgCache &c = g_img_cache[s->device];
        auto it = c.map.find(im.content_id);
        if (it != c.map.end() && (it->second.w != im.w || it->second.h != im.h)) {   // an id reused for other pixels: not cacheable
            it = c.map.end();
        } else if (it == c.map.end()) {
            ImgEntry e;
            e.bytes = bytes; e.w = im.w; e.h = im.h;
            RT_CUDA(cudaMalloc((void **)&e.d, bytes));
            const cudaError_t ce = cudaMemcpy(e.d, im.rgb, bytes, cudaMemcpyHostToDevice);
            if (ce != cudaSuccess) { cudaFree(e.d); RT_CUDA(ce); }
            s->h2d_bytes += bytes;
            c.total += bytes;
            it = c.map.emplace(im.content_id, e).first;
        }
        if (it != c.map.end()) {
            ++it->second.refs;
            it->second.stamp = ++c.clock;
            s->cached_images.push_back(im.content_id);
            s->bytes += bytes;
            out.w = im.w; out.h = im.h; out.rgb = it->second.d;
            return RT_OK;
        }
    }
    const unsigned char *d = nullptr;
    int rc = dev_upload<unsigned char>(s, im.rgb, (size_t)im.w * im.h * 3, &d);
    if (rc) return rc;
    out.w = im.w; out.h = im.h; out.rgb = d;
    return RT_OK;
}

int check_material(const RtMaterial &m, const RtSceneDesc &d) {
    if (m.type < 0 || m.type > 2 || m.texture_type < 0 || m.texture_type > 2) return fail(RT_ERR_INVALID, "material enum out of range");
    if (m.image >= (int)d.n_textures || m.normal_map >= (int)d.n_normal_maps) return fail(RT_ERR_INVALID, "material image index out of range");
    return RT_OK;
}

int select_device(int device, int *sm_count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(RT_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= n) return fail(RT_ERR_INVALID, "device index out of range");
    // attribute queries, not cudaGetDeviceProperties: that call takes 3-90 ms when it has to wait behind other driver
    // activity, and it sat in every scene upload (profiles/r01_notes.md, e2e spikes)
    int major = 0, minor = 0, sms = 0;
    RT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    RT_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    RT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (major != 10) return fail(RT_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(major * 10 + minor) + ", this library is built for sm_100a only");
    RT_CUDA(cudaSetDevice(device));
    if (sm_count) *sm_count = sms;
    return RT_OK;
}

int fill_camera(const RtCamera &c, DCamera &d) {
    for (int i = 0; i < 16; ++i) { d.mvi[i] = c.modelview_inverse[i]; d.pi[i] = c.projection_inverse[i]; }
    d.depth_near = c.depth_near;
    // cameraSpaceToWorldSpace(Vec3(0,0,0)) (matrixUtilities.h:53-58): MV^-1 * (0,0,0,1), dehomogenised in fp64
    const double *m = c.modelview_inverse;
    double r[4];
    r[0] = m[0] * 0.0 + m[4] * 0.0 + m[8] * 0.0 + m[12] * 1.0;
    r[1] = m[1] * 0.0 + m[5] * 0.0 + m[9] * 0.0 + m[13] * 1.0;
    r[2] = m[2] * 0.0 + m[6] * 0.0 + m[10] * 0.0 + m[14] * 1.0;
    r[3] = m[3] * 0.0 + m[7] * 0.0 + m[11] * 0.0 + m[15] * 1.0;
    for (int k = 0; k < 3; ++k) d.pos[k] = (float)(r[k] / r[3]);
    return RT_OK;
}

int persistent_grid(RtScene *s, const void *kernel, int threads) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 4; }
    return s->sm_count * per_sm;
}

#define WF_F4_PLANES (10 + RT_LC_MAXC / 4)   /* ray0 ray1 hit0..hit4 park0 park1 + the candidate-triangle planes + mesh_hit */
#define WF_Q_PLANES 5   /* live, hit, parked, overflow, mesh-walk */
#define WF_MAX_BLOCK 256u
// Queue positions beyond the path count: every warp of a producing kernel may leave the rest of one block of each queue it
// appends to unused (wf_push), and a queue has at most three producing kernels per level (classify + the two sample passes).
size_t wf_padding(int sm_count) { return (size_t)3 * (size_t)sm_count * RT_WF_MINB * 4 * WF_MAX_BLOCK; }
int ensure_wavefront(Scratch *s, int sm_count, size_t paths, int max_bounces) {
    if (paths <= s->wf_cap && max_bounces <= s->wf_bounces) return RT_OK;
    paths = std::max(paths, s->wf_cap); max_bounces = std::max(max_bounces, s->wf_bounces);   // grow-only in both dimensions
    if (s->wf_f4) cudaFree(s->wf_f4);
    if (s->wf_rng) cudaFree(s->wf_rng);
    if (s->wf_q) cudaFree(s->wf_q);
    if (s->wf_rec) cudaFree(s->wf_rec);
    s->wf_f4 = nullptr; s->wf_rng = nullptr; s->wf_q = nullptr; s->wf_rec = nullptr; s->wf_cap = 0; s->wf_pos_cap = 0; s->wf_bounces = 0;
    const size_t pos = paths + wf_padding(sm_count);
    RT_CUDA(cudaMalloc((void **)&s->wf_f4, WF_F4_PLANES * pos * sizeof(float4)));
    RT_CUDA(cudaMalloc((void **)&s->wf_rng, 2 * pos * sizeof(uint2)));
    RT_CUDA(cudaMalloc((void **)&s->wf_q, WF_Q_PLANES * pos * sizeof(unsigned int)));
    RT_CUDA(cudaMalloc((void **)&s->wf_rec, 3 * (size_t)std::max(1, max_bounces) * paths * sizeof(float4)));
    if (!s->wf_ctr) RT_CUDA(cudaMalloc((void **)&s->wf_ctr, WF_NCTR * (RT_MAX_BOUNCES + 2) * sizeof(unsigned int)));
    s->wf_cap = paths; s->wf_pos_cap = pos; s->wf_bounces = max_bounces;
    return RT_OK;
}

// per-chunk buffers of one lane (the scene's own Scratch, or its `second`)
int ensure_lane(Scratch *s, size_t sample_floats, size_t cam_paths, size_t chunk_pixels) {
    if (cam_paths > s->cam_cap) {
        if (s->cam_rays) cudaFree(s->cam_rays);
        if (s->cam_keys) cudaFree(s->cam_keys);
        s->cam_rays = nullptr; s->cam_keys = nullptr; s->cam_cap = 0;
        RT_CUDA(cudaMalloc((void **)&s->cam_rays, cam_paths * sizeof(float4)));
        RT_CUDA(cudaMalloc((void **)&s->cam_keys, cam_paths * sizeof(unsigned int)));
        s->cam_cap = cam_paths;
    }
    if (cam_paths && chunk_pixels > s->pix_cap) {
        if (s->pix_xy) cudaFree(s->pix_xy);
        s->pix_xy = nullptr; s->pix_cap = 0;
        RT_CUDA(cudaMalloc((void **)&s->pix_xy, chunk_pixels * sizeof(unsigned int)));
        s->pix_cap = chunk_pixels;
    }
    if (sample_floats > s->samples_cap) {
        if (s->samples) cudaFree(s->samples);
        s->samples = nullptr; s->samples_cap = 0;
        RT_CUDA(cudaMalloc((void **)&s->samples, sample_floats * sizeof(float)));
        s->samples_cap = sample_floats;
    }
    return RT_OK;
}
int ensure_scratch(RtScene *s, size_t sample_floats, size_t n_tiles, size_t cam_paths, size_t chunk_pixels) {
    int rc = ensure_lane(s, sample_floats, cam_paths, chunk_pixels);
    if (rc) return rc;
    if (!s->counters) RT_CUDA(cudaMalloc((void **)&s->counters, 16 * sizeof(unsigned long long)));
    if (n_tiles + 1 > s->tiles_cap) {
        if (s->d_tiles) cudaFree(s->d_tiles);
        if (s->d_tile_off) cudaFree(s->d_tile_off);
        s->d_tiles = nullptr; s->d_tile_off = nullptr; s->tiles_cap = 0;
        RT_CUDA(cudaMalloc((void **)&s->d_tiles, (n_tiles + 1) * sizeof(TileRec)));
        RT_CUDA(cudaMalloc((void **)&s->d_tile_off, (n_tiles + 2) * sizeof(unsigned int)));
        s->tiles_cap = n_tiles + 1;
    }
    if (!s->ev0) { RT_CUDA(cudaEventCreate(&s->ev0)); RT_CUDA(cudaEventCreate(&s->ev1)); }
    return RT_OK;
}

int ensure_out(RtScene *s, size_t floats, size_t bytes) {
    if (floats > s->out_cap) {
        if (s->out_img) cudaFree(s->out_img);
        s->out_img = nullptr; s->out_cap = 0;
        RT_CUDA(cudaMalloc((void **)&s->out_img, floats * sizeof(float)));
        s->out_cap = floats;
    }
    if (bytes > s->out_bytes_cap) {
        if (s->out_bytes) cudaFree(s->out_bytes);
        s->out_bytes = nullptr; s->out_bytes_cap = 0;
        RT_CUDA(cudaMalloc((void **)&s->out_bytes, bytes));
        s->out_bytes_cap = bytes;
    }
    return RT_OK;
}

// Spheres, squares, lights and the culling hierarchy over the analytic primitives. update == false: first upload
// (arrays come from the scene's arena); update == true: the same arrays are rewritten in place.
template <class T> int dev_put(RtScene *s, bool update, const T *host, size_t n, const T **field) {
    if (!update) return dev_upload(s, host, n, field);
    if (n) RT_CUDA(h2d(s, const_cast<T *>(*field), host, n * sizeof(T)));
    return RT_OK;
}
int upload_analytic(RtScene *s, const RtSceneDesc *desc, bool update) {
    DScene &d = s->d;
    int rc;
    // spheres
    {
        std::vector<float4> a(desc->n_spheres), b(desc->n_spheres);
        std::vector<DMaterial> m(desc->n_spheres);
        for (uint32_t i = 0; i < desc->n_spheres; ++i) {
            const RtSphere &sp = desc->spheres[i];
            if ((rc = check_material(sp.material, *desc))) return rc;
            a[i] = make_float4(sp.center[0], sp.center[1], sp.center[2], sp.radius);
            b[i] = make_float4(sp.material.motion[0], sp.material.motion[1], sp.material.motion[2], sp.material.transparency);
            m[i] = to_dmat(sp.material);
        }
        if ((rc = dev_put(s, update, a.data(), a.size(), &d.sph_a))) return rc;
        if ((rc = dev_put(s, update, b.data(), b.size(), &d.sph_b))) return rc;
        if ((rc = dev_put(s, update, m.data(), m.size(), &d.sph_mat))) return rc;
    }
    // squares
    if (desc->n_squares) {
        std::vector<SquareIn> in(desc->n_squares);
        std::vector<DMaterial> m(desc->n_squares);
        std::vector<float> tr(desc->n_squares);
        for (uint32_t i = 0; i < desc->n_squares; ++i) {
            const RtSquare &q = desc->squares[i];
            if ((rc = check_material(q.material, *desc))) return rc;
            for (int k = 0; k < 3; ++k) { in[i].v0[k] = q.v0[k]; in[i].v1[k] = q.v1[k]; in[i].v3[k] = q.v3[k]; in[i].right[k] = q.right[k]; in[i].up[k] = q.up[k]; in[i].motion[k] = q.material.motion[k]; }
            in[i].glass = q.material.type == RT_MAT_GLASS;
            m[i] = to_dmat(q.material);
            tr[i] = q.material.transparency;
        }
        const SquareIn *d_in = s->d_square_in;
        DSquare *d_sq = const_cast<DSquare *>(d.squares);
        if ((rc = dev_put(s, update, in.data(), in.size(), &d_in))) return rc;
        s->d_square_in = const_cast<SquareIn *>(d_in);
        if (!update && (rc = dev_alloc(s, in.size(), &d_sq))) return rc;
        k_precompute_squares<<<(unsigned)((in.size() + 127) / 128), 128>>>(d_in, d_sq, (int)in.size());
        RT_CUDA(cudaGetLastError());
        d.squares = d_sq;
        if ((rc = dev_put(s, update, m.data(), m.size(), &d.sq_mat))) return rc;
        if ((rc = dev_put(s, update, tr.data(), tr.size(), &d.sq_transparency))) return rc;
    }
    // lights
    {
        std::vector<DLight> l(desc->n_lights);
        for (uint32_t i = 0; i < desc->n_lights; ++i) {
            for (int k = 0; k < 3; ++k) { l[i].pos[k] = desc->lights[i].pos[k]; l[i].color[k] = desc->lights[i].color[k]; }
            l[i].radius = desc->lights[i].radius;
        }
        if ((rc = dev_put(s, update, l.data(), l.size(), &d.lights))) return rc;
    }
    // culling hierarchy over the analytic primitives (variant 3)
    {
        AnalyticAccel aa;
        build_analytic_accel(*desc, aa);
        if (aa.nodes.size() / 4 > RT_ABVH_MAX_NODES) aa.root = -1;   // cannot happen for <= 128 primitives (n - 1 inner nodes); the staged copy relies on it
        d.abvh_root = aa.root;
        d.abvh_n_nodes = (int)(aa.nodes.size() / 4);
        for (int k = 0; k < 3; ++k) d.abvh_c[k] = aa.center[k];
        d.abvh_r = aa.radius;
        for (int k = 0; k < 3; ++k) d.abvh_cs[k] = aa.center_s[k];
        d.abvh_rs = aa.radius_s;
        if (!update) d.abvh_flat = nullptr;
        if (aa.root >= 0) {
            if (!update) {
                // a binary hierarchy over n primitives has at most n - 1 inner nodes (4 float4 each) and n leaf entries,
                // whatever its shape: room for every later update
                const size_t n = (size_t)desc->n_spheres + desc->n_squares;
                float4 *nodes = nullptr; uint32_t *prims = nullptr;
                s->abvh_node_cap = 4 * std::max<size_t>(1, n); s->abvh_prim_cap = std::max<size_t>(1, n);
                if ((rc = dev_alloc(s, s->abvh_node_cap, &nodes))) return rc;
                if ((rc = dev_alloc(s, s->abvh_prim_cap, &prims))) return rc;
                d.abvh_nodes = nodes; d.abvh_prims = prims;
            }
            if (aa.nodes.size() > s->abvh_node_cap || aa.tris.size() > s->abvh_prim_cap) return fail(RT_ERR_INVALID, "analytic hierarchy outgrew its arrays");
            RT_CUDA(h2d(s, const_cast<float4 *>(d.abvh_nodes), aa.nodes.data(), aa.nodes.size() * sizeof(float4)));
            RT_CUDA(h2d(s, const_cast<uint32_t *>(d.abvh_prims), aa.tris.data(), aa.tris.size() * sizeof(uint32_t)));
            if (!aa.flat.empty()) {   // same primitive counts on every update: allocated once
                if (!update) { float4 *fl = nullptr; if ((rc = dev_alloc(s, aa.flat.size(), &fl))) return rc; d.abvh_flat = fl; }
                if (d.abvh_flat) RT_CUDA(h2d(s, const_cast<float4 *>(d.abvh_flat), aa.flat.data(), aa.flat.size() * sizeof(float4)));
            }
        }
    }

    return RT_OK;
}

}  // namespace

extern "C" {

int rt_abi_version(void) { return HAI719_RT_ABI_VERSION; }

const char *rt_last_error(void) { return g_err.c_str(); }

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

void rt_scene_retain(RtScene *s) { if (s) s->refs.fetch_add(1); }

void rt_scene_destroy(RtScene *s) {
    if (!s) return;
    if (s->refs.fetch_sub(1) > 1) return;   // an accumulator or a retained handle still renders from these arrays
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();   // nothing of this scene may still be in flight when its arrays go
    img_cache_release(s);
    s->arena_reset();
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        g_scratch_pool[s->device].push_back(static_cast<Scratch &>(*s));
    }
    delete s;
}

int rt_release_cached_memory(int device) {
    std::vector<Scratch> idle;
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        auto it = g_scratch_pool.find(device);
        if (it != g_scratch_pool.end()) idle.swap(it->second);
    }
    RT_CUDA(cudaSetDevice(device));
    for (Scratch &sc : idle) sc.release();
    {   // images no scene references any more
        std::lock_guard<std::mutex> lock(g_img_mu);
        auto it = g_img_cache.find(device);
        if (it != g_img_cache.end()) img_cache_trim(it->second, 0);
    }
    {   // the tile tables rt_untile_device keeps per geometry on this device
        std::lock_guard<std::mutex> lock(g_untile_mu);
        for (auto it = g_untile.begin(); it != g_untile.end();) {
            if (it->first.device != device) { ++it; continue; }
            if (it->second.tiles) cudaFree(it->second.tiles);
            if (it->second.off) cudaFree(it->second.off);
            it = g_untile.erase(it);
        }
    }
    return RT_OK;
}

size_t rt_scene_device_bytes(const RtScene *s) { return s ? s->bytes : 0; }
size_t rt_scene_h2d_bytes(const RtScene *s) { return s ? s->h2d_bytes : 0; }

int rt_scene_update_analytic(RtScene *s, const RtSceneDesc *desc) {
    if (!s || !desc) return fail(RT_ERR_INVALID, "null argument");
    if (desc->abi_version != HAI719_RT_ABI_VERSION) return fail(RT_ERR_INVALID, "RtSceneDesc.abi_version mismatch");
    if ((int)desc->n_spheres != s->d.n_spheres || (int)desc->n_squares != s->d.n_squares || (int)desc->n_lights != s->d.n_lights ||
        (int)desc->n_meshes != s->d.n_meshes || desc->n_textures != s->n_textures || desc->n_normal_maps != s->n_normal_maps)
        return fail(RT_ERR_INVALID, "rt_scene_update_analytic: primitive / light / mesh / image counts differ from the uploaded scene");
    if ((desc->n_spheres && !desc->spheres) || (desc->n_squares && !desc->squares) || (desc->n_lights && !desc->lights))
        return fail(RT_ERR_INVALID, "count > 0 with a null array");
    RT_CUDA(cudaSetDevice(s->device));
    RT_CUDA(cudaDeviceSynchronize());   // no render of the old primitives may still be in flight
    s->stage_used = 0;
    const bool had_accel = s->d.abvh_root >= 0;
    int rc = upload_analytic(s, desc, true);
    if (rc) return rc;
    if (had_accel != (s->d.abvh_root >= 0)) return fail(RT_ERR_INVALID, "rt_scene_update_analytic: hierarchy appeared or vanished");
    s->d.dark_sky = desc->dark_sky;
    RT_CUDA(cudaDeviceSynchronize());
    return RT_OK;
}

int rt_scene_check(const RtSceneDesc *desc) {
    if (!desc) return fail(RT_ERR_INVALID, "null argument");
    if (desc->abi_version != HAI719_RT_ABI_VERSION) return fail(RT_ERR_INVALID, "RtSceneDesc.abi_version mismatch");
    if ((desc->n_spheres && !desc->spheres) || (desc->n_squares && !desc->squares) || (desc->n_meshes && !desc->meshes) ||
        (desc->n_lights && !desc->lights) || (desc->n_textures && !desc->textures) || (desc->n_normal_maps && !desc->normal_maps))
        return fail(RT_ERR_INVALID, "count > 0 with a null array");
    int rc;
    for (uint32_t i = 0; i < desc->n_spheres; ++i) if ((rc = check_material(desc->spheres[i].material, *desc))) return rc;
    for (uint32_t i = 0; i < desc->n_squares; ++i) if ((rc = check_material(desc->squares[i].material, *desc))) return rc;
    for (uint32_t i = 0; i < desc->n_meshes; ++i) if ((rc = check_material(desc->meshes[i].material, *desc))) return rc;
    if (desc->n_meshes) {
        PackedMeshes pk;
        const std::string why = pack_meshes(*desc, pk);
        if (!why.empty()) return fail(RT_ERR_INVALID, why);
    }
    return RT_OK;
}

int rt_scene_create(const RtSceneDesc *desc, int device, RtScene **out) {
    if (!desc || !out) return fail(RT_ERR_INVALID, "null argument");
    *out = nullptr;
    if (desc->abi_version != HAI719_RT_ABI_VERSION) return fail(RT_ERR_INVALID, "RtSceneDesc.abi_version mismatch");
    if ((desc->n_spheres && !desc->spheres) || (desc->n_squares && !desc->squares) || (desc->n_meshes && !desc->meshes) ||
        (desc->n_lights && !desc->lights) || (desc->n_textures && !desc->textures) || (desc->n_normal_maps && !desc->normal_maps))
        return fail(RT_ERR_INVALID, "count > 0 with a null array");
    int sm_count = 0;
    int rc = select_device(device, &sm_count);
    if (rc) return rc;

    RtScene *s = new RtScene;
    s->device = device;
    s->sm_count = sm_count;
    {
        std::lock_guard<std::mutex> lock(g_scratch_mu);
        std::vector<Scratch> &pool = g_scratch_pool[device];
        if (!pool.empty()) { static_cast<Scratch &>(*s) = pool.back(); pool.pop_back(); }
    }
    struct Guard { RtScene *s; bool keep = false; ~Guard() { if (!keep) rt_scene_destroy(s); } } guard{s};
    s->stage_used = 0;   // the copies of the scene that used this staging buffer before completed when its upload returned
    DScene &d = s->d;
    d.n_spheres = (int)desc->n_spheres; d.n_squares = (int)desc->n_squares; d.n_meshes = (int)desc->n_meshes; d.n_lights = (int)desc->n_lights;
    d.dark_sky = desc->dark_sky;
    s->n_textures = desc->n_textures; s->n_normal_maps = desc->n_normal_maps;

    // images
    std::vector<DImage> tex(desc->n_textures), nrm(desc->n_normal_maps);
    for (uint32_t i = 0; i < desc->n_textures; ++i) if ((rc = upload_image(s, desc->textures[i], tex[i]))) return rc;
    for (uint32_t i = 0; i < desc->n_normal_maps; ++i) if ((rc = upload_image(s, desc->normal_maps[i], nrm[i]))) return rc;
    if ((rc = dev_upload(s, tex.data(), tex.size(), &d.textures))) return rc;
    if ((rc = dev_upload(s, nrm.data(), nrm.size(), &d.normal_maps))) return rc;
    if ((rc = upload_image(s, desc->skybox, d.sky))) return rc;

    if ((rc = upload_analytic(s, desc, false))) return rc;
    // meshes: one shared node array + shared per-ref triangle constants (rt_pack.hpp, rt::DMesh)
    if (desc->n_meshes) {
        PackedMeshes pk;
        const std::string why = pack_meshes(*desc, pk);
        if (!why.empty()) return fail(RT_ERR_INVALID, why);
        if ((rc = dev_upload(s, pk.lo.data(), pk.lo.size(), &d.node_lo))) return rc;
        if ((rc = dev_upload(s, pk.hi.data(), pk.hi.size(), &d.node_hi))) return rc;
        float4 *pl = nullptr, *ed = nullptr; float2 *dn = nullptr;
        if ((rc = dev_alloc(s, (size_t)pk.total_refs, &pl))) return rc;
        if ((rc = dev_alloc(s, (size_t)3 * pk.total_refs, &ed))) return rc;
        if ((rc = dev_alloc(s, (size_t)pk.total_refs, &dn))) return rc;
        d.tri_plane = pl; d.tri_edge = ed; d.tri_den = dn;
        Accel ac;
        build_accel(*desc, pk, ac);
        if ((rc = dev_upload(s, ac.nodes.data(), ac.nodes.size(), &d.bvh_nodes))) return rc;
        if ((rc = dev_upload(s, ac.nodes4.data(), ac.nodes4.size(), &d.bvh4_nodes))) return rc;
        if ((rc = dev_upload(s, ac.tris.data(), ac.tris.size(), &d.bvh_tris))) return rc;
        float4 *ab = nullptr;
        if (!ac.tris.empty()) {
            if ((rc = dev_alloc(s, (size_t)3 * ac.tris.size(), &ab))) return rc;
            RT_CUDA(cudaMemsetAsync(ab, 0, (size_t)3 * ac.tris.size() * sizeof(float4)));
        }
        d.always_bound = ab;
        if ((rc = dev_upload(s, ac.ref_next.data(), ac.ref_next.size(), &d.ref_next))) return rc;
        if ((rc = dev_upload(s, ac.ref_leaf.data(), ac.ref_leaf.size(), &d.ref_leaf))) return rc;
        if ((rc = dev_upload(s, ac.node_parent.data(), ac.node_parent.size(), &d.node_parent))) return rc;
        std::vector<DMesh> dm(desc->n_meshes);
        std::vector<DMaterial> m(desc->n_meshes);
        std::vector<float> tr(desc->n_meshes);
        for (uint32_t i = 0; i < desc->n_meshes; ++i) {
            const RtSceneMesh &src = desc->meshes[i];
            if ((rc = check_material(src.material, *desc))) return rc;
            DMesh &o = dm[i];
            memset(&o, 0, sizeof o);
            o.node_begin = pk.node_begin[i];
            o.node_end = pk.node_end[i];
            o.bvh_root = ac.mesh_root[i]; o.bvh4_root = ac.mesh_root4[i]; o.always_first = ac.always_first[i]; o.always_count = ac.always_count[i];
            o.color_type = src.color_type;
            if ((rc = dev_upload(s, src.triangles, (size_t)3 * src.n_triangles, &o.triangles))) return rc;
            if (src.color_type == RT_COLOR_VERTEX && (rc = dev_upload(s, src.vert_colors, (size_t)3 * src.n_vertices, &o.vert_colors))) return rc;
            if (src.color_type == RT_COLOR_FACE && (rc = dev_upload(s, src.face_colors, (size_t)3 * src.n_triangles, &o.face_colors))) return rc;
            if (src.n_leaf_refs) {
                DevTmp d_pos, d_refs;
                RT_CUDA(d_pos.put(s, src.positions, (size_t)3 * src.n_vertices * sizeof(float)));
                RT_CUDA(d_refs.put(s, src.leaf_refs, (size_t)src.n_leaf_refs * sizeof(RtTriRef)));
                const uint32_t rb = pk.ref_begin[i];
                k_precompute_tris<<<(src.n_leaf_refs + 127) / 128, 128>>>((const float *)d_pos.p, (const RtTriRef *)d_refs.p, src.n_leaf_refs,
                                                                         pl + rb, ed + 3 * (size_t)rb, dn + rb);
                RT_CUDA(cudaGetLastError());   // the staging arrays live in the scene's arena: no need to wait for the kernel here
            }
            if (o.always_count > 0u && ab) {
                k_always_bounds<<<(o.always_count + 63) / 64, 64>>>(d.bvh_tris, o.always_first, o.always_count, ed, dn, ab);
                RT_CUDA(cudaGetLastError());
            }
            m[i] = to_dmat(src.material);
            tr[i] = src.material.transparency;
        }
        if ((rc = dev_upload(s, dm.data(), dm.size(), &d.meshes))) return rc;
        if ((rc = dev_upload(s, m.data(), m.size(), &d.mesh_mat))) return rc;
        if ((rc = dev_upload(s, tr.data(), tr.size(), &d.mesh_transparency))) return rc;
    }
    RT_CUDA(cudaDeviceSynchronize());
    guard.keep = true;
    *out = s;
    return RT_OK;
}

int64_t rt_render_pixel_count(const RtRenderParams *p) {
    if (!p) return -1;
    Rect r;
    if (resolve_rect(*p, r)) return -1;
    std::vector<TileRec> tiles; std::vector<unsigned int> off;
    build_tiles(*p, r, tiles, off);
    return (int64_t)off.back();
}

int64_t rt_tile_layout(const RtRenderParams *p, int32_t *out, int64_t cap) {
    if (!p) return -1;
    Rect r;
    if (resolve_rect(*p, r)) return -1;
    std::vector<TileRec> tiles; std::vector<unsigned int> off;
    build_tiles(*p, r, tiles, off);
    if (out)
        for (int64_t i = 0; i < (int64_t)tiles.size() && i < cap; ++i) {
            out[4 * i] = tiles[i].x0; out[4 * i + 1] = tiles[i].y0; out[4 * i + 2] = tiles[i].w; out[4 * i + 3] = tiles[i].h;
        }
    return (int64_t)tiles.size();
}

// The render behind rt_render_device (sample_base 0, no running sum) and rt_accum_add (samples sample_base ..
// sample_base + spp - 1 of every pixel, continuing the per-pixel sums in d_sum).
static int render_device_impl(RtScene *s, const RtCamera *camera, const RtRenderParams *p, float *d_gamma, float *d_linear,
                              void *cuda_stream, RtStats *stats, unsigned int sample_base, float *d_sum, int image_mode = 0) {
    if (!s || !camera || !p) return fail(RT_ERR_INVALID, "null argument");
    Rect r;
    int rc = resolve_rect(*p, r);
    if (rc) return rc;
    RT_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    build_tiles(*p, r, s->h_tiles, s->h_tile_off);
    const size_t n_tiles = s->h_tiles.size();
    const unsigned long long n_pixels = s->h_tile_off.back();
    if (stats) { memset(stats, 0, sizeof *stats); stats->n_tiles = (uint32_t)n_tiles; stats->n_samples = n_pixels * (unsigned long long)p->spp; }
    if (n_pixels == 0) return RT_OK;

    // chunking: whole pixels, at most ~16 Mi paths in flight (192 MiB of samples)
    // Kernel selection (measured on B200, profiles/r01_notes.md): the wavefront (6) wins wherever the analytic
    // culling hierarchy exists (configs 2, 4, 5: 1.2-1.3x over the state machine; config 3 since the end of round 2); mesh scenes
    // without it (no lights and fewer than 24 analytic primitives) run the single state-machine kernel over the exact culling
    // hierarchies (3); scenes with a handful of analytic primitives and nothing else (config 1) the plain one-path-per-lane kernel (1).
    int kind_req = p->variant & 0xFF;
    // (Until the end of round 2 lit mesh scenes with fewer than 24 analytic primitives — the pond scene — stayed on the state machine: the two
    // were level at 2 spp. At the pond scene's full 16 spp the wavefront, with its kernels compiled per mode and the any-hit overflow samples, takes
    // 350 ms against 442: adjacent samples of a pixel make coherent batches. profiles/r02_notes.md, r04j.)
    // By size (profiles/r04n_*, kernel ms, state machine / wavefront): 2 spp = 16.6 M paths 60.5 / 59.7, 4 spp 118.8 / 99.0, 8 spp 227 / 188.5, 16 spp 442 / 350 — and
    // on 8 GPUs, where a rank renders 16.6 M paths of the 16-spp frame, 59.5 / 64.2 (max over ranks, r04d / r04m): with fewer than ~24 M paths in a
    // call the deeper bounces hold a handful of batches per warp and the wavefront's kernels end unevenly. Such calls keep the state machine.
    if (kind_req == 0 && s->d.abvh_root >= 0 &&
        (s->d.n_spheres + s->d.n_squares >= 24 || n_pixels * (unsigned long long)p->spp >= (24ull << 20))) kind_req = 6;
    const bool wavefront = kind_req == 6 && p->max_bounces > 0;
    // paths per chunk. Wavefront: every kernel of a chunk ends in a tail during which SMs drain, so fewer, larger chunks
    // are faster (config 2, ms per frame at 32 spp: 4 Mi 36.5, 8 Mi 32.9, 16 Mi 31.3, 32 Mi 30.3); 32 Mi paths are ~17 GB
    // of path state at 6 bounces (~230 B + 48 B per bounce and path), a tenth of the HBM. Scenes without meshes take 64 Mi (config 2 at 64 spp: 39.8 -> 38.4 ms; the mesh scenes did not
    // gain: config 4 122.8 -> 124.6, config 5 116.1 -> 114.6, profiles/r02_notes.md). Chunks are EQUAL: a frame that needs k chunks
    // gets k chunks of n / k pixels, not k - 1 full ones and a small rest whose kernels are mostly tail.
    unsigned long long max_paths = wavefront ? ((s->d.n_meshes == 0 ? 64ull : 32ull) << 20) : (16ull << 20);
    if (const char *e = getenv("HAI719_CHUNK_LOG2")) { const int l = atoi(e); if (l >= 16 && l <= 26) max_paths = 1ull << l; }   // tuning experiments
    // The wavefront renders the chunks of a frame on TWO streams (lanes) at once, each with its own per-chunk buffers: every
    // one of a chunk's ~19 persistent kernels ends in a tail of 40-60 us during which SMs drain (one batch of 32 paths takes
    // that long through the kernel; measured as the intercept of kernel time against chunk size, profiles/r02_notes.md), and
    // the other lane's kernels fill it. The paths in flight stay what they were: a lane's chunk is half the budget.
    // Measured (profiles/r02_notes.md, r03a): config 2 38.5 -> 38.4 ms, config 4 121.5 -> 120.0, config 5 115.1 -> 114.0 — within
    // 1 %, for twice the per-chunk buffers. OFF by default; HAI719_LANES=2 switches it on (A/B).
    int n_lanes = 1;
    if (const char *e = getenv("HAI719_LANES")) { if (atoi(e) == 2 && wavefront && n_pixels * (unsigned long long)p->spp >= (2ull << 20)) n_lanes = 2; }
    unsigned long long chunk_pixels = std::max<unsigned long long>(1, (max_paths / (unsigned long long)n_lanes) / (unsigned long long)p->spp);
    chunk_pixels = std::min(chunk_pixels, n_pixels);
    {
        unsigned long long n_chunks = (n_pixels + chunk_pixels - 1) / chunk_pixels;
        if (n_lanes == 2) n_chunks += n_chunks & 1ull;   // both lanes get the same number of chunks
        chunk_pixels = (n_pixels + n_chunks - 1) / n_chunks;
    }
    // bit 28 of variant: generate camera rays inside the render kernel instead of the k_camera_rays pass (A/B switch)
    const bool cam_split = wavefront || ((p->variant >> 28) & 1) == 0 && (p->variant & 0xFF) != 1 && !((p->variant & 0xFF) == 0 && s->d.n_meshes == 0 && s->d.abvh_root < 0);
    if ((rc = ensure_scratch(s, (size_t)(chunk_pixels * p->spp * 3ull), n_tiles, cam_split ? (size_t)(chunk_pixels * p->spp) : 0, (size_t)chunk_pixels))) return rc;
    if (n_lanes == 2) {
        if (!s->second) s->second = new Scratch();
        if ((rc = ensure_lane(s->second, (size_t)(chunk_pixels * p->spp * 3ull), cam_split ? (size_t)(chunk_pixels * p->spp) : 0, (size_t)chunk_pixels))) return rc;
        if (!s->aux_stream) RT_CUDA(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
        if (!s->ev_fork) { RT_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming)); RT_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming)); }
    }
    RT_CUDA(cudaMemcpyAsync(s->d_tiles, s->h_tiles.data(), n_tiles * sizeof(TileRec), cudaMemcpyHostToDevice, st));
    RT_CUDA(cudaMemcpyAsync(s->d_tile_off, s->h_tile_off.data(), (n_tiles + 1) * sizeof(unsigned int), cudaMemcpyHostToDevice, st));

    DCamera cam;
    fill_camera(*camera, cam);
    const bool want_stats = p->collect_stats != 0 && stats != nullptr;
    // variant: low byte = kernel (0 auto, 1 k_render_paths: one path per lane to completion,
    // 2 k_render_regen: ray-level state machine with path regeneration and warp-voted KD traversal);
    // bits 8..15 = regeneration threshold of kernel 2 (idle lanes needed before a refill; 0 = 16)
    if (p->variant < 0 || (p->variant & 0xFF) > 6 || (p->variant >> 30)) return fail(RT_ERR_INVALID, "unknown kernel variant");
    // bits 16..19: CTAs per SM of kernel 3 — 0 auto, 1 = 4 (<= 128 registers), 2 = 6 (<= 80), 3 = 8 (<= 64, a few spills)
    int occ = (p->variant >> 16) & 0xF;
    if (occ == 0) occ = 3;   // 8 CTAs/SM beat 6 and 4 on every config (profiles/r01_notes.md)
    const int minb = occ - 1;
    int kind = kind_req;
    // 5 = occluder candidates per (hit, light): needs the analytic hierarchy and at least one light
    const bool lc_ok = s->d.abvh_root >= 0 && s->d.n_lights > 0;
    if (kind == 6 && !wavefront) kind = 0;   // rayTraceRecursive(ray, 0) / 0: no bounce level to run
    // automatic choice below the wavefront's threshold (measured, profiles/r01_notes.md): the state machine over the exact
    // culling hierarchies (3) for anything with meshes - the pond scene: 73 ms against 78 for the light-cone kernel (5),
    // which is only run when asked for - and one path per lane (1) for a handful of analytic primitives
    if (kind == 0) kind = (s->d.n_meshes > 0 || s->d.abvh_root >= 0) ? 3 : 1;
    if (kind == 5 && !lc_ok) kind = 3;
    const bool regen = kind >= 2, accel = kind >= 3, voted = kind == 4, lc = kind == 5;
    typedef void (*RenderKernel)(const DScene, const DCamera, const RenderArgs);
    RenderKernel fn;
    if (lc)         fn = want_stats ? k_render_regen<true, 3, 4> : minb == 0 ? k_render_regen<false, 3, 4> : minb == 1 ? k_render_regen<false, 3, 6> : k_render_regen<false, 3, 8>;
    else if (voted)      fn = want_stats ? k_render_regen<true, 2, 4> : minb == 0 ? k_render_regen<false, 2, 4> : minb == 1 ? k_render_regen<false, 2, 6> : k_render_regen<false, 2, 8>;
    else if (accel) fn = want_stats ? k_render_regen<true, 1, 4> : minb == 0 ? k_render_regen<false, 1, 4> : minb == 1 ? k_render_regen<false, 1, 6> : k_render_regen<false, 1, 8>;
    else if (regen) fn = want_stats ? k_render_regen<true, 0, 4> : k_render_regen<false, 0, 4>;
    else            fn = want_stats ? k_render_paths<true> : k_render_paths<false>;
    const void *kern = (const void *)fn;
    const int grid = persistent_grid(s, kern, 128);
    // wavefront kernels (variant 6): LC = cone candidates per light, possible whenever the analytic hierarchy exists
    typedef void (*TraceKernel)(const DScene, const DCamera, const WfArgs);
    typedef void (*LightKernel)(const DScene, const WfArgs);
    const bool wf_lc = s->d.abvh_root >= 0;
    // a scene without lights has no light stage: the trace kernel can scatter itself (NOLIGHT), one kernel per bounce level.
    // Variant bit 29 asks for it; measured slower (profiles/r02_notes.md), so it is not the default.
    const bool wf_nolight = s->d.n_lights == 0 && ((p->variant >> 29) & 1) != 0;
    // Scenes with meshes: analytic phase, then the mesh walk for the rays that touch a mesh, in a kernel of its own
    // (k_wf_trace<.., MESH>). Measured (profiles/r02_notes.md, r02k-r02m): the pool scene (no lights: the trace stage is 80 % of
    // the frame) 64.6 -> 61.2 ms per 4-spp frame; config 5 62.5 -> 65.2 and the pond scene 67.0 -> 68.5 (lit scenes: trace is a
    // quarter of the frame, and nearly every ray touches a mesh's root box there, so the second kernel repeats the pass over the
    // ray records for all of them). Hence: on for scenes without lights; wavefront variant bit 28 flips the choice (A/B).
    const bool wf_split = wf_lc && !wf_nolight && s->d.n_meshes > 0 && ((s->d.n_lights == 0) != (((p->variant >> 28) & 1) != 0));
    TraceKernel wf_trace = want_stats ? (wf_lc ? (wf_nolight ? k_wf_trace<true, true, true, 0> : wf_split ? k_wf_trace<true, true, false, 1> : k_wf_trace<true, true, false, 0>)
                                               : (wf_nolight ? k_wf_trace<true, false, true, 0> : k_wf_trace<true, false, false, 0>))
                                      : (wf_lc ? (wf_nolight ? k_wf_trace<false, true, true, 0> : wf_split ? k_wf_trace<false, true, false, 1> : k_wf_trace<false, true, false, 0>)
                                               : (wf_nolight ? k_wf_trace<false, false, true, 0> : k_wf_trace<false, false, false, 0>));
    if (wf_lc && !wf_nolight && s->d.n_meshes == 0 && RT_OPT_LC_COLLECT) wf_trace = want_stats ? k_wf_trace<true, true, false, 3> : k_wf_trace<false, true, false, 3>;
    TraceKernel wf_trace_mesh = want_stats ? k_wf_trace<true, true, false, 2> : k_wf_trace<false, true, false, 2>;
    // light stage: one kernel (no masks to classify by) or walk/classify + sample (see k_wf_light)
    LightKernel wf_light = want_stats ? (wf_lc ? k_wf_light<true, true, 1> : k_wf_light<true, false, 0>) : (wf_lc ? k_wf_light<false, true, 1> : k_wf_light<false, false, 0>);
    LightKernel wf_light_b = want_stats ? k_wf_light<true, true, 2> : k_wf_light<false, true, 2>;
    if (s->d.n_lights == 1 && RT_OPT_LC_COLLECT) wf_light_b = want_stats ? k_wf_light<true, true, 3> : k_wf_light<false, true, 3>;
    LightKernel wf_light_over = wf_light_b;   // the overflow queue's launch
    if (s->d.n_lights == 1 && s->d.n_meshes > 0 && RT_OPT_LC_COLLECT) wf_light_over = want_stats ? k_wf_light<true, true, 3, false, 1> : k_wf_light<false, true, 3, false, 1>;
    if (wf_lc && s->d.n_meshes == 0 && RT_OPT_LC_COLLECT) {
        wf_light = want_stats ? k_wf_light<true, true, 1, true> : k_wf_light<false, true, 1, true>;
        if (s->d.n_lights == 1) wf_light_b = want_stats ? k_wf_light<true, true, 3, true> : k_wf_light<false, true, 3, true>;
    }
    // no lights: the light stage is only the scatter (k_wf_scatter); wavefront variant bit 27 keeps the general kernel (A/B)
    const bool wf_scatter_only = s->d.n_lights == 0 && !wf_nolight && ((p->variant >> 27) & 1) == 0;
    if (wf_scatter_only) wf_light = want_stats ? k_wf_scatter<true> : k_wf_scatter<false>;
    int wf_grid_t = 0, wf_grid_l = 0;
    if (wavefront) {
        if ((rc = ensure_wavefront(s, s->sm_count, (size_t)(chunk_pixels * p->spp), p->max_bounces))) return rc;
        if (n_lanes == 2 && (rc = ensure_wavefront(s->second, s->sm_count, (size_t)(chunk_pixels * p->spp), p->max_bounces))) return rc;
        wf_grid_t = persistent_grid(s, (const void *)wf_trace, 128);
        wf_grid_l = persistent_grid(s, (const void *)wf_light, 128);
    }

    RenderArgs a{};
    a.tiles = s->d_tiles; a.tile_off = s->d_tile_off; a.n_tiles = (int)n_tiles;
    a.width = p->width; a.height = p->height; a.spp = p->spp; a.max_bounces = p->max_bounces; a.nb_ech = p->nb_ech;
    a.regen_min = ((p->variant >> 8) & 0xFF) ? std::min(32, (p->variant >> 8) & 0xFF) : 16;
    // bits 20..27: traversal threshold of kernel 5. Measured (profiles/r01_notes.md): 1 is best — traverse as soon as
    // any lane wants to; the lanes then catch up with the ones sampling shadows and the warp falls into cohorts by itself
    a.t_min = ((p->variant >> 20) & 0xFF) ? std::min(32, (p->variant >> 20) & 0xFF) : 1;
    a.seed = p->seed; a.samples = s->samples; a.work_counter = s->counters; a.stats = want_stats ? s->counters + 1 : nullptr;
    a.sample_base = sample_base;
    uint32_t launches = 0;
    SceneBinding binding;
    if ((rc = binding.bind(s, st))) return rc;
    RT_CUDA(cudaMemsetAsync(s->counters, 0, 16 * sizeof(unsigned long long), st));
    if (stats) RT_CUDA(cudaEventRecord(s->ev0, st));
    cudaStream_t const st_caller = st;
    if (n_lanes == 2) {   // the second lane starts after everything queued on the caller's stream so far (tile table, constants)
        RT_CUDA(cudaEventRecord(s->ev_fork, st_caller));
        RT_CUDA(cudaStreamWaitEvent(s->aux_stream, s->ev_fork, 0));
    }
    unsigned long long chunk_index = 0;
    for (unsigned long long pb = 0; pb < n_pixels; pb += chunk_pixels, ++chunk_index) {
        const unsigned long long np = std::min(chunk_pixels, n_pixels - pb);
        const int lane = n_lanes == 2 ? (int)(chunk_index & 1ull) : 0;
        Scratch *const L = lane ? s->second : static_cast<Scratch *>(s);   // this chunk's buffers ...
        st = lane ? s->aux_stream : st_caller;                             // ... and stream
        a.samples = L->samples;
        a.pixel_begin = pb;
        a.n_paths = np * (unsigned long long)p->spp;
        if (pb && !wavefront) RT_CUDA(cudaMemsetAsync(s->counters, 0, sizeof(unsigned long long), st));
        const unsigned long long batches = (a.n_paths + 31) / 32;
        const int g = (int)std::min<unsigned long long>((unsigned long long)grid, (batches + 3) / 4);
        if (cam_split) {
            a.cam_rays = L->cam_rays; a.cam_keys = L->cam_keys;
            k_pixel_xy<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(a, (unsigned int)np, L->pix_xy);
            RT_CUDA(cudaGetLastError());
            k_camera_rays<<<(unsigned)((a.n_paths + 255) / 256), 256, 0, st>>>(cam, a, L->pix_xy, L->cam_rays, L->cam_keys);
            ++launches;
            RT_CUDA(cudaGetLastError());
            ++launches;
        }
        if (wavefront) {
            WfArgs w{};
            const size_t pc = L->wf_pos_cap;
            w.n_paths = (unsigned int)a.n_paths; w.max_bounces = p->max_bounces; w.nb_ech = p->nb_ech;
            // mesh scenes: a batch can cost 100x another one (rays that walk a mesh vs rays that miss its box), so warps take
            // one batch at a time as before; analytic scenes: cheap, even batches, where the fetch atomics were the bottleneck
            w.max_grab = s->d.n_meshes > 0 ? 1u : 8u;
            // windows of the live queue sorted by ray direction from bounce 1 on (wf_next_batch_sorted); HAI719_WF_SORT=0 / 1 overrides (A/B)
            w.flat = s->d.abvh_flat ? 2 : 0;   // from bounce 1 on; HAI719_WF_FLAT=0 (never) / 1 (from bounce 0) / 2 overrides (A/B)
            if (const char *e = getenv("HAI719_WF_FLAT")) w.flat = s->d.abvh_flat ? atoi(e) : 0;
            w.sort = RT_WF_SORT && w.max_grab > 1u ? 1 : 0;
            if (const char *e = getenv("HAI719_WF_SORT")) w.sort = RT_WF_SORT && atoi(e) != 0 ? 1 : 0;
            if (const char *e = getenv("HAI719_WF_GRAB")) { const int g_ = atoi(e); if (g_ >= 1 && g_ <= 8) w.max_grab = (unsigned int)g_; }
            // output positions a warp reserves per atomic: 256 where the queue atomics were the bottleneck (large chunks);
            // small chunks take small blocks, so that their queues are not mostly padding
            w.block = a.n_paths >= (4ull << 20) ? WF_MAX_BLOCK : (a.n_paths >= (256ull << 10) ? 64u : 32u);
            w.cam_rays = L->cam_rays; w.cam_keys = L->cam_keys;
            float4 *const f4 = L->wf_f4;
            float4 *rayA0 = f4, *rayA1 = f4 + pc;
            w.hit0 = f4 + 2 * pc; w.hit1 = f4 + 3 * pc; w.hit2 = f4 + 4 * pc; w.hit3 = f4 + 5 * pc; w.hit4 = f4 + 6 * pc;
            w.park0 = f4 + 7 * pc; w.park1 = f4 + 8 * pc; w.park2 = f4 + 9 * pc; w.park_stride = pc;
            w.mesh_hit = f4 + (9 + RT_LC_MAXC / 4) * pc; w.q_mesh = L->wf_q + 4 * pc;
            uint2 *rngA = L->wf_rng;
            w.rng_hit = L->wf_rng + pc;
            unsigned int *qA = L->wf_q;
            w.q_hit = L->wf_q + pc; w.q_park = L->wf_q + 2 * pc; w.q_over = L->wf_q + 3 * pc; w.which_park = 0;
            w.rec = L->wf_rec; w.rec_stride = L->wf_cap;
            w.ctr = L->wf_ctr; w.samples = L->samples; w.stats = a.stats;
            RT_CUDA(cudaMemsetAsync(L->wf_ctr, 0, WF_NCTR * (RT_MAX_BOUNCES + 2) * sizeof(unsigned int), st));
            const int gt = (int)std::min<unsigned long long>((unsigned long long)wf_grid_t, (batches + 3) / 4);
            const int gl = (int)std::min<unsigned long long>((unsigned long long)wf_grid_l, (batches + 3) / 4);
            for (int level = 0; level < p->max_bounces; ++level) {
                w.level = level;
                if (wf_nolight) {
                    // no hit queue: its arrays are the second live queue, the two alternate by level (a level's kernel reads
                    // one and appends to the other)
                    const bool odd = (level & 1) != 0;
                    w.q_live = odd ? w.q_hit : qA; w.ray0 = odd ? w.hit0 : rayA0; w.ray1 = odd ? w.hit1 : rayA1; w.rng_ray = odd ? w.rng_hit : rngA;
                    w.q_next = odd ? qA : w.q_hit; w.nray0 = odd ? rayA0 : w.hit0; w.nray1 = odd ? rayA1 : w.hit1; w.nrng = odd ? rngA : w.rng_hit;
                    wf_trace<<<gt, 128, 0, st>>>(s->d, cam, w);
                    RT_CUDA(cudaGetLastError());
                    ++launches;
                    continue;
                }
                // the light kernels of level L append the live queue of level L + 1 to the arrays trace L has finished reading
                w.q_live = qA; w.ray0 = rayA0; w.ray1 = rayA1; w.rng_ray = rngA;
                w.q_next = qA; w.nray0 = rayA0; w.nray1 = rayA1; w.nrng = rngA;
                wf_trace<<<gt, 128, 0, st>>>(s->d, cam, w);
                RT_CUDA(cudaGetLastError());
                if (wf_split) {
                    wf_trace_mesh<<<gt, 128, 0, st>>>(s->d, cam, w);
                    RT_CUDA(cudaGetLastError());
                    ++launches;
                }
                wf_light<<<gl, 128, 0, st>>>(s->d, w);
                RT_CUDA(cudaGetLastError());
                launches += 2;
                if (wf_lc && !wf_scatter_only) {
                    w.which_park = 0;
                    wf_light_b<<<gl, 128, 0, st>>>(s->d, w);
                    RT_CUDA(cudaGetLastError());
                    ++launches;
                    if (s->d.n_meshes > 0) {   // lights whose candidate-triangle list overflowed, in warps of their own
                        w.which_park = 1;
                        wf_light_over<<<gl, 128, 0, st>>>(s->d, w);
                        RT_CUDA(cudaGetLastError());
                        ++launches;
                        w.which_park = 0;
                    }
                }
            }
        } else {
            fn<<<g, 128, 0, st>>>(s->d, cam, a);
            RT_CUDA(cudaGetLastError());
            ++launches;
        }
        if (wavefront) k_wf_tally<<<1, 32, 0, st>>>(L->wf_ctr, (unsigned int)a.n_paths, p->max_bounces, s->counters + 11);
        k_resolve<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(a, cam_split ? L->pix_xy : nullptr, (unsigned int)np, d_linear, d_gamma, d_sum, sample_base,
                                                                 image_mode, r.x0, r.y0, r.x1 - r.x0);
        RT_CUDA(cudaGetLastError());
        ++launches;
    }
    st = st_caller;
    if (n_lanes == 2) {   // the caller's stream continues when the second lane is done
        RT_CUDA(cudaEventRecord(s->ev_join, s->aux_stream));
        RT_CUDA(cudaStreamWaitEvent(st_caller, s->ev_join, 0));
    }
    binding.done();
    if (stats) {
        RT_CUDA(cudaEventRecord(s->ev1, st));
        RT_CUDA(cudaEventSynchronize(s->ev1));
        float ms = 0.f;
        RT_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        stats->kernel_ms = ms;
        stats->n_launches = launches;
        stats->n_chunks = (uint32_t)((n_pixels + chunk_pixels - 1) / chunk_pixels);
        if (wavefront && !want_stats) {   // ray counts from the queue counters (k_wf_tally): exact, and free
            unsigned long long c[2];
            RT_CUDA(cudaMemcpy(c, s->counters + 11, sizeof c, cudaMemcpyDeviceToHost));
            stats->n_closest_rays = c[0];
            stats->n_shadow_rays = c[1] * (unsigned long long)s->d.n_lights * (unsigned long long)p->nb_ech;
        }
        if (want_stats) {
            unsigned long long c[11];
            RT_CUDA(cudaMemcpy(c, s->counters, sizeof c, cudaMemcpyDeviceToHost));
            stats->n_closest_rays = c[1]; stats->n_shadow_rays = c[2]; stats->n_sphere_tests = c[3]; stats->n_square_tests = c[4];
            stats->n_mesh_tests = c[5]; stats->n_node_visits = c[6]; stats->n_tri_tests = c[7]; stats->n_tri_full = c[8];
            stats->n_tex_fetches = c[9]; stats->n_random = c[10];
        }
    }
    return RT_OK;
}

int rt_render_device(RtScene *s, const RtCamera *camera, const RtRenderParams *p, float *d_gamma, float *d_linear,
                     void *cuda_stream, RtStats *stats) {
    return render_device_impl(s, camera, p, d_gamma, d_linear, cuda_stream, stats, 0u, nullptr);
}

int rt_render_device_image(RtScene *s, const RtCamera *camera, const RtRenderParams *p, float *d_gamma_image, float *d_linear_image,
                           void *cuda_stream, RtStats *stats) {
    return render_device_impl(s, camera, p, d_gamma_image, d_linear_image, cuda_stream, stats, 0u, nullptr, 1);
}


int rt_untile_device(const RtRenderParams *p, const float *d_packed, const int64_t *pixel_offsets, float *d_image,
                     int device, void *cuda_stream) {
    if (!p || !d_packed || !d_image) return fail(RT_ERR_INVALID, "null argument");
    Rect r;
    int rc = resolve_rect(*p, r);
    if (rc) return rc;
    RT_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int nr = p->n_ranks > 1 ? p->n_ranks : 1;
    for (int rk = 0; rk < nr; ++rk) {
        UntileKey key;
        memset(&key, 0, sizeof key);
        key.device = device; key.w = p->width; key.h = p->height; key.x0 = r.x0; key.y0 = r.y0; key.x1 = r.x1; key.y1 = r.y1;
        key.tw = r.tw; key.th = r.th; key.nr = nr; key.rk = rk;
        UntileTab tab;
        {
            std::lock_guard<std::mutex> lock(g_untile_mu);
            auto it = g_untile.find(key);
            if (it == g_untile.end()) {
                RtRenderParams q = *p;
                q.rank = rk;
                std::vector<TileRec> tiles; std::vector<unsigned int> off;
                build_tiles(q, r, tiles, off);
                UntileTab t;
                t.n = (int)tiles.size();
                t.pixels = off.back();
                if (g_untile.size() >= 256) {   // a viewer that keeps resizing: bound the cache (entries of this device go)
                    for (auto jt = g_untile.begin(); jt != g_untile.end();) {
                        if (jt->first.device != device) { ++jt; continue; }
                        if (jt->second.tiles) cudaFree(jt->second.tiles);
                        if (jt->second.off) cudaFree(jt->second.off);
                        jt = g_untile.erase(jt);
                    }
                }
                if (t.n) {
                    cudaError_t e = cudaMalloc((void **)&t.tiles, tiles.size() * sizeof(TileRec));
                    if (e == cudaSuccess) e = cudaMalloc((void **)&t.off, off.size() * sizeof(unsigned int));
                    if (e == cudaSuccess) e = cudaMemcpy(t.tiles, tiles.data(), tiles.size() * sizeof(TileRec), cudaMemcpyHostToDevice);
                    if (e == cudaSuccess) e = cudaMemcpy(t.off, off.data(), off.size() * sizeof(unsigned int), cudaMemcpyHostToDevice);
                    if (e != cudaSuccess) {
                        if (t.tiles) cudaFree(t.tiles);
                        if (t.off) cudaFree(t.off);
                        RT_CUDA(e);
                    }
                }
                it = g_untile.emplace(key, t).first;
            }
            tab = it->second;
        }
        if (!tab.n) continue;
        const float *src = d_packed + 3 * (pixel_offsets ? pixel_offsets[rk] : 0);
        k_untile<<<(tab.pixels + 255) / 256, 256, 0, st>>>(tab.tiles, tab.off, tab.n, src, d_image, r.x0, r.y0, r.x1 - r.x0, tab.pixels);
        RT_CUDA(cudaGetLastError());
    }
    return RT_OK;
}

int rt_render(RtScene *s, const RtCamera *camera, const RtRenderParams *p, float *gamma_rgb, float *linear_rgb, RtStats *stats) {
    if (!s || !camera || !p) return fail(RT_ERR_INVALID, "null argument");
    if (!gamma_rgb && !linear_rgb) return fail(RT_ERR_INVALID, "no output buffer");
    Rect r;
    int rc = resolve_rect(*p, r);
    if (rc) return rc;
    RT_CUDA(cudaSetDevice(s->device));
    const int64_t np = rt_render_pixel_count(p);
    if (np <= 0) { if (stats) memset(stats, 0, sizeof *stats); return RT_OK; }
    const size_t rect_px = (size_t)(r.x1 - r.x0) * (size_t)(r.y1 - r.y0);
    const int n_out = (gamma_rgb ? 1 : 0) + (linear_rgb ? 1 : 0);
    const bool sharded = p->n_ranks > 1;
    RtStats local;
    if (!sharded) {
        // the resolve kernel writes every pixel at its place in the row-major image: no packed buffer, no untile pass
        if ((rc = ensure_out(s, rect_px * 3 * n_out, 0))) return rc;
        float *dg = gamma_rgb ? s->out_img : nullptr;
        float *dl = linear_rgb ? s->out_img + (gamma_rgb ? rect_px * 3 : 0) : nullptr;
        rc = render_device_impl(s, camera, p, dg, dl, nullptr, stats ? stats : &local, 0u, nullptr, 1);
        if (rc) return rc;
        if (gamma_rgb) RT_CUDA(cudaMemcpy(gamma_rgb, dg, rect_px * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        if (linear_rgb) RT_CUDA(cudaMemcpy(linear_rgb, dl, rect_px * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        return RT_OK;
    }
    // one rank of a sharded render with a HOST image: other ranks' tiles must stay untouched, so the packed pixels come
    // back and are scattered row by row on the host (ranks on one box write one device image instead: rt_render_multi)
    if ((rc = ensure_out(s, (size_t)np * 3 * n_out, 0))) return rc;
    float *dg = gamma_rgb ? s->out_img : nullptr;
    float *dl = linear_rgb ? s->out_img + (gamma_rgb ? (size_t)np * 3 : 0) : nullptr;
    rc = rt_render_device(s, camera, p, dg, dl, nullptr, stats ? stats : &local);
    if (rc) return rc;
    std::vector<float> h((size_t)np * 3);
    float *outs[2] = {gamma_rgb, linear_rgb};
    const float *srcs[2] = {dg, dl};
    const int rw = r.x1 - r.x0;
    for (int k = 0; k < 2; ++k) {
        if (!outs[k]) continue;
        RT_CUDA(cudaMemcpy(h.data(), srcs[k], h.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (size_t t = 0; t < s->h_tiles.size(); ++t) {
            const TileRec &tr = s->h_tiles[t];
            const float *src = h.data() + 3 * (size_t)s->h_tile_off[t];
            for (int y = 0; y < tr.h; ++y)
                memcpy(outs[k] + 3 * ((size_t)(tr.y0 - r.y0 + y) * rw + (tr.x0 - r.x0)), src + 3 * (size_t)y * tr.w, (size_t)tr.w * 3 * sizeof(float));
        }
    }
    return RT_OK;
}

int rt_quantize_device(const float *d_values, size_t n, uint8_t *d_bytes, int device, void *cuda_stream) {
    if (!d_values || !d_bytes) return fail(RT_ERR_INVALID, "null argument");
    if (n == 0) return RT_OK;
    RT_CUDA(cudaSetDevice(device));
    k_quantize<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(d_values, n, d_bytes);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

int rt_render_rgb8(RtScene *s, const RtCamera *camera, const RtRenderParams *p, uint8_t *rgb8, RtStats *stats) {
    if (!s || !camera || !p || !rgb8) return fail(RT_ERR_INVALID, "null argument");
    Rect r;
    int rc = resolve_rect(*p, r);
    if (rc) return rc;
    RT_CUDA(cudaSetDevice(s->device));
    const int64_t np = rt_render_pixel_count(p);
    if (np <= 0) { if (stats) memset(stats, 0, sizeof *stats); return RT_OK; }
    const size_t rect_px = (size_t)(r.x1 - r.x0) * (size_t)(r.y1 - r.y0);
    RtStats local;
    if (p->n_ranks > 1) {
        // only this rank's tiles are written: quantise the packed pixels, scatter rows on the host
        if ((rc = ensure_out(s, (size_t)np * 3, (size_t)np * 3))) return rc;
        rc = rt_render_device(s, camera, p, s->out_img, nullptr, nullptr, stats ? stats : &local);
        if (rc) return rc;
        const int rw = r.x1 - r.x0;
        std::vector<uint8_t> hb((size_t)np * 3);
        if ((rc = rt_quantize_device(s->out_img, (size_t)np * 3, s->out_bytes, s->device, nullptr))) return rc;
        RT_CUDA(cudaMemcpy(hb.data(), s->out_bytes, hb.size(), cudaMemcpyDeviceToHost));
        for (size_t t = 0; t < s->h_tiles.size(); ++t) {
            const TileRec &tr = s->h_tiles[t];
            const uint8_t *src = hb.data() + 3 * (size_t)s->h_tile_off[t];
            for (int y = 0; y < tr.h; ++y)
                memcpy(rgb8 + 3 * ((size_t)(tr.y0 - r.y0 + y) * rw + (tr.x0 - r.x0)), src + 3 * (size_t)y * tr.w, (size_t)tr.w * 3);
        }
        return RT_OK;
    }
    if ((rc = ensure_out(s, rect_px * 3, rect_px * 3))) return rc;
    rc = render_device_impl(s, camera, p, s->out_img, nullptr, nullptr, stats ? stats : &local, 0u, nullptr, 1);
    if (rc) return rc;
    if ((rc = rt_quantize_device(s->out_img, rect_px * 3, s->out_bytes, s->device, nullptr))) return rc;
    RT_CUDA(cudaMemcpy(rgb8, s->out_bytes, rect_px * 3, cudaMemcpyDeviceToHost));
    return RT_OK;
}

// ---- multi-GPU render inside one call (replaces the thread-per-row block main.cpp:229-238) -----------
// One host thread per device; the scene is replicated (scenes[i] on its own device); tile t of the 32x32 grid goes to
// device (tx + K ty) % n (tile_owner: round-robin on diagonals; the cost per pixel varies > 100x across an image). Every device's resolve kernel stores
// its pixels straight into the framebuffer on scenes[0]'s device through peer-mapped memory (NVLink / NVSwitch): the
// "gather" of the framebuffer is those stores; there is no packed buffer, no collective and no untile pass.
namespace {
int enable_peer(int from, int to) {
    if (from == to) return RT_OK;
    int can = 0;
    RT_CUDA(cudaDeviceCanAccessPeer(&can, from, to));
    if (!can) return fail(RT_ERR_CUDA, "device " + std::to_string(from) + " cannot map memory of device " + std::to_string(to) + " (no peer access)");
    RT_CUDA(cudaSetDevice(from));
    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) RT_CUDA(e);
    cudaGetLastError();
    return RT_OK;
}
}  // namespace

int rt_render_multi_device(RtScene *const *scenes, int n, const RtCamera *camera, const RtRenderParams *p, float *d_gamma_image,
                           float *d_linear_image, RtStats *stats) {
    if (!scenes || n < 1 || !camera || !p) return fail(RT_ERR_INVALID, "null argument");
    if (!d_gamma_image && !d_linear_image) return fail(RT_ERR_INVALID, "no output buffer");
    if (p->n_ranks > 1) return fail(RT_ERR_INVALID, "rt_render_multi shards over its devices itself: n_ranks must be <= 1");
    for (int i = 0; i < n; ++i) {
        if (!scenes[i]) return fail(RT_ERR_INVALID, "null scene");
        for (int j = 0; j < i; ++j)
            if (scenes[j]->device == scenes[i]->device) return fail(RT_ERR_INVALID, "rt_render_multi: two scenes on the same device");
    }
    int rc;
    for (int i = 1; i < n; ++i)
        if ((rc = enable_peer(scenes[i]->device, scenes[0]->device))) return rc;
    std::vector<RtStats> st((size_t)n);
    std::vector<int> rcs((size_t)n, RT_OK);
    std::vector<std::string> errs((size_t)n);
    auto work = [&](int i) {
        RtScene *s = scenes[i];
        RtRenderParams q = *p;
        q.rank = i; q.n_ranks = n;
        int r = RT_OK;
        if (cudaSetDevice(s->device) != cudaSuccess) r = fail(RT_ERR_CUDA, "cudaSetDevice failed");
        if (!r && !s->stream && cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) r = fail(RT_ERR_CUDA, "cudaStreamCreate failed");
        if (!r) r = render_device_impl(s, camera, &q, d_gamma_image, d_linear_image, s->stream, &st[i], 0u, nullptr, 1);
        if (!r && cudaStreamSynchronize(s->stream) != cudaSuccess) r = fail(RT_ERR_CUDA, std::string("render on device ") + std::to_string(s->device) + ": " + cudaGetErrorString(cudaGetLastError()));
        rcs[i] = r;
        if (r) errs[i] = g_err;
    };
    std::vector<std::thread> threads;
    for (int i = 1; i < n; ++i) threads.emplace_back(work, i);
    work(0);
    for (std::thread &t : threads) t.join();
    for (int i = 0; i < n; ++i)
        if (rcs[i]) return fail(rcs[i], errs[i]);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int i = 0; i < n; ++i) {
            const RtStats &x = st[i];
            stats->n_samples += x.n_samples; stats->n_closest_rays += x.n_closest_rays; stats->n_shadow_rays += x.n_shadow_rays;
            stats->n_sphere_tests += x.n_sphere_tests; stats->n_square_tests += x.n_square_tests; stats->n_mesh_tests += x.n_mesh_tests;
            stats->n_node_visits += x.n_node_visits; stats->n_tri_tests += x.n_tri_tests; stats->n_tri_full += x.n_tri_full;
            stats->n_tex_fetches += x.n_tex_fetches; stats->n_random += x.n_random;
            stats->kernel_ms = std::max(stats->kernel_ms, x.kernel_ms);   // the devices run side by side
            stats->n_launches += x.n_launches; stats->n_tiles += x.n_tiles; stats->n_chunks = std::max(stats->n_chunks, x.n_chunks);
        }
    }
    RT_CUDA(cudaSetDevice(scenes[0]->device));
    return RT_OK;
}

int rt_render_multi(RtScene *const *scenes, int n, const RtCamera *camera, const RtRenderParams *p, float *gamma_rgb, float *linear_rgb,
                    RtStats *stats) {
    if (!scenes || n < 1 || !scenes[0] || !camera || !p) return fail(RT_ERR_INVALID, "null argument");
    if (!gamma_rgb && !linear_rgb) return fail(RT_ERR_INVALID, "no output buffer");
    Rect r;
    int rc = resolve_rect(*p, r);
    if (rc) return rc;
    RtScene *s0 = scenes[0];
    RT_CUDA(cudaSetDevice(s0->device));
    const size_t rect_px = (size_t)(r.x1 - r.x0) * (size_t)(r.y1 - r.y0);
    const int n_out = (gamma_rgb ? 1 : 0) + (linear_rgb ? 1 : 0);
    if ((rc = ensure_out(s0, rect_px * 3 * n_out, 0))) return rc;
    float *dg = gamma_rgb ? s0->out_img : nullptr;
    float *dl = linear_rgb ? s0->out_img + (gamma_rgb ? rect_px * 3 : 0) : nullptr;
    if ((rc = rt_render_multi_device(scenes, n, camera, p, dg, dl, stats))) return rc;
    if (gamma_rgb) RT_CUDA(cudaMemcpy(gamma_rgb, dg, rect_px * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (linear_rgb) RT_CUDA(cudaMemcpy(linear_rgb, dl, rect_px * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return RT_OK;
}

// ---- framebuffer shared between PROCESSES (one rank per GPU, e.g. under torchrun): CUDA IPC ---------------------------
// Rank 0 allocates the image (rt_ipc_alloc) and hands the 64-byte handle to the other ranks by any means; they map it
// (rt_ipc_open) and pass the mapped pointer to rt_render_device_image: their resolve kernels then store into rank 0's
// memory over NVLink, exactly as the threads of rt_render_multi do inside one process.
int rt_ipc_alloc(int device, size_t bytes, void **d_ptr, unsigned char *handle64) {
    if (!d_ptr || !handle64 || bytes == 0) return fail(RT_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    int rc = select_device(device, nullptr);
    if (rc) return rc;
    void *p = nullptr;
    RT_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); RT_CUDA(e); }
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return RT_OK;
}
int rt_ipc_open(int device, const unsigned char *handle64, void **d_ptr) {
    if (!d_ptr || !handle64) return fail(RT_ERR_INVALID, "null argument");
    int rc = select_device(device, nullptr);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    RT_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_OK;
}
int rt_ipc_close(int device, void *d_ptr) {
    if (!d_ptr) return RT_OK;
    RT_CUDA(cudaSetDevice(device));
    RT_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return RT_OK;
}
int rt_ipc_free(int device, void *d_ptr) {
    if (!d_ptr) return RT_OK;
    RT_CUDA(cudaSetDevice(device));
    RT_CUDA(cudaFree(d_ptr));
    return RT_OK;
}

// ---- progressive accumulation -----------------------------------------------------
// The reference renders a frame in one go from a key press (main.cpp:200-263, 321-326) and shows nothing in between.
// An accumulator keeps each pixel's running sample sum on the device; every rt_accum_add() traces `spp` MORE samples per
// pixel (sample indices continue where the last pass stopped, so the random streams are those of one long render) and
// refreshes the mean. After passes of s1, s2, ... samples the frame is bit-identical to ONE render at s1 + s2 + ... spp.
struct RtAccum {
    RtScene *scene = nullptr;       // retained: the device arrays outlive the caller's rt_scene_destroy
    int device = 0;
    RtRenderParams geo{};
    Rect rect{};
    int64_t np = 0;                 // packed pixels of this rank
    size_t rect_px = 0;
    unsigned int done = 0;          // samples per pixel so far
    float *d_sum = nullptr, *d_gamma = nullptr, *d_linear = nullptr, *d_image = nullptr;
    uint8_t *d_bytes = nullptr;
    ~RtAccum() {
        if (d_sum) cudaFree(d_sum);
        if (d_gamma) cudaFree(d_gamma);
        if (d_linear) cudaFree(d_linear);
        if (d_image) cudaFree(d_image);
        if (d_bytes) cudaFree(d_bytes);
    }
};

int rt_accum_create(RtScene *s, const RtRenderParams *geometry, RtAccum **out) {
    if (!s || !geometry || !out) return fail(RT_ERR_INVALID, "null argument");
    *out = nullptr;
    RtRenderParams g = *geometry;
    if (g.spp < 1) g.spp = 1;       // the per-pass count is an argument of rt_accum_add
    Rect r;
    int rc = resolve_rect(g, r);
    if (rc) return rc;
    RT_CUDA(cudaSetDevice(s->device));
    RtAccum *a = new (std::nothrow) RtAccum;
    if (!a) return fail(RT_ERR_OOM, "out of host memory");
    struct Guard { RtAccum *a; bool keep = false; ~Guard() { if (!keep) delete a; } } guard{a};
    a->geo = g; a->rect = r; a->device = s->device;
    a->np = rt_render_pixel_count(&g);
    a->rect_px = (size_t)(r.x1 - r.x0) * (size_t)(r.y1 - r.y0);
    const size_t nb = (size_t)std::max<int64_t>(a->np, 1) * 3 * sizeof(float);
    RT_CUDA(cudaMalloc((void **)&a->d_sum, nb));
    RT_CUDA(cudaMalloc((void **)&a->d_gamma, nb));
    RT_CUDA(cudaMalloc((void **)&a->d_linear, nb));
    RT_CUDA(cudaMalloc((void **)&a->d_image, a->rect_px * 3 * sizeof(float)));
    RT_CUDA(cudaMalloc((void **)&a->d_bytes, a->rect_px * 3));
    RT_CUDA(cudaMemset(a->d_sum, 0, nb));
    RT_CUDA(cudaMemset(a->d_gamma, 0, nb));
    RT_CUDA(cudaMemset(a->d_linear, 0, nb));
    rt_scene_retain(s);
    a->scene = s;
    guard.keep = true;
    *out = a;
    return RT_OK;
}

void rt_accum_destroy(RtAccum *a) {
    if (!a) return;
    cudaSetDevice(a->device);
    RtScene *s = a->scene;
    delete a;
    rt_scene_destroy(s);   // drops the accumulator's reference; frees the scene if the caller already let go of it
}

int rt_accum_reset(RtAccum *a) {
    if (!a) return fail(RT_ERR_INVALID, "null argument");
    a->done = 0;                    // the next pass overwrites the sums (k_resolve ignores them when prior == 0)
    return RT_OK;
}

uint32_t rt_accum_samples(const RtAccum *a) { return a ? a->done : 0u; }

int rt_accum_add(RtAccum *a, const RtCamera *camera, int32_t spp, RtStats *stats) {
    if (!a || !camera) return fail(RT_ERR_INVALID, "null argument");
    if (spp < 1 || (unsigned long long)a->done + (unsigned long long)spp > 0x7FFFFFFFull) return fail(RT_ERR_INVALID, "bad sample count");
    RtRenderParams p = a->geo;
    p.spp = spp;
    RtStats local;
    const int rc = render_device_impl(a->scene, camera, &p, a->d_gamma, a->d_linear, nullptr, stats ? stats : &local, a->done, a->d_sum);
    if (rc) return rc;
    a->done += (unsigned int)spp;
    return RT_OK;
}

int rt_accum_read(RtAccum *a, float *gamma_rgb, float *linear_rgb, uint8_t *rgb8) {
    if (!a) return fail(RT_ERR_INVALID, "null argument");
    if (a->done == 0) return fail(RT_ERR_INVALID, "no samples accumulated yet");
    if (a->np <= 0) return RT_OK;
    RtScene *s = a->scene;
    RT_CUDA(cudaSetDevice(s->device));
    int rc;
    if (a->geo.n_ranks > 1) {
        // only this rank's tiles are written, as in rt_render / rt_render_rgb8
        std::vector<TileRec> tiles; std::vector<unsigned int> off;
        build_tiles(a->geo, a->rect, tiles, off);
        const int rw = a->rect.x1 - a->rect.x0;
        std::vector<float> h((size_t)a->np * 3);
        float *outs[2] = {gamma_rgb, linear_rgb};
        const float *srcs[2] = {a->d_gamma, a->d_linear};
        for (int k = 0; k < 2; ++k) {
            if (!outs[k]) continue;
            RT_CUDA(cudaMemcpy(h.data(), srcs[k], h.size() * sizeof(float), cudaMemcpyDeviceToHost));
            for (size_t t = 0; t < tiles.size(); ++t)
                for (int y = 0; y < tiles[t].h; ++y)
                    memcpy(outs[k] + 3 * ((size_t)(tiles[t].y0 - a->rect.y0 + y) * rw + (tiles[t].x0 - a->rect.x0)),
                           h.data() + 3 * ((size_t)off[t] + (size_t)y * tiles[t].w), (size_t)tiles[t].w * 3 * sizeof(float));
        }
        if (rgb8) {
            std::vector<uint8_t> hb((size_t)a->np * 3);
            if ((rc = rt_quantize_device(a->d_gamma, (size_t)a->np * 3, a->d_bytes, s->device, nullptr))) return rc;
            RT_CUDA(cudaMemcpy(hb.data(), a->d_bytes, hb.size(), cudaMemcpyDeviceToHost));
            for (size_t t = 0; t < tiles.size(); ++t)
                for (int y = 0; y < tiles[t].h; ++y)
                    memcpy(rgb8 + 3 * ((size_t)(tiles[t].y0 - a->rect.y0 + y) * rw + (tiles[t].x0 - a->rect.x0)),
                           hb.data() + 3 * ((size_t)off[t] + (size_t)y * tiles[t].w), (size_t)tiles[t].w * 3);
        }
        return RT_OK;
    }
    int64_t off0 = 0;
    if (linear_rgb) {
        if ((rc = rt_untile_device(&a->geo, a->d_linear, &off0, a->d_image, s->device, nullptr))) return rc;
        RT_CUDA(cudaMemcpy(linear_rgb, a->d_image, a->rect_px * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (gamma_rgb || rgb8) {
        if ((rc = rt_untile_device(&a->geo, a->d_gamma, &off0, a->d_image, s->device, nullptr))) return rc;
        if (gamma_rgb) RT_CUDA(cudaMemcpy(gamma_rgb, a->d_image, a->rect_px * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        if (rgb8) {
            if ((rc = rt_quantize_device(a->d_image, a->rect_px * 3, a->d_bytes, s->device, nullptr))) return rc;
            RT_CUDA(cudaMemcpy(rgb8, a->d_bytes, a->rect_px * 3, cudaMemcpyDeviceToHost));
        }
    }
    return RT_OK;
}

int rt_measure_fp32_peak(int device, double *unfused, double *fused) {
    int sm = 0;
    int rc = select_device(device, &sm);
    if (rc) return rc;
    const int blocks = sm * 8, threads = 256, iters = 4096;
    float *d = nullptr;
    RT_CUDA(cudaMalloc((void **)&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (mode) k_fp32_peak<true><<<blocks, threads>>>(d, 0.999f, 0.001f, iters);
            else k_fp32_peak<false><<<blocks, threads>>>(d, 0.999f, 0.001f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
            if (rep > 0 && ms > 0.f) best[mode] = std::max(best[mode], flops / (ms * 1e-3) / 1e12);
        }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    cudaFree(d);
    RT_CUDA(e);
    if (unfused) *unfused = best[0];
    if (fused) *fused = best[1];
    return RT_OK;
}

int rt_trace_primary(RtScene *s, const RtCamera *camera, const RtRenderParams *p, uint32_t *ids) {
    if (!s || !camera || !p || !ids) return fail(RT_ERR_INVALID, "null argument");
    Rect r;
    int rc = resolve_rect(*p, r);
    if (rc) return rc;
    RT_CUDA(cudaSetDevice(s->device));
    const int rw = r.x1 - r.x0, rh = r.y1 - r.y0;
    const size_t n = (size_t)rw * rh;
    uint32_t *d = nullptr;
    RT_CUDA(cudaMalloc((void **)&d, n * 4 * sizeof(uint32_t)));
    DCamera cam;
    fill_camera(*camera, cam);
    SceneBinding binding;
    if ((rc = binding.bind(s, nullptr))) { cudaFree(d); return rc; }
    k_primary_ids<<<(unsigned)((n + 127) / 128), 128>>>(s->d, cam, p->width, p->height, p->seed, r.x0, r.y0, rw, rh, d);
    cudaError_t e = cudaGetLastError();
    binding.done();
    if (e == cudaSuccess) e = cudaMemcpy(ids, d, n * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    RT_CUDA(e);
    return RT_OK;
}

namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t put(const void *h, size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e == cudaSuccess && h && bytes) e = cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice);
        return e;
    }
};
}  // namespace

int rt_trace_rays(RtScene *s, size_t n, const float *org, const float *dir, const float *time, uint32_t *ids, float *aux) {
    if (!s || !org || !dir || !ids) return fail(RT_ERR_INVALID, "null argument");
    if (n == 0) return RT_OK;
    RT_CUDA(cudaSetDevice(s->device));
    DevBuf o, d, t, i, a;
    RT_CUDA(o.put(org, n * 12)); RT_CUDA(d.put(dir, n * 12));
    if (time) RT_CUDA(t.put(time, n * 4));
    RT_CUDA(i.put(nullptr, n * 16));
    if (aux) RT_CUDA(a.put(nullptr, n * 32));
    SceneBinding binding;
    if (int brc = binding.bind(s, nullptr)) return brc;
    k_trace_rays<<<(unsigned)((n + 127) / 128), 128>>>(s->d, n, (const float *)o.p, (const float *)d.p, time ? (const float *)t.p : nullptr, (uint32_t *)i.p, aux ? (float *)a.p : nullptr);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaMemcpy(ids, i.p, n * 16, cudaMemcpyDeviceToHost));
    if (aux) RT_CUDA(cudaMemcpy(aux, a.p, n * 32, cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_shade_rays(RtScene *s, size_t n, const float *org, const float *dir, const float *time, const RtRenderParams *p, float *rgb) {
    if (!s || !org || !dir || !rgb || !p) return fail(RT_ERR_INVALID, "null argument");
    if (p->max_bounces < 0 || p->max_bounces > RT_MAX_BOUNCES || p->nb_ech < 1) return fail(RT_ERR_INVALID, "bad max_bounces / nb_ech");
    if (n == 0) return RT_OK;
    RT_CUDA(cudaSetDevice(s->device));
    DevBuf o, d, t, c;
    RT_CUDA(o.put(org, n * 12)); RT_CUDA(d.put(dir, n * 12));
    if (time) RT_CUDA(t.put(time, n * 4));
    RT_CUDA(c.put(nullptr, n * 12));
    SceneBinding binding;
    if (int brc = binding.bind(s, nullptr)) return brc;
    k_shade_rays<<<(unsigned)((n + 127) / 128), 128>>>(s->d, n, (const float *)o.p, (const float *)d.p, time ? (const float *)t.p : nullptr, p->seed, p->max_bounces, p->nb_ech, (float *)c.p);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaMemcpy(rgb, c.p, n * 12, cudaMemcpyDeviceToHost));
    return RT_OK;
}

}  // extern "C"
