/* hai719_rt.h — C ABI of the B200 (sm_100a) render path.
 *
 * The reference (Kuuro-neko/HAI719-Raytracing) has no plugin/FFI boundary: its render is the body
 * of ray_trace_from_camera() (main.cpp:200-263), reachable only from a GLUT key handler. This
 * header is the boundary that body is replaced behind. Inputs are what that function reads —
 * the selected Scene (Scene.h:57-65), the camera's inverse matrices (matrixUtilities.h:15-19),
 * w, h, nsamples (main.cpp:64,201) and the compile-time knobs of Constants.h:10-12 made runtime —
 * and the output is what it produces: `image`, w*h gamma-corrected float RGB (main.cpp:202,193-196).
 *
 * Plain C: POD structs, pointers and sizes; no C++/torch types. The host-side C++ API
 * (hai719-raytracing_b200/host/, same class names as the reference) fills RtSceneDesc with
 * flatten(); any other host language can do the same through its FFI.
 *
 * All functions return 0 on success or a negative RtStatus; rt_last_error() gives the text for
 * the calling thread. There is NO CPU fallback: without a CUDA device every entry point that
 * needs one fails with RT_ERR_NO_DEVICE.
 */
#ifndef HAI719_RT_H
#define HAI719_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
/* the libraries are built with -fvisibility=hidden: only what is declared here is exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define HAI719_RT_ABI_VERSION 3

typedef enum RtStatus {
    RT_OK = 0,
    RT_ERR_INVALID = -1,    /* bad argument / inconsistent description */
    RT_ERR_NO_DEVICE = -2,  /* no CUDA device, or not an sm_100 part */
    RT_ERR_CUDA = -3,       /* a CUDA runtime call failed (text in rt_last_error) */
    RT_ERR_OOM = -4
} RtStatus;

/* ---- scene description (host memory, read during rt_scene_create only) ------------------- */

/* Material.h:10-20 */
enum { RT_MAT_DIFFUSE = 0, RT_MAT_GLASS = 1, RT_MAT_MIRROR = 2 };
enum { RT_TEX_NONE = 0, RT_TEX_CHECKER = 1, RT_TEX_IMAGE = 2 };
/* Mesh.h:72-76 */
enum { RT_COLOR_VERTEX = 0, RT_COLOR_FACE = 1, RT_COLOR_NONE = 2 };

/* The fields of `struct Material` (Material.h:22-58) that the tracer reads. Pointers to images
 * become indices into RtSceneDesc.textures / .normal_maps (-1 = none). */
typedef struct RtMaterial {
    int32_t type;           /* RT_MAT_* */
    int32_t texture_type;   /* RT_TEX_* */
    float diffuse[3];
    float transparency;
    float index_medium;
    float checker1[3];
    float checker2[3];
    float texture_scale_x, texture_scale_y;
    int32_t emissive;
    float light_color[3];
    float light_intensity;
    int32_t image;          /* index into textures, used when texture_type == RT_TEX_IMAGE */
    int32_t normal_map;     /* index into normal_maps, or -1 (Material::has_normal_map) */
    float motion[3];        /* motion_blur_translation */
} RtMaterial;

/* Sphere.h:41-47 */
typedef struct RtSphere {
    float center[3];
    float radius;
    RtMaterial material;
} RtSphere;

/* Square.h:20-63. v0,v1,v3 are the TRANSFORMED vertices[0,1,3].position that Square::intersect
 * reads (Square.h:68-72); right/up are the m_right_vector / m_up_vector MEMBERS, which transforms
 * never update and which only the normal-map tangent frame uses (Scene.h:284). */
typedef struct RtSquare {
    float v0[3], v1[3], v3[3];
    float right[3], up[3];
    RtMaterial material;
} RtSquare;

/* Scene.h:28-42 (pos, radius, material are the only fields the tracer reads) */
typedef struct RtLight {
    float pos[3];
    float radius;
    float color[3];
} RtLight;

/* ppmLoader::ImageRGB (imageLoader.h:17-26): 8-bit RGB, row-major, w*h*3 bytes. w < 1 or h < 1
 * means "no image" (the reference's own test, Material.cpp:74, Scene.h:150). */
typedef struct RtImage {
    int32_t w, h;
    const uint8_t *rgb;
    /* Identity of the pixel contents, for the device-side image cache: 0 = unknown, the pixels are uploaded by every
     * rt_scene_create; non-zero = the caller vouches that two images with the same id, w and h hold the same pixels
     * (the host loader numbers every file it loads), and a copy already resident on the device is reused: a scene that
     * is re-uploaded every frame then does not move its textures again (Cornell box: 8.4 of 8.45 MB per upload). */
    uint64_t content_id;
} RtImage;

/* One node of the host-built KD-tree (KDTree.cpp:6-29), flattened in PRE-ORDER (node, left
 * subtree, right subtree — the reference's visiting order, KDTree.cpp:49-61). Null children
 * are simply absent. `skip` is the index of the first node after this node's subtree, so the
 * device walks the array front to back with no stack: box missed -> jump to skip, else next.
 * Leaves (KDTree::Node::leaf()) carry a range of RtSceneMesh.leaf_refs, in build order. */
typedef struct RtKdNode {
    float bmin[3];
    uint32_t skip;          /* inner: first index after the subtree; leaf: own index + 1 */
    float bmax[3];
    uint32_t first_ref;     /* leaf: first leaf ref; inner: 0 */
    uint32_t n_refs;        /* leaf: number of refs (may be 0) */
    uint32_t is_leaf;
} RtKdNode;

/* One triangle reference of a leaf: MeshTriangle::v[0..3] (Mesh.h:48-68); v[3] is the index of
 * the triangle in the mesh (Mesh.cpp:71-73), which shading uses (Scene.h:292-297). */
typedef struct RtTriRef {
    uint32_t v[3];
    uint32_t tri_index;
} RtTriRef;

typedef struct RtSceneMesh {
    uint32_t n_vertices, n_triangles;
    const float *positions;      /* 3*n_vertices, as transformed by the scene builder        */
    const uint32_t *triangles;   /* 3*n_triangles vertex indices                              */
    int32_t color_type;          /* RT_COLOR_*                                                */
    const float *vert_colors;    /* 3*n_vertices if RT_COLOR_VERTEX else NULL                 */
    const float *face_colors;    /* 3*n_triangles if RT_COLOR_FACE else NULL                  */
    float root_bmin[3], root_bmax[3]; /* KDTree::aabb (== Mesh::aabb), gate of KDTree::intersect */
    uint32_t n_nodes;            /* 0 = empty tree (root == nullptr): never hit               */
    const RtKdNode *nodes;
    uint32_t n_leaf_refs;
    const RtTriRef *leaf_refs;
    RtMaterial material;
} RtSceneMesh;

typedef struct RtSceneDesc {
    uint32_t abi_version;        /* HAI719_RT_ABI_VERSION */
    uint32_t n_spheres;  const RtSphere *spheres;
    uint32_t n_squares;  const RtSquare *squares;
    uint32_t n_meshes;   const RtSceneMesh *meshes;
    uint32_t n_lights;   const RtLight *lights;
    uint32_t n_textures; const RtImage *textures;
    uint32_t n_normal_maps; const RtImage *normal_maps;
    RtImage skybox;
    int32_t dark_sky;
} RtSceneDesc;

/* ---- camera and render parameters ---------------------------------------------------------- */

/* MatrixUtilities::modelviewInverse / projectionInverse / nearAndFarPlanes[0]
 * (matrixUtilities.h:15-19), column-major doubles exactly as gluInvertMatrix leaves them. */
typedef struct RtCamera {
    double modelview_inverse[16];
    double projection_inverse[16];
    double depth_near;           /* GL_DEPTH_RANGE[0], 0.0 by default */
} RtCamera;

typedef struct RtRenderParams {
    int32_t width, height;       /* full image, as glutGet(GLUT_WINDOW_WIDTH/HEIGHT), main.cpp:201 */
    int32_t spp;                 /* nsamples (main.cpp:64; DEFAULT_NSAMPLES 20)                      */
    int32_t max_bounces;         /* MAXBOUNCES (Constants.h:11), 6                                   */
    int32_t nb_ech;              /* NB_ECH shadow samples per light (Constants.h:12), 10             */
    uint32_t seed;               /* deterministic stream, see "Random numbers" below                 */
    int32_t x0, y0, x1, y1;      /* pixel rectangle to render; all 0 = whole image                   */
    /* tile sharding: the rectangle is cut into tile_w x tile_h tiles, numbered row-major; this
     * call renders the tiles (tx, ty) of the tile grid with (tx + K * ty) % n_ranks == rank, K = 3 (5 if 3 divides
     * n_ranks, 1 if 15 does): round-robin along rows, rows shifted against each other. n_ranks <= 1 renders everything. */
    int32_t rank, n_ranks;
    int32_t tile_w, tile_h;      /* 0 = default 32 x 32                                              */
    int32_t collect_stats;       /* fill the work counters of RtStats (slower)                       */
    int32_t variant;             /* 0 = auto. Low byte = kernel: 1 one path per lane (k_render_paths); 2 ray-level
                                    state machine with path regeneration (k_render_regen), reference-order KD walk;
                                    3 the same over the exact culling hierarchies; 4 warp-voted walk of those;
                                    5 occluder candidates per (hit, light) + gated traversal / shadow-sample
                                    phases; 6 wavefront (camera rays / trace + shade / light + scatter kernels per
                                    bounce level, path state at queue positions). Tuning bits: 8..15 regeneration
                                    threshold, 16..19 CTAs per SM, 20..27 traversal threshold of kernel 5, 28 = no
                                    separate camera-ray pass (kernels 2-5). Wavefront only: 27 = general light kernel
                                    instead of the scatter-only kernel for scenes without lights, 28 = flip the choice
                                    of the split trace (analytic phase + mesh-walk kernel; default: on without lights),
                                    29 = scatter inside the trace kernel for scenes without lights (slower; off).
                                    All variants give identical bits (tests/test_gpu_parity.py, test_gpu_round2.py).
                                    Environment HAI719_CHUNK_LOG2=16..26 (read per call) overrides the number of
                                    paths rendered per chunk (default 2^25 wavefront, 2^26 for wavefront scenes without
                                    meshes, 2^24 otherwise; the chunks of a frame are equal): a memory / speed knob,
                                    results do not depend on it. HAI719_LANES=2 renders the chunks of a frame on two
                                    streams (measured within 1 %, off). The wavefront picks, per scene, kernel
                                    instantiations compiled for what the scene needs (one light: sample kernel without
                                    the next-light walk; no meshes: no mesh code at all): same arithmetic, same bits. */
} RtRenderParams;

typedef struct RtStats {
    uint64_t n_samples;          /* rayTrace() calls = pixels * spp                                  */
    uint64_t n_closest_rays;     /* computeIntersection() calls                                      */
    uint64_t n_shadow_rays;      /* computeShadow() calls                                            */
    uint64_t n_sphere_tests, n_square_tests, n_mesh_tests;
    uint64_t n_node_visits;      /* KD nodes whose box was tested                                    */
    uint64_t n_tri_tests;        /* triangle tests (culled + full)                                   */
    uint64_t n_tri_full;         /* triangle tests that passed the facing + t >= 0 checks            */
    uint64_t n_tex_fetches;
    uint64_t n_random;           /* random_float() draws                                             */
    double   kernel_ms;          /* device time of the trace kernels of this call (CUDA events)      */
    uint32_t n_launches;         /* kernels launched by this call                                    */
    uint32_t n_tiles;            /* tiles rendered by this rank                                      */
    uint32_t n_chunks;           /* launches of the render kernel (the image is rendered in chunks of <= 16 Mi paths) */
    uint32_t reserved;
} RtStats;

/* Random numbers. The reference draws from a time-seeded mt19937 shared by all threads without a
 * lock (Functions.cpp:4-8) plus a thread_local one for jitter (main.cpp:181), so its images are
 * not reproducible. This implementation draws, in the reference's call order of random_float(),
 * from a counter-based stream keyed per path:
 *     fmix32(h): h^=h>>16; h*=0x85EBCA6B; h^=h>>13; h*=0xC2B2AE35; h^=h>>16
 *     key  = fmix32( fmix32(seed ^ (pixel+1)*0x9E3779B9) + (sample+1)*0x85EBCA6B )
 *     draw_i = (fmix32(key + i*0x9E3779B9) >> 8) * 2^-24
 * pixel = x + y*width in full-image coordinates; draws 0,1,2 are the jitter u, v and the ray time
 * (main.cpp:189-192); random_float() continues from 3. Results therefore do not depend on the
 * rectangle, the tile size or the number of ranks. */

typedef struct RtScene RtScene;  /* opaque, device-resident, bound to one device */

int rt_abi_version(void);
int rt_device_count(void);                       /* number of usable sm_100 devices, <= 0 if none */
const char *rt_last_error(void);

/* The consistency checks rt_scene_create applies to a description, on their own and without a device: ABI version,
 * counts against null arrays, material enums and image indices, KD skip links and leaf ranges, leaf-ref and triangle
 * vertex indices. RT_OK or RT_ERR_INVALID (reason in rt_last_error). Host arithmetic only. */
int rt_scene_check(const RtSceneDesc *desc);

/* Copies everything it needs to `device`; the caller keeps ownership of the host arrays. */
int rt_scene_create(const RtSceneDesc *desc, int device, RtScene **out);
/* Frees the scene's arrays. Its render scratch (sample, camera-ray and path-state buffers, grow-only) is kept in a
 * per-device pool and handed to the next rt_scene_create on that device; rt_release_cached_memory frees the pool. */
void rt_scene_destroy(RtScene *scene);
/* A scene handle is reference counted: rt_scene_create returns it with one reference, rt_scene_retain adds one, and
 * every accumulator (rt_accum_create) holds one of its own. rt_scene_destroy drops one reference; the device arrays
 * are released with the last. A caller may therefore destroy / re-create its scene while a preview still refines the
 * old one: the preview keeps rendering the arrays it was created on until it is destroyed itself. */
void rt_scene_retain(RtScene *scene);
int rt_release_cached_memory(int device);
/* SURVEY 8(f)-1, incremental re-upload for animated scenes: rewrite the spheres, squares, lights and their culling
 * hierarchy IN PLACE from `desc` (same counts as the uploaded scene; meshes, textures and normal maps of `desc` are
 * ignored and stay as uploaded). Waits for renders in flight. A render after it equals a render of a fresh upload. */
int rt_scene_update_analytic(RtScene *scene, const RtSceneDesc *desc);
size_t rt_scene_device_bytes(const RtScene *scene);
/* Bytes rt_scene_create really copied host -> device for this handle (images found in the device cache are not copied). */
size_t rt_scene_h2d_bytes(const RtScene *scene);

/* Number of pixels this rank renders under `params` (rectangle + tile sharding), i.e. the
 * element count / 3 of the packed output of rt_render_device(). */
int64_t rt_render_pixel_count(const RtRenderParams *params);

/* The tiles this rank renders under `params`, in packed order: 4 ints per tile {x0, y0, w, h} in full-image
 * pixel coordinates. tiles == NULL returns the count. No device needed (host arithmetic only). */
int64_t rt_tile_layout(const RtRenderParams *params, int32_t *tiles, int64_t cap_tiles);

/* Drop-in for the render part of ray_trace_from_camera(): HOST output buffers.
 * gamma_rgb  : rect_h*rect_w*3 floats, row-major, after gamma_correct (the reference's `image`).
 * linear_rgb : same shape, the value before gamma (sum/nsamples); may be NULL.
 * With n_ranks > 1 pixels of other ranks' tiles are left untouched. */
int rt_render(RtScene *scene, const RtCamera *camera, const RtRenderParams *params,
              float *gamma_rgb, float *linear_rgb, RtStats *stats);

/* Output stage of ray_trace_from_camera() (main.cpp:252-262) on the device: the reference writes
 * (int)(255.f * min(1.f, c)) per channel; this render returns exactly those values as bytes
 * (rect_h*rect_w*3, row-major, host memory), quantised on the GPU so that only 3 bytes per pixel
 * cross PCIe instead of 12. A channel that is negative or NaN — which the reference would print as
 * a negative number — becomes 0; every value a valid PPM can hold is identical.
 * With n_ranks > 1 only this rank's tiles are written. */
int rt_render_rgb8(RtScene *scene, const RtCamera *camera, const RtRenderParams *params, uint8_t *rgb8, RtStats *stats);

/* The same quantisation of n floats already on the device (asynchronous on `cuda_stream`). */
int rt_quantize_device(const float *d_values, size_t n, uint8_t *d_bytes, int device, void *cuda_stream);

/* Same render with DEVICE output, asynchronous on `cuda_stream` (a cudaStream_t; NULL = default
 * stream). d_gamma_rgb / d_linear_rgb (either may be NULL) receive this rank's pixels PACKED
 * tile after tile (tiles in increasing tile index, each tile row-major, edge tiles clipped), the
 * layout rt_untile_device() reads after a gather. stats->kernel_ms is valid only if
 * stats->... was requested with collect_stats or after a stream synchronise. */
int rt_render_device(RtScene *scene, const RtCamera *camera, const RtRenderParams *params,
                     float *d_gamma_rgb, float *d_linear_rgb, void *cuda_stream, RtStats *stats);

/* The same render, but every pixel is stored at its place in the ROW-MAJOR image of the rendered rectangle
 * (rect_h * rect_w * 3 floats, row 0 = top) instead of this rank's packed tile order. The image pointers may be
 * peer-mapped memory of ANOTHER GPU (cudaDeviceEnablePeerAccess in-process, rt_ipc_open across processes): with
 * tile sharding (rank / n_ranks) every rank's resolve kernel then writes its tiles straight into one shared
 * framebuffer over NVLink — no packed buffer, no gather collective, no untile pass. Pixels of other ranks' tiles
 * are not touched. Asynchronous on `cuda_stream` unless stats != NULL. */
int rt_render_device_image(RtScene *scene, const RtCamera *camera, const RtRenderParams *params,
                           float *d_gamma_image, float *d_linear_image, void *cuda_stream, RtStats *stats);

/* Multi-GPU render in ONE call — what replaces the thread-per-scanline block of ray_trace_from_camera()
 * (main.cpp:229-238) on a box with several B200s. scenes[0..n-1] are the same scene uploaded to n DISTINCT devices;
 * one host thread per device renders the tiles with (tx + K * ty) % n == i (params->rank / n_ranks must be unset) and its
 * resolve kernel stores them directly into the framebuffer on scenes[0]'s device through peer-mapped memory.
 * rt_render_multi copies that framebuffer to the HOST buffers (either may be NULL), like rt_render;
 * rt_render_multi_device leaves it in the device buffers given (on scenes[0]'s device, row-major rectangle).
 * Both are synchronous. stats: counts summed over the devices, kernel_ms = the slowest device. The image is
 * bit-identical to rt_render on one device (the random streams are keyed by absolute pixel). */
int rt_render_multi(RtScene *const *scenes, int n_devices, const RtCamera *camera, const RtRenderParams *params,
                    float *gamma_rgb, float *linear_rgb, RtStats *stats);
int rt_render_multi_device(RtScene *const *scenes, int n_devices, const RtCamera *camera, const RtRenderParams *params,
                           float *d_gamma_image, float *d_linear_image, RtStats *stats);

/* One framebuffer shared by one PROCESS per GPU (torchrun, MPI): rank 0 allocates it and publishes the 64-byte CUDA
 * IPC handle, the others map it and hand the mapped pointer to rt_render_device_image. */
int rt_ipc_alloc(int device, size_t bytes, void **d_ptr, unsigned char *handle64);
int rt_ipc_open(int device, const unsigned char *handle64, void **d_ptr);
int rt_ipc_close(int device, void *d_ptr);   /* a pointer from rt_ipc_open */
int rt_ipc_free(int device, void *d_ptr);    /* a pointer from rt_ipc_alloc */

/* SURVEY 8(f)-4, progressive accumulation for an interactive preview. The reference renders a frame in one go on a
 * key press (main.cpp:200-263, 321-326) and shows nothing until it is finished. An accumulator keeps every pixel's
 * running sample sum on the device: each rt_accum_add() traces `spp` MORE samples per pixel — their indices continue
 * where the previous pass stopped, so the random streams are those of one long render — and refreshes the mean.
 * After passes of s1, s2, ... samples the frame read back is bit-identical to ONE rt_render at spp = s1 + s2 + ...
 * (trace_line's additions happen in the same order, main.cpp:188-195). `geometry` supplies everything of
 * RtRenderParams except spp (image size, rectangle, tile sharding, bounces, seed, variant). rt_accum_reset() forgets
 * the samples (camera moved, scene updated). The accumulator holds a reference on the scene (rt_scene_retain), so
 * rt_scene_destroy by the caller cannot pull the device arrays from under it. All calls are synchronous. */
typedef struct RtAccum RtAccum;
int rt_accum_create(RtScene *scene, const RtRenderParams *geometry, RtAccum **out);
void rt_accum_destroy(RtAccum *accum);
int rt_accum_reset(RtAccum *accum);
int rt_accum_add(RtAccum *accum, const RtCamera *camera, int32_t spp, RtStats *stats);
uint32_t rt_accum_samples(const RtAccum *accum);   /* samples per pixel accumulated so far */
/* Current mean frame to HOST buffers (any may be NULL): rect_h*rect_w*3 gamma-corrected floats, the same before
 * gamma, and the reference's 8-bit values (see rt_render_rgb8). With n_ranks > 1 only this rank's tiles are written. */
int rt_accum_read(RtAccum *accum, float *gamma_rgb, float *linear_rgb, uint8_t *rgb8);

/* Scatter packed per-rank tile buffers (as gathered on one device: rank r's buffer starts at
 * float offset 3*pixel_offsets[r]) into a row-major rect_h*rect_w*3 image. */
int rt_untile_device(const RtRenderParams *params, const float *d_packed, const int64_t *pixel_offsets,
                     float *d_image, int device, void *cuda_stream);

/* Primary-hit identification of sample 0's camera ray of every pixel in the rectangle
 * (Scene::computeIntersection, Scene.h:202-230). ids: 4 uint32 per pixel
 * {type (0 miss, 1 sphere, 2 square, 3 mesh), objectIndex, tIndex (meshes), float bits of t}. */
int rt_trace_primary(RtScene *scene, const RtCamera *camera, const RtRenderParams *params, uint32_t *ids);

/* Scene::computeIntersection on caller-supplied rays (directions are normalised as the Ray
 * constructor does, Line.h:13-16). ids as above; aux (8 floats per ray, may be NULL):
 * sphere {theta, phi, n.xyz}, square {u, v, n.xyz}, mesh {w0, w1, w2, n.xyz}. */
int rt_trace_rays(RtScene *scene, size_t n, const float *origins, const float *directions,
                  const float *times, uint32_t *ids, float *aux);

/* Scene::rayTrace (Scene.h:345-350) on caller-supplied rays; ray i draws from
 * key(seed, pixel = i, sample = 0) starting at counter 3. rgb: 3 floats per ray. */
int rt_shade_rays(RtScene *scene, size_t n, const float *origins, const float *directions,
                  const float *times, const RtRenderParams *params, float *rgb);

/* Measurement helper (not on the render path): FP32 peak of `device` in Tflop/s, measured with
 * register-resident dependent chains over all SMs. unfused = separate FMUL + FADD (what the parity
 * build issues, -fmad=false); fused = FFMA counted as 2 flops. MEASURED_PEAKS.json has no FP32
 * figure, so bench.py takes its roofline denominator from here. */
int rt_measure_fp32_peak(int device, double *unfused_tflops, double *fused_tflops);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HAI719_RT_H */
