/* hai719_host.h — C entry points of the C++ host API (hai719-raytracing_b200/host/), so that
 * tests, bench.py and non-C++ callers can drive the same objects a C++ user of the drop-in would:
 * build a Scene with the reference's setup_*() builders (Scene.h:358-1882) or the OFF/PPM
 * loaders, get the default camera of main.cpp:418 (Camera() + move(0,0,-3.1)), flatten
 * to the RtSceneDesc of hai719_rt.h, and render through ray_trace_from_camera()'s replacement.
 *
 * All functions return 0 on success, negative on error (hai_last_error() has the text), unless
 * documented otherwise. Nothing here computes an image on the CPU.
 */
#ifndef HAI719_HOST_H
#define HAI719_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "hai719_rt.h"
#ifdef __cplusplus
extern "C" {
#endif
/* the libraries are built with -fvisibility=hidden: only what is declared here is exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct HaiScene HaiScene;

const char *hai_last_error(void);

HaiScene *hai_scene_new(const char *asset_root);
void hai_scene_free(HaiScene *s);
/* scene_id: 0..10 = main.cpp:421-432 order, 11 = flamingo_lake, 100 = BASELINE config 5.
 * seed feeds the scene-construction randomness (random spheres). */
int hai_scene_setup(HaiScene *s, int scene_id, float aspect_ratio, uint32_t seed);
/* canonical dump (same word layout as the oracle's ref_scene_dump); out == NULL returns the size */
size_t hai_scene_dump(HaiScene *s, uint32_t *out, size_t cap_words);
/* Flatten; the returned description stays valid until the scene is changed or freed. */
const RtSceneDesc *hai_scene_flatten(HaiScene *s);
/* per-mesh KD statistics: out6 = {nodes, leaves, empty_leaves, refs, max_leaf, max_depth} */
int hai_scene_kd_stats(HaiScene *s, int mesh, uint64_t *out6);
int hai_scene_counts(HaiScene *s, uint32_t *out8); /* spheres, squares, meshes, lights, textures, normals, sky_w, sky_h */

/* Camera() ; resize(w,h) ; move(0,0,-3.1) ; apply() ; MatrixUtilities::updateMatrices() */
int hai_default_camera(int w, int h, RtCamera *out);

/* Upload (cached per scene+device until the scene changes) and render; host output buffers as
 * in rt_render(). This is the call a user of the drop-in makes: the e2e number in bench.py. */
int hai_render(HaiScene *s, int device, const RtCamera *cam, const RtRenderParams *params, float *gamma_rgb,
               float *linear_rgb, RtStats *stats);
/* The same on several devices of one box at once (rt_render_multi: tiles round-robin over the devices, every device
 * writing its tiles into devices[0]'s framebuffer over NVLink). Device copies are cached per device like hai_render's. */
int hai_render_multi(HaiScene *s, const int *devices, int n_devices, const RtCamera *cam, const RtRenderParams *params,
                     float *gamma_rgb, float *linear_rgb, RtStats *stats);
/* ray_trace_from_camera() with RenderOptions::devices: upload to every device, render, write the P3 file, free. */
int hai_ray_trace_from_camera_multi(HaiScene *s, const int *devices, int n_devices, int w, int h, int nsamples, uint32_t seed,
                                    const char *ppm_path, float *gamma_rgb);
/* The device-resident scene handle (uploading if needed), for rt_render_device() etc. */
RtScene *hai_scene_device(HaiScene *s, int device);

/* Replace the scene by the one a scene description file builds (grammar: host/SceneFile.cpp; file names inside it are
 * relative to the asset root given to hai_scene_new). Errors come back as "<file>:<line>: <what>". */
int hai_scene_load_file(HaiScene *s, const char *filename);

/* Animated scenes: move sphere `index` by (dx, dy, dz) on the host; hai_scene_update_device() then pushes the analytic
 * primitives (spheres, squares, lights) of the host scene to the device copies that exist, in place
 * (rt_scene_update_analytic) — meshes and textures are not uploaded again. */
int hai_scene_move_sphere(HaiScene *s, int index, float dx, float dy, float dz);
int hai_scene_update_device(HaiScene *s);

/* Drop the cached device copy, so the next hai_render()/hai_scene_device() uploads again (what a
 * fresh ray_trace_from_camera(scene, ...) call does every time). */
void hai_scene_invalidate_device(HaiScene *s);

/* ray_trace_from_camera() end to end, including the P3 file (main.cpp:252-262); ppm_path may be NULL */
int hai_ray_trace_from_camera(HaiScene *s, int device, int w, int h, int nsamples, uint32_t seed, const char *ppm_path,
                              float *gamma_rgb);

/* The same with the output stage on the GPU: 8-bit RGB (h*w*3 bytes, row 0 = top) = the values the reference
 * writes to rendu.ppm. ppm_path (may be NULL): the reference's P3 text when p6 == 0, 8-bit RGB PNG when p6 == 2,
 * binary P6 for any other value. */
int hai_ray_trace_from_camera_rgb8(HaiScene *s, int device, int w, int h, int nsamples, uint32_t seed, const char *ppm_path,
                                   int p6, uint8_t *rgb8);

/* The output stage on its own (replaces the file loop of main.cpp:252-262 for bytes that are already
 * quantised): h*w*3 bytes, row 0 = top, written as format 0 = P3 text (the reference's file, byte for byte),
 * 1 = binary P6, 2 = 8-bit RGB PNG (stored deflate blocks, no library). Host only, no GPU involved. */
int hai_write_image_rgb8(const char *path, int format, int w, int h, const uint8_t *rgb8);
/* Lossless fp32 output: h*w*3 floats (row 0 = top; e.g. the linear_rgb of rt_render) as an uncompressed scanline
 * OpenEXR file with FLOAT channels B, G, R. Host only. */
int hai_write_exr(const char *path, int w, int h, const float *rgb);

/* Interactive preview (host/Preview.h): the reference's mouse handlers (main.cpp:344-388; button 0
 * left = rotate, 1 middle = zoom, 2 right = move; state 0 down, 1 up) drive a Camera placed like main.cpp:418, and
 * every hai_preview_pass() adds pass_spp samples per pixel to the frame on the GPU (rt_accum_*). Moving the camera
 * restarts the accumulation. hai_preview_frame() returns the current mean: h*w*3 bytes and/or h*w*3 gamma floats.
 * A preview holds its own reference on the scene's device copy (rt_scene_retain): after hai_scene_setup /
 * hai_scene_load_file / hai_scene_invalidate_device / hai_scene_free it keeps refining the scene it was created on;
 * create a new preview to see the new scene. */
typedef struct HaiPreview HaiPreview;
HaiPreview *hai_preview_new(HaiScene *s, int device, int w, int h, uint32_t seed);
void hai_preview_free(HaiPreview *p);
int hai_preview_mouse(HaiPreview *p, int button, int state, int x, int y);
int hai_preview_motion(HaiPreview *p, int x, int y);
int hai_preview_resize(HaiPreview *p, int w, int h);
int hai_preview_invalidate(HaiPreview *p);   /* after hai_scene_update_device(): restart the accumulation */
int hai_preview_pass(HaiPreview *p, int pass_spp, uint32_t *samples_out);
int hai_preview_frame(HaiPreview *p, uint8_t *rgb8, float *gamma_rgb);
int hai_preview_camera(HaiPreview *p, RtCamera *out);   /* the matrices the next pass will use */

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
