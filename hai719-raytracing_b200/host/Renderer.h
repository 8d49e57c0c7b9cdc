// Host API — the headless replacement of ray_trace_from_camera() (main.cpp:200-263).
//
// Reference flow:  w,h from GLUT -> camera.apply() -> matrixUtilities.updated()/updateMatrices()
//                  -> one std::thread per scanline running trace_line() -> "Done in" print
//                  -> P3 ./rendu.ppm
// This flow:       w,h from the caller -> camera.apply() -> MatrixUtilities::updateMatrices(camera)
//                  -> Scene::flatten() -> rt_scene_create() (cached in DeviceScene)
//                  -> rt_render() on the GPU -> same P3 writer.
// There is no CPU path: if the CUDA library reports an error, RenderError is thrown.
#ifndef HAI719_HOST_RENDERER_H
#define HAI719_HOST_RENDERER_H
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "Camera.h"
#include "Scene.h"
#include "hai719_rt.h"
#include "matrixUtilities.h"

namespace hai719 {

struct RenderError : std::runtime_error {
    int status;
    RenderError(int status, const std::string &what) : std::runtime_error(what), status(status) {}
};

struct RenderOptions {
    int max_bounces = MAXBOUNCES;
    int nb_ech = NB_ECH;
    uint32_t seed = 0;
    int device = 0;
    // Several B200s of one box: devices = {0, 1, ..., 7} renders the frame on all of them inside the one call (one host
    // thread per device, 32x32 tiles round-robin, every device storing its tiles straight into devices[0]'s framebuffer
    // over NVLink: rt_render_multi). This is what replaces the reference's thread-per-scanline block (main.cpp:229-238).
    // Empty = the single `device` above. The image is bit-identical whatever the device count.
    std::vector<int> devices;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;     // pixel rectangle, all 0 = full image
    int rank = 0, n_ranks = 1;              // tile sharding (see RtRenderParams)
    int tile_w = 0, tile_h = 0;
    bool collect_stats = false;
    int variant = 0;
    std::string ppm_path = "./rendu.ppm";   // "" = do not write (the reference always writes)
    // Output stage. P3: the reference's ASCII file, byte for byte (main.cpp:252-262) — ~400 MB of text
    // at 8K. P6: the same 8-bit values as binary PPM, quantised on the GPU (rt_render_rgb8): 3 bytes per pixel
    // cross PCIe and reach the file. The reference's own ppmLoader reads both. PNG: the same bytes as 8-bit truecolour PNG
    // (write_png below) for viewers that do not read PPM.
    enum Format { P3 = 0, P6 = 1, PNG = 2 } format = P3;
    // Float render only: also write the LINEAR image (sum / nsamples, before gamma_correct) as OpenEXR (write_exr below).
    std::string exr_path;
    bool verbose = true;                    // the reference's two std::cout lines
};

// A scene uploaded to one device. Re-create after changing the Scene.
class DeviceScene {
public:
    DeviceScene(const Scene &scene, int device = 0);
    DeviceScene(const RtSceneDesc &flat, int device);       // an already flattened scene (MultiDeviceScene uploads one flatten() n times)
    ~DeviceScene();
    DeviceScene(const DeviceScene &) = delete;
    DeviceScene &operator=(const DeviceScene &) = delete;
    RtScene *handle() const { return handle_; }
    // Push the scene's spheres, squares and lights again, in place (same counts); meshes and textures stay as uploaded.
    void update_analytic(const Scene &scene);
    int device() const { return device_; }
private:
    RtScene *handle_ = nullptr;
    int device_ = 0;
};

// The same scene on several devices of one box (uploads run side by side, one thread per device).
class MultiDeviceScene {
public:
    MultiDeviceScene(const Scene &scene, const std::vector<int> &devices);
    size_t size() const { return copies_.size(); }
    const DeviceScene &operator[](size_t i) const { return *copies_[i]; }
    std::vector<RtScene *> handles() const;
    void update_analytic(const Scene &scene);
private:
    std::vector<std::unique_ptr<DeviceScene>> copies_;
};

RtRenderParams make_params(int w, int h, unsigned int nsamples, const RenderOptions &opt);

// Render `image` (w*h gamma-corrected Vec3, row 0 = top, like main.cpp:202) from the camera.
void ray_trace_from_camera(const DeviceScene &scene, Camera &camera, int w, int h, unsigned int nsamples,
                           std::vector<Vec3> &image, const RenderOptions &opt = RenderOptions(), RtStats *stats = nullptr);
// The same on every device of `scenes` at once (rt_render_multi); opt.rank / n_ranks must be unset.
void ray_trace_from_camera(const MultiDeviceScene &scenes, Camera &camera, int w, int h, unsigned int nsamples,
                           std::vector<Vec3> &image, const RenderOptions &opt = RenderOptions(), RtStats *stats = nullptr);
// Convenience: uploads the scene (to opt.devices if given, else opt.device), renders, frees.
void ray_trace_from_camera(const Scene &scene, Camera &camera, int w, int h, unsigned int nsamples,
                           std::vector<Vec3> &image, const RenderOptions &opt = RenderOptions(), RtStats *stats = nullptr);

// main.cpp:252-262 — "P3\n w h\n255\n" then (int)(255.f*min(1.f,c)) per channel, space separated.
bool write_ppm_p3(const std::string &filename, int w, int h, const std::vector<Vec3> &image);
// The same file from already quantised bytes (identical output whenever every channel is in [0, 1] or above, i.e. for
// every image a PPM reader accepts), formatted through a 256-entry table: ~20x faster than operator<< per value.
bool write_ppm_p3(const std::string &filename, int w, int h, const std::vector<unsigned char> &rgb8);
// Binary PPM: "P6\n w h\n255\n" + w*h*3 bytes.
bool write_ppm_p6(const std::string &filename, int w, int h, const std::vector<unsigned char> &rgb8);
// 8-bit RGB PNG of the same bytes, self-contained: IHDR, IDAT chunks of at most 1 MiB holding a zlib
// stream of STORED deflate blocks (filter type 0 on every row; no compression library in the image, and the renders are
// noise-like anyway), IEND; CRC-32 and Adler-32 computed here. Any PNG reader decodes it to exactly rgb8.
bool write_png(const std::string &filename, int w, int h, const std::vector<unsigned char> &rgb8);
// Lossless fp32 output: single-part scanline OpenEXR 2, no compression, three FLOAT channels B, G, R (the
// file stores channels in alphabetical order), data window = display window = the image. rgb: h*w*3 floats, row 0 = top.
bool write_exr(const std::string &filename, int w, int h, const float *rgb);

// Render straight to 8-bit RGB (rect_h*rect_w*3, row 0 = top): quantisation on the device, see rt_render_rgb8().
void ray_trace_from_camera_rgb8(const DeviceScene &scene, Camera &camera, int w, int h, unsigned int nsamples,
                                std::vector<unsigned char> &rgb8, const RenderOptions &opt = RenderOptions(), RtStats *stats = nullptr);

}  // namespace hai719
#endif
