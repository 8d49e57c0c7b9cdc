#include "Renderer.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>

namespace hai719 {

namespace {
void check(int status, const char *what) {
    if (status != RT_OK) throw RenderError(status, std::string(what) + ": " + rt_last_error());
}
}  // namespace

DeviceScene::DeviceScene(const Scene &scene, int device) : device_(device) {
    FlatScene flat;
    scene.flatten(flat);
    check(rt_scene_create(&flat.desc, device, &handle_), "rt_scene_create");
}

DeviceScene::~DeviceScene() { rt_scene_destroy(handle_); }

void DeviceScene::update_analytic(const Scene &scene) {
    FlatScene flat;
    scene.flatten(flat);
    check(rt_scene_update_analytic(handle_, &flat.desc), "rt_scene_update_analytic");
}

RtRenderParams make_params(int w, int h, unsigned int nsamples, const RenderOptions &o) {
    RtRenderParams p{};
    p.width = w; p.height = h; p.spp = (int32_t)nsamples;
    p.max_bounces = o.max_bounces; p.nb_ech = o.nb_ech; p.seed = o.seed;
    p.x0 = o.x0; p.y0 = o.y0; p.x1 = o.x1; p.y1 = o.y1;
    p.rank = o.rank; p.n_ranks = o.n_ranks; p.tile_w = o.tile_w; p.tile_h = o.tile_h;
    p.collect_stats = o.collect_stats ? 1 : 0;
    p.variant = o.variant;
    return p;
}

void ray_trace_from_camera(const DeviceScene &scene, Camera &camera, int w, int h, unsigned int nsamples,
                           std::vector<Vec3> &image, const RenderOptions &opt, RtStats *stats) {
    camera.apply();
    MatrixUtilities mu;
    mu.updateMatrices(camera);
    RtCamera cam;
    mu.fill(cam);
    const RtRenderParams p = make_params(w, h, nsamples, opt);
    const bool full = (p.x0 | p.y0 | p.x1 | p.y1) == 0;
    const int rw = full ? w : p.x1 - p.x0, rh = full ? h : p.y1 - p.y0;
    if (opt.verbose)
        std::cout << "Ray tracing a " << rw << " x " << rh << " image on CUDA device " << scene.device() << " with "
                  << nsamples << " samples per pixel" << std::endl;
    image.assign((size_t)rw * (size_t)rh, Vec3(0, 0, 0));
    static_assert(sizeof(Vec3) == 3 * sizeof(float), "Vec3 must be three packed floats");
    const auto t0 = std::chrono::steady_clock::now();
    check(rt_render(scene.handle(), &cam, &p, reinterpret_cast<float *>(image.data()), nullptr, stats), "rt_render");
    const auto t1 = std::chrono::steady_clock::now();
    if (opt.verbose) std::cout << "  Done in " << std::chrono::duration<double>(t1 - t0).count() << " seconds" << std::endl;
    if (!opt.ppm_path.empty() && !write_ppm_p3(opt.ppm_path, rw, rh, image))
        std::cout << "Could not open file: " << opt.ppm_path << std::endl;
}

void ray_trace_from_camera(const Scene &scene, Camera &camera, int w, int h, unsigned int nsamples,
                           std::vector<Vec3> &image, const RenderOptions &opt, RtStats *stats) {
    DeviceScene dev(scene, opt.device);
    ray_trace_from_camera(dev, camera, w, h, nsamples, image, opt, stats);
}

void ray_trace_from_camera_rgb8(const DeviceScene &scene, Camera &camera, int w, int h, unsigned int nsamples,
                                std::vector<unsigned char> &rgb8, const RenderOptions &opt, RtStats *stats) {
    camera.apply();
    MatrixUtilities mu;
    mu.updateMatrices(camera);
    RtCamera cam;
    mu.fill(cam);
    const RtRenderParams p = make_params(w, h, nsamples, opt);
    const bool full = (p.x0 | p.y0 | p.x1 | p.y1) == 0;
    const int rw = full ? w : p.x1 - p.x0, rh = full ? h : p.y1 - p.y0;
    if (opt.verbose)
        std::cout << "Ray tracing a " << rw << " x " << rh << " image on CUDA device " << scene.device() << " with "
                  << nsamples << " samples per pixel" << std::endl;
    rgb8.assign((size_t)rw * (size_t)rh * 3, 0);
    const auto t0 = std::chrono::steady_clock::now();
    check(rt_render_rgb8(scene.handle(), &cam, &p, rgb8.data(), stats), "rt_render_rgb8");
    const auto t1 = std::chrono::steady_clock::now();
    if (opt.verbose) std::cout << "  Done in " << std::chrono::duration<double>(t1 - t0).count() << " seconds" << std::endl;
    if (!opt.ppm_path.empty()) {
        const bool ok = opt.format == RenderOptions::P6 ? write_ppm_p6(opt.ppm_path, rw, rh, rgb8) : write_ppm_p3(opt.ppm_path, rw, rh, rgb8);
        if (!ok) std::cout << "Could not open file: " << opt.ppm_path << std::endl;
    }
}

bool write_ppm_p3(const std::string &filename, int w, int h, const std::vector<unsigned char> &rgb8) {
    std::ofstream f(filename.c_str(), std::ios::binary);
    if (f.fail()) return false;
    f << "P3" << std::endl << w << " " << h << std::endl << 255 << std::endl;
    char table[256][4];
    int len[256];
    for (int v = 0; v < 256; ++v) len[v] = std::snprintf(table[v], 4, "%d", v);
    std::vector<char> buf;
    buf.reserve(1 << 20);
    const size_t n = (size_t)w * (size_t)h * 3;
    for (size_t i = 0; i < n; ++i) {
        const unsigned char v = rgb8[i];
        buf.insert(buf.end(), table[v], table[v] + len[v]);
        buf.push_back(' ');
        if (buf.size() > (1 << 20) - 8) { f.write(buf.data(), (std::streamsize)buf.size()); buf.clear(); }
    }
    f.write(buf.data(), (std::streamsize)buf.size());
    f << std::endl;
    return !f.fail();
}

bool write_ppm_p6(const std::string &filename, int w, int h, const std::vector<unsigned char> &rgb8) {
    std::ofstream f(filename.c_str(), std::ios::binary);
    if (f.fail()) return false;
    f << "P6\n" << w << " " << h << "\n255\n";
    f.write(reinterpret_cast<const char *>(rgb8.data()), (std::streamsize)((size_t)w * (size_t)h * 3));
    return !f.fail();
}

bool write_ppm_p3(const std::string &filename, int w, int h, const std::vector<Vec3> &image) {
    std::ofstream f(filename.c_str(), std::ios::binary);
    if (f.fail()) return false;
    f << "P3" << std::endl << w << " " << h << std::endl << 255 << std::endl;
    for (size_t i = 0, n = (size_t)w * (size_t)h; i < n; ++i)
        for (unsigned int c = 0; c < 3; ++c) f << (int)(255.f * std::min<float>(1.f, image[i][c])) << " ";
    f << std::endl;
    return true;
}

}  // namespace hai719
