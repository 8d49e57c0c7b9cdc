// C wrappers over the C++ host API (declared in include/hai719_host.h).
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include "Camera.h"
#include "Preview.h"
#include "Renderer.h"
#include "Scene.h"
#include "SceneFile.h"
#include "errors.h"
#include "hai719_host.h"
#include "matrixUtilities.h"

struct HaiScene {
    Scene scene;
    FlatScene flat;
    bool flat_valid = false;
    std::map<int, std::unique_ptr<hai719::DeviceScene>> on_device;
    void touch() { flat_valid = false; on_device.clear(); }
};

namespace {
thread_local std::string g_err;
template <class F> int guarded(F &&f) {
    // the C ABI reports through return codes; the C++ API's own default (exit on a missing mesh, like the reference)
    // is restored when the call returns
    struct Restore { bool before; ~Restore() { hai719::set_fatal_throws(before); } } restore{hai719::fatal_throws()};
    try {
        hai719::set_fatal_throws(true);
        f();
        return 0;
    } catch (const hai719::RenderError &e) {
        g_err = e.what();
        return e.status;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
}  // namespace

extern "C" {

const char *hai_last_error(void) { return g_err.c_str(); }

HaiScene *hai_scene_new(const char *asset_root) {
    HaiScene *s = new HaiScene;
    if (asset_root) s->scene.asset_root = asset_root;
    return s;
}
void hai_scene_free(HaiScene *s) { delete s; }

int hai_scene_setup(HaiScene *s, int scene_id, float ar, uint32_t seed) {
    return guarded([&] {
        s->touch();
        seed_scene_random(seed);
        if (!s->scene.setup_by_id(scene_id, ar)) throw std::runtime_error("unknown scene id " + std::to_string(scene_id));
    });
}

size_t hai_scene_dump(HaiScene *s, uint32_t *out, size_t cap) {
    std::vector<uint32_t> w;
    s->scene.dump(w);
    if (out) std::memcpy(out, w.data(), std::min(cap, w.size()) * 4);
    return w.size();
}

const RtSceneDesc *hai_scene_flatten(HaiScene *s) {
    const int rc = guarded([&] {
        if (!s->flat_valid) { s->scene.flatten(s->flat); s->flat_valid = true; }
    });
    return rc == 0 ? &s->flat.desc : nullptr;
}

int hai_scene_kd_stats(HaiScene *s, int mesh, uint64_t *o) {
    if (mesh < 0 || (size_t)mesh >= s->scene.meshes.size() || !s->scene.meshes[mesh].kdtree) { g_err = "no such KD-tree"; return -1; }
    const KDTree::Stats st = s->scene.meshes[mesh].kdtree->stats();
    o[0] = st.nodes; o[1] = st.leaves; o[2] = st.empty_leaves; o[3] = st.refs; o[4] = st.max_leaf; o[5] = st.max_depth;
    return 0;
}

int hai_scene_counts(HaiScene *s, uint32_t *o) {
    const Scene &c = s->scene;
    o[0] = (uint32_t)c.spheres.size(); o[1] = (uint32_t)c.squares.size(); o[2] = (uint32_t)c.meshes.size();
    o[3] = (uint32_t)c.lights.size(); o[4] = (uint32_t)c.textures.size(); o[5] = (uint32_t)c.normals.size();
    o[6] = (uint32_t)std::max(0, c.skybox.w); o[7] = (uint32_t)std::max(0, c.skybox.h);
    return 0;
}

int hai_default_camera(int w, int h, RtCamera *out) {
    return guarded([&] {
        Camera camera;
        camera.resize(w, h);
        camera.move(0., 0., -3.1);
        camera.apply();
        MatrixUtilities mu;
        mu.updateMatrices(camera);
        mu.fill(*out);
    });
}

RtScene *hai_scene_device(HaiScene *s, int device) {
    RtScene *h = nullptr;
    guarded([&] {
        auto it = s->on_device.find(device);
        if (it == s->on_device.end())
            it = s->on_device.emplace(device, std::unique_ptr<hai719::DeviceScene>(new hai719::DeviceScene(s->scene, device))).first;
        h = it->second->handle();
    });
    return h;
}

void hai_scene_invalidate_device(HaiScene *s) { s->on_device.clear(); }

int hai_scene_load_file(HaiScene *s, const char *filename) {
    if (!filename) { g_err = "null file name"; return -1; }
    s->touch();
    std::string err;
    if (!hai719::load_scene_file(s->scene, filename, &err)) { g_err = err; return -1; }
    return 0;
}

int hai_scene_move_sphere(HaiScene *s, int index, float dx, float dy, float dz) {
    if (index < 0 || (size_t)index >= s->scene.spheres.size()) { g_err = "no such sphere"; return -1; }
    s->scene.spheres[index].m_center += Vec3(dx, dy, dz);
    s->flat_valid = false;
    return 0;
}

int hai_scene_update_device(HaiScene *s) {
    return guarded([&] {
        for (auto &kv : s->on_device) kv.second->update_analytic(s->scene);
    });
}

int hai_render(HaiScene *s, int device, const RtCamera *cam, const RtRenderParams *params, float *gamma_rgb,
               float *linear_rgb, RtStats *stats) {
    RtScene *h = hai_scene_device(s, device);
    if (!h) return -1;
    const int rc = rt_render(h, cam, params, gamma_rgb, linear_rgb, stats);
    if (rc != RT_OK) g_err = rt_last_error();
    return rc;
}

int hai_render_multi(HaiScene *s, const int *devices, int n_devices, const RtCamera *cam, const RtRenderParams *params,
                     float *gamma_rgb, float *linear_rgb, RtStats *stats) {
    if (!s || !devices || n_devices < 1) { g_err = "hai_render_multi: bad argument"; return -1; }
    std::vector<RtScene *> handles;
    for (int i = 0; i < n_devices; ++i) {
        RtScene *h = hai_scene_device(s, devices[i]);
        if (!h) return -1;
        handles.push_back(h);
    }
    const int rc = rt_render_multi(handles.data(), n_devices, cam, params, gamma_rgb, linear_rgb, stats);
    if (rc != RT_OK) g_err = rt_last_error();
    return rc;
}

int hai_ray_trace_from_camera_multi(HaiScene *s, const int *devices, int n_devices, int w, int h, int nsamples, uint32_t seed,
                                    const char *ppm_path, float *gamma_rgb) {
    return guarded([&] {
        if (!s || !devices || n_devices < 1) throw std::runtime_error("hai_ray_trace_from_camera_multi: bad argument");
        Camera camera;
        camera.resize(w, h);
        camera.move(0., 0., -3.1);
        hai719::RenderOptions opt;
        opt.seed = seed;
        opt.devices.assign(devices, devices + n_devices);
        opt.ppm_path = ppm_path ? ppm_path : "";
        opt.verbose = false;
        std::vector<Vec3> image;
        hai719::ray_trace_from_camera(s->scene, camera, w, h, (unsigned)nsamples, image, opt);   // uploads to every device, renders, frees
        if (gamma_rgb) std::memcpy(gamma_rgb, image.data(), image.size() * sizeof(Vec3));
    });
}

int hai_ray_trace_from_camera(HaiScene *s, int device, int w, int h, int nsamples, uint32_t seed, const char *ppm_path,
                              float *gamma_rgb) {
    return guarded([&] {
        RtScene *dev = hai_scene_device(s, device);
        if (!dev) throw std::runtime_error(g_err);
        Camera camera;
        camera.resize(w, h);
        camera.move(0., 0., -3.1);
        hai719::RenderOptions opt;
        opt.seed = seed;
        opt.device = device;
        opt.ppm_path = ppm_path ? ppm_path : "";
        opt.verbose = false;
        std::vector<Vec3> image;
        hai719::ray_trace_from_camera(*s->on_device[device], camera, w, h, (unsigned)nsamples, image, opt);
        if (gamma_rgb) std::memcpy(gamma_rgb, image.data(), image.size() * sizeof(Vec3));
    });
}

int hai_ray_trace_from_camera_rgb8(HaiScene *s, int device, int w, int h, int nsamples, uint32_t seed, const char *ppm_path,
                                   int p6, uint8_t *rgb8) {
    return guarded([&] {
        RtScene *dev = hai_scene_device(s, device);
        if (!dev) throw std::runtime_error(g_err);
        Camera camera;
        camera.resize(w, h);
        camera.move(0., 0., -3.1);
        hai719::RenderOptions opt;
        opt.seed = seed;
        opt.device = device;
        opt.ppm_path = ppm_path ? ppm_path : "";
        opt.format = p6 == 2 ? hai719::RenderOptions::PNG : p6 ? hai719::RenderOptions::P6 : hai719::RenderOptions::P3;
        opt.verbose = false;
        std::vector<unsigned char> bytes;
        hai719::ray_trace_from_camera_rgb8(*s->on_device[device], camera, w, h, (unsigned)nsamples, bytes, opt);
        if (rgb8) std::memcpy(rgb8, bytes.data(), bytes.size());
    });
}

int hai_write_image_rgb8(const char *path, int format, int w, int h, const uint8_t *rgb8) {
    return guarded([&] {
        if (!path || !rgb8 || w <= 0 || h <= 0) throw std::runtime_error("hai_write_image_rgb8: bad argument");
        const std::vector<unsigned char> bytes(rgb8, rgb8 + (size_t)w * (size_t)h * 3);
        const bool ok = format == 2 ? hai719::write_png(path, w, h, bytes)
                        : format == 1 ? hai719::write_ppm_p6(path, w, h, bytes)
                        : format == 0 ? hai719::write_ppm_p3(path, w, h, bytes) : false;
        if (!ok) throw std::runtime_error(std::string("hai_write_image_rgb8: could not write ") + path);
    });
}

int hai_write_exr(const char *path, int w, int h, const float *rgb) {
    return guarded([&] {
        if (!path || !rgb || w <= 0 || h <= 0) throw std::runtime_error("hai_write_exr: bad argument");
        if (!hai719::write_exr(path, w, h, rgb)) throw std::runtime_error(std::string("hai_write_exr: could not write ") + path);
    });
}

// ---- interactive preview (host/Preview.h) ----------------------------------------------------------
struct HaiPreview {
    Camera camera;
    std::unique_ptr<hai719::Preview> preview;
};

HaiPreview *hai_preview_new(HaiScene *s, int device, int w, int h, uint32_t seed) {
    HaiPreview *p = nullptr;
    const int rc = guarded([&] {
        RtScene *dev = hai_scene_device(s, device);
        if (!dev) throw std::runtime_error(g_err);
        p = new HaiPreview;
        p->camera.resize(w, h);
        p->camera.move(0., 0., -3.1);      // main.cpp:418
        hai719::RenderOptions opt;
        opt.seed = seed;
        opt.device = device;
        opt.verbose = false;
        p->preview.reset(new hai719::Preview(*s->on_device[device], p->camera, w, h, opt));
    });
    if (rc) { delete p; return nullptr; }
    return p;
}

void hai_preview_free(HaiPreview *p) { delete p; }

int hai_preview_mouse(HaiPreview *p, int button, int state, int x, int y) {
    return guarded([&] { if (!p) throw std::runtime_error("null preview"); p->preview->mouse(button, state, x, y); });
}

int hai_preview_motion(HaiPreview *p, int x, int y) {
    return guarded([&] { if (!p) throw std::runtime_error("null preview"); p->preview->motion(x, y); });
}

int hai_preview_resize(HaiPreview *p, int w, int h) {
    return guarded([&] { if (!p) throw std::runtime_error("null preview"); p->preview->resize(w, h); });
}

int hai_preview_invalidate(HaiPreview *p) {
    return guarded([&] { if (!p) throw std::runtime_error("null preview"); p->preview->invalidate(); });
}

int hai_preview_pass(HaiPreview *p, int pass_spp, uint32_t *samples_out) {
    return guarded([&] {
        if (!p) throw std::runtime_error("null preview");
        if (pass_spp < 1) throw std::runtime_error("pass_spp must be >= 1");
        const unsigned int n = p->preview->pass((unsigned)pass_spp);
        if (samples_out) *samples_out = n;
    });
}

int hai_preview_frame(HaiPreview *p, uint8_t *rgb8, float *gamma_rgb) {
    return guarded([&] {
        if (!p) throw std::runtime_error("null preview");
        if (rgb8) {
            const std::vector<unsigned char> &px = p->preview->frame_rgb8();
            std::memcpy(rgb8, px.data(), px.size());
        }
        if (gamma_rgb) {
            std::vector<Vec3> image;
            p->preview->frame(image);
            std::memcpy(gamma_rgb, image.data(), image.size() * sizeof(Vec3));
        }
    });
}

int hai_preview_camera(HaiPreview *p, RtCamera *out) {
    return guarded([&] {
        if (!p || !out) throw std::runtime_error("null argument");
        p->camera.apply();
        MatrixUtilities mu;
        mu.updateMatrices(p->camera);
        mu.fill(*out);
    });
}

}  // extern "C"
