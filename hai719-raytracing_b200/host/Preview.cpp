#include "Preview.h"

namespace hai719 {

namespace {
void check(int status, const char *what) {
    if (status != RT_OK) throw RenderError(status, std::string(what) + ": " + rt_last_error());
}
}  // namespace

Preview::Preview(const DeviceScene &scene, Camera &camera, int w, int h, const RenderOptions &opt)
    : scene_(scene.handle()), camera_(camera), opt_(opt), w_(w), h_(h) {
    rt_scene_retain(scene_);
    try { rebuild(); } catch (...) { rt_scene_destroy(scene_); throw; }
}

Preview::~Preview() {
    rt_accum_destroy(accum_);
    rt_scene_destroy(scene_);   // drops this preview's reference
}

void Preview::rebuild() {
    if (accum_) { rt_accum_destroy(accum_); accum_ = nullptr; }
    const RtRenderParams p = make_params(w_, h_, 1, opt_);
    check(rt_accum_create(scene_, &p, &accum_), "rt_accum_create");
    dirty_ = true;
    rgb8_samples_ = 0;
}

void Preview::invalidate() { dirty_ = true; }

void Preview::mouse(int button, int state, int x, int y) {
    if (state == Up) {
        rotate_ = move_ = zoom_ = false;
        return;
    }
    if (button == Left) {
        camera_.beginRotate(x, y);
        move_ = false; rotate_ = true; zoom_ = false;
    } else if (button == Right) {
        last_x_ = x; last_y_ = y;
        move_ = true; rotate_ = false; zoom_ = false;
    } else if (button == Middle) {
        if (!zoom_) {
            last_zoom_ = y;
            move_ = false; rotate_ = false; zoom_ = true;
        }
    }
}

void Preview::motion(int x, int y) {
    if (rotate_) {
        camera_.rotate(x, y);
    } else if (move_) {
        camera_.move((x - last_x_) / static_cast<float>(w_), (last_y_ - y) / static_cast<float>(h_), 0.0);
        last_x_ = x; last_y_ = y;
    } else if (zoom_) {
        camera_.zoom(float(y - last_zoom_) / h_);
        last_zoom_ = y;
    } else {
        return;
    }
    dirty_ = true;
}

void Preview::resize(int w, int h) {
    camera_.resize(w, h);
    w_ = w; h_ = h;
    rebuild();
}

unsigned int Preview::samples() const { return dirty_ ? 0u : rt_accum_samples(accum_); }

unsigned int Preview::pass(unsigned int pass_spp, RtStats *stats) {
    if (dirty_) {
        check(rt_accum_reset(accum_), "rt_accum_reset");
        dirty_ = false;
    }
    camera_.apply();
    MatrixUtilities mu;
    mu.updateMatrices(camera_);
    RtCamera cam;
    mu.fill(cam);
    check(rt_accum_add(accum_, &cam, (int32_t)pass_spp, stats), "rt_accum_add");
    rgb8_samples_ = 0;
    return rt_accum_samples(accum_);
}

const std::vector<unsigned char> &Preview::frame_rgb8() {
    const unsigned int n = samples();
    if (n == 0) throw RenderError(RT_ERR_INVALID, "Preview::frame_rgb8: no pass since the camera moved");
    if (rgb8_samples_ != n) {
        const RtRenderParams p = make_params(w_, h_, 1, opt_);
        const bool full = (p.x0 | p.y0 | p.x1 | p.y1) == 0;
        const int rw = full ? w_ : p.x1 - p.x0, rh = full ? h_ : p.y1 - p.y0;
        rgb8_.assign((size_t)rw * (size_t)rh * 3, 0);
        check(rt_accum_read(accum_, nullptr, nullptr, rgb8_.data()), "rt_accum_read");
        rgb8_samples_ = n;
    }
    return rgb8_;
}

void Preview::frame(std::vector<Vec3> &image) {
    if (samples() == 0) throw RenderError(RT_ERR_INVALID, "Preview::frame: no pass since the camera moved");
    const RtRenderParams p = make_params(w_, h_, 1, opt_);
    const bool full = (p.x0 | p.y0 | p.x1 | p.y1) == 0;
    const int rw = full ? w_ : p.x1 - p.x0, rh = full ? h_ : p.y1 - p.y0;
    image.assign((size_t)rw * (size_t)rh, Vec3(0, 0, 0));
    check(rt_accum_read(accum_, reinterpret_cast<float *>(image.data()), nullptr, nullptr), "rt_accum_read");
}

bool Preview::save(const std::string &filename, RenderOptions::Format format) {
    const std::vector<unsigned char> &px = frame_rgb8();
    const int rw = (opt_.x0 | opt_.y0 | opt_.x1 | opt_.y1) == 0 ? w_ : opt_.x1 - opt_.x0;
    const int rh = (opt_.x0 | opt_.y0 | opt_.x1 | opt_.y1) == 0 ? h_ : opt_.y1 - opt_.y0;
    return format == RenderOptions::P6 ? write_ppm_p6(filename, rw, rh, px) : write_ppm_p3(filename, rw, rh, px);
}

}  // namespace hai719
