#include "Renderer.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>

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

DeviceScene::DeviceScene(const RtSceneDesc &flat, int device) : device_(device) {
    check(rt_scene_create(&flat, device, &handle_), "rt_scene_create");
}

DeviceScene::~DeviceScene() { rt_scene_destroy(handle_); }

MultiDeviceScene::MultiDeviceScene(const Scene &scene, const std::vector<int> &devices) {
    if (devices.empty()) throw RenderError(RT_ERR_INVALID, "MultiDeviceScene: no device given");
    FlatScene flat;
    scene.flatten(flat);
    copies_.resize(devices.size());
    std::vector<std::exception_ptr> errors(devices.size());
    std::vector<std::thread> threads;
    auto upload = [&](size_t i) {
        try { copies_[i].reset(new DeviceScene(flat.desc, devices[i])); } catch (...) { errors[i] = std::current_exception(); }
    };
    for (size_t i = 1; i < devices.size(); ++i) threads.emplace_back(upload, i);
    upload(0);
    for (std::thread &t : threads) t.join();
    for (const std::exception_ptr &e : errors) if (e) std::rethrow_exception(e);
}

std::vector<RtScene *> MultiDeviceScene::handles() const {
    std::vector<RtScene *> h;
    for (const auto &c : copies_) h.push_back(c->handle());
    return h;
}

void MultiDeviceScene::update_analytic(const Scene &scene) {
    for (auto &c : copies_) c->update_analytic(scene);
}

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
    std::vector<float> linear;
    if (!opt.exr_path.empty()) linear.assign((size_t)rw * (size_t)rh * 3, 0.f);
    check(rt_render(scene.handle(), &cam, &p, reinterpret_cast<float *>(image.data()), linear.empty() ? nullptr : linear.data(), stats),
          "rt_render");
    const auto t1 = std::chrono::steady_clock::now();
    if (opt.verbose) std::cout << "  Done in " << std::chrono::duration<double>(t1 - t0).count() << " seconds" << std::endl;
    if (!opt.exr_path.empty() && !write_exr(opt.exr_path, rw, rh, linear.data()))
        std::cout << "Could not open file: " << opt.exr_path << std::endl;
    if (!opt.ppm_path.empty() && !write_ppm_p3(opt.ppm_path, rw, rh, image))
        std::cout << "Could not open file: " << opt.ppm_path << std::endl;
}

void ray_trace_from_camera(const MultiDeviceScene &scenes, Camera &camera, int w, int h, unsigned int nsamples,
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
        std::cout << "Ray tracing a " << rw << " x " << rh << " image on " << scenes.size() << " CUDA devices with "
                  << nsamples << " samples per pixel" << std::endl;
    image.assign((size_t)rw * (size_t)rh, Vec3(0, 0, 0));
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<float> linear;
    if (!opt.exr_path.empty()) linear.assign((size_t)rw * (size_t)rh * 3, 0.f);
    const std::vector<RtScene *> handles = scenes.handles();
    check(rt_render_multi(handles.data(), (int)handles.size(), &cam, &p, reinterpret_cast<float *>(image.data()),
                          linear.empty() ? nullptr : linear.data(), stats), "rt_render_multi");
    const auto t1 = std::chrono::steady_clock::now();
    if (opt.verbose) std::cout << "  Done in " << std::chrono::duration<double>(t1 - t0).count() << " seconds" << std::endl;
    if (!opt.exr_path.empty() && !write_exr(opt.exr_path, rw, rh, linear.data()))
        std::cout << "Could not open file: " << opt.exr_path << std::endl;
    if (!opt.ppm_path.empty() && !write_ppm_p3(opt.ppm_path, rw, rh, image))
        std::cout << "Could not open file: " << opt.ppm_path << std::endl;
}

void ray_trace_from_camera(const Scene &scene, Camera &camera, int w, int h, unsigned int nsamples,
                           std::vector<Vec3> &image, const RenderOptions &opt, RtStats *stats) {
    if (opt.devices.size() > 1) {
        MultiDeviceScene devs(scene, opt.devices);
        ray_trace_from_camera(devs, camera, w, h, nsamples, image, opt, stats);
        return;
    }
    DeviceScene dev(scene, opt.devices.empty() ? opt.device : opt.devices[0]);
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
        const bool ok = opt.format == RenderOptions::PNG  ? write_png(opt.ppm_path, rw, rh, rgb8)
                        : opt.format == RenderOptions::P6 ? write_ppm_p6(opt.ppm_path, rw, rh, rgb8)
                                                          : write_ppm_p3(opt.ppm_path, rw, rh, rgb8);
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

namespace {
struct Crc32 {
    uint32_t table[256];
    Crc32() {
        for (uint32_t n = 0; n < 256; ++n) {
            uint32_t c = n;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[n] = c;
        }
    }
    uint32_t update(uint32_t crc, const unsigned char *p, size_t n) const {
        for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
        return crc;
    }
};
void put_be32(std::vector<unsigned char> &v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16)); v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}
// one PNG chunk: length, type, data, CRC-32 over type + data
void write_chunk(std::ofstream &f, const Crc32 &crc, const char type[4], const unsigned char *data, size_t n) {
    std::vector<unsigned char> head;
    put_be32(head, (uint32_t)n);
    head.insert(head.end(), type, type + 4);
    f.write(reinterpret_cast<const char *>(head.data()), 8);
    if (n) f.write(reinterpret_cast<const char *>(data), (std::streamsize)n);
    uint32_t c = crc.update(0xFFFFFFFFu, head.data() + 4, 4);
    c = crc.update(c, data, n) ^ 0xFFFFFFFFu;
    std::vector<unsigned char> tail;
    put_be32(tail, c);
    f.write(reinterpret_cast<const char *>(tail.data()), 4);
}
}  // namespace

bool write_png(const std::string &filename, int w, int h, const std::vector<unsigned char> &rgb8) {
    if (w <= 0 || h <= 0 || rgb8.size() < (size_t)w * (size_t)h * 3) return false;
    std::ofstream f(filename.c_str(), std::ios::binary);
    if (f.fail()) return false;
    static const Crc32 crc;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    f.write(reinterpret_cast<const char *>(sig), 8);
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (uint32_t)w);
    put_be32(ihdr, (uint32_t)h);
    const unsigned char rest[5] = {8, 2, 0, 0, 0};   // 8 bits per channel, truecolour, deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), rest, rest + 5);
    write_chunk(f, crc, "IHDR", ihdr.data(), ihdr.size());
    // zlib stream: 0x78 0x01, stored blocks of <= 65535 bytes over the filtered scanlines (a 0 byte + the row), Adler-32
    const size_t row = (size_t)w * 3, raw_n = (row + 1) * (size_t)h;
    std::vector<unsigned char> z;
    z.reserve(raw_n + 5 * (raw_n / 65535 + 1) + 6);
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1u, b = 0u;                         // Adler-32, reduced often enough not to overflow
    size_t in_block = 0, produced = 0;
    auto open_block = [&](size_t len, bool last) {
        z.push_back(last ? 1 : 0);
        z.push_back((unsigned char)(len & 0xFF)); z.push_back((unsigned char)(len >> 8));
        z.push_back((unsigned char)(~len & 0xFF)); z.push_back((unsigned char)((~len >> 8) & 0xFF));
        in_block = len;
    };
    auto put = [&](const unsigned char *p, size_t n) {
        while (n) {
            if (in_block == 0) {
                const size_t left = raw_n - produced, len = left < 65535 ? left : 65535;
                open_block(len, len == left);
            }
            const size_t k = n < in_block ? n : in_block;
            z.insert(z.end(), p, p + k);
            for (size_t i = 0; i < k; ++i) { a += p[i]; b += a; if ((i & 2047u) == 2047u) { a %= 65521u; b %= 65521u; } }
            a %= 65521u; b %= 65521u;
            p += k; n -= k; in_block -= k; produced += k;
        }
    };
    const unsigned char filter0 = 0;
    for (int y = 0; y < h; ++y) { put(&filter0, 1); put(rgb8.data() + (size_t)y * row, row); }
    put_be32(z, (b << 16) | a);
    for (size_t off = 0; off < z.size(); off += (1u << 20)) {
        const size_t n = z.size() - off < (1u << 20) ? z.size() - off : (1u << 20);
        write_chunk(f, crc, "IDAT", z.data() + off, n);
    }
    write_chunk(f, crc, "IEND", nullptr, 0);
    return !f.fail();
}

namespace {
void put_le32(std::vector<unsigned char> &v, uint32_t x) {
    for (int i = 0; i < 4; ++i) v.push_back((unsigned char)(x >> (8 * i)));
}
void put_le64(std::vector<unsigned char> &v, uint64_t x) {
    for (int i = 0; i < 8; ++i) v.push_back((unsigned char)(x >> (8 * i)));
}
void put_f32(std::vector<unsigned char> &v, float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    put_le32(v, u);
}
void put_str(std::vector<unsigned char> &v, const char *s) {
    for (; *s; ++s) v.push_back((unsigned char)*s);
    v.push_back(0);
}
// attribute: name, type, byte size, then the value appended by the caller
void put_attr(std::vector<unsigned char> &v, const char *name, const char *type, uint32_t size) {
    put_str(v, name);
    put_str(v, type);
    put_le32(v, size);
}
}  // namespace

bool write_exr(const std::string &filename, int w, int h, const float *rgb) {
    if (w <= 0 || h <= 0 || !rgb) return false;
    std::ofstream f(filename.c_str(), std::ios::binary);
    if (f.fail()) return false;
    std::vector<unsigned char> head;
    put_le32(head, 20000630u);   // magic 0x76 0x2f 0x31 0x01
    put_le32(head, 2u);          // version 2, single-part scanline, short names
    put_attr(head, "channels", "chlist", 3 * 18 + 1);
    for (const char *c : {"B", "G", "R"}) {
        put_str(head, c);
        put_le32(head, 2u);                                   // FLOAT
        head.push_back(0); head.push_back(0); head.push_back(0); head.push_back(0);   // pLinear + 3 reserved bytes
        put_le32(head, 1u); put_le32(head, 1u);               // x / y sampling
    }
    head.push_back(0);
    put_attr(head, "compression", "compression", 1); head.push_back(0);   // NO_COMPRESSION: one scanline per block
    for (const char *name : {"dataWindow", "displayWindow"}) {
        put_attr(head, name, "box2i", 16);
        put_le32(head, 0u); put_le32(head, 0u); put_le32(head, (uint32_t)(w - 1)); put_le32(head, (uint32_t)(h - 1));
    }
    put_attr(head, "lineOrder", "lineOrder", 1); head.push_back(0);       // increasing y
    put_attr(head, "pixelAspectRatio", "float", 4); put_f32(head, 1.f);
    put_attr(head, "screenWindowCenter", "v2f", 8); put_f32(head, 0.f); put_f32(head, 0.f);
    put_attr(head, "screenWindowWidth", "float", 4); put_f32(head, 1.f);
    head.push_back(0);           // end of header
    const uint64_t row_bytes = (uint64_t)w * 12u, block = 8u + row_bytes;
    const uint64_t first = head.size() + 8u * (uint64_t)h;
    for (int y = 0; y < h; ++y) put_le64(head, first + (uint64_t)y * block);
    f.write(reinterpret_cast<const char *>(head.data()), (std::streamsize)head.size());
    std::vector<unsigned char> line;
    std::vector<float> planar((size_t)w * 3);
    for (int y = 0; y < h; ++y) {
        line.clear();
        put_le32(line, (uint32_t)y);
        put_le32(line, (uint32_t)row_bytes);
        f.write(reinterpret_cast<const char *>(line.data()), 8);
        const float *src = rgb + (size_t)y * (size_t)w * 3;
        for (int x = 0; x < w; ++x) {
            planar[x] = src[3 * x + 2];                       // B
            planar[(size_t)w + x] = src[3 * x + 1];           // G
            planar[2 * (size_t)w + x] = src[3 * x];           // R
        }
        f.write(reinterpret_cast<const char *>(planar.data()), (std::streamsize)row_bytes);   // little-endian host
    }
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
