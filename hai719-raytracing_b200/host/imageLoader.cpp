#include "imageLoader.h"

#include <atomic>

namespace ppmLoader {
namespace {
// skip line breaks, then at most one '#' comment line (imageLoader.cpp:10-18)
void skip_comment(std::ifstream &f) {
    int c;
    while ((c = f.peek()) == '\n' || c == '\r') f.get();
    if (c == '#') { std::string junk; std::getline(f, junk); }
}
bool fail(ImageRGB &img, const std::string &why) {
    std::cout << why << std::endl;
    img.w = img.h = 0;
    img.data.clear();
    img.content_id = 0;
    return false;
}
}  // namespace

bool load_ppm(ImageRGB &img, const std::string &name) {
    std::ifstream f(name.c_str(), std::ios::binary);
    if (f.fail()) return fail(img, "Could not open file: " + name);
    std::string magic;
    int bits = 0;
    skip_comment(f); f >> magic;
    skip_comment(f); f >> img.w;
    skip_comment(f); f >> img.h;
    skip_comment(f); f >> bits;
    const bool p3 = magic == "P3", p6 = magic == "P6";
    if (!p3 && !p6) return fail(img, "Unsupported magic number");
    if (img.w < 1) return fail(img, "Unsupported width: " + std::to_string(img.w));
    if (img.h < 1) return fail(img, "Unsupported height: " + std::to_string(img.h));
    if (bits < 1 || bits > 255) return fail(img, "Unsupported number of bits: " + std::to_string(bits));
    // Hardened: the pixel count must fit in what is left of the file (3 bytes per pixel in P6, at least
    // "0 " per sample in P3) before anything is allocated, and a pixel block that ends early is a failed load (empty
    // image, as for an unreadable file) instead of a half-filled one.
    if (f.fail()) return fail(img, "Unreadable PPM header: " + name);
    const std::streamoff here = f.tellg();
    f.seekg(0, std::ios::end);
    const long long left = (long long)(f.tellg() - here);
    f.seekg(here, std::ios::beg);
    const long long samples = (long long)img.w * (long long)img.h * 3;
    if (img.w > left || img.h > left || (p6 ? samples > left - 1 : 2 * samples - 1 > left))
        return fail(img, "PPM pixel data ends early (" + std::to_string(img.w) + " x " + std::to_string(img.h) + "): " + name);
    img.data.assign((size_t)img.w * (size_t)img.h, RGB{0, 0, 0});
    if (p6) {
        f.get();  // the single whitespace byte after maxval
        f.read((char *)img.data.data(), (std::streamsize)(img.data.size() * 3));
    } else {
        for (RGB &px : img.data) {
            int v;
            f >> v; px.r = (unsigned char)v;
            f >> v; px.g = (unsigned char)v;
            f >> v; px.b = (unsigned char)v;
        }
    }
    if (f.fail()) return fail(img, "PPM pixel data ends early or is not numeric: " + name);
    static std::atomic<unsigned long long> next_id{1};
    img.content_id = next_id.fetch_add(1);
    return true;
}
}  // namespace ppmLoader
