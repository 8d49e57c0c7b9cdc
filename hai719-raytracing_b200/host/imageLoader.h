// Host API — PPM loader with the reference's names (src/imageLoader.h:11-38): P3 and P6,
// '#' comment lines between header fields, 8-bit RGB kept as bytes (the device samples them in
// software, nearest texel, exactly like Material::texture). A file that cannot be opened or
// parsed leaves the image EMPTY (w = h = 0) after printing a message — the reference leaves w/h
// unset in that case (imageLoader.cpp:24-28); "empty" is the defined version of that.
#ifndef HAI719_HOST_IMAGELOADER_H
#define HAI719_HOST_IMAGELOADER_H
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

namespace ppmLoader {
struct RGB { unsigned char r, g, b; };
struct ImageRGB {
    int w = 0, h = 0;
    std::vector<RGB> data;
    // identity of the pixel contents for the device image cache (RtImage::content_id): load_ppm gives every file it loads
    // a fresh number; code that edits `data` afterwards must set it to 0 (= always upload) or to a number of its own
    unsigned long long content_id = 0;
};
bool load_ppm(ImageRGB &img, const std::string &name);
}  // namespace ppmLoader
#endif
