// Host API — Material (src/Material.h:10-58). Same field names; EVERY field has a defined
// initial value (the reference's constructor leaves emissive, light_*, shininess, image and
// normals indeterminate, Material.cpp:5-11 — SURVEY A.3). scatter/emit/texture/get_normal are the
// device's job (csrc/rt_shade.cuh); on the host the struct is pure data that flatten() copies.
#ifndef HAI719_HOST_MATERIAL_H
#define HAI719_HOST_MATERIAL_H
#include "Vec3.h"
#include "imageLoader.h"

enum MaterialType { Material_Diffuse_Blinn_Phong, Material_Glass, Material_Mirror };
enum TextureType { Texture_None, Texture_Checkerboard, Texture_Image };

struct Material {
    Vec3 ambient_material;
    Vec3 diffuse_material;
    Vec3 specular_material;
    double shininess = 0.;
    Vec3 motion_blur_translation = Vec3(0.f);

    float index_medium = 1.f;
    float transparency = 0.f;

    MaterialType type = Material_Diffuse_Blinn_Phong;
    TextureType texture_type = Texture_None;

    Vec3 checkerboard_color1;
    Vec3 checkerboard_color2;
    float texture_scale_x = 1.f;
    float texture_scale_y = 1.f;

    bool emissive = false;
    Vec3 light_color;
    float light_intensity = 0.f;

    ppmLoader::ImageRGB *image = nullptr;
    ppmLoader::ImageRGB *normals = nullptr;
    bool has_normal_map = false;

    void set_texture(ppmLoader::ImageRGB *img) { image = img; }
    void set_normals(ppmLoader::ImageRGB *img) { normals = img; has_normal_map = true; }
};
#endif
