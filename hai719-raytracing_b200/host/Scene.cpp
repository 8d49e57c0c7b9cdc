// Host API — scene containers, the reference's scene builders restated as data, flatten() and
// the canonical dump. See Scene.h for what is (deliberately) not here.
#include "Scene.h"
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include "errors.h"

// ------------------------------------------------------------------------------------------------
// small builder vocabulary
// ------------------------------------------------------------------------------------------------
namespace {

// A transform step on a Mesh/Square; applied in list order, exactly the calls the reference's
// setup_*() functions make (translate / scale / rotate_x / rotate_y / rotate_z).
struct Op {
    enum Kind { T, S, RX, RY, RZ } kind;
    float a, b, c;
};
inline Op T(float x, float y, float z) { return {Op::T, x, y, z}; }
inline Op S(float x, float y, float z) { return {Op::S, x, y, z}; }
inline Op S(float s) { return {Op::S, s, s, s}; }
inline Op RX(float deg) { return {Op::RX, deg, 0, 0}; }
inline Op RY(float deg) { return {Op::RY, deg, 0, 0}; }
inline Op RZ(float deg) { return {Op::RZ, deg, 0, 0}; }

void place(Mesh &m, std::initializer_list<Op> ops) {
    for (const Op &o : ops) switch (o.kind) {
            case Op::T: m.translate(Vec3(o.a, o.b, o.c)); break;
            case Op::S: m.scale(Vec3(o.a, o.b, o.c)); break;
            case Op::RX: m.rotate_x(o.a); break;
            case Op::RY: m.rotate_y(o.a); break;
            case Op::RZ: m.rotate_z(o.a); break;
        }
    m.build_arrays();
}

// The two 2x2 quads every wall/floor in the reference starts from.
void centred_quad(Square &s) { s.setQuad(Vec3(-1.f, -1.f, 0.f), Vec3(1.f, 0.f, 0.f), Vec3(0.f, 1.f, 0.f), 2.f, 2.f); }
void low_quad(Square &s) { s.setQuad(Vec3(-1.f, -0.2f, 0.f), Vec3(1.f, 0.f, 0.f), Vec3(0.f, 1.f, 0.f), 2.f, 2.f); }

void checker(Material &m, Vec3 c1, Vec3 c2, float sx, float sy) {
    m.texture_type = Texture_Checkerboard;
    m.checkerboard_color1 = c1;
    m.checkerboard_color2 = c2;
    m.texture_scale_x = sx;
    m.texture_scale_y = sy;
}
void emitter(Material &m, float intensity) {
    m.emissive = true;
    m.light_color = Vec3(1.f);
    m.light_intensity = intensity;
}
void glass(Material &m, Vec3 kd, float ior) {
    m.type = Material_Glass;
    m.diffuse_material = kd;
    m.index_medium = ior;
}
void mirror(Material &m, Vec3 kd) {
    m.type = Material_Mirror;
    m.diffuse_material = kd;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// containers
// ------------------------------------------------------------------------------------------------
std::string Scene::path(const std::string &rel) const {
    if (asset_root.empty()) return rel;
    return asset_root.back() == '/' ? asset_root + rel : asset_root + "/" + rel;
}

void Scene::loadSkybox(const std::string &filename) { ppmLoader::load_ppm(skybox, path(filename)); }

int Scene::load_texture(const std::string &filename) {
    textures.emplace_back();
    ppmLoader::load_ppm(textures.back(), path(filename));
    return (int)textures.size() - 1;
}

int Scene::load_normal_map(const std::string &filename) {
    normals.emplace_back();
    ppmLoader::load_ppm(normals.back(), path(filename));
    return (int)normals.size() - 1;
}

// Scene::clear (Scene.h:181-188) leaves skybox and dark_sky alone; so does this.
void Scene::clear() {
    meshes.clear();
    spheres.clear();
    squares.clear();
    lights.clear();
    textures.clear();
    normals.clear();
}

void Scene::computeKDTrees() {
    for (Mesh &m : meshes) m.computeKDTree();
}

Square &Scene::new_square() { squares.emplace_back(); return squares.back(); }
Sphere &Scene::new_sphere(Vec3 c, float r) { spheres.emplace_back(c, r); return spheres.back(); }
Light &Scene::new_light(Vec3 pos, float radius) {
    lights.emplace_back();
    Light &l = lights.back();
    l.pos = pos;
    l.radius = radius;
    l.powerCorrection = 2.f;
    l.material = Vec3(1.f, 1.f, 1.f);
    return l;
}
Mesh &Scene::new_mesh(const std::string &off_file) {
    meshes.emplace_back();
    meshes.back().loadOFF(path(off_file));
    return meshes.back();
}

// Scene::addBox (Scene.h:92-146): unit quads (width = height = 1 whatever `size` is — only the
// corner -size/2 scales), one per enabled face in the order bottom, top, front, back, left, right;
// `rotation` is accepted and ignored, `facing_out` only flips the unused m_normal member.
void Scene::addBox(std::vector<Material> const &materials, bool faces[6], Vec3 const &pos, Vec3 const /*rotation*/,
                   float const size, bool facing_out) {
    const Vec3 corner(-size / 2.);
    const Vec3 right(size, 0.f, 0.f), up(0.f, 0.f, size);
    static const float rx[6] = {0.f, 180.f, 90.f, -90.f, 90.f, 90.f};
    static const float ry[6] = {0.f, 0.f, 0.f, 0.f, 90.f, -90.f};
    size_t k = 0;
    for (int f = 0; f < 6; ++f) {
        if (!faces[f]) continue;
        Square &s = new_square();
        s.setQuad(corner, right, up, 1.f, 1.f);
        if (f >= 1) s.rotate_x(rx[f]);
        if (f >= 4) s.rotate_y(ry[f]);
        s.translate(pos);
        s.build_arrays();
        if (!facing_out) s.m_normal *= -1.f;
        s.material = materials[k++];
    }
}

// ------------------------------------------------------------------------------------------------
// builders. Geometry, transform order and material values are the reference's (cited per scene);
// fields the tracer never reads (specular, shininess) are left at their defaults.
// ------------------------------------------------------------------------------------------------

// Scene.h:358-382
void Scene::setup_single_sphere() {
    clear();
    loadSkybox("img/textures/space.ppm");
    new_light(Vec3(-5, 5, 5), 2.5f);
    mirror(new_sphere(Vec3(0.f, 0.f, 0.f), 1.f).material, Vec3(1.f));
}

// Scene.h:384-419
void Scene::setup_single_square() {
    clear();
    dark_sky = false;
    new_light(Vec3(-5, 5, 5), 2.5f);
    {
        Square &s = new_square();
        s.setQuad(Vec3(-1.f, -1.f, 0.f), Vec3(1.f, 0.f, 0.f), Vec3(0.f, 1.f, 0.f), 6.f, 2.f);
        s.build_arrays();
        s.material.diffuse_material = Vec3(1.f, 0.f, 0.f);
    }
    {
        Square &s = new_square();
        centred_quad(s);
        place(s, {T(0, 0, -2), S(2, 2, 1), RY(-90)});
        s.material.diffuse_material = Vec3(0.f, 1.f, 0.f);
    }
}

// Scene.h:421-619. No Light entries: the room is lit by the emissive bottom face of a small box
// under the ceiling. Wall geometry depends on the aspect ratio.
void Scene::setup_cornell_box(float ar) {
    clear();
    skybox = ppmLoader::ImageRGB();
    const int brick = load_texture("img/planeTextures/brickwall.ppm");
    const int brick_n = load_normal_map("img/normalMaps/brickwall_normal.ppm");
    const int floor_n = load_normal_map("img/normalMaps/n1.ppm");
    const int sand = load_texture("img/planeTextures/sand.ppm");
    load_normal_map("img/normalMaps/water_normal.ppm");  // loaded by the reference, never bound

    Material white;
    white.diffuse_material = Vec3(0.9f);
    Material lamp;
    emitter(lamp, 60.f);
    std::vector<Material> box = {lamp, white, white, white, white};
    bool faces[6] = {true, false, true, true, true, true};
    addBox(box, faces, Vec3(0.f, 1.95f, 0.f), Vec3(45.f), 1.f, false);

    auto bricks = [&](Material &m) {
        m.texture_type = Texture_Image;
        m.set_texture(&textures[brick]);
        m.set_normals(&normals[brick_n]);
    };
    const float w2 = 2. * ar;      // half-extent scale of the long walls
    const float zoff = -2. * (-ar); // side walls sit at x = +-2*ar
    {   // back wall
        Square &s = new_square();
        centred_quad(s);
        place(s, {S(w2, 2, 1), T(0, 0, -2)});
        s.material.diffuse_material = Vec3(1.f);
        bricks(s.material);
        s.material.texture_scale_x = 1. * ar;
        s.material.texture_scale_y = 1.f;
    }
    {   // left wall
        Square &s = new_square();
        centred_quad(s);
        place(s, {RX(180), S(2, 2, 1), T(0, 0, zoff), RY(90)});
        s.material.diffuse_material = Vec3(1.f, 0.f, 0.f);
        bricks(s.material);
    }
    {   // right wall
        Square &s = new_square();
        centred_quad(s);
        place(s, {RX(180), T(0, 0, zoff), S(2, 2, 1), RY(-90)});
        s.material.diffuse_material = Vec3(0.f, 1.f, 0.f);
        bricks(s.material);
    }
    {   // floor
        Square &s = new_square();
        centred_quad(s);
        place(s, {T(0, 0, -2), S(w2, 2, 1), RX(-90)});
        s.material.diffuse_material = Vec3(246. / 255., 204. / 255., 162. / 255.);
        s.material.texture_type = Texture_Image;
        s.material.set_texture(&textures[sand]);
        s.material.set_normals(&normals[floor_n]);
    }
    {   // ceiling
        Square &s = new_square();
        centred_quad(s);
        place(s, {T(0, 0, -2), S(w2, 2, 1), RX(90)});
        s.material.diffuse_material = Vec3(1.f);
        checker(s.material, Vec3(0.95f), Vec3(0.5f), 8. * ar, 8.f);
    }
    {   // front wall (behind the camera)
        Square &s = new_square();
        centred_quad(s);
        place(s, {T(0, 0, -2), S(w2, 2, 1), RY(180)});
        s.material.diffuse_material = Vec3(1.f);
        bricks(s.material);
    }
    {
        Sphere &s = new_sphere(Vec3(1.0f, -1.25f, 0.5f), 0.75f);
        glass(s.material, Vec3(1.f), 1.4f);
        s.material.transparency = 1.0f;
    }
    {
        Sphere &s = new_sphere(Vec3(-1.0f, -1.25f, -0.5f), 0.75f);
        mirror(s.material, Vec3(0.7f));
        s.material.index_medium = 0.f;
    }
}

// Scene.h:714-827
void Scene::setup_mesh() {
    clear();
    loadSkybox("img/textures/space.ppm");
    new_light(Vec3(0.0f, 3.f, 2.0f));
    new_sphere(Vec3(0.f, 0.f, -16.f), 2.f).material.diffuse_material = Vec3(0.1f, 0.6f, 0.2f);
    mirror(new_sphere(Vec3(4.f, 0.f, -8.f), 2.f).material, Vec3(0.8f));
    {
        Mesh &m = new_mesh("mesh/blob-closed.off");
        place(m, {T(0.f, 0.9f, -4.f), S(1.5f), RX(180), RY(180)});
        glass(m.material, Vec3(0.1f, 0.2f, 0.5f), 1.333f);
        m.material.transparency = 0.9f;
    }
    new_sphere(Vec3(0.2f, -1.f, -4.8f), 0.3f).material.diffuse_material = Vec3(1.f);    // eye
    new_sphere(Vec3(0.2f, -1.f, -4.55f), 0.1f).material.diffuse_material = Vec3(0.f);   // pupil
    new_sphere(Vec3(-0.7f, -1.f, -4.95f), 0.3f).material.diffuse_material = Vec3(1.f);
    new_sphere(Vec3(-0.7f, -1.f, -4.7f), 0.1f).material.diffuse_material = Vec3(0.f);
    {
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(50, 50, 1), RX(-90)});
        s.material.diffuse_material = Vec3(0.8f, 0.8f, 0.f);
    }
    computeKDTrees();
}

// Scene.h:621-712
void Scene::setup_rt_in_a_weekend() {
    clear();
    loadSkybox("img/textures/sky.ppm");
    const int sun = load_texture("img/sphereTextures/s2.ppm");
    new_light(Vec3(0.0f, 3.f, -8.0f));
    new_light(Vec3(-4.f, 3.f, -8.0f));
    new_light(Vec3(4.f, 3.f, -8.0f));
    glass(new_sphere(Vec3(-4.f, 0.f, -8.f), 2.f).material, Vec3(0.8f), 1.5f);
    {
        Sphere &s = new_sphere(Vec3(0.f, 0.5f, -8.f), 1.5f);
        s.material.diffuse_material = Vec3(0.1f, 0.2f, 0.5f);
        s.material.texture_type = Texture_Image;
        s.material.set_texture(&textures[sun]);
        s.material.emissive = true;        // textured emitter: light_color stays 0, texture * 15
        s.material.light_intensity = 15.f;
        s.material.motion_blur_translation = Vec3(0.f, 1.f, 0.f);
    }
    mirror(new_sphere(Vec3(4.f, 0.f, -8.f), 2.f).material, Vec3(0.8f));
    {
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(50, 50, 1), RX(-90)});
        s.material.diffuse_material = Vec3(0.1f, 0.2f, 0.5f);
        checker(s.material, Vec3(1.f), Vec3(0.1f, 0.2f, 0.5f), 100.f, 100.f);
    }
}

// Scene.h:829-924. 3 fixed + 79 random spheres over a floor, one light, gradient sky. Every
// random sphere gets motion blur (0, height, 0). Randomness: random_float() (scene stream, see
// Functions.h) and rand() % 3 for the material type. Multi-argument Vec3(...) constructors are
// drawn right to left, as g++ evaluates the reference's.
void Scene::setup_random_spheres() {
    clear();
    dark_sky = false;
    new_light(Vec3(-1.0f, 8.f, 2.0f));
    {
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -4), S(100, 100, 1), RX(-90)});
        s.material.diffuse_material = Vec3(0.8f, 0.8f, 0.f);
    }
    mirror(new_sphere(Vec3(-3.f, 0.f, -22.f), 4.f).material, Vec3(0.8f));
    mirror(new_sphere(Vec3(4.f, -2.f, -15.f), 2.f).material, Vec3(0.8f));
    glass(new_sphere(Vec3(-1.f, -2.5f, -8.f), 1.5f).material, Vec3(0.8f), 1.0f);

    auto rgb01 = []() {
        const float b = random_float(0.f, 1.f), g = random_float(0.f, 1.f), r = random_float(0.f, 1.f);
        return Vec3(r, g, b);
    };
    for (int i = 0; i < 79; ++i) {
        const float height = random_float(0.25f, 1.f);
        const float radius = random_float(0.25f, 1.5f);
        const int type = rand() % 3;
        const float z = random_float(-50.f, -2.f);
        const float x = random_float(-30.f, 30.f);
        Sphere &s = new_sphere(Vec3(x, -4 + radius + height, z), radius);
        Material &m = s.material;
        switch (type) {
            case 0:
                m.type = Material_Mirror;
                m.diffuse_material = rgb01();
                m.specular_material = rgb01();
                m.shininess = random_float(32.f, 100.f);
                break;
            case 1:
                m.type = Material_Glass;
                m.diffuse_material = Vec3(random_float(0.7f, 1.f));
                m.specular_material = Vec3(random_float(0.7f, 1.f));
                m.shininess = random_float(32.f, 70.f);
                m.transparency = random_float(0.7f, 1.f);
                m.index_medium = random_float(1.f, 2.f);
                break;
            default:
                m.diffuse_material = rgb01();
                m.specular_material = rgb01();
                m.shininess = random_float(0.f, 30.f);
                break;
        }
        m.motion_blur_translation = Vec3(0.f, height, 0.f);
    }
}

// Scene.h:926-998
void Scene::setup_debug_refraction() {
    clear();
    dark_sky = false;
    new_light(Vec3(-1.0f, 8.f, 2.0f));
    const float at[4][2] = {{-2, 2}, {-2, -2}, {2, 2}, {2, -2}};
    const Vec3 col[4] = {Vec3(1, 0, 0), Vec3(0, 1, 0), Vec3(0, 0, 1), Vec3(1, 1, 1)};
    for (int k = 0; k < 4; ++k) {
        Square &s = new_square();
        centred_quad(s);
        place(s, {S(2, 2, 1), T(at[k][0], at[k][1], -2)});
        s.material.diffuse_material = col[k];
    }
    Sphere &s = new_sphere(Vec3(0.f, 0.f, 0.f), 0.75f);
    glass(s.material, Vec3(1.f), 1.4f);
    s.material.transparency = 1.0f;
}

// Scene.h:1000-1078
void Scene::setup_flamingo() {
    clear();
    dark_sky = false;
    new_light(Vec3(-1.0f, 8.f, 2.0f));
    new_light(Vec3(1.0f, 8.f, 2.0f));
    {
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(50, 50, 1), RX(-90)});
        s.material.diffuse_material = Vec3(0.8f, 0.8f, 0.f);
        checker(s.material, Vec3(0.8f, 0.8f, 0.f), Vec3(0.6f, 0.6f, 0.f), 100.f, 100.f);
    }
    glass(new_sphere(Vec3(-4.f, 0.f, -8.f), 2.f).material, Vec3(0.8f), 1.5f);
    mirror(new_sphere(Vec3(4.f, 0.f, -8.f), 2.f).material, Vec3(0.8f));
    {
        Mesh &m = new_mesh("mesh/flamingo_lowpoly_colored.off");
        place(m, {S(2.5f), RX(90), RY(90), RZ(180), T(0.f, 1.f, -8.f)});
        m.material.diffuse_material = Vec3(0.1f, 0.2f, 0.5f);
    }
    computeKDTrees();
}

// Scene.h:1080-1207
void Scene::setup_raccoon() {
    clear();
    loadSkybox("img/textures/sky.ppm");
    const int fire = load_texture("img/sphereTextures/s2.ppm");
    const int wind = load_texture("img/sphereTextures/s4.ppm");
    const int water = load_texture("img/sphereTextures/s7.ppm");
    new_light(Vec3(-1.0f, 8.f, 2.0f));
    {   // carpet, checker part
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(2, 4, 1), RX(-90), T(0, 0, -4)});
        s.material.diffuse_material = Vec3(0.5f, 0.f, 0.5f);
        checker(s.material, Vec3(0.5f, 0.f, 0.5f), Vec3(0.6f, 0.f, 0.6f), 16.f, 16.f);
    }
    {   // carpet, red border
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(2.5f, 5, 1), RX(-90), T(0.f, -0.0001f, -3.5f)});
        s.material.diffuse_material = Vec3(0.9f, 0.2f, 0.f);
    }
    {
        Mesh &m = new_mesh("mesh/raccoon_low_poly_colored.off");
        place(m, {RY(-90), S(2.f), T(0.f, -2.f, -5.f)});
        m.material.diffuse_material = Vec3(0.1f, 0.2f, 0.5f);
    }
    {
        Mesh &m = new_mesh("mesh/magic_staff_low_poly_colored.off");
        place(m, {RY(-90), RZ(90), S(0.15f), T(1.f, 0.2f, -2.7f)});
        m.material.diffuse_material = Vec3(0.1f, 0.2f, 0.5f);
    }
    {   // staff orb
        Sphere &s = new_sphere(Vec3(-1.85f, 0.35f, -2.7f), 0.14f);
        glass(s.material, Vec3(0.451f, 0.6627f, 0.7608f), 1.5f);
        s.material.transparency = 0.65f;
    }
    auto orb = [&](Sphere &s, int tex) {
        s.material.texture_type = Texture_Image;
        s.material.set_texture(&textures[tex]);
    };
    {
        Sphere &s = new_sphere(Vec3(4.f, 3.f, -8.f), 1.3f);
        mirror(s.material, Vec3(0.8f, 0.f, 0.f));
        orb(s, fire);
    }
    {
        Sphere &s = new_sphere(Vec3(-4.f, 2.f, -5.f), 0.9f);
        glass(s.material, Vec3(1.f), 1.0f);
        s.material.transparency = 0.4f;
        orb(s, wind);
    }
    {
        Sphere &s = new_sphere(Vec3(-0.2f, 3.f, -1.f), 1.4f);
        glass(s.material, Vec3(0.5f, 0.53f, 0.8f), 1.0f);
        s.material.transparency = 0.8f;
        orb(s, water);
    }
    computeKDTrees();
}

// Scene.h:1209-1262 — BASELINE config 3: pond.off (11 110 face-coloured triangles), low-poly
// flamingo (832 vertex-coloured triangles), mirror water quad, sky texture.
void Scene::setup_flamingo_pond() {
    clear();
    loadSkybox("img/textures/sky.ppm");
    new_light(Vec3(-1.0f, 8.f, -19.0f));
    {
        Mesh &m = new_mesh("mesh/pond.off");
        place(m, {S(3.f), T(1.f, -5.f, -3.f)});
        m.material.diffuse_material = Vec3(0.1f, 0.2f, 0.5f);
    }
    {
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(5, 3.5f, 1), RX(-90), T(1.f, 0.f, 2.8f)});
        mirror(s.material, Vec3(0.5f, 0.53f, 0.8f));
    }
    {
        Mesh &m = new_mesh("mesh/flamingo_lowpoly_colored.off");
        place(m, {S(0.8f), RX(90), RY(115), RZ(180), T(3.f, -1.2f, -1.f)});
        m.material.diffuse_material = Vec3(0.1f, 0.2f, 0.5f);
    }
    computeKDTrees();
}

// Scene.h:1264-1327
void Scene::setup_flamingo_lake() {
    clear();
    loadSkybox("img/textures/sky.ppm");
    load_texture("img/sphereTextures/s2.ppm");  // loaded, never bound
    const int water_n = load_normal_map("img/normalMaps/water_normal.ppm");
    new_light(Vec3(1.0f, 2.f, 1.0f));
    {
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(50, 50, 1), RX(-90)});
        s.material.diffuse_material = Vec3(0.1f, 0.5f, 0.1f);
        checker(s.material, Vec3(1.f), Vec3(0.1f, 0.2f, 0.5f), 100.f, 100.f);
    }
    {   // water: glass quad with a normal map, texture scale 10
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(50, 50, 1), RX(-90), T(0.f, 0.3f, 0.f)});
        glass(s.material, Vec3(0.1f, 0.2f, 0.5f), 1.0f);
        s.material.texture_scale_x = 10.f;
        s.material.texture_scale_y = 10.f;
        s.material.set_normals(&normals[water_n]);
    }
    {
        Mesh &m = new_mesh("mesh/flamingo_float.off");
        m.centerAndScaleToUnit();
        place(m, {RX(270), T(0.f, -1.5f, -1.f)});
        m.material.diffuse_material = Vec3(237. / 255., 149. / 255., 218. / 255.);
    }
    computeKDTrees();
}

// Scene.h:1329-1882 — BASELINE config 4: an enclosed tiled pool hall. 28 quads (12 of them
// emissive panels at intensity 30), two tiny spheres, three meshes, no Light entries.
void Scene::setup_backrooms_pool() {
    clear();
    loadSkybox("img/textures/sky.ppm");
    const int tiles = load_texture("img/planeTextures/white_pool_tiles.ppm");
    const int tiles_n = load_normal_map("img/normalMaps/pool_tiles_normal.ppm");
    const int water_n = load_normal_map("img/normalMaps/water_normal.ppm");
    const float panel_power = 30.f;

    auto panel = [&](std::initializer_list<Op> ops) {
        Square &s = new_square();
        low_quad(s);
        place(s, ops);
        s.material.diffuse_material = Vec3(1.f);
        emitter(s.material, panel_power);
    };
    auto tiled = [&](std::initializer_list<Op> ops, float sx, float sy) {
        Square &s = new_square();
        low_quad(s);
        place(s, ops);
        s.material.diffuse_material = Vec3(0.1f, 0.5f, 0.1f);
        s.material.texture_type = Texture_Image;
        s.material.texture_scale_x = sx;
        s.material.texture_scale_y = sy;
        s.material.set_texture(&textures[tiles]);
        s.material.set_normals(&normals[tiles_n]);
    };

    for (float z : {-12.75f, -8.75f, -4.75f, -0.75f})  // four ceiling panels
        panel({T(0, 0, -2), S(0.5f, 0.5f, 1), RX(90), T(0.f, 2.95f, z)});
    {   // water surface
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(4, 8, 1), RX(-90), T(0.f, -0.75f, 0.f)});
        glass(s.material, Vec3(170. / 255., 213. / 255., 219. / 255.), 1.0f);
        s.material.transparency = 0.99f;
        s.material.set_normals(&normals[water_n]);
    }
    tiled({T(0, 0, -2), S(4, 8, 1), RX(-90), T(0.f, -1.f, 0.f)}, 1.f, 2.f);                 // pool floor
    {   // ceiling, plain
        Square &s = new_square();
        low_quad(s);
        place(s, {T(0, 0, -2), S(4, 8, 1), RX(90), T(0.f, 3.f, -12.75f)});
        s.material.diffuse_material = Vec3(0.8f);
    }
    tiled({T(0, 0, -2), S(0.5f, 8, 1), RX(-90), RZ(90), T(2.f, -2.5f, 0.f)}, 0.25f, 2.f);    // right pool wall
    tiled({T(0, 0, -2), S(2, 8, 1), RX(-90), RZ(90), T(2.f, 4.f, 0.f)}, 1.f, 2.f);           // right upper wall
    tiled({T(0, 0, -2), S(2, 8, 1), RX(-90), RZ(-90), T(-2.f, 4.f, 0.f)}, 1.f, 2.f);         // left upper wall
    tiled({T(0, 0, -2), S(0.5f, 8, 1), RX(-90), RZ(-90), T(-2.f, -2.5f, 0.f)}, 0.25f, 2.f);  // left pool wall
    tiled({T(0, 0, -2), S(1, 8, 1), RX(-90), T(5.f, 0.f, 0.f)}, 1.f, 2.f);                   // right deck
    tiled({T(0, 0, -2), S(1, 8, 1), RX(90), T(5.f, 0.f, -12.75f)}, 1.f, 2.f);                // right deck ceiling
    tiled({T(0, 0, -2), S(1, 8, 1), RX(-90), T(-5.f, 0.f, 0.f)}, 1.f, 2.f);                  // left deck
    tiled({T(0, 0, -2), S(1, 8, 1), RX(90), T(5.f, 0.f, -12.75f)}, 1.f, 2.f);                // (duplicate in the reference)
    tiled({T(0, 0, -2), S(1, 8, 1), RX(90), T(-5.f, 0.f, -12.75f)}, 1.f, 2.f);               // left deck ceiling
    tiled({T(0, 0, -2), S(8, 2, 1), RY(-90), T(4.f, -1.6f, -6.4f)}, 2.f, 1.f);               // right side wall
    for (float z : {-0.75f, -4.75f, -8.75f, -12.75f})
        panel({T(0, 0, -2), S(0.5f, 0.5f, 1), RY(-90), T(3.95f, 0.9f, z)});
    tiled({T(0, 0, -2), S(8, 2, 1), RY(90), T(-4.f, -1.6f, -6.4f)}, 2.f, 1.f);               // left side wall
    for (float z : {-0.75f, -4.75f, -8.75f, -12.75f})
        panel({T(0, 0, -2), S(0.5f, 0.5f, 1), RY(90), T(-3.95f, 0.8f, z)});
    tiled({T(0, 0, -2), S(8, 8, 1), RX(-180), T(0.f, 4.f, 0.f)}, 2.f, 2.f);                  // front
    tiled({T(0, 0, -2), S(8, 8, 1), T(0.f, -3.f, -12.f)}, 2.f, 2.f);                         // back

    {
        Mesh &m = new_mesh("mesh/flamingo_float_colored.off");
        m.centerAndScaleToUnit();
        place(m, {RX(0), RY(225), T(-0.5f, -1.35f, -2.f), S(1.8f)});
        m.material.diffuse_material = Vec3(237. / 255., 149. / 255., 218. / 255.);
    }
    new_sphere(Vec3(0.05f, -1.4f, -3.1f), 0.05f).material.diffuse_material = Vec3(1.f);   // eye
    new_sphere(Vec3(0.05f, -1.4f, -3.05f), 0.01f).material.diffuse_material = Vec3(0.f);  // pupil
    {
        Mesh &m = new_mesh("mesh/rubber_duck_colored.off");
        m.centerAndScaleToUnit();
        place(m, {RY(-35), T(2.f, -1.65f, -2.f), S(1.3f)});
        m.material.diffuse_material = Vec3(1.f, 1.f, 0.f);
    }
    {
        Mesh &m = new_mesh("mesh/pool_ladder.off");
        m.centerAndScaleToUnit();
        place(m, {RY(90), T(-3.f, -1.445f, -3.f), S(1.3f)});
        mirror(m.material, Vec3(0.5f, 0.5f, 0.5f));
    }
    computeKDTrees();
}

// BASELINE.json config 5 — no such builder exists in the reference. Composed
// from its parts: setup_random_spheres (79 motion-blurred spheres) + triceratops.off +
// gorilla.off with KD-trees. The oracle composes the same scene through the reference's own
// classes (oracle/ref_driver.cpp : setup_config5).
void Scene::setup_motion_blur_meshes() {
    setup_random_spheres();
    {
        Mesh &m = new_mesh("mesh/triceratops.off");
        m.centerAndScaleToUnit();
        place(m, {S(2.0f), RY(200.f), T(-3.2f, -2.6f, -3.5f)});
        m.material.diffuse_material = Vec3(0.35f, 0.55f, 0.25f);
    }
    {
        Mesh &m = new_mesh("mesh/gorilla.off");
        m.centerAndScaleToUnit();
        place(m, {S(1.8f), RY(160.f), T(3.0f, -2.3f, -4.5f)});
        m.material.diffuse_material = Vec3(0.45f, 0.35f, 0.3f);
    }
    computeKDTrees();
}

bool Scene::setup_by_id(int id, float ar) {
    switch (id) {
        case 0: setup_single_sphere(); return true;
        case 1: setup_single_square(); return true;
        case 2: setup_cornell_box(ar); return true;
        case 3: setup_mesh(); return true;
        case 4: setup_rt_in_a_weekend(); return true;
        case 5: setup_random_spheres(); return true;
        case 6: setup_debug_refraction(); return true;
        case 7: setup_flamingo(); return true;
        case 8: setup_raccoon(); return true;
        case 9: setup_flamingo_pond(); return true;
        case 10: setup_backrooms_pool(); return true;
        case 11: setup_flamingo_lake(); return true;
        case 100: setup_motion_blur_meshes(); return true;
        default: return false;
    }
}

// ------------------------------------------------------------------------------------------------
// flatten
// ------------------------------------------------------------------------------------------------
namespace {
template <class Img> int image_index(const std::vector<Img> &pool, const Img *p) {
    if (!p || pool.empty()) return -1;
    const std::ptrdiff_t d = p - pool.data();
    return (d >= 0 && (size_t)d < pool.size()) ? (int)d : -1;
}
void put3(float *dst, const Vec3 &v) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; }

RtMaterial flat_material(const Scene &s, const Material &m) {
    RtMaterial r{};
    r.type = (int32_t)m.type;
    r.texture_type = (int32_t)m.texture_type;
    put3(r.diffuse, m.diffuse_material);
    r.transparency = m.transparency;
    r.index_medium = m.index_medium;
    put3(r.checker1, m.checkerboard_color1);
    put3(r.checker2, m.checkerboard_color2);
    r.texture_scale_x = m.texture_scale_x;
    r.texture_scale_y = m.texture_scale_y;
    r.emissive = m.emissive ? 1 : 0;
    put3(r.light_color, m.light_color);
    r.light_intensity = m.light_intensity;
    r.image = m.texture_type == Texture_Image ? image_index(s.textures, m.image) : -1;
    r.normal_map = m.has_normal_map ? image_index(s.normals, m.normals) : -1;
    put3(r.motion, m.motion_blur_translation);
    return r;
}
RtImage flat_image(const ppmLoader::ImageRGB &im) {
    RtImage r{0, 0, nullptr, 0};
    if (im.w >= 1 && im.h >= 1 && im.data.size() >= (size_t)im.w * (size_t)im.h) {
        r.w = im.w;
        r.h = im.h;
        r.rgb = (const uint8_t *)im.data.data();
        r.content_id = im.content_id;
    }
    return r;
}
}  // namespace

void Scene::flatten(FlatScene &f) const {
    f = FlatScene();
    for (const Sphere &s : spheres) {
        RtSphere r{};
        put3(r.center, s.m_center);
        r.radius = s.m_radius;
        r.material = flat_material(*this, s.material);
        f.spheres.push_back(r);
    }
    for (const Square &s : squares) {
        if (s.vertices.size() < 4) hai719::fatal("Square without setQuad()");
        RtSquare r{};
        put3(r.v0, s.vertices[0].position);
        put3(r.v1, s.vertices[1].position);
        put3(r.v3, s.vertices[3].position);
        put3(r.right, s.m_right_vector);
        put3(r.up, s.m_up_vector);
        r.material = flat_material(*this, s.material);
        f.squares.push_back(r);
    }
    for (const Light &l : lights) {
        RtLight r{};
        put3(r.pos, l.pos);
        r.radius = l.radius;
        put3(r.color, l.material);
        f.lights.push_back(r);
    }
    for (const auto &t : textures) f.textures.push_back(flat_image(t));
    for (const auto &t : normals) f.normal_maps.push_back(flat_image(t));

    const size_t nm = meshes.size();
    f.positions.resize(nm); f.vert_colors.resize(nm); f.face_colors.resize(nm);
    f.triangles.resize(nm); f.nodes.resize(nm); f.leaf_refs.resize(nm);
    for (size_t i = 0; i < nm; ++i) {
        const Mesh &m = meshes[i];
        RtSceneMesh r{};
        r.n_vertices = (uint32_t)m.vertices.size();
        r.n_triangles = (uint32_t)m.triangles.size();
        for (const MeshVertex &v : m.vertices) for (unsigned k = 0; k < 3; ++k) f.positions[i].push_back(v.position[k]);
        for (const MeshTriangle &t : m.triangles) for (unsigned k = 0; k < 3; ++k) f.triangles[i].push_back(t[k]);
        r.positions = f.positions[i].data();
        r.triangles = f.triangles[i].data();
        r.color_type = (int32_t)m.colorType;
        if (m.colorType == ColorType_Vertex) {
            for (const Vec3 &c : m.vertColors) for (unsigned k = 0; k < 3; ++k) f.vert_colors[i].push_back(c[k]);
            r.vert_colors = f.vert_colors[i].data();
        } else if (m.colorType == ColorType_Face) {
            for (const Vec3 &c : m.faceColors) for (unsigned k = 0; k < 3; ++k) f.face_colors[i].push_back(c[k]);
            r.face_colors = f.face_colors[i].data();
        }
        if (m.kdtree) {
            const KDTree &kd = *m.kdtree;
            put3(r.root_bmin, kd.aabb.p0);
            put3(r.root_bmax, kd.aabb.p1);
            r.n_nodes = (uint32_t)kd.nodes.size();
            r.nodes = kd.nodes.data();
            r.n_leaf_refs = (uint32_t)kd.leaf_refs.size();
            r.leaf_refs = kd.leaf_refs.data();
        } else {
            // Mesh::intersectOld (Mesh.h:257-277): mesh box as the gate, then every triangle in
            // file order == one leaf holding them all.
            put3(r.root_bmin, m.aabb.p0);
            put3(r.root_bmax, m.aabb.p1);
            RtKdNode n{};
            put3(n.bmin, m.aabb.p0);
            put3(n.bmax, m.aabb.p1);
            n.is_leaf = 1; n.skip = 1; n.first_ref = 0; n.n_refs = r.n_triangles;
            f.nodes[i].push_back(n);
            for (uint32_t t = 0; t < r.n_triangles; ++t)
                f.leaf_refs[i].push_back(RtTriRef{{m.triangles[t][0], m.triangles[t][1], m.triangles[t][2]}, t});
            r.n_nodes = 1;
            r.nodes = f.nodes[i].data();
            r.n_leaf_refs = r.n_triangles;
            r.leaf_refs = f.leaf_refs[i].data();
        }
        r.material = flat_material(*this, m.material);
        f.meshes.push_back(r);
    }
    RtSceneDesc &d = f.desc;
    d = RtSceneDesc{};
    d.abi_version = HAI719_RT_ABI_VERSION;
    d.n_spheres = (uint32_t)f.spheres.size();       d.spheres = f.spheres.data();
    d.n_squares = (uint32_t)f.squares.size();       d.squares = f.squares.data();
    d.n_meshes = (uint32_t)f.meshes.size();         d.meshes = f.meshes.data();
    d.n_lights = (uint32_t)f.lights.size();         d.lights = f.lights.data();
    d.n_textures = (uint32_t)f.textures.size();     d.textures = f.textures.data();
    d.n_normal_maps = (uint32_t)f.normal_maps.size(); d.normal_maps = f.normal_maps.data();
    d.skybox = flat_image(skybox);
    d.dark_sky = dark_sky ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// canonical dump (word-for-word the layout of oracle/ref_driver.cpp : dump_scene)
// ------------------------------------------------------------------------------------------------
namespace {
struct Words {
    std::vector<uint32_t> &w;
    void u(uint32_t v) { w.push_back(v); }
    void f(float v) { uint32_t b; std::memcpy(&b, &v, 4); w.push_back(b); }
    void v3(const Vec3 &v) { f(v[0]); f(v[1]); f(v[2]); }
    void v3(const float *v) { f(v[0]); f(v[1]); f(v[2]); }
    void img(const ppmLoader::ImageRGB &im) {
        if (im.w < 1 || im.h < 1 || im.data.size() < (size_t)im.w * (size_t)im.h) { u(0); u(0); u(0); u(0); return; }
        uint64_t h = 1469598103934665603ull;
        const unsigned char *b = (const unsigned char *)im.data.data();
        for (size_t i = 0, n = (size_t)im.w * (size_t)im.h * 3; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        u((uint32_t)im.w); u((uint32_t)im.h); u((uint32_t)h); u((uint32_t)(h >> 32));
    }
    void material(const Scene &s, const Material &m) {
        u((uint32_t)m.type); u((uint32_t)m.texture_type);
        v3(m.diffuse_material); f(m.transparency); f(m.index_medium);
        v3(m.checkerboard_color1); v3(m.checkerboard_color2);
        f(m.texture_scale_x); f(m.texture_scale_y);
        u(m.emissive ? 1u : 0u);
        if (m.emissive) { v3(m.light_color); f(m.light_intensity); } else { v3(Vec3(0.f)); f(0.f); }
        u((uint32_t)(m.texture_type == Texture_Image ? image_index(s.textures, m.image) : -1));
        u((uint32_t)(m.has_normal_map ? image_index(s.normals, m.normals) : -1));
        u(m.has_normal_map ? 1u : 0u);
        v3(m.motion_blur_translation);
    }
};
}  // namespace

void Scene::dump(std::vector<uint32_t> &out) const {
    Words d{out};
    d.u(0x44533748u);
    d.u(1u);
    d.u(dark_sky ? 1u : 0u);
    d.img(skybox);
    d.u((uint32_t)textures.size());
    for (auto &t : textures) d.img(t);
    d.u((uint32_t)normals.size());
    for (auto &t : normals) d.img(t);
    d.u((uint32_t)lights.size());
    for (auto &l : lights) { d.v3(l.pos); d.f(l.radius); d.v3(l.material); }
    d.u((uint32_t)spheres.size());
    for (auto &s : spheres) { d.v3(s.m_center); d.f(s.m_radius); d.material(*this, s.material); }
    d.u((uint32_t)squares.size());
    for (auto &s : squares) {
        for (int k = 0; k < 4; ++k) d.v3(s.vertices[k].position);
        d.v3(s.m_right_vector);
        d.v3(s.m_up_vector);
        d.material(*this, s.material);
    }
    d.u((uint32_t)meshes.size());
    for (auto &m : meshes) {
        d.u((uint32_t)m.vertices.size());
        d.u((uint32_t)m.triangles.size());
        d.u((uint32_t)m.colorType);
        d.u(m.kdtree ? 1u : 0u);
        d.v3(m.aabb.p0);
        d.v3(m.aabb.p1);
        d.material(*this, m.material);
        for (auto &v : m.vertices) d.v3(v.position);
        for (auto &t : m.triangles) { d.u(t[0]); d.u(t[1]); d.u(t[2]); }
        if (m.colorType == ColorType_Vertex) for (auto &c : m.vertColors) d.v3(c);
        if (m.colorType == ColorType_Face) for (auto &c : m.faceColors) d.v3(c);
        if (m.kdtree) {
            const KDTree &kd = *m.kdtree;
            d.v3(kd.aabb.p0);
            d.v3(kd.aabb.p1);
            d.u((uint32_t)kd.nodes.size());
            for (size_t i = 0; i < kd.nodes.size(); ++i) {
                const RtKdNode &n = kd.nodes[i];
                d.v3(n.bmin);
                d.v3(n.bmax);
                // children of a pre-order node: left starts at i+1 (if the subtree is non-empty),
                // right starts where the left subtree ends
                const bool has_left = !n.is_leaf && (i + 1 < n.skip);
                const bool has_right = has_left ? (kd.nodes[i + 1].skip < n.skip) : false;
                // an inner node whose single child is its RIGHT one looks the same in pre-order;
                // the builder records which (bit 31 of first_ref) — see note in KDTree.cpp
                const bool only_right = !n.is_leaf && (n.first_ref & 0x80000000u);
                const bool left = only_right ? false : has_left;
                const bool right = only_right ? has_left : has_right;
                const bool leaf = !left && !right;
                d.u((leaf ? 1u : 0u) | (left ? 2u : 0u) | (right ? 4u : 0u));
                d.u(n.is_leaf ? n.n_refs : 0u);
                if (n.is_leaf)
                    for (uint32_t k = 0; k < n.n_refs; ++k) {
                        const RtTriRef &t = kd.leaf_refs[n.first_ref + k];
                        d.u(t.v[0]); d.u(t.v[1]); d.u(t.v[2]); d.u(t.tri_index);
                    }
            }
        }
    }
}
