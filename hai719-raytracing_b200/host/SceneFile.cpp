// Scene description files. The reference hard-codes its scenes in C++ (Scene.h:358-1882); this is a
// line-based text format that builds a Scene through the SAME host calls the builders make (Square::setQuad,
// Mesh::loadOFF/translate/scale/rotate_*, Scene::load_texture, ...), so a file that restates a built-in scene yields a
// bit-identical Scene::dump() (tests/test_host_scene.py). Errors are returned with their line number; nothing exits.
//
//   hai719scene 1
//   sky dark | gradient | image <file.ppm>               (several sky lines may be given; 'image' only loads the picture)
//   light <x> <y> <z> [<radius>]                         (Scene::new_light: radius 1.5, white, Scene.h:28-42)
//   texture <file.ppm>      normalmap <file.ppm>         (indexed in order of appearance, from 0)
//   material <name> [diffuse|glass|mirror] [kd r g b] [ks r g b] [ka r g b] [shininess s] [transparency t] [ior n]
//            [emissive r g b intensity] [checker r g b r g b sx sy] [image <texture> sx sy] [normals <normalmap>]
//            [motion x y z]
//   sphere <material> <cx> <cy> <cz> <radius>
//   square <material> <bottom-left xyz> <right xyz> <up xyz> <width> <height> [transforms]   (Square::setQuad)
//   mesh <material> <file.off> [center_unit] [transforms]
//   transforms, applied in the order written: translate x y z | scale x y z | rotate_x deg | rotate_y deg | rotate_z deg
//   '#' starts a comment. File names are relative to Scene::asset_root like the builders' ("img/...", "mesh/...").
#include "SceneFile.h"

#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "errors.h"

namespace hai719 {

namespace {

struct Cursor {
    std::vector<std::string> tok;
    size_t i = 0;
    int line = 0;
    bool more() const { return i < tok.size(); }
    const std::string &peek() const { return tok[i]; }
    std::string word(const char *what) {
        if (!more()) throw std::runtime_error(std::string("missing ") + what);
        return tok[i++];
    }
    float num(const char *what) {
        const std::string w = word(what);
        char *end = nullptr;
        const float v = std::strtof(w.c_str(), &end);
        if (end == w.c_str() || *end != '\0') throw std::runtime_error(std::string("'") + w + "' is not a number (" + what + ")");
        return v;
    }
    int index(const char *what, size_t count) {
        const float v = num(what);
        if (v < 0.f || v != (float)(int)v || (size_t)v >= count) throw std::runtime_error(std::string(what) + " index out of range");
        return (int)v;
    }
    Vec3 vec(const char *what) { const float x = num(what), y = num(what), z = num(what); return Vec3(x, y, z); }
};

// the builders' place(): transforms in list order, then Mesh::build_arrays()
void transforms(Cursor &c, Mesh &m) {
    while (c.more()) {
        const std::string k = c.word("transform");
        if (k == "translate") m.translate(c.vec("translate"));
        else if (k == "scale") m.scale(c.vec("scale"));
        else if (k == "rotate_x") m.rotate_x(c.num("angle"));
        else if (k == "rotate_y") m.rotate_y(c.num("angle"));
        else if (k == "rotate_z") m.rotate_z(c.num("angle"));
        else throw std::runtime_error("unknown transform '" + k + "'");
    }
    m.build_arrays();
}

struct MaterialSpec { Material m; int image = -1, normals = -1; };

Material resolve(Scene &s, const MaterialSpec &spec) {
    Material m = spec.m;
    if (spec.image >= 0) m.set_texture(&s.textures[(size_t)spec.image]);
    if (spec.normals >= 0) m.set_normals(&s.normals[(size_t)spec.normals]);
    return m;
}

}  // namespace

bool load_scene_file(Scene &scene, const std::string &filename, std::string *error) {
    std::ifstream in(filename.c_str());
    if (!in) { if (error) *error = "cannot open " + filename; return false; }
    const bool throws_before = fatal_throws();
    set_fatal_throws(true);   // loaders report through exceptions while a file is read, never exit()
    int line_no = 0;
    try {
        // pass 1: images (materials keep pointers into the pools, so the pools must be complete first)
        std::vector<std::string> lines;
        for (std::string ln; std::getline(in, ln);) lines.push_back(ln);
        scene.clear();
        scene.textures.clear();
        scene.normals.clear();
        scene.skybox = ppmLoader::ImageRGB();
        scene.dark_sky = true;
        std::vector<Cursor> stmts;
        bool header = false;
        for (const std::string &raw : lines) {
            ++line_no;
            std::istringstream ss(raw.substr(0, raw.find('#')));
            Cursor c;
            c.line = line_no;
            for (std::string t; ss >> t;) c.tok.push_back(t);
            if (c.tok.empty()) continue;
            if (!header) {
                if (c.tok.size() != 2 || c.tok[0] != "hai719scene" || c.tok[1] != "1") throw std::runtime_error("expected the header 'hai719scene 1'");
                header = true;
                continue;
            }
            if (c.tok[0] == "texture") { c.i = 1; scene.load_texture(c.word("file name")); if (c.more()) throw std::runtime_error("trailing tokens"); continue; }
            if (c.tok[0] == "normalmap") { c.i = 1; scene.load_normal_map(c.word("file name")); if (c.more()) throw std::runtime_error("trailing tokens"); continue; }
            stmts.push_back(c);
        }
        if (!header) { line_no = 0; throw std::runtime_error("empty file"); }

        // pass 2: everything else, in file order
        std::map<std::string, MaterialSpec> materials;
        for (Cursor &c : stmts) {
            line_no = c.line;
            const std::string kind = c.word("statement");
            if (kind == "sky") {
                const std::string how = c.word("sky mode");
                if (how == "dark") scene.dark_sky = true;
                else if (how == "gradient") scene.dark_sky = false;
                else if (how == "image") scene.loadSkybox(c.word("file name"));   // like Scene::loadSkybox: dark_sky is left alone
                else throw std::runtime_error("sky must be dark, gradient or image <file>");
            } else if (kind == "light") {
                const Vec3 pos = c.vec("light position");
                scene.lights.emplace_back();
                Light &l = scene.lights.back();
                l.pos = pos;
                l.radius = c.more() ? c.num("light radius") : 1.5f;
                l.powerCorrection = 2.f;
                l.material = Vec3(1.f, 1.f, 1.f);
            } else if (kind == "material") {
                const std::string name = c.word("material name");
                MaterialSpec spec;
                while (c.more()) {
                    const std::string k = c.word("material attribute");
                    if (k == "diffuse") spec.m.type = Material_Diffuse_Blinn_Phong;
                    else if (k == "glass") spec.m.type = Material_Glass;
                    else if (k == "mirror") spec.m.type = Material_Mirror;
                    else if (k == "kd") spec.m.diffuse_material = c.vec("kd");
                    else if (k == "ks") spec.m.specular_material = c.vec("ks");
                    else if (k == "ka") spec.m.ambient_material = c.vec("ka");
                    else if (k == "shininess") spec.m.shininess = c.num("shininess");
                    else if (k == "transparency") spec.m.transparency = c.num("transparency");
                    else if (k == "ior") spec.m.index_medium = c.num("ior");
                    else if (k == "motion") spec.m.motion_blur_translation = c.vec("motion");
                    else if (k == "emissive") { spec.m.emissive = true; spec.m.light_color = c.vec("emissive colour"); spec.m.light_intensity = c.num("intensity"); }
                    else if (k == "checker") {
                        spec.m.texture_type = Texture_Checkerboard;
                        spec.m.checkerboard_color1 = c.vec("checker colour 1");
                        spec.m.checkerboard_color2 = c.vec("checker colour 2");
                        spec.m.texture_scale_x = c.num("scale x"); spec.m.texture_scale_y = c.num("scale y");
                    } else if (k == "image") {
                        spec.m.texture_type = Texture_Image;
                        spec.image = c.index("texture", scene.textures.size());
                        spec.m.texture_scale_x = c.num("scale x"); spec.m.texture_scale_y = c.num("scale y");
                    } else if (k == "normals") spec.normals = c.index("normalmap", scene.normals.size());
                    else throw std::runtime_error("unknown material attribute '" + k + "'");
                }
                materials[name] = spec;
            } else if (kind == "sphere" || kind == "square" || kind == "mesh") {
                const std::string mname = c.word("material name");
                const auto it = materials.find(mname);
                if (it == materials.end()) throw std::runtime_error("unknown material '" + mname + "'");
                if (kind == "sphere") {
                    const Vec3 centre = c.vec("sphere centre");
                    const float r = c.num("radius");
                    scene.spheres.emplace_back(centre, r);
                    scene.spheres.back().material = resolve(scene, it->second);
                    if (c.more()) throw std::runtime_error("trailing tokens");
                } else if (kind == "square") {
                    const Vec3 bl = c.vec("bottom-left"), right = c.vec("right vector"), up = c.vec("up vector");
                    const float w = c.num("width"), h = c.num("height");
                    scene.squares.emplace_back();
                    Square &q = scene.squares.back();
                    q.setQuad(bl, right, up, w, h);
                    transforms(c, q);
                    q.material = resolve(scene, it->second);
                } else {
                    const std::string file = c.word("OFF file");
                    scene.meshes.emplace_back();
                    Mesh &m = scene.meshes.back();
                    m.loadOFF(scene.asset_root.empty() ? file : scene.asset_root + "/" + file);
                    if (c.more() && c.peek() == "center_unit") { c.word("center_unit"); m.centerAndScaleToUnit(); }
                    transforms(c, m);
                    m.material = resolve(scene, it->second);
                }
            } else {
                throw std::runtime_error("unknown statement '" + kind + "'");
            }
            if (c.more()) throw std::runtime_error("trailing tokens after '" + kind + "'");
        }
        scene.computeKDTrees();
    } catch (const std::exception &e) {
        set_fatal_throws(throws_before);
        if (error) *error = filename + ":" + std::to_string(line_no) + ": " + e.what();
        return false;
    }
    set_fatal_throws(throws_before);
    return true;
}

}  // namespace hai719
