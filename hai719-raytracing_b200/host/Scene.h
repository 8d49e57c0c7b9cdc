// Host API — Scene (src/Scene.h:56-1882) without OpenGL and without the CPU tracer.
//
// Same containers and builder names as the reference, so code that populated a reference Scene
// populates this one. The integrator methods (rayTrace, rayTraceRecursive, computeIntersection,
// computeShadow, skyboxTexture; Scene.h:149-161,190-350) are NOT host functions here: they are
// the CUDA kernels behind include/hai719_rt.h. The bridge is flatten(), which turns the scene
// into the POD RtSceneDesc once; render() / ray_trace_from_camera() (Renderer.h) then call the C
// ABI. Containers are public (they are private-by-default in the reference, which forces its
// only client, main.cpp, to go through setup_*()).
#ifndef HAI719_HOST_SCENE_H
#define HAI719_HOST_SCENE_H
#include <cstdint>
#include <string>
#include <vector>
#include "Constants.h"
#include "Functions.h"
#include "KDTree.hpp"
#include "Material.h"
#include "Mesh.h"
#include "Sphere.h"
#include "Square.h"
#include "Vec3.h"
#include "hai719_rt.h"
#include "imageLoader.h"

enum LightType { LightType_Spherical, LightType_Quad };

struct Light {
    Vec3 material;
    bool isInCamSpace = false;
    LightType type = LightType_Spherical;
    Vec3 pos;
    float radius = 0.f;
    float powerCorrection = 1.f;
};

// Owns the arrays an RtSceneDesc points into.
struct FlatScene {
    RtSceneDesc desc;
    std::vector<RtSphere> spheres;
    std::vector<RtSquare> squares;
    std::vector<RtLight> lights;
    std::vector<RtImage> textures, normal_maps;
    std::vector<RtSceneMesh> meshes;
    std::vector<std::vector<float>> positions, vert_colors, face_colors;
    std::vector<std::vector<uint32_t>> triangles;
    std::vector<std::vector<RtKdNode>> nodes;        // only for meshes without a KD-tree
    std::vector<std::vector<RtTriRef>> leaf_refs;    // (single brute-force leaf)
};

class Scene {
public:
    std::vector<Mesh> meshes;
    std::vector<Sphere> spheres;
    std::vector<Square> squares;
    std::vector<Light> lights;
    std::vector<ppmLoader::ImageRGB> textures;
    std::vector<ppmLoader::ImageRGB> normals;
    ppmLoader::ImageRGB skybox;
    bool dark_sky = true;
    std::string asset_root;  // prefix for the builders' relative "img/..." and "mesh/..." paths ("" = cwd)

    Scene() {}

    void addBox(std::vector<Material> const &materials, bool faces[6], Vec3 const &pos, Vec3 const rotation,
                float const size = 1.f, bool facing_out = true);
    void loadSkybox(const std::string &filename);
    int load_texture(const std::string &filename);
    int load_normal_map(const std::string &filename);
    void clear();
    void computeKDTrees();

    // the reference's builders (Scene.h:358-1882), registered order of main.cpp:421-432
    void setup_single_sphere();
    void setup_single_square();
    void setup_cornell_box(float aspect_ratio);
    void setup_mesh();
    void setup_rt_in_a_weekend();
    void setup_random_spheres();
    void setup_debug_refraction();
    void setup_flamingo();
    void setup_raccoon();
    void setup_flamingo_pond();
    void setup_flamingo_lake();   // defined but unregistered in the reference (Scene.h:1264)
    void setup_backrooms_pool();
    // BASELINE.json config 5: random spheres with motion blur + triceratops.off + gorilla.off
    void setup_motion_blur_meshes();
    // by id: 0..10 as main.cpp:421-432, 11 = flamingo_lake, 100 = config 5. Returns false if unknown.
    bool setup_by_id(int id, float aspect_ratio);

    // Scene -> POD. Pointers inside `out.desc` refer to `out`'s vectors and to this scene's
    // images: both must outlive the rt_scene_create() call that consumes the description.
    void flatten(FlatScene &out) const;
    // Canonical word dump of everything the tracer reads (format shared with the oracle's
    // ref_scene_dump; tests compare the two bit for bit).
    void dump(std::vector<uint32_t> &words) const;

private:
    std::string path(const std::string &rel) const;
    Square &new_square();
    Sphere &new_sphere(Vec3 center, float radius);
    Light &new_light(Vec3 pos, float radius = 1.5f);
    Mesh &new_mesh(const std::string &off_file);
};
#endif
