// Host API — the reference's compile-time knobs (src/Constants.h:5-18), kept as the DEFAULTS of
// the runtime RtRenderParams. MULTI_THREADED / MONORAY have no meaning on the GPU path.
//
// The values live as typed constants in hai719::defaults; the reference's macro names are aliases
// of them so that code written against its Constants.h compiles unchanged.
#ifndef HAI719_HOST_CONSTANTS_H
#define HAI719_HOST_CONSTANTS_H
namespace hai719 {
namespace defaults {
constexpr int selected_scene = 2;              // scene shown at start-up (main.cpp:66)
constexpr unsigned int nsamples = 20;          // samples per pixel of a "press r" render
constexpr int max_bounces = 6;                 // RtRenderParams.max_bounces
constexpr int shadow_samples = 10;             // RtRenderParams.nb_ech: shadow rays per light and hit
constexpr int kdtree_max_depth = 100;          // host KD build (KDTree.cpp:100-151)
constexpr int kdtree_triangles_per_leaf = 40;
constexpr double epsilon = 0.00001;            // a double, as in the reference: comparisons promote to fp64
constexpr float triangle_scaling = 1.000001f;  // Mesh.h:23 — applied on the device at flatten time
}  // namespace defaults
}  // namespace hai719
#define DEFAULT_SELECTED_SCENE (hai719::defaults::selected_scene)
#define DEFAULT_NSAMPLES (hai719::defaults::nsamples)
#define MAXBOUNCES (hai719::defaults::max_bounces)
#define NB_ECH (hai719::defaults::shadow_samples)
#define KDTREE_MAX_DEPTH (hai719::defaults::kdtree_max_depth)
#define KDTREE_TRIANGLES_PER_LEAF (hai719::defaults::kdtree_triangles_per_leaf)
#define EPSILON (hai719::defaults::epsilon)
#define TRIANGLE_SCALING (hai719::defaults::triangle_scaling)
#endif
