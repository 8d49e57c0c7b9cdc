// Host API — the reference's compile-time knobs (src/Constants.h:5-18), kept as the DEFAULTS of
// the runtime RtRenderParams. MULTI_THREADED / MONORAY have no meaning on the GPU path.
#ifndef HAI719_HOST_CONSTANTS_H
#define HAI719_HOST_CONSTANTS_H
#define DEFAULT_SELECTED_SCENE 2
#define DEFAULT_NSAMPLES 20
#define MAXBOUNCES 6
#define NB_ECH 10
#define KDTREE_MAX_DEPTH 100
#define KDTREE_TRIANGLES_PER_LEAF 40
#define EPSILON 0.00001          // a double, as in the reference: comparisons promote to fp64
#define TRIANGLE_SCALING 1.000001f  // Mesh.h:23 — applied on the device at flatten time
#endif
