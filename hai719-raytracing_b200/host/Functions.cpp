#include "Functions.h"
#include <cstdlib>

namespace {
uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
struct SceneStream { uint32_t key; uint32_t ctr; };
SceneStream make_stream(uint32_t seed) {
    const uint32_t pixel = 0xFFFFFFFFu, sample = 0u;
    return {fmix32(fmix32(seed ^ ((pixel + 1u) * 0x9E3779B9u)) + (sample + 1u) * 0x85EBCA6Bu), 0u};
}
thread_local SceneStream g_stream = make_stream(0);
}  // namespace

void seed_scene_random(uint32_t seed) {
    g_stream = make_stream(seed);
    srand(seed);
}

float random_float() {
    const uint32_t r = fmix32(g_stream.key + (g_stream.ctr++) * 0x9E3779B9u);
    return (float)(r >> 8) * (1.0f / 16777216.0f);
}

float random_float(float lo, float hi) { return lo + (hi - lo) * random_float(); }

// The reference builds Vec3(random_float(-1,1), random_float(-1,1), random_float(-1,1)) and g++
// evaluates the three arguments right to left (Functions.cpp:15): z is drawn first.
Vec3 random_unit_vector() {
    const float z = random_float(-1, 1);
    const float y = random_float(-1, 1);
    const float x = random_float(-1, 1);
    Vec3 p(x, y, z);
    p.normalize();
    return p;
}

float min(float a, float b) { return a < b ? a : b; }
float max(float a, float b) { return a > b ? a : b; }
float clamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
Vec3 reflect(const Vec3 &d, const Vec3 &n) { return d - 2 * Vec3::dot(d, n) * n; }
void gamma_correct(Vec3 &c) {
    for (unsigned int i = 0; i < 3; ++i) c[i] = std::pow(c[i], 1.0 / 2.2);
}
