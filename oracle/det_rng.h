// TEST INFRASTRUCTURE (oracle) — the deterministic counter-based random stream that both
// sides of a parity test use in place of the reference's time-seeded, racy std::mt19937
// (Functions.cpp:4-8, main.cpp:181).
//
// Definition (restated independently in include/hai719_rt.h for the product):
//   fmix32(h): h ^= h>>16; h *= 0x85EBCA6B; h ^= h>>13; h *= 0xC2B2AE35; h ^= h>>16
//   key(seed, pixel, sample) = fmix32( fmix32(seed ^ (pixel+1)*0x9E3779B9) + (sample+1)*0x85EBCA6B )
//   draw #i of a path        = (fmix32(key + i*0x9E3779B9) >> 8) * 2^-24        in [0, 1)
// pixel = x + y*image_width in FULL-image coordinates, so a crop or a tile shard draws the
// same numbers. Draw order per path is the reference's call order of random_float().
// Scene-construction randomness (Scene.h:895-922) uses pixel = 0xFFFFFFFF, sample = 0.
#ifndef ORACLE_DET_RNG_H
#define ORACLE_DET_RNG_H
#include <cstdint>

namespace oracle {
static inline uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
static inline uint32_t path_key(uint32_t seed, uint32_t pixel, uint32_t sample) {
    return fmix32(fmix32(seed ^ ((pixel + 1u) * 0x9E3779B9u)) + (sample + 1u) * 0x85EBCA6Bu);
}
static inline float draw(uint32_t key, uint32_t i) {
    return (float)(fmix32(key + i * 0x9E3779B9u) >> 8) * (1.0f / 16777216.0f);
}
struct DetCtx { uint32_t key; uint32_t ctr; uint64_t total; };
DetCtx &ctx();               // thread-local, defined in ref_driver.cpp
float det_next();            // draws ctx().ctr++ from ctx().key
}  // namespace oracle
#endif
