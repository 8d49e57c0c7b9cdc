// Host API — helpers of src/Functions.h used at scene-construction time.
//
// random_float() here is the HOST side of the deterministic stream documented in
// include/hai719_rt.h (key(seed, pixel = 0xFFFFFFFF, sample = 0), counter advancing per call);
// the reference's is a time-seeded mt19937 (Functions.cpp:4-8). seed_scene_random() rewinds it and
// also calls srand(), because setup_random_spheres picks material types with rand()
// (Scene.h:895). The per-path draws of the render itself happen on the device.
#ifndef HAI719_HOST_FUNCTIONS_H
#define HAI719_HOST_FUNCTIONS_H
#include <cstdint>
#include "Vec3.h"

void seed_scene_random(uint32_t seed);
float random_float();
float random_float(float min, float max);
Vec3 random_unit_vector();
float min(float a, float b);
float max(float a, float b);
float clamp(float x, float min, float max);
Vec3 reflect(const Vec3 &direction_in, const Vec3 &n);
void gamma_correct(Vec3 &color);
#endif
