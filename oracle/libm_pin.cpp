// TEST INFRASTRUCTURE (oracle) — libm pin for the two FLOAT functions the path calls.
//
// Scene::skyboxTexture (Scene.h:155-156) calls atan2f/asinf (float overloads). glibc 2.39's are
// not correctly rounded: against the correctly rounded value they differ by one ulp on 16 % /
// 7 % of arguments (measured here), so the sky texel a direction maps to would depend on the libm
// version the oracle happens to link. The oracle therefore defines both as the correctly
// rounded result (fp64 function rounded once to float) — what glibc >= 2.41 (CORE-MATH) returns,
// and what the product computes on the device. oracle/Makefile links with -Bsymbolic-functions
// so the reference objects inside libref_*.so bind to these definitions; nothing outside the
// library sees them (the library is dlopen'ed RTLD_LOCAL).
// Double-precision calls (acos/atan2 in Sphere.h:129-130, pow, fmod, sqrt) go to libm untouched.
#include <cmath>
extern "C" {
float asinf(float x) noexcept { return (float)std::asin((double)x); }
float atan2f(float y, float x) noexcept { return (float)std::atan2((double)y, (double)x); }
// probes so a test can confirm which definition the library really calls
float ref_probe_asinf(float x) { return asinf(x); }
float ref_probe_atan2f(float y, float x) { return atan2f(y, x); }
}
