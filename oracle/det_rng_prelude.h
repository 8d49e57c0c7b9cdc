// TEST INFRASTRUCTURE (oracle) — force-included (-include) in front of the reference's
// UNMODIFIED src/Functions.cpp when building the deterministic oracle.
//
// Functions.cpp:4-8 is
//     static std::uniform_real_distribution<float> distribution(0.0, 1.0);
//     static std::mt19937 generator(time(nullptr));
//     return distribution(generator);
// After <random> has been included for real (guards make the later #include a no-op), the
// token `uniform_real_distribution` is redirected to a type whose call operator ignores the
// engine and returns the next number of the counter-based stream in oracle/det_rng.h. Every
// other line of the reference file compiles as written, so random_float(a,b),
// random_unit_vector() (and its right-to-left argument evaluation under g++) are the
// reference's own.
#include <random>
#include <ctime>
#include "Functions.h"
#include "det_rng.h"
namespace std {
template <class T> struct oracle_det_distribution {
    oracle_det_distribution(double, double) {}
    template <class Engine> T operator()(Engine &) { return oracle::det_next(); }
};
}  // namespace std
#define uniform_real_distribution oracle_det_distribution
