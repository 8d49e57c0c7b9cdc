#include "errors.h"
#include <cstdlib>
#include <iostream>
namespace hai719 {
static thread_local bool g_throw = false;
void set_fatal_throws(bool on) { g_throw = on; }
bool fatal_throws() { return g_throw; }
void fatal(const std::string &what) {
    if (g_throw) throw std::runtime_error(what);
    std::cerr << "hai719: " << what << std::endl;
    std::exit(EXIT_FAILURE);
}
}  // namespace hai719
