// Host API — error policy. The reference calls exit(EXIT_FAILURE) when a mesh file is missing
// (Mesh.cpp:11-13). The drop-in keeps that for C++ callers by default; the C wrappers
// (capi_host.cpp) switch to exceptions so that a bad path becomes an error code instead.
#ifndef HAI719_HOST_ERRORS_H
#define HAI719_HOST_ERRORS_H
#include <stdexcept>
#include <string>
namespace hai719 {
void set_fatal_throws(bool on);
bool fatal_throws();
[[noreturn]] void fatal(const std::string &what);
}
#endif
