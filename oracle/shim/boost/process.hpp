// Stub of boost/process.hpp for the oracle build (the VMD pipe is never used).
#pragma once
#include <chrono>
#include <string>
#include <iostream>
namespace boost { namespace process {
struct std_out_t { template <class T> std_out_t operator>(T const&) const { return *this; } };
static const std_out_t std_out{};
inline std::string search_path(std::string const& s) { return s; }
class child {
  public:
    child() {}
    template <class... A> explicit child(A const&...) {}
    void terminate() {}
    void wait() {}
};
} }
