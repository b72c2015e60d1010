// Stub of boost/math/special_functions/factorials.hpp for the oracle build.
#pragma once
namespace boost { namespace math {
template <class T> inline T factorial(unsigned n) {
    T r = 1;
    for (unsigned i = 2; i <= n; i++) r *= static_cast<T>(i);
    return r;
}
} }
