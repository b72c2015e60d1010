// Stub of boost/functional/hash.hpp for the oracle build (test infrastructure only).
// Provides hash_value / hash_combine with the classic Boost combiner so that the
// reference's std::hash specialisations (include/LatticeDNAOrigami/hash.hpp:7-105) compile.
#pragma once
#include <cstddef>
#include <functional>
#include <utility>
#include <vector>
#include <string>
namespace boost {
template <class T> inline std::size_t hash_value(T const& v) { return std::hash<T>{}(v); }
template <class T> inline void hash_combine(std::size_t& seed, T const& v) {
    seed ^= hash_value(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
template <class T> struct hash {
    std::size_t operator()(T const& v) const { return hash_value(v); }
};
} // namespace boost
