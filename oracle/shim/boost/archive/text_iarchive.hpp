// Stub of boost/archive/text_iarchive.hpp for the oracle build (test infrastructure only).
// Archives are never used on the Monte Carlo path; every operation throws.
#pragma once
#include <cmath>
#include <chrono>
#include <iostream>
#include <stdexcept>
#include <boost/serialization/access.hpp>
namespace boost { namespace archive {
class text_iarchive {
  public:
    explicit text_iarchive(std::istream&) {}
    template <class T> text_iarchive& operator<<(T const&) { throw std::runtime_error("boost archive stub"); }
    template <class T> text_iarchive& operator>>(T&) { throw std::runtime_error("boost archive stub"); }
    template <class T> text_iarchive& operator&(T&) { throw std::runtime_error("boost archive stub"); }
};
} }
