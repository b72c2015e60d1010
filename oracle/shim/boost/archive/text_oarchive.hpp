// Stub of boost/archive/text_oarchive.hpp for the oracle build (test infrastructure only).
// Archives are never used on the Monte Carlo path; every operation throws.
#pragma once
#include <cmath>
#include <chrono>
#include <iostream>
#include <stdexcept>
#include <boost/serialization/access.hpp>
namespace boost { namespace archive {
class text_oarchive {
  public:
    explicit text_oarchive(std::ostream&) {}
    template <class T> text_oarchive& operator<<(T const&) { throw std::runtime_error("boost archive stub"); }
    template <class T> text_oarchive& operator>>(T&) { throw std::runtime_error("boost archive stub"); }
    template <class T> text_oarchive& operator&(T&) { throw std::runtime_error("boost archive stub"); }
};
} }
