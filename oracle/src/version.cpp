// ORACLE — stands in for the cmake-generated version.cpp (cmake/version.cpp.in).
#include "LatticeDNAOrigami/version.hpp"
char const* const GIT_COMMIT = "oracle-build";
