// Stub of a boost/serialization header for the oracle build (test infrastructure only).
#pragma once
#include <cmath>
#include <chrono>
namespace boost { namespace serialization { class access {}; } }
