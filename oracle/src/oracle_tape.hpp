// ORACLE — test infrastructure only. Value-level RNG tape shared by
// tape_random_gens.cpp and ref_driver.cpp.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
namespace oracle_tape {
struct Draw {
    int32_t kind; // 0 = uniform_real, 1 = uniform_int
    int32_t lo;
    int32_t hi;
    int32_t ival;
    double real;
};
using Tape = std::vector<Draw>;
extern thread_local Tape* g_record;
extern thread_local Tape* g_replay;
extern thread_local size_t g_replay_pos;
extern thread_local bool g_quiet_seed;
// the RandomGens object every recorded draw came from (PTMWUS: the multi-window driver's own generator serves the
// exchange tests, the inner simulation's generator the Monte Carlo moves)
extern thread_local std::vector<const void*>* g_record_src;
} // namespace oracle_tape
