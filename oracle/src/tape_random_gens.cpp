// ORACLE — test infrastructure only; never linked into the product library.
//
// Replacement for the reference's src/random_gens.cpp (random_gens.cpp:12-49): same
// RandomGens class (declared in include/LatticeDNAOrigami/random_gens.hpp:17-30), same
// mt19937_64 + libstdc++ distributions, plus a value-level tape:
//   record mode : every returned draw is appended to a global tape
//   replay mode : draws are served from a supplied tape; a request whose kind or
//                 (lo,hi) differs from the taped one raises std::runtime_error
// The tape is what lets the CUDA engine be checked bit-exactly, because libstdc++'s
// distribution algorithms are implementation-defined.

#include <iostream>
#include <random>
#include <stdexcept>
#include <string>

#include "LatticeDNAOrigami/random_gens.hpp"
#include "oracle_tape.hpp"

namespace oracle_tape {
thread_local Tape* g_record {nullptr};
thread_local Tape* g_replay {nullptr};
thread_local size_t g_replay_pos {0};
thread_local bool g_quiet_seed {true};
thread_local std::vector<const void*>* g_record_src {nullptr};
} // namespace oracle_tape

namespace randomGen {

using oracle_tape::Draw;

RandomGens::RandomGens() {
    std::random_device true_random_engine {};
    auto seed {true_random_engine()};
    if (not oracle_tape::g_quiet_seed) {
        std::cout << "Truly random seed: " << seed << "\n";
    }
    m_random_engine.seed(seed);
}

RandomGens::~RandomGens() {
    for (auto const key: m_uniform_int_dists) {
        delete &m_uniform_int_dists.at(key.first);
    }
}

void RandomGens::set_seed(int seed) { m_random_engine.seed(seed); }

double RandomGens::uniform_real() {
    if (oracle_tape::g_replay != nullptr) {
        auto& t {*oracle_tape::g_replay};
        if (oracle_tape::g_replay_pos >= t.size()) {
            throw std::runtime_error {"tape exhausted (real)"};
        }
        Draw const& d {t[oracle_tape::g_replay_pos++]};
        if (d.kind != 0) {
            throw std::runtime_error {"tape mismatch: real requested, int taped"};
        }
        if (oracle_tape::g_record != nullptr) oracle_tape::g_record->push_back(d);
        if (oracle_tape::g_record_src != nullptr) oracle_tape::g_record_src->push_back(this);
        return d.real;
    }
    double v {m_uniform_real_dist(m_random_engine)};
    if (oracle_tape::g_record != nullptr) {
        oracle_tape::g_record->push_back(Draw {0, 0, 0, 0, v});
        if (oracle_tape::g_record_src != nullptr) oracle_tape::g_record_src->push_back(this);
    }
    return v;
}

int RandomGens::uniform_int(int lower, int upper) {
    if (oracle_tape::g_replay != nullptr) {
        auto& t {*oracle_tape::g_replay};
        if (oracle_tape::g_replay_pos >= t.size()) {
            throw std::runtime_error {"tape exhausted (int)"};
        }
        Draw const& d {t[oracle_tape::g_replay_pos++]};
        if (d.kind != 1 or d.lo != lower or d.hi != upper) {
            throw std::runtime_error {
                    "tape mismatch: int(" + std::to_string(lower) + "," +
                    std::to_string(upper) + ") requested"};
        }
        if (oracle_tape::g_record != nullptr) oracle_tape::g_record->push_back(d);
        if (oracle_tape::g_record_src != nullptr) oracle_tape::g_record_src->push_back(this);
        return d.ival;
    }
    int v;
    pair<int, int> key {lower, upper};
    if (m_uniform_int_dists.find(key) != m_uniform_int_dists.end()) {
        auto dist {m_uniform_int_dists.at(key)};
        v = dist(m_random_engine);
    }
    else {
        auto dist {new std::uniform_int_distribution<int> {lower, upper}};
        v = (*dist)(m_random_engine);
        m_uniform_int_dists.insert({key, *dist});
    }
    if (oracle_tape::g_record != nullptr) {
        oracle_tape::g_record->push_back(Draw {1, lower, upper, v, 0.0});
        if (oracle_tape::g_record_src != nullptr) oracle_tape::g_record_src->push_back(this);
    }
    return v;
}
} // namespace randomGen
