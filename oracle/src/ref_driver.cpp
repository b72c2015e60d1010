// ORACLE — test infrastructure only; never linked into the product library.
//
// C-ABI driver around the UNMODIFIED reference sources (compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/liboracle_ref.so). It drives the
// reference exactly as apps/main.cpp:19-99 does (InputParameters(argc, argv) ->
// origami::setup_origami -> ConstantTGCMCSimulation) and exposes
//   * configuration, counters, energy and its enthalpy/entropy/stacking split,
//   * GCMCSimulation::simulate() (simulation.cpp:568-653) in caller-sized chunks,
//   * the value-level RNG tape recorded by tape_random_gens.cpp (record / replay),
//   * per-movetype attempt / accept counters (movetypes.hpp:49-52),
//   * leaf functions pinned by the reference's own tests (nearest_neighbour.cpp, ideal_random_walk.cpp),
//   * the temperature-replica-exchange acceptance rule (ptmc_simulation.cpp:275-313).
// Built with -fno-access-control so protected members can be read; nothing is modified.

#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "LatticeDNAOrigami/bias_functions.hpp"
#include "LatticeDNAOrigami/constant_temp_simulation.hpp"
#include "LatticeDNAOrigami/domain.hpp"
#include "LatticeDNAOrigami/files.hpp"
#include "LatticeDNAOrigami/ideal_random_walk.hpp"
#include "LatticeDNAOrigami/movetypes.hpp"
#include "LatticeDNAOrigami/nearest_neighbour.hpp"
#include "LatticeDNAOrigami/order_params.hpp"
#include "LatticeDNAOrigami/origami_system.hpp"
#include "LatticeDNAOrigami/parser.hpp"
#include "LatticeDNAOrigami/simulation.hpp"
#include "oracle_tape.hpp"

using domainContainer::Domain;
using utility::Occupancy;
using utility::VectorThree;

namespace {

struct Handle {
    parser::InputParameters* params {nullptr};
    origami::OrigamiSystem* origami {nullptr};
    constantTemp::ConstantTGCMCSimulation* sim {nullptr};
    oracle_tape::Tape tape {};
    oracle_tape::Tape replay {};
    long long step {0};
    std::string err {};
};

void set_err(char* err, int errlen, std::string const& msg) {
    if (err != nullptr and errlen > 0) {
        std::strncpy(err, msg.c_str(), errlen - 1);
        err[errlen - 1] = 0;
    }
}

// Silence the reference's std::cout chatter while inside the driver
struct CoutSilencer {
    std::streambuf* old;
    std::ostringstream sink;
    CoutSilencer(): old {std::cout.rdbuf(sink.rdbuf())} {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};

int state_code(Occupancy s) {
    switch (s) {
    case Occupancy::unassigned: return 0;
    case Occupancy::unbound: return 1;
    case Occupancy::bound: return 2;
    case Occupancy::misbound: return 3;
    }
    return -1;
}

} // namespace

extern "C" {

void* oref_create(const char* inp_path, int with_sim, char* err, int errlen) {
    auto h = new Handle {};
    try {
        CoutSilencer quiet {};
        std::string a0 {"oracle"}, a1 {"-i"}, a2 {inp_path};
        char* argv[] {&a0[0], &a1[0], &a2[0]};
        h->params = new parser::InputParameters {3, argv};
        h->origami = origami::setup_origami(*h->params);
        if (with_sim) {
            oracle_tape::g_record = nullptr;
            oracle_tape::g_replay = nullptr;
            h->sim = new constantTemp::ConstantTGCMCSimulation {
                    *h->origami,
                    h->origami->get_system_order_params(),
                    h->origami->get_system_biases(),
                    *h->params};
            h->sim->m_logging_stream = new std::ostringstream {};
        }
    } catch (std::exception const& e) {
        set_err(err, errlen, e.what());
        delete h;
        return nullptr;
    }
    return h;
}

void oref_destroy(void* vh) {
    auto h = static_cast<Handle*>(vh);
    CoutSilencer quiet {};
    delete h->sim;
    delete h->origami;
    delete h->params;
    delete h;
}

const char* oref_last_error(void* vh) { return static_cast<Handle*>(vh)->err.c_str(); }

// ---- stepping -----------------------------------------------------------------------

// Runs `steps` MC steps (simulation.cpp:568). Draws are appended to the handle's tape;
// when a replay tape is installed they are served from it. Returns 0 or -1 (see last_error).
int oref_simulate(void* vh, long long steps) {
    auto h = static_cast<Handle*>(vh);
    try {
        CoutSilencer quiet {};
        oracle_tape::g_record = &h->tape;
        long long next {h->sim->simulate(steps, h->step, false)};
        h->step = next - 1;
        oracle_tape::g_record = nullptr;
    } catch (std::exception const& e) {
        oracle_tape::g_record = nullptr;
        h->err = e.what();
        return -1;
    }
    return 0;
}

long long oref_step(void* vh) { return static_cast<Handle*>(vh)->step; }

void oref_seed(void* vh, int seed) {
    static_cast<Handle*>(vh)->sim->m_random_gens.set_seed(seed);
}

long long oref_tape_len(void* vh) { return static_cast<Handle*>(vh)->tape.size(); }

void oref_tape_copy(void* vh, oracle_tape::Draw* out) {
    auto h = static_cast<Handle*>(vh);
    std::memcpy(out, h->tape.data(), h->tape.size() * sizeof(oracle_tape::Draw));
}

void oref_tape_clear(void* vh) { static_cast<Handle*>(vh)->tape.clear(); }

void oref_tape_set_replay(void* vh, oracle_tape::Draw const* draws, long long n) {
    auto h = static_cast<Handle*>(vh);
    h->replay.assign(draws, draws + n);
    oracle_tape::g_replay = n > 0 ? &h->replay : nullptr;
    oracle_tape::g_replay_pos = 0;
}

int oref_num_movetypes(void* vh) { return static_cast<Handle*>(vh)->sim->m_movetypes.size(); }

void oref_move_stats(void* vh, long long* attempts, long long* accepts) {
    auto h = static_cast<Handle*>(vh);
    for (size_t i {0}; i != h->sim->m_movetypes.size(); i++) {
        attempts[i] = h->sim->m_movetypes[i]->get_attempts();
        accepts[i] = h->sim->m_movetypes[i]->get_accepts();
    }
}

// ---- state ----------------------------------------------------------------------------

int oref_num_chains(void* vh) { return static_cast<Handle*>(vh)->origami->m_domains.size(); }

int oref_num_domains(void* vh) { return static_cast<Handle*>(vh)->origami->num_domains(); }

// Working-order chains (origami_system.cpp:173-191). pos/ore are 3 ints per domain,
// state per domain (0 unassigned, 1 unbound, 2 bound, 3 misbound), bound = (chain index, domain) or -1,-1.
void oref_get_state(
        void* vh,
        int* chain_index,
        int* chain_ident,
        int* chain_len,
        int* pos,
        int* ore,
        int* state,
        int* bound) {
    auto h = static_cast<Handle*>(vh);
    auto& o {*h->origami};
    size_t k {0};
    for (size_t i {0}; i != o.m_domains.size(); i++) {
        chain_index[i] = o.m_chain_indices[i];
        chain_ident[i] = o.m_chain_identities[i];
        chain_len[i] = o.m_domains[i].size();
        for (auto d: o.m_domains[i]) {
            for (int a {0}; a != 3; a++) {
                pos[3 * k + a] = d->m_pos.at(a);
                ore[3 * k + a] = d->m_ore.at(a);
            }
            state[k] = state_code(d->m_state);
            if (d->m_bound_domain != nullptr) {
                bound[2 * k] = d->m_bound_domain->m_c;
                bound[2 * k + 1] = d->m_bound_domain->m_d;
            }
            else {
                bound[2 * k] = -1;
                bound[2 * k + 1] = -1;
            }
            k++;
        }
    }
}

// origami_system.cpp:327-341 (set_config)
int oref_set_state(
        void* vh,
        int nchains,
        int const* chain_index,
        int const* chain_ident,
        int const* chain_len,
        int const* pos,
        int const* ore) {
    auto h = static_cast<Handle*>(vh);
    try {
        origami::Chains chains {};
        size_t k {0};
        for (int i {0}; i != nchains; i++) {
            origami::Chain c {};
            c.index = chain_index[i];
            c.identity = chain_ident[i];
            for (int d {0}; d != chain_len[i]; d++) {
                c.positions.push_back({pos[3 * k], pos[3 * k + 1], pos[3 * k + 2]});
                c.orientations.push_back({ore[3 * k], ore[3 * k + 1], ore[3 * k + 2]});
                k++;
            }
            chains.push_back(c);
        }
        h->origami->set_config(chains);
    } catch (std::exception const& e) {
        h->err = e.what();
        return -1;
    }
    return 0;
}

double oref_energy(void* vh) { return static_cast<Handle*>(vh)->origami->energy(); }

// counters: staples, domains, bound pairs, fully bound pairs, self-bound pairs, misbound pairs,
// stacked pairs, unassigned domains, current unique chain index
void oref_counters(void* vh, int* out) {
    auto& o {*static_cast<Handle*>(vh)->origami};
    out[0] = o.num_staples();
    out[1] = o.num_domains();
    out[2] = o.num_bound_domain_pairs();
    out[3] = o.num_fully_bound_domain_pairs();
    out[4] = o.num_self_bound_domain_pairs();
    out[5] = o.num_misbound_domain_pairs();
    out[6] = o.num_stacked_domain_pairs();
    out[7] = o.num_unassigned_domains();
    out[8] = o.m_current_c_i;
}

// origami_system.cpp:204-246: out = enthalpy, entropy, stacking
void oref_energy_split(void* vh, double* out) {
    auto& o {*static_cast<Handle*>(vh)->origami};
    o.update_enthalpy_and_entropy();
    out[0] = o.hybridization_enthalpy();
    out[1] = o.hybridization_entropy();
    out[2] = o.stacking_energy();
}

int oref_check_all_constraints(void* vh) {
    auto h = static_cast<Handle*>(vh);
    try {
        h->origami->check_all_constraints();
    } catch (std::exception const& e) {
        h->err = e.what();
        return -1;
    }
    return 0;
}

void oref_center(void* vh, int centering_domain) {
    static_cast<Handle*>(vh)->origami->center(centering_domain);
}

// origami_system.cpp:618-628
int oref_update_temp(void* vh, double temp, double stacking_mult) {
    auto h = static_cast<Handle*>(vh);
    try {
        h->origami->update_temp(temp, stacking_mult);
    } catch (std::exception const& e) {
        h->err = e.what();
        return -1;
    }
    return 0;
}

void oref_update_staple_us(void* vh, double temp, double mult) {
    static_cast<Handle*>(vh)->origami->update_staple_us(temp, mult);
}

int oref_num_staple_types(void* vh) {
    return static_cast<Handle*>(vh)->origami->m_identities.size() - 1;
}

void oref_staple_us(void* vh, double* out) {
    auto& o {*static_cast<Handle*>(vh)->origami};
    for (size_t i {0}; i != o.m_staple_us.size(); i++) out[i] = o.m_staple_us[i];
}

// ---- potential tables (origami_potential.cpp:1057-1221, 1319-1350) ---------------------

// Returns 0 and fills out[4] = {hyb energy, hyb enthalpy, hyb entropy, stacking energy}
// for identity pair (a, b); -1 if the pair is not tabulated.
int oref_pair_energies(void* vh, int ident_a, int ident_b, double* out) {
    auto& p {static_cast<Handle*>(vh)->origami->m_pot};
    std::pair<int, int> key {ident_a, ident_b};
    if (p.m_hybridization_energies.count(key) == 0) return -1;
    out[0] = p.m_hybridization_energies.at(key);
    out[1] = p.m_hybridization_enthalpies.at(key);
    out[2] = p.m_hybridization_entropies.at(key);
    out[3] = p.m_stacking_energies.at(key);
    return 0;
}

void oref_init_energies(void* vh, double* out) {
    auto& p {static_cast<Handle*>(vh)->origami->m_pot};
    out[0] = p.init_energy();
    out[1] = p.init_enthalpy();
    out[2] = p.init_entropy();
}

// ---- order parameters and biases ------------------------------------------------------

int oref_order_param(void* vh, const char* tag, int* value) {
    auto h = static_cast<Handle*>(vh);
    try {
        auto& op {h->origami->get_system_order_params().get_order_param(tag)};
        op.calc_param();
        *value = op.get_param();
    } catch (std::exception const& e) {
        h->err = e.what();
        return -1;
    }
    return 0;
}

// The stored value and `defined` flag of an order parameter, without re-evaluating it (calc_param above refreshes a
// per-domain parameter, which changes what the reference does next)
int oref_order_param_stored(void* vh, const char* tag, int* value, int* defined) {
    auto h = static_cast<Handle*>(vh);
    try {
        auto& op {h->origami->get_system_order_params().get_order_param(tag)};
        *value = op.get_param();
        *defined = op.defined() ? 1 : 0;
    } catch (std::exception const& e) {
        h->err = e.what();
        return -1;
    }
    return 0;
}

double oref_total_bias(void* vh) {
    auto h = static_cast<Handle*>(vh);
    h->origami->get_system_order_params().update_move_params();
    h->origami->get_system_biases().calc_move();
    return h->origami->get_system_biases().get_total_bias();
}

// The stored total as the .ene writer reads it (files.cpp:707-720), without re-evaluating anything
double oref_total_bias_stored(void* vh) { return static_cast<Handle*>(vh)->origami->get_system_biases().get_total_bias(); }

// ---- single-domain operations (origami_system.cpp:343-385, 478-541) -------------------

// Finds the domain (unique chain index c, domain index d)
static Domain* find_domain(Handle* h, int c, int d) { return h->origami->get_domain(c, d); }

// out[0] = delta energy, returns 1 if constraints violated (and clears the flag), 0 otherwise
int oref_check_domain(void* vh, int c, int d, int const* pos, int const* ore, double* out) {
    auto h = static_cast<Handle*>(vh);
    Domain* dom {find_domain(h, c, d)};
    out[0] = h->origami->check_domain_constraints(
            *dom, {pos[0], pos[1], pos[2]}, {ore[0], ore[1], ore[2]});
    int violated {h->origami->m_constraints_violated ? 1 : 0};
    h->origami->m_constraints_violated = false;
    return violated;
}

int oref_set_domain(void* vh, int c, int d, int const* pos, int const* ore, double* out) {
    auto h = static_cast<Handle*>(vh);
    Domain* dom {find_domain(h, c, d)};
    out[0] = h->origami->set_domain_config(
            *dom, {pos[0], pos[1], pos[2]}, {ore[0], ore[1], ore[2]});
    int violated {h->origami->m_constraints_violated ? 1 : 0};
    h->origami->m_constraints_violated = false;
    return violated;
}

double oref_unassign_domain(void* vh, int c, int d) {
    auto h = static_cast<Handle*>(vh);
    return h->origami->unassign_domain(*find_domain(h, c, d));
}

// ---- leaf functions pinned by the reference's own tests --------------------------------

// nearest_neighbour.cpp:34-56 via calc_unitless_hybridization_thermo (:162-176)
void oref_nn_unitless_thermo(const char* seq, double temp, double cation_M, double* out) {
    auto t {nearestNeighbour::calc_unitless_hybridization_thermo(seq, temp, cation_M)};
    out[0] = t.enthalpy;
    out[1] = t.entropy;
}

double oref_nn_unitless_energy(const char* seq, double temp, double cation_M) {
    return nearestNeighbour::calc_unitless_hybridization_energy(seq, temp, cation_M);
}

// nearest_neighbour.cpp:117-160; writes complements separated by '\n'
int oref_nn_longest_contig_complement(const char* a, const char* b, char* out, int outlen) {
    auto v {nearestNeighbour::find_longest_contig_complement(a, b)};
    std::string s {};
    for (auto const& x: v) {
        s += x;
        s += "\n";
    }
    std::strncpy(out, s.c_str(), outlen - 1);
    out[outlen - 1] = 0;
    return v.size();
}

// ideal_random_walk.cpp:14-73
double oref_num_walks(int const* start, int const* end, int steps) {
    idealRandomWalk::IdealRandomWalks w {};
    return static_cast<double>(w.num_walks(
            {start[0], start[1], start[2]}, {end[0], end[1], end[2]}, steps));
}

} // extern "C"

// ---- debugging aid: regrowth stack and initial active endpoints of the last CTRG move ----------------
#include "LatticeDNAOrigami/rg_movetypes.hpp"
extern "C" int oref_debug_rg(void* vh, int movetype, int* out, int cap) {
    auto h = static_cast<Handle*>(vh);
    auto* mt = dynamic_cast<movetypes::CTRGRegrowthMCMovetype*>(h->sim->m_movetypes[movetype].get());
    if (mt == nullptr) return -1;
    int n = 0;
    auto put = [&](int v) { if (n < cap) out[n++] = v; };
    put(static_cast<int>(mt->m_regrow_ds.size()));
    for (auto d: mt->m_regrow_ds) {
        put(d->m_c);
        put(d->m_d);
        put(mt->m_constraintpoints.m_segs[d]);
        put(mt->m_constraintpoints.get_dir(d));
    }
    put(-999);
    for (auto const& kv: mt->m_constraintpoints.m_initial_active_endpoints) {
        for (auto const& ep: kv.second) {
            put(kv.first.first);
            put(kv.first.second);
            put(ep.first);
            put(ep.second.at(0));
            put(ep.second.at(1));
            put(ep.second.at(2));
        }
    }
    return n;
}

// ---- replica exchange: the reference's own acceptance probability --------------------------------------
// Builds a UTPTGCMCSimulation from the handle's parameters (which must describe a ut_parallel_tempering
// run; rank 0 of the thread-backed MPI shim, nothing is communicated by the constructor) and evaluates
// PTGCMCSimulation::calc_acceptance_p (ptmc_simulation.cpp:275-313) for one pair of replicas.
// dep1 / dep2 = {enthalpy, bias, stacking}; staple_u / staple_n have n_staple entries each.
#include "LatticeDNAOrigami/ptmc_simulation.hpp"
extern "C" int oref_pt_acceptance_p(
        void* vh, double temp1, double temp2, double umult1, double umult2, double bmult1, double bmult2,
        double smult1, double smult2, const double* dep1, const double* dep2, int n_staple,
        const double* staple_u1, const double* staple_u2, const double* staple_n1, const double* staple_n2,
        double* p_out) {
    auto h = static_cast<Handle*>(vh);
    try {
        CoutSilencer quiet {};
        static ptmc::UTPTGCMCSimulation* pt {nullptr};
        static Handle* owner {nullptr};
        if (owner != h) {
            pt = new ptmc::UTPTGCMCSimulation {
                    *h->origami, h->origami->get_system_order_params(), h->origami->get_system_biases(), *h->params};
            owner = h;
        }
        std::vector<std::pair<double, double>> control {{temp1, temp2}, {umult1, umult2}, {bmult1, bmult2}, {smult1, smult2}};
        std::vector<std::pair<double, double>> dependent {{dep1[0], dep2[0]}, {dep1[1], dep2[1]}, {dep1[2], dep2[2]}};
        std::vector<double> u1(staple_u1, staple_u1 + n_staple), u2(staple_u2, staple_u2 + n_staple);
        std::vector<double> n1(staple_n1, staple_n1 + n_staple), n2(staple_n2, staple_n2 + n_staple);
        std::vector<std::pair<std::vector<double>, std::vector<double>>> per_staple {{u1, u2}, {n1, n2}};
        *p_out = pt->calc_acceptance_p(control, dependent, per_staple);
    } catch (std::exception const& e) {
        h->err = e.what();
        return -1;
    }
    return 0;
}

// ---- replica exchange: the reference's own PT drivers, one thread per rank ----------------------------
//
// Runs the UNMODIFIED PTGCMCSimulation subclasses (ptmc_simulation.cpp:106-150) with `n_ranks` replicas inside
// this process: each rank is a std::thread over the thread-backed boost::mpi shim (shim/boost/mpi) and does what
// apps/main.cpp does (InputParameters -> setup_origami -> simulation constructor -> run()). Every rank records
// its own value-level tape; rank 0's tape also carries the master's exchange draws (App. A20), whose positions
// are marked by a thin subclass that brackets attempt_exchange() - nothing of the reference is modified.

#include <memory>
#include <thread>
#include <type_traits>

#include "LatticeDNAOrigami/ptmc_simulation.hpp"
#include "LatticeDNAOrigami/us_simulation.hpp"
#include <boost/mpi/communicator.hpp>

namespace {

struct NullBuf: std::streambuf {
    int overflow(int c) override { return c; }
    std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};

struct PtRank {
    oracle_tape::Tape tape {};
    std::vector<long long> marks {}; // rank 0: [begin, end) tape positions of every attempt_exchange call
    std::vector<const void*> src {}; // generator of every taped draw (umbrella-sampling drivers)
    std::vector<unsigned char> exchange_flag {}; // 1: the draw came from the multi-window driver's own generator
    std::vector<int> chain_index, chain_ident, chain_len, pos, ore;
    std::vector<long long> attempts, accepts;
    double energy {0};
    std::string err {};
};

struct PtHandle {
    std::vector<PtRank> ranks;
    std::string err {};
};

template <class Base>
struct MarkedPT: Base {
    using Base::Base;
    PtRank* rec {nullptr};
    void attempt_exchange(int swap_i) override {
        long long b {static_cast<long long>(rec->tape.size())};
        Base::attempt_exchange(swap_i);
        rec->marks.push_back(b);
        rec->marks.push_back(static_cast<long long>(rec->tape.size()));
    }
};

template <class Sim>
void pt_rank_body(parser::InputParameters& params, origami::OrigamiSystem& origami, PtRank& rec, int seed, bool record) {
    MarkedPT<Sim> sim {origami, origami.get_system_order_params(), origami.get_system_biases(), params};
    sim.rec = &rec;
    sim.m_random_gens.set_seed(seed);
    if (record) oracle_tape::g_record = &rec.tape;
    sim.run();
    oracle_tape::g_record = nullptr;
    for (auto& mt: sim.m_movetypes) {
        rec.attempts.push_back(mt->get_attempts());
        rec.accepts.push_back(mt->get_accepts());
    }
}

// Umbrella-sampling drivers (us_simulation.cpp:75-91, 475-484, 647-660). The multi-window classes own an inner
// SimpleUSGCMCSimulation with its own RandomGens (the one that drives the MC moves); the outer object's generator only
// serves PTMWUS' exchange tests. Both are seeded per rank.
template <class Sim>
void us_rank_body(parser::InputParameters& params, origami::OrigamiSystem& origami, PtRank& rec, int seed, bool record) {
    Sim sim {origami, origami.get_system_order_params(), origami.get_system_biases(), params};
    sim.m_random_gens.set_seed(seed);
    simulation::GCMCSimulation* mc {&sim};
    if constexpr (std::is_base_of<us::MWUSGCMCSimulation, Sim>::value) {
        sim.m_us_sim->m_random_gens.set_seed(seed + 1000003);
        mc = sim.m_us_sim;
    }
    if (record) {
        oracle_tape::g_record = &rec.tape;
        oracle_tape::g_record_src = &rec.src;
    }
    sim.run();
    oracle_tape::g_record = nullptr;
    oracle_tape::g_record_src = nullptr;
    if (mc != &sim) {
        for (auto p: rec.src) rec.exchange_flag.push_back(p == static_cast<const void*>(&sim.m_random_gens) ? 1 : 0);
    }
    for (auto& mt: mc->m_movetypes) {
        rec.attempts.push_back(mt->get_attempts());
        rec.accepts.push_back(mt->get_accepts());
    }
}

void pt_rank_main(const char* inp_path, int rank, int seed, bool record, PtRank* rec) {
    boost::mpi::shim_rank() = rank;
    try {
        std::string a0 {"oracle"}, a1 {"-i"}, a2 {inp_path};
        char* argv[] {&a0[0], &a1[0], &a2[0]};
        parser::InputParameters params {3, argv};
        std::unique_ptr<origami::OrigamiSystem> origami {origami::setup_origami(params)};
        std::string const& st = params.m_simulation_type;
        if (st == "t_parallel_tempering") pt_rank_body<ptmc::TPTGCMCSimulation>(params, *origami, *rec, seed, record);
        else if (st == "ut_parallel_tempering") pt_rank_body<ptmc::UTPTGCMCSimulation>(params, *origami, *rec, seed, record);
        else if (st == "hut_parallel_tempering") pt_rank_body<ptmc::HUTPTGCMCSimulation>(params, *origami, *rec, seed, record);
        else if (st == "st_parallel_tempering") pt_rank_body<ptmc::STPTGCMCSimulation>(params, *origami, *rec, seed, record);
        else if (st == "2d_parallel_tempering") pt_rank_body<ptmc::TwoDPTGCMCSimulation>(params, *origami, *rec, seed, record);
        else if (st == "umbrella_sampling") us_rank_body<us::SimpleUSGCMCSimulation>(params, *origami, *rec, seed, record);
        else if (st == "mw_umbrella_sampling") us_rank_body<us::MWUSGCMCSimulation>(params, *origami, *rec, seed, record);
        else if (st == "ptmw_umbrella_sampling") us_rank_body<us::PTMWUSGCMCSimulation>(params, *origami, *rec, seed, record);
        else throw std::runtime_error {"oref_pt_run: not a replica-exchange or umbrella-sampling simulation type"};
        auto& o = *origami;
        for (size_t i {0}; i != o.m_domains.size(); i++) {
            rec->chain_index.push_back(o.m_chain_indices[i]);
            rec->chain_ident.push_back(o.m_chain_identities[i]);
            rec->chain_len.push_back(o.m_domains[i].size());
            for (auto d: o.m_domains[i]) {
                for (int a {0}; a != 3; a++) {
                    rec->pos.push_back(d->m_pos.at(a));
                    rec->ore.push_back(d->m_ore.at(a));
                }
            }
        }
        rec->energy = o.energy();
    } catch (std::exception const& e) {
        rec->err = e.what();
        // a rank that dies would leave the others blocked in recv: there is no recovery, report and abort the run
        std::cerr << "oracle PT rank " << rank << ": " << e.what() << std::endl;
        std::abort();
    }
}

} // namespace

extern "C" {

// seeds[n_ranks]: every rank's mt19937_64 seed (RandomGens::set_seed after construction, so that the ranks do not
// share one stream as they would with a common random_seed, simulation.cpp:199-203)
void* oref_pt_run(const char* inp_path, int n_ranks, const int* seeds, int record_tapes, char* err, int errlen) {
    auto h = new PtHandle {};
    h->ranks.resize(n_ranks);
    NullBuf nullbuf {};
    std::streambuf* old {std::cout.rdbuf(&nullbuf)};
    boost::mpi::shim_world::get().size = n_ranks;
    {
        std::unique_lock<std::mutex> lk(boost::mpi::shim_world::get().mu);
        boost::mpi::shim_world::get().box.clear();
    }
    std::vector<std::thread> threads;
    for (int r {0}; r != n_ranks; r++) threads.emplace_back(pt_rank_main, inp_path, r, seeds[r], record_tapes != 0, &h->ranks[r]);
    for (auto& t: threads) t.join();
    boost::mpi::shim_world::get().size = 1;
    std::cout.rdbuf(old);
    for (auto& r: h->ranks) {
        if (!r.err.empty()) {
            set_err(err, errlen, r.err);
            delete h;
            return nullptr;
        }
    }
    return h;
}
void oref_pt_destroy(void* vh) { delete static_cast<PtHandle*>(vh); }
long long oref_pt_tape_len(void* vh, int rank) { return static_cast<PtHandle*>(vh)->ranks[rank].tape.size(); }
void oref_pt_tape_copy(void* vh, int rank, oracle_tape::Draw* out) {
    auto& t = static_cast<PtHandle*>(vh)->ranks[rank].tape;
    std::memcpy(out, t.data(), t.size() * sizeof(oracle_tape::Draw));
}
long long oref_pt_num_flags(void* vh, int rank) { return static_cast<PtHandle*>(vh)->ranks[rank].exchange_flag.size(); }
void oref_pt_flags(void* vh, int rank, unsigned char* out) {
    auto& f = static_cast<PtHandle*>(vh)->ranks[rank].exchange_flag;
    std::memcpy(out, f.data(), f.size());
}
long long oref_pt_num_marks(void* vh) { return static_cast<PtHandle*>(vh)->ranks[0].marks.size(); }
void oref_pt_marks(void* vh, long long* out) {
    auto& m = static_cast<PtHandle*>(vh)->ranks[0].marks;
    std::memcpy(out, m.data(), m.size() * sizeof(long long));
}
int oref_pt_num_chains(void* vh, int rank) { return static_cast<PtHandle*>(vh)->ranks[rank].chain_index.size(); }
int oref_pt_num_domains(void* vh, int rank) { return static_cast<PtHandle*>(vh)->ranks[rank].pos.size() / 3; }
void oref_pt_get_state(void* vh, int rank, int* chain_index, int* chain_ident, int* chain_len, int* pos, int* ore) {
    auto& r = static_cast<PtHandle*>(vh)->ranks[rank];
    std::memcpy(chain_index, r.chain_index.data(), r.chain_index.size() * sizeof(int));
    std::memcpy(chain_ident, r.chain_ident.data(), r.chain_ident.size() * sizeof(int));
    std::memcpy(chain_len, r.chain_len.data(), r.chain_len.size() * sizeof(int));
    std::memcpy(pos, r.pos.data(), r.pos.size() * sizeof(int));
    std::memcpy(ore, r.ore.data(), r.ore.size() * sizeof(int));
}
double oref_pt_energy(void* vh, int rank) { return static_cast<PtHandle*>(vh)->ranks[rank].energy; }
int oref_pt_num_movetypes(void* vh, int rank) { return static_cast<PtHandle*>(vh)->ranks[rank].attempts.size(); }
void oref_pt_move_stats(void* vh, int rank, long long* attempts, long long* accepts) {
    auto& r = static_cast<PtHandle*>(vh)->ranks[rank];
    std::memcpy(attempts, r.attempts.data(), r.attempts.size() * sizeof(long long));
    std::memcpy(accepts, r.accepts.data(), r.accepts.size() * sizeof(long long));
}

} // extern "C"
