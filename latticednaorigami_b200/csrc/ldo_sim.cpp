// ldo_sim.cpp — simulation drivers on top of one ldo_engine: the host-side equivalents of
// ConstantTGCMCSimulation::run, AnnealingGCMCSimulation::run and PTGCMCSimulation::run, with the
// reference's text output files. See include/ldo_host.h for the reference file:line of each piece.

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <random>
#include <sstream>

#include "../../include/ldo_host.h"
#include "ldo_host.hpp"

#ifndef LDO_HOSTSIM
#include <dlfcn.h>
#endif

using namespace ldohost;

namespace {
thread_local std::string g_host_error;

struct ReplicaFiles {
    std::unique_ptr<std::ofstream> trj, counts, staples, staplestates, times, ene, ops, randstate;
    std::unique_ptr<std::ofstream> log; // m_logging_stream of the PT drivers: <filebase>-<rank>.out (ptmc_simulation.cpp:56)
    std::unique_ptr<std::ofstream> vcf, states, ores; // setup_config_files (simulation.cpp:150-180)
};
} // namespace

struct ldo_sim {
    InputParameters params;
    std::unique_ptr<OrigamiInputFile> sysfile;
    std::vector<EnergyTables> tables;
    std::vector<double> temps;
    std::vector<MovetypeSpec> movetypes;
    std::vector<OrderParamSpec> ops;
    std::vector<BiasSpec> biases;
    ldo_engine* eng {nullptr};
    int R {0};
    int rank {0};
    int n_ranks {1};
    long long step {0};
    bool is_pt {false};
    int pt_variant {LDO_PT_T};
    int v1_dim {0}, v2_dim {0}; // 2d_parallel_tempering: temperatures x stacking multipliers
    int num_reps {1};
    int n_ladders {1};
    std::vector<int> q2r;
    std::vector<long long> attempts, accepts;
    std::vector<ReplicaFiles> files;
    std::vector<int> file_owner; // file set -> replica (umbrella sampling: files follow windows)
    std::string filebase_postfix; // "_iter-n" of the umbrella-sampling drivers
    std::vector<std::string> window_postfix; // "_win-a--b" per window
    std::vector<int> ops_out_idx;
    // umbrella sampling (us_simulation.hpp:60-150)
    bool is_us {false}, is_mwus {false}, is_ptmwus {false};
    bool is_enum {false}; // simulation_type=enumerate (enumerate.cpp)
    long long enum_leaves {0}; // conformations visited by the last enumeration / their number with multiplicities
    double enum_num_configs {0};
    double enum_average_energy {0}, enum_average_bias {0};
    int n_windows {1};
    int grid_bias {-1};
    std::vector<int> window_biases;
    std::vector<int> grid_lo, grid_n;
    WindowsFile windows;
    std::chrono::steady_clock::time_point start;
    // multi-GPU replica exchange: NCCL communicator (one rank per GPU) and the engine's exchange buffers
    void* nccl_comm {nullptr};
    void *dep_send {nullptr}, *dep_recv {nullptr};
    int dep_nq {0};
    bool exchange_resident {false}; // slot -> replica map and counters currently live on the device

    ~ldo_sim();
    void check(int rc) {
        if (rc != 0) throw std::runtime_error(std::string("engine: ") + ldo_last_error(eng));
    }
};

namespace {

int domain_type_code(std::string const& t) {
    if (t == "HalfTurn") return LDO_DOMAIN_HALFTURN;
    if (t == "ThreeQuarterTurn") return LDO_DOMAIN_THREEQUARTERTURN;
    throw NotImplemented {t + ": No such domain type"}; // origami_system.cpp:423-425
}

int op_type_code(std::string const& t) {
    if (t == "NumStaples") return LDO_OP_NUM_STAPLES;
    if (t == "NumStaplesType") return LDO_OP_NUM_STAPLES_TYPE;
    if (t == "StapleTypeFullyBound") return LDO_OP_STAPLE_TYPE_FULLY_BOUND;
    if (t == "NumBoundDomainPairs") return LDO_OP_NUM_BOUND_DOMAIN_PAIRS;
    if (t == "NumMisboundDomainPairs") return LDO_OP_NUM_MISBOUND_DOMAIN_PAIRS;
    if (t == "NumStackedPairs") return LDO_OP_NUM_STACKED_PAIRS;
    if (t == "NumLinearHelices") return LDO_OP_NUM_LINEAR_HELICES;
    if (t == "NumStackedJuncts") return LDO_OP_NUM_STACKED_JUNCTS;
    if (t == "Sum") return LDO_OP_SUM;
    if (t == "Dist") return LDO_OP_DIST;
    if (t == "AdjacentSite") return LDO_OP_ADJACENT_SITE;
    throw SimulationMisuse {t + ": order parameter type does not exist"}; // order_params.cpp:546-549
}

int bias_type_code(std::string const& t) {
    if (t == "LinearStepWell") return LDO_BIAS_LINEAR_STEP_WELL;
    if (t == "SquareWell") return LDO_BIAS_SQUARE_WELL;
    return LDO_BIAS_GRID;
}

// Ladder slot held by this rank at local slot index b: the slots of a ladder are dealt to the ranks in serpentine
// order (ldo_exchange_pt, ldo_b200.h), which balances the temperature-dependent cost of a move across GPUs
int ladder_slot_of(ldo_sim const& s, int b) { return b * s.n_ranks + ((b & 1) ? s.n_ranks - 1 - s.rank : s.rank); }

// File base of replica r for a run whose file base is `filebase` (the output files; the .randstate files read back)
std::string replica_filebase_of(ldo_sim& s, std::string const& filebase, int r) {
    if (s.is_us) {
        // MWUSGCMCSimulation::setup_window_variables (us_simulation.cpp:486-501) + "_iter-n" (:107,131)
        int ladder {r / s.n_windows}, w {r % s.n_windows};
        std::string base {filebase};
        if (s.is_mwus) base += s.window_postfix[w];
        if (s.R / s.n_windows > 1) base += "_rep-" + std::to_string(s.rank * (s.R / s.n_windows) + ladder);
        return base + s.filebase_postfix;
    }
    // PTGCMCSimulation appends "-<rank>" (ptmc_simulation.cpp:48); batches of independent replicas do the same
    if (s.R * s.n_ranks == 1) return filebase;
    return filebase + "-" + std::to_string(s.rank * s.R + r);
}
std::string replica_filebase(ldo_sim& s, int r) { return replica_filebase_of(s, s.params.m_output_filebase, r); }

void open_output_files(ldo_sim& s) {
    InputParameters const& p = s.params;
    if (p.m_output_filebase.empty()) return;
    s.files.resize(s.R);
    int n_scaf = static_cast<int>(s.sysfile->identities[0].size());
    int max_domains = n_scaf + p.m_max_total_staples * p.m_max_staple_size;
    for (int r {0}; r != s.R; r++) {
        std::string base {replica_filebase(s, r)};
        {
            // OrigamiVSFOutputFile (files.cpp:519-527), always written (simulation.cpp:56-62)
            std::ofstream vsf {base + ".vsf"};
            if (!vsf) throw FileError {"Cannot open output file " + base + ".vsf"};
            vsf << "atom 0:" << n_scaf << " radius 0.25 type scaffold\n";
            vsf << "atom " << n_scaf << ":" << max_domains - 1 << " radius 0.25 type staple";
        }
        ReplicaFiles& f = s.files[r];
        if (p.m_configs_output_freq != 0) f.trj.reset(new std::ofstream {base + ".trj"});
        if (p.m_vtf_output_freq != 0) {
            f.vcf.reset(new std::ofstream {base + ".vcf"});
            f.states.reset(new std::ofstream {base + ".states"});
            f.ores.reset(new std::ofstream {base + ".ores"});
        }
        if (p.m_counts_output_freq != 0) {
            f.counts.reset(new std::ofstream {base + ".counts"});
            f.staples.reset(new std::ofstream {base + ".staples"});
            f.staplestates.reset(new std::ofstream {base + ".staplestates"});
        }
        if (p.m_times_output_freq != 0) {
            f.times.reset(new std::ofstream {base + ".times"});
            *f.times << "step time\n";
        }
        if (p.m_energies_output_freq != 0) {
            f.ene.reset(new std::ofstream {base + ".ene"});
            *f.ene << "step tenergy henthalpy hentropy stacking bias\n";
            f.ene->precision(10);
        }
        if (p.m_order_params_output_freq != 0) {
            f.ops.reset(new std::ofstream {base + ".ops"});
            for (auto const& tag: p.m_ops_to_output) *f.ops << tag << ", "; // header quirk (App. A14)
            *f.ops << "\n";
        }
        // the log of a replica-exchange replica goes to <filebase>-<rank>.out (opened whatever logging_freq says,
        // ptmc_simulation.cpp:56); batches of independent replicas log the same way, a single replica logs to stdout
        if (s.is_pt || (!s.is_us && p.m_logging_freq != 0 && s.R * s.n_ranks != 1)) f.log.reset(new std::ofstream {base + ".out"});
        // RandomEngineStateOutputFile (simulation.cpp:137-144, files.cpp:781-793): one line per write, the state of the
        // replica's generator as decimal numbers - here the Philox state of ldo_get_rng_state
        if (p.m_rand_engine_state_output_freq != 0) f.randstate.reset(new std::ofstream {base + ".randstate"});
    }
    s.ops_out_idx.clear();
    for (auto const& tag: p.m_ops_to_output) {
        int found {-1};
        for (size_t i {0}; i != s.ops.size(); i++)
            if (s.ops[i].tag == tag) found = static_cast<int>(i);
        if (found < 0) throw SimulationMisuse {"ops_to_output: unknown order parameter " + tag};
        s.ops_out_idx.push_back(found);
    }
}

// GCMCSimulation constructor (simulation.cpp:204-212) + RandomEngineStateInputFile::read_state (files.cpp:232-246): with no
// seed specified and read_rand_engine_state set, the generator continues from line restart_step (counted from 0) of
// rand_engine_state_file. A batch of replicas reads the names open_output_files gives a batch with the file's base
// ("<base>-<replica>.randstate", window postfixes for the umbrella-sampling drivers).
void restore_rng_states(ldo_sim& s) {
    InputParameters const& p = s.params;
    std::cout << "Loading random engine state\n";
    int const W {ldo_rng_state_words()};
    std::vector<unsigned long long> words(static_cast<size_t>(s.R) * W);
    for (int r {0}; r != s.R; r++) {
        std::string name {p.m_rand_engine_state_file};
        std::string const ext {".randstate"};
        if (name.size() >= ext.size() && name.compare(name.size() - ext.size(), ext.size(), ext) == 0) {
            name = replica_filebase_of(s, name.substr(0, name.size() - ext.size()), r) + ext;
        }
        else if (s.R * s.n_ranks != 1) {
            throw FileError {"rand_engine_state_file of a batch of replicas must end in .randstate (read as <filebase>-<replica>.randstate)"};
        }
        std::ifstream in {name};
        if (!in) throw FileError {"Random engine state input file " + name + " does not exist"};
        std::string line;
        for (int i {0}; i <= p.m_restart_step; i++) {
            if (!std::getline(in, line)) {
                throw FileError {"Step " + std::to_string(p.m_restart_step) + " not found in trajectory input file" + name};
            }
        }
        std::istringstream is {line};
        for (int k {0}; k != W; k++) {
            if (!(is >> words[static_cast<size_t>(r) * W + k])) throw FileError {"Random engine state in " + name + " is not a Philox state of this engine"};
        }
    }
    s.check(ldo_set_rng_state(s.eng, 0, s.R, words.data()));
}

bool due(int freq, long long step) { return freq != 0 && step % freq == 0; }

// Output at `step` for every replica (simulation.cpp:641-646 and the writers of files.cpp:529-778)
void write_outputs(ldo_sim& s, long long step) {
    InputParameters const& p = s.params;
    if (s.files.empty()) return;
    bool w_trj {due(p.m_configs_output_freq, step)}, w_counts {due(p.m_counts_output_freq, step)};
    bool w_times {due(p.m_times_output_freq, step)}, w_ene {due(p.m_energies_output_freq, step)};
    bool w_ops {due(p.m_order_params_output_freq, step)}, w_vtf {due(p.m_vtf_output_freq, step)};
    bool w_rand {due(p.m_rand_engine_state_output_freq, step)};
    if (w_rand) {
        int const W {ldo_rng_state_words()};
        std::vector<unsigned long long> words(static_cast<size_t>(s.R) * W);
        s.check(ldo_get_rng_state(s.eng, 0, s.R, words.data()));
        for (int r {0}; r != s.R; r++) {
            std::ofstream* f {s.files[r].randstate.get()};
            if (!f) continue;
            for (int k {0}; k != W; k++) *f << (k ? " " : "") << words[static_cast<size_t>(r) * W + k];
            *f << "\n";
            f->flush();
        }
    }
    if (!(w_trj || w_counts || w_times || w_ene || w_ops || w_vtf)) return;
    int nst {static_cast<int>(s.sysfile->identities.size()) - 1};
    std::vector<double> ene;
    std::vector<int> counters, staple_counts, opv;
    if (w_ene) {
        ene.resize(5 * static_cast<size_t>(s.R));
        s.check(ldo_get_energies(s.eng, ene.data()));
    }
    if (w_counts) {
        counters.resize(9 * static_cast<size_t>(s.R));
        staple_counts.resize(static_cast<size_t>(std::max(nst, 1)) * s.R);
        s.check(ldo_get_counters(s.eng, counters.data()));
        s.check(ldo_get_staple_counts(s.eng, staple_counts.data()));
    }
    if (w_ops && !s.ops.empty()) {
        opv.resize(s.ops.size() * s.R);
        s.check(ldo_get_order_params(s.eng, opv.data()));
    }
    double dt {std::chrono::duration<double>(std::chrono::steady_clock::now() - s.start).count()};
    int max_c, max_d;
    ldo_state_capacity(s.eng, &max_c, &max_d);
    std::vector<int> ci(max_c), cid(max_c), cl(max_c), pos(3 * max_d), ore(3 * max_d), st(max_d), bd(2 * max_d);
    for (int r {0}; r != s.R; r++) {
        ReplicaFiles& f = s.files[r];
        int rr {s.file_owner.empty() ? r : s.file_owner[r]}; // replica whose data this file set follows
        bool need_state {(w_trj && f.trj) || (w_counts && f.staplestates) || (w_vtf && f.vcf)};
        int nc {0};
        if (need_state) s.check(ldo_get_state(s.eng, rr, &nc, ci.data(), cid.data(), cl.data(), pos.data(), ore.data(), st.data(), bd.data()));
        if (w_trj && f.trj) {
            // OrigamiTrajOutputFile::write (files.cpp:529-548)
            std::ofstream& o = *f.trj;
            o << step << "\n";
            int k {0};
            for (int c {0}; c != nc; c++) {
                o << ci[c] << " " << cid[c] << "\n";
                for (int d {0}; d != cl[c]; d++)
                    for (int a {0}; a != 3; a++) o << pos[3 * (k + d) + a] << " ";
                o << "\n";
                for (int d {0}; d != cl[c]; d++)
                    for (int a {0}; a != 3; a++) o << ore[3 * (k + d) + a] << " ";
                o << "\n";
                k += cl[c];
            }
            o << "\n";
            o.flush();
        }
        if (w_vtf && f.vcf) {
            // OrigamiVCFOutputFile / OrigamiOrientationOutputFile / OrigamiStateOutputFile::write (files.cpp:550-645):
            // every chain padded to max_staple_size entries, the frame padded to the maximum number of domains
            int n_scaf {static_cast<int>(s.sysfile->identities[0].size())};
            int max_domains {n_scaf + p.m_max_total_staples * p.m_max_staple_size};
            std::ofstream &v = *f.vcf, &so = *f.states, &oo = *f.ores;
            v << "timestep\n";
            int written {0}, k {0};
            for (int c {0}; c != nc; c++) {
                int n {0};
                for (int d {0}; d != cl[c]; d++, n++) {
                    for (int a {0}; a != 3; a++) v << pos[3 * (k + d) + a] << " ";
                    v << "\n";
                    for (int a {0}; a != 3; a++) oo << (st[k + d] != 0 ? ore[3 * (k + d) + a] : 0) << " ";
                    so << (st[k + d] >= 1 && st[k + d] <= 3 ? st[k + d] : 0) << " ";
                }
                for (; n < p.m_max_staple_size; n++) {
                    v << "0 0 0 \n";
                    oo << "0 0 0 ";
                    so << "-1 ";
                }
                written += n;
                k += cl[c];
            }
            for (int d {written}; d < max_domains; d++) {
                v << "0 0 0 \n";
                oo << "0 0 0 ";
                so << "-1 ";
            }
            v << "\n";
            oo << "\n";
            so << "\n";
            v.flush();
            oo.flush();
            so.flush();
        }
        if (w_counts && f.counts) {
            int const* c = &counters[9 * static_cast<size_t>(rr)];
            int unique {0};
            for (int t {0}; t != nst; t++)
                if (staple_counts[static_cast<size_t>(rr) * nst + t] > 0) unique++;
            *f.counts << step << " " << c[0] << " " << unique << " " << c[2] << " " << c[3] << " " << c[5] << " \n";
            f.counts->flush();
            *f.staples << step << " ";
            for (int t {0}; t != nst; t++) *f.staples << staple_counts[static_cast<size_t>(rr) * nst + t] << " ";
            *f.staples << "\n";
            f.staples->flush();
            // OrigamiStaplesFullyBoundOutputFile::write (files.cpp:667-690)
            std::vector<int> full(nst, 0);
            int k {0};
            for (int c {0}; c != nc; c++) {
                if (c > 0) {
                    bool all_bound {true};
                    for (int d {0}; d != cl[c]; d++)
                        if (st[k + d] != 2) all_bound = false;
                    if (all_bound) full[cid[c] - 1] = 1;
                }
                k += cl[c];
            }
            *f.staplestates << step << " ";
            for (int t {0}; t != nst; t++) *f.staplestates << full[t] << " ";
            *f.staplestates << "\n";
            f.staplestates->flush();
        }
        if (w_times && f.times) {
            *f.times << step << " " << dt << "\n";
            f.times->flush();
        }
        if (w_ene && f.ene) {
            double const* e = &ene[5 * static_cast<size_t>(rr)];
            *f.ene << step << " " << e[0] << " " << e[1] << " " << e[2] << " " << e[3] << " " << e[4] << " \n";
            f.ene->flush();
        }
        if (w_ops && f.ops) {
            *f.ops << step;
            for (int idx: s.ops_out_idx) *f.ops << " " << opv[static_cast<size_t>(rr) * s.ops.size() + idx];
            *f.ops << "\n";
            f.ops->flush();
        }
    }
}

// GCMCSimulation::write_log_entry (simulation.cpp:667-705) for every replica, to stdout (one replica) or the replica's
// .out file. att0 / acc0: the move counters before the step that is reported.
void write_log_entries(ldo_sim& s, long long step, std::vector<long long> const& att0, std::vector<long long> const& acc0) {
    size_t n {s.movetypes.size()};
    int nst {static_cast<int>(s.sysfile->identities.size()) - 1};
    std::vector<long long> att(n * s.R), acc(n * s.R);
    s.check(ldo_get_move_stats(s.eng, att.data(), acc.data()));
    std::vector<int> counters(9 * static_cast<size_t>(s.R)), staple_counts(static_cast<size_t>(s.R) * std::max(nst, 1)), temp_idx(s.R);
    std::vector<double> ene(5 * static_cast<size_t>(s.R)), um(s.R), bm(s.R), sm(s.R);
    s.check(ldo_get_counters(s.eng, counters.data()));
    s.check(ldo_get_staple_counts(s.eng, staple_counts.data()));
    s.check(ldo_get_energies(s.eng, ene.data()));
    s.check(ldo_get_control(s.eng, 0, s.R, temp_idx.data(), um.data(), bm.data(), sm.data()));
    for (int r {0}; r != s.R; r++) {
        std::ostream* sink {&std::cout};
        if (!s.files.empty() && s.files[r].log) sink = s.files[r].log.get();
        else if (s.R * s.n_ranks != 1) continue; // a batch without output files has nowhere to log to
        // formatted in a stream of its own: the entry must not inherit whatever precision / flags the process has left
        // on std::cout (the umbrella-sampling summaries set some)
        std::ostringstream entry;
        std::ostream* o {&entry};
        int const* c {&counters[9 * static_cast<size_t>(r)]};
        int unique {0};
        for (int t {0}; t != nst; t++) unique += staple_counts[static_cast<size_t>(r) * nst + t] > 0 ? 1 : 0;
        int moved {-1};
        for (size_t i {0}; i != n; i++)
            if (att[r * n + i] != att0[r * n + i]) moved = static_cast<int>(i);
        bool accepted {moved >= 0 && acc[r * n + moved] != acc0[r * n + moved]};
        *o << "Step: " << step << "\n";
        *o << "Temperature: " << s.temps[temp_idx[r]] << "\n";
        *o << "Bound staples: " << c[0] << "\n";
        *o << "Unique bound staples: " << unique << "\n";
        *o << "Fully bound domain pairs: " << c[3] << "\n";
        *o << "Misbound domain pairs: " << c[5] << "\n";
        *o << "Stacked domain pairs: " << c[6] << "\n";
        // (never updated by the reference, App. A5)
        *o << "Linear helix triplets: " << 0 << "\n";
        *o << "Stacked junction quadruplets: " << 0 << "\n";
        *o << "Staple counts: ";
        for (int t {0}; t != nst; t++) *o << staple_counts[static_cast<size_t>(r) * nst + t] << " ";
        *o << "\n";
        *o << "System energy: " << ene[5 * static_cast<size_t>(r)] << "\n";
        // every bias is of the move-update kind (no bias can be registered as per-domain)
        *o << "Total external bias: " << ene[5 * static_cast<size_t>(r) + 4] << "\n";
        *o << "Domain update external bias: " << 0 << "\n";
        *o << "Move update external bias: " << ene[5 * static_cast<size_t>(r) + 4] << "\n";
        *o << "Movetype: " << (moved >= 0 ? s.movetypes[moved].label : std::string {}) << "\n";
        *o << "Accepted: " << std::boolalpha << accepted << std::noboolalpha << "\n";
        *o << "\n";
        *sink << entry.str();
        sink->flush();
    }
}

long long next_output_step(ldo_sim& s, long long cur, long long end) {
    InputParameters const& p = s.params;
    if (s.files.empty()) return end;
    long long next {end};
    int freqs[] {p.m_configs_output_freq, p.m_counts_output_freq, p.m_times_output_freq, p.m_energies_output_freq, p.m_order_params_output_freq,
                 p.m_vtf_output_freq, p.m_rand_engine_state_output_freq, s.is_us ? 0 : p.m_logging_freq};
    for (int f: freqs) {
        if (f == 0) continue;
        long long n {(cur / f + 1) * f};
        if (n < next) next = n;
    }
    return next;
}

// GCMCSimulation::simulate (simulation.cpp:568-653) for every replica; returns false when max_duration hit
bool simulate(ldo_sim& s, long long steps) {
    InputParameters const& p = s.params;
    long long end {s.step + steps};
    while (s.step < end) {
        long long stop {next_output_step(s, s.step, end)};
        // chunks are bounded so that the wall-clock limit is honoured with useful granularity
        long long chunk {std::min<long long>(stop - s.step, 100000)};
        // write_log_entry names the movetype of the step it reports and whether it was accepted (simulation.cpp:628-630,
        // 701-702): the move that lands on a logging step runs in a launch of its own, between two reads of the counters
        bool log_here {!s.is_us && due(p.m_logging_freq, s.step + chunk) && s.step + chunk == stop};
        std::vector<long long> att0, acc0;
        if (log_here) {
            if (chunk > 1) s.check(ldo_run(s.eng, chunk - 1, p.m_centering_freq, p.m_centering_domain, p.m_constraint_check_freq));
            att0.resize(s.movetypes.size() * s.R);
            acc0.resize(att0.size());
            s.check(ldo_get_move_stats(s.eng, att0.data(), acc0.data()));
            s.check(ldo_run(s.eng, 1, p.m_centering_freq, p.m_centering_domain, p.m_constraint_check_freq));
        }
        else {
            s.check(ldo_run(s.eng, chunk, p.m_centering_freq, p.m_centering_domain, p.m_constraint_check_freq));
        }
        s.step += chunk;
        std::vector<int> status(s.R), detail(s.R);
        s.check(ldo_get_status(s.eng, status.data(), detail.data()));
        for (int r {0}; r != s.R; r++) {
            if (status[r] != 0) {
                throw OrigamiMisuse {
                        "replica " + std::to_string(s.rank * s.R + r) + " stopped with status " +
                        std::to_string(status[r]) + " (detail " + std::to_string(detail[r]) + ")"};
            }
        }
        double dt {std::chrono::duration<double>(std::chrono::steady_clock::now() - s.start).count()};
        if (dt > p.m_max_duration) {
            // (simulation.cpp:621-625: the limit is tested before the step's log entry and output)
            std::cout << "Maximum time allowed reached" << std::endl;
            return false;
        }
        if (log_here) write_log_entries(s, s.step, att0, acc0);
        write_outputs(s, s.step);
    }
    return true;
}

// write_log_summary (simulation.cpp:706-718, movetypes.cpp:87-96)
// GCMCSimulation::write_log_summary (simulation.cpp:706-718): <filebase>.moves holds, for every movetype but the
// orientation rotation (whose write_log_summary is empty, orientation_movetype.cpp:28), the header of
// movetypes.cpp:88-96 and the movetype's own breakdown from its trackers (met_movetypes.cpp:145-190, 442-468;
// cb_movetypes.cpp:264-290, 626-675, 816-870; rg_movetypes.cpp:593-627, 754-786; transform_movetypes.cpp:64-139).
void write_move_summary(ldo_sim& s) {
    if (s.params.m_output_filebase.empty()) return;
    size_t n {s.movetypes.size()};
    std::vector<long long> att(n * s.R), acc(n * s.R);
    s.check(ldo_get_move_stats(s.eng, att.data(), acc.data()));
    int nst {static_cast<int>(s.sysfile->identities.size()) - 1};
    std::vector<int> sticky(2 * n);
    std::vector<unsigned int> counts(n * 2 * LDO_TRACKER_BINS * 2);
    std::vector<int> lk_sticky(6 * n), lk_entries(6 * LDO_LINKER_TRACKER_CAP);
    std::vector<double> mults(static_cast<size_t>(s.R) * std::max(nst, 1));
    for (int r {0}; r != s.R; r++) {
        std::ofstream o {replica_filebase(s, r) + ".moves"};
        bool typed {ldo_get_move_trackers(s.eng, r, sticky.data(), counts.data()) == 0};
        int lk_n {0}, lk_dropped {0};
        if (typed) s.check(ldo_get_linker_trackers(s.eng, r, lk_sticky.data(), &lk_n, lk_entries.data(), &lk_dropped));
        if (lk_dropped != 0) std::cerr << "warning: " << lk_dropped << " linker tracker updates of replica " << r << " were dropped (list full)\n";
        auto cnt = [&](size_t i, int field, int value, int what) {
            return static_cast<int>(counts[((i * 2 + field) * LDO_TRACKER_BINS + value) * 2 + what]);
        };
        auto block = [&](std::string const& title, int value, int ats, int acs) {
            double freq {static_cast<double>(acs) / ats};
            o << "    " << title << ": " << value << "\n";
            o << "        Attempts: " << ats << "\n";
            o << "        Accepts: " << acs << "\n";
            o << "        Frequency: " << freq << "\n";
        };
        for (size_t i {0}; i != n; i++) {
            int type {s.movetypes[i].desc.type};
            if (type == LDO_MT_ORIENTATION_ROTATION) continue;
            long long a {att[static_cast<size_t>(r) * n + i]}, c {acc[static_cast<size_t>(r) * n + i]};
            o << "Movetype: " << s.movetypes[i].label << "\n";
            o << "    Attempts: " << a << "\n";
            o << "    Accepts: " << c << "\n";
            o << "    Frequency: " << static_cast<double>(c) / a << "\n";
            if (!typed) continue;
            if (type == LDO_MT_MET_STAPLE_EXCHANGE) {
                std::set<int> types;
                for (int f {0}; f != 2; f++)
                    for (int v {0}; v != LDO_TRACKER_BINS; v++)
                        if (cnt(i, f, v, 0) != 0) types.insert(v);
                s.check(ldo_get_exchange_mults(s.eng, static_cast<int>(i), mults.data()));
                o << "       Exchange multipliers\n        ";
                for (size_t k {0}; k != types.size(); k++) o << mults[static_cast<size_t>(r) * nst + k] << ", ";
                o << "\n";
                for (int st: types) {
                    o << "    Staple type: " << st << "\n";
                    int iats {cnt(i, 0, st, 0)}, iacs {cnt(i, 0, st, 1)};
                    float ifreq {static_cast<float>(iacs) / iats};
                    o << "        Insertion attempts: " << iats << "\n";
                    o << "        Insertion accepts: " << iacs << "\n";
                    o << "        Insertion frequency: " << ifreq << "\n";
                    int dats {cnt(i, 1, st, 0)}, dacs {cnt(i, 1, st, 1)};
                    float dfreq {static_cast<float>(dacs) / dats};
                    o << "        Deletion attempts: " << dats << "\n";
                    o << "        Deletion accepts: " << dacs << "\n";
                    o << "        Deletion frequency: " << dfreq << "\n";
                }
            }
            else if (type == LDO_MT_MET_STAPLE_REGROWTH || type == LDO_MT_CB_STAPLE_REGROWTH) {
                // Met: every tracker's staple type is listed, counted only when staples were present (:449-457);
                // CB: only the trackers with staples present are listed, after an empty line (:271-280)
                bool met {type == LDO_MT_MET_STAPLE_REGROWTH};
                std::set<int> types;
                for (int f {0}; f != (met ? 2 : 1); f++)
                    for (int v {0}; v != LDO_TRACKER_BINS; v++)
                        if (cnt(i, f, v, 0) != 0) types.insert(v);
                if (!met) o << "\n";
                for (int st: types) block("Staple type", st, cnt(i, 0, st, 0), cnt(i, 0, st, 1));
            }
            else if (type == LDO_MT_CTRG_SCAFFOLD_REGROWTH || type == LDO_MT_CTRG_JUMP_SCAFFOLD_REGROWTH ||
                     type == LDO_MT_CTCB_SCAFFOLD_REGROWTH || type == LDO_MT_CTCB_JUMP_SCAFFOLD_REGROWTH) {
                for (int v {0}; v != LDO_TRACKER_BINS; v++)
                    if (cnt(i, 0, v, 0) != 0) block("Number of scaffold domains", v, cnt(i, 0, v, 0), cnt(i, 0, v, 1));
                o << "\n";
                if (type == LDO_MT_CTCB_SCAFFOLD_REGROWTH || type == LDO_MT_CTCB_JUMP_SCAFFOLD_REGROWTH) {
                    for (int v {0}; v != LDO_TRACKER_BINS; v++)
                        if (cnt(i, 1, v, 0) != 0) block("Number of staples", v, cnt(i, 1, v, 0), cnt(i, 1, v, 1));
                }
            }
            else if (type == LDO_MT_CTCB_LINKER_REGROWTH || type == LDO_MT_CTCB_CLUSTERED_LINKER_REGROWTH || type == LDO_MT_CTRG_LINKER_REGROWTH) {
                // LinkerRegrowthMCMovetype::write_log_summary (transform_movetypes.cpp:64-139): three tables keyed by pairs,
                // each in ascending order of the pair, an empty line between them
                char const* titles[] {"Number of linker/central domains", "Number of linker/central staples", "Sum/number of displacement/turns"};
                for (int table {0}; table != 3; table++) {
                    std::map<std::pair<int, int>, std::pair<int, int>> rows;
                    for (int k {0}; k != lk_n; k++) {
                        int const* en {&lk_entries[6 * static_cast<size_t>(k)]};
                        if (en[0] == static_cast<int>(i) && en[1] == table) rows[{en[2], en[3]}] = {en[4], en[5]};
                    }
                    if (table != 0) o << "\n";
                    for (auto const& row: rows) {
                        double freq {static_cast<double>(row.second.second) / row.second.first};
                        o << "    " << titles[table] << ": " << row.first.first << "/" << row.first.second << "\n";
                        o << "        Attempts: " << row.second.first << "\n";
                        o << "        Accepts: " << row.second.second << "\n";
                        o << "        Frequency: " << freq << "\n";
                    }
                }
            }
        }
    }
}

void set_all_control(ldo_sim& s, int temp_idx) {
    std::vector<int> ti(s.R, temp_idx);
    s.check(ldo_set_control(s.eng, 0, s.R, ti.data(), nullptr, nullptr, nullptr));
}

// ---------------------------------------------------------------------------------------------
// Umbrella sampling drivers (us_simulation.cpp): single window, multi-window, replica-exchange MW
// ---------------------------------------------------------------------------------------------

using GridPoint = std::vector<int>;

struct UsWindowState { // per window slot: USGCMCSimulation members m_S_n, m_s_i, m_f_i, m_E_w, m_p_i, m_w_i
    std::set<GridPoint> S_n, s_i;
    std::map<GridPoint, long long> f_i;
    std::map<GridPoint, double> E_w, p_i, w_i;
    std::vector<GridPoint> old_only;
    long long steps {0};
    std::unique_ptr<std::ofstream> us_stream;
};

std::vector<UsWindowState> g_unused_us_state; // (state lives in ldo_sim via a side table keyed by pointer)
std::map<ldo_sim*, std::vector<UsWindowState>> g_us_states;

// Upper bound of an order parameter, for the dense grid box the device uses in place of the reference's map
int op_upper_bound(ldo_sim& s, int op) {
    OrderParamSpec const& o = s.ops[op];
    int n_scaf {static_cast<int>(s.sysfile->identities[0].size())};
    int max_domains {n_scaf + s.params.m_max_total_staples * s.params.m_max_staple_size};
    if (o.type == "NumStaples") return s.params.m_max_total_staples;
    if (o.type == "NumStaplesType") return s.params.m_max_type_staples;
    if (o.type == "StapleTypeFullyBound") return 1;
    if (o.type == "NumBoundDomainPairs") return max_domains / 2;
    if (o.type == "NumMisboundDomainPairs") return max_domains / 2;
    if (o.type == "NumStackedPairs") return 2 * max_domains;
    if (o.type == "Sum") {
        int sum {0};
        for (int k: o.sum_ops) sum += op_upper_bound(s, k);
        return sum;
    }
    return 0;
}

void us_setup(ldo_sim& s) {
    InputParameters const& p = s.params;
    for (size_t b {0}; b != s.biases.size(); b++)
        if (s.biases[b].tag == p.m_us_grid_bias_tag && s.biases[b].type == "Grid") s.grid_bias = static_cast<int>(b);
    if (s.grid_bias < 0) throw SimulationMisuse {"us_grid_bias_tag does not name a Grid bias function"};
    size_t total {1};
    for (int op: s.biases[s.grid_bias].ops) {
        s.grid_lo.push_back(0);
        s.grid_n.push_back(op_upper_bound(s, op) + 1);
        total *= s.grid_n.back();
    }
    if (total > 1024) throw SimulationMisuse {"grid bias box exceeds the device capacity of 1024 points"};
    s.n_windows = 1;
    if (s.is_mwus) {
        s.windows = read_windows_file(p.m_windows_file);
        s.n_windows = static_cast<int>(s.windows.mins.size());
        if (s.n_windows < 1 || s.R % s.n_windows != 0) throw SimulationMisuse {"replica count must be a multiple of the number of windows"};
        int wb {-1};
        for (size_t b {0}; b != s.biases.size(); b++)
            if (s.biases[b].tag == s.windows.bias_tag) wb = static_cast<int>(b);
        if (wb < 0 || s.biases[wb].type == "Grid") throw SimulationMisuse {"windows file does not name a well bias function"};
        s.window_biases = {wb};
        for (int w {0}; w != s.n_windows; w++) {
            // "_win-" + mins joined "-" + "-" + maxs each prefixed "-" (us_simulation.cpp:488-497)
            std::string post {"_win-"};
            for (int v: s.windows.mins[w]) post += std::to_string(v) + "-";
            for (int v: s.windows.maxs[w]) post += "-" + std::to_string(v);
            s.window_postfix.push_back(post);
        }
        // (the window limits themselves are applied by us_apply_window_limits once the starting configuration is set)
    }
    // empty grid (every point off-grid, bias 0) for every replica
    std::vector<double> vals(total, std::nan(""));
    for (int r {0}; r != s.R; r++) s.check(ldo_set_grid_bias(s.eng, r, s.grid_bias, s.grid_lo.data(), s.grid_n.data(), vals.data()));
    auto& st = g_us_states[&s];
    st.clear();
    st.resize(s.R);
    if (s.is_ptmwus) {
        s.q2r.resize(s.R);
        for (int r {0}; r != s.R; r++) s.q2r[r] = r % s.n_windows;
        s.attempts.assign(static_cast<size_t>(s.R / s.n_windows) * std::max(s.n_windows - 1, 1), 0);
        s.accepts = s.attempts;
    }
}

// MWUSGCMCSimulation::setup_window_restraints (us_simulation.cpp:503-516) overrides min_op / max_op of the window
// bias AFTER SystemBiases was constructed and evaluated (origami_system.cpp:69-70): the stored bias value is still the
// one of the file's limits until the first move re-evaluates it, so that move sees the whole jump as its bias change.
// Same order here: the configuration is loaded (biases evaluated with the file's limits) before the limits change.
void us_apply_window_limits(ldo_sim& s) {
    if (!s.is_mwus) return;
    for (int r {0}; r != s.R; r++) {
        int w {r % s.n_windows};
        s.check(ldo_set_window(s.eng, r, s.window_biases[0], s.windows.mins[w][0], s.windows.maxs[w][0]));
    }
}

size_t us_point_index(ldo_sim& s, GridPoint const& pt) {
    size_t idx {0};
    for (size_t k {0}; k != pt.size(); k++) idx = idx * s.grid_n[k] + static_cast<size_t>(pt[k] - s.grid_lo[k]);
    return idx;
}

// replica currently holding window slot `slot` (slot = ladder * n_windows + window)
int us_replica_of_slot(ldo_sim& s, int slot) {
    if (!s.is_ptmwus) return slot;
    int ladder {slot / s.n_windows}, w {slot % s.n_windows};
    return ladder * s.n_windows + s.q2r[static_cast<size_t>(ladder) * s.n_windows + w];
}

void us_upload_bias(ldo_sim& s, int slot) {
    UsWindowState& ws = g_us_states[&s][slot];
    size_t total {1};
    for (int n: s.grid_n) total *= n;
    std::vector<double> vals(total, std::nan(""));
    for (auto const& kv: ws.E_w) vals[us_point_index(s, kv.first)] = kv.second;
    s.check(ldo_set_grid_bias(s.eng, us_replica_of_slot(s, slot), s.grid_bias, s.grid_lo.data(), s.grid_n.data(), vals.data()));
}

// USGCMCSimulation::output_weights (us_simulation.cpp:207-244); entries sorted by point
void us_output_weights(ldo_sim& s, int slot, std::string const& filename) {
    UsWindowState& ws = g_us_states[&s][slot];
    std::ofstream f {filename};
    f << "{\n    \"biases\": [";
    bool first {true};
    for (auto const& kv: ws.E_w) {
        f << (first ? "\n" : ",\n");
        first = false;
        f << "        {\n            \"point\": [\n";
        for (size_t i {0}; i + 1 < kv.first.size(); i++) f << "                " << std::to_string(kv.first[i]) << ",\n";
        f << "                " << std::to_string(kv.first.back()) << "\n            ],\n";
        f << "            \"bias\": " << std::to_string(kv.second) << "\n        }";
    }
    f << "\n    ]\n}\n";
}

void us_clear_visits(ldo_sim& s) {
    size_t total {1};
    for (int n: s.grid_n) total *= n;
    std::vector<long long> counts(total);
    for (int r {0}; r != s.R; r++) s.check(ldo_get_grid_visits(s.eng, r, s.grid_bias, counts.data(), 1));
}

// process_iteration (us_simulation.cpp:154-170) for one window slot
void us_process_iteration(ldo_sim& s, int slot, int n, long long steps) {
    InputParameters const& p = s.params;
    UsWindowState& ws = g_us_states[&s][slot];
    size_t total {1};
    for (int k: s.grid_n) total *= k;
    std::vector<long long> counts(total);
    s.check(ldo_get_grid_visits(s.eng, us_replica_of_slot(s, slot), s.grid_bias, counts.data(), 1));
    ws.s_i.clear();
    ws.f_i.clear();
    for (size_t idx {0}; idx != total; idx++) {
        if (counts[idx] == 0) continue;
        GridPoint pt(s.grid_n.size());
        size_t rem {idx};
        for (size_t k {s.grid_n.size()}; k-- > 0;) {
            pt[k] = static_cast<int>(rem % s.grid_n[k]) + s.grid_lo[k];
            rem /= s.grid_n[k];
        }
        ws.s_i.insert(pt);
        ws.f_i[pt] = counts[idx];
    }
    // fill_grid_sets (:307-340): points seen before but not in this iteration
    ws.old_only.clear();
    for (auto const& pt: ws.S_n)
        if (!ws.s_i.count(pt)) ws.old_only.push_back(pt);
    ws.S_n.insert(ws.s_i.begin(), ws.s_i.end());
    // estimate_current_weights (:286-305)
    auto bias_of = [&](GridPoint const& pt) {
        auto it = ws.E_w.find(pt);
        return it == ws.E_w.end() ? 0.0 : it->second;
    };
    double ave {0};
    for (auto const& pt: ws.s_i) ave += ws.f_i[pt] * std::exp(bias_of(pt));
    for (auto const& pt: ws.s_i) ws.p_i[pt] = ws.f_i[pt] * std::exp(bias_of(pt)) / ave;
    for (auto const& pt: ws.old_only) ws.p_i[pt] = 0;
    // update_grids (:418-434)
    for (auto const& pt: ws.s_i) ws.w_i[pt] = static_cast<double>(ws.f_i[pt]) / steps;
    for (auto const& pt: ws.old_only) ws.w_i[pt] = 0;
    // SimpleUSGCMCSimulation::update_bias (:385-416)
    for (auto const& pt: ws.S_n) {
        double old_bias {ws.E_w[pt]};
        double p_k_n {ws.p_i[pt]};
        double D_bias, new_bias;
        if (p_k_n == 0) {
            D_bias = -p.m_max_D_bias;
            new_bias = old_bias + D_bias;
        }
        else {
            new_bias = std::log(p_k_n);
            D_bias = new_bias - old_bias;
        }
        double updated {new_bias};
        if (std::abs(D_bias) > p.m_max_D_bias) updated = D_bias > 0 ? old_bias + p.m_max_D_bias : old_bias - p.m_max_D_bias;
        ws.E_w[pt] = updated;
    }
    us_upload_bias(s, slot);
    // output_summary (:436-452)
    // the summary goes to the window's .out stream (MWUS, us_simulation.cpp:545-547) or to stdout (single window:
    // m_us_stream stays &cout, us_simulation.hpp:77)
    bool to_stdout {!ws.us_stream && !s.is_mwus && !std::getenv("LDO_QUIET")};
    if (ws.us_stream || to_stdout) {
        std::ostream& o = ws.us_stream ? static_cast<std::ostream&>(*ws.us_stream) : std::cout;
        o << "Iteration: " << n << "\n\nGridpoint w, P, E:\n";
        for (auto const& pt: ws.S_n) {
            for (int c: pt) o << c << " ";
            o << std::setprecision(3) << ": " << std::setw(10) << ws.w_i[pt] << std::setw(10) << ws.p_i[pt] << std::setw(10) << ws.E_w[pt] << "\n";
        }
        o << "\n";
        o.flush();
    }
}

void us_open_iteration_files(ldo_sim& s, std::string const& postfix) {
    s.filebase_postfix = postfix;
    s.files.clear();
    open_output_files(s);
    s.file_owner.resize(s.R);
    for (int slot {0}; slot != s.R; slot++) s.file_owner[slot] = us_replica_of_slot(s, slot);
}

void us_run(ldo_sim& s) {
    InputParameters const& p = s.params;
    auto& st = g_us_states[&s];
    int n_ladders {s.R / s.n_windows};
    // per-window summary streams "<filebase><window postfix>.out" (us_simulation.cpp:545-547)
    if (!p.m_output_filebase.empty()) {
        for (int slot {0}; slot != s.R; slot++) {
            s.filebase_postfix = "";
            // The constructor messages ("No biases read in", ...) are written before MWUS redirects the stream
            // (us_simulation.cpp:68-90 vs :545-547): they go to stdout, the .out file starts with the first summary
            if (!std::getenv("LDO_QUIET")) {
                std::cout << (p.m_read_biases ? "Reading biases from file\n" : "No biases read in\n")
                          << (p.m_restart_from_config ? "Reading starting config from file\n" : "Starting from configuration in system file\n")
                          << "Starting new iteration\n";
            }
            if (s.is_mwus) st[slot].us_stream.reset(new std::ofstream {replica_filebase(s, slot) + ".out"});
        }
    }
    if (p.m_restart_us_iter) {
        throw NotImplemented {"restart_us_iter reads Boost text archives (.S_n / .s_i / .f_i, us_simulation.cpp:300-340): not available on the device path; restart with read_biases and restart_from_config"};
    }
    if (p.m_read_biases) {
        // USGCMCSimulation constructor: read_weights + GridBiasFunction::replace_biases (us_simulation.cpp:50-66, 192-205;
        // bias_functions.cpp:237-240); the multi-window drivers read <biases_filebase><window postfix>.biases (:526-529)
        for (int slot {0}; slot != s.R; slot++) {
            std::string file {s.is_mwus ? p.m_biases_filebase + s.window_postfix[slot % s.n_windows] + ".biases" : p.m_biases_file};
            std::ifstream f {file};
            if (!f) throw FileError {"Restart bias file " + file + " does not exist"};
            std::stringstream buf;
            buf << f.rdbuf();
            try {
                Json root {Json::parse(buf.str())};
                st[slot].E_w.clear();
                for (size_t k {0}; k != root["biases"].size(); k++) {
                    GridPoint pt;
                    for (size_t c {0}; c != root["biases"][k]["point"].size(); c++) pt.push_back(root["biases"][k]["point"][c].as_int());
                    st[slot].E_w[pt] = root["biases"][k]["bias"].as_double();
                }
            } catch (std::exception const& e) {
                throw FileError {"Restart bias file " + file + " is not well formed:\n" + e.what()};
            }
            us_upload_bias(s, slot);
        }
    }
    double saved_max_duration {s.params.m_max_duration};
    // run_equilibration (:93-116)
    if (!(s.is_ptmwus && (p.m_restart_us_iter || p.m_read_biases))) {
        us_open_iteration_files(s, "_iter-equil");
        s.params.m_max_duration = static_cast<double>(p.m_max_equil_dur);
        s.start = std::chrono::steady_clock::now();
        s.step = 0;
        simulate(s, p.m_equil_steps);
        us_clear_visits(s);
    }
    for (int n {0}; n != p.m_max_num_iters; n++) {
        // prepare_iteration (:127-147)
        us_open_iteration_files(s, "_iter-" + std::to_string(n));
        if (!p.m_output_filebase.empty())
            for (int slot {0}; slot != s.R; slot++) us_output_weights(s, slot, replica_filebase(s, slot) + "-inp.biases");
        s.params.m_max_duration = static_cast<double>(p.m_max_iter_dur);
        s.start = std::chrono::steady_clock::now();
        s.step = 0;
        long long steps_done {0};
        if (s.is_ptmwus) {
            // run_swaps (:662-706); swap file "<filebase>_iter-n.swp" (:900-918), first ladder
            std::ofstream swp;
            if (!p.m_output_filebase.empty()) swp.open(p.m_output_filebase + "_iter-" + std::to_string(n) + ".swp");
            auto write_swap_entry = [&](long long step) {
                if (!swp.is_open() || p.m_configs_output_freq == 0 || step % p.m_configs_output_freq != 0) return;
                for (int w {0}; w != s.n_windows; w++) swp << s.q2r[w] << " ";
                swp << "\n";
            };
            write_swap_entry(0);
            for (long long swap_i {1}; swap_i != p.m_iter_swaps + 1; swap_i++) {
                bool more {simulate(s, p.m_exchange_interval)};
                steps_done += p.m_exchange_interval;
                if (!more) break;
                write_swap_entry(s.step);
                s.check(ldo_exchange_windows(
                        s.eng, swap_i, n_ladders, s.n_windows, s.grid_bias, static_cast<int>(s.window_biases.size()),
                        s.window_biases.data(), s.q2r.data(), s.attempts.data(), s.accepts.data()));
                for (int slot {0}; slot != s.R; slot++) s.file_owner[slot] = us_replica_of_slot(s, slot);
            }
            write_swap_entry(s.step);
            // write_acceptance_freqs (:884-898)
            std::cout << "Iter " << n << "\nWindow 1, Window 2, Swaps, Attempts, Frequency\n";
            for (int w {0}; w + 1 < s.n_windows; w++) {
                std::cout << w << " " << w + 1 << " " << s.accepts[w] << " " << s.attempts[w] << " "
                          << static_cast<double>(s.accepts[w]) / s.attempts[w] << "\n";
            }
            std::cout << "\n";
            std::fill(s.attempts.begin(), s.attempts.end(), 0);
            std::fill(s.accepts.begin(), s.accepts.end(), 0);
        }
        else {
            simulate(s, p.m_iter_steps);
            // run_simulation stores simulate()'s return value - the last step number plus one (simulation.cpp:574,652) -
            // in m_steps, which update_grids divides the visit counts by (us_simulation.cpp:146-151, 418-422)
            steps_done = s.step + 1;
        }
        // process_iteration
        for (int slot {0}; slot != s.R; slot++) {
            us_process_iteration(s, slot, n, steps_done);
            if (!p.m_output_filebase.empty()) us_output_weights(s, slot, replica_filebase(s, slot) + ".biases");
        }
    }
    s.params.m_max_duration = saved_max_duration;
    s.filebase_postfix = "";
}

} // namespace

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time (only a multi-GPU run needs it: the single-GPU library has no NCCL dependency, and a
// process that already carries a libnccl - PyTorch's - shares that one). Minimal declarations of nccl.h.
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclId {
    char internal[128];
};
struct NcclApi {
    void* lib {nullptr};
    int (*GetUniqueId)(NcclId*) {nullptr};
    int (*CommInitRank)(void**, int, NcclId, int) {nullptr};
    int (*AllGather)(const void*, void*, size_t, int, void*, void*) {nullptr};
    int (*CommDestroy)(void*) {nullptr};
    const char* (*GetErrorString)(int) {nullptr};
};
const int NCCL_FLOAT64 {8};

NcclApi& nccl() {
    static NcclApi api;
#ifndef LDO_HOSTSIM
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name: {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(api.lib, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(api.lib, "ncclCommInitRank"));
        api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, void*)>(dlsym(api.lib, "ncclAllGather"));
        api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
    });
#endif
    return api;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact enumeration (simulation_type=enumerate; enumerate.cpp). The staple sets (StapleEnumerator, :1070-1133) and the
// growthpoint sets (GrowthpointEnumerator and its two subclasses, :813-1068) are enumerated here exactly as the reference
// does, duplicates included and skipped the same way; every growthpoint set is handed to the device
// (ldo_enumerate_conformations), which grows all conformations on all replica slots at once. A domain is (unique chain
// index, domain index) like the reference's Domain*.
// ---------------------------------------------------------------------------------------------------------------
struct HostEnumerator {
    using Dom = std::pair<int, int>;
    using Growthpoint = std::pair<std::pair<int, int>, std::pair<int, int>>;
    ldo_sim& s;
    InputParameters const& p;
    std::vector<std::vector<int>> const& identities;
    bool quiet;
    // ConformationalEnumerator state
    std::vector<std::pair<int, int>> live; // staples present: (unique chain index, identity), in the order they were added
    int next_c_i {0};
    std::map<int, std::vector<int>> identity_to_indices;
    std::map<Dom, Dom> conf_growthpoints;
    std::map<int, int> ident_unassigned;
    long double prefix {1};
    std::map<std::vector<int>, long double> state_weights;
    long double partition_f {0}, average_energy {0}, average_bias {0}, num_configs {0};
    long long leaves {0};
    // GrowthpointEnumerator state
    bool misbinding;
    std::vector<std::pair<int, int>> staples; // identity, copies
    std::vector<std::vector<Growthpoint>> enumerated_growthpoints;
    std::vector<Growthpoint> growthpoints;
    std::vector<Dom> unbound_system_domains; // MisbindingGrowthpointEnumerator
    std::vector<std::vector<Dom>> comp_system_domains; // NoMisbindingGrowthpointEnumerator, per staple identity
    // StapleEnumerator state
    int num_staple_combos {0}, cur_max_total_staples {0}, max_type_staples {0};

    explicit HostEnumerator(ldo_sim& sim):
            s {sim}, p {sim.params}, identities {sim.sysfile->identities}, quiet {std::getenv("LDO_QUIET") != nullptr} {
        if (p.m_max_staple_size != 2 && p.m_misbinding_pot != "Disallowed") throw NotImplemented {"No enumerator available for system"}; // :26-34
        misbinding = p.m_misbinding_pot != "Disallowed";
        for (int d_ident: identities[0]) ident_unassigned[d_ident] = 1; // :204-206
        int scaffold_len {static_cast<int>(identities[0].size())};
        if (misbinding) {
            for (int d {0}; d != scaffold_len; d++) unbound_system_domains.push_back({0, d}); // :875
        }
        else {
            // complementary_scaffold_domains (origami_system.cpp:125-128, 630-655): per staple domain, the first
            // scaffold domain of the opposite identity
            for (size_t c_ident {1}; c_ident != identities.size(); c_ident++) {
                comp_system_domains.push_back({});
                for (int sd: identities[c_ident]) {
                    for (int d {0}; d != scaffold_len; d++) {
                        if (sd == -identities[0][d]) {
                            comp_system_domains.back().push_back({0, d});
                            break;
                        }
                    }
                }
            }
        }
    }

    int chain_number(int c_i) const { // device numbering: 0 scaffold, 1 + k for the k-th staple present
        if (c_i == 0) return 0;
        for (size_t k {0}; k != live.size(); k++)
            if (live[k].first == c_i) return static_cast<int>(k) + 1;
        throw std::logic_error("enumeration: unknown chain");
    }
    int chain_identity(int c_i) const { return c_i == 0 ? 0 : live[chain_number(c_i) - 1].second; }

    // ---- ConformationalEnumerator bookkeeping (:260-342) ----
    void add_staple(int staple) {
        prefix *= p.m_staple_M; // m_reduced_fugacity (origami_system.cpp:58)
        next_c_i++;
        live.push_back({next_c_i, staple});
        size_t len {identities[staple].size()};
        prefix /= static_cast<long double>(std::pow(6, 2 * len - 1));
        identity_to_indices[staple].push_back(next_c_i);
        for (int d_ident: identities[staple]) {
            if (ident_unassigned.count(d_ident) == 0) ident_unassigned[d_ident] = 1;
            else ident_unassigned[d_ident]++;
        }
    }
    void remove_staple(int staple) {
        prefix /= p.m_staple_M;
        size_t len {identities[staple].size()};
        prefix *= static_cast<long double>(std::pow(6, 2 * len - 1));
        int c_i {identity_to_indices[staple].back()};
        identity_to_indices[staple].pop_back();
        for (size_t k {0}; k != live.size(); k++) {
            if (live[k].first == c_i) {
                live.erase(live.begin() + k);
                break;
            }
        }
        for (int d_ident: identities[staple]) ident_unassigned[d_ident]--;
    }
    std::vector<Dom> add_growthpoint(int new_c_ident, int new_d_i, Dom old_domain) {
        int new_c_i {identity_to_indices[new_c_ident].back()};
        identity_to_indices[new_c_ident].pop_back();
        conf_growthpoints[old_domain] = {new_c_i, new_d_i};
        std::vector<Dom> rest;
        for (int d {0}; d != static_cast<int>(identities[new_c_ident].size()); d++)
            if (d != new_d_i) rest.push_back({new_c_i, d});
        return rest;
    }
    void remove_growthpoint(Dom old_domain) {
        Dom new_domain {conf_growthpoints[old_domain]};
        conf_growthpoints.erase(old_domain);
        identity_to_indices[chain_identity(new_domain.first)].push_back(new_domain.first);
    }

    // create_domains_stack / create_staple_stack (:591-637): the order in which the domains are popped
    void staple_stack(Dom domain, std::vector<Dom>& order) {
        int len {static_cast<int>(identities[chain_identity(domain.first)].size())};
        order.push_back(domain);
        for (int d_i {domain.second + 1}; d_i != len; d_i++) {
            Dom d {domain.first, d_i};
            order.push_back(d);
            if (conf_growthpoints.count(d)) staple_stack(conf_growthpoints[d], order);
        }
        for (int d_i {domain.second - 1}; d_i != -1; d_i--) {
            Dom d {domain.first, d_i};
            order.push_back(d);
            if (conf_growthpoints.count(d)) staple_stack(conf_growthpoints[d], order);
        }
    }

    // ConformationalEnumerator::enumerate (:218-258) - on the device
    void enumerate_conformations() {
        // create_domains_stack: the scaffold's domains with the staples that grow off them (:591-608), or - staples only -
        // those staples alone (:793-811)
        std::vector<Dom> order;
        for (int d {0}; d != static_cast<int>(identities[0].size()); d++) {
            if (!p.m_enumerate_staples_only) order.push_back({0, d});
            if (conf_growthpoints.count({0, d})) staple_stack(conf_growthpoints[{0, d}], order);
        }
        std::vector<int> staple_type, stack_chain, stack_d, go_c, go_d, gn_c, gn_d;
        for (auto const& st: live) staple_type.push_back(st.second);
        for (Dom const& d: order) {
            stack_chain.push_back(chain_number(d.first));
            stack_d.push_back(d.second);
        }
        for (auto const& gp: conf_growthpoints) {
            go_c.push_back(chain_number(gp.first.first));
            go_d.push_back(gp.first.second);
            gn_c.push_back(chain_number(gp.second.first));
            gn_d.push_back(gp.second.second);
        }
        int n_ident {0};
        for (auto const& chain: identities)
            for (int d_ident: chain) n_ident = std::max(n_ident, std::abs(d_ident));
        std::vector<int> unassigned(2 * n_ident + 1, 0);
        for (auto const& kv: ident_unassigned) unassigned[kv.first + n_ident] = kv.second;
        ldo_enum_job job {};
        job.n_staples = static_cast<int>(staple_type.size());
        job.staple_type = staple_type.data();
        job.n_stack = static_cast<int>(order.size());
        job.stack_chain = stack_chain.data();
        job.stack_d = stack_d.data();
        job.n_growthpoints = static_cast<int>(go_c.size());
        job.gp_old_chain = go_c.data();
        job.gp_old_d = go_d.data();
        job.gp_new_chain = gn_c.data();
        job.gp_new_d = gn_d.data();
        job.n_ident = n_ident;
        job.ident_unassigned = unassigned.data();
        job.overcount = p.m_max_staple_size == 2 ? 0 : 1; // :26-31
        job.n_out_ops = static_cast<int>(s.ops_out_idx.size());
        job.out_ops = s.ops_out_idx.data();
        job.split_depth = 0;
        job.staples_only = p.m_enumerate_staples_only ? 1 : 0;
        job.scaffold_pos = s.sysfile->chains[0].positions.data();
        job.scaffold_ore = s.sysfile->chains[0].orientations.data();
        const int max_keys {4096};
        std::vector<int> keys(static_cast<size_t>(max_keys) * job.n_out_ops);
        std::vector<double> weights(max_keys);
        double sums[4];
        int n_keys {0};
        long long n_leaves {0};
        s.check(ldo_enumerate_conformations(s.eng, &job, max_keys, &n_keys, keys.data(), weights.data(), sums, &n_leaves));
        // calc_and_save_weights (:639-664) with the staple set's prefix
        for (int k {0}; k != n_keys; k++) {
            std::vector<int> key(keys.begin() + static_cast<size_t>(k) * job.n_out_ops, keys.begin() + static_cast<size_t>(k + 1) * job.n_out_ops);
            state_weights[key] += prefix * static_cast<long double>(weights[k]);
        }
        partition_f += prefix * static_cast<long double>(sums[0]);
        average_energy += prefix * static_cast<long double>(sums[1]);
        average_bias += prefix * static_cast<long double>(sums[2]);
        num_configs += static_cast<long double>(sums[3]);
        leaves += n_leaves;
    }

    // ---- GrowthpointEnumerator (:819-1068) ----
    void enumerate_growthpoints() {
        enumerated_growthpoints.clear();
        staples.clear();
        for (size_t c_ident {1}; c_ident != identities.size(); c_ident++) {
            int n {0};
            for (auto const& st: live) n += st.second == static_cast<int>(c_ident) ? 1 : 0;
            if (n != 0) staples.push_back({static_cast<int>(c_ident), n});
        }
        iterate_staple_identities();
    }
    void iterate_staple_identities() {
        for (size_t i {0}; i != staples.size(); i++) {
            int staple_ident {staples[i].first};
            int num_remaining {staples[i].second - 1};
            if (num_remaining == 0) staples.erase(staples.begin() + i);
            else staples[i] = {staple_ident, num_remaining};
            iterate_domain_growthpoints(staple_ident);
            bool staple_remain {false};
            for (size_t k {0}; k != staples.size(); k++) {
                if (staples[k].first == staple_ident) {
                    staples[k].second++;
                    staple_remain = true;
                    break;
                }
            }
            if (!staple_remain) staples.insert(staples.begin() + i, {staple_ident, num_remaining + 1});
        }
    }
    void iterate_domain_growthpoints(int staple_ident) {
        std::vector<Dom>& avail = misbinding ? unbound_system_domains : comp_system_domains[staple_ident - 1];
        for (size_t j {0}; j != avail.size(); j++) {
            Dom old_domain {avail[j]};
            avail.erase(avail.begin() + j);
            size_t staple_length {identities[staple_ident].size()};
            for (size_t d_i {0}; d_i != staple_length; d_i++) {
                if (!misbinding) {
                    int old_ident {identities[chain_identity(old_domain.first)][old_domain.second]};
                    if (identities[staple_ident][d_i] != -old_ident) continue; // :1001-1005
                }
                recurse_or_enumerate_conf(staple_ident, static_cast<int>(d_i), staple_length, old_domain);
            }
            avail.insert(avail.begin() + j, old_domain);
        }
    }
    void recurse_or_enumerate_conf(int staple_ident, int d_i, size_t staple_length, Dom old_domain) {
        growthpoints.push_back({{staple_ident, d_i}, {old_domain.first, old_domain.second}});
        std::vector<Dom> new_unbound_domains {add_growthpoint(staple_ident, d_i, old_domain)};
        if (!staples.empty()) {
            if (misbinding) {
                unbound_system_domains.insert(unbound_system_domains.end(), new_unbound_domains.begin(), new_unbound_domains.end());
                iterate_staple_identities();
                for (size_t k {0}; k != staple_length - 1; k++) unbound_system_domains.pop_back();
            }
            else {
                iterate_staple_identities();
            }
        }
        else if (!growthpoints_repeated()) {
            enumerate_conformations();
            if (!quiet) std::cout << "   Growthpoint set " << enumerated_growthpoints.size() + 1 << "\n";
            enumerated_growthpoints.push_back(growthpoints);
        }
        remove_growthpoint(old_domain);
        growthpoints.pop_back();
    }
    bool growthpoints_repeated() {
        for (auto const& gps: enumerated_growthpoints) {
            if (gps.size() != growthpoints.size()) continue;
            bool repeated {true};
            for (auto const& gp: gps) {
                if (std::count(growthpoints.begin(), growthpoints.end(), gp) == 0) {
                    repeated = false;
                    break;
                }
            }
            if (repeated) return true;
        }
        return false;
    }

    // ---- StapleEnumerator (:1082-1133) ----
    void enumerate_staples() {
        if (p.m_min_total_staples == 0) {
            enumerate_conformations();
            cur_max_total_staples = 1;
        }
        else {
            cur_max_total_staples = p.m_min_total_staples;
        }
        max_type_staples = p.m_max_type_staples;
        while (cur_max_total_staples <= p.m_max_total_staples) {
            recurse(0, 1, 0);
            cur_max_total_staples++;
        }
    }
    void recurse(int cur_num_staples, int staple_type_i, int cur_num_staples_i) {
        int num_staple_types {static_cast<int>(identities.size()) - 1};
        if (cur_num_staples == cur_max_total_staples) {
            num_staple_combos++;
            if (!quiet) std::cout << "Staple set " << num_staple_combos << "\n";
            enumerate_growthpoints();
            return;
        }
        while (staple_type_i != num_staple_types + 1) {
            bool excluded {std::find(p.m_excluded_staples.begin(), p.m_excluded_staples.end(), staple_type_i) != p.m_excluded_staples.end()};
            if (cur_num_staples_i < max_type_staples && !excluded) {
                add_staple(staple_type_i);
                cur_num_staples_i++;
                cur_num_staples++;
                recurse(cur_num_staples, staple_type_i, cur_num_staples_i);
                cur_num_staples--;
                cur_num_staples_i--;
                remove_staple(staple_type_i);
            }
            staple_type_i++;
            cur_num_staples_i = 0;
        }
    }
};

// enumerate_main (enumerate.cpp:19-81): weights normalised and written to <filebase>.weights, summary on stdout
void enumerate_run(ldo_sim& s) {
    InputParameters const& p = s.params;
    s.ops_out_idx.clear();
    for (auto const& tag: p.m_ops_to_output) {
        int found {-1};
        for (size_t i {0}; i != s.ops.size(); i++)
            if (s.ops[i].tag == tag) found = static_cast<int>(i);
        if (found < 0) throw SimulationMisuse {"ops_to_output: unknown order parameter " + tag};
        s.ops_out_idx.push_back(found);
    }
    bool quiet {std::getenv("LDO_QUIET") != nullptr};
    if (!quiet) std::cout << "Running enumeration\n";
    HostEnumerator en {s};
    en.enumerate_staples();
    if (!p.m_output_filebase.empty()) {
        // print_weights (:361-376); the reference's table is an unordered_map, the states are written in key order here
        std::ofstream out {p.m_output_filebase + ".weights"};
        for (auto const& tag: p.m_ops_to_output) out << tag << " ";
        out << "\n";
        for (auto const& kv: en.state_weights) {
            out << "(" << kv.first[0];
            for (size_t i {1}; i != kv.first.size(); i++) out << " " << kv.first[i];
            out << ") " << static_cast<double>(kv.second / en.partition_f) << "\n";
        }
        out << "\n";
    }
    if (!quiet) {
        std::cout << en.num_configs << "\n\n";
        std::cout << en.average_energy / en.partition_f << "\n";
        std::cout << en.average_bias / en.partition_f << "\n";
    }
    s.enum_leaves = en.leaves;
    s.enum_num_configs = static_cast<double>(en.num_configs);
    s.enum_average_energy = static_cast<double>(en.average_energy / en.partition_f);
    s.enum_average_bias = static_cast<double>(en.average_bias / en.partition_f);
}

void nccl_check(int rc, const char* what) {
    if (rc == 0) return;
    NcclApi& a = nccl();
    throw std::runtime_error(std::string("NCCL ") + what + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "error " + std::to_string(rc)));
}

NcclApi& nccl_or_throw() {
    NcclApi& a = nccl();
    if (!a.lib || !a.GetUniqueId || !a.CommInitRank || !a.AllGather) {
        throw std::runtime_error("NCCL (libnccl.so.2) is not available: multi-GPU replica exchange needs it");
    }
    return a;
}

// One exchange round with nothing but stream-ordered work (ptmc_simulation.cpp:113-141): exchange_interval moves on every
// local replica, collection of the exchange records, NCCL all-gather over the ranks, on-device swap decisions, energy
// rebuild. The host only waits when an output file is due at the end of the interval.
bool exchange_round(ldo_sim& s, long long swap_i) {
    InputParameters const& p = s.params;
    long long end {s.step + p.m_exchange_interval};
    if (next_output_step(s, s.step, end) < end || due(p.m_logging_freq, end)) {
        // an output step falls inside the interval, or the last step of the round is logged (its move runs in a launch
        // of its own): the blocking path writes them where the reference does
        if (!simulate(s, p.m_exchange_interval)) return false;
    }
    else {
        s.check(ldo_run_async(s.eng, p.m_exchange_interval, p.m_centering_freq, p.m_centering_domain, p.m_constraint_check_freq));
        s.step = end;
        if (next_output_step(s, end - 1, end) == end && !s.files.empty()) {
            std::vector<int> status(s.R), detail(s.R);
            s.check(ldo_get_status(s.eng, status.data(), detail.data()));
            for (int r {0}; r != s.R; r++) {
                if (status[r] != 0) {
                    throw OrigamiMisuse {"replica " + std::to_string(s.rank * s.R + r) + " stopped with status " + std::to_string(status[r]) +
                                         " (detail " + std::to_string(detail[r]) + ")"};
                }
            }
            write_outputs(s, s.step);
        }
    }
    if (!s.exchange_resident) {
        s.check(ldo_exchange_state_set(s.eng, static_cast<int>(s.q2r.size()), static_cast<int>(s.attempts.size()), s.q2r.data(),
                                       s.attempts.data(), s.accepts.data()));
        s.exchange_resident = true;
    }
    s.check(ldo_exchange_collect_async(s.eng));
    if (s.n_ranks > 1) {
        if (!s.nccl_comm) throw SimulationMisuse {"replica exchange on several ranks needs ldo_sim_comm_init (or the caller's own all-gather)"};
        nccl_check(nccl().AllGather(s.dep_send, s.dep_recv, static_cast<size_t>(s.R) * s.dep_nq, NCCL_FLOAT64, s.nccl_comm, ldo_stream(s.eng)),
                   "all-gather");
    }
    s.check(ldo_exchange_pt_async(s.eng, s.pt_variant, s.v2_dim, swap_i, s.n_ladders, s.num_reps, s.rank, s.n_ranks));
    return true;
}

void exchange_download(ldo_sim& s) {
    if (!s.exchange_resident) return;
    s.check(ldo_exchange_state_get(s.eng, static_cast<int>(s.q2r.size()), static_cast<int>(s.attempts.size()), s.q2r.data(),
                                   s.attempts.data(), s.accepts.data()));
}

} // namespace

ldo_sim::~ldo_sim() {
    if (nccl_comm && nccl().CommDestroy) nccl().CommDestroy(nccl_comm);
    if (eng) ldo_engine_destroy(eng);
}

extern "C" {

const char* ldo_host_last_error(void) { return g_host_error.c_str(); }

ldo_sim* ldo_sim_create(const char* inp_path, int n_replicas, int device, int rank, int n_ranks) {
    std::unique_ptr<ldo_sim> s {new ldo_sim {}};
    try {
        s->params = InputParameters {inp_path};
        InputParameters& p = s->params;
        s->sysfile.reset(new OrigamiInputFile {p.m_origami_input_filename});
        OrigamiInputFile& sf = *s->sysfile;
        s->R = n_replicas;
        s->rank = rank;
        s->n_ranks = n_ranks > 0 ? n_ranks : 1;

        // potentials (origami_potential.cpp:952-1013)
        if (p.m_binding_pot != "FourBody") throw NotImplemented {p.m_binding_pot + ": No such binding potential"};
        int misbind;
        if (p.m_misbinding_pot == "Opposing") misbind = LDO_MISBIND_OPPOSING;
        else if (p.m_misbinding_pot == "Disallowed") misbind = LDO_MISBIND_DISALLOWED;
        else throw NotImplemented {p.m_misbinding_pot + ": No such misbinding potential"};
        if (p.m_stacking_pot == "SequenceSpecific") {
            throw NotImplemented {"SequenceSpecific stacking is not available on the device path (unfinished in the reference)"};
        }
        if (p.m_stacking_pot != "Constant") throw NotImplemented {p.m_stacking_pot + ": No such stacking potential"};
        int domain_type {domain_type_code(p.m_domain_type)};

        // temperatures
        std::string const& st = p.m_simulation_type;
        if (st == "constant_temp" || st == "umbrella_sampling" || st == "mw_umbrella_sampling" || st == "ptmw_umbrella_sampling") {
            s->temps = {p.m_temp};
            s->is_us = st != "constant_temp";
            s->is_mwus = st == "mw_umbrella_sampling" || st == "ptmw_umbrella_sampling";
            s->is_ptmwus = st == "ptmw_umbrella_sampling";
        }
        else if (st == "annealing") {
            for (double t {p.m_max_temp}; t >= p.m_min_temp; t -= p.m_temp_interval) s->temps.push_back(t);
            if (s->temps.empty()) throw SimulationMisuse {"annealing: empty temperature range"};
        }
        else if (st == "t_parallel_tempering" || st == "ut_parallel_tempering" || st == "hut_parallel_tempering" || st == "st_parallel_tempering") {
            s->is_pt = true;
            s->pt_variant = st == "t_parallel_tempering" ? LDO_PT_T : st == "ut_parallel_tempering" ? LDO_PT_UT : st == "hut_parallel_tempering" ? LDO_PT_HUT : LDO_PT_ST;
            s->num_reps = p.m_num_reps;
            if (static_cast<int>(p.m_temps.size()) != s->num_reps) throw SimulationMisuse {"temps must list num_reps values"};
            if (s->num_reps % s->n_ranks != 0) throw SimulationMisuse {"num_reps must be a multiple of the number of ranks"};
            if (n_replicas % (s->num_reps / s->n_ranks) != 0) {
                throw SimulationMisuse {"replicas per rank must be a multiple of num_reps / ranks"};
            }
            s->n_ladders = n_replicas / (s->num_reps / s->n_ranks);
            s->temps = p.m_temps;
        }
        else if (st == "2d_parallel_tempering") {
            // TwoDPTGCMCSimulation (ptmc_simulation.cpp:428-493): replica (rank) v1_i * v2_dim + v2_i runs at
            // temps[v1_i] with stacking_mults[v2_i]
            s->is_pt = true;
            s->pt_variant = LDO_PT_2D;
            s->v1_dim = static_cast<int>(p.m_temps.size());
            s->v2_dim = static_cast<int>(p.m_stacking_mults.size());
            if (s->v1_dim < 1 || s->v2_dim < 1) throw SimulationMisuse {"2d_parallel_tempering needs temps and stacking_mults"};
            s->num_reps = s->v1_dim * s->v2_dim;
            if (s->num_reps % s->n_ranks != 0) throw SimulationMisuse {"temps x stacking_mults must be a multiple of the number of ranks"};
            if (n_replicas % (s->num_reps / s->n_ranks) != 0) {
                throw SimulationMisuse {"replicas per rank must be a multiple of (temps x stacking_mults) / ranks"};
            }
            s->n_ladders = n_replicas / (s->num_reps / s->n_ranks);
            s->temps = p.m_temps;
        }
        else if (st == "enumerate") {
            s->temps = {p.m_temp};
            s->is_enum = true;
        }
        else {
            throw NotImplemented {st + ": simulation type not available on the device path"};
        }
        for (double t: s->temps) s->tables.push_back(calc_energy_tables(sf, p, t));

        // engine
        std::vector<int> type_len, idents;
        for (auto const& chain: sf.identities) {
            type_len.push_back(static_cast<int>(chain.size()));
            for (int id: chain) idents.push_back(id);
        }
        ldo_system_desc d {};
        d.n_types = static_cast<int>(sf.identities.size());
        d.type_len = type_len.data();
        d.idents = idents.data();
        d.cyclic = sf.cyclic ? 1 : 0;
        d.domain_type = domain_type;
        d.misbinding_pot = misbind;
        d.apply_mean_field_cor = p.m_apply_mean_field_cor ? 1 : 0;
        d.max_total_staples = std::max(p.m_max_total_staples, static_cast<int>(sf.chains.size()) - 1);
        // capacity follows the limits the moves can reach, clamped by what a scaffold can ever hold
        int hard_cap {0};
        for (size_t t {1}; t < sf.identities.size(); t++) hard_cap += p.m_max_type_staples;
        if (d.max_total_staples > hard_cap && hard_cap >= static_cast<int>(sf.chains.size()) - 1) d.max_total_staples = hard_cap;
        d.max_type_staples = p.m_max_type_staples;
        d.max_staple_size = p.m_max_staple_size;
        d.staple_M = p.m_staple_M;
        d.stacking_ene = p.m_stacking_ene;
        if (ldo_engine_create(&d, n_replicas, device, &s->eng) != 0) {
            throw std::runtime_error(std::string("engine: ") + ldo_last_error(nullptr));
        }
        // the move rules use the reference's limits, not the capacity clamp
        // (insertion refuses at num_staples == max_total_staples, met_movetypes.cpp:311-318)

        // tables
        int n_ident {s->tables[0].n_ident};
        size_t tsz {s->tables[0].hyb_energy.size()};
        std::vector<double> e, h, en, init;
        for (auto const& t: s->tables) {
            e.insert(e.end(), t.hyb_energy.begin(), t.hyb_energy.end());
            h.insert(h.end(), t.hyb_enthalpy.begin(), t.hyb_enthalpy.end());
            en.insert(en.end(), t.hyb_entropy.begin(), t.hyb_entropy.end());
            init.push_back(t.init_energy);
            init.push_back(t.init_enthalpy);
            init.push_back(t.init_entropy);
        }
        (void)tsz;
        s->check(ldo_set_temperature_tables(s->eng, static_cast<int>(s->temps.size()), n_ident, s->temps.data(), e.data(), h.data(), en.data(), init.data()));

        // moveset
        if (!p.m_movetype_filename.empty()) {
            s->movetypes = read_movetype_file(p.m_movetype_filename);
            std::vector<ldo_movetype_desc> md;
            for (auto const& m: s->movetypes) md.push_back(m.desc);
            s->check(ldo_set_moveset(s->eng, static_cast<int>(md.size()), md.data(), p.m_allow_nonsensical_ps ? 1 : 0));
            // the breakdowns of the .moves summary (write_move_summary) need the typed trackers
            if (!p.m_output_filebase.empty() && !s->is_enum) s->check(ldo_enable_move_trackers(s->eng, 1));
        }

        // order parameters and biases
        s->check(ldo_set_domain_update_biases(s->eng, p.m_domain_update_biases_present ? 1 : 0));
        if (!p.m_ops_filename.empty()) {
            s->ops = read_order_params_file(p.m_ops_filename);
            std::vector<ldo_order_param_desc> od(s->ops.size());
            for (size_t i {0}; i != s->ops.size(); i++) {
                od[i].type = op_type_code(s->ops[i].type);
                od[i].staple = s->ops[i].staple;
                od[i].chain1 = s->ops[i].chain1;
                od[i].domain1 = s->ops[i].domain1;
                od[i].chain2 = s->ops[i].chain2;
                od[i].domain2 = s->ops[i].domain2;
                od[i].n_sum = static_cast<int>(s->ops[i].sum_ops.size());
                od[i].sum_ops = s->ops[i].sum_ops.data();
                // per-domain updates happen in OrigamiSystemWithBias only, which setup_origami builds when
                // domain_update_biases_present is set (origami_system.cpp:1006-1028); with a plain OrigamiSystem such
                // parameters keep their initial value for ever - the device does the same through this flag
                od[i].update_per_domain = s->ops[i].update_per_domain ? 1 : 0;
            }
            s->check(ldo_set_order_params(s->eng, static_cast<int>(od.size()), od.data()));
        }
        if (!p.m_bias_funcs_filename.empty()) {
            s->biases = read_bias_functions_file(p.m_bias_funcs_filename, s->ops);
            std::vector<ldo_bias_desc> bd(s->biases.size());
            for (size_t i {0}; i != s->biases.size(); i++) {
                BiasSpec const& b = s->biases[i];
                bd[i].type = bias_type_code(b.type);
                bd[i].n_ops = static_cast<int>(b.ops.size());
                bd[i].ops = b.ops.data();
                bd[i].min_op = b.min_op;
                bd[i].max_op = b.max_op;
                bd[i].well_bias = b.well_bias;
                bd[i].min_bias = b.min_bias;
                bd[i].slope = b.slope;
                bd[i].outside_bias = b.outside_bias;
            }
            s->check(ldo_set_biases(s->eng, static_cast<int>(bd.size()), bd.data()));
        }

        if (s->is_us) us_setup(*s);

        // control variables
        std::vector<int> ti(n_replicas, 0);
        std::vector<double> um(n_replicas, p.m_staple_u_mult), bm(n_replicas, p.m_bias_funcs_mult), sm(n_replicas, 1.0);
        if (s->is_pt) {
            std::vector<double> cm {p.m_chem_pot_mults}, bmm {p.m_bias_mults}, smm {p.m_stacking_mults};
            cm.resize(s->num_reps, 1.0);
            bmm.resize(s->num_reps, 1.0);
            smm.resize(s->num_reps, 1.0);
            std::vector<int> ladder_ti(s->num_reps);
            for (int k {0}; k != s->num_reps; k++) ladder_ti[k] = k;
            if (s->pt_variant == LDO_PT_2D) {
                // initialize_control_qs (ptmc_simulation.cpp:454-493)
                for (int k {0}; k != s->num_reps; k++) {
                    ladder_ti[k] = k / s->v2_dim;
                    cm[k] = p.m_staple_u_mult;
                    bmm[k] = 1;
                    smm[k] = p.m_stacking_mults[k % s->v2_dim];
                }
            }
            // (the ladder keeps stacking_mults for every variant: calc_acceptance_p reads m_control_qs[3] whether
            // or not the variant applies the multiplier to the system, ptmc_simulation.cpp:283-284,307)
            s->check(ldo_set_exchange_ladder(s->eng, s->num_reps, ladder_ti.data(), cm.data(), bmm.data(), smm.data()));
            int slots_per_rank {s->num_reps / s->n_ranks};
            for (int r {0}; r != n_replicas; r++) {
                int k {ladder_slot_of(*s, r % slots_per_rank)};
                ti[r] = ladder_ti[k];
                // What every replica holds BEFORE the first exchange (initialize_control_qs, ptmc_simulation.cpp:337-357 and
                // :454-493): the bias multiplier is stored at index m_bias_i = 1, which is the staple-u-multiplier slot
                // (App. A17), so a replica starts with staple_u_mult = bias_mults[rank] (1-D) / 1 (2-D) and, for
                // hut_parallel_tempering, would start with bias multiplier 0 - had update_bias_mult any effect (below).
                // master_send overwrites the exchanged quantities with the ladder values after the first exchange
                // (m_exchange_q_is, :602-647); quantities a variant does not exchange keep these initial values.
                um[r] = s->pt_variant == LDO_PT_2D ? 1.0 : bmm[k];
                // No variant ever changes the bias multiplier: hut_parallel_tempering calls
                // OrigamiSystem::update_bias_mult (:665-672), an empty virtual nothing overrides (origami_system.hpp:165),
                // so SystemBiases keeps the constructor's bias_funcs_mult and bias_mults only decorate the .swp header.
                bm[r] = p.m_bias_funcs_mult;
                // st_/2d_: update_temp(temp, stacking_mult); t_/ut_/hut_: update_temp(temp), multiplier 1 (:651-680)
                sm[r] = (s->pt_variant == LDO_PT_ST || s->pt_variant == LDO_PT_2D) ? smm[k] : 1.0;
            }
            s->q2r.resize(static_cast<size_t>(s->n_ladders) * s->num_reps);
            for (int l {0}; l != s->n_ladders; l++)
                for (int k {0}; k != s->num_reps; k++) s->q2r[static_cast<size_t>(l) * s->num_reps + k] = k;
            // 1-D: one counter per neighbour pair; 2-D: [direction][v1][v2] (m_attempt_count, ptmc_simulation.cpp:438-443)
            size_t n_counters {s->pt_variant == LDO_PT_2D ? 2 * static_cast<size_t>(s->num_reps) : static_cast<size_t>(s->num_reps - 1)};
            s->attempts.assign(static_cast<size_t>(s->n_ladders) * n_counters, 0);
            s->accepts.assign(static_cast<size_t>(s->n_ladders) * n_counters, 0);
        }
        s->check(ldo_set_control(s->eng, 0, n_replicas, ti.data(), um.data(), bm.data(), sm.data()));

        // starting configuration
        Chains chains {sf.chains};
        if (!p.m_restart_traj_file.empty()) chains = read_trj_config(p.m_restart_traj_file, p.m_restart_step);
        std::vector<int> ci, cid, cl, pos, ore;
        for (auto const& c: chains) {
            ci.push_back(c.index);
            cid.push_back(c.identity);
            cl.push_back(static_cast<int>(c.positions.size() / 3));
            pos.insert(pos.end(), c.positions.begin(), c.positions.end());
            ore.insert(ore.end(), c.orientations.begin(), c.orientations.end());
        }
        s->check(ldo_set_state(s->eng, -1, static_cast<int>(chains.size()), ci.data(), cid.data(), cl.data(), pos.data(), ore.data()));
        if (s->is_pt && p.m_restart_from_config) {
            // PTGCMCSimulation constructor (ptmc_simulation.cpp:38-46): replica `rank` restarts from frame restart_step of
            // restart_traj_filebase-<rank><restart_traj_postfix>; every ladder of the batch starts from the same files
            int slots_per_rank {s->num_reps / s->n_ranks};
            for (int b {0}; b != slots_per_rank; b++) {
                int k {ladder_slot_of(*s, b)};
                Chains rc {read_trj_config(p.m_restart_traj_filebase + "-" + std::to_string(k) + p.m_restart_traj_postfix, p.m_restart_step)};
                std::vector<int> rci, rcid, rcl, rpos, rore;
                for (auto const& c: rc) {
                    rci.push_back(c.index);
                    rcid.push_back(c.identity);
                    rcl.push_back(static_cast<int>(c.positions.size() / 3));
                    rpos.insert(rpos.end(), c.positions.begin(), c.positions.end());
                    rore.insert(rore.end(), c.orientations.begin(), c.orientations.end());
                }
                for (int l {0}; l != s->n_ladders; l++) {
                    s->check(ldo_set_state(s->eng, l * slots_per_rank + b, static_cast<int>(rc.size()), rci.data(), rcid.data(), rcl.data(),
                                           rpos.data(), rore.data()));
                }
            }
        }
        if (s->is_pt && p.m_restart_from_swap) {
            // m_q_to_repi from the last row of restart_swap_file (ptmc_simulation.cpp:61-83). The replicas themselves
            // start with the control variables of their RANK, as in the reference (initialize_control_qs does not look
            // at the restored map); the first exchange puts them where the map says.
            std::ifstream swp {p.m_restart_swap_file};
            if (!swp) throw FileError {"Restart swap file " + p.m_restart_swap_file + " does not exist"};
            std::string line, last;
            while (std::getline(swp, line)) last = line;
            std::istringstream ls {last};
            for (int k {0}; k != s->num_reps; k++) {
                int repi {0};
                ls >> repi;
                for (int l {0}; l != s->n_ladders; l++) s->q2r[static_cast<size_t>(l) * s->num_reps + k] = repi;
            }
        }
        std::vector<int> status(n_replicas), detail(n_replicas);
        s->check(ldo_get_status(s->eng, status.data(), detail.data()));
        if (status[0] != 0) {
            throw OrigamiMisuse {"Constaints in violation after move complete (status " + std::to_string(status[0]) + ")"};
        }
        if (s->is_mwus && p.m_restart_from_config) {
            // MWUSGCMCSimulation::setup_window_variables + USGCMCSimulation constructor (us_simulation.cpp:531-549, 74-81,
            // 246-250): window w continues from frame restart_step of restart_traj_filebase<window postfix><restart_traj_postfix>
            // (or restart_traj_files[w] / restart_steps[w]) through set_config, which leaves the stored order parameters
            // and biases of the system-file configuration in place until the first move
            for (int r {0}; r != s->R; r++) {
                int w {r % s->n_windows};
                std::string file {p.m_restart_traj_filebase + s->window_postfix[w] + p.m_restart_traj_postfix};
                int step {p.m_restart_step};
                if (!p.m_restart_traj_files.empty()) {
                    if (static_cast<int>(p.m_restart_traj_files.size()) != s->n_windows || static_cast<int>(p.m_restart_steps.size()) != s->n_windows) {
                        throw SimulationMisuse {"restart_traj_files / restart_steps must list one entry per window"};
                    }
                    file = p.m_restart_traj_files[w];
                    step = p.m_restart_steps[w];
                }
                Chains rc {read_trj_config(file, step)};
                std::vector<int> rci, rcid, rcl, rpos, rore;
                for (auto const& c: rc) {
                    rci.push_back(c.index);
                    rcid.push_back(c.identity);
                    rcl.push_back(static_cast<int>(c.positions.size() / 3));
                    rpos.insert(rpos.end(), c.positions.begin(), c.positions.end());
                    rore.insert(rore.end(), c.orientations.begin(), c.orientations.end());
                }
                s->check(ldo_replace_config(s->eng, r, static_cast<int>(rc.size()), rci.data(), rcid.data(), rcl.data(), rpos.data(), rore.data()));
            }
        }
        if (s->is_us) us_apply_window_limits(*s);

        // seed (simulation.cpp:199-203; random_gens.cpp:12-19 when unspecified)
        unsigned long long seed;
        if (p.m_random_seed != -1) {
            seed = static_cast<unsigned long long>(p.m_random_seed);
            // the reference reports the seed on stdout (simulation.cpp:203); LDO_QUIET keeps stdout clean for
            // callers that print machine-readable output themselves (bench.py)
            if (!std::getenv("LDO_QUIET")) std::cout << "Using specified seed: " << p.m_random_seed << "\n";
        }
        else {
            // every rank must key the shared exchange stream identically (k_exchange draws from (seed, swap,
            // ladder, pair)): a per-rank random_device seed would make the ranks disagree on the swaps
            if (s->n_ranks > 1 && s->is_pt) throw SimulationMisuse {"random_seed must be set when replica exchange runs on more than one rank"};
            std::random_device rd {};
            seed = (static_cast<unsigned long long>(rd()) << 32) ^ rd();
            if (!std::getenv("LDO_QUIET")) std::cout << "Truly random seed: " << seed << "\n";
        }
        if (s->is_pt) {
            // replica k of ladder l keeps stream l * num_reps + k whichever rank holds it
            int slots_per_rank {s->num_reps / s->n_ranks};
            std::vector<unsigned int> sub(n_replicas);
            for (int r {0}; r != n_replicas; r++) {
                int l {r / slots_per_rank}, k {ladder_slot_of(*s, r % slots_per_rank)};
                sub[r] = static_cast<unsigned int>(l * s->num_reps + k);
            }
            s->check(ldo_seed_subsequences(s->eng, seed, sub.data()));
        }
        else {
            s->check(ldo_seed(s->eng, seed, static_cast<unsigned int>(s->rank * n_replicas)));
        }
        if (p.m_random_seed == -1 && p.m_read_rand_engine_state) restore_rng_states(*s);
        s->start = std::chrono::steady_clock::now();
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return nullptr;
    }
    return s.release();
}

void ldo_sim_destroy(ldo_sim* s) {
    g_us_states.erase(s);
    delete s;
}
ldo_engine* ldo_sim_engine(ldo_sim* s) { return s->eng; }

int ldo_sim_exchange_advance(ldo_sim* s) {
    try {
        if (!s->is_pt) throw SimulationMisuse {"not a replica-exchange simulation"};
        if (!simulate(*s, s->params.m_exchange_interval)) return 1;
        s->check(ldo_exchange_collect(s->eng, nullptr));
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_comm_unique_id(void* id_out) {
    try {
        NcclId id {};
        nccl_check(nccl_or_throw().GetUniqueId(&id), "unique id");
        std::memcpy(id_out, &id, sizeof(id));
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_sim_comm_init(ldo_sim* s, const void* unique_id) {
    try {
#ifdef LDO_HOSTSIM
        (void)unique_id;
        throw std::runtime_error("ldo_sim_comm_init: the NCCL path needs the CUDA build");
#else
        if (s->n_ranks < 2) return 0;
        NcclId id {};
        std::memcpy(&id, unique_id, sizeof(id));
        nccl_check(nccl_or_throw().CommInitRank(&s->nccl_comm, s->n_ranks, id, s->rank), "communicator");
        s->check(ldo_exchange_buffers(s->eng, s->R * s->n_ranks, &s->dep_send, &s->dep_recv, &s->dep_nq));
#endif
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_sim_exchange_round(ldo_sim* s, long long swap_i) {
    try {
        if (!s->is_pt) throw SimulationMisuse {"not a replica-exchange simulation"};
        if (!exchange_round(*s, swap_i)) return 1;
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_sim_exchange_apply(ldo_sim* s, long long swap_i, const double* dependent_all) {
    try {
        exchange_download(*s);
        s->exchange_resident = false;
        if (s->pt_variant == LDO_PT_2D) {
            s->check(ldo_exchange_pt_2d(
                    s->eng, swap_i, s->n_ladders, s->v1_dim, s->v2_dim, s->rank, s->n_ranks,
                    dependent_all, s->q2r.data(), s->attempts.data(), s->accepts.data()));
        }
        else {
            s->check(ldo_exchange_pt(
                    s->eng, s->pt_variant, swap_i, s->n_ladders, s->num_reps, s->rank, s->n_ranks,
                    dependent_all, s->q2r.data(), s->attempts.data(), s->accepts.data()));
        }
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_sim_exchange_state(ldo_sim* s, int* slot_to_replica, long long* attempts, long long* accepts) {
    try {
        exchange_download(*s);
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    if (slot_to_replica) std::memcpy(slot_to_replica, s->q2r.data(), sizeof(int) * s->q2r.size());
    if (attempts) std::memcpy(attempts, s->attempts.data(), sizeof(long long) * s->attempts.size());
    if (accepts) std::memcpy(accepts, s->accepts.data(), sizeof(long long) * s->accepts.size());
    return 0;
}

int ldo_sim_run(ldo_sim* s) {
    try {
        InputParameters const& p = s->params;
        if (!s->is_us && !s->is_enum) open_output_files(*s);
        s->start = std::chrono::steady_clock::now();
        std::string const& st = p.m_simulation_type;
        if (s->is_enum) {
            enumerate_run(*s);
            return 0; // no move statistics
        }
        else if (st == "constant_temp") {
            simulate(*s, p.m_ct_steps);
        }
        else if (st == "annealing") {
            // AnnealingGCMCSimulation::run (annealing_simulation.cpp:38-49)
            // `step += simulate(m_steps_per_temp, step)`: simulate returns the last step number plus one
            // (simulation.cpp:574,652), so the step numbers of the output files jump between temperatures:
            // 1..n, 2(n+1)+1.., ... Reproduced, including for the centring / constraint-check frequencies.
            for (size_t i {0}; i != s->temps.size(); i++) {
                set_all_control(*s, static_cast<int>(i));
                long long start {s->step};
                if (!simulate(*s, p.m_steps_per_temp)) break;
                s->step = start + (start + p.m_steps_per_temp + 1);
                s->check(ldo_set_step(s->eng, s->step));
            }
        }
        else if (s->is_us) {
            us_run(*s);
        }
        else if (s->is_pt) {
            if (s->n_ranks != 1 && !s->nccl_comm) throw SimulationMisuse {"replica exchange on several ranks: call ldo_sim_comm_init first"};
            // PTGCMCSimulation::run (ptmc_simulation.cpp:106-150); .swp as :92-104, :315-322 (written by rank 0)
            bool master {s->rank == 0};
            std::ofstream swp;
            if (master && !p.m_output_filebase.empty()) {
                swp.open(p.m_output_filebase + ".swp");
                std::vector<double> cm {p.m_chem_pot_mults}, bmm {p.m_bias_mults}, smm {p.m_stacking_mults};
                cm.resize(s->num_reps, 1.0);
                bmm.resize(s->num_reps, 1.0);
                smm.resize(s->num_reps, 1.0);
                for (int k {0}; k != s->num_reps; k++) {
                    if (s->pt_variant == LDO_PT_2D) {
                        // m_exchange_q_is = {temp, staple_u_mult, stacking_mult} (ptmc_simulation.cpp:446-448)
                        swp << p.m_temps[k / s->v2_dim] << "/" << p.m_staple_u_mult << "/" << p.m_stacking_mults[k % s->v2_dim] << "/ ";
                        continue;
                    }
                    swp << p.m_temps[k] << "/";
                    if (s->pt_variant == LDO_PT_UT || s->pt_variant == LDO_PT_HUT) swp << cm[k] << "/";
                    if (s->pt_variant == LDO_PT_HUT) swp << bmm[k] << "/";
                    if (s->pt_variant == LDO_PT_ST) swp << smm[k] << "/";
                    swp << " ";
                }
                swp << "\n";
            }
            auto write_swap_entry = [&](long long step) {
                if (!swp.is_open() || p.m_configs_output_freq == 0) return;
                if (step % p.m_configs_output_freq == 0) {
                    exchange_download(*s); // the map lives on the device between the rows that are written
                    for (int k {0}; k != s->num_reps; k++) swp << s->q2r[k] << " ";
                    swp << "\n";
                }
            };
            for (long long swap_i {1}; swap_i != p.m_swaps + 1; swap_i++) {
                // the round's moves, then the row of the map as it stood during them, then the exchange
                // (ptmc_simulation.cpp:113-141); the decisions are only enqueued here, so the row is read first
                InputParameters const& pp = s->params;
                long long end {s->step + pp.m_exchange_interval};
                double dt {std::chrono::duration<double>(std::chrono::steady_clock::now() - s->start).count()};
                // (the wall-clock limit is a per-process decision: with several ranks it would need the master's kill
                // message of ptmc_simulation.cpp:120-127; it is honoured on a single rank only)
                if (s->n_ranks == 1 && dt > p.m_max_pt_dur) {
                    std::cout << "Maximum time allowed reached\n";
                    break;
                }
                if (master && p.m_configs_output_freq != 0 && end % p.m_configs_output_freq == 0) write_swap_entry(end);
                if (!exchange_round(*s, swap_i)) {
                    if (master) std::cout << "Maximum time allowed reached\n";
                    break;
                }
            }
            s->check(ldo_synchronize(s->eng));
            exchange_download(*s);
            write_swap_entry(s->step);
            if (!master) {
                write_move_summary(*s);
                return 0;
            }
            if (s->pt_variant == LDO_PT_2D) {
                // TwoDPTGCMCSimulation::write_acceptance_freqs (ptmc_simulation.cpp:562-593), first ladder
                int v1 {s->v1_dim}, v2 {s->v2_dim};
                for (int a {0}; a != v1 - 1; a++)
                    for (int b {0}; b != v2; b++) {
                        long long sw {s->accepts[a * v2 + b]}, at {s->attempts[a * v2 + b]};
                        std::cout << p.m_temps[a] << " " << p.m_temps[a + 1] << " " << p.m_stacking_mults[b] << " " << sw << " " << at
                                  << " " << static_cast<double>(sw) / at << " \n";
                    }
                std::cout << "\n";
                for (int b {0}; b != v2 - 1; b++)
                    for (int a {0}; a != v1; a++) {
                        long long sw {s->accepts[v1 * v2 + a * v2 + b]}, at {s->attempts[v1 * v2 + a * v2 + b]};
                        std::cout << p.m_stacking_mults[b] << " " << p.m_stacking_mults[b + 1] << " " << p.m_temps[a] << " " << sw << " "
                                  << at << " " << static_cast<double>(sw) / at << " \n";
                    }
                std::cout << "\n";
            }
            // write_acceptance_freqs (ptmc_simulation.cpp:414-426), first ladder
            for (int i {0}; s->pt_variant != LDO_PT_2D && i + 1 < s->num_reps; i++) {
                std::cout << p.m_temps[i] << " " << p.m_temps[i + 1] << " " << s->accepts[i] << " " << s->attempts[i] << " "
                          << static_cast<double>(s->accepts[i]) / s->attempts[i] << " \n";
            }
            std::cout << "\n";
        }
        else {
            throw NotImplemented {st + ": simulation type not available on the device path"};
        }
        write_move_summary(*s);
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_sim_num_temps(ldo_sim* s) { return static_cast<int>(s->temps.size()); }
int ldo_sim_num_order_params(ldo_sim* s) { return static_cast<int>(s->ops.size()); }
const char* ldo_sim_order_param_tag(ldo_sim* s, int i) { return s->ops[i].tag.c_str(); }
int ldo_sim_num_movetypes(ldo_sim* s) { return static_cast<int>(s->movetypes.size()); }
const char* ldo_sim_movetype_label(ldo_sim* s, int i) { return s->movetypes[i].label.c_str(); }
int ldo_sim_num_staple_types(ldo_sim* s) { return static_cast<int>(s->sysfile->identities.size()) - 1; }
int ldo_sim_enumeration_summary(ldo_sim* s, double* out) {
    out[0] = s->enum_num_configs;
    out[1] = s->enum_average_energy;
    out[2] = s->enum_average_bias;
    out[3] = static_cast<double>(s->enum_leaves);
    return s->is_enum ? 0 : -1;
}
long long ldo_sim_step(ldo_sim* s) { return s->step; }

int ldo_sim_pair_energies(ldo_sim* s, int temp_idx, int a, int b, double* out) {
    EnergyTables const& t = s->tables[temp_idx];
    if (std::abs(a) > t.n_ident || std::abs(b) > t.n_ident) return -1;
    size_t k {t.index(a, b)};
    if (!t.present[k]) return -1;
    out[0] = t.hyb_energy[k];
    out[1] = t.hyb_enthalpy[k];
    out[2] = t.hyb_entropy[k];
    return 0;
}

int ldo_sim_init_energies(ldo_sim* s, int temp_idx, double* out) {
    EnergyTables const& t = s->tables[temp_idx];
    out[0] = t.init_energy;
    out[1] = t.init_enthalpy;
    out[2] = t.init_entropy;
    return 0;
}

int ldo_host_energy_tables(const char* inp_path, double temp, int* n_ident, double* hyb_energy, double* hyb_enthalpy,
                           double* hyb_entropy, char* present, double* init) {
    try {
        InputParameters p {inp_path};
        OrigamiInputFile sf {p.m_origami_input_filename};
        EnergyTables t {calc_energy_tables(sf, p, temp)};
        *n_ident = t.n_ident;
        size_t n {t.hyb_energy.size()};
        if (hyb_energy) std::memcpy(hyb_energy, t.hyb_energy.data(), sizeof(double) * n);
        if (hyb_enthalpy) std::memcpy(hyb_enthalpy, t.hyb_enthalpy.data(), sizeof(double) * n);
        if (hyb_entropy) std::memcpy(hyb_entropy, t.hyb_entropy.data(), sizeof(double) * n);
        if (present) std::memcpy(present, t.present.data(), n);
        if (init) {
            init[0] = t.init_energy;
            init[1] = t.init_enthalpy;
            init[2] = t.init_entropy;
        }
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_host_inp_value(const char* inp_path, const char* key, char* out, int outlen) {
    try {
        InputParameters p {inp_path};
        std::ostringstream os;
        os.precision(17);
        std::string k {key};
        auto join = [&](auto const& v) {
            for (size_t i {0}; i != v.size(); i++) os << (i ? " " : "") << v[i];
        };
        if (k == "origami_input_filename") os << p.m_origami_input_filename;
        else if (k == "domain_type") os << p.m_domain_type;
        else if (k == "temp") os << p.m_temp;
        else if (k == "staple_M") os << p.m_staple_M;
        else if (k == "cation_M") os << p.m_cation_M;
        else if (k == "stacking_ene") os << p.m_stacking_ene;
        else if (k == "max_total_staples") os << p.m_max_total_staples;
        else if (k == "max_type_staples") os << p.m_max_type_staples;
        else if (k == "max_staple_size") os << p.m_max_staple_size;
        else if (k == "simulation_type") os << p.m_simulation_type;
        else if (k == "random_seed") os << p.m_random_seed;
        else if (k == "centering_freq") os << p.m_centering_freq;
        else if (k == "constraint_check_freq") os << p.m_constraint_check_freq;
        else if (k == "max_duration") os << p.m_max_duration;
        else if (k == "ct_steps") os << p.m_ct_steps;
        else if (k == "temps") join(p.m_temps);
        else if (k == "chem_pot_mults") join(p.m_chem_pot_mults);
        else if (k == "bias_mults") join(p.m_bias_mults);
        else if (k == "stacking_mults") join(p.m_stacking_mults);
        else if (k == "num_reps") os << p.m_num_reps;
        else if (k == "exchange_interval") os << p.m_exchange_interval;
        else if (k == "swaps") os << p.m_swaps;
        else if (k == "ops_to_output") join(p.m_ops_to_output);
        else if (k == "apply_mean_field_cor") os << (p.m_apply_mean_field_cor ? "true" : "false");
        else if (k == "vcf_per_domain") os << (p.m_vcf_per_domain ? "true" : "false");
        else if (k == "restart_traj_postfix") os << p.m_restart_traj_postfix;
        else if (k == "max_rel_P_diff") os << p.m_max_rel_P_diff;
        else if (k == "output_filebase") os << p.m_output_filebase;
        else throw FileError {"ldo_host_inp_value: key not exposed: " + k};
        std::strncpy(out, os.str().c_str(), outlen - 1);
        out[outlen - 1] = 0;
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_host_nn_unitless_thermo(const char* seq, double temp, double cation_M, double* out) {
    try {
        ThermoOfHybrid t {calc_unitless_hybridization_thermo(seq, temp, cation_M)};
        out[0] = t.enthalpy;
        out[1] = t.entropy;
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
    return 0;
}

int ldo_host_longest_contig_complement(const char* a, const char* b, char* out, int outlen) {
    try {
        auto v = find_longest_contig_complement(a, b);
        std::string s;
        for (auto const& x: v) s += x + "\n";
        std::strncpy(out, s.c_str(), outlen - 1);
        out[outlen - 1] = 0;
        return static_cast<int>(v.size());
    } catch (std::exception const& e) {
        g_host_error = e.what();
        return -1;
    }
}

int ldo_host_no_walks(const int* start, const int* end, int steps) {
    int dr {std::abs(end[0] - start[0]) + std::abs(end[1] - start[1]) + std::abs(end[2] - start[2])};
    return (dr > steps || (steps - dr) % 2 != 0) ? 1 : 0;
}

} // extern "C"
