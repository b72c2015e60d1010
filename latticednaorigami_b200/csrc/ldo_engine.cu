// ldo_engine.cu — kernels and C-ABI (include/ldo_b200.h) of the replica-batched Monte Carlo engine.
//
// Data layout in HBM (per engine = per GPU):
//   states[R]   SysState<K>   persistent configuration of every replica (array of structs; one
//                             replica is a contiguous, 16-byte aligned blob so that a warp stages it
//                             to / from shared memory with coalesced 128-bit loads and stores)
//   aux[R]      RepAux        RNG state, control variables, bias state, move statistics, step counter
//   tables      TempTables[n_temps] + fp64 hybridization tables (shared by all replicas at a temperature)
// One warp owns one replica for the whole launch: the state lives in shared memory while the warp
// runs its `n_steps` moves (small systems), or stays in HBM/L2 and is accessed in place (large
// systems whose state + move scratch exceed the per-warp shared-memory budget).
//
// With LDO_HOSTSIM defined (tests/hostsim only) the same per-replica functions are driven by a plain
// host loop with one emulated lane, for sanitizer/debug runs of the device logic on a machine without
// a GPU. That build is never part of the shipped library: the product path is CUDA only.

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "ldo_moves.cuh"
#include "ldo_enum.cuh"
#include "../../include/ldo_b200.h"

using namespace ldo;
#ifdef LDO_HOSTSIM
extern "C" long long ldo_dbg_counts[16] = {0};
#if defined(LDO_CALLER_PROFILE)
extern "C" unsigned long long ldo_dbg_callers[2 * 4096] = {0};
#endif
#endif

// ---------------------------------------------------------------------------------------------
// Backend
// ---------------------------------------------------------------------------------------------
#ifdef LDO_HOSTSIM
#define LDO_GLOBAL
typedef int dev_stream_t;
static int dev_malloc(void** p, size_t n) {
    *p = calloc(1, n ? n : 1);
    return *p ? 0 : -1;
}
static void dev_free(void* p) { free(p); }
static int dev_h2d(void* d, const void* h, size_t n, dev_stream_t) {
    memcpy(d, h, n);
    return 0;
}
static int dev_d2h(void* h, const void* d, size_t n, dev_stream_t) {
    memcpy(h, d, n);
    return 0;
}
static int dev_memset(void* d, int v, size_t n, dev_stream_t) {
    memset(d, v, n);
    return 0;
}
static int dev_sync(dev_stream_t) { return 0; }
static const char* dev_err() { return "hostsim"; }
#else
#include <cuda_runtime.h>
#define LDO_GLOBAL __global__
typedef cudaStream_t dev_stream_t;
static thread_local cudaError_t g_last_cuda = cudaSuccess;
static int chk(cudaError_t e) {
    if (e != cudaSuccess) {
        g_last_cuda = e;
        return -1;
    }
    return 0;
}
static int dev_malloc(void** p, size_t n) {
    if (chk(cudaMalloc(p, n ? n : 1))) return -1;
    // the engine's stream is non-blocking with respect to the legacy stream this memset runs on:
    // wait for it, or a later copy on the engine stream could be overwritten by the zero fill
    if (chk(cudaMemset(*p, 0, n ? n : 1))) return -1;
    return chk(cudaDeviceSynchronize());
}
static void dev_free(void* p) { cudaFree(p); }
static int dev_h2d(void* d, const void* h, size_t n, dev_stream_t s) {
    if (chk(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s))) return -1;
    return chk(cudaStreamSynchronize(s));
}
static int dev_d2h(void* h, const void* d, size_t n, dev_stream_t s) {
    if (chk(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s))) return -1;
    return chk(cudaStreamSynchronize(s));
}
static int dev_memset(void* d, int v, size_t n, dev_stream_t s) { return chk(cudaMemsetAsync(d, v, n, s)); }
static int dev_sync(dev_stream_t s) { return chk(cudaStreamSynchronize(s)); }
static const char* dev_err() { return cudaGetErrorString(g_last_cuda); }
#endif

// ---------------------------------------------------------------------------------------------
// Per-replica auxiliary state and shared (read-only) engine constants
// ---------------------------------------------------------------------------------------------

#define LDO_MAX_LISTS 256 // work lists of the staged kernel (one per SM, indexed by %smid modulo the count)
#define LDO_GRID_CAP 1024 // grid-bias points per replica (sum over grid biases)
// Exchange record of a replica (what slave_send ships, ptmc_simulation.cpp:163-175): enthalpy, bias, stacking, the
// replica's own staple chemical-potential multiplier (the reference ships m_staple_us = reduced_u * T * multiplier,
// origami_system.cpp:624-628; T is the slot's), then the staple counts per identity
#define LDO_DEP_FIXED 4

struct __attribute__((aligned(16))) RepAux {
    Rng rng;
    Control ctl;
    // OrigamiPotential::update_temp keeps one table set per (temperature, stacking multiplier) it has seen and, on a
    // cache hit, restores the hybridization and stacking tables but NOT m_init_enthalpy / m_init_entropy / m_init_energy
    // (origami_potential.cpp:1019-1042 vs :1060-1063): a replica that returns to a temperature it has visited keeps the
    // initiation terms of the last temperature it saw for the first time. Replica exchange revisits temperatures all the
    // time, so the quirk is reproduced: `init_temp_idx` names the table whose initiation terms are current (-1: those of
    // ctl.temp_idx), `table_cache_mask` has one bit per ladder slot key already seen (exchange_ladder).
    int init_temp_idx;
    unsigned long long table_cache_mask;
    int window_swapped; // set by the window exchange for the two replicas of an accepted swap (OP_AFTER_WINDOW_SWAP)
    BiasState bs;
    long long step;
    // not staged to shared memory (touched twice per move): everything from here on stays in HBM/L2
    __attribute__((aligned(16))) MoveStats stats;
};
#define LDO_AUX_HOT_BYTES offsetof(RepAux, stats)

struct Shared {
    SysConst sc;
    MoveSet ms;
    OpsBiasConst ob;
    int n_temps;
    int has_grid;
};

enum {
    OP_RUN = 0,
    OP_LOAD_CONFIG = 1,
    OP_UPDATE_ENERGY = 2,
    OP_CHECK_CONSTRAINTS = 3,
    OP_CENTER = 4,
    OP_OBSERVE = 5,
    OP_RECOMPUTE = 6,
    OP_REFRESH_OPS = 7,
    OP_REINIT_BIASES = 8,
    OP_AFTER_WINDOW_SWAP = 9
};

struct OpArgs {
    int op;
    int n_replicas;
    int only_replica; // -1 = all
    long long n_steps;
    int centering_freq, centering_domain, constraint_check_freq;
    // LOAD_CONFIG staging: one configuration applied to the selected replicas
    int cfg_n_chains;
    int cfg_keep_bias; // OrigamiSystem::set_config on a live system: stored order parameters / biases stay (ldo_replace_config)
    const int* cfg_chain_index;
    const int* cfg_chain_ident;
    const int* cfg_chain_len;
    const int* cfg_pos;
    const int* cfg_ore;
    // OBSERVE / RECOMPUTE outputs
    double* out_energies; // [R][5]
    int* out_counters; // [R][9]
    int* out_ops; // [R][n_ops]
    int* out_staples; // [R][n_types - 1]
    double* out_dependent; // [R][LDO_DEP_FIXED + n_types - 1] (exchange quantities)
    int* out_status; // [R][2]
    double* out_recomputed; // [R]
    int* out_recomputed_stacked; // [R]
};

template <class K>
struct DevPtrs {
    SysState<K>* states;
    MoveScratch<K>* scratch; // hot move scratch, only for the in-place (non-staged) path
    ColdScratch<K>* cold; // [R] selection / topology scratch, always in global memory
    Engine<K>* engines; // per-replica engine objects, only for the in-place path
    RepAux* aux;
    const Shared* shared;
    const TempTables* tables;
    double* grid_vals; // [R][LDO_GRID_CAP]
    long long* grid_visits; // [R][LDO_GRID_CAP]
    long long* run_timing; // [R][3] diagnostics of the last run launch: start ns, end ns (globaltimer), SM id
    int* queue; // [LDO_MAX_LISTS] work-queue heads of the staged kernel (reset before every launch)
    int* order; // [R] replicas in the order they are handed out by a run launch (most expensive first)
    TrackStats* track; // [R] typed move trackers, null unless enabled (ldo_enable_move_trackers)
};

// The device pointers seen by the Tracked<K> instantiation of the kernels (same layout, ldo_core.cuh)
template <class K>
inline DevPtrs<Tracked<K>> tracked_ptrs(const DevPtrs<K>& p) {
    static_assert(sizeof(DevPtrs<Tracked<K>>) == sizeof(DevPtrs<K>) && sizeof(SysState<Tracked<K>>) == sizeof(SysState<K>) &&
                  sizeof(MoveScratch<Tracked<K>>) == sizeof(MoveScratch<K>) && sizeof(ColdScratch<Tracked<K>>) == sizeof(ColdScratch<K>) &&
                  sizeof(Engine<Tracked<K>>) == sizeof(Engine<K>), "Tracked<K> must not change any layout");
    DevPtrs<Tracked<K>> q;
    memcpy(&q, &p, sizeof(q));
    return q;
}

// ---------------------------------------------------------------------------------------------
// Per-replica operations (host+device; executed warp-uniformly)
// ---------------------------------------------------------------------------------------------

template <class K>
LDO_HD void rep_init_engine(Engine<K>& eng, SysState<K>* st, MoveScratch<K>* ms, RepAux* aux, const DevPtrs<K>& P, int r) {
    const Shared* sh = P.shared;
    eng.sys.init(st, &sh->sc, P.tables[aux->ctl.temp_idx]);
    eng.sys.serial_terms = (aux->rng.tape != nullptr || sh->ms.reference_draw_order != 0) ? 1 : 0;
    if (aux->init_temp_idx >= 0 && aux->init_temp_idx != aux->ctl.temp_idx) {
        const TempTables& it = P.tables[aux->init_temp_idx];
        eng.sys.tt.init_energy = it.init_energy;
        eng.sys.tt.init_enthalpy = it.init_enthalpy;
        eng.sys.tt.init_entropy = it.init_entropy;
    }
    eng.m = ms;
    eng.mc = &P.cold[r];
    eng.rng = &aux->rng;
    eng.ms = &sh->ms;
    eng.ob = &sh->ob;
    eng.bs = &aux->bs;
    eng.sys.bsp = &aux->bs;
    eng.sys.obp = &sh->ob;
    eng.grid_vals = P.grid_vals ? P.grid_vals + (size_t)aux->bs.grid_slot * LDO_GRID_CAP : nullptr;
    eng.ctl = aux->ctl;
    eng.stats = &P.aux[r].stats;
    eng.trk = P.track ? &P.track[r] : nullptr;
}

template <class K>
LDO_HD void rep_refresh_stack_energy(SysState<K>* st, const Shared* sh, const Control& ctl) {
    // OrigamiPotential::update_temp scales the Constant stacking energy (origami_potential.cpp:1026-1029,1219)
    (void)sh;
#if defined(__CUDA_ARCH__)
    st->stack_e = ldo_c_sc.stacking_ene * ctl.stacking_mult / ctl.temp;
#else
    st->stack_e = sh->sc.stacking_ene * ctl.stacking_mult / ctl.temp;
#endif
}

// Bias bookkeeping restart (SystemBiases constructor: every m_bias evaluated once, bias_functions.cpp:60-66,423-426)
template <class K>
LDO_HD void rep_init_biases(Engine<K>& eng) {
    // every parameter and every bias is evaluated once, whichever kind (order_params.cpp:31, bias_functions.cpp:414-426)
    eng.BS()->op_undefined = 0;
    for (int i = 0; i < eng.OB().n_ops; i++) eng.BS()->op_val[i] = eng.calc_op(i);
    eng.BS()->move_update_bias = 0;
    eng.BS()->pd_disabled = 0;
    for (int b = 0; b < eng.OB().n_biases; b++) {
        eng.BS()->bias_val[b] = eng.calc_bias_fn(b);
        eng.BS()->move_update_bias += eng.BS()->bias_val[b];
    }
}

// set_config (origami_system.cpp:327-341) + the constructor's initialisation (:41-71, 659-693)
template <class K>
LDO_HD void rep_load_config(Engine<K>& eng, const OpArgs& a) {
    System<K>& sys = eng.sys;
    SysState<K>* s = sys.S();
    const SysConst* sc = &sys.SC();
    s->status = LDO_OK;
    s->status_detail = 0;
    s->constraints_violated = 0;
    s->weight_pass = 0;
    eng.BS()->pd_disabled = 1; // ops and biases are (re)built from the loaded configuration below
    sys.table_clear();
    for (int c = 0; c < K::C; c++) s->chain_used[c] = 0;
    for (int t = 0; t < K::T; t++) s->type_count[t] = 0;
    s->n_chains = 0;
    s->num_staples = 0;
    s->num_domains = 0;
    s->num_bound_pairs = 0;
    s->num_fully_bound_pairs = 0;
    s->num_self_bound_pairs = 0;
    s->num_stacked_pairs = 0;
    s->num_unassigned = 0;
    s->energy = 0;
    // scaffold (initialize_scaffold, :680-693)
    {
        int len = sc->type_len[0];
        s->chain_used[0] = 1;
        s->chain_uid[0] = a.cfg_chain_index[0];
        s->chain_type[0] = 0;
        s->chain_len[0] = (uint16_t)len;
        s->order[0] = 0;
        s->n_chains = 1;
        s->type_count[0] = 1;
        for (int i = 0; i < len; i++) {
            s->dom[i].state = ST_UNASSIGNED;
            s->dom[i].link = chain_link_flags(i, len, sc->cyclic != 0);
            s->bound[i] = -1;
            s->ident[i] = sc->idents[sc->type_off[0] + i];
            s->dchain[i] = 0;
            s->dindex[i] = (uint16_t)i;
            s->num_domains++;
            s->num_unassigned++;
        }
    }
    // staples (initialize_staples, :659-678)
    int max_uid = a.cfg_chain_index[0];
    for (int i = 1; i < a.cfg_n_chains; i++) {
        sys.add_chain_with_uid(a.cfg_chain_ident[i], a.cfg_chain_index[i]);
        if (a.cfg_chain_index[i] > max_uid) max_uid = a.cfg_chain_index[i];
    }
    int ns = a.cfg_n_chains - 1;
    if (sc->apply_mean_field_cor) s->energy += ns * log(6.0);
    s->energy += ns * sys.TT().init_energy;
    s->current_c_i = max_uid;
    if (s->status != LDO_OK) return;
    // positions and orientations, then set_all_domains (:588-616)
    int k = 0;
    for (int w = 0; w < s->n_chains; w++) {
        int c = s->order[w];
        int base = sys.chain_base(c);
        for (int i = 0; i < s->chain_len[c]; i++) {
            if (abs(a.cfg_pos[3 * k]) > LDO_COORD_MAX_XY || abs(a.cfg_pos[3 * k + 1]) > LDO_COORD_MAX_XY ||
                abs(a.cfg_pos[3 * k + 2]) > LDO_COORD_MAX_Z) {
                sys.fail(LDO_ERR_COORD_RANGE, base + i);
                return;
            }
            V3 p = v3(a.cfg_pos[3 * k], a.cfg_pos[3 * k + 1], a.cfg_pos[3 * k + 2]);
            V3 o = v3(a.cfg_ore[3 * k], a.cfg_ore[3 * k + 1], a.cfg_ore[3 * k + 2]);
            int oc = ore_code(o);
            if (oc == 7) oc = ORE_ZERO;
            sys.write_dom(base + i, p, oc);
            k++;
        }
    }
    sys.set_all_domains();
    if (a.cfg_keep_bias) eng.BS()->pd_disabled = 0;
    else rep_init_biases(eng);
}

template <class K>
LDO_HD void rep_observe(Engine<K>& eng, const OpArgs& a, int r) {
    System<K>& sys = eng.sys;
    const SysState<K>* s = sys.S();
    int nst = sys.SC().n_types - 1;
    if (a.out_energies) {
        double H, S, stk;
        sys.enthalpy_and_entropy(&H, &S, &stk);
        double* o = a.out_energies + (size_t)r * 5;
        o[0] = s->energy;
        o[1] = H;
        o[2] = S;
        o[3] = stk;
        o[4] = eng.total_bias();
    }
    if (a.out_dependent) {
        // PTGCMCSimulation::update_dependent_qs (ptmc_simulation.cpp:152-161) + staple counts
        double H, S, stk;
        sys.enthalpy_and_entropy(&H, &S, &stk);
        double* o = a.out_dependent + (size_t)r * (LDO_DEP_FIXED + nst);
        o[0] = H;
        o[1] = eng.total_bias();
        o[2] = stk;
        o[3] = eng.CTL().staple_u_mult;
        for (int t = 0; t < nst; t++) o[LDO_DEP_FIXED + t] = (double)s->type_count[t + 1];
    }
    if (a.out_counters) {
        int* o = a.out_counters + (size_t)r * 9;
        o[0] = s->num_staples;
        o[1] = s->num_domains;
        o[2] = s->num_bound_pairs;
        o[3] = s->num_fully_bound_pairs;
        o[4] = s->num_self_bound_pairs;
        o[5] = s->num_bound_pairs - s->num_fully_bound_pairs;
        o[6] = s->num_stacked_pairs;
        o[7] = s->num_unassigned;
        o[8] = s->current_c_i;
    }
    if (a.out_ops) {
        // the stored values, as OrigamiOrderParamsOutputFile::write prints them (files.cpp:772-778): reading them must
        // not change what the next move sees
        for (int i = 0; i < eng.OB().n_ops; i++) a.out_ops[(size_t)r * eng.OB().n_ops + i] = eng.BS()->op_val[i];
    }
    if (a.out_staples) {
        for (int t = 0; t < nst; t++) a.out_staples[(size_t)r * nst + t] = s->type_count[t + 1];
    }
    if (a.out_status) {
        a.out_status[2 * (size_t)r] = s->status;
        a.out_status[2 * (size_t)r + 1] = s->status_detail;
    }
}

// USGCMCSimulation::update_internal (us_simulation.cpp:262-266): visit count of the current grid point
template <class K>
LDO_HD void rep_count_grid_visit(Engine<K>& eng, long long* visits) {
    for (int b = 0; b < eng.OB().n_biases; b++) {
        if (eng.OB().biases[b].type != BIAS_GRID) continue;
        int off = eng.BS()->grid_off[b];
        if (off < 0) continue;
        const BiasDef& bd = eng.OB().biases[b];
        int idx = 0;
        bool inside = true;
        for (int k = 0; k < bd.n_ops; k++) {
            int v = eng.BS()->op_val[bd.op_idx[k]] - eng.BS()->grid_lo[b][k];
            if (v < 0 || v >= eng.BS()->grid_n[b][k]) {
                inside = false;
                break;
            }
            idx = idx * eng.BS()->grid_n[b][k] + v;
        }
        if (inside) visits[off + idx] += 1;
    }
}

// `engp` is one engine object per replica (shared by the warp's lanes: shared memory when staged), so
// that the move state is not replicated 32 times in per-lane local memory.
template <class K>
LDO_HD void rep_execute(Engine<K>* engp, SysState<K>* st, MoveScratch<K>* ms, RepAux* aux, const DevPtrs<K>& P, const OpArgs& a, int r) {
    Engine<K>& eng = *engp;
    rep_init_engine(eng, st, ms, aux, P, r);
    LDO_SYNCWARP();
    switch (a.op) {
    case OP_RUN: {
        if (st->status != LDO_OK) break;
        long long step = aux->step;
        long long* visits = (P.shared->has_grid && P.grid_visits) ? P.grid_visits + (size_t)aux->bs.grid_slot * LDO_GRID_CAP : nullptr;
        for (long long n = 0; n < a.n_steps; n++) {
            step++;
            eng.mc_step();
            if (st->status != LDO_OK) break;
            if (a.centering_freq != 0 && step % a.centering_freq == 0) eng.sys.center(a.centering_domain);
            if (a.constraint_check_freq != 0 && step % a.constraint_check_freq == 0) {
                if (!eng.sys.check_all_constraints()) break;
            }
            // USGCMCSimulation::update_internal counts the grid bias' STORED point (us_simulation.cpp:262-266): the
            // parameters as the last move left them. They are current except right after a window exchange, where they
            // still describe the configuration that left until a move re-evaluates them (an accepted orientation
            // rotation does not, orientation_movetype.cpp:30-65) - the reference counts that stale point, so do we.
            if (visits) rep_count_grid_visit(eng, visits);
        }
        aux->step = step;
        break;
    }
    case OP_LOAD_CONFIG: {
        rep_refresh_stack_energy(st, P.shared, aux->ctl);
        rep_load_config(eng, a);
        break;
    }
    case OP_UPDATE_ENERGY: {
        // OrigamiSystem::update_temp -> update_energy (origami_system.cpp:618-622, 808-826)
        rep_refresh_stack_energy(st, P.shared, aux->ctl);
        if (st->status == LDO_OK) eng.sys.update_energy();
        eng.update_move_params();
        break;
    }
    case OP_CHECK_CONSTRAINTS: {
        if (st->status == LDO_OK) eng.sys.check_all_constraints();
        break;
    }
    case OP_CENTER: {
        if (st->status == LDO_OK) eng.sys.center(a.centering_domain);
        break;
    }
    case OP_OBSERVE: {
        rep_observe(eng, a, r);
        break;
    }
    case OP_REFRESH_OPS: {
        eng.update_move_params();
        break;
    }
    case OP_REINIT_BIASES: {
        rep_init_biases(eng);
        break;
    }
    case OP_AFTER_WINDOW_SWAP: {
        // The reference ships the two configurations (us_simulation.cpp:848-863) and rebuilds the receiving system with
        // set_config -> initialize_staples (origami_system.cpp:327-341, 659-678), which restarts the unique chain counter
        // from the largest index among the received chains. Same here for the two replicas of an accepted swap.
        if (aux->window_swapped) {
            int mx = st->chain_uid[0];
            for (int w = 1; w < st->n_chains; w++) {
                int uid = st->chain_uid[st->order[w]];
                mx = uid > mx ? uid : mx;
            }
            st->current_c_i = mx;
            aux->window_swapped = 0;
        }
        break;
    }
    case OP_RECOMPUTE: {
        // tear down and rebuild the running energy on the staged copy (never written back)
        eng.sys.update_energy();
        a.out_recomputed[r] = st->energy;
        a.out_recomputed_stacked[r] = st->num_stacked_pairs;
        break;
    }
    }
}

// Exact enumeration (ldo_enum.cuh): replica r is worker r of the job; its state is scratch and is never written back
template <class K>
LDO_HD void rep_enumerate(Engine<K>* engp, SysState<K>* st, MoveScratch<K>* ms, RepAux* aux, const DevPtrs<K>& P, const EnumJob* job, EnumAcc* accs, int r) {
    rep_init_engine(*engp, st, ms, aux, P, r);
    rep_refresh_stack_energy(st, P.shared, aux->ctl);
    LDO_SYNCWARP();
    Enumerator<K> en(*engp, *job, accs[r]);
    en.run(r);
    LDO_SYNCWARP();
}

// ---------------------------------------------------------------------------------------------
// Kernels: one warp per replica
// ---------------------------------------------------------------------------------------------
#ifndef LDO_HOSTSIM

template <class T>
__device__ inline void warp_copy16(T* dst, const T* src) {
    static_assert(sizeof(T) % 16 == 0, "struct must be a multiple of 16 bytes");
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    const int n = sizeof(T) / 16;
    for (int i = threadIdx.x & 31; i < n; i += 32) d[i] = s[i];
    __syncwarp();
}

template <class K>
struct __align__(16) WarpSmem {
    SysState<K> st;
    __align__(16) unsigned char aux[LDO_AUX_HOT_BYTES]; // RepAux up to (not including) stats
    MoveScratch<K> ms;
    Engine<K> eng;
};

namespace ldo {
template <class K>
struct SmemLayout {
    static const unsigned stride = sizeof(WarpSmem<K>);
    static const unsigned state = offsetof(WarpSmem<K>, st);
    static const unsigned scratch = offsetof(WarpSmem<K>, ms);
    static const unsigned engine = offsetof(WarpSmem<K>, eng);
    static const unsigned rng = offsetof(WarpSmem<K>, aux) + offsetof(RepAux, rng);
    static const unsigned bias = offsetof(WarpSmem<K>, aux) + offsetof(RepAux, bs);
};
} // namespace ldo

__device__ inline void warp_copy_bytes16(void* dst, const void* src, int bytes) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x & 31; i < bytes / 16; i += 32) d[i] = s[i];
    __syncwarp();
}
static_assert(LDO_AUX_HOT_BYTES % 16 == 0, "staged part of RepAux must be a multiple of 16 bytes");

// Staged: the replica state is copied HBM -> shared memory (coalesced 128-bit), all moves run on
// shared memory, and the state is copied back once at the end.
// Launch shape of the staged kernel: LDO_BLOCK_WARPS warps per block, at least LDO_MIN_BLOCKS blocks per
// SM (this caps the registers per thread; see DESIGN.md for the measured occupancy trade-off)
#ifndef LDO_MIN_BLOCKS
#define LDO_MIN_BLOCKS (28 / LDO_BLOCK_WARPS)
#endif
// Persistent: the grid is sized to what is resident on the chip and every warp pulls replicas from a
// queue until it is empty, so that ensembles larger than one wave keep all warp slots busy. Run launches
// hand replicas out most-expensive-first (k_build_order: cost = duration in the previous run launch, the
// longest-processing-time rule for the tail), dealt round-robin into one list per SM so that every SM
// works through the same mix of expensive and cheap replicas; a warp whose list is empty steals from the
// next lists.
__device__ inline int take_item(int* heads, int n_items, int n_lists, int my_list, int grouped) {
    for (int j = 0; j < n_lists; j++) {
        // own list first; then the next lists (dealt mode), or the most expensive lists first (grouped mode)
        int l = (grouped && j > 0) ? j - 1 : my_list + j;
        if (l >= n_lists) l -= n_lists;
        int len = (n_items - l + n_lists - 1) / n_lists; // items l, l + n_lists, ...
        if (len <= 0 || *(volatile int*)&heads[l] >= len) continue;
        int k = atomicAdd(&heads[l], 1);
        if (k < len) return k * n_lists + l;
    }
    return -1;
}

template <class K>
__global__ void __launch_bounds__(32 * LDO_BLOCK_WARPS, LDO_MIN_BLOCKS) k_exec_staged(DevPtrs<K> P, OpArgs a, int n_items, int n_lists, int grouped) {
    WarpSmem<K>* ws = reinterpret_cast<WarpSmem<K>*>(ldo_smem_raw);
    int warp = threadIdx.x >> 5;
    WarpSmem<K>& w = ws[warp];
    RepAux* aux = reinterpret_cast<RepAux*>(w.aux);
#ifdef LDO_SMEM_WINDOW_BASE
    // the accessors address shared memory with literal window offsets (ldo_core.cuh): refuse to run otherwise
    if ((unsigned)__cvta_generic_to_shared(ldo_smem_raw) != LDO_SMEM_WINDOW_BASE) {
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_replicas; r += gridDim.x * blockDim.x) {
            P.states[r].status = LDO_ERR_INTERNAL;
            P.states[r].status_detail = 900;
        }
        return;
    }
#endif
    unsigned my_list;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(my_list));
    my_list %= (unsigned)n_lists;
    for (;;) {
        int idx = 0;
        if ((threadIdx.x & 31) == 0) idx = take_item(P.queue, n_items, n_lists, (int)my_list, grouped);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx < 0) break;
        int r = a.only_replica >= 0 ? a.only_replica : (a.op == OP_RUN ? P.order[idx] : idx);
        long long t0 = 0;
        if (a.op == OP_RUN) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        warp_copy16(&w.st, &P.states[r]);
        warp_copy_bytes16(w.aux, &P.aux[r], LDO_AUX_HOT_BYTES);
        rep_execute<K>(&w.eng, &w.st, &w.ms, aux, P, a, r);
        __syncwarp();
        if (a.op != OP_OBSERVE && a.op != OP_RECOMPUTE) {
            warp_copy16(&P.states[r], &w.st);
            warp_copy_bytes16(&P.aux[r], w.aux, LDO_AUX_HOT_BYTES);
        }
        if (a.op == OP_RUN && (threadIdx.x & 31) == 0) {
            long long t1;
            unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.run_timing[3 * r] = t0;
            P.run_timing[3 * r + 1] = t1;
            P.run_timing[3 * r + 2] = smid;
        }
        __syncwarp();
    }
}

// Enumeration workers: launch shape and shared-memory layout of the staged kernel (the accessors address the replica
// through the same window offsets), one worker per warp; the recursion runs on the per-thread stack
// (cudaLimitStackSize is raised by the host around the launch).
template <class K>
__global__ void __launch_bounds__(32 * LDO_BLOCK_WARPS) k_enum(DevPtrs<K> P, const EnumJob* job, EnumAcc* accs, int n_workers) {
    WarpSmem<K>* ws = reinterpret_cast<WarpSmem<K>*>(ldo_smem_raw);
    int warp = threadIdx.x >> 5;
    WarpSmem<K>& w = ws[warp];
    RepAux* aux = reinterpret_cast<RepAux*>(w.aux);
    int r = blockIdx.x * LDO_BLOCK_WARPS + warp;
    if (r >= n_workers) return;
    warp_copy16(&w.st, &P.states[r]);
    warp_copy_bytes16(w.aux, &P.aux[r], LDO_AUX_HOT_BYTES);
    rep_enumerate<K>(&w.eng, &w.st, &w.ms, aux, P, job, accs, r);
}

// Reports where the dynamic shared memory of a launch shaped like the staged kernel starts in the CTA's window
__global__ void k_probe_smem_base(unsigned* out) {
    if (threadIdx.x == 0) *out = (unsigned)__cvta_generic_to_shared(ldo_smem_raw);
}

// Order in which the next run launch hands out replicas: descending duration of the previous run launch
// (counting sort over 256 cost buckets, one block). All zeros (first launch) gives the identity.
// `grouped`: list l (items l, l + n_lists, ... of `order`) receives a CONTIGUOUS run of the cost-sorted replicas instead of
// every n_lists-th one, so that the warps of one SM work on replicas of the same regime (same hot code).
__global__ void __launch_bounds__(1024) k_build_order(const long long* run_timing, int* order, int* queue, int n, int n_lists, int grouped) {
    __shared__ long long s_max;
    __shared__ int hist[256], cursor[256];
    int t = threadIdx.x;
    if (t == 0) s_max = 0;
    if (t < 256) hist[t] = 0;
    if (t < LDO_MAX_LISTS) queue[t] = 0;
    __syncthreads();
    long long m = 0;
    for (int r = t; r < n; r += blockDim.x) {
        long long d = run_timing[3 * r + 1] - run_timing[3 * r];
#ifdef LDO_NO_LPT
        d = 0;
        order[r] = r;
#endif
        m = d > m ? d : m;
    }
    atomicMax((unsigned long long*)&s_max, (unsigned long long)m);
    __syncthreads();
#ifdef LDO_NO_LPT
    return;
#endif
    long long mx = s_max + 1;
    for (int r = t; r < n; r += blockDim.x) {
        long long d = run_timing[3 * r + 1] - run_timing[3 * r];
        atomicAdd(&hist[255 - (int)(d * 255 / mx)], 1);
    }
    __syncthreads();
    if (t == 0) {
        int acc = 0;
        for (int b = 0; b < 256; b++) {
            cursor[b] = acc;
            acc += hist[b];
        }
    }
    __syncthreads();
    // the order inside a bucket is arbitrary: replicas are independent, results do not depend on it
    for (int r = t; r < n; r += blockDim.x) {
        long long d = run_timing[3 * r + 1] - run_timing[3 * r];
        int q = atomicAdd(&cursor[255 - (int)(d * 255 / mx)], 1);
        if (grouped && n_lists > 1) {
            int base = n / n_lists, rem = n % n_lists, l, k;
            if (q < rem * (base + 1)) {
                l = q / (base + 1);
                k = q % (base + 1);
            }
            else {
                l = rem + (q - rem * (base + 1)) / base;
                k = (q - rem * (base + 1)) % base;
            }
            q = k * n_lists + l;
        }
        order[q] = r;
    }
}

// In place: state and scratch stay in HBM / L2 (large systems)
template <class K>
// 7 blocks of 4 warps per SM (72 registers, as the staged kernel): 4096 replicas are one wave
#ifndef LDO_INPLACE_MIN_BLOCKS
#define LDO_INPLACE_MIN_BLOCKS 7
#endif
__global__ void __launch_bounds__(128, LDO_INPLACE_MIN_BLOCKS) k_exec_inplace(DevPtrs<K> P, OpArgs a, int warps_per_block, SysState<K>* recompute_tmp) {
    int warp = threadIdx.x >> 5;
    int r = blockIdx.x * warps_per_block + warp;
    if (r >= a.n_replicas) return;
    if (a.only_replica >= 0 && r != a.only_replica) return;
    SysState<K>* st = &P.states[r];
    if (a.op == OP_RECOMPUTE) {
        warp_copy16(&recompute_tmp[r], &P.states[r]);
        st = &recompute_tmp[r];
    }
    rep_execute<K>(&P.engines[r], st, &P.scratch[r], &P.aux[r], P, a, r);
}

#endif // !LDO_HOSTSIM

// ---------------------------------------------------------------------------------------------
// Replica exchange (ptmc_simulation.cpp:255-412), decided on device
// ---------------------------------------------------------------------------------------------

struct ExchangeArgs {
    int variant;
    int v2_dim; // 2-D exchange (LDO_PT_2D): slots form a [ladder_len / v2_dim][v2_dim] grid; 0 otherwise
    long long swap_i;
    int n_ladders, ladder_len;
    int rank, n_ranks, n_local, n_global;
    int n_staple_types;
    unsigned long long seed;
    const double* dependent; // [n_global][LDO_DEP_FIXED + nst]
    int* slot_to_replica; // [n_ladders][ladder_len], updated in place (m_q_to_repi)
    // control variables per slot of a ladder [ladder_len] (m_control_qs), fixed for the whole run
    const int* slot_temp_idx;
    const double* slot_temp;
    const double* slot_staple_u_mult;
    const double* slot_bias_mult;
    const double* slot_stacking_mult;
    long long* attempts; // [n_ladders][ladder_len - 1]
    long long* accepts;
    const double* reduced_staple_u; // [nst]: ln M - (2L-1) ln 6 (origami_system.cpp:965-990)
    RepAux* aux; // local replicas
    // replay mode (ldo_set_exchange_tape, single ladder): the uniform reals of the master's test_acceptance calls
    // in the order the reference drew them, and where every exchange round's draws start;
    // exchange_tape_state[0] = next position, [1] = end of this round's draws, [2] = draws wanted beyond them,
    // [3] = draws of finished rounds left unused
    const double* exchange_tape;
    const long long* exchange_tape_offsets; // [n_rounds + 1]
    long long exchange_tape_rounds;
    long long* exchange_tape_state;
};

// One thread per ladder: the neighbour tests of one ladder are sequential in the reference only
// through the shared RNG; here every (swap, ladder, pair) owns a Philox counter, so all ranks
// reproduce the same decisions without communication.
// The slots of a ladder are dealt to the ranks in serpentine order (0 1 .. G-1, G-1 .. 1 0, 0 1 ..): the cost of
// a move falls monotonically along the temperature ladder, and this gives every rank the same sum (plain
// round-robin leaves rank 0 with the coldest slot of every group: 7 % more work at 4 GPUs). Replica k of ladder
// l lives on rank exchange_rank_of(k) at local index l * S + k / n_ranks with S = ladder_len / n_ranks; the
// all-gathered buffer is rank-major.
LDO_HD inline int exchange_rank_of(const ExchangeArgs& x, int k) {
    int b = k / x.n_ranks, pos = k % x.n_ranks;
    return (b & 1) ? x.n_ranks - 1 - pos : pos;
}
LDO_HD inline int exchange_local_index(const ExchangeArgs& x, int l, int k) {
    int S = x.ladder_len / x.n_ranks;
    return l * S + k / x.n_ranks;
}
LDO_HD inline size_t exchange_gathered_index(const ExchangeArgs& x, int l, int k) {
    return (size_t)exchange_rank_of(x, k) * x.n_local + exchange_local_index(x, l, k);
}

// calc_acceptance_p (ptmc_simulation.cpp:275-313). d1 / d2 = {enthalpy, bias, stacking, staple counts...} of the
// two replicas, each in units of its own kT. The staple sum runs over the n_types - 1 staple identities (the
// reference's loop bound reads one past the end, App. A1).
LDO_HD inline double exchange_acceptance_p(int n_staple_types, const double* reduced_staple_u, double temp1, double temp2,
                                           double um1, double um2, double sm1, double sm2, const double* d1, const double* d2) {
    double DBU_DN = 0;
    for (int t = 0; t < n_staple_types; t++) {
        double N1 = d1[LDO_DEP_FIXED + t], N2 = d2[LDO_DEP_FIXED + t];
        double u1 = reduced_staple_u[t] * temp1 * um1;
        double u2 = reduced_staple_u[t] * temp2 * um2;
        DBU_DN += (u2 / temp2 - u1 / temp1) * (N2 - N1);
    }
    double DB = 1 / temp2 - 1 / temp1;
    double DH = d2[0] * temp2 - d1[0] * temp1;
    double Dstacking = d2[2] * temp2 - d1[2] * temp1;
    double DBM = sm2 / temp2 - sm1 / temp1;
    double DBias = d2[1] * temp2 - d1[1] * temp1;
    return fmin(1.0, exp(DB * (DH + DBias) + DBM * Dstacking - DBU_DN));
}

// Swap test of the slots si, sj of ladder l (calc_acceptance_p + test_acceptance, ptmc_simulation.cpp:255-313);
// `counter` indexes attempts / accepts
LDO_HD inline void exchange_pair(const ExchangeArgs& x, int l, int* q2r, int i, int j, size_t counter) {
    int nq = LDO_DEP_FIXED + x.n_staple_types;
    {
        size_t si = (size_t)i, sj = (size_t)j;
        x.attempts[counter]++;
        int rep1 = q2r[i], rep2 = q2r[j];
        const double* d1 = x.dependent + exchange_gathered_index(x, l, rep1) * nq;
        const double* d2 = x.dependent + exchange_gathered_index(x, l, rep2) * nq;
        double temp1 = x.slot_temp[si], temp2 = x.slot_temp[sj];
        double sm1 = x.slot_stacking_mult[si], sm2 = x.slot_stacking_mult[sj];
        // each replica's chemical potentials as it computed them itself: its own multiplier (the slot's once the variant
        // exchanges it; what it was initialised with otherwise and in the first round, App. A17)
        double um1 = d1[3], um2 = d2[3];
        double p_accept = exchange_acceptance_p(x.n_staple_types, x.reduced_staple_u, temp1, temp2, um1, um2, sm1, sm2, d1, d2);
        bool accept;
        if (p_accept == 1) {
            accept = true;
        }
        else if (x.exchange_tape) {
            // a test the reference decided without a draw (its p rounded to exactly 1, ours to 1 - 1e-16) finds no draw
            // of this round left: it is taken as accepted, and counted
            long long* ts = x.exchange_tape_state;
            double prob = 0.0;
            if (ts[0] < ts[1]) prob = x.exchange_tape[ts[0]++];
            else ts[2] += 1;
            accept = p_accept > prob;
        }
        else {
            Rng g;
            g.tape = nullptr;
            g.key0 = (uint32_t)x.seed;
            g.key1 = (uint32_t)(x.seed >> 32);
            g.subseq = (uint32_t)(l * x.ladder_len + i);
            g.stream = 0x45584348u; // "EXCH"
            uint32_t o[4];
            philox4x32_10(g, (unsigned long long)x.swap_i, o);
            unsigned long long u = ((unsigned long long)o[0] << 32) | o[1];
            double prob = (double)(u >> 11) * (1.0 / 9007199254740992.0);
            accept = p_accept > prob;
        }
        if (accept) {
            x.accepts[counter]++;
            q2r[i] = rep2;
            q2r[j] = rep1;
        }
    }
}

LDO_HD inline void exchange_ladder(const ExchangeArgs& x, int l) {
    int* q2r = x.slot_to_replica + (size_t)l * x.ladder_len;
    if (x.exchange_tape) {
        long long* ts = x.exchange_tape_state;
        long long r = x.swap_i - 1;
        if (r >= 0 && r < x.exchange_tape_rounds) {
            ts[0] = x.exchange_tape_offsets[r];
            ts[1] = x.exchange_tape_offsets[r + 1];
        }
        else {
            ts[1] = ts[0];
        }
    }
    if (x.variant == LDO_PT_2D) {
        // TwoDPTGCMCSimulation::attempt_exchange (ptmc_simulation.cpp:495-560): slot (i, j) = i * v2 + j with
        // i the temperature index and j the stacking-multiplier index; four alternating pair sets
        // (ptmc_simulation.hpp:158-164); counters are [swap_i % 2][i][j]
        int v2 = x.v2_dim, v1 = x.ladder_len / x.v2_dim;
        int set = (int)(x.swap_i % 4), swap_v = (int)(x.swap_i % 2);
        int i_start = set == 2 ? 1 : 0, j_start = set == 3 ? 1 : 0;
        int i_incr = (set & 1) ? 1 : 2, j_incr = (set & 1) ? 2 : 1;
        int i_end = (set & 1) ? v1 : v1 - 1, j_end = (set & 1) ? v2 - 1 : v2;
        int rep_incr = (set & 1) ? 1 : v2;
        for (int i = i_start; i < i_end; i += i_incr) {
            for (int j = j_start; j < j_end; j += j_incr) {
                int a = i * v2 + j;
                exchange_pair(x, l, q2r, a, a + rep_incr, ((size_t)l * 2 + swap_v) * x.ladder_len + a);
            }
        }
    }
    else {
        int swap_set = (int)(x.swap_i % 2);
        for (int i = swap_set; i < x.ladder_len - 1; i += 2) {
            exchange_pair(x, l, q2r, i, i + 1, (size_t)l * (x.ladder_len - 1) + i);
        }
    }
    if (x.exchange_tape) x.exchange_tape_state[3] += x.exchange_tape_state[1] - x.exchange_tape_state[0];
    // master_send (ptmc_simulation.cpp:212-226): every replica receives the control variables of its slot
    for (int i = 0; i < x.ladder_len; i++) {
        if (exchange_rank_of(x, q2r[i]) != x.rank) continue;
        int loc = exchange_local_index(x, l, q2r[i]);
        size_t si = (size_t)i;
        Control& c = x.aux[loc].ctl;
        {
            // table-cache bookkeeping of the replica (see RepAux): the key of update_temp is (temperature, stacking
            // multiplier), the multiplier being 1 for the variants that call update_temp(temp) (ptmc_simulation.cpp:651-680)
            bool with_mult = x.variant == LDO_PT_ST || x.variant == LDO_PT_2D;
            RepAux& ra = x.aux[loc];
            int cur_key = -1, new_key = -1;
            for (int k = 0; k < x.ladder_len && (cur_key < 0 || new_key < 0); k++) {
                bool same_mult_cur = !with_mult || x.slot_stacking_mult[k] == c.stacking_mult;
                bool same_mult_new = !with_mult || x.slot_stacking_mult[k] == x.slot_stacking_mult[si];
                if (cur_key < 0 && x.slot_temp_idx[k] == c.temp_idx && same_mult_cur) cur_key = k;
                if (new_key < 0 && x.slot_temp_idx[k] == x.slot_temp_idx[si] && same_mult_new) new_key = k;
            }
            if (x.ladder_len <= 64 && new_key >= 0) {
                if (cur_key >= 0) ra.table_cache_mask |= 1ull << cur_key; // the round that just ran used it
                if (!((ra.table_cache_mask >> new_key) & 1ull)) {
                    ra.table_cache_mask |= 1ull << new_key;
                    ra.init_temp_idx = x.slot_temp_idx[si];
                }
                else if (ra.init_temp_idx < 0) {
                    ra.init_temp_idx = c.temp_idx;
                }
            }
        }
        if (x.variant == LDO_PT_ST) {
            c.temp_idx = x.slot_temp_idx[si];
            c.temp = x.slot_temp[si];
            c.stacking_mult = x.slot_stacking_mult[si];
        }
        else if (x.variant == LDO_PT_2D) {
            // TwoDPTGCMCSimulation::update_control_qs (ptmc_simulation.cpp:595-601)
            c.temp_idx = x.slot_temp_idx[si];
            c.temp = x.slot_temp[si];
            c.stacking_mult = x.slot_stacking_mult[si];
            c.staple_u_mult = x.slot_staple_u_mult[si];
        }
        else {
            c.temp_idx = x.slot_temp_idx[si];
            c.temp = x.slot_temp[si];
            if (x.variant == LDO_PT_UT || x.variant == LDO_PT_HUT) c.staple_u_mult = x.slot_staple_u_mult[si];
            // hut_parallel_tempering: HUTPTGCMCSimulation::update_control_qs passes the exchanged bias multiplier to
            // OrigamiSystem::update_bias_mult, which is an empty virtual that nothing overrides
            // (origami_system.hpp:165): the system keeps bias_funcs_mult. The multiplier is relabelled in the
            // .swp header only - nothing to apply here.
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Window exchange for replica-exchange multi-window umbrella sampling (us_simulation.cpp:770-864)
// ---------------------------------------------------------------------------------------------

struct WindowExchangeArgs {
    long long swap_i;
    int n_ladders, n_windows;
    int grid_bias; // index of the Grid bias whose point / values decide the swap
    int n_window_biases; // well biases overridden per window
    int window_bias[LDO_MAX_BIASES];
    unsigned long long seed;
    int* window_to_replica; // [n_ladders][n_windows] (m_win_to_configi)
    long long* attempts; // [n_ladders][n_windows - 1]
    long long* accepts;
    RepAux* aux;
    const double* grid_vals;
    const OpsBiasConst* ob;
    // replay mode (ldo_set_exchange_tape): the master's test_acceptance draws in the order the reference made them
    const double* exchange_tape;
    long long exchange_tape_len;
    long long* exchange_tape_state;
};

LDO_HD inline double window_grid_value(const WindowExchangeArgs& x, const BiasState& owner, const int* point) {
    int b = x.grid_bias;
    int off = owner.grid_off[b];
    if (off < 0) return 0;
    const BiasDef& bd = x.ob->biases[b];
    int idx = 0;
    for (int k = 0; k < bd.n_ops; k++) {
        int v = point[k] - owner.grid_lo[b][k];
        if (v < 0 || v >= owner.grid_n[b][k]) return 0;
        idx = idx * owner.grid_n[b][k] + v;
    }
    double g = x.grid_vals[(size_t)owner.grid_slot * LDO_GRID_CAP + off + idx];
    return g == g ? g : 0; // absent points read as 0 (unordered_map::operator[], us_simulation.cpp:824-827)
}

// One thread per ladder of windows. An accepted swap exchanges the window-specific state of the two
// replicas (limits of the window biases, grid boxes and grid slot = bias values and visit histogram);
// the reference ships the configurations instead (:848-863), which is the same relabelled.
LDO_HD inline void window_exchange_ladder(const WindowExchangeArgs& x, int l) {
    int* w2r = x.window_to_replica + (size_t)l * x.n_windows;
    const BiasDef& gb = x.ob->biases[x.grid_bias];
    int swap_set = (int)(x.swap_i % 2);
    for (int i = swap_set; i < x.n_windows - 1; i += 2) {
        x.attempts[(size_t)l * (x.n_windows - 1) + i]++;
        RepAux& a1 = x.aux[l * x.n_windows + w2r[i]];
        RepAux& a2 = x.aux[l * x.n_windows + w2r[i + 1]];
        int p1[LDO_MAX_GRID_DIM], p2[LDO_MAX_GRID_DIM];
        bool same = true;
        for (int k = 0; k < gb.n_ops; k++) {
            p1[k] = a1.bs.op_val[gb.op_idx[k]];
            p2[k] = a2.bs.op_val[gb.op_idx[k]];
            if (p1[k] != p2[k]) same = false;
        }
        // both points must lie inside both windows (:800-815); window limits are those of the well biases
        bool inside = true;
        for (int j = 0; j < x.n_window_biases && inside; j++) {
            int wb = x.window_bias[j];
            int op = x.ob->biases[wb].op_idx[0];
            int v1 = a1.bs.op_val[op], v2 = a2.bs.op_val[op];
            if (v2 < a1.bs.win_min[wb] || v2 > a1.bs.win_max[wb]) inside = false;
            if (v1 < a2.bs.win_min[wb] || v1 > a2.bs.win_max[wb]) inside = false;
        }
        if (!inside) continue;
        bool accepted = true;
        if (!same) {
            double d1 = window_grid_value(x, a1.bs, p1) - window_grid_value(x, a2.bs, p1);
            double d2 = window_grid_value(x, a2.bs, p2) - window_grid_value(x, a1.bs, p2);
            double p_accept = fmin(1.0, exp(d1 + d2));
            if (p_accept != 1 && x.exchange_tape) {
                long long* ts = x.exchange_tape_state;
                double prob = 0.0;
                if (ts[0] < x.exchange_tape_len) prob = x.exchange_tape[ts[0]++];
                else ts[2] += 1;
                accepted = p_accept > prob;
            }
            else if (p_accept != 1) {
                Rng g;
                g.tape = nullptr;
                g.key0 = (uint32_t)x.seed;
                g.key1 = (uint32_t)(x.seed >> 32);
                g.subseq = (uint32_t)(l * x.n_windows + i);
                g.stream = 0x57494e44u; // "WIND"
                uint32_t o[4];
                philox4x32_10(g, (unsigned long long)x.swap_i, o);
                unsigned long long u = ((unsigned long long)o[0] << 32) | o[1];
                accepted = p_accept > (double)(u >> 11) * (1.0 / 9007199254740992.0);
            }
        }
        if (!accepted) continue;
        x.accepts[(size_t)l * (x.n_windows - 1) + i]++;
        int t = w2r[i];
        w2r[i] = w2r[i + 1];
        w2r[i + 1] = t;
        // Everything that belongs to the WINDOW changes hands, because in the reference the rank keeps its
        // SystemOrderParams / SystemBiases objects, its bias grid and histogram and its random generator while the
        // configuration moves: the stored parameter and bias values (stale with respect to the new configuration until
        // the next move re-evaluates them, exactly as in the reference), the window limits, the grid box and slot, and -
        // in replay mode - the window's tape.
        BiasState tb = a1.bs;
        a1.bs = a2.bs;
        a2.bs = tb;
        const TapeDraw* tt = a1.rng.tape;
        a1.rng.tape = a2.rng.tape;
        a2.rng.tape = tt;
        long long tl = a1.rng.tape_len;
        a1.rng.tape_len = a2.rng.tape_len;
        a2.rng.tape_len = tl;
        tl = a1.rng.tape_pos;
        a1.rng.tape_pos = a2.rng.tape_pos;
        a2.rng.tape_pos = tl;
        a1.window_swapped = 1;
        a2.window_swapped = 1;
    }
}

#ifndef LDO_HOSTSIM
__global__ void k_window_exchange(WindowExchangeArgs x) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < x.n_ladders) window_exchange_ladder(x, l);
}
__global__ void k_exchange(ExchangeArgs x) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < x.n_ladders) exchange_ladder(x, l);
}
#endif

// ---------------------------------------------------------------------------------------------
// Checkpoint blobs (ldo_checkpoint_save / ldo_checkpoint_load): pack / unpack between the engine's arrays and a
// contiguous staging buffer of per-replica blobs, so that the host copy is one transfer
// ---------------------------------------------------------------------------------------------
template <class K>
struct BlobArgs {
    SysState<K>* states;
    RepAux* aux;
    double* grid_vals;
    long long* grid_visits;
    unsigned char* stage;
    int first, count;
    unsigned blob_bytes;
    bool with_grid;
};
// `lane` / `nlanes`: the words of one blob are spread over the lanes of a warp (1 on the host)
template <class K>
LDO_HD inline void blob_copy_words(void* dst, const void* src, size_t bytes, int lane, int nlanes) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (size_t i = lane; i < bytes / 4; i += nlanes) d[i] = s[i];
}
template <class K>
LDO_HD inline void blob_pack_one(const BlobArgs<K>& b, int i, int lane = 0, int nlanes = 1) {
    int r = b.first + i;
    unsigned char* o = b.stage + (size_t)i * b.blob_bytes;
    blob_copy_words<K>(o, &b.states[r], sizeof(SysState<K>), lane, nlanes);
    blob_copy_words<K>(o + sizeof(SysState<K>), &b.aux[r], sizeof(RepAux), lane, nlanes);
    if (b.with_grid) {
        size_t slot = (size_t)b.aux[r].bs.grid_slot * LDO_GRID_CAP;
        unsigned char* g = o + sizeof(SysState<K>) + sizeof(RepAux);
        blob_copy_words<K>(g, b.grid_vals + slot, sizeof(double) * LDO_GRID_CAP, lane, nlanes);
        blob_copy_words<K>(g + sizeof(double) * LDO_GRID_CAP, b.grid_visits + slot, sizeof(long long) * LDO_GRID_CAP, lane, nlanes);
    }
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
    if (lane == 0) {
        // tapes are not part of a checkpoint: no device pointer leaves the engine
        RepAux* a = reinterpret_cast<RepAux*>(o + sizeof(SysState<K>));
        a->rng.tape = nullptr;
        a->rng.tape_len = 0;
        a->rng.tape_pos = 0;
    }
}
template <class K>
LDO_HD inline void blob_unpack_one(const BlobArgs<K>& b, int i, int lane = 0, int nlanes = 1) {
    int r = b.first + i;
    const unsigned char* o = b.stage + (size_t)i * b.blob_bytes;
    blob_copy_words<K>(&b.states[r], o, sizeof(SysState<K>), lane, nlanes);
    blob_copy_words<K>(&b.aux[r], o + sizeof(SysState<K>), sizeof(RepAux), lane, nlanes);
    if (b.with_grid) {
        size_t slot = (size_t)r * LDO_GRID_CAP;
        const unsigned char* g = o + sizeof(SysState<K>) + sizeof(RepAux);
        blob_copy_words<K>(b.grid_vals + slot, g, sizeof(double) * LDO_GRID_CAP, lane, nlanes);
        blob_copy_words<K>(b.grid_visits + slot, g + sizeof(double) * LDO_GRID_CAP, sizeof(long long) * LDO_GRID_CAP, lane, nlanes);
    }
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
    if (lane == 0) {
        RepAux& a = b.aux[r];
        a.rng.tape = nullptr;
        a.rng.tape_len = 0;
        a.rng.tape_pos = 0;
        a.bs.grid_slot = r;
    }
}
#ifndef LDO_HOSTSIM
template <class K>
__global__ void k_blob_pack(BlobArgs<K> b) {
    int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i < b.count) blob_pack_one(b, i, threadIdx.x & 31, 32);
}
template <class K>
__global__ void k_blob_unpack(BlobArgs<K> b) {
    int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i < b.count) blob_unpack_one(b, i, threadIdx.x & 31, 32);
}
#endif

// ---------------------------------------------------------------------------------------------
// Engine object
// ---------------------------------------------------------------------------------------------

#ifdef LDO_SMALL_INPLACE
// profiling variant: snodin-class systems run in place (HBM / L2 / L1) like the large ones
typedef Caps<80, 26, 7, 16, false, 48, 40, 40> CapsSmall;
#define LDO_SMALL_STAGED false
#else
typedef Caps<80, 26, 7, 16, true, 48, 40, 40> CapsSmall; // snodin-class systems: staged in shared memory
#define LDO_SMALL_STAGED true
#endif
typedef Caps<512, 176, 11, 96, false, 511, 512, 1026> CapsLarge; // large scaffolds: in place in HBM/L2

static thread_local std::string g_create_error;
// Constant memory is per device: the engine whose descriptions currently sit in each device's copy. One host thread
// drives one engine / GPU (ldo_b200.h), so entries of different devices are touched by different threads.
#define LDO_MAX_DEVICES 64
static const void* g_const_owner[LDO_MAX_DEVICES] = {nullptr};
static long long g_const_version[LDO_MAX_DEVICES] = {0};
#include <atomic>
static std::atomic<long long> g_const_counter {0};

struct EngineBase {
    virtual ~EngineBase() {}
    std::string err;
    int R = 0;
    int device = 0;
    Shared shared;
    int n_ident = 0;
    std::vector<double> temps;
    std::vector<double> reduced_staple_u;
    dev_stream_t stream;
    int fail(const std::string& m) {
        err = m;
        return -1;
    }
    virtual int set_tables(int n_temps, int n_ident, const double* temps, const double* e, const double* h, const double* s, const double* init) = 0;
    virtual int push_shared() = 0;
    virtual int exec(OpArgs& a, bool sync) = 0;
    virtual long long* run_timing_ptr() = 0;
    virtual int get_aux(int first, int count, RepAux* out) = 0;
    virtual int put_aux(int first, int count, const RepAux* in) = 0;
    virtual int get_state_raw(int replica, std::vector<unsigned char>& blob) = 0;
    virtual int decode_state(int replica, int* n_chains, int* ci, int* cid, int* cl, int* pos, int* ore, int* st, int* bd) = 0;
    virtual void capacity(int* c, int* d) = 0;
    virtual int set_grid(int replica, int bias, const int* lo, const int* n, const double* vals) = 0;
    virtual int get_visits(int replica, int bias, long long* counts, int clear) = 0;
    virtual int attach_tape(int replica, const ldo_tape_draw* draws, long long n) = 0;
    virtual int exchange(ExchangeArgs& x, const double* dependent_host, int* slot_to_replica, long long* attempts, long long* accepts) = 0;
    virtual int exchange_state_set(int n_slots, int n_counters, const int* q2r, const long long* att, const long long* acc) = 0;
    virtual int exchange_state_get(int n_slots, int n_counters, int* q2r, long long* att, long long* acc) = 0;
    virtual int exchange_resident(ExchangeArgs& x) = 0;
    virtual int exchange_buffers(int n_global, void** send, void** recv, int* nq) = 0;
    virtual int window_exchange(WindowExchangeArgs& x, int* window_to_replica, long long* attempts, long long* accepts) = 0;
    virtual int alloc_outputs() = 0;
    virtual int set_exchange_tape(const double* reals, long long n, const long long* offsets, long long n_rounds) = 0;
    virtual int exchange_tape_state(long long* missing, long long* unused) = 0;
    virtual size_t blob_size() = 0;
    virtual size_t state_bytes() = 0;
    virtual int get_blobs(int first, int count, void* host) = 0;
    virtual int put_blobs(int first, int count, const void* host) = 0;
    virtual int enumerate(EnumJob& job, std::vector<EnumAcc>& accs) = 0;
    virtual int enable_trackers(bool on) = 0;
    virtual int get_trackers(int replica, TrackStats* out) = 0;
    std::vector<EnumAcc> enum_accs; // per-worker tables of the last enumeration job (host copy)
    int domain_update_biases = 0; // ldo_set_domain_update_biases
    int keep_bias_on_load = 0; // set around the load of ldo_replace_config
    long long launches = 0; // kernels launched so far (ldo_launch_count)
    long long const_version = 0;
    // device output buffers
    double* d_energies = nullptr;
    int* d_counters = nullptr;
    int* d_ops = nullptr;
    int* d_staples = nullptr;
    double* d_dependent = nullptr;
    double* d_recomputed = nullptr;
    int* d_recomputed_stacked = nullptr;
    int* d_status = nullptr;
    // control-variable ladder for replica exchange (m_control_qs, ptmc_simulation.cpp:341-346)
    std::vector<int> ladder_temp_idx;
    std::vector<double> ladder_staple_u_mult, ladder_bias_mult, ladder_stacking_mult;
};

template <class K, bool STAGED>
struct EngineImpl: EngineBase {
    DevPtrs<K> P;
    Shared* d_shared = nullptr;
    TempTables* d_tables = nullptr;
    double* d_table_data = nullptr;
    std::vector<void*> tape_bufs;
    std::vector<void*> parked_tapes;
    SysState<K>* d_recompute_tmp = nullptr;
    int* d_cfg = nullptr;
    size_t cfg_cap = 0;
    // exchange
    double* d_dep_all = nullptr;
    int dep_all_n = 0;
    int* d_q2r = nullptr;
    long long* d_att = nullptr;
    long long* d_acc = nullptr;
    int* d_slot_tidx = nullptr;
    double* d_slot_vals = nullptr;
    double* d_red_u = nullptr;
    int exch_cap = 0;
    int warps_per_block = 4;
    int n_sm = 1;
    int resident_blocks = 1; // staged kernel: blocks resident on the whole device (persistent grid)
    int order_grouped = 0; // LDO_ORDER_MODE=grouped: one SM works on replicas of one cost class (profiling knob)

    EngineImpl() {
        memset(&P, 0, sizeof(P));
        memset(&shared, 0, sizeof(shared));
    }
    ~EngineImpl() override {
        dev_free(P.states);
        dev_free(P.scratch);
        dev_free(P.cold);
        dev_free(P.engines);
        dev_free(P.aux);
        dev_free(d_shared);
        dev_free(d_tables);
        dev_free(d_table_data);
        dev_free(P.grid_vals);
        dev_free(P.grid_visits);
        dev_free(P.run_timing);
        dev_free(P.queue);
        dev_free(P.order);
        dev_free(d_recompute_tmp);
        dev_free(d_cfg);
        dev_free(d_energies);
        dev_free(d_counters);
        dev_free(d_ops);
        dev_free(d_staples);
        dev_free(d_dependent);
        dev_free(d_recomputed);
        dev_free(d_recomputed_stacked);
        dev_free(d_status);
        dev_free(d_dep_all);
        dev_free(d_q2r);
        dev_free(d_att);
        dev_free(d_acc);
        dev_free(d_slot_tidx);
        dev_free(d_slot_vals);
        dev_free(d_red_u);
        dev_free(d_blob_stage);
        dev_free(d_exch_tape);
        dev_free(d_exch_tape_state);
        dev_free(d_exch_tape_offsets);
        dev_free(d_enum_job);
        dev_free(d_enum_acc);
        dev_free(P.track);
#ifndef LDO_HOSTSIM
        if (enum_stack_raised) cudaDeviceSetLimit(cudaLimitStackSize, enum_old_stack);
#endif
        for (void* p: tape_bufs) dev_free(p);
        for (void* p: parked_tapes) dev_free(p);
        if (g_const_owner[device % LDO_MAX_DEVICES] == this) g_const_owner[device % LDO_MAX_DEVICES] = nullptr;
#ifndef LDO_HOSTSIM
        cudaStreamDestroy(stream);
#endif
    }

    int init(int n_replicas) {
        R = n_replicas;
#ifndef LDO_HOSTSIM
        if (chk(cudaSetDevice(device))) return fail(dev_err());
        if (chk(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking))) return fail(dev_err());
#else
        stream = 0;
#endif
        if (dev_malloc((void**)&P.states, sizeof(SysState<K>) * R)) return fail(dev_err());
        if (dev_malloc((void**)&P.cold, sizeof(ColdScratch<K>) * R)) return fail(dev_err());
        if (!STAGED) {
            if (dev_malloc((void**)&P.engines, sizeof(Engine<K>) * R)) return fail(dev_err());
            if (dev_malloc((void**)&P.scratch, sizeof(MoveScratch<K>) * R)) return fail(dev_err());
            if (dev_malloc((void**)&d_recompute_tmp, sizeof(SysState<K>) * R)) return fail(dev_err());
        }
#ifdef LDO_HOSTSIM
        if (STAGED && dev_malloc((void**)&P.scratch, sizeof(MoveScratch<K>))) return fail(dev_err());
        if (STAGED && dev_malloc((void**)&d_recompute_tmp, sizeof(SysState<K>))) return fail(dev_err());
#endif
        if (dev_malloc((void**)&P.aux, sizeof(RepAux) * R)) return fail(dev_err());
        if (dev_malloc((void**)&d_shared, sizeof(Shared))) return fail(dev_err());
        P.shared = d_shared;
        tape_bufs.assign(R, nullptr);
        // default per-replica aux
        std::vector<RepAux> aux(R);
        memset(aux.data(), 0, sizeof(RepAux) * R);
        for (int r = 0; r < R; r++) {
            aux[r].ctl.temp_idx = 0;
            aux[r].ctl.temp = 0;
            aux[r].ctl.staple_u_mult = 1;
            aux[r].ctl.bias_mult = 1;
            aux[r].ctl.stacking_mult = 1;
            aux[r].init_temp_idx = -1;
            aux[r].table_cache_mask = 0;
            aux[r].rng.subseq = (uint32_t)r;
            for (int b = 0; b < LDO_MAX_BIASES; b++) aux[r].bs.grid_off[b] = -1;
            aux[r].bs.grid_slot = r;
        }
        if (dev_h2d(P.aux, aux.data(), sizeof(RepAux) * R, stream)) return fail(dev_err());
        return alloc_outputs();
    }

    int alloc_outputs() override {
        int nst = shared.sc.n_types - 1;
        if (dev_malloc((void**)&d_energies, sizeof(double) * 5 * R)) return fail(dev_err());
        if (dev_malloc((void**)&d_counters, sizeof(int) * 9 * R)) return fail(dev_err());
        if (dev_malloc((void**)&d_ops, sizeof(int) * LDO_MAX_OPS * R)) return fail(dev_err());
        if (dev_malloc((void**)&d_staples, sizeof(int) * (nst > 0 ? nst : 1) * R)) return fail(dev_err());
        if (dev_malloc((void**)&d_dependent, sizeof(double) * (LDO_DEP_FIXED + nst) * R)) return fail(dev_err());
        if (dev_malloc((void**)&d_recomputed, sizeof(double) * R)) return fail(dev_err());
        if (dev_malloc((void**)&P.run_timing, sizeof(long long) * 3 * R)) return fail(dev_err());
        if (dev_malloc((void**)&P.queue, sizeof(int) * LDO_MAX_LISTS)) return fail(dev_err());
        if (dev_malloc((void**)&P.order, sizeof(int) * R)) return fail(dev_err());
        if (dev_malloc((void**)&d_recomputed_stacked, sizeof(int) * R)) return fail(dev_err());
        if (dev_malloc((void**)&d_status, sizeof(int) * 2 * R)) return fail(dev_err());
        return 0;
    }

    int push_shared() override {
        const_version = ++g_const_counter;
        if (dev_h2d(d_shared, &shared, sizeof(Shared), stream)) return fail(dev_err());
        if (shared.has_grid && !P.grid_vals) {
            if (dev_malloc((void**)&P.grid_vals, sizeof(double) * LDO_GRID_CAP * R)) return fail(dev_err());
            if (dev_malloc((void**)&P.grid_visits, sizeof(long long) * LDO_GRID_CAP * R)) return fail(dev_err());
        }
        return 0;
    }

    int set_tables(int n_temps, int nid, const double* tv, const double* e, const double* h, const double* s, const double* init) override {
        n_ident = nid;
        shared.sc.n_ident = nid;
        shared.n_temps = n_temps;
        size_t tsz = (size_t)(2 * nid + 1) * (2 * nid + 1);
        dev_free(d_table_data);
        dev_free(d_tables);
        d_table_data = nullptr;
        d_tables = nullptr;
        if (dev_malloc((void**)&d_table_data, sizeof(double) * tsz * 3 * n_temps)) return fail(dev_err());
        if (dev_malloc((void**)&d_tables, sizeof(TempTables) * n_temps)) return fail(dev_err());
        std::vector<TempTables> tt(n_temps);
        temps.assign(tv, tv + n_temps);
        for (int t = 0; t < n_temps; t++) {
            tt[t].temp = tv[t];
            tt[t].init_energy = init[3 * t];
            tt[t].init_enthalpy = init[3 * t + 1];
            tt[t].init_entropy = init[3 * t + 2];
            tt[t].hyb_energy = d_table_data + (size_t)(3 * t) * tsz;
            tt[t].hyb_enthalpy = d_table_data + (size_t)(3 * t + 1) * tsz;
            tt[t].hyb_entropy = d_table_data + (size_t)(3 * t + 2) * tsz;
            if (dev_h2d((void*)tt[t].hyb_energy, e + t * tsz, sizeof(double) * tsz, stream)) return fail(dev_err());
            if (dev_h2d((void*)tt[t].hyb_enthalpy, h + t * tsz, sizeof(double) * tsz, stream)) return fail(dev_err());
            if (dev_h2d((void*)tt[t].hyb_entropy, s + t * tsz, sizeof(double) * tsz, stream)) return fail(dev_err());
        }
        if (dev_h2d(d_tables, tt.data(), sizeof(TempTables) * n_temps, stream)) return fail(dev_err());
        P.tables = d_tables;
        // default control: temperature slot 0
        std::vector<RepAux> aux(R);
        if (get_aux(0, R, aux.data())) return -1;
        for (int r = 0; r < R; r++) {
            if (aux[r].ctl.temp_idx >= n_temps) aux[r].ctl.temp_idx = 0;
            aux[r].ctl.temp = tv[aux[r].ctl.temp_idx];
        }
        if (put_aux(0, R, aux.data())) return -1;
        return push_shared();
    }

    int exec(OpArgs& a, bool sync) override {
        a.n_replicas = R;
        if (!P.tables) return fail("temperature tables not set");
#ifdef LDO_HOSTSIM
        for (int r = 0; r < R; r++) {
            if (a.only_replica >= 0 && r != a.only_replica) continue;
            SysState<K>* st = &P.states[r];
            MoveScratch<K>* ms = STAGED ? P.scratch : &P.scratch[r];
            if (a.op == OP_RECOMPUTE) {
                SysState<K>* tmp = STAGED ? d_recompute_tmp : &d_recompute_tmp[r];
                memcpy(tmp, st, sizeof(SysState<K>));
                st = tmp;
            }
            if (tracked_kernels()) {
                typedef Tracked<K> KT;
                Engine<KT> eng;
                rep_execute<KT>(&eng, reinterpret_cast<SysState<KT>*>(st), reinterpret_cast<MoveScratch<KT>*>(ms), &P.aux[r], tracked_ptrs(P), a, r);
            }
            else {
                Engine<K> eng;
                rep_execute<K>(&eng, st, ms, &P.aux[r], P, a, r);
            }
        }
        (void)sync;
        return 0;
#else
        int wpb = warps_per_block;
        int blocks = (R + wpb - 1) / wpb;
        if (upload_constants()) return -1;
        if constexpr (STAGED) {
            size_t smem = sizeof(WarpSmem<K>) * wpb;
            int n_items = a.only_replica >= 0 ? 1 : R;
            int n_lists = a.op == OP_RUN && n_items > 1 ? (n_sm < LDO_MAX_LISTS ? n_sm : LDO_MAX_LISTS) : 1;
            int grouped = order_grouped && n_lists > 1 && R >= n_lists ? 1 : 0;
            if (a.op == OP_RUN && a.only_replica < 0) {
                k_build_order<<<1, 1024, 0, stream>>>(P.run_timing, P.order, P.queue, R, n_lists, grouped);
                launches++;
            }
            else if (chk(cudaMemsetAsync(P.queue, 0, sizeof(int) * LDO_MAX_LISTS, stream))) return fail(dev_err());
            blocks = (n_items + wpb - 1) / wpb;
            if (blocks > resident_blocks) blocks = resident_blocks;
            if (tracked_kernels()) k_exec_staged<Tracked<K>><<<blocks, wpb * 32, smem, stream>>>(tracked_ptrs(P), a, n_items, n_lists, grouped);
            else k_exec_staged<K><<<blocks, wpb * 32, smem, stream>>>(P, a, n_items, n_lists, grouped);
        }
        else {
            if (tracked_kernels()) {
                k_exec_inplace<Tracked<K>><<<blocks, wpb * 32, 0, stream>>>(tracked_ptrs(P), a, wpb, reinterpret_cast<SysState<Tracked<K>>*>(d_recompute_tmp));
            }
            else k_exec_inplace<K><<<blocks, wpb * 32, 0, stream>>>(P, a, wpb, d_recompute_tmp);
        }
        if (chk(cudaGetLastError())) return fail(dev_err());
        launches++;
        if (sync && dev_sync(stream)) return fail(dev_err());
        return 0;
#endif
    }

    // the Tracked<K> instantiation runs while the trackers are on (with LDO_ONE_KERNEL / LDO_NO_TRACKERS: never)
    bool tracked_kernels() const {
#if defined(LDO_ONE_KERNEL) || defined(LDO_NO_TRACKERS)
        return false;
#else
        return P.track != nullptr;
#endif
    }
    int enable_trackers(bool on) override {
        if (on && !P.track) {
            if (dev_malloc((void**)&P.track, sizeof(TrackStats) * R)) return fail(dev_err());
#ifdef LDO_HOSTSIM
            memset(P.track, 0, sizeof(TrackStats) * R);
#else
            if (chk(cudaMemsetAsync(P.track, 0, sizeof(TrackStats) * R, stream))) return fail(dev_err());
#endif
        }
        else if (!on && P.track) {
            if (dev_sync(stream)) return fail(dev_err());
            dev_free(P.track);
            P.track = nullptr;
        }
        return 0;
    }
    int get_trackers(int replica, TrackStats* out) override {
        if (!P.track) return fail("move trackers are not enabled");
        if (dev_d2h(out, &P.track[replica], sizeof(TrackStats), stream) || dev_sync(stream)) return fail(dev_err());
        return 0;
    }

    // One growthpoint set of the exact enumeration: every replica slot is a worker (ldo_enum.cuh)
    EnumJob* d_enum_job = nullptr;
    EnumAcc* d_enum_acc = nullptr;
    bool enum_stack_raised = false;
    size_t enum_old_stack = 0;
    int enumerate(EnumJob& job, std::vector<EnumAcc>& accs) override {
        if (!P.tables) return fail("temperature tables not set");
        job.n_workers = R;
        // (16 KB per worker: sized once and reused; only the headers are cleared - entries beyond n_keys are never read)
        if ((int)accs.size() != R) accs.resize(R);
        for (EnumAcc& a: accs) memset(&a, 0, offsetof(EnumAcc, keys));
#ifdef LDO_HOSTSIM
        for (int r = 0; r < R; r++) {
            SysState<K>* tmp = STAGED ? d_recompute_tmp : &d_recompute_tmp[r];
            memcpy(tmp, &P.states[r], sizeof(SysState<K>));
            RepAux aux = P.aux[r];
            MoveScratch<K>* ms = STAGED ? P.scratch : &P.scratch[r];
            Engine<K> eng;
            rep_enumerate<K>(&eng, tmp, ms, &aux, P, &job, accs.data(), r);
        }
        return 0;
#else
        if constexpr (!STAGED) {
            return fail("exact enumeration runs on systems that fit the shared-memory replica (snodin class and smaller)");
        }
        else {
            if (upload_constants()) return -1;
            if (!d_enum_job && dev_malloc((void**)&d_enum_job, sizeof(EnumJob))) return fail(dev_err());
            if (!d_enum_acc && dev_malloc((void**)&d_enum_acc, sizeof(EnumAcc) * R)) return fail(dev_err());
            if (dev_h2d(d_enum_job, &job, sizeof(EnumJob), stream)) return fail(dev_err());
            if (chk(cudaMemset2DAsync(d_enum_acc, sizeof(EnumAcc), 0, offsetof(EnumAcc, keys), R, stream))) return fail(dev_err());
            // The recursion is three frames per placed domain deep. Changing the limit reallocates the local memory of
            // the whole device, so it is raised when a job needs more and stays raised while the engine lives.
            size_t cur_stack = 0, want = 2048 + 1024 * (size_t)job.n_stack;
            if (chk(cudaDeviceGetLimit(&cur_stack, cudaLimitStackSize))) return fail(dev_err());
            if (want > cur_stack) {
                if (!enum_stack_raised) enum_old_stack = cur_stack;
                if (dev_sync(stream) || chk(cudaDeviceSetLimit(cudaLimitStackSize, want))) return fail(dev_err());
                enum_stack_raised = true;
            }
            int wpb = warps_per_block;
            size_t smem = sizeof(WarpSmem<K>) * wpb;
            if (chk(cudaFuncSetAttribute(k_enum<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return fail(dev_err());
            k_enum<K><<<(R + wpb - 1) / wpb, wpb * 32, smem, stream>>>(P, d_enum_job, d_enum_acc, R);
            if (chk(cudaGetLastError())) return fail(dev_err());
            launches++;
            // the workers' tables are mostly empty: headers first, then only as many entries as the fullest table holds
            const size_t head = offsetof(EnumAcc, keys);
            if (chk(cudaMemcpy2DAsync(accs.data(), sizeof(EnumAcc), d_enum_acc, sizeof(EnumAcc), head, R, cudaMemcpyDeviceToHost, stream)) || dev_sync(stream)) {
                return fail(dev_err());
            }
            int max_keys = 0;
            for (const EnumAcc& a: accs) max_keys = a.n_keys > max_keys ? a.n_keys : max_keys;
            if (max_keys > 0) {
                char* h = reinterpret_cast<char*>(accs.data());
                char* d = reinterpret_cast<char*>(d_enum_acc);
                if (chk(cudaMemcpy2DAsync(h + offsetof(EnumAcc, keys), sizeof(EnumAcc), d + offsetof(EnumAcc, keys), sizeof(EnumAcc),
                                          sizeof(int) * LDO_ENUM_MAX_OPS * max_keys, R, cudaMemcpyDeviceToHost, stream)) ||
                    chk(cudaMemcpy2DAsync(h + offsetof(EnumAcc, w), sizeof(EnumAcc), d + offsetof(EnumAcc, w), sizeof(EnumAcc), sizeof(double) * max_keys, R,
                                          cudaMemcpyDeviceToHost, stream)) ||
                    dev_sync(stream)) {
                    return fail(dev_err());
                }
            }
            return 0;
        }
#endif
    }

    // The system / moveset / order-parameter descriptions live in constant memory, which is shared by
    // every engine of the process: re-upload when another engine (or a new description) owns it.
    int upload_constants() {
#ifndef LDO_HOSTSIM
        int dslot = device % LDO_MAX_DEVICES;
        if (g_const_owner[dslot] == this && g_const_version[dslot] == const_version) return 0;
        if (chk(cudaDeviceSynchronize())) return fail(dev_err());
        if (chk(cudaMemcpyToSymbol(ldo_c_sc, &shared.sc, sizeof(SysConst)))) return fail(dev_err());
        if (chk(cudaMemcpyToSymbol(ldo_c_ms, &shared.ms, sizeof(MoveSet)))) return fail(dev_err());
        if (chk(cudaMemcpyToSymbol(ldo_c_ob, &shared.ob, sizeof(OpsBiasConst)))) return fail(dev_err());
        if (chk(cudaDeviceSynchronize())) return fail(dev_err());
        g_const_owner[dslot] = this;
        g_const_version[dslot] = const_version;
#endif
        return 0;
    }

    int configure_launch() {
#ifndef LDO_HOSTSIM
        if constexpr (STAGED) {
            // as many warps per block as fit the opt-in shared memory limit, capped at 4 (128 threads)
            int max_smem = 0;
            cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
            size_t per_warp = sizeof(WarpSmem<K>);
            int wpb = (int)(max_smem / per_warp);
            if (wpb < 1) return fail("replica state does not fit shared memory");
            if (wpb > LDO_BLOCK_WARPS) wpb = LDO_BLOCK_WARPS;
            warps_per_block = wpb;
            if (chk(cudaFuncSetAttribute(k_exec_staged<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)))) {
                return fail(dev_err());
            }
#if !defined(LDO_ONE_KERNEL) && !defined(LDO_NO_TRACKERS)
            static_assert(sizeof(WarpSmem<Tracked<K>>) == sizeof(WarpSmem<K>), "Tracked<K> must not change the shared-memory layout");
            if (chk(cudaFuncSetAttribute(k_exec_staged<Tracked<K>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)))) {
                return fail(dev_err());
            }
#endif
#ifdef LDO_SMEM_WINDOW_BASE
            {
                unsigned* d_base = nullptr;
                unsigned h_base = 0;
                if (chk(cudaFuncSetAttribute(k_probe_smem_base, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)))) return fail(dev_err());
                if (dev_malloc((void**)&d_base, sizeof(unsigned))) return fail(dev_err());
                k_probe_smem_base<<<1, wpb * 32, per_warp * wpb, stream>>>(d_base);
                if (chk(cudaGetLastError()) || dev_d2h(&h_base, d_base, sizeof(unsigned), stream) || dev_sync(stream)) return fail(dev_err());
                dev_free(d_base);
                if (h_base != LDO_SMEM_WINDOW_BASE) return fail("dynamic shared memory does not start at the window offset the kernels were built for");
            }
#endif
            int per_sm = 0;
            if (chk(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_exec_staged<K>, wpb * 32, per_warp * wpb))) return fail(dev_err());
#if !defined(LDO_ONE_KERNEL) && !defined(LDO_NO_TRACKERS)
            {
                int per_sm_t = 0;
                if (chk(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, k_exec_staged<Tracked<K>>, wpb * 32, per_warp * wpb))) return fail(dev_err());
                if (per_sm_t < per_sm) per_sm = per_sm_t; // one persistent grid size for both instantiations
            }
#endif
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
            if (const char* om = getenv("LDO_ORDER_MODE")) order_grouped = strcmp(om, "grouped") == 0;
            // profiling knob (profiles/sweep_occupancy.py): fewer resident blocks per SM than fit
            if (const char* cap = getenv("LDO_MAX_BLOCKS_PER_SM")) {
                int c = atoi(cap);
                if (c >= 1 && c < per_sm) per_sm = c;
            }
            resident_blocks = per_sm * n_sm;
            if (resident_blocks < 1) return fail("staged kernel does not fit on the device");
        }
#endif
        return 0;
    }

    // Opaque checkpoint blobs, one per replica, contiguous: [SysState][RepAux][grid values][grid visits]
    // (the grid part only when a Grid bias is configured). A blob is self-contained: the tape pointer is
    // cleared (tapes are not part of a checkpoint) and the replica's grid-bias slice travels with it, so a
    // blob saved from replica i can be loaded into replica j of another engine of the same description
    // (on load the replica owns grid slot j again, whatever window exchange had relabelled before).
    long long* run_timing_ptr() override { return P.run_timing; }
    size_t grid_blob_bytes() const { return shared.has_grid && P.grid_vals ? (sizeof(double) + sizeof(long long)) * LDO_GRID_CAP : 0; }
    size_t blob_size() override { return sizeof(SysState<K>) + sizeof(RepAux) + grid_blob_bytes(); }
    size_t state_bytes() override { return sizeof(SysState<K>); }
    unsigned char* d_blob_stage = nullptr;
    size_t blob_stage_cap = 0;
    int ensure_blob_stage(size_t bytes) {
        if (bytes <= blob_stage_cap) return 0;
        dev_free(d_blob_stage);
        dev_free(d_exch_tape);
        dev_free(d_exch_tape_state);
        dev_free(d_exch_tape_offsets);
        d_blob_stage = nullptr;
        blob_stage_cap = 0;
        if (dev_malloc((void**)&d_blob_stage, bytes)) return fail(dev_err());
        blob_stage_cap = bytes;
        return 0;
    }
    int get_blobs(int first, int count, void* host) override {
        size_t bs = blob_size();
        if (ensure_blob_stage(bs * count)) return -1;
        BlobArgs<K> b {P.states, P.aux, P.grid_vals, P.grid_visits, d_blob_stage, first, count, (unsigned)bs, grid_blob_bytes() != 0};
#ifdef LDO_HOSTSIM
        for (int i = 0; i < count; i++) blob_pack_one(b, i);
        memcpy(host, d_blob_stage, bs * count);
#else
        k_blob_pack<K><<<(count + 3) / 4, 128, 0, stream>>>(b);
        if (chk(cudaGetLastError())) return fail(dev_err());
        launches++;
        if (chk(cudaMemcpyAsync(host, d_blob_stage, bs * count, cudaMemcpyDeviceToHost, stream))) return fail(dev_err());
        if (dev_sync(stream)) return fail(dev_err());
#endif
        return 0;
    }
    int put_blobs(int first, int count, const void* host) override {
        size_t bs = blob_size();
        if (ensure_blob_stage(bs * count)) return -1;
        BlobArgs<K> b {P.states, P.aux, P.grid_vals, P.grid_visits, d_blob_stage, first, count, (unsigned)bs, grid_blob_bytes() != 0};
#ifdef LDO_HOSTSIM
        memcpy(d_blob_stage, host, bs * count);
        for (int i = 0; i < count; i++) blob_unpack_one(b, i);
#else
        if (chk(cudaMemcpyAsync(d_blob_stage, host, bs * count, cudaMemcpyHostToDevice, stream))) return fail(dev_err());
        k_blob_unpack<K><<<(count + 3) / 4, 128, 0, stream>>>(b);
        if (chk(cudaGetLastError())) return fail(dev_err());
        launches++;
#endif
        // the tapes of the overwritten replicas are gone with their aux records
        for (int i = 0; i < count; i++) {
            if (tape_bufs[first + i]) {
#ifndef LDO_HOSTSIM
                if (dev_sync(stream)) return fail(dev_err());
#endif
                dev_free(tape_bufs[first + i]);
                tape_bufs[first + i] = nullptr;
            }
        }
        return 0;
    }

    int get_aux(int first, int count, RepAux* out) override {
        if (dev_d2h(out, P.aux + first, sizeof(RepAux) * count, stream)) return fail(dev_err());
        return 0;
    }
    int put_aux(int first, int count, const RepAux* in) override {
        if (dev_h2d(P.aux + first, in, sizeof(RepAux) * count, stream)) return fail(dev_err());
        return 0;
    }
    int get_state_raw(int replica, std::vector<unsigned char>& blob) override {
        blob.resize(sizeof(SysState<K>));
        if (dev_d2h(blob.data(), &P.states[replica], sizeof(SysState<K>), stream)) return fail(dev_err());
        return 0;
    }
    void capacity(int* c, int* d) override {
        *c = K::C;
        *d = K::D;
    }
    int decode_state(int replica, int* n_chains, int* ci, int* cid, int* cl, int* pos, int* ore, int* st, int* bd) override {
        std::vector<unsigned char> blob;
        if (get_state_raw(replica, blob)) return -1;
        const SysState<K>* s = reinterpret_cast<const SysState<K>*>(blob.data());
        const SysConst& sc = shared.sc;
        *n_chains = s->n_chains;
        int k = 0;
        for (int w = 0; w < s->n_chains; w++) {
            int c = s->order[w];
            ci[w] = s->chain_uid[c];
            cid[w] = s->chain_type[c];
            cl[w] = s->chain_len[c];
            int base = c == 0 ? 0 : sc.n_scaffold + (c - 1) * sc.lmax;
            for (int i = 0; i < s->chain_len[c]; i++) {
                const DomRec& r = s->dom[base + i];
                pos[3 * k] = vx(rec_pos(r));
                pos[3 * k + 1] = vy(rec_pos(r));
                pos[3 * k + 2] = vz(rec_pos(r));
                V3 o = r.ore == ORE_ZERO ? v3(0, 0, 0) : ore_vec(r.ore);
                ore[3 * k] = vx(o);
                ore[3 * k + 1] = vy(o);
                ore[3 * k + 2] = vz(o);
                st[k] = r.state;
                int b = s->bound[base + i];
                if (b >= 0 && r.state != ST_UNASSIGNED && r.state != ST_UNBOUND) {
                    bd[2 * k] = s->chain_uid[s->dchain[b]];
                    bd[2 * k + 1] = s->dindex[b];
                }
                else {
                    bd[2 * k] = -1;
                    bd[2 * k + 1] = -1;
                }
                k++;
            }
        }
        return 0;
    }

    int load_config(int replica, int n_chains, const int* ci, const int* cid, const int* cl, const int* pos, const int* ore) {
        int nd = 0;
        for (int i = 0; i < n_chains; i++) nd += cl[i];
        const SysConst& sc = shared.sc;
        if (n_chains < 1 || n_chains > K::C) return fail("too many chains for this engine's capacity");
        if (cid[0] != 0 || cl[0] != sc.n_scaffold) return fail("first chain must be the scaffold");
        for (int i = 1; i < n_chains; i++) {
            if (cid[i] < 1 || cid[i] >= sc.n_types) return fail("chain identity out of range");
            if (cl[i] != sc.type_len[cid[i]]) return fail("chain length does not match its identity");
        }
        size_t ints = (size_t)3 * n_chains + (size_t)6 * nd;
        if (ints > cfg_cap) {
            dev_free(d_cfg);
            d_cfg = nullptr;
            if (dev_malloc((void**)&d_cfg, sizeof(int) * ints)) return fail(dev_err());
            cfg_cap = ints;
        }
        std::vector<int> h(ints);
        memcpy(&h[0], ci, sizeof(int) * n_chains);
        memcpy(&h[n_chains], cid, sizeof(int) * n_chains);
        memcpy(&h[2 * n_chains], cl, sizeof(int) * n_chains);
        memcpy(&h[3 * n_chains], pos, sizeof(int) * 3 * nd);
        memcpy(&h[3 * n_chains + 3 * nd], ore, sizeof(int) * 3 * nd);
        if (dev_h2d(d_cfg, h.data(), sizeof(int) * ints, stream)) return fail(dev_err());
        OpArgs a;
        memset(&a, 0, sizeof(a));
        a.op = OP_LOAD_CONFIG;
        a.only_replica = replica;
        a.cfg_keep_bias = keep_bias_on_load;
        a.cfg_n_chains = n_chains;
        a.cfg_chain_index = d_cfg;
        a.cfg_chain_ident = d_cfg + n_chains;
        a.cfg_chain_len = d_cfg + 2 * n_chains;
        a.cfg_pos = d_cfg + 3 * n_chains;
        a.cfg_ore = d_cfg + 3 * n_chains + 3 * nd;
        return exec(a, true);
    }

    int set_grid(int replica, int bias, const int* lo, const int* n, const double* vals) override {
        if (!shared.has_grid || !P.grid_vals) return fail("no Grid bias configured");
        if (bias < 0 || bias >= shared.ob.n_biases || shared.ob.biases[bias].type != BIAS_GRID) return fail("not a Grid bias");
        RepAux aux;
        if (get_aux(replica, 1, &aux)) return -1;
        // offsets: grid biases are packed in bias order
        int off = 0;
        for (int b = 0; b < bias; b++) {
            if (shared.ob.biases[b].type != BIAS_GRID || aux.bs.grid_off[b] < 0) continue;
            int sz = 1;
            for (int k = 0; k < shared.ob.biases[b].n_ops; k++) sz *= aux.bs.grid_n[b][k];
            off = aux.bs.grid_off[b] + sz;
        }
        int sz = 1;
        for (int k = 0; k < shared.ob.biases[bias].n_ops; k++) {
            aux.bs.grid_lo[bias][k] = lo[k];
            aux.bs.grid_n[bias][k] = n[k];
            sz *= n[k];
        }
        if (off + sz > LDO_GRID_CAP) return fail("grid bias exceeds LDO_GRID_CAP points");
        aux.bs.grid_off[bias] = off;
        if (put_aux(replica, 1, &aux)) return -1;
        if (dev_h2d(P.grid_vals + (size_t)aux.bs.grid_slot * LDO_GRID_CAP + off, vals, sizeof(double) * sz, stream)) return fail(dev_err());
        return 0;
    }
    int get_visits(int replica, int bias, long long* counts, int clear) override {
        if (!shared.has_grid || !P.grid_visits) return fail("no Grid bias configured");
        if (bias < 0 || bias >= shared.ob.n_biases || shared.ob.biases[bias].type != BIAS_GRID) return fail("not a Grid bias");
        RepAux aux;
        if (get_aux(replica, 1, &aux)) return -1;
        int off = aux.bs.grid_off[bias];
        if (off < 0) return fail("grid not set for this replica");
        int sz = 1;
        for (int k = 0; k < shared.ob.biases[bias].n_ops; k++) sz *= aux.bs.grid_n[bias][k];
        long long* p = P.grid_visits + (size_t)aux.bs.grid_slot * LDO_GRID_CAP + off;
        if (dev_d2h(counts, p, sizeof(long long) * sz, stream)) return fail(dev_err());
        if (clear) {
            if (dev_memset(p, 0, sizeof(long long) * sz, stream)) return fail(dev_err());
            if (dev_sync(stream)) return fail(dev_err());
        }
        return 0;
    }

    int attach_tape(int replica, const ldo_tape_draw* draws, long long n) override {
        static_assert(sizeof(ldo_tape_draw) == sizeof(TapeDraw), "tape layout");
        RepAux aux;
        if (get_aux(replica, 1, &aux)) return -1;
        // free the buffer this replica is reading NOW: window exchange hands tapes from replica to replica
        for (size_t j = 0; j < tape_bufs.size(); j++) {
            if (tape_bufs[j] != nullptr && tape_bufs[j] == (const void*)aux.rng.tape) {
                dev_free(tape_bufs[j]);
                tape_bufs[j] = nullptr;
            }
        }
        if (tape_bufs[replica] != nullptr) {
            // still referenced by another replica after a swap: park it, it is freed with the engine
            parked_tapes.push_back(tape_bufs[replica]);
            tape_bufs[replica] = nullptr;
        }
        aux.rng.tape = nullptr;
        aux.rng.tape_len = 0;
        aux.rng.tape_pos = 0;
        if (n > 0) {
            void* p = nullptr;
            if (dev_malloc(&p, sizeof(TapeDraw) * n)) return fail(dev_err());
            if (dev_h2d(p, draws, sizeof(TapeDraw) * n, stream)) return fail(dev_err());
            tape_bufs[replica] = p;
            aux.rng.tape = (const TapeDraw*)p;
            aux.rng.tape_len = n;
        }
        return put_aux(replica, 1, &aux);
    }

    int ensure_exchange(int n_global, int n_slots) {
        int nq = LDO_DEP_FIXED + (shared.sc.n_types - 1);
        if (n_global > dep_all_n) {
            dev_free(d_dep_all);
            d_dep_all = nullptr;
            if (dev_malloc((void**)&d_dep_all, sizeof(double) * nq * n_global)) return fail(dev_err());
            dep_all_n = n_global;
        }
        if (n_slots > exch_cap) {
            dev_free(d_q2r);
            dev_free(d_att);
            dev_free(d_acc);
            dev_free(d_slot_tidx);
            dev_free(d_slot_vals);
            d_q2r = nullptr;
            d_att = d_acc = nullptr;
            d_slot_tidx = nullptr;
            d_slot_vals = nullptr;
            if (dev_malloc((void**)&d_q2r, sizeof(int) * n_slots)) return fail(dev_err());
            if (dev_malloc((void**)&d_att, sizeof(long long) * n_slots)) return fail(dev_err());
            if (dev_malloc((void**)&d_acc, sizeof(long long) * n_slots)) return fail(dev_err());
            if (dev_malloc((void**)&d_slot_tidx, sizeof(int) * n_slots)) return fail(dev_err());
            if (dev_malloc((void**)&d_slot_vals, sizeof(double) * 4 * n_slots)) return fail(dev_err());
            exch_cap = n_slots;
            slot_cache.clear();
            slot_tidx_cache.clear();
            exch_resident_slots = 0;
        }
        if (!d_red_u) {
            int nst = shared.sc.n_types - 1;
            if (dev_malloc((void**)&d_red_u, sizeof(double) * (nst > 0 ? nst : 1))) return fail(dev_err());
            if (nst > 0 && dev_h2d(d_red_u, reduced_staple_u.data(), sizeof(double) * nst, stream)) return fail(dev_err());
        }
        return 0;
    }
    double* d_exch_tape = nullptr;
    long long* d_exch_tape_state = nullptr;
    long long* d_exch_tape_offsets = nullptr;
    long long exch_tape_rounds = 0;
    long long exch_tape_total = 0;
    int set_exchange_tape(const double* reals, long long n, const long long* offsets, long long n_rounds) override {
        dev_free(d_exch_tape);
        dev_free(d_exch_tape_offsets);
        d_exch_tape = nullptr;
        d_exch_tape_offsets = nullptr;
        exch_tape_rounds = 0;
        if (n_rounds <= 0) return 0;
        if (offsets[0] != 0 || offsets[n_rounds] != n) return fail("exchange tape offsets must run from 0 to n");
        if (!d_exch_tape_state && dev_malloc((void**)&d_exch_tape_state, sizeof(long long) * 4)) return fail(dev_err());
        if (dev_malloc((void**)&d_exch_tape, sizeof(double) * (n > 0 ? n : 1))) return fail(dev_err());
        if (dev_malloc((void**)&d_exch_tape_offsets, sizeof(long long) * (n_rounds + 1))) return fail(dev_err());
        if (n > 0 && dev_h2d(d_exch_tape, reals, sizeof(double) * n, stream)) return fail(dev_err());
        if (dev_h2d(d_exch_tape_offsets, offsets, sizeof(long long) * (n_rounds + 1), stream)) return fail(dev_err());
        long long st[4] = {0, 0, 0, 0};
        if (dev_h2d(d_exch_tape_state, st, sizeof(st), stream)) return fail(dev_err());
        exch_tape_rounds = n_rounds;
        exch_tape_total = n;
        return 0;
    }
    int exchange_tape_state(long long* missing, long long* unused) override {
        long long st[4] = {0, 0, 0, 0};
        if (d_exch_tape && dev_d2h(st, d_exch_tape_state, sizeof(st), stream)) return fail(dev_err());
        *missing = st[2];
        *unused = st[3];
        return 0;
    }
    int exchange_buffers(int n_global, void** send, void** recv, int* nq) override {
        if (ensure_exchange(n_global, 1)) return -1;
        *send = d_dependent;
        *recv = d_dep_all;
        *nq = LDO_DEP_FIXED + (shared.sc.n_types - 1);
        return 0;
    }

    int window_exchange(WindowExchangeArgs& x, int* window_to_replica, long long* attempts, long long* accepts) override {
        int n_slots = x.n_ladders * x.n_windows;
        int n_pairs = x.n_ladders * (x.n_windows - 1);
        if (n_slots != R) return fail("n_ladders * n_windows must equal the number of replicas");
        if (ensure_exchange(R, n_slots)) return -1;
        if (dev_h2d(d_q2r, window_to_replica, sizeof(int) * n_slots, stream)) return fail(dev_err());
        if (dev_h2d(d_att, attempts, sizeof(long long) * n_pairs, stream)) return fail(dev_err());
        if (dev_h2d(d_acc, accepts, sizeof(long long) * n_pairs, stream)) return fail(dev_err());
        x.window_to_replica = d_q2r;
        x.attempts = d_att;
        x.accepts = d_acc;
        x.aux = P.aux;
        x.grid_vals = P.grid_vals;
        x.ob = &d_shared->ob;
        x.exchange_tape = d_exch_tape;
        x.exchange_tape_len = d_exch_tape ? exch_tape_total : 0;
        x.exchange_tape_state = d_exch_tape_state;
#ifdef LDO_HOSTSIM
        for (int l = 0; l < x.n_ladders; l++) window_exchange_ladder(x, l);
#else
        int threads = 128;
        k_window_exchange<<<(x.n_ladders + threads - 1) / threads, threads, 0, stream>>>(x);
        if (chk(cudaGetLastError())) return fail(dev_err());
        launches++;
#endif
        if (dev_d2h(window_to_replica, d_q2r, sizeof(int) * n_slots, stream)) return fail(dev_err());
        if (dev_d2h(attempts, d_att, sizeof(long long) * n_pairs, stream)) return fail(dev_err());
        if (dev_d2h(accepts, d_acc, sizeof(long long) * n_pairs, stream)) return fail(dev_err());
        return 0;
    }

    // Slot control variables (x.slot_* are HOST arrays of the caller): uploaded when they change
    std::vector<double> slot_cache;
    std::vector<int> slot_tidx_cache;
    int upload_slots(const ExchangeArgs& x) {
        int L = x.ladder_len;
        std::vector<double> sv(4 * (size_t)L);
        memcpy(&sv[0], x.slot_temp, sizeof(double) * L);
        memcpy(&sv[L], x.slot_staple_u_mult, sizeof(double) * L);
        memcpy(&sv[2 * L], x.slot_bias_mult, sizeof(double) * L);
        memcpy(&sv[3 * L], x.slot_stacking_mult, sizeof(double) * L);
        std::vector<int> ti(x.slot_temp_idx, x.slot_temp_idx + L);
        if (sv == slot_cache && ti == slot_tidx_cache) return 0;
        if (dev_h2d(d_slot_tidx, ti.data(), sizeof(int) * L, stream)) return fail(dev_err());
        if (dev_h2d(d_slot_vals, sv.data(), sizeof(double) * 4 * L, stream)) return fail(dev_err());
        slot_cache = sv;
        slot_tidx_cache = ti;
        return 0;
    }
    // The exchange kernel on the device-resident map / counters and the gathered records already in d_dep_all
    // (or, single GPU, the local ones); nothing is copied to the host and nothing synchronises
    int exchange_resident(ExchangeArgs& x) override {
        int n_slots = x.n_ladders * x.ladder_len;
        if (ensure_exchange(x.n_global, 2 * n_slots)) return -1;
        if (exch_resident_slots != n_slots) return fail("ldo_exchange_state_set not called for this ladder shape");
        if (x.n_global == R) {
            int nq = LDO_DEP_FIXED + x.n_staple_types;
#ifdef LDO_HOSTSIM
            memcpy(d_dep_all, d_dependent, sizeof(double) * nq * R);
#else
            if (chk(cudaMemcpyAsync(d_dep_all, d_dependent, sizeof(double) * nq * R, cudaMemcpyDeviceToDevice, stream))) return fail(dev_err());
#endif
        }
        if (upload_slots(x)) return -1;
        return exchange_launch(x);
    }
    int exch_resident_slots = 0;
    int exchange_state_set(int n_slots, int n_counters, const int* q2r, const long long* att, const long long* acc) override {
        if (ensure_exchange(1, 2 * n_slots)) return -1;
        if (n_counters > 2 * n_slots) return fail("too many exchange counters");
        if (dev_h2d(d_q2r, q2r, sizeof(int) * n_slots, stream)) return fail(dev_err());
        if (dev_h2d(d_att, att, sizeof(long long) * n_counters, stream)) return fail(dev_err());
        if (dev_h2d(d_acc, acc, sizeof(long long) * n_counters, stream)) return fail(dev_err());
        exch_resident_slots = n_slots;
        return 0;
    }
    int exchange_state_get(int n_slots, int n_counters, int* q2r, long long* att, long long* acc) override {
        if (exch_resident_slots != n_slots) return fail("no resident exchange state of this shape");
        if (q2r && dev_d2h(q2r, d_q2r, sizeof(int) * n_slots, stream)) return fail(dev_err());
        if (att && dev_d2h(att, d_att, sizeof(long long) * n_counters, stream)) return fail(dev_err());
        if (acc && dev_d2h(acc, d_acc, sizeof(long long) * n_counters, stream)) return fail(dev_err());
        return 0;
    }
    int exchange_launch(ExchangeArgs& x) {
        int L = x.ladder_len;
        ExchangeArgs dx = x;
        dx.dependent = d_dep_all;
        dx.slot_to_replica = d_q2r;
        dx.attempts = d_att;
        dx.accepts = d_acc;
        dx.slot_temp_idx = d_slot_tidx;
        dx.slot_temp = d_slot_vals;
        dx.slot_staple_u_mult = d_slot_vals + L;
        dx.slot_bias_mult = d_slot_vals + 2 * L;
        dx.slot_stacking_mult = d_slot_vals + 3 * L;
        dx.reduced_staple_u = d_red_u;
        dx.aux = P.aux;
        dx.exchange_tape = d_exch_tape;
        dx.exchange_tape_offsets = d_exch_tape_offsets;
        dx.exchange_tape_rounds = exch_tape_rounds;
        dx.exchange_tape_state = d_exch_tape_state;
        if (d_exch_tape && x.n_ladders != 1) return fail("an exchange tape replays a single ladder");
#ifdef LDO_HOSTSIM
        for (int l = 0; l < x.n_ladders; l++) exchange_ladder(dx, l);
#else
        int threads = 128;
        k_exchange<<<(x.n_ladders + threads - 1) / threads, threads, 0, stream>>>(dx);
        if (chk(cudaGetLastError())) return fail(dev_err());
        launches++;
#endif
        return 0;
    }
    int exchange(ExchangeArgs& x, const double* dependent_host, int* slot_to_replica, long long* attempts, long long* accepts) override {
        int n_slots = x.n_ladders * x.ladder_len;
        int n_pairs = x.variant == LDO_PT_2D ? 2 * n_slots : x.n_ladders * (x.ladder_len - 1);
        int nq = LDO_DEP_FIXED + x.n_staple_types;
        if (ensure_exchange(x.n_global, 2 * n_slots)) return -1;
        if (dependent_host) {
            if (dev_h2d(d_dep_all, dependent_host, sizeof(double) * nq * x.n_global, stream)) return fail(dev_err());
        }
        else if (x.n_global == R) {
            // single-GPU: the local dependent quantities are the global ones
#ifdef LDO_HOSTSIM
            memcpy(d_dep_all, d_dependent, sizeof(double) * nq * R);
#else
            if (chk(cudaMemcpyAsync(d_dep_all, d_dependent, sizeof(double) * nq * R, cudaMemcpyDeviceToDevice, stream))) return fail(dev_err());
#endif
        }
        if (exchange_state_set(n_slots, n_pairs, slot_to_replica, attempts, accepts)) return -1;
        if (upload_slots(x)) return -1;
        if (exchange_launch(x)) return -1;
        return exchange_state_get(n_slots, n_pairs, slot_to_replica, attempts, accepts);
    }
};

struct ldo_engine {
    EngineBase* b;
    int (*load_config)(ldo_engine*, int, int, const int*, const int*, const int*, const int*, const int*);
    unsigned long long seed;
};

template <class K, bool STAGED>
static int load_config_thunk(ldo_engine* e, int replica, int n, const int* ci, const int* cid, const int* cl, const int* pos, const int* ore) {
    return static_cast<EngineImpl<K, STAGED>*>(e->b)->load_config(replica, n, ci, cid, cl, pos, ore);
}

template <class K, bool STAGED>
static int make_engine(const ldo_system_desc* d, const SysConst& sc, int n_replicas, int device, ldo_engine** out) {
    auto* impl = new EngineImpl<K, STAGED>();
    impl->device = device;
    impl->shared.sc = sc;
    // staple chemical potentials / T (origami_system.cpp:965-990)
    for (int t = 1; t < sc.n_types; t++) {
        impl->reduced_staple_u.push_back(log(d->staple_M) - (2.0 * sc.type_len[t] - 1) * log(6.0));
    }
    if (impl->init(n_replicas) || impl->configure_launch()) {
        g_create_error = impl->err;
        delete impl;
        return -1;
    }
    ldo_engine* e = new ldo_engine();
    e->b = impl;
    e->load_config = &load_config_thunk<K, STAGED>;
    e->seed = 0;
    *out = e;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------

extern "C" {

int ldo_engine_create(const ldo_system_desc* d, int n_replicas, int device, ldo_engine** out) {
    if (!d || !out || n_replicas < 1) {
        g_create_error = "bad arguments";
        return -1;
    }
    if (d->n_types < 1 || d->n_types > LDO_MAX_TYPES) {
        g_create_error = "too many chain identities";
        return -1;
    }
    SysConst sc;
    memset(&sc, 0, sizeof(sc));
    sc.n_types = d->n_types;
    int off = 0, n_ident = 0, max_staple_len = 0;
    for (int t = 0; t < d->n_types; t++) {
        sc.type_len[t] = d->type_len[t];
        sc.type_off[t] = off;
        for (int i = 0; i < d->type_len[t]; i++) {
            if (off + i >= LDO_MAX_IDENTS) {
                g_create_error = "too many domain identities";
                return -1;
            }
            int id = d->idents[off + i];
            sc.idents[off + i] = (short)id;
            if (abs(id) > n_ident) n_ident = abs(id);
        }
        off += d->type_len[t];
        if (t > 0 && d->type_len[t] > max_staple_len) max_staple_len = d->type_len[t];
    }
    sc.n_scaffold = d->type_len[0];
    sc.lmax = d->max_staple_size > max_staple_len ? d->max_staple_size : max_staple_len;
    if (sc.lmax < 1) sc.lmax = 1;
    sc.cyclic = d->cyclic;
    sc.domain_type = d->domain_type;
    sc.misbinding_pot = d->misbinding_pot;
    sc.apply_mean_field_cor = d->apply_mean_field_cor;
    sc.max_total_staples = d->max_total_staples;
    sc.max_type_staples = d->max_type_staples;
    sc.n_ident = n_ident;
    sc.staple_M = d->staple_M;
    sc.stacking_ene = d->stacking_ene;
    int need_d = sc.n_scaffold + d->max_total_staples * sc.lmax;
    int need_c = 1 + d->max_total_staples;
    if (need_d <= CapsSmall::D && need_c <= CapsSmall::C && d->n_types <= CapsSmall::T) {
        return make_engine<CapsSmall, LDO_SMALL_STAGED>(d, sc, n_replicas, device, out);
    }
    if (need_d <= CapsLarge::D && need_c <= CapsLarge::C && d->n_types <= CapsLarge::T) {
        return make_engine<CapsLarge, false>(d, sc, n_replicas, device, out);
    }
    g_create_error = "system exceeds the compiled capacities (domains/chains/types)";
    return -1;
}

void ldo_engine_destroy(ldo_engine* e) {
    if (!e) return;
    delete e->b;
    delete e;
}

const char* ldo_last_error(const ldo_engine* e) { return e ? e->b->err.c_str() : g_create_error.c_str(); }
int ldo_num_replicas(const ldo_engine* e) { return e->b->R; }
int ldo_set_step(ldo_engine* e, long long step) {
    EngineBase* b = e->b;
    std::vector<RepAux> aux(b->R);
    if (b->get_aux(0, b->R, aux.data())) return -1;
    for (int r = 0; r < b->R; r++) aux[r].step = step;
    return b->put_aux(0, b->R, aux.data());
}
int ldo_get_reduced_staple_u(ldo_engine* e, double* out) {
    for (size_t t = 0; t < e->b->reduced_staple_u.size(); t++) out[t] = e->b->reduced_staple_u[t];
    return (int)e->b->reduced_staple_u.size();
}

int ldo_set_temperature_tables(ldo_engine* e, int n_temps, int n_ident, const double* temps, const double* hyb_energy,
                               const double* hyb_enthalpy, const double* hyb_entropy, const double* init) {
    if (n_ident < e->b->shared.sc.n_ident) return e->b->fail("n_ident smaller than the largest domain identity");
    return e->b->set_tables(n_temps, n_ident, temps, hyb_energy, hyb_enthalpy, hyb_entropy, init);
}

int ldo_set_moveset(ldo_engine* e, int n, const ldo_movetype_desc* mts, int allow_nonsensical_ps) {
    EngineBase* b = e->b;
    if (n < 1 || n > LDO_MAX_MOVETYPES) return b->fail("bad movetype count");
    MoveSet& ms = b->shared.ms;
    int keep_order = ms.reference_draw_order;
    memset(&ms, 0, sizeof(ms));
    ms.reference_draw_order = keep_order;
    ms.n = n;
    ms.allow_nonsensical_ps = allow_nonsensical_ps;
    double cum = 0;
    int nst = b->shared.sc.n_types - 1;
    int em_off = 0;
    for (int i = 0; i < n; i++) {
        MoveDef& md = ms.mt[i];
        md.type = mts[i].type;
        if (md.type < 0 || md.type > MT_CTRG_LINKER_REGROWTH) {
            return b->fail("movetype not available on device");
        }
        cum += mts[i].freq;
        md.cum_prob = cum;
        md.max_regrowth = mts[i].max_regrowth;
        md.max_seg_regrowth = mts[i].max_seg_regrowth;
        md.max_num_recoils = mts[i].max_num_recoils;
        md.max_c_attempts = mts[i].max_c_attempts;
        md.max_disp = mts[i].max_disp;
        md.max_turns = mts[i].max_turns;
        md.max_linker_length = mts[i].max_linker_length;
        md.num_transforms = mts[i].num_transforms;
        md.adaptive_exchange = mts[i].adaptive_exchange;
        md.exchange_mults_off = 0;
        if (md.type == MT_MET_STAPLE_EXCHANGE) {
            if (em_off + nst > LDO_MAX_TYPES) return b->fail("too many exchange multipliers");
            md.exchange_mults_off = em_off;
            for (int t = 0; t < nst; t++) {
                ms.exchange_mults[em_off + t] = t < mts[i].n_exchange_mults ? mts[i].exchange_mults[t] : 1.0;
            }
            em_off += nst;
        }
        if ((md.type == MT_CTCB_SCAFFOLD_REGROWTH || md.type == MT_CTCB_JUMP_SCAFFOLD_REGROWTH) && md.max_regrowth < 2) {
            return b->fail("CTCB options out of range (max_regrowth >= 2)");
        }
        bool linker = md.type == MT_CTCB_LINKER_REGROWTH || md.type == MT_CTCB_CLUSTERED_LINKER_REGROWTH || md.type == MT_CTRG_LINKER_REGROWTH;
        if (linker && (md.num_transforms < 1 || md.num_transforms > LDO_MAX_TRANSFORMS || md.max_disp < 0 || md.max_turns < 0 ||
                       md.max_linker_length < 1 || (md.type != MT_CTCB_CLUSTERED_LINKER_REGROWTH && md.max_regrowth < 1) ||
                       b->shared.sc.n_scaffold < 3)) {
            return b->fail("linker regrowth options out of range (num_transforms 1..16, max_disp >= 0, max_turns >= 0, max_linker_length >= 1, max_regrowth >= 1, scaffold of 3+ domains)");
        }
        if ((md.type == MT_CTRG_SCAFFOLD_REGROWTH || md.type == MT_CTRG_JUMP_SCAFFOLD_REGROWTH || md.type == MT_CTRG_LINKER_REGROWTH) &&
            (md.max_c_attempts < 1 || md.max_c_attempts > 36 || md.max_regrowth < 2 || md.max_num_recoils < 0 ||
             md.max_num_recoils >= LDO_RG_OWN_SLOTS) && !(md.type == MT_CTRG_LINKER_REGROWTH && md.max_c_attempts >= 1 &&
             md.max_c_attempts <= 36 && md.max_num_recoils >= 0 && md.max_num_recoils < LDO_RG_OWN_SLOTS)) {
            return b->fail("CTRG options out of range (max_c_attempts 1..36, max_regrowth >= 2, max_num_recoils 0..3)");
        }
    }
    // every replica starts its adaptive multipliers from the movetype file's
    {
        std::vector<RepAux> aux(b->R);
        if (b->get_aux(0, b->R, aux.data())) return -1;
        for (int r = 0; r < b->R; r++) memcpy(aux[r].stats.exchange_mults, ms.exchange_mults, sizeof(ms.exchange_mults));
        if (b->put_aux(0, b->R, aux.data())) return -1;
    }
    return b->push_shared();
}

int ldo_get_exchange_mults(ldo_engine* e, int movetype, double* out) {
    EngineBase* b = e->b;
    const MoveSet& ms = b->shared.ms;
    if (movetype < 0 || movetype >= ms.n || ms.mt[movetype].type != MT_MET_STAPLE_EXCHANGE) return b->fail("not a staple-exchange movetype");
    int nst = b->shared.sc.n_types - 1;
    std::vector<RepAux> aux(b->R);
    if (b->get_aux(0, b->R, aux.data())) return -1;
    const MoveDef& md = ms.mt[movetype];
    for (int r = 0; r < b->R; r++) {
        for (int t = 0; t < nst; t++) {
            out[(size_t)r * nst + t] = md.adaptive_exchange ? aux[r].stats.exchange_mults[md.exchange_mults_off + t] : ms.exchange_mults[md.exchange_mults_off + t];
        }
    }
    return 0;
}

int ldo_set_reference_draw_order(ldo_engine* e, int on) {
    e->b->shared.ms.reference_draw_order = on ? 1 : 0;
    return e->b->push_shared();
}

int ldo_set_domain_update_biases(ldo_engine* e, int present) {
    e->b->domain_update_biases = present ? 1 : 0;
    return 0;
}

int ldo_set_order_params(ldo_engine* e, int n, const ldo_order_param_desc* ops) {
    EngineBase* b = e->b;
    if (n < 0 || n > LDO_MAX_OPS) return b->fail("too many order parameters");
    OpsBiasConst& ob = b->shared.ob;
    ob.n_ops = n;
    ob.n_pd_ops = 0;
    // the per-domain hooks exist in OrigamiSystemWithBias only (domain_update_biases_present, origami_system.cpp:1006-1028);
    // without it per-domain parameters are registered but never updated again
    for (int i = 0; i < n && b->domain_update_biases; i++) {
        if ((ops[i].type == OP_DIST || ops[i].type == OP_ADJACENT_SITE) && ops[i].update_per_domain) ob.n_pd_ops++;
    }
    for (int i = 0; i < n; i++) {
        ob.ops[i].type = ops[i].type;
        ob.ops[i].per_domain = 0;
        ob.ops[i].arg = ops[i].staple;
        ob.ops[i].arg2 = 0;
        ob.ops[i].n_sum = 0;
        if (ops[i].type == OP_DIST || ops[i].type == OP_ADJACENT_SITE) {
            int n_scaf = b->shared.sc.n_scaffold;
            if (ops[i].chain1 != 0 || ops[i].chain2 != 0 || ops[i].domain1 < 0 || ops[i].domain1 >= n_scaf ||
                ops[i].domain2 < 0 || ops[i].domain2 >= n_scaf) {
                return b->fail("Dist / AdjacentSite order parameters must refer to scaffold domains (chain 0)");
            }
            ob.ops[i].arg = ops[i].domain1;
            ob.ops[i].arg2 = ops[i].domain2;
            ob.ops[i].per_domain = ops[i].update_per_domain ? 1 : 0;
        }
        else if (ops[i].update_per_domain) {
            return b->fail("only Dist / AdjacentSite order parameters can be updated per domain");
        }
        if (ops[i].type == OP_SUM) {
            if (ops[i].n_sum > LDO_MAX_SUM) return b->fail("Sum order parameter has too many terms");
            for (int k = 0; k < ops[i].n_sum; k++) {
                if (ops[i].sum_ops[k] >= 0 && ops[i].sum_ops[k] < i && ob.ops[ops[i].sum_ops[k]].per_domain) {
                    return b->fail("Sum of per-domain order parameters is not available (its evaluation order is not reproducible)");
                }
            }
            ob.ops[i].n_sum = ops[i].n_sum;
            for (int k = 0; k < ops[i].n_sum; k++) {
                if (ops[i].sum_ops[k] < 0 || ops[i].sum_ops[k] >= i) return b->fail("Sum refers to a later order parameter");
                ob.ops[i].sum_idx[k] = ops[i].sum_ops[k];
            }
        }
        if ((ops[i].type == OP_NUM_STAPLES_TYPE || ops[i].type == OP_STAPLE_TYPE_FULLY_BOUND) &&
            (ops[i].staple < 1 || ops[i].staple >= b->shared.sc.n_types)) {
            return b->fail("order parameter staple identity out of range");
        }
    }
    return b->push_shared();
}

int ldo_set_biases(ldo_engine* e, int n, const ldo_bias_desc* biases) {
    EngineBase* b = e->b;
    if (n < 0 || n > LDO_MAX_BIASES) return b->fail("too many bias functions");
    OpsBiasConst& ob = b->shared.ob;
    ob.n_biases = n;
    b->shared.has_grid = 0;
    std::vector<RepAux> aux(b->R);
    if (b->get_aux(0, b->R, aux.data())) return -1;
    for (int i = 0; i < n; i++) {
        BiasDef& bd = ob.biases[i];
        bd.type = biases[i].type;
        bd.n_ops = biases[i].n_ops;
        if (bd.type != BIAS_LINEAR_STEP_WELL && bd.type != BIAS_SQUARE_WELL && bd.type != BIAS_GRID) return b->fail("unknown bias function type");
        if (bd.n_ops < 1 || bd.n_ops > LDO_MAX_GRID_DIM) return b->fail("bias function has a bad number of order parameters");
        if (bd.type != BIAS_GRID && bd.n_ops != 1) return b->fail("well bias functions take exactly one order parameter");
        for (int k = 0; k < bd.n_ops; k++) {
            if (biases[i].ops[k] < 0 || biases[i].ops[k] >= ob.n_ops) return b->fail("bias refers to an unknown order parameter");
            bd.op_idx[k] = biases[i].ops[k];
        }
        bd.min_op = biases[i].min_op;
        bd.max_op = biases[i].max_op;
        bd.well_bias = biases[i].well_bias;
        bd.min_bias = biases[i].min_bias;
        bd.slope = biases[i].slope;
        bd.outside_bias = biases[i].outside_bias;
        if (bd.type == BIAS_GRID) b->shared.has_grid = 1;
        for (int r = 0; r < b->R; r++) {
            aux[r].bs.win_min[i] = bd.min_op;
            aux[r].bs.win_max[i] = bd.max_op;
            aux[r].bs.grid_off[i] = -1;
        }
    }
    if (b->put_aux(0, b->R, aux.data())) return -1;
    return b->push_shared();
}

int ldo_set_window(ldo_engine* e, int replica, int bias, int min_op, int max_op) {
    EngineBase* b = e->b;
    if (replica < 0 || replica >= b->R || bias < 0 || bias >= b->shared.ob.n_biases) return b->fail("bad replica or bias index");
    RepAux aux;
    if (b->get_aux(replica, 1, &aux)) return -1;
    aux.bs.win_min[bias] = min_op;
    aux.bs.win_max[bias] = max_op;
    return b->put_aux(replica, 1, &aux);
}

int ldo_set_grid_bias(ldo_engine* e, int replica, int bias, const int* lo, const int* n, const double* values) {
    if (replica < 0 || replica >= e->b->R) return e->b->fail("bad replica index");
    return e->b->set_grid(replica, bias, lo, n, values);
}

int ldo_get_grid_visits(ldo_engine* e, int replica, int bias, long long* counts, int clear) {
    if (replica < 0 || replica >= e->b->R) return e->b->fail("bad replica index");
    return e->b->get_visits(replica, bias, counts, clear);
}

static int refresh_energy(ldo_engine* e, bool sync = true) {
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.op = OP_UPDATE_ENERGY;
    a.only_replica = -1;
    return e->b->exec(a, sync);
}

int ldo_set_control(ldo_engine* e, int first, int count, const int* temp_idx, const double* staple_u_mult,
                    const double* bias_mult, const double* stacking_mult) {
    EngineBase* b = e->b;
    if (first < 0 || count < 0 || first + count > b->R) return b->fail("bad replica range");
    std::vector<RepAux> aux(count);
    if (b->get_aux(first, count, aux.data())) return -1;
    for (int i = 0; i < count; i++) {
        if (temp_idx) {
            if (temp_idx[i] < 0 || temp_idx[i] >= b->shared.n_temps) return b->fail("temperature index out of range");
            aux[i].ctl.temp_idx = temp_idx[i];
            aux[i].ctl.temp = b->temps[temp_idx[i]];
            // a control variable set from outside is a fresh potential (no table cache history)
            aux[i].init_temp_idx = -1;
            aux[i].table_cache_mask = 0;
        }
        if (staple_u_mult) aux[i].ctl.staple_u_mult = staple_u_mult[i];
        if (bias_mult) aux[i].ctl.bias_mult = bias_mult[i];
        if (stacking_mult) aux[i].ctl.stacking_mult = stacking_mult[i];
    }
    if (b->put_aux(first, count, aux.data())) return -1;
    return refresh_energy(e);
}

int ldo_get_control(ldo_engine* e, int first, int count, int* temp_idx, double* staple_u_mult, double* bias_mult,
                    double* stacking_mult) {
    EngineBase* b = e->b;
    if (first < 0 || count < 0 || first + count > b->R) return b->fail("bad replica range");
    std::vector<RepAux> aux(count);
    if (b->get_aux(first, count, aux.data())) return -1;
    for (int i = 0; i < count; i++) {
        if (temp_idx) temp_idx[i] = aux[i].ctl.temp_idx;
        if (staple_u_mult) staple_u_mult[i] = aux[i].ctl.staple_u_mult;
        if (bias_mult) bias_mult[i] = aux[i].ctl.bias_mult;
        if (stacking_mult) stacking_mult[i] = aux[i].ctl.stacking_mult;
    }
    return 0;
}

int ldo_seed(ldo_engine* e, unsigned long long seed, unsigned int first_subsequence) {
    EngineBase* b = e->b;
    e->seed = seed;
    std::vector<RepAux> aux(b->R);
    if (b->get_aux(0, b->R, aux.data())) return -1;
    for (int r = 0; r < b->R; r++) {
        aux[r].rng.key0 = (uint32_t)seed;
        aux[r].rng.key1 = (uint32_t)(seed >> 32);
        aux[r].rng.subseq = first_subsequence + (uint32_t)r;
        aux[r].rng.stream = 0;
        aux[r].rng.counter = 0;
        aux[r].rng.buf_n = 0;
    }
    return b->put_aux(0, b->R, aux.data());
}

int ldo_seed_subsequences(ldo_engine* e, unsigned long long seed, const unsigned int* subsequences) {
    EngineBase* b = e->b;
    e->seed = seed;
    std::vector<RepAux> aux(b->R);
    if (b->get_aux(0, b->R, aux.data())) return -1;
    for (int r = 0; r < b->R; r++) {
        aux[r].rng.key0 = (uint32_t)seed;
        aux[r].rng.key1 = (uint32_t)(seed >> 32);
        aux[r].rng.subseq = subsequences[r];
        aux[r].rng.stream = 0;
        aux[r].rng.counter = 0;
        aux[r].rng.buf_n = 0;
    }
    return b->put_aux(0, b->R, aux.data());
}

// Philox state of a replica as LDO_RNG_STATE_WORDS numbers: key0, key1, subsequence, stream, counter, number of buffered
// words, the buffered words. What RandomEngineStateOutputFile / RandomEngineStateInputFile carry for the reference's
// mt19937_64 (files.cpp:220-246, 781-793; simulation.cpp:204-212).
int ldo_rng_state_words(void) { return 6 + 4 * LDO_PHILOX_BLOCKS; }

int ldo_get_rng_state(ldo_engine* e, int first, int count, unsigned long long* words) {
    EngineBase* b = e->b;
    if (first < 0 || count < 0 || first + count > b->R) return b->fail("bad replica range");
    std::vector<RepAux> aux(count);
    if (count && b->get_aux(first, count, aux.data())) return -1;
    const int W = ldo_rng_state_words();
    for (int r = 0; r < count; r++) {
        const Rng& g = aux[r].rng;
        unsigned long long* w = words + (size_t)r * W;
        w[0] = g.key0;
        w[1] = g.key1;
        w[2] = g.subseq;
        w[3] = g.stream;
        w[4] = g.counter;
        w[5] = (unsigned long long)g.buf_n;
        for (int k = 0; k < 4 * LDO_PHILOX_BLOCKS; k++) w[6 + k] = g.buf[k];
    }
    return 0;
}

int ldo_set_rng_state(ldo_engine* e, int first, int count, const unsigned long long* words) {
    EngineBase* b = e->b;
    if (first < 0 || count < 0 || first + count > b->R) return b->fail("bad replica range");
    const int W = ldo_rng_state_words();
    for (int r = 0; r < count; r++) {
        const unsigned long long* w = words + (size_t)r * W;
        if (w[0] > 0xffffffffull || w[1] > 0xffffffffull || w[2] > 0xffffffffull || w[3] > 0xffffffffull ||
            w[5] > (unsigned long long)(4 * LDO_PHILOX_BLOCKS))
            return b->fail("not a Philox state");
        for (int k = 0; k < 4 * LDO_PHILOX_BLOCKS; k++)
            if (w[6 + k] > 0xffffffffull) return b->fail("not a Philox state");
    }
    std::vector<RepAux> aux(count);
    if (count && b->get_aux(first, count, aux.data())) return -1;
    for (int r = 0; r < count; r++) {
        Rng& g = aux[r].rng;
        const unsigned long long* w = words + (size_t)r * W;
        g.key0 = (uint32_t)w[0];
        g.key1 = (uint32_t)w[1];
        g.subseq = (uint32_t)w[2];
        g.stream = (uint32_t)w[3];
        g.counter = w[4];
        g.buf_n = (int)w[5];
        for (int k = 0; k < 4 * LDO_PHILOX_BLOCKS; k++) g.buf[k] = (uint32_t)w[6 + k];
    }
    // the exchange stream is keyed by the engine's seed: it follows the key of the restored states
    if (count && first == 0) e->seed = (unsigned long long)aux[0].rng.key0 | ((unsigned long long)aux[0].rng.key1 << 32);
    return count ? b->put_aux(first, count, aux.data()) : 0;
}

int ldo_attach_tape(ldo_engine* e, int replica, const ldo_tape_draw* draws, long long n) {
    if (replica < 0 || replica >= e->b->R) return e->b->fail("bad replica index");
    return e->b->attach_tape(replica, draws, n);
}

int ldo_tape_position(ldo_engine* e, int replica, long long* pos) {
    if (replica < 0 || replica >= e->b->R) return e->b->fail("bad replica index");
    RepAux aux;
    if (e->b->get_aux(replica, 1, &aux)) return -1;
    *pos = aux.rng.tape_pos;
    return 0;
}

int ldo_set_state(ldo_engine* e, int replica, int n_chains, const int* chain_index, const int* chain_ident,
                  const int* chain_len, const int* pos, const int* ore) {
    if (replica < -1 || replica >= e->b->R) return e->b->fail("bad replica index");
    return e->load_config(e, replica, n_chains, chain_index, chain_ident, chain_len, pos, ore);
}

int ldo_replace_config(ldo_engine* e, int replica, int n_chains, const int* chain_index, const int* chain_ident,
                       const int* chain_len, const int* pos, const int* ore) {
    if (replica < -1 || replica >= e->b->R) return e->b->fail("bad replica index");
    e->b->keep_bias_on_load = 1;
    int rc = e->load_config(e, replica, n_chains, chain_index, chain_ident, chain_len, pos, ore);
    e->b->keep_bias_on_load = 0;
    return rc;
}

int ldo_state_capacity(const ldo_engine* e, int* max_chains, int* max_domains) {
    e->b->capacity(max_chains, max_domains);
    return 0;
}

int ldo_get_state(ldo_engine* e, int replica, int* n_chains, int* chain_index, int* chain_ident, int* chain_len,
                  int* pos, int* ore, int* state, int* bound) {
    if (replica < 0 || replica >= e->b->R) return e->b->fail("bad replica index");
    return e->b->decode_state(replica, n_chains, chain_index, chain_ident, chain_len, pos, ore, state, bound);
}

static int run_impl(ldo_engine* e, long long n_steps, int cf, int cd, int ccf, bool sync) {
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.op = OP_RUN;
    a.only_replica = -1;
    a.n_steps = n_steps;
    a.centering_freq = cf;
    a.centering_domain = cd;
    a.constraint_check_freq = ccf;
    if (e->b->shared.ms.n < 1) return e->b->fail("moveset not set");
    return e->b->exec(a, sync);
}

int ldo_run(ldo_engine* e, long long n_steps, int centering_freq, int centering_domain, int constraint_check_freq) {
    return run_impl(e, n_steps, centering_freq, centering_domain, constraint_check_freq, true);
}
int ldo_run_async(ldo_engine* e, long long n_steps, int centering_freq, int centering_domain, int constraint_check_freq) {
    return run_impl(e, n_steps, centering_freq, centering_domain, constraint_check_freq, false);
}
int ldo_synchronize(ldo_engine* e) {
    if (dev_sync(e->b->stream)) return e->b->fail(dev_err());
    return 0;
}
void* ldo_stream(ldo_engine* e) {
#ifdef LDO_HOSTSIM
    return nullptr;
#else
    return (void*)e->b->stream;
#endif
}

static int observe(ldo_engine* e, OpArgs& a) {
    a.op = OP_OBSERVE;
    a.only_replica = -1;
    return e->b->exec(a, true);
}

int ldo_get_status(ldo_engine* e, int* status, int* detail) {
    EngineBase* b = e->b;
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.out_status = b->d_status;
    if (observe(e, a)) return -1;
    std::vector<int> h(2 * (size_t)b->R);
    if (dev_d2h(h.data(), b->d_status, sizeof(int) * 2 * b->R, b->stream)) return b->fail(dev_err());
    for (int r = 0; r < b->R; r++) {
        if (status) status[r] = h[2 * r];
        if (detail) detail[r] = h[2 * r + 1];
    }
    return 0;
}

int ldo_get_energies(ldo_engine* e, double* out) {
    EngineBase* b = e->b;
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.out_energies = b->d_energies;
    if (observe(e, a)) return -1;
    if (dev_d2h(out, b->d_energies, sizeof(double) * 5 * b->R, b->stream)) return b->fail(dev_err());
    return 0;
}

int ldo_get_counters(ldo_engine* e, int* out) {
    EngineBase* b = e->b;
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.out_counters = b->d_counters;
    if (observe(e, a)) return -1;
    if (dev_d2h(out, b->d_counters, sizeof(int) * 9 * b->R, b->stream)) return b->fail(dev_err());
    return 0;
}

int ldo_get_staple_counts(ldo_engine* e, int* out) {
    EngineBase* b = e->b;
    int nst = b->shared.sc.n_types - 1;
    if (nst < 1) return 0;
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.out_staples = b->d_staples;
    if (observe(e, a)) return -1;
    if (dev_d2h(out, b->d_staples, sizeof(int) * nst * b->R, b->stream)) return b->fail(dev_err());
    return 0;
}

int ldo_get_order_params(ldo_engine* e, int* out) {
    EngineBase* b = e->b;
    int n = b->shared.ob.n_ops;
    if (n < 1) return 0;
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.out_ops = b->d_ops;
    if (observe(e, a)) return -1;
    if (dev_d2h(out, b->d_ops, sizeof(int) * n * b->R, b->stream)) return b->fail(dev_err());
    return 0;
}

int ldo_get_move_stats(ldo_engine* e, long long* attempts, long long* accepts) {
    EngineBase* b = e->b;
    std::vector<RepAux> aux(b->R);
    if (b->get_aux(0, b->R, aux.data())) return -1;
    int n = b->shared.ms.n;
    for (int r = 0; r < b->R; r++) {
        for (int i = 0; i < n; i++) {
            attempts[(size_t)r * n + i] = aux[r].stats.attempts[i];
            accepts[(size_t)r * n + i] = aux[r].stats.accepts[i];
        }
    }
    return 0;
}

static_assert(LDO_TRK_BINS == LDO_TRACKER_BINS, "tracker bins of the ABI and of the device code differ");
int ldo_enable_move_trackers(ldo_engine* e, int on) { return e->b->enable_trackers(on != 0); }

int ldo_get_move_trackers(ldo_engine* e, int replica, int* sticky, unsigned int* counts) {
    EngineBase* b = e->b;
    if (replica < 0 || replica >= b->R) return b->fail("bad replica index");
    TrackStats t;
    if (b->get_trackers(replica, &t)) return -1;
    int n = b->shared.ms.n;
    for (int i = 0; i < n; i++) {
        sticky[2 * i] = t.sticky_a[i];
        sticky[2 * i + 1] = t.sticky_b[i];
        memcpy(counts + (size_t)i * 2 * LDO_TRK_BINS * 2, t.cnt[i], sizeof(unsigned) * 2 * LDO_TRK_BINS * 2);
    }
    return 0;
}

static_assert(LDO_TRK_LK_CAP == LDO_LINKER_TRACKER_CAP, "linker tracker capacity of the ABI and of the device code differ");
int ldo_get_linker_trackers(ldo_engine* e, int replica, int* sticky, int* n_entries, int* entries, int* dropped) {
    EngineBase* b = e->b;
    if (replica < 0 || replica >= b->R) return b->fail("bad replica index");
    TrackStats t;
    if (b->get_trackers(replica, &t)) return -1;
    int n = b->shared.ms.n;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 6; k++) sticky[6 * i + k] = t.lk_sticky[i][k];
    *n_entries = t.lk_n;
    *dropped = t.lk_dropped;
    for (int k = 0; k < t.lk_n; k++) {
        unsigned key = t.lk_key[k];
        int* o = entries + 6 * k;
        o[0] = (int)(key >> 28);
        o[1] = (int)(key >> 26 & 3);
        o[2] = (int)(key << 6) >> 19; // two signed 13-bit values
        o[3] = (int)(key << 19) >> 19;
        o[4] = (int)t.lk_cnt[k][0];
        o[5] = (int)t.lk_cnt[k][1];
    }
    return 0;
}

int ldo_get_run_timing(ldo_engine* e, long long* out) {
    EngineBase* b = e->b;
    if (dev_d2h(out, b->run_timing_ptr(), sizeof(long long) * 3 * b->R, b->stream)) return b->fail(dev_err());
    return 0;
}

int ldo_recompute_energies(ldo_engine* e, double* energy, int* stacked_pairs) {
    EngineBase* b = e->b;
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.op = OP_RECOMPUTE;
    a.only_replica = -1;
    a.out_recomputed = b->d_recomputed;
    a.out_recomputed_stacked = b->d_recomputed_stacked;
    if (b->exec(a, true)) return -1;
    if (dev_d2h(energy, b->d_recomputed, sizeof(double) * b->R, b->stream)) return b->fail(dev_err());
    if (dev_d2h(stacked_pairs, b->d_recomputed_stacked, sizeof(int) * b->R, b->stream)) return b->fail(dev_err());
    return 0;
}

int ldo_check_all_constraints(ldo_engine* e) {
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.op = OP_CHECK_CONSTRAINTS;
    a.only_replica = -1;
    return e->b->exec(a, true);
}

int ldo_center(ldo_engine* e, int centering_domain) {
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.op = OP_CENTER;
    a.only_replica = -1;
    a.centering_domain = centering_domain;
    return e->b->exec(a, true);
}

int ldo_exchange_collect_async(ldo_engine* e) {
    EngineBase* b = e->b;
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.out_dependent = b->d_dependent;
    a.op = OP_OBSERVE;
    a.only_replica = -1;
    return b->exec(a, false);
}

int ldo_exchange_collect(ldo_engine* e, double* dependent_local) {
    EngineBase* b = e->b;
    int nq = LDO_DEP_FIXED + (b->shared.sc.n_types - 1);
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.out_dependent = b->d_dependent;
    if (observe(e, a)) return -1;
    if (dependent_local) {
        if (dev_d2h(dependent_local, b->d_dependent, sizeof(double) * nq * b->R, b->stream)) return b->fail(dev_err());
    }
    return 0;
}

int ldo_set_exchange_ladder(ldo_engine* e, int ladder_len, const int* temp_idx, const double* staple_u_mult,
                            const double* bias_mult, const double* stacking_mult) {
    EngineBase* b = e->b;
    if (ladder_len < 1) return b->fail("bad ladder length");
    b->ladder_temp_idx.assign(temp_idx, temp_idx + ladder_len);
    for (int i = 0; i < ladder_len; i++) {
        if (temp_idx[i] < 0 || temp_idx[i] >= b->shared.n_temps) return b->fail("ladder temperature index out of range");
    }
    b->ladder_staple_u_mult.assign(ladder_len, 1.0);
    b->ladder_bias_mult.assign(ladder_len, 1.0);
    b->ladder_stacking_mult.assign(ladder_len, 1.0);
    if (staple_u_mult) b->ladder_staple_u_mult.assign(staple_u_mult, staple_u_mult + ladder_len);
    if (bias_mult) b->ladder_bias_mult.assign(bias_mult, bias_mult + ladder_len);
    if (stacking_mult) b->ladder_stacking_mult.assign(stacking_mult, stacking_mult + ladder_len);
    return 0;
}

static int exchange_pt_impl(ldo_engine* e, int variant, int v2_dim, long long swap_i, int n_ladders, int ladder_len, int rank,
                            int n_ranks, const double* dependent, int* slot_to_replica, long long* attempts,
                            long long* accepts);

double ldo_exchange_acceptance_p(int n_staple_types, const double* reduced_staple_u, double temp1, double temp2,
                                 double staple_u_mult1, double staple_u_mult2, double stacking_mult1, double stacking_mult2,
                                 const double* dependent1, const double* dependent2) {
    return exchange_acceptance_p(n_staple_types, reduced_staple_u, temp1, temp2, staple_u_mult1, staple_u_mult2,
                                 stacking_mult1, stacking_mult2, dependent1, dependent2);
}

int ldo_exchange_pt(ldo_engine* e, int variant, long long swap_i, int n_ladders, int ladder_len, int rank,
                    int n_ranks, const double* dependent, int* slot_to_replica, long long* attempts,
                    long long* accepts) {
    if (variant < LDO_PT_T || variant > LDO_PT_ST) return e->b->fail("ldo_exchange_pt: 1-D variants only (see ldo_exchange_pt_2d)");
    return exchange_pt_impl(e, variant, 0, swap_i, n_ladders, ladder_len, rank, n_ranks, dependent, slot_to_replica, attempts, accepts);
}

int ldo_exchange_pt_2d(ldo_engine* e, long long swap_i, int n_ladders, int v1_dim, int v2_dim, int rank,
                       int n_ranks, const double* dependent, int* slot_to_replica, long long* attempts,
                       long long* accepts) {
    if (v1_dim < 1 || v2_dim < 1) return e->b->fail("ldo_exchange_pt_2d: bad grid");
    return exchange_pt_impl(e, LDO_PT_2D, v2_dim, swap_i, n_ladders, v1_dim * v2_dim, rank, n_ranks, dependent, slot_to_replica, attempts, accepts);
}

static int exchange_fill_args(ldo_engine* e, ExchangeArgs& x, std::vector<double>& slot_temp, int variant, int v2_dim, long long swap_i,
                              int n_ladders, int ladder_len, int rank, int n_ranks) {
    EngineBase* b = e->b;
    if ((int)b->ladder_temp_idx.size() != ladder_len) return b->fail("ldo_set_exchange_ladder not called for this ladder length");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || ladder_len % n_ranks != 0) return b->fail("ladder_len must be a multiple of n_ranks");
    if (n_ladders * (ladder_len / n_ranks) != b->R) return b->fail("this rank must hold ladder_len / n_ranks slots of every ladder");
    slot_temp.resize(ladder_len);
    for (int i = 0; i < ladder_len; i++) slot_temp[i] = b->temps[b->ladder_temp_idx[i]];
    memset(&x, 0, sizeof(x));
    x.variant = variant;
    x.v2_dim = v2_dim;
    x.swap_i = swap_i;
    x.n_ladders = n_ladders;
    x.ladder_len = ladder_len;
    x.rank = rank;
    x.n_ranks = n_ranks;
    x.n_local = b->R;
    x.n_global = n_ladders * ladder_len;
    x.n_staple_types = b->shared.sc.n_types - 1;
    x.seed = e->seed;
    x.slot_temp_idx = b->ladder_temp_idx.data();
    x.slot_temp = slot_temp.data();
    x.slot_staple_u_mult = b->ladder_staple_u_mult.data();
    x.slot_bias_mult = b->ladder_bias_mult.data();
    x.slot_stacking_mult = b->ladder_stacking_mult.data();
    return 0;
}

static int exchange_pt_impl(ldo_engine* e, int variant, int v2_dim, long long swap_i, int n_ladders, int ladder_len, int rank,
                            int n_ranks, const double* dependent, int* slot_to_replica, long long* attempts,
                            long long* accepts) {
    ExchangeArgs x;
    std::vector<double> slot_temp;
    if (exchange_fill_args(e, x, slot_temp, variant, v2_dim, swap_i, n_ladders, ladder_len, rank, n_ranks)) return -1;
    if (e->b->exchange(x, dependent, slot_to_replica, attempts, accepts)) return -1;
    // PTGCMCSimulation::run calls update_control_qs() at the top of every round, which always ends
    // in update_energy() (App. A19): rebuild the running energy with the (possibly new) tables
    return refresh_energy(e);
}

int ldo_exchange_state_set(ldo_engine* e, int n_slots, int n_counters, const int* slot_to_replica, const long long* attempts,
                           const long long* accepts) {
    return e->b->exchange_state_set(n_slots, n_counters, slot_to_replica, attempts, accepts);
}
int ldo_exchange_state_get(ldo_engine* e, int n_slots, int n_counters, int* slot_to_replica, long long* attempts, long long* accepts) {
    return e->b->exchange_state_get(n_slots, n_counters, slot_to_replica, attempts, accepts);
}
int ldo_exchange_pt_async(ldo_engine* e, int variant, int v2_dim, long long swap_i, int n_ladders, int ladder_len, int rank, int n_ranks) {
    if (variant < LDO_PT_T || variant > LDO_PT_2D) return e->b->fail("bad exchange variant");
    ExchangeArgs x;
    std::vector<double> slot_temp;
    if (exchange_fill_args(e, x, slot_temp, variant, variant == LDO_PT_2D ? v2_dim : 0, swap_i, n_ladders, ladder_len, rank, n_ranks)) return -1;
    if (e->b->exchange_resident(x)) return -1;
    return refresh_energy(e, false);
}

int ldo_exchange_windows(ldo_engine* e, long long swap_i, int n_ladders, int n_windows, int grid_bias,
                         int n_window_biases, const int* window_biases, int* window_to_replica,
                         long long* attempts, long long* accepts) {
    EngineBase* b = e->b;
    if (!b->shared.has_grid) return b->fail("no Grid bias configured");
    if (grid_bias < 0 || grid_bias >= b->shared.ob.n_biases || b->shared.ob.biases[grid_bias].type != BIAS_GRID) {
        return b->fail("grid_bias is not a Grid bias");
    }
    if (n_window_biases < 0 || n_window_biases > LDO_MAX_BIASES) return b->fail("bad window bias count");
    // the decisions read every window's STORED grid point (get_current_point, us_simulation.cpp:179-181, 773-786)
    OpArgs a;
    memset(&a, 0, sizeof(a));
    a.only_replica = -1;
    WindowExchangeArgs x;
    memset(&x, 0, sizeof(x));
    x.swap_i = swap_i;
    x.n_ladders = n_ladders;
    x.n_windows = n_windows;
    x.grid_bias = grid_bias;
    x.n_window_biases = n_window_biases;
    for (int j = 0; j < n_window_biases; j++) x.window_bias[j] = window_biases[j];
    x.seed = e->seed;
    if (b->window_exchange(x, window_to_replica, attempts, accepts)) return -1;
    a.op = OP_AFTER_WINDOW_SWAP;
    return b->exec(a, true);
}

int ldo_set_exchange_tape(ldo_engine* e, const double* reals, long long n, const long long* round_offsets, long long n_rounds) {
    return e->b->set_exchange_tape(reals, n, round_offsets, n_rounds);
}
int ldo_exchange_tape_status(ldo_engine* e, long long* missing, long long* unused) { return e->b->exchange_tape_state(missing, unused); }
long long ldo_launch_count(const ldo_engine* e) { return e->b->launches; }
const char* ldo_build_info(void) {
#ifdef LDO_HOSTSIM
    return "hostsim (tests only: host emulation of the device sources, one emulated lane)";
#else
    return "cuda sm_100a";
#endif
}
unsigned long ldo_state_bytes(const ldo_engine* e) { return e->b->state_bytes(); }

unsigned long ldo_checkpoint_size(const ldo_engine* e) { return e->b->blob_size(); }
int ldo_checkpoint_save(ldo_engine* e, int first, int count, void* host) {
    if (first < 0 || count < 0 || first + count > e->b->R) return e->b->fail("bad replica range");
    return e->b->get_blobs(first, count, host);
}
int ldo_checkpoint_load(ldo_engine* e, int first, int count, const void* host) {
    if (first < 0 || count < 0 || first + count > e->b->R) return e->b->fail("bad replica range");
    return e->b->put_blobs(first, count, host);
}

int ldo_exchange_buffers(ldo_engine* e, int n_global, void** send_dev, void** recv_dev, int* doubles_per_replica) {
    return e->b->exchange_buffers(n_global, send_dev, recv_dev, doubles_per_replica);
}

int ldo_enumerate_conformations(ldo_engine* e, const ldo_enum_job* j, int max_keys, int* n_keys, int* keys, double* weights, double* sums, long long* n_leaves) {
    EngineBase* b = e->b;
    if (j->n_staples < 0 || j->n_staples > LDO_ENUM_MAX_STAPLES) return b->fail("enumeration: too many staples in the set");
    if (j->n_stack < (j->staples_only ? 0 : 2) || j->n_stack > LDO_ENUM_MAX_DOMAINS) return b->fail("enumeration: between 2 and 32 domains");
    if (j->staples_only && b->shared.sc.n_scaffold > LDO_ENUM_MAX_DOMAINS) return b->fail("enumeration: scaffold of at most 32 domains");
    if (j->n_growthpoints < 0 || j->n_growthpoints > LDO_ENUM_MAX_STAPLES) return b->fail("enumeration: too many growthpoints");
    if (j->n_out_ops < 1 || j->n_out_ops > LDO_ENUM_MAX_OPS) return b->fail("enumeration: between 1 and 6 order parameters to output");
    if (j->n_ident < 0 || j->n_ident > LDO_ENUM_MAX_IDENT) return b->fail("enumeration: domain identities out of range");
    if (j->split_depth > LDO_ENUM_MAX_SPLIT) return b->fail("enumeration: split depth above 6");
    EnumJob job;
    memset(&job, 0, sizeof(job));
    job.n_staples = j->n_staples;
    for (int k = 0; k < j->n_staples; k++) {
        if (j->staple_type[k] < 1 || j->staple_type[k] >= b->shared.sc.n_types) return b->fail("enumeration: staple type out of range");
        job.staple_type[k] = j->staple_type[k];
    }
    job.n_stack = j->n_stack;
    for (int k = 0; k < j->n_stack; k++) {
        if (j->stack_chain[k] < 0 || j->stack_chain[k] > j->n_staples || j->stack_d[k] < 0) return b->fail("enumeration: bad domain in the stack");
        job.stack_chain[k] = (short)j->stack_chain[k];
        job.stack_d[k] = (short)j->stack_d[k];
    }
    job.n_gp = j->n_growthpoints;
    for (int k = 0; k < j->n_growthpoints; k++) {
        job.gp_old_chain[k] = (short)j->gp_old_chain[k];
        job.gp_old_d[k] = (short)j->gp_old_d[k];
        job.gp_new_chain[k] = (short)j->gp_new_chain[k];
        job.gp_new_d[k] = (short)j->gp_new_d[k];
    }
    for (int i = -j->n_ident; i <= j->n_ident; i++) job.ident_unassigned[i + LDO_ENUM_MAX_IDENT] = j->ident_unassigned[i + j->n_ident];
    job.overcount = j->overcount;
    job.n_out_ops = j->n_out_ops;
    for (int k = 0; k < j->n_out_ops; k++) {
        if (j->out_ops[k] < 0 || j->out_ops[k] >= b->shared.ob.n_ops) return b->fail("enumeration: order parameter index out of range");
        job.out_op[k] = j->out_ops[k];
    }
    job.staples_only = j->staples_only ? 1 : 0;
    if (job.staples_only) {
        for (int i = 0; i < b->shared.sc.n_scaffold; i++) {
            for (int k = 0; k < 3; k++) job.scaf_pos[i][k] = j->scaffold_pos[3 * i + k];
            int oc = ore_code(v3(j->scaffold_ore[3 * i], j->scaffold_ore[3 * i + 1], j->scaffold_ore[3 * i + 2]));
            job.scaf_ore[i] = oc == 7 ? ORE_ZERO : oc;
        }
    }
    // levels of the recursion: one per domain of the stack but the first, a growthpoint pair shares one (staples only: the
    // domains grown off the scaffold take no level)
    int levels = j->staples_only ? j->n_stack - j->n_growthpoints : j->n_stack - 1 - j->n_growthpoints;
    if (levels < 0) levels = 0;
    int split = 0;
    if (j->split_depth > 0) {
        split = j->split_depth;
    }
    else if (b->R > 1) {
        long long n = 1, per_worker = 8; // prefixes per worker: a deeper cut evens the subtrees out but costs more walks (profiles/enum_time.py)
        if (const char* pw = getenv("LDO_ENUM_PREFIXES_PER_WORKER")) per_worker = atoll(pw) > 0 ? atoll(pw) : per_worker;
        while (split < levels && split < LDO_ENUM_MAX_SPLIT && n < per_worker * b->R) {
            n *= 36;
            split++;
        }
    }
    job.split_depth = split;
    job.n_prefixes = 1;
    for (int l = 0; l < split; l++) job.n_prefixes *= 36;
    std::vector<EnumAcc>& accs = b->enum_accs;
    if (b->enumerate(job, accs)) return -1;
    // merge in worker order, keys sorted at the end: the result does not depend on the number of workers beyond rounding
    std::map<std::vector<int>, double> table;
    double z = 0, avg_e = 0, avg_b = 0, n_configs = 0;
    long long leaves = 0;
    for (const EnumAcc& a: accs) {
        if (a.status != 0) {
            char msg[96];
            snprintf(msg, sizeof msg, "enumeration failed on the device: status %d detail %d", a.status, a.status_detail);
            return b->fail(msg);
        }
        z += a.z;
        avg_e += a.avg_e;
        avg_b += a.avg_b;
        n_configs += a.n_configs;
        leaves += a.n_leaves;
        for (int k = 0; k < a.n_keys; k++) table[std::vector<int>(a.keys[k], a.keys[k] + job.n_out_ops)] += a.w[k];
    }
    if ((int)table.size() > max_keys) return b->fail("enumeration: more states than the caller's table holds");
    int k = 0;
    for (auto const& kv: table) {
        for (int i = 0; i < job.n_out_ops; i++) keys[k * job.n_out_ops + i] = kv.first[i];
        weights[k] = kv.second;
        k++;
    }
    *n_keys = k;
    sums[0] = z;
    sums[1] = avg_e;
    sums[2] = avg_b;
    sums[3] = n_configs;
    if (n_leaves) *n_leaves = leaves;
    return 0;
}

} // extern "C"
