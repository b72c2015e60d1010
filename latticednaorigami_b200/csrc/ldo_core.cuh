// ldo_core.cuh — per-replica system state and the FourBody / misbinding / stacking / hybridization
// delta-energy evaluation of the LatticeDNAOrigami model, written for one warp per replica.
//
// Execution model. Every replica is owned by one warp. All 32 lanes execute the serial Monte Carlo
// logic redundantly on identical data (warp-uniform control flow: loads broadcast, stores write the
// same value); the places where a move evaluates several candidate lattice sites are distributed
// over lanes (`for (k = LDO_LANE; k < n; k += LDO_NLANES)`) as far as they are read-only: the lattice
// lookup, the ideal-walk predicates and the misbinding weight of a site. A candidate that binds its
// complement is evaluated by the whole warp (eval_place): the pair is entered into the domain records,
// the potential evaluated, and the records restored - the reference's set-then-roll-back
// (origami_system.cpp:343-355) without the occupancy-map and counter updates.
//
// What this restates (reference file:line):
//   * Domain records, chain walk, twist/kink/junction constraints     domain.hpp:14-92, domain.cpp:9-118
//   * occupancy maps pos->state / pos->unbound domain                  origami_system.hpp:184-187, hash.hpp:16-24
//     (rebuilt as one open-addressing table: packed position -> occupant domain id)
//   * set/check/unassign one domain, add/delete chain, centre,
//     full constraint check, energy rebuild, enthalpy/entropy split   origami_system.cpp:193-871
//   * FourBody binding potential, Opposing/Disallowed misbinding       origami_potential.cpp:25-950, 1282-1350
//
// The same source compiles for the device (nvcc, sm_100a) and — for the host-emulation debugging
// harness under tests/hostsim only — as plain C++ with one emulated lane.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define LDO_HD __host__ __device__
// Out-of-line device functions: keeps the per-replica kernel's code size (instruction cache) and
// ptxas time under control; the call overhead is negligible next to the shared-memory latency chains.
#define LDO_HDN __host__ __device__ __noinline__
// Small helpers with many call sites: out of line unless LDO_INLINE_HELPERS (A/B twin, profiles/ab_r2.txt) - the kernel
// is bound by instruction supply and their inlined copies were a tenth of its code
#ifdef LDO_INLINE_HELPERS
#define LDO_HDC __host__ __device__
#else
#define LDO_HDC __host__ __device__ __noinline__
#endif
// Functions with a single call site on the hot path: merged into their caller (no code growth, one call / return and two
// instruction-line jumps fewer per use) unless LDO_SPLIT_SINGLE_SITE (A/B twin, profiles/ab_r2.txt)
#ifdef LDO_SPLIT_SINGLE_SITE
#define LDO_HDS __host__ __device__ __noinline__
#else
#define LDO_HDS __host__ __device__ __forceinline__
#endif
#else
#define LDO_HD
#define LDO_HDN
#define LDO_HDC
#define LDO_HDS
#endif

// Warps (= replicas) per block of the staged kernel (the in-place kernel uses 4-warp blocks)
#ifndef LDO_BLOCK_WARPS
#define LDO_BLOCK_WARPS 1
#endif

#if defined(__CUDA_ARCH__)
// %laneid: one S2R instead of S2R tid + mask
__device__ __forceinline__ int ldo_laneid() {
    unsigned l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return (int)l;
}
#define LDO_LANE ldo_laneid()
#define LDO_NLANES 32
#define LDO_SYNCWARP() __syncwarp()
#else
#define LDO_LANE 0
#define LDO_NLANES 1
#define LDO_SYNCWARP() ((void)0)
#endif

// Operation counters of the host emulation (profiles/count_ops.py): how often a move calls the expensive primitives.
#if defined(LDO_HOSTSIM)
extern "C" long long ldo_dbg_counts[16];
#define LDO_COUNT(i) (ldo_dbg_counts[i]++)
#if defined(LDO_CALLER_PROFILE) // profiles/step_callers.py: histogram of the call sites of System::step
extern "C" unsigned long long ldo_dbg_callers[2 * 4096];
static __attribute__((noinline)) void ldo_dbg_note_caller(void* ra) {
    unsigned long long a = (unsigned long long)ra;
    unsigned h = (unsigned)((a * 0x9E3779B97F4A7C15ull) >> 52);
    while (ldo_dbg_callers[2 * h] && ldo_dbg_callers[2 * h] != a) h = (h + 1) & 4095;
    ldo_dbg_callers[2 * h] = a;
    ldo_dbg_callers[2 * h + 1]++;
}
#endif
#else
#define LDO_COUNT(i) ((void)0)
#endif
// 0 occupant lookups, 1 table puts, 2 table erases, 3 bind_domain (complementary), 4 check_stacking, 5 eval_place,
// 6 cp_walks_remain_seg, 7 rg_site_lookup, 8 rg_compute_slot, 9 rg_fill_feeler_memo, 10 rg_feeler_general, 11 step

namespace ldo {

// ---------------------------------------------------------------------------------------------
// Lattice vectors (utility.hpp:66-95, utility.cpp:26-161)
// ---------------------------------------------------------------------------------------------

// A lattice vector is ONE 32-bit word: x + (y << 11) + (z << 22) in wrap-around (mod 2^32) arithmetic.
// The map is linear, so sums, differences and negation of vectors are single integer instructions on
// the packed words and equality is one compare, which is what the potential does almost exclusively;
// components are extracted only for |.|_1 (ideal-walk counts) and I/O. Positions obey |x|,|y| <= 511 and
// |z| <= 255 (checked when a domain is placed), so every difference of two positions (|dx|,|dy| <= 1022,
// |dz| <= 510) still has a unique packed word and unpacks exactly.
struct V3 {
    uint32_t k;
};
#define LDO_COORD_MAX_XY 511
#define LDO_COORD_MAX_Z 255

LDO_HD inline V3 v3(int x, int y, int z) {
    V3 v;
    v.k = (uint32_t)x + ((uint32_t)y << 11) + ((uint32_t)z << 22);
    return v;
}
LDO_HD inline int vx(V3 v) { return (int)(v.k << 21) >> 21; }
LDO_HD inline int vy(V3 v) { return (int)((v.k - (uint32_t)vx(v)) << 10) >> 21; }
LDO_HD inline int vz(V3 v) { return (int)(v.k - (uint32_t)vx(v) - ((uint32_t)vy(v) << 11)) >> 22; }
LDO_HD inline V3 operator+(V3 a, V3 b) {
    V3 v;
    v.k = a.k + b.k;
    return v;
}
LDO_HD inline V3 operator-(V3 a, V3 b) {
    V3 v;
    v.k = a.k - b.k;
    return v;
}
LDO_HD inline V3 operator-(V3 a) {
    V3 v;
    v.k = 0u - a.k;
    return v;
}
LDO_HD inline bool operator==(V3 a, V3 b) { return a.k == b.k; }
LDO_HD inline bool operator!=(V3 a, V3 b) { return a.k != b.k; }
LDO_HD inline int abssum(V3 a) {
    int x = (int)(a.k << 21) >> 21;
    uint32_t r = a.k - (uint32_t)x;
    int y = (int)(r << 10) >> 21;
    int z = (int)(r - ((uint32_t)y << 11)) >> 22;
    return abs(x) + abs(y) + abs(z);
}
LDO_HD inline bool in_coord_range(V3 p) {
    return abs(vx(p)) <= LDO_COORD_MAX_XY && abs(vy(p)) <= LDO_COORD_MAX_XY && abs(vz(p)) <= LDO_COORD_MAX_Z;
}

// Orientation codes index utility::vectors (utility.hpp:156-162): +x,-x,+y,-y,+z,-z; 6 = zero vector
enum { ORE_ZERO = 6 };

LDO_HD inline V3 ore_vec(int code) {
    V3 v;
    uint32_t unit = 1u << (11 * ((code >> 1) & 3)); // axis 0,1,2 -> bit 0, 11, 22
    v.k = code >= ORE_ZERO ? 0u : ((code & 1) ? 0u - unit : unit);
    return v;
}

// Returns 0..5 for unit vectors, ORE_ZERO for (0,0,0), 7 for anything else. Branch-free: a unit vector is +-(1 << t) with
// t = 0, 11 or 22 the position of its lowest set bit.
LDO_HD inline int ore_code(V3 v) {
    uint32_t k = v.k;
#if defined(__CUDA_ARCH__)
    int t = __ffs((int)k) - 1; // -1 for k == 0
#else
    int t = k ? __builtin_ctz(k) : -1;
#endif
    uint32_t neg = k >> 31;
    uint32_t unit = 1u << (t & 31);
    bool is_unit = (k == (neg ? 0u - unit : unit)) & ((t == 0) | (t == 11) | (t == 22));
    int code = 2 * ((t + 1) >> 3) + (int)neg; // t = 0, 11, 22 -> axis 0, 1, 2
    return is_unit ? code : (k == 0u ? ORE_ZERO : 7);
}

// VectorThree::rotate_half (utility.cpp:68-86): only acts when |axis| is a basis vector
LDO_HD inline V3 rotate_half(V3 v, V3 axis) {
    int a = ore_code(axis);
    if (a > 5) return v;
    int c = ore_code(v);
    if (c <= 5) return (c >> 1) == (a >> 1) ? v : -v; // unit vector: parallel stays, perpendicular flips
    int x = vx(v), y = vy(v), z = vz(v);
    if ((a >> 1) == 0) return v3(x, -y, -z);
    if ((a >> 1) == 1) return v3(-x, y, -z);
    return v3(-x, -y, z);
}

// VectorThree::rotate(axis, turns) (utility.cpp:103-142)
LDO_HD inline V3 rotate_turns(V3 v, V3 axis, int turns) {
    if (turns % 2 == 0) return rotate_half(v, axis);
    bool odd_turns_even = ((turns - 1) / 2 % 2 == 0);
    bool turns_neg = turns < 0;
    int dir = (!turns_neg && odd_turns_even) ? 1 : -1;
    int a = ore_code(axis);
    if (a > 5) return v;
    if (a & 1) dir *= -1;
    int x = vx(v), y = vy(v), z = vz(v);
    if ((a >> 1) == 0) return v3(x, -dir * z, dir * y);
    if ((a >> 1) == 1) return v3(-dir * z, y, dir * x);
    return v3(-dir * y, dir * x, z);
}

// VectorThree::rotate(origin, axis, turns) (utility.cpp:88-101)
LDO_HD inline V3 rotate_about(V3 v, V3 origin, V3 axis, int turns) {
    if (turns == 0) return v;
    return rotate_turns(v - origin, axis, turns) + origin;
}

// ---------------------------------------------------------------------------------------------
// Constants, capacities
// ---------------------------------------------------------------------------------------------

// utility::Occupancy (utility.hpp:63)
enum : uint8_t { ST_UNASSIGNED = 0, ST_UNBOUND = 1, ST_BOUND = 2, ST_MISBOUND = 3 };

enum { DOMAIN_HALFTURN = 0, DOMAIN_THREEQUARTERTURN = 1 };
enum { MISBIND_OPPOSING = 0, MISBIND_DISALLOWED = 1 };

// Replica status codes (mirror the reference's exception sites)
enum {
    LDO_OK = 0,
    LDO_ERR_TAPE_EXHAUSTED = 1, // replay tape ran out
    LDO_ERR_TAPE_MISMATCH = 2, // replayed request differs from the taped one (kind or bounds)
    LDO_ERR_COORD_RANGE = 3, // |coordinate| exceeded the packed-key range
    LDO_ERR_TABLE_FULL = 4, // occupancy table full
    LDO_ERR_CAPACITY = 5, // fixed-size move scratch exceeded
    LDO_ERR_UNASSIGNED_AT_CHECK = 6, // origami_system.cpp:272-279
    LDO_ERR_STACK_COUNT = 7, // origami_system.cpp:296-300
    LDO_ERR_ENERGY_DRIFT = 8, // origami_system.cpp:305-309
    LDO_ERR_CONSTRAINTS = 9, // origami_system.cpp:579-583
    LDO_ERR_DISTANCE = 10, // origami_system.cpp:367-369
    LDO_ERR_BIND_BOUND = 11, // origami_system.cpp:804
    LDO_ERR_SET_ASSIGNED = 12, // origami_system.cpp:526
    LDO_ERR_NONSENSICAL_P = 13, // met_movetypes.cpp:240,291
    LDO_ERR_UNBOUND_STAPLE = 14, // movetypes.cpp:273
    LDO_ERR_INTERNAL = 15
};


// One domain: packed position, orientation code, occupancy state. 8 bytes = one LDS.64.
struct __attribute__((aligned(8))) DomRec {
    uint32_t k; // V3::k
    int8_t ore;
    uint8_t state;
    uint16_t link; // LINK_* flags: where the chain neighbours of this domain are (fixed when the chain is created)
};
// Domain::m_forward_domain / m_backward_domain (domain.hpp:30-31): neighbour at d + 1 / d - 1, or - for the two ends
// of a cyclic scaffold (origami_system.cpp:687-692) - at the other end of the chain
enum : uint16_t { LINK_FWD = 1, LINK_BAC = 2, LINK_FWD_WRAP = 4, LINK_BAC_WRAP = 8 };
LDO_HD inline uint16_t chain_link_flags(int i, int len, bool cyclic_scaffold) {
    unsigned f = (i + 1 < len ? LINK_FWD : 0) | (i > 0 ? LINK_BAC : 0);
    if (cyclic_scaffold && i + 1 == len) f |= LINK_FWD_WRAP;
    if (cyclic_scaffold && i == 0) f |= LINK_BAC_WRAP;
    return (uint16_t)f;
}
LDO_HD inline V3 rec_pos(const DomRec& r) {
    V3 v;
    v.k = r.k;
    return v;
}

// System description shared by every replica (read-only on the device)
#define LDO_MAX_TYPES 192
#define LDO_MAX_IDENTS 1024
struct SysConst {
    int n_types; // chain identities including the scaffold (identity 0)
    int n_scaffold; // scaffold length
    int lmax; // staple slot stride (max_staple_size)
    int cyclic;
    int domain_type;
    int misbinding_pot;
    int apply_mean_field_cor;
    int max_total_staples;
    int max_type_staples;
    int n_ident; // identities run over [-n_ident, n_ident]; tables are (2 n_ident + 1)^2
    double staple_M; // reduced fugacity (origami_system.cpp:58)
    double stacking_ene; // Constant stacking potential, kb K (origami_potential.cpp:1219-1221)
    int type_len[LDO_MAX_TYPES];
    int type_off[LDO_MAX_TYPES];
    short idents[LDO_MAX_IDENTS]; // flattened m_identities
};

#if defined(__CUDACC__)
// Device copy of the system description: constant memory, read with uniform addresses by every lane
__constant__ SysConst ldo_c_sc;
#endif

// Staged replicas live in the dynamic shared memory of the block, one WarpSmem block per warp (ldo_engine.cu).
// The accessors of System / Engine form their addresses from the shared-memory symbol itself, so that the
// compiler emits LDS/STS with 32-bit addresses instead of generic loads and stores (which need a 64-bit
// address and a descriptor in uniform registers each time). SmemLayout<K> (byte offsets inside one warp's
// block) is defined next to WarpSmem.
#if defined(__CUDACC__) && !defined(LDO_HOSTSIM)
extern __shared__ __align__(16) unsigned char ldo_smem_raw[];
#endif
#if defined(__CUDA_ARCH__) && !defined(LDO_GENERIC_ACCESS)
// With one warp per block the warp's shared block starts at the symbol itself: every shared-memory address is
// a constant plus the CTA's window base
#if LDO_BLOCK_WARPS == 1
#define LDO_WARP_IN_BLOCK 0u
#else
#define LDO_WARP_IN_BLOCK (threadIdx.x >> 5)
#endif
#if defined(LDO_SMEM_FROM_SYMBOL)
#define LDO_SMEM_AT(K, T, offset, fallback) \
    (K::STAGED ? reinterpret_cast<T*>(ldo_smem_raw + LDO_WARP_IN_BLOCK * SmemLayout<K>::stride + (offset)) : (fallback))
#elif !defined(LDO_SMEM_FROM_THIS)
// Forming an address from the shared-memory symbol costs `S2R SR_CgaCtaId; MOV; LEA` in the prologue of every
// out-of-line function (ptxas composes the CTA's rank in its cluster into the window address). The dynamic
// shared memory of a non-cluster launch starts at a fixed offset of the CTA's window (after the 1 KB the system
// reserves on sm_100), so the accessors use that offset as a literal: every address is an immediate. The
// value is verified, not assumed: the engine probes it at creation (k_probe_smem_base) and the staged kernel
// checks it at entry (LDO_ERR_INTERNAL on every replica otherwise).
#define LDO_SMEM_WINDOW_BASE 0x400u
#define LDO_SMEM_AT(K, T, offset, fallback)                                                                         \
    (K::STAGED ? reinterpret_cast<T*>(__cvta_shared_to_generic(                                                    \
                         LDO_SMEM_WINDOW_BASE + LDO_WARP_IN_BLOCK * SmemLayout<K>::stride + (unsigned)(offset)))    \
               : (fallback))
#else
// A/B variant (profiles/README.md): the accessors are member functions of the engine object (or of its first
// member, the System), which for a staged replica sits at SmemLayout<K>::engine inside the warp's block, so the
// block's shared-space address follows from `this` with one conversion instead of being re-derived in every
// function from the symbol, the CTA's rank in its cluster and the warp index (2 S2R + 6 more instructions).
// Fewer executed instructions, but `this` then stays live everywhere: larger code, more spills, 6 % slower.
#define LDO_SMEM_AT(K, T, offset, fallback)                                                                            \
    (K::STAGED ? reinterpret_cast<T*>(__cvta_shared_to_generic(                                                       \
                         (unsigned)__cvta_generic_to_shared(this) - SmemLayout<K>::engine + (unsigned)(offset)))       \
               : (fallback))
#endif
#else
#define LDO_SMEM_AT(K, T, offset, fallback) (fallback)
#endif
#define LDO_SMEM_PTR(K, T, member, fallback) LDO_SMEM_AT(K, T, SmemLayout<K>::member, fallback)
template <class K>
struct SmemLayout;

// Per-temperature energy tables (origami_potential.cpp:1057-1221); shared by replicas at that T
struct TempTables {
    double temp;
    double init_energy, init_enthalpy, init_entropy;
    const double* hyb_energy; // [(2n+1)^2], index (a+n)*(2n+1) + (b+n)
    const double* hyb_enthalpy;
    const double* hyb_entropy;
};

// Control variables of one replica (what replica exchange permutes, ptmc_simulation.hpp:73-81)
struct Control {
    int temp_idx;
    double temp;
    double staple_u_mult;
    double bias_mult;
    double stacking_mult;
};

template <int D_, int C_, int HBITS_, int T_, bool STAGED_, int LV_, int E_, int SEG_>
struct Caps {
    static const bool STAGED = STAGED_; // replica state lives in shared memory during a launch
    static const int LV = LV_; // domains regrown by one move (regrowth levels)
    static const int E = E_; // active endpoints of one move
    static const int SEG = SEG_; // scaffold segments of one move
    static const int D = D_; // domain slots
    static const int C = C_; // chain slots (scaffold + staples)
    static const int T = T_; // chain identities (scaffold + staple types)
    static const int HBITS = HBITS_;
    static const int H = 1 << HBITS_; // occupancy table slots
    static const bool TRACK = false; // typed move trackers compiled in (Tracked<K>)
};
// The same capacities with the typed move trackers (TrackStats, ldo_moves.cuh) compiled in: a second instantiation of
// the kernels, launched only while ldo_enable_move_trackers is on, so that the run kernel of a production launch carries
// neither the tracker hooks nor their code. Every structure templated on the capacities has the same layout in both.
template <class K0>
struct Tracked : K0 {
    static const bool TRACK = true;
};

#define LDO_HEMPTY 0x80000000u // z = -512: outside the coordinate range

// Persistent configuration of one replica
template <class K>
struct alignas(16) SysState { // staged to / from shared memory with 128-bit copies
    DomRec dom[K::D];
    short bound[K::D]; // partner domain id or -1
    short ident[K::D]; // domain identity (m_d_ident)
    uint16_t dchain[K::D]; // chain slot
    uint16_t dindex[K::D]; // index in chain (m_d)
    uint32_t hkey[K::H];
    short hval[K::H];
    int chain_uid[K::C]; // unique chain index (m_c)
    uint16_t chain_type[K::C]; // chain identity (m_c_ident)
    uint16_t chain_len[K::C];
    uint8_t chain_used[K::C];
    uint16_t order[K::C]; // working order -> chain slot (m_domains order, App. B)
    int n_chains; // scaffold included
    int current_c_i; // m_current_c_i
    int num_staples, num_domains;
    int num_bound_pairs, num_fully_bound_pairs, num_self_bound_pairs;
    int num_stacked_pairs, num_unassigned;
    int constraints_violated;
    int status;
    int status_detail;
    // != 0 inside the weight passes of the recoil-growth moves: placements and removals keep the lattice, the domain
    // records and the pair counters current but leave the running energy and the stacked-pair count alone (the
    // caller restores both from its snapshots); see Engine::rg_regrow_and_test
    int weight_pass;
    double energy;
    double stack_e; // stacking_ene * stacking_mult / T for this replica
    int type_count[K::T]; // staples per identity (|m_identity_to_index[i]|)
};

struct DeltaConfig {
    double e;
    int stacked;
    bool violated;
};

// Table key = the packed position itself
LDO_HD inline uint32_t pack_pos(V3 p) { return p.k; }

// ---------------------------------------------------------------------------------------------
// Order parameters and biases: the kind evaluated once per move (`update_per_domain: false`, Engine::calc_op /
// calc_move_bias in ldo_moves.cuh) and the kind updated with every domain placement (System::pd_*)
// ---------------------------------------------------------------------------------------------

enum {
    OP_NUM_STAPLES = 0,
    OP_NUM_STAPLES_TYPE = 1,
    OP_STAPLE_TYPE_FULLY_BOUND = 2,
    OP_NUM_BOUND_DOMAIN_PAIRS = 3,
    OP_NUM_MISBOUND_DOMAIN_PAIRS = 4,
    OP_NUM_STACKED_PAIRS = 5,
    OP_NUM_LINEAR_HELICES = 6,
    OP_NUM_STACKED_JUNCTS = 7,
    OP_SUM = 8,
    OP_DIST = 9, // DistOrderParam (order_params.cpp:34-48), move-update kind, scaffold domains
    OP_ADJACENT_SITE = 10 // AdjacentSiteOrderParam (order_params.cpp:84-104)
};
enum { BIAS_LINEAR_STEP_WELL = 0, BIAS_SQUARE_WELL = 1, BIAS_GRID = 2 };

#define LDO_MAX_OPS 16
#define LDO_MAX_BIASES 8
#define LDO_MAX_SUM 8
#define LDO_MAX_GRID_DIM 3

struct OpDef {
    int type;
    int per_domain; // update_per_domain: true (Dist / AdjacentSite on scaffold domains): order_params.cpp:566-575
    int arg; // staple identity for the *Type ops; first scaffold domain of Dist / AdjacentSite
    int arg2; // second scaffold domain of Dist / AdjacentSite
    int n_sum;
    int sum_idx[LDO_MAX_SUM];
};

struct BiasDef {
    int type;
    int n_ops;
    int op_idx[LDO_MAX_GRID_DIM];
    int min_op, max_op; // LinearStepWell / SquareWell
    double well_bias, min_bias, slope, outside_bias;
};

struct OpsBiasConst {
    int n_ops;
    int n_biases;
    int n_pd_ops; // per-domain order parameters; 0 switches every per-domain hook off
    OpDef ops[LDO_MAX_OPS];
    BiasDef biases[LDO_MAX_BIASES];
};

// Per-replica bias state: window limits (MWUS overrides min_op/max_op per window,
// us_simulation.cpp:503-516) and dense grid-bias boxes (GridBiasFunction, bias_functions.cpp:242-283)
struct BiasState {
    int op_val[LDO_MAX_OPS]; // m_param of every order parameter
    unsigned op_undefined; // bit i: OrderParam::m_defined == false (a Dist / AdjacentSite domain is unassigned)
    double bias_val[LDO_MAX_BIASES]; // BiasFunction::m_bias
    double move_update_bias; // SystemBiases::m_move_update_bias (m_domain_update_bias is always 0, see System::pd_update)
    int pd_disabled; // != 0 while a configuration is being loaded (the reference builds ops / biases after set_all_domains)
    int win_min[LDO_MAX_BIASES], win_max[LDO_MAX_BIASES];
    int grid_lo[LDO_MAX_BIASES][LDO_MAX_GRID_DIM];
    int grid_n[LDO_MAX_BIASES][LDO_MAX_GRID_DIM];
    int grid_off[LDO_MAX_BIASES]; // offset into the slot's grid value / visit arrays, -1 = none
    // Which slot of the engine-wide grid arrays this replica currently uses. Window exchange swaps the
    // window-specific fields (limits, boxes, slot) of two replicas instead of shipping configurations.
    int grid_slot;
};

#if defined(__CUDACC__)
__constant__ OpsBiasConst ldo_c_ob;
#endif

template <class K>
LDO_HD inline uint32_t hash_slot(uint32_t key) {
    return (key * 0x9E3779B1u) >> (32 - K::HBITS);
}

// ---------------------------------------------------------------------------------------------
// System: state + constants + tables
// ---------------------------------------------------------------------------------------------

template <class K>
struct System {
    SysState<K>* s;
    const SysConst* sc;
    TempTables tt;
    BiasState* bsp; // per-domain order parameters (the hooks below)
    const OpsBiasConst* obp;
    // != 0: the FourBody terms are evaluated one after the other in the reference's order and their energies added
    // term by term (stacking_and_steric_terms); set for replicas that replay a tape or follow the reference's draw order
    int serial_terms;

    LDO_HD int SERIAL_TERMS() const { return *LDO_SMEM_AT(K, const int, SmemLayout<K>::engine + offsetof(System<K>, serial_terms), &serial_terms); }
    LDO_HD BiasState* BSP() const { return LDO_SMEM_PTR(K, BiasState, bias, bsp); }
    LDO_HD const OpsBiasConst& OBC() const {
#if defined(__CUDA_ARCH__)
        return ldo_c_ob;
#else
        return *obp;
#endif
    }

    LDO_HD SysState<K>* S() const {
        return LDO_SMEM_PTR(K, SysState<K>, state, s);
    }
    // The engine's System object of a staged replica sits at a fixed place of the warp's shared block
    LDO_HD const TempTables& TT() const { return *LDO_SMEM_AT(K, const TempTables, SmemLayout<K>::engine + offsetof(System<K>, tt), &tt); }
    LDO_HD const SysConst& SC() const {
#if defined(__CUDA_ARCH__)
        return ldo_c_sc;
#else
        return *sc;
#endif
    }

    LDO_HD void init(SysState<K>* s_, const SysConst* sc_, const TempTables& tt_) {
        s = s_;
        sc = sc_;
        tt = tt_;
        bsp = nullptr;
        obp = nullptr;
        serial_terms = 0;
    }

    LDO_HDN void fail(int code, int detail = 0) {
        if (S()->status == LDO_OK) {
            S()->status = code;
            S()->status_detail = detail;
        }
    }

    // ---- accessors ----
    LDO_HD V3 pos(int d) const { return rec_pos(S()->dom[d]); }
    LDO_HD V3 ore(int d) const { return ore_vec(S()->dom[d].ore); }
    LDO_HD int orc(int d) const { return S()->dom[d].ore; } // orientation code 0..6
    LDO_HD int state(int d) const { return S()->dom[d].state; }
    LDO_HD int bound(int d) const { return S()->bound[d]; }
    LDO_HD int chain(int d) const { return S()->dchain[d]; }
    LDO_HD int dindex(int d) const { return S()->dindex[d]; }
    LDO_HD int ident(int d) const { return S()->ident[d]; }
    LDO_HD int chain_base(int c) const { return c == 0 ? 0 : SC().n_scaffold + (c - 1) * SC().lmax; }
    LDO_HD int dom_id(int c, int i) const { return chain_base(c) + i; }

    // Domain::m_forward_domain / m_backward_domain: one 16-bit load of the record's link flags
    LDO_HD int fwd(int d) const {
        unsigned f = S()->dom[d].link;
        return (f & LINK_FWD) ? d + 1 : ((f & LINK_FWD_WRAP) ? 0 : -1);
    }
    LDO_HD int bac(int d) const {
        unsigned f = S()->dom[d].link;
        return (f & LINK_BAC) ? d - 1 : ((f & LINK_BAC_WRAP) ? (int)S()->chain_len[0] - 1 : -1);
    }
    // Domain::operator+ (domain.cpp:9-31). Out of line: inlining the single step at its ~150 call sites made the kernel
    // 4 % slower (instruction-cache bound, profiles/ab_r2.txt); LDO_STEP_INLINE keeps the A/B twin buildable.
#ifdef LDO_STEP_INLINE
    LDO_HD
#else
    LDO_HDN
#endif
#if defined(LDO_CALLER_PROFILE)
    __attribute__((noinline))
#endif
    int step(int d, int incr) const {
        LDO_COUNT(11);
#if defined(LDO_CALLER_PROFILE)
        ldo_dbg_note_caller(__builtin_return_address(0));
#endif
        if (d < 0) return d;
        if (incr == 1) return fwd(d);
        if (incr == -1) return bac(d);
        return step_n(d, incr);
    }
    LDO_HDN int step_n(int d, int incr) const {
#pragma unroll 1
        while (incr > 0 && d >= 0) {
            d = fwd(d);
            incr--;
        }
#pragma unroll 1
        while (incr < 0 && d >= 0) {
            d = bac(d);
            incr++;
        }
        return d;
    }

    // ---- energy tables ----
    LDO_HD int pair_index(int a, int b) const {
        int n = SC().n_ident;
        return (a + n) * (2 * n + 1) + (b + n);
    }
    LDO_HD double hyb_energy(int di, int dj) const { return TT().hyb_energy[pair_index(ident(di), ident(dj))]; }
    LDO_HD double hyb_enthalpy(int di, int dj) const { return TT().hyb_enthalpy[pair_index(ident(di), ident(dj))]; }
    LDO_HD double hyb_entropy(int di, int dj) const { return TT().hyb_entropy[pair_index(ident(di), ident(dj))]; }
    LDO_HD double stack_energy() const { return S()->stack_e; }

    // ---- occupancy table ----
    // Returns the occupant domain id at p, or -1 (origami_system.cpp:193-202, 130-132)
    LDO_HDN int occupant(V3 p) const {
        LDO_COUNT(0);
        uint32_t key = pack_pos(p);
        uint32_t i = hash_slot<K>(key);
#pragma unroll 1
        for (int n = 0; n < K::H; n++) {
            uint32_t k = S()->hkey[i];
            if (k == key) return S()->hval[i];
            if (k == LDO_HEMPTY) return -1;
            i = (i + 1) & (K::H - 1);
        }
        return -1;
    }
    LDO_HDN void table_put(V3 p, int d) {
        LDO_COUNT(1);
        if (!in_coord_range(p)) fail(LDO_ERR_COORD_RANGE, d);
        uint32_t key = pack_pos(p);
        uint32_t i = hash_slot<K>(key);
#pragma unroll 1
        for (int n = 0; n < K::H; n++) {
            uint32_t k = S()->hkey[i];
            if (k == key || k == LDO_HEMPTY) {
                S()->hkey[i] = key;
                S()->hval[i] = (short)d;
                return;
            }
            i = (i + 1) & (K::H - 1);
        }
        fail(LDO_ERR_TABLE_FULL, d);
    }
    // Linear-probing erase with backward shift (no tombstones: erases are as frequent as inserts)
    LDO_HDS void table_erase(V3 p) {
        LDO_COUNT(2);
        uint32_t key = pack_pos(p);
        uint32_t i = hash_slot<K>(key);
        int n = 0;
#pragma unroll 1
        while (S()->hkey[i] != key) {
            if (S()->hkey[i] == LDO_HEMPTY || ++n == K::H) return;
            i = (i + 1) & (K::H - 1);
        }
        uint32_t j = i;
#pragma unroll 1
        for (;;) {
            j = (j + 1) & (K::H - 1);
            uint32_t kj = S()->hkey[j];
            if (kj == LDO_HEMPTY) break;
            uint32_t h = hash_slot<K>(kj);
            // move kj into the hole i unless its home slot lies cyclically in (i, j]
            bool home_between = (i <= j) ? (i < h && h <= j) : (i < h || h <= j);
            if (!home_between) {
                S()->hkey[i] = kj;
                S()->hval[i] = S()->hval[j];
                i = j;
            }
        }
        S()->hkey[i] = LDO_HEMPTY;
    }
    LDO_HD void table_clear() {
#pragma unroll 1
        for (int i = 0; i < K::H; i++) S()->hkey[i] = LDO_HEMPTY;
    }

    // ---- domain constraint checkers (domain.cpp:33-118) ----
    LDO_HDN bool check_twist(int d1, V3 ndr, int d2) const {
        if (SC().domain_type == DOMAIN_HALFTURN) {
            // rotate_half on orientation codes: a unit vector parallel to the axis stays, a perpendicular one flips
            int a = ore_code(ndr), c1 = orc(d1), c2 = orc(d2);
            if (a > 5 || c1 >= ORE_ZERO || (c1 >> 1) == (a >> 1)) return c1 == c2;
            return (c1 ^ 1) == c2;
        }
        return rotate_turns(ore(d1), ndr, -1) == ore(d2);
    }
    LDO_HDS bool check_kink(int d1, V3 ndr, int d2) const {
        V3 o1 = ore(d1), o2 = ore(d2);
        if (ndr == -o1) return false;
        if (ndr == o1) {
            if (SC().domain_type == DOMAIN_HALFTURN) {
                if (o2 == -o1) return false;
            }
            else {
                if (o1 == o2 || o1 == -o2) return false;
            }
            return true;
        }
        if (ndr == o2 || ndr == -o2) return false;
        return true;
    }
    LDO_HDS bool check_junction_constraint(int j1, int j2, int k1, int k2) const {
        if (SC().domain_type == DOMAIN_HALFTURN) return true;
        V3 ndr_k1 = pos(k2) - pos(k1);
        if (ndr_k1 == ore(k1)) {
            V3 ndr_1 = pos(j2) - pos(j1);
            if (dindex(j1) > dindex(j2)) ndr_1 = -ndr_1;
            if (!check_twist(k1, ndr_1, k2)) return false;
        }
        return true;
    }

    // ---- free helpers (origami_potential.cpp:25-127) ----
    LDO_HD bool exists_bound(int d) const { return d >= 0 && state(d) == ST_BOUND; }
    LDO_HDN bool doubly_contiguous(int d1, int d2) const {
        if (chain(d1) != chain(d2) || dindex(d2) != dindex(d1) + 1) return false;
        int b1 = bound(d1), b2 = bound(d2);
        if (chain(b1) != chain(b2)) return false;
        return abs(dindex(b2) - dindex(b1)) == 1;
    }
    LDO_HDN bool pair_stacked(int d1, int d2) const {
        if (dindex(d1) > dindex(d2)) {
            int t = d1;
            d1 = d2;
            d2 = t;
        }
        V3 ndr = pos(d2) - pos(d1);
        V3 o1 = ore(d1);
        if (ndr != o1 && ndr != -o1) return check_twist(d1, ndr, d2);
        return false;
    }
    LDO_HDS int junction_stacking_penalty(int j1, int j2, int j3, int j4, int k1, int k2) const {
        int penalty = 0;
        V3 ndr_k1 = pos(k2) - pos(k1);
        if (ndr_k1 == ore(k1)) {
            V3 ndr_1 = pos(j2) - pos(j1);
            if (dindex(j1) > dindex(j2)) ndr_1 = -ndr_1;
            V3 ndr_3 = pos(j4) - pos(j3);
            if (dindex(j3) > dindex(j4)) ndr_3 = -ndr_3;
            if (ndr_1 == ndr_3) penalty = 2;
            else if (ndr_1 == -ndr_3) penalty = 0;
            else penalty = 1;
        }
        return penalty;
    }

    // ---- triplet terms (origami_potential.cpp:158-210) ----
    LDO_HDN void triplet_single_stacking(DeltaConfig& dc, int h1, int h2, int h3) const {
        V3 ndr_1 = pos(h2) - pos(h1);
        if (ndr_1 == ore(h1)) return;
        V3 ndr_2 = pos(h3) - pos(h2);
        if (ndr_1 != ndr_2) {
            dc.e -= stack_energy();
            dc.stacked -= 1;
        }
    }
    LDO_HDN void triplet_double_stacking(DeltaConfig& dc, int h1, int h2, int h3) const {
        V3 ndr_1 = pos(h2) - pos(h1);
        V3 ndr_2 = pos(h3) - pos(h2);
        if (ndr_1 != ndr_2) {
            dc.e -= stack_energy() / 2;
            dc.e -= stack_energy() / 2;
            dc.stacked -= 1;
        }
    }
    LDO_HDN void triply_contig_helix(DeltaConfig& dc, int h1, int h2, int h3) const {
        V3 ndr_1 = pos(h2) - pos(h1);
        V3 ndr_2 = pos(h3) - pos(h2);
        if (ndr_1 != ndr_2) dc.violated = true;
    }

    // origami_potential.cpp:419-462
    LDO_HDN void check_junction(DeltaConfig& dc, int j1, int j2, int j3, int j4, int k1, int k2) const {
        if (dindex(k1) > dindex(k2)) {
            int t = k1;
            k1 = k2;
            k2 = t;
            t = j1;
            j1 = j4;
            j4 = t;
            t = j2;
            j2 = j3;
            j3 = t;
        }
        if (!(pair_stacked(j1, j2) && pair_stacked(j3, j4))) return;
        if (!check_junction_constraint(j1, j2, k1, k2)) {
            dc.violated = true;
            return;
        }
        int penalty = junction_stacking_penalty(j1, j2, j3, j4, k1, k2);
        if (penalty == 1) {
            dc.e -= stack_energy() / 2;
            dc.e -= stack_energy() / 2;
            dc.stacked -= 1;
        }
        else if (penalty == 2) {
            dc.e -= stack_energy();
            dc.e -= stack_energy();
            dc.stacked -= 2;
        }
    }

    // Second-junction-pair scan shared by the three single-junction routines
    // (origami_potential.cpp:713-745, 809-841, 905-933): pairs (a, a_next) on the kink chain and on
    // the chain bound to a. `forward_role`: true when the pair found is (j3, j4), false for (j2, j1).
    LDO_HDN void scan_second_pairs(
            DeltaConfig& dc,
            int a,
            int a_next,
            bool found_is_j34,
            int fj1,
            int fj2,
            int k1,
            int k2) const {
        int sel_a[3], sel_b[3];
        int n = 0;
        sel_a[n] = a;
        sel_b[n] = a_next;
        n++;
        int ab = bound(a);
        int ab_for = fwd(ab), ab_bac = bac(ab);
        if (exists_bound(a_next)) {
            int anb = bound(a_next);
            if (ab_for != anb) {
                sel_a[n] = ab;
                sel_b[n] = ab_for;
                n++;
            }
            if (ab_bac != anb) {
                sel_a[n] = ab;
                sel_b[n] = ab_bac;
                n++;
            }
        }
        else {
            sel_a[n] = ab;
            sel_b[n] = ab_for;
            n++;
            sel_a[n] = ab;
            sel_b[n] = ab_bac;
            n++;
        }
#pragma unroll 1
        for (int q = 0; q < n; q++) {
            if (!exists_bound(sel_b[q])) continue;
            if (found_is_j34) check_junction(dc, fj1, fj2, sel_a[q], sel_b[q], k1, k2);
            else check_junction(dc, sel_b[q], sel_a[q], fj1, fj2, k1, k2);
        }
    }

    // origami_potential.cpp:654-748 (passed domains are the second junction pair j3, j4)
    LDO_HDN void backward_single_junction(DeltaConfig& dc, int d1, int d2) const {
        int j3 = d1, j4 = d2;
        int sel_k2[3], sel_k1[3];
        int n = 0;
        int j3_bac = bac(j3);
        sel_k2[n] = j3;
        sel_k1[n] = j3_bac;
        n++;
        int j3b = bound(j3);
        int j4b = bound(j4);
        int j3b_for = fwd(j3b), j3b_bac = bac(j3b);
        if (j3b_for == j4b) {
            sel_k2[n] = j3b;
            sel_k1[n] = j3b_bac;
            n++;
        }
        else if (exists_bound(j3_bac) && exists_bound(j3b_bac) && bound(j3_bac) == j3b_bac) {
            sel_k2[n] = j3b;
            sel_k1[n] = j3b_for;
            n++;
        }
        else if (exists_bound(j3_bac) && exists_bound(j3b_for) && bound(j3_bac) == j3b_for) {
            sel_k2[n] = j3b;
            sel_k1[n] = j3b_bac;
            n++;
        }
        else {
            sel_k2[n] = j3b;
            sel_k1[n] = j3b_for;
            n++;
            sel_k2[n] = j3b;
            sel_k1[n] = j3b_bac;
            n++;
        }
#pragma unroll 1
        for (int q = 0; q < n; q++) {
            int k2 = sel_k2[q], k1 = sel_k1[q];
            if (!exists_bound(k1)) continue;
            if (pair_stacked(k1, k2)) continue;
            int dir = dindex(k1) - dindex(k2);
            int k1_next = step(k1, dir);
            scan_second_pairs(dc, k1, k1_next, false, j3, j4, k1, k2);
        }
    }

    // origami_potential.cpp:750-844 (passed domains are the first junction pair j1, j2)
    LDO_HDN void forward_single_junction(DeltaConfig& dc, int d1, int d2) const {
        int j1 = d1, j2 = d2;
        int sel_k1[3], sel_k2[3];
        int n = 0;
        int j2_for = fwd(j2);
        sel_k1[n] = j2;
        sel_k2[n] = j2_for;
        n++;
        int j1b = bound(j1);
        int j2b = bound(j2);
        int j2b_for = fwd(j2b), j2b_bac = bac(j2b);
        if (fwd(j1b) == j2b) {
            sel_k1[n] = j2b;
            sel_k2[n] = j2b_for;
            n++;
        }
        else if (exists_bound(j2_for) && exists_bound(j2b_bac) && bound(j2_for) == j2b_bac) {
            sel_k1[n] = j2b;
            sel_k2[n] = j2b_for;
            n++;
        }
        else if (exists_bound(j2_for) && exists_bound(j2b_for) && bound(j2_for) == j2b_for) {
            sel_k1[n] = j2b;
            sel_k2[n] = j2b_bac;
            n++;
        }
        else {
            sel_k1[n] = j2b;
            sel_k2[n] = j2b_for;
            n++;
            sel_k1[n] = j2b;
            sel_k2[n] = j2b_bac;
            n++;
        }
#pragma unroll 1
        for (int q = 0; q < n; q++) {
            int k1 = sel_k1[q], k2 = sel_k2[q];
            if (!exists_bound(k2)) continue;
            if (pair_stacked(k1, k2)) continue;
            int dir = dindex(k2) - dindex(k1);
            int k2_next = step(k2, dir);
            scan_second_pairs(dc, k2, k2_next, true, j1, j2, k1, k2);
        }
    }

    // origami_potential.cpp:846-934 (passed domains are the kink pair)
    LDO_HDS void central_single_junction(DeltaConfig& dc, int d1, int d2) const {
        int k1 = d1, k2 = d2;
        int k1b = bound(k1), k2b = bound(k2);
        if (chain(k1b) == chain(k2b) && abs(dindex(k1b) - dindex(k2b)) == 1) return;

        int sel_j2[3], sel_j1[3];
        int n = 0;
        int k1_bac = bac(k1);
        sel_j2[n] = k1;
        sel_j1[n] = k1_bac;
        n++;
        int k1b_for = fwd(k1b), k1b_bac = bac(k1b);
        if (exists_bound(k1_bac)) {
            int kbb = bound(k1_bac);
            if (k1b_for != kbb) {
                sel_j2[n] = k1b;
                sel_j1[n] = k1b_for;
                n++;
            }
            if (k1b_bac != kbb) {
                sel_j2[n] = k1b;
                sel_j1[n] = k1b_bac;
                n++;
            }
        }
        else {
            sel_j2[n] = k1b;
            sel_j1[n] = k1b_for;
            n++;
            sel_j2[n] = k1b;
            sel_j1[n] = k1b_bac;
            n++;
        }
#pragma unroll 1
        for (int q = 0; q < n; q++) {
            int j2 = sel_j2[q], j1 = sel_j1[q];
            if (!exists_bound(j1)) continue;
            int k2_for = fwd(k2);
            scan_second_pairs(dc, k2, k2_for, true, j1, j2, k1, k2);
        }
    }

    // origami_potential.cpp:464-519
    LDO_HDN void backward_triplet_combos(DeltaConfig& dc, int d1, int d2) const {
        int h2 = d1, h3 = d2;
        if (!pair_stacked(h2, h3)) return;
        int h1 = step(d1, -1);
        bool h2_h3_dc = doubly_contiguous(h2, h3);
        if (exists_bound(h1)) {
            bool h1_h2_dc = doubly_contiguous(h1, h2);
            if (pair_stacked(h1, h2)) {
                if (h1_h2_dc && h2_h3_dc) {
                    triply_contig_helix(dc, h1, h2, h3);
                    if (dc.violated) return;
                }
                triplet_double_stacking(dc, h1, h2, h3);
            }
            else {
                triplet_single_stacking(dc, h1, h2, h3);
            }
        }
        int h2_prev = h1;
        int d1b = bound(d1);
        h1 = step(d1b, 1);
        if (exists_bound(h1) && bound(h1) != h2_prev && bound(h1) != h3) {
            if (pair_stacked(d1b, h1)) triplet_double_stacking(dc, h1, h2, h3);
        }
        h1 = step(d1b, -1);
        if (exists_bound(h1) && bound(h1) != h2_prev && bound(h1) != h3) {
            if (pair_stacked(h1, d1b)) triplet_double_stacking(dc, h1, h2, h3);
            else triplet_single_stacking(dc, h1, d1b, h3);
        }
    }

    // origami_potential.cpp:521-585
    LDO_HDN void forward_triplet_combos(DeltaConfig& dc, int d1, int d2) const {
        int h2 = d2, h1 = d1;
        bool first_pair_stacked = pair_stacked(h1, h2);
        int h3 = step(d2, 1);
        bool h1_h2_dc = doubly_contiguous(h1, h2);
        if (exists_bound(h3)) {
            bool second_pair_stacked = pair_stacked(h2, h3);
            if (first_pair_stacked && second_pair_stacked) {
                bool h2_h3_dc = doubly_contiguous(h2, h3);
                if (h1_h2_dc && h2_h3_dc) {
                    triply_contig_helix(dc, h1, h2, h3);
                    if (dc.violated) return;
                }
                triplet_double_stacking(dc, h1, h2, h3);
            }
            else if (second_pair_stacked) {
                triplet_single_stacking(dc, h1, h2, h3);
            }
        }
        int h2_next = h3;
        int d2b = bound(d2);
        h3 = step(d2b, 1);
        if (exists_bound(h3) && bound(h3) != h2_next && bound(h3) != h1) {
            if (pair_stacked(d2b, h3)) {
                if (first_pair_stacked) triplet_double_stacking(dc, h1, h2, h3);
                else triplet_single_stacking(dc, h1, h2, h3);
            }
        }
        h3 = step(d2b, -1);
        if (exists_bound(h3) && bound(h3) != h2_next && bound(h3) != h1) {
            if (pair_stacked(h3, d2b)) {
                if (first_pair_stacked) triplet_double_stacking(dc, h1, h2, h3);
                else triplet_single_stacking(dc, h1, h2, h3);
            }
            else if (first_pair_stacked) {
                triplet_single_stacking(dc, h3, d2b, h1);
            }
        }
    }

    // origami_potential.cpp:587-652
    // di_prev .. dj_forw: the chain neighbours of the pair (computed once per evaluation, see stacking_and_steric_terms)
    LDO_HDS void central_triplet_combos(DeltaConfig& dc, int di, int dj, int di_prev, int di_forw, int dj_prev, int dj_forw) const {
        int h1 = di_prev;
        int h2 = di;
        int h3 = dj_forw;
        int h2_next = di_forw;
        if (exists_bound(h1) && exists_bound(h3) && bound(h3) != h2_next && !doubly_contiguous(h1, h2)) {
            if (pair_stacked(dj, h3)) {
                if (pair_stacked(h1, h2)) triplet_double_stacking(dc, h1, h2, h3);
                else triplet_single_stacking(dc, h1, h2, h3);
            }
        }
        h3 = dj_prev;
        if (exists_bound(h1) && exists_bound(h3) && bound(h3) != h1 && !doubly_contiguous(h1, h2)) {
            if (pair_stacked(h1, h2)) {
                if (pair_stacked(h3, dj)) triplet_double_stacking(dc, h1, h2, h3);
                else triplet_single_stacking(dc, h3, dj, h1);
            }
            else if (pair_stacked(h3, dj)) {
                triplet_single_stacking(dc, h1, h2, h3);
            }
        }
        h2 = di;
        h3 = di_forw;
        h1 = dj_forw;
        int h2_prev = di_prev;
        if (exists_bound(h1) && exists_bound(h3) && bound(h1) != h3 && !doubly_contiguous(h2, h3)) {
            if (pair_stacked(dj, h1) && pair_stacked(h2, h3)) triplet_double_stacking(dc, h1, h2, h3);
        }
        h1 = dj_prev;
        if (exists_bound(h1) && exists_bound(h3) && bound(h1) != h2_prev && !doubly_contiguous(h2, h3)) {
            if (pair_stacked(h2, h3)) {
                if (pair_stacked(h1, dj)) triplet_double_stacking(dc, h1, h2, h3);
                else triplet_single_stacking(dc, h1, dj, h3);
            }
        }
    }

    // origami_potential.cpp:288-320
    LDO_HDS void regular_pair_constraints(DeltaConfig& dc, int d1, int d2, int i) const {
        V3 ndr = pos(d2) - pos(d1);
        if (!check_kink(d1, ndr, d2)) {
            dc.violated = true;
            return;
        }
        if (pair_stacked(d1, d2)) {
            dc.e += stack_energy();
            dc.stacked += 1;
            if (i == -1) backward_single_junction(dc, d1, d2);
            else forward_single_junction(dc, d1, d2);
        }
        else {
            central_single_junction(dc, d1, d2);
        }
        if (i == -1) backward_triplet_combos(dc, d1, d2);
        else forward_triplet_combos(dc, d1, d2);
    }

    // origami_potential.cpp:322-357
    LDO_HDS void doubly_contig_helix_pair(DeltaConfig& dc, int d1, int d2, int i, int j) const {
        if (j == 1) return;
        V3 ndr = pos(d2) - pos(d1);
        V3 o1 = ore(d1);
        if (ndr == o1 || ndr == -o1) {
            dc.violated = true;
            return;
        }
        if (check_twist(d1, ndr, d2)) {
            dc.e += stack_energy();
            dc.stacked += 1;
        }
        else {
            dc.violated = true;
            return;
        }
        if (i == -1) {
            backward_triplet_combos(dc, d1, d2);
            backward_single_junction(dc, d1, d2);
        }
        else {
            forward_triplet_combos(dc, d1, d2);
            forward_single_junction(dc, d1, d2);
        }
    }

    // origami_potential.cpp:359-417
    LDO_HDS void doubly_contig_junction_pair(DeltaConfig& dc, int d1, int d2, int j) const {
        int k1 = d1, k2 = d2;
        V3 ndr = pos(k2) - pos(k1);
        if (ore(k1) != ndr) {
            dc.violated = true;
            return;
        }
        if (j == 1) return;
        int d1b = bound(d1);
        int sel_j1[2], sel_j2[2];
        sel_j1[0] = step(d1, -1);
        sel_j2[0] = d1;
        sel_j1[1] = step(d1b, 1);
        sel_j2[1] = d1b;
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            int j1 = sel_j1[q], j2 = sel_j2[q];
            int j3 = k2;
            int j4 = step(k2, 1);
            if (exists_bound(j4)) {
                if (exists_bound(j1)) check_junction(dc, j1, j2, j3, j4, k1, k2);
            }
            j3 = bound(k2);
            j4 = step(j3, -1);
            if (exists_bound(j4)) {
                if (exists_bound(j1)) check_junction(dc, j1, j2, j3, j4, k1, k2);
            }
        }
    }

    // origami_potential.cpp:231-286, in its three independent parts: the pair (cd + i, cd + i + 1) for i = -1, 0 and the
    // triplet centred on cd
    LDO_HDS void check_constraints_pair(DeltaConfig& dc, int cd, int i, int j, int prev, int forw) const {
        int d1 = i == -1 ? prev : cd;
        int d2 = i == -1 ? cd : forw;
        if (!(exists_bound(d1) && exists_bound(d2))) return;
        int b1 = bound(d1), b2 = bound(d2);
        int rel = chain(b1) == chain(b2) ? (int)dindex(b1) - (int)dindex(b2) : 0;
        if (rel == -1) doubly_contig_helix_pair(dc, d1, d2, i, j);
        else if (rel == 1) doubly_contig_junction_pair(dc, d1, d2, j);
        else regular_pair_constraints(dc, d1, d2, i);
    }
    LDO_HDS void check_constraints_middle(DeltaConfig& dc, int cd, int prev, int forw) const {
        if (exists_bound(prev) && exists_bound(forw)) {
            if (pair_stacked(cd, forw)) {
                if (pair_stacked(prev, cd)) {
                    if (doubly_contiguous(prev, cd) && doubly_contiguous(cd, forw)) {
                        triply_contig_helix(dc, prev, cd, forw);
                        if (dc.violated) return;
                    }
                    triplet_double_stacking(dc, prev, cd, forw);
                }
                else {
                    triplet_single_stacking(dc, prev, cd, forw);
                }
            }
        }
    }

    // JunctionBindingPotential::calc_stacking_and_steric_terms (origami_potential.cpp:212-229):
    // check_constraints(di, 0), check_constraints(dj, 1), check_central_triplet_stacking_combos(di, dj). The seven
    // parts - two pairs and the middle triplet for each of the two domains, and the central triplet combinations - read
    // the domain records only and add independent terms, so each is evaluated by one lane and the contributions are
    // reduced across the warp (every term of the potential is a whole number of stacked pairs: the energy follows as
    // stacked x stacking energy). Called warp-uniformly. The reference evaluates the parts in this order and stops at the
    // first violation (origami_potential.cpp:218-223, 256-258): check_stacking hands the terms counted up to there to
    // unassign_domain / set_checked_domain_config, which apply them whatever the flag says, so only the parts up to and
    // including the first violated one are added (trajectories of the reference pass through configurations its own
    // full constraint check refuses). It also adds the energies term by term (+s, -s/2 -s/2, ...: its running sum
    // passes through 1.5 s, 2.5 s, 3 s, which round), and a move whose terms cancel leaves a residue of an ulp that
    // decides whether p == 1 consumes a draw: replicas that replay a tape (serial_terms) therefore evaluate the parts
    // one after the other with one running sum - bit for bit the reference's arithmetic; production replicas use
    // the lanes and stacked x s. Both found by the replay campaign (tests/stress_replay.py).
    LDO_HDS void stacking_task(DeltaConfig& dc, int task, int di, int dj, int di_prev, int di_forw, int dj_prev, int dj_forw) const {
        if (task == 6) {
            central_triplet_combos(dc, di, dj, di_prev, di_forw, dj_prev, dj_forw);
            return;
        }
        int cd = task < 3 ? di : dj, j = task < 3 ? 0 : 1, t = task < 3 ? task : task - 3;
        int prev = task < 3 ? di_prev : dj_prev, forw = task < 3 ? di_forw : dj_forw;
        if (t < 2) check_constraints_pair(dc, cd, t - 1, j, prev, forw);
        else check_constraints_middle(dc, cd, prev, forw);
    }
    LDO_HDN void stacking_and_steric_terms(DeltaConfig& dc, int di, int dj) const {
        // the four chain neighbours every part looks at, once (they were 20 calls of step per evaluation)
        int di_prev = bac(di), di_forw = fwd(di), dj_prev = bac(dj), dj_forw = fwd(dj);
        DeltaConfig part;
        part.e = 0;
        part.stacked = 0;
        part.violated = false;
#if defined(__CUDA_ARCH__) && !defined(LDO_SERIAL_POTENTIAL)
        // one pass with a part per lane, or (serial_terms) seven passes with the whole warp on one part after the other:
        // one call site for both (the mode is re-read from shared memory rather than kept in a register across the parts)
#define LDO_SERIAL_TERMS_NOW() (SERIAL_TERMS() != 0)
#define LDO_TERMS_LANE() LDO_LANE
#else
#define LDO_SERIAL_TERMS_NOW() true
#define LDO_TERMS_LANE() 0
#endif
        int pass = 0;
#pragma unroll 1
        for (;;) {
            int task = LDO_SERIAL_TERMS_NOW() ? pass : LDO_TERMS_LANE();
            if (task < 7) stacking_task(part, task, di, dj, di_prev, di_forw, dj_prev, dj_forw);
            // serial: stop at the first violation (origami_potential.cpp:218-223, 256-258)
            if (!LDO_SERIAL_TERMS_NOW() || part.violated || ++pass == 7) break;
        }
#if defined(__CUDA_ARCH__) && !defined(LDO_SERIAL_POTENTIAL)
        if (!LDO_SERIAL_TERMS_NOW()) {
            // the parts after the first violated one are not evaluated by the reference: they do not count
            const int lane = LDO_LANE;
            unsigned vmask = __ballot_sync(0xffffffffu, part.violated);
            int stacked = __reduce_add_sync(0xffffffffu, (lane < 7 && (vmask & ((1u << lane) - 1u)) == 0u) ? part.stacked : 0);
            part.violated = vmask != 0u;
            part.stacked = stacked;
            part.e = stacked * stack_energy();
        }
#endif
#undef LDO_SERIAL_TERMS_NOW
#undef LDO_TERMS_LANE
        dc.violated = dc.violated || part.violated;
        dc.stacked += part.stacked;
        dc.e += part.e;
    }

    // BindingPotential::check_stacking (origami_potential.cpp:149-156)
    LDO_HDN DeltaConfig check_stacking(int di, int dj) const {
        LDO_COUNT(4);
        DeltaConfig dc;
        dc.e = 0;
        dc.stacked = 0;
        dc.violated = false;
        stacking_and_steric_terms(dc, di, dj);
        return dc;
    }

    // OrigamiPotential::bind_domain (origami_potential.cpp:1282-1293) on the (possibly overlaid) pair
#ifdef LDO_X_BIND_MERGE // experiment twin (profiles/ab_r2.txt)
    LDO_HDS
#else
    LDO_HDN
#endif
    DeltaConfig bind_domain(int di) const {
        int dj = bound(di);
        DeltaConfig dc;
        dc.e = 0;
        dc.stacked = 0;
        dc.violated = false;
        int ci = orc(di), cj = orc(dj);
        bool opposing = ci == (cj < ORE_ZERO ? (cj ^ 1) : cj);
        if (ident(di) == -ident(dj)) {
            LDO_COUNT(3);
            // BindingPotential::bind_domains (origami_potential.cpp:131-147)
            if (!opposing) {
                dc.violated = true;
                return dc;
            }
            stacking_and_steric_terms(dc, di, dj);
            if (dc.violated) {
                dc.e = 0;
                dc.stacked = 0;
                return dc;
            }
            dc.e += hyb_energy(di, dj);
        }
        else {
            // origami_potential.cpp:938-950
            if (SC().misbinding_pot == MISBIND_DISALLOWED || !opposing) {
                dc.violated = true;
                return dc;
            }
            dc.e += hyb_energy(di, dj);
        }
        return dc;
    }

    // Placing d with orientation code o on the site of the unbound, NON-complementary domain j: the misbinding
    // potentials look at the two orientations and the pair's hybridization energy only
    // (origami_potential.cpp:938-950), so this needs no placement and is safe to call from any lane.
    LDO_HD DeltaConfig eval_misbind(int d, int j, int o) const {
        DeltaConfig dc;
        dc.e = 0;
        dc.stacked = 0;
        dc.violated = false;
        int cj = S()->dom[j].ore;
        bool opposing = o == (cj < ORE_ZERO ? (cj ^ 1) : cj);
        if (SC().misbinding_pot == MISBIND_DISALLOWED || !opposing) dc.violated = true;
        else dc.e = hyb_energy(d, j);
        return dc;
    }

    // -----------------------------------------------------------------------------------------
    // Evaluation of placing unassigned domain d at (p, o): what OrigamiSystem::check_domain_constraints
    // (origami_system.cpp:343-355, 828-871) returns; the state is the same before and after. Warp-uniform:
    // every lane of the warp calls it with the same arguments (the lane-parallel evaluators of the moves look
    // sites up on their own lanes and leave the complementary bindings to the whole warp). `new_state` receives
    // the state d would take.
    // -----------------------------------------------------------------------------------------
    LDO_HDN DeltaConfig eval_place(int d, V3 p, int o, int* new_state, int* partner) {
        LDO_COUNT(5);
        DeltaConfig dc;
        dc.e = 0;
        dc.stacked = 0;
        dc.violated = false;
        int j = occupant(p);
        *partner = -1;
        if (j < 0) {
            *new_state = ST_UNBOUND;
            return dc;
        }
        int sj = S()->dom[j].state;
        if (sj != ST_UNBOUND) {
            dc.violated = true;
            *new_state = ST_UNASSIGNED;
            return dc;
        }
        *partner = j;
        bool comp = S()->ident[d] == -S()->ident[j];
        if (!comp) {
            *new_state = ST_MISBOUND;
            return eval_misbind(d, j, o);
        }
        // The pair is entered into the domain records for the duration of the evaluation and taken out again
        // (what the reference does, origami_system.cpp:343-355, without the occupancy-map and counter updates:
        // the potential reads domain records only). Every lane writes the same values; the barriers keep a
        // lane that is ahead from restoring the records while another still reads them.
        LDO_SYNCWARP();
        DomRec saved = S()->dom[d];
        short saved_bd = S()->bound[d], saved_bj = S()->bound[j];
        int st = comp ? ST_BOUND : ST_MISBOUND;
        LDO_SYNCWARP();
        S()->dom[d].k = p.k;
        S()->dom[d].ore = (int8_t)o;
        S()->dom[d].state = (uint8_t)st;
        S()->dom[j].state = (uint8_t)st;
        S()->bound[d] = (short)j;
        S()->bound[j] = (short)d;
        *new_state = st;
        LDO_SYNCWARP();
        dc = bind_domain(d);
        LDO_SYNCWARP();
        S()->dom[d] = saved;
        S()->dom[j].state = ST_UNBOUND;
        S()->bound[d] = saved_bd;
        S()->bound[j] = saved_bj;
        LDO_SYNCWARP();
        if (SC().apply_mean_field_cor && !dc.violated && comp) {
            // origami_system.cpp:858-868 (counter already incremented in the reference at this point)
            int nfb = S()->num_fully_bound_pairs + 1;
            if (nfb == 1) dc.e += 2 * log(6.0);
            else if (nfb == 2) dc.e += log(3.0);
        }
        return dc;
    }

    // -----------------------------------------------------------------------------------------
    // Per-domain-update order parameters (OrigamiSystemWithBias, origami_system.cpp:873-957; SystemOrderParams::
    // update_one_domain, order_params.cpp:603-612): Dist / AdjacentSite parameters between scaffold domains marked
    // update_per_domain are recomputed (calc_param: value and `defined`) whenever one of their domains is unassigned or
    // placed by set_checked_domain_config - and NOT by set_domain_config, which leaves them as the preceding unassignment
    // left them (undefined) until a later move touches the domain again. Biases on them are read at the end of the move
    // like any other (SystemBiases::calc_move): the reference can never register a bias function as per-domain - its
    // dependency test reads the empty "bias_funcs" entry as one tag without domains (bias_functions.cpp:371-384,
    // utility.cpp:238-248) - so calc_one_domain / check_one_domain always return 0 and no candidate check carries a bias.
    // check_param's writes to m_defined are never read before the next calc_param of the same parameter.
    // -----------------------------------------------------------------------------------------
    LDO_HD bool pd_active(int d) const { return OBC().n_pd_ops != 0 && S()->dchain[d] == 0 && !BSP()->pd_disabled; }
    LDO_HDN void pd_update(int d) {
        BiasState* bs = BSP();
        int di_ = S()->dindex[d];
#pragma unroll 1
        for (int i = 0; i < OBC().n_ops; i++) {
            const OpDef& o = OBC().ops[i];
            if (!o.per_domain || (o.arg != di_ && o.arg2 != di_)) continue;
            const DomRec& a = S()->dom[o.arg];
            const DomRec& b = S()->dom[o.arg2];
            if (a.state == ST_UNASSIGNED || b.state == ST_UNASSIGNED) {
                bs->op_undefined |= 1u << i;
            }
            else {
                bs->op_undefined &= ~(1u << i);
                int dist = abssum(rec_pos(b) - rec_pos(a));
                bs->op_val[i] = o.type == OP_DIST ? dist : (dist == 1 ? 1 : 0);
            }
        }
    }

    // ---- mutation ----
    LDO_HD void write_dom(int d, V3 p, int o) {
        DomRec& r = S()->dom[d];
        r.k = p.k;
        r.ore = (int8_t)o;
    }

    // update_domain + update_occupancies (origami_system.cpp:762-806)
    LDO_HDN void commit_place(int d, V3 p, int o) {
        write_dom(d, p, o);
        int j = occupant(p);
        if (j < 0) {
            S()->dom[d].state = ST_UNBOUND;
            S()->bound[d] = -1;
            table_put(p, d);
            return;
        }
        if (S()->dom[j].state != ST_UNBOUND) {
            fail(LDO_ERR_BIND_BOUND, d);
            return;
        }
        S()->num_bound_pairs += 1;
        uint8_t ns;
        if (S()->ident[d] == -S()->ident[j]) {
            ns = ST_BOUND;
            S()->num_fully_bound_pairs += 1;
        }
        else {
            if (S()->dchain[d] == S()->dchain[j]) S()->num_self_bound_pairs += 1;
            ns = ST_MISBOUND;
        }
        S()->dom[d].state = ns;
        S()->dom[j].state = ns;
        S()->bound[d] = (short)j;
        S()->bound[j] = (short)d;
    }

    // OrigamiSystem::set_domain_config (origami_system.cpp:517-541). Sets constraints_violated.
    LDO_HDN double set_domain_config_base(int d, V3 p, int o) {
        if (S()->dom[d].state != ST_UNASSIGNED) {
            fail(LDO_ERR_SET_ASSIGNED, d);
            return 0;
        }
        int ns, partner;
        DeltaConfig dc = eval_place(d, p, o, &ns, &partner);
        if (ns == ST_UNASSIGNED) {
            // site already bound/misbound: position and orientation are left untouched (:838-845)
            S()->constraints_violated = 1;
            return dc.e;
        }
        if (partner < 0) {
            // unassigned site: the sticky violation flag is NOT cleared (App. A8, :852-855)
            if (S()->constraints_violated) {
                write_dom(d, p, o); // reverted by internal_unassign_domain, pos/ore stay
                return dc.e;
            }
            commit_place(d, p, o);
            S()->num_unassigned--;
            return dc.e;
        }
        S()->constraints_violated = dc.violated ? 1 : 0;
        if (dc.violated) {
            write_dom(d, p, o);
            return dc.e;
        }
        commit_place(d, p, o);
        S()->energy += dc.e;
        S()->num_stacked_pairs += dc.stacked;
        S()->num_unassigned--;
        return dc.e;
    }

    // OrigamiSystem::check_domain_constraints (origami_system.cpp:343-355): side-effect free apart
    // from the violation flag and (as in the reference) the stored position/orientation of d.
    LDO_HDN double check_domain_constraints(int d, V3 p, int o) {
        // OrigamiSystemWithBias::check_domain_constraints (origami_system.cpp:892-921) adds the per-domain bias check
        int ns, partner;
        DeltaConfig dc = eval_place(d, p, o, &ns, &partner);
        if (ns == ST_UNASSIGNED) {
            S()->constraints_violated = 1;
                return dc.e;
        }
        write_dom(d, p, o);
        if (partner >= 0) S()->constraints_violated = dc.violated ? 1 : 0;
        return dc.e;
    }

    // OrigamiSystem::set_checked_domain_config (origami_system.cpp:478-515)
    LDO_HDS double set_checked_domain_config_base(int d, V3 p, int o) {
        commit_place(d, p, o);
        if (S()->weight_pass) {
            S()->num_unassigned--;
            return 0;
        }
        double delta_e = 0;
        int st = S()->dom[d].state;
        if (st == ST_MISBOUND) {
            delta_e += hyb_energy(d, S()->bound[d]);
        }
        else if (st == ST_BOUND) {
            delta_e += hyb_energy(d, S()->bound[d]);
            DeltaConfig dc = check_stacking(d, S()->bound[d]);
            delta_e += dc.e;
            S()->num_stacked_pairs += dc.stacked;
        }
        if (SC().apply_mean_field_cor && st == ST_BOUND) {
            if (S()->num_fully_bound_pairs == 1) delta_e += 2 * log(6.0);
            else if (S()->num_fully_bound_pairs == 2) delta_e += log(3.0);
        }
        S()->energy += delta_e;
        S()->num_unassigned--;
        return delta_e;
    }

    // internal_unassign_domain + unassign_domain (origami_system.cpp:375-385, 695-760)
    LDO_HDS double unassign_domain_base(int d) {
        int st = S()->dom[d].state;
        double e = 0;
        int stacked = 0;
        if (st == ST_BOUND || st == ST_MISBOUND) {
            int j = S()->bound[d];
            bool with_energy = S()->weight_pass == 0;
            S()->num_bound_pairs -= 1;
            if (st == ST_BOUND) {
                S()->num_fully_bound_pairs -= 1;
                if (with_energy) {
                    DeltaConfig dc = check_stacking(d, j);
                    e = -dc.e;
                    stacked = -dc.stacked;
                }
            }
            else if (S()->dchain[j] == S()->dchain[d]) {
                S()->num_self_bound_pairs -= 1;
            }
            if (with_energy) e += -hyb_energy(d, j);
            S()->bound[d] = -1;
            S()->bound[j] = -1;
            S()->dom[d].state = ST_UNASSIGNED;
            S()->dom[j].state = ST_UNBOUND;
            const DomRec& r = S()->dom[d];
            table_put(rec_pos(r), j);
            if (with_energy && SC().apply_mean_field_cor && st == ST_BOUND) {
                if (S()->num_fully_bound_pairs == 0) e -= 2 * log(6.0);
                else if (S()->num_fully_bound_pairs == 1) e -= log(3.0);
            }
        }
        else if (st == ST_UNBOUND) {
            const DomRec& r = S()->dom[d];
            table_erase(rec_pos(r));
            S()->dom[d].state = ST_UNASSIGNED;
        }
        else {
            // double unassignment is allowed (:720-724): the two counter updates cancel
            S()->num_unassigned--;
        }
        S()->energy += e;
        S()->num_stacked_pairs += stacked;
        S()->num_unassigned++;
        return e;
    }

    // OrigamiSystemWithBias::set_checked_domain_config / unassign_domain (origami_system.cpp:928-945); set_domain_config
    // (:947-954) only adds calc_one_domain, which is always 0 (see above)
    LDO_HD double set_domain_config(int d, V3 p, int o) { return set_domain_config_base(d, p, o); }
    LDO_HDC double set_checked_domain_config(int d, V3 p, int o) {
        double e = set_checked_domain_config_base(d, p, o);
        if (pd_active(d)) pd_update(d);
        return e;
    }
    LDO_HDC double unassign_domain(int d) {
        double e = unassign_domain_base(d);
        if (pd_active(d)) pd_update(d);
        return e;
    }

    // ---- chains ----
    // add_chain(c_i_ident, c_i) (origami_system.cpp:402-443). Returns the chain slot.
    LDO_HDN int add_chain_with_uid(int type, int uid) {
        int c = -1;
#pragma unroll 1
        for (int k = 1; k < K::C; k++) {
            if (!S()->chain_used[k]) {
                c = k;
                break;
            }
        }
        int len = SC().type_len[type];
        if (c < 0 || len > SC().lmax || chain_base(c) + len > K::D) {
            fail(LDO_ERR_CAPACITY, type);
            return -1;
        }
        S()->chain_used[c] = 1;
        S()->chain_uid[c] = uid;
        S()->chain_type[c] = (uint16_t)type;
        S()->chain_len[c] = (uint16_t)len;
        S()->order[S()->n_chains] = (uint16_t)c;
        S()->n_chains++;
        S()->type_count[type]++;
        S()->num_staples++;
        int base = chain_base(c);
#pragma unroll 1
        for (int i = 0; i < len; i++) {
            int d = base + i;
            S()->dom[d].k = 0;
            S()->dom[d].ore = ORE_ZERO;
            S()->dom[d].state = ST_UNASSIGNED;
            S()->dom[d].link = chain_link_flags(i, len, false);
            S()->bound[d] = -1;
            S()->ident[d] = SC().idents[SC().type_off[type] + i];
            S()->dchain[d] = (uint16_t)c;
            S()->dindex[d] = (uint16_t)i;
            S()->num_domains++;
            S()->num_unassigned++;
        }
        return c;
    }
    // add_chain(c_i_ident) (origami_system.cpp:387-400)
    LDO_HD int add_chain(int type) {
        S()->current_c_i += 1;
        if (SC().apply_mean_field_cor) S()->energy += log(6.0);
        S()->energy += TT().init_energy;
        return add_chain_with_uid(type, S()->current_c_i);
    }
    // delete_chain (origami_system.cpp:445-472); the chain's domains must be unassigned
    LDO_HDN void delete_chain(int c) {
        int w = -1;
#pragma unroll 1
        for (int k = 0; k < S()->n_chains; k++) {
            if (S()->order[k] == c) {
                w = k;
                break;
            }
        }
        if (w < 0) {
            fail(LDO_ERR_INTERNAL, c);
            return;
        }
#pragma unroll 1
        for (int k = w; k + 1 < S()->n_chains; k++) S()->order[k] = S()->order[k + 1];
        S()->n_chains--;
        int len = S()->chain_len[c];
        S()->type_count[S()->chain_type[c]]--;
        S()->num_domains -= len;
        S()->num_staples--;
        S()->num_unassigned -= len;
        S()->chain_used[c] = 0;
        if (SC().apply_mean_field_cor) S()->energy -= log(6.0);
        S()->energy -= TT().init_energy;
    }

    // k-th staple of a given identity in insertion order (m_identity_to_index[type][k]; App. B)
    LDO_HD int staple_of_type(int type, int k) const {
#pragma unroll 1
        for (int w = 1; w < S()->n_chains; w++) {
            int c = S()->order[w];
            if (S()->chain_type[c] == type) {
                if (k == 0) return c;
                k--;
            }
        }
        return -1;
    }

    // Domain at position `index` of the concatenation of chains in working order (movetypes.cpp:98-113)
    LDO_HD int domain_by_flat_index(int index) const {
#pragma unroll 1
        for (int w = 0; w < S()->n_chains; w++) {
            int c = S()->order[w];
            int len = S()->chain_len[c];
            if (index < len) return chain_base(c) + index;
            index -= len;
        }
        return -1;
    }

    // ---- whole-system passes ----
    // OrigamiSystem::center (origami_system.cpp:553-571)
    LDO_HDN void center(int centering_domain) {
        const DomRec& c0 = S()->dom[chain_base(0) + centering_domain];
        V3 ref = rec_pos(c0);
        table_clear();
#pragma unroll 1
        for (int w = 0; w < S()->n_chains; w++) {
            int c = S()->order[w];
            int base = chain_base(c);
#pragma unroll 1
            for (int i = 0; i < S()->chain_len[c]; i++) {
                DomRec& r = S()->dom[base + i];
                r.k -= ref.k;
                if (r.state != ST_UNASSIGNED) table_put(rec_pos(r), base + i);
            }
        }
        // a site shared by a bound pair must resolve to a valid occupant: either is fine
    }

    // set_all_domains() + check_distance_constraints (origami_system.cpp:357-373, 573-586)
    LDO_HDN bool set_all_domains() {
#pragma unroll 1
        for (int w = 0; w < S()->n_chains; w++) {
            int c = S()->order[w];
            int base = chain_base(c);
#pragma unroll 1
            for (int i = 0; i < S()->chain_len[c]; i++) {
                int d = base + i;
                const DomRec r = S()->dom[d];
                set_domain_config(d, rec_pos(r), r.ore);
                if (S()->constraints_violated) {
                    fail(LDO_ERR_CONSTRAINTS, d);
                    return false;
                }
            }
        }
#pragma unroll 1
        for (int w = 0; w < S()->n_chains; w++) {
            int c = S()->order[w];
            int base = chain_base(c);
#pragma unroll 1
            for (int i = 0; i < S()->chain_len[c]; i++) {
                int d = base + i;
                int n = step(d, 1);
                if (n < 0) continue;
                const DomRec& a = S()->dom[d];
                const DomRec& b = S()->dom[n];
                if (abssum(rec_pos(b) - rec_pos(a)) != 1) {
                    fail(LDO_ERR_DISTANCE, d);
                    return false;
                }
            }
        }
        return true;
    }

    LDO_HD void unassign_all() {
#pragma unroll 1
        for (int w = 0; w < S()->n_chains; w++) {
            int c = S()->order[w];
            int base = chain_base(c);
#pragma unroll 1
            for (int i = 0; i < S()->chain_len[c]; i++) unassign_domain(base + i);
        }
    }

    // OrigamiSystem::check_all_constraints (origami_system.cpp:267-325)
    LDO_HDN bool check_all_constraints() {
        if (S()->num_unassigned != 0) {
            fail(LDO_ERR_UNASSIGNED_AT_CHECK);
            return false;
        }
#pragma unroll 1
        for (int w = 0; w < S()->n_chains; w++) {
            int c = S()->order[w];
            int base = chain_base(c);
#pragma unroll 1
            for (int i = 0; i < S()->chain_len[c]; i++) {
                if (S()->dom[base + i].state == ST_UNASSIGNED) {
                    fail(LDO_ERR_UNASSIGNED_AT_CHECK, base + i);
                    return false;
                }
                unassign_domain(base + i);
            }
        }
        int ns = S()->n_chains - 1;
        if (SC().apply_mean_field_cor) S()->energy -= ns * log(6.0);
        S()->energy -= ns * TT().init_energy;
        if (S()->num_stacked_pairs != 0) {
            int sp = S()->num_stacked_pairs;
            set_all_domains();
            fail(LDO_ERR_STACK_COUNT, sp);
            return false;
        }
        double eps = 0.000001;
        if (S()->energy < -eps || S()->energy > eps) {
            set_all_domains();
            fail(LDO_ERR_ENERGY_DRIFT);
            return false;
        }
        S()->energy = 0;
        if (!set_all_domains()) return false;
        if (SC().apply_mean_field_cor) S()->energy += ns * log(6.0);
        S()->energy += ns * TT().init_energy;
        return true;
    }

    // OrigamiSystem::update_energy (origami_system.cpp:808-826), after new tables were installed
    LDO_HDN bool update_energy() {
        unassign_all();
        S()->energy = 0;
        int ns = S()->n_chains - 1;
        if (SC().apply_mean_field_cor) S()->energy += ns * log(6.0);
        S()->energy += ns * TT().init_energy;
        return set_all_domains();
    }

    // OrigamiSystem::update_enthalpy_and_entropy (origami_system.cpp:204-246)
    LDO_HDN void enthalpy_and_entropy(double* enthalpy, double* entropy, double* stacking) const {
        double H = 0, Sent = 0;
#pragma unroll 1
        for (int w = 0; w < S()->n_chains; w++) {
            int c = S()->order[w];
            int base = chain_base(c);
#pragma unroll 1
            for (int i = 0; i < S()->chain_len[c]; i++) {
                int d = base + i;
                int st = S()->dom[d].state;
                if (st == ST_BOUND || st == ST_MISBOUND) {
                    int j = S()->bound[d];
                    // the pair is counted when its first member (in working order) is visited
                    bool j_first = false;
                    int cj = S()->dchain[j];
                    if (cj == c) {
                        j_first = S()->dindex[j] < i;
                    }
                    else {
#pragma unroll 1
                        for (int w2 = 0; w2 < w; w2++) {
                            if (S()->order[w2] == cj) {
                                j_first = true;
                                break;
                            }
                        }
                    }
                    if (!j_first) {
                        H += hyb_enthalpy(d, j);
                        Sent += hyb_entropy(d, j);
                    }
                }
            }
        }
        int ns = S()->n_chains - 1;
        if (SC().apply_mean_field_cor) {
            Sent -= ns * log(6.0);
            if (S()->num_fully_bound_pairs >= 1) Sent -= 2 * log(6.0);
            if (S()->num_fully_bound_pairs >= 2) Sent -= log(3.0);
        }
        H += ns * TT().init_enthalpy;
        Sent += ns * TT().init_entropy;
        *enthalpy = H;
        *entropy = Sent;
        *stacking = S()->energy - (H - Sent);
    }
};

} // namespace ldo
