// ldo_moves.cuh — random numbers, order parameters / biases, topology constraint points and the
// Monte Carlo movetypes of the LatticeDNAOrigami model, one warp per replica (see ldo_core.cuh for
// the execution model).
//
// What this restates (reference file:line):
//   * RandomGens::uniform_real / uniform_int                      random_gens.cpp:29-49
//     -> counter-based Philox4x32-10, or a value-level replay tape for parity
//   * SystemOrderParams::update_move_params, move-update biases   order_params.cpp:143-445,595-601
//                                                                 bias_functions.cpp:139-283,447-487
//   * MCMovetype skeleton: attempt / reset_origami / acceptance   movetypes.cpp:41-277
//   * OrientationRotation                                         orientation_movetype.cpp:30-65
//   * MetStapleExchange, MetStapleRegrowth, growth helpers        met_movetypes.cpp:50-513, movetypes.cpp:322-395
//   * CBStapleRegrowth (6-site Rosenbluth weights over lanes)     cb_movetypes.cpp:44-404
//   * StapleNetwork + Constraintpoints                            top_constraint_points.cpp:36-601
//   * IdealRandomWalks::num_walks == 0 predicate                  ideal_random_walk.cpp:14-73 (App. A7)
//   * CT segment selection (contiguous, non-contiguous)           movetypes.cpp:449-719
//   * CTRG recoil growth: scaffold + jump variants                rg_movetypes.cpp:14-856
//   * GCMCSimulation::simulate step                               simulation.cpp:568-665
#pragma once

#include "ldo_core.cuh"

namespace ldo {

// ---------------------------------------------------------------------------------------------
// Random numbers
// ---------------------------------------------------------------------------------------------

struct TapeDraw {
    int32_t kind; // 0 = uniform_real, 1 = uniform_int
    int32_t lo, hi, ival;
    double real;
};

#ifndef LDO_PHILOX_BLOCKS
#define LDO_PHILOX_BLOCKS 8
#endif
// LDO_NO_TAPE builds (profiling variants) drop the replay branches from the draw functions
#ifdef LDO_NO_TAPE
#define LDO_TAPE_MODE(g) false
#else
#define LDO_TAPE_MODE(g) ((g)->tape != nullptr)
#endif
// true when the move code must follow the reference's serial draw order: replay, or Philox with the
// reference_draw_order switch of the moveset
#define LDO_SERIAL_DRAWS() (LDO_TAPE_MODE(RNG()) || MS().reference_draw_order != 0)
struct Rng {
    // replay tape (parity mode) when tape != nullptr
    const TapeDraw* tape;
    long long tape_len;
    long long tape_pos;
    // Philox4x32-10: key = seed, counter = (draw index lo, draw index hi, replica id, stream)
    uint32_t key0, key1;
    uint32_t subseq;
    uint32_t stream;
    unsigned long long counter;
    // unconsumed words of the last refill: LDO_PHILOX_BLOCKS blocks of 4 words, block j computed by lane j
    // from counter + j (a real takes 2 words, an int 1)
    uint32_t buf[4 * LDO_PHILOX_BLOCKS];
    int buf_n;
};

LDO_HD inline void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
    unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
    unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
}

LDO_HD inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t subseq, uint32_t stream, unsigned long long ctr, uint32_t* out) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = subseq, c3 = stream;
#pragma unroll 1
    for (int i = 0; i < 10; i++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}
LDO_HD inline void philox4x32_10(const Rng& r, unsigned long long ctr, uint32_t* out) {
    philox4x32_10(r.key0, r.key1, r.subseq, r.stream, ctr, out);
}

// ---------------------------------------------------------------------------------------------
// Moveset
// ---------------------------------------------------------------------------------------------

enum {
    MT_ORIENTATION_ROTATION = 0,
    MT_MET_STAPLE_EXCHANGE = 1,
    MT_MET_STAPLE_REGROWTH = 2,
    MT_CB_STAPLE_REGROWTH = 3,
    MT_CTCB_SCAFFOLD_REGROWTH = 4,
    MT_CTCB_JUMP_SCAFFOLD_REGROWTH = 5,
    MT_CTRG_SCAFFOLD_REGROWTH = 6,
    MT_CTRG_JUMP_SCAFFOLD_REGROWTH = 7,
    MT_CTCB_LINKER_REGROWTH = 8,
    MT_CTCB_CLUSTERED_LINKER_REGROWTH = 9,
    MT_CTRG_LINKER_REGROWTH = 10
};
#define LDO_MAX_TRANSFORMS 16 // num_transforms capacity of the linker moves

#define LDO_MAX_MOVETYPES 12
struct MoveDef {
    int type;
    double cum_prob; // m_cumulative_probs (simulation.cpp:233-237)
    int max_regrowth, max_seg_regrowth, max_num_recoils, max_c_attempts;
    int max_disp, max_turns, max_linker_length, num_transforms; // linker / transform moves
    int adaptive_exchange;
    int exchange_mults_off; // offset into MoveSet::exchange_mults
};
struct MoveSet {
    int n;
    int allow_nonsensical_ps;
    // Production (Philox) mode only: 1 = keep the reference's serial trial order in the three places where the
    // draws are otherwise re-associated to lanes (rg_select_open_config, the single-draw feeler test,
    // rg_count_avail_parallel), i.e. run Philox through exactly the branches a replay tape runs. Same ensemble,
    // used by the tests to isolate the lane-parallel branches (ldo_set_reference_draw_order).
    int reference_draw_order;
    MoveDef mt[LDO_MAX_MOVETYPES];
    double exchange_mults[LDO_MAX_TYPES];
};

// Extended-precision accumulator for the configurational-bias Rosenbluth weights. The reference keeps
// m_bias / m_new_bias in x87 long double (cb_movetypes.hpp:126) and only the final min(1, ratio) is
// rounded to double (movetypes.cpp:142-143); whether that rounds to exactly 1 decides if a random number
// is drawn, so fp64 products are not enough for replay parity. A double-double (hi + lo, ~106 bits,
// error-free products through fma) rounds to the same double as the 64-bit x87 mantissa except within
// 2^-64 of a rounding boundary.
struct DD {
    double hi, lo;
};
LDO_HD inline DD dd_from(double a) {
    DD r;
    r.hi = a;
    r.lo = 0;
    return r;
}
LDO_HD inline DD dd_renorm(double a, double b) {
    DD r;
    r.hi = a + b;
    r.lo = b - (r.hi - a);
    return r;
}
LDO_HD inline DD dd_mul(DD a, double b) {
    double p = a.hi * b;
    double e = fma(a.hi, b, -p) + a.lo * b;
    return dd_renorm(p, e);
}
LDO_HD inline DD dd_mul_dd(DD a, DD b) {
    double p = a.hi * b.hi;
    double e = fma(a.hi, b.hi, -p) + (a.hi * b.lo + a.lo * b.hi);
    return dd_renorm(p, e);
}
LDO_HD inline DD dd_add(DD a, double b) {
    double s = a.hi + b;
    double bb = s - a.hi;
    double e = (a.hi - (s - bb)) + (b - bb);
    return dd_renorm(s, e + a.lo);
}
// double(a / b) for a double a and a double-double b (the reference divides two long doubles)
LDO_HD inline double dd_quot(double a, DD b) {
    double q1 = a / b.hi;
    double p = q1 * b.hi;
    double e = fma(q1, b.hi, -p) + q1 * b.lo;
    double r = (a - p) - e;
    return q1 + r / b.hi;
}
LDO_HD inline DD dd_div(DD a, double b) {
    double q1 = a.hi / b;
    double r = fma(-q1, b, a.hi) + a.lo;
    return dd_renorm(q1, r / b);
}
// a / b rounded like the reference's (long double ratio -> fmin(1, .) -> double)
LDO_HD inline double dd_ratio_min1(DD a, DD b) {
    double q1 = a.hi / b.hi;
    // remainder a - q1 * b in double-double
    double p = q1 * b.hi;
    double e = fma(q1, b.hi, -p) + q1 * b.lo;
    double r = (a.hi - p) - e + a.lo;
    DD q = dd_renorm(q1, r / b.hi);
    if (q.hi > 1.0 || (q.hi == 1.0 && q.lo >= 0)) return 1.0;
    return q.hi;
}

// 36-bit configuration masks: population count and position of the n-th (0-based) set bit
LDO_HD inline int popc36(unsigned long long m) {
#if defined(__CUDA_ARCH__)
    return __popcll(m);
#else
    int n = 0;
#pragma unroll 1
    for (; m; m &= m - 1) n++;
    return n;
#endif
}
LDO_HD inline int nth_set_bit36(unsigned long long m, int n) {
#if defined(__CUDA_ARCH__)
    // rank-select by halving with popc (the __fns intrinsic expands to a long software routine)
    unsigned x = (unsigned)m;
    int pos = 0;
    int nlo = __popc(x);
    if (n >= nlo) {
        n -= nlo;
        x = (unsigned)(m >> 32);
        pos = 32;
    }
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        unsigned low = x & ((1u << w) - 1u);
        int c = __popc(low);
        if (n >= c) {
            n -= c;
            x >>= w;
            pos += w;
        }
        else {
            x = low;
        }
    }
    return pos;
#else
#pragma unroll 1
    for (int k = 0; k < n; k++) m &= m - 1;
    int i = 0;
#pragma unroll 1
    while (i < 36 && !((m >> i) & 1ull)) i++;
    return i;
#endif
}

// Open probabilities of the 36 trial configurations of one regrown domain, stored per neighbour
// site of the reference domain (the 6 orientations of a site share the lattice lookup):
//   kind 0: site blocked or binding violates a constraint -> p = 0 for all orientations
//   kind 1: empty site -> p (1, or 0 when no ideal walk remains) for every orientation
//   kind 2: site holds an unbound domain -> p for the single opposing orientation `ore`, else 0
struct RgSlot {
    double p[6];
    uint8_t kind[6];
    int8_t ore[6];
};
#define LDO_RG_OWN_SLOTS 4
#define LDO_RG_SLOTS (LDO_RG_OWN_SLOTS + 6)

#if defined(__CUDACC__)
__constant__ MoveSet ldo_c_ms;
#endif

// ---------------------------------------------------------------------------------------------
// Per-move scratch (MCMovetype / RegrowthMCMovetype / CBMCMovetype / CTRGRegrowthMCMovetype members)
// and Constraintpoints (top_constraint_points.hpp:112-288), fixed capacity
// ---------------------------------------------------------------------------------------------

// Effect of placing a parent domain on an EMPTY site q on the active endpoints (update_endpoints,
// top_constraint_points.cpp:340-349), applied on the fly by the read-only walk / endpoint predicates:
// endpoints of (rm_chain, rm_seg) with domain index rm_d are removed; one endpoint is added when the
// parent has an inactive endpoint.
struct EpOverlay {
    int rm_chain, rm_seg, rm_d; // rm_chain < 0: no overlay
    int add_chain, add_seg, add_d; // add_chain < 0: nothing added
    V3 add_pos;
};

// Hot part: touched inside the trial loops; lives in shared memory for staged replicas
template <class K>
struct MoveScratch {
    static const int A = 2 * K::LV + 8; // assigned-domain list capacity
    static const int E = K::E; // active endpoints capacity
    static const int S = K::SEG; // scaffold segments capacity

    // MCMovetype (movetypes.hpp:146-166)
    short modified[K::LV + 1];
    int n_modified;
    short assigned[A];
    int n_assigned;
    int added_chain; // chain slot or -1 (at most one chain is added per move)
    int rejected;
    double modifier;

    // CTRG (rg_movetypes.hpp:83-115)
    short regrow[K::LV + 1];
    int n_regrow;
    uint8_t c_attempts_q[K::LV + 1];
    unsigned long long avail_q[K::LV + 1];
    double c_opens[K::LV + 1];
    // m_erased_endpoints_q: stack of position lists
    int eq_depth;
    short eq_start[A + 1];
    int eq_npos;
    uint32_t eq_pos[E]; // packed positions (V3::k)

    // Constraintpoints
    int8_t seg_of[K::D]; // m_segs (-1 = absent)
    int8_t scaf_dir[S]; // m_domain_to_dir for (scaffold, seg); staples: seg 0 -> +1, seg 1 -> -1
    short stem_gp[K::D]; // m_stemdomains[d] (growthpoint of stem domain d) or -1
    short gp_stem[K::D]; // m_growthpoints[d] (domain to grow from d) or -1
    short inactive[K::D]; // m_inactive_endpoints[d] or -1
    int8_t stem_seg0[K::D]; // m_stemd_to_segs[d] = {s, s+1}, -1 = none
    int n_ep; // m_active_endpoints, flattened, per-key order preserved
    short ep_chain[E];
    int8_t ep_seg[E];
    short ep_d[E];
    uint32_t ep_pos[E];
    int n_erased; // m_erased_endpoints
    uint32_t erased_pos[8];

    // CTRG trial-configuration caches: own[level % 4] and feeler memo keyed by the parent's site
    RgSlot slots[LDO_RG_SLOTS];

    // lane-parallel candidate evaluation results (6 neighbour sites)
    double site_w[8];
    int8_t site_o[8];
    int8_t site_kind[8];
};

// Cold part: selection / topology set-up and the saved configurations, touched a few times per move;
// always in global memory (L2 resident)
template <class K>
struct ColdScratch {
    static const int E = K::E;
    static const int S = K::SEG;
    DomRec prev[K::D]; // m_prev_pos / m_prev_ore (written once per modified domain, read on rejection)
    DomRec oldc[K::D]; // m_old_pos / m_old_ore
    DomRec newc[K::D]; // m_new_pos / m_new_ore
    short sel_scaf[K::D + 1];
    int n_sel;
    uint8_t in_sel[K::D]; // membership in m_scaffold_domains
    uint8_t checked_chain[K::C]; // m_checked_staples
    int n_ep0; // m_initial_active_endpoints
    short ep0_chain[E];
    int8_t ep0_seg[E];
    short ep0_d[E];
    uint32_t ep0_pos[E];

    // StapleNetwork scratch (top_constraint_points.hpp:40-108)
    uint8_t net_chain[K::C];
    short net_growth_idx[K::C];
    int n_pot_gps, n_pot_iaes, n_pot_ds;
    short pot_gps[K::D + 1][2];
    short pot_iaes[K::D + 1][2];
    short pot_ds[K::D + 1];
    short scan_stack[K::C][2];

    // non-contiguous selection (movetypes.cpp:512-648)
    short seg_start[S + 1]; // segment s occupies seg_dom[seg_start[s] .. seg_start[s+1])
    short seg_dom[2 * K::D + 2];
    short stems[K::D + 1];
    short stem_queue[4 * K::D + 8];

    // CTRG: m_c_attempts_wq / m_avail_cis_wq (rg_movetypes.hpp:107-108), written twice and read once per level per move
    uint8_t c_attempts_wq[K::LV + 1];
    unsigned long long avail_wq[K::LV + 1];
    // Trial-probability slots of every level as computed when the level's configuration was last chosen (growth) or
    // examined (old configuration): calc_weights meets every level again in exactly that environment - the domains
    // before it placed, the ones after it unassigned, the same active endpoints - so the six site evaluations (and the
    // binding evaluations among them) need not be repeated. Pure-function caching: same values, same draws.
    RgSlot slot_cache[K::LV + 1];
    uint8_t slot_cached[K::LV + 1];
    // energy change and stacked-pair change of every level's placement during growth: a recoil takes the placement
    // back in the environment it was made in, so its energy change is the negative of these (no second evaluation)
    double set_de[K::LV + 1];
    short set_sp[K::LV + 1];

    // CTCB: chains of the internally bound staple networks (m_regrowth_staples) and the growth work stack
    uint8_t regrow_chain[K::C];
    short work[K::LV + 1][4];

    // Linker regrowth (transform_movetypes.hpp:92-100): the two linkers (element 0 = end of the central
    // segment), their fixed endpoints, the central segment's domains with those of its staples, and the
    // accepted trial transformations of transform_segment
    short lnk[2][K::LV + 1];
    int n_lnk[2];
    short lnk_ep[2]; // m_linker_endpoints (-1 = none)
    short cen_dom[K::D + 1]; // central_domains; the first C()->n_sel entries are sel_scaf (central_segment)
    int n_cen;
    uint8_t lk_mark[K::D]; // bit 0: in central_segment, bit 1: in central_domains
    uint8_t cen_chain[K::C]; // central_staples (find_staples)
    int n_tf;
    uint32_t tf_center[LDO_MAX_TRANSFORMS], tf_disp[LDO_MAX_TRANSFORMS];
    int8_t tf_axis[LDO_MAX_TRANSFORMS], tf_turns[LDO_MAX_TRANSFORMS];
    double tf_bfactor[LDO_MAX_TRANSFORMS];
};

// Move statistics (MovetypeTracking, movetypes.hpp:49-52)
struct MoveStats {
    long long attempts[LDO_MAX_MOVETYPES];
    long long accepts[LDO_MAX_MOVETYPES];
    // MetStapleExchangeMCMovetype::m_exchange_mults of the movetypes with adaptive_exchange (met_movetypes.cpp:127,
    // 232-234, 279-282): state of the movetype object, i.e. of the replica; laid out like MoveSet::exchange_mults
    double exchange_mults[LDO_MAX_TYPES];
};

// Typed move trackers (m_tracker / m_tracking of the movetypes, movetypes.hpp:327-339, utility.hpp:100-147): what the
// reference breaks its .moves summary down by - staple type for the staple moves (insertions and deletions apart),
// number of scaffold domains (and of staples) for the scaffold regrowth moves. Optional (ldo_enable_move_trackers): kept
// outside RepAux so that checkpoints and the run kernel's staged state do not grow. The tracker fields of a movetype
// object persist between its moves in the reference (a field a move does not set keeps its last value): sticky_a/_b.
// The transform / linker movetypes track six numbers (CTCBLinkerRegrowthTracking, utility.hpp:136-144: linker domains,
// linker staples, central domains, central staples, sum of the displacement, turns; central_domains_connected is never
// set) and summarise them as three tables keyed by pairs (transform_movetypes.cpp:64-139): the pairs met so far are kept in
// a list (lk_key = movetype, table, two 13-bit values), the fields of the move under way in lk_now / lk_set.
#define LDO_TRK_BINS 64
#define LDO_TRK_LK_CAP 1024
struct TrackStats {
    int sticky_a[LDO_MAX_MOVETYPES], sticky_b[LDO_MAX_MOVETYPES];
    unsigned cnt[LDO_MAX_MOVETYPES][2][LDO_TRK_BINS][2]; // [movetype][field][value][attempts, accepts]
    int lk_sticky[LDO_MAX_MOVETYPES][6];
    int lk_now[6];
    unsigned lk_set;
    int lk_n, lk_dropped;
    unsigned lk_key[LDO_TRK_LK_CAP];
    unsigned lk_cnt[LDO_TRK_LK_CAP][2];
};
LDO_HD inline unsigned trk_lk_key(int movetype, int table, int a, int b) {
    return ((unsigned)movetype << 28) | ((unsigned)table << 26) | (((unsigned)a & 0x1fffu) << 13) | ((unsigned)b & 0x1fffu);
}

// ---------------------------------------------------------------------------------------------
// The per-replica engine
// ---------------------------------------------------------------------------------------------

template <class K>
struct Engine {
    System<K> sys; // first member: SmemLayout<K>::engine is also the offset of the System object
    MoveScratch<K>* m;
    ColdScratch<K>* mc;
    Rng* rng;
    const MoveSet* ms;
    const OpsBiasConst* ob;
    BiasState* bs;
    const double* grid_vals; // replica's grid-bias values (NaN = off grid)
    Control ctl;
    MoveStats* stats;
    TrackStats* trk; // null unless trackers are enabled

    struct Work {
        // RG per-domain working state (rg_movetypes.hpp:118-126)
        int di, d, ref_d, stemd, dir, c_attempts, d_max_c_attempts;
        unsigned long long avail;
        int max_recoils, max_c_attempts;
        double delta_e, weight, weight_new;
        // slot holding the current domain's trial probabilities; feeler memo bookkeeping
        int cur_slot, memo_level, memo_key, memo_mask, last_pc, last_kind;
        int slot_cache_on; // calc_weights may take the slots of C()->slot_cache (see there)
        int in_growth; // inside recoil_regrow's own growth (not a feeler): recoils may reuse the level's cached slot / energy
        int trk_a, trk_b; // tracker fields of the move under way (TrackStats)
        int sp_start; // stacked pairs / energy before the move (mc_step: rejected moves restore them)
        double e_start;
    };
    Work wk;

    // ---- accessors: shared-memory hints and constant-memory copies on the device ----
    LDO_HD MoveScratch<K>* M() const {
        return LDO_SMEM_PTR(K, MoveScratch<K>, scratch, m);
    }
#define LDO_ENG_FIELD(T, member) (*LDO_SMEM_AT(K, T, SmemLayout<K>::engine + offsetof(Engine<K>, member), const_cast<T*>(&member)))
    LDO_HD Work* W() const { return &LDO_ENG_FIELD(Work, wk); }
    LDO_HD ColdScratch<K>* C() const { return LDO_ENG_FIELD(ColdScratch<K>*, mc); }
    LDO_HD const Control& CTL() const { return LDO_ENG_FIELD(Control, ctl); }
    LDO_HD const double* GV() const { return LDO_ENG_FIELD(const double*, grid_vals); }
    LDO_HD Rng* RNG() const {
        return LDO_SMEM_PTR(K, Rng, rng, rng);
    }
    LDO_HD BiasState* BS() const {
        return LDO_SMEM_PTR(K, BiasState, bias, bs);
    }
    LDO_HD MoveStats* STATS() const { return LDO_ENG_FIELD(MoveStats*, stats); } // global memory: touched twice per move
#ifdef LDO_NO_TRACKERS // A/B knob (profiles/ab_r2.txt): the tracker hooks compiled out
    LDO_HD TrackStats* TRK() const { return nullptr; }
#elif defined(LDO_ONE_KERNEL) // A/B knob: one kernel with the hooks tested at run time (as before Tracked<K>)
    LDO_HD TrackStats* TRK() const { return LDO_ENG_FIELD(TrackStats*, trk); }
#else
    LDO_HD TrackStats* TRK() const {
        if (!K::TRACK) return nullptr; // compile-time: the hooks exist in the Tracked<K> instantiation only
        return LDO_ENG_FIELD(TrackStats*, trk);
    }
#endif
    LDO_HD const MoveSet& MS() const {
#if defined(__CUDA_ARCH__)
        return ldo_c_ms;
#else
        return *ms;
#endif
    }
    LDO_HD const OpsBiasConst& OB() const {
#if defined(__CUDA_ARCH__)
        return ldo_c_ob;
#else
        return *ob;
#endif
    }

    // ---- RNG (random_gens.cpp:29-49) ----
    // The two draw functions are out-of-line calls (LDO_HDC): inlined at their ~60 call sites they were 42 KB of code
    LDO_HDN void philox_refill() {
        Rng* g = RNG();
        unsigned long long c = g->counter;
#pragma unroll 1
        for (int j = LDO_LANE; j < LDO_PHILOX_BLOCKS; j += LDO_NLANES) philox4x32_10(g->key0, g->key1, g->subseq, g->stream, c + j, g->buf + 4 * j);
        LDO_SYNCWARP();
        g->counter = c + LDO_PHILOX_BLOCKS;
        g->buf_n = 4 * LDO_PHILOX_BLOCKS;
        LDO_SYNCWARP();
    }
    LDO_HD uint32_t next_word() {
        Rng* g = RNG();
        if (g->buf_n == 0) philox_refill();
        return g->buf[--g->buf_n];
    }
    // replay tape (parity mode)
    LDO_HDN double tape_real() {
        Rng* g = RNG();
        if (g->tape_pos >= g->tape_len) {
            sys.fail(LDO_ERR_TAPE_EXHAUSTED);
            return 0.5;
        }
        const TapeDraw& t = g->tape[g->tape_pos];
        if (t.kind != 0) {
            sys.fail(LDO_ERR_TAPE_MISMATCH, (int)g->tape_pos);
            return 0.5;
        }
        g->tape_pos++;
        return t.real;
    }
    LDO_HDN int tape_int(int lo, int hi) {
        Rng* g = RNG();
        if (g->tape_pos >= g->tape_len) {
            sys.fail(LDO_ERR_TAPE_EXHAUSTED);
            return lo;
        }
        const TapeDraw& t = g->tape[g->tape_pos];
        if (t.kind != 1 || t.lo != lo || t.hi != hi) {
            sys.fail(LDO_ERR_TAPE_MISMATCH, (int)g->tape_pos);
            return lo;
        }
        g->tape_pos++;
        return t.ival;
    }
    LDO_HDC double uniform_real() {
        if (LDO_TAPE_MODE(RNG())) return tape_real();
        uint32_t hi = next_word();
        uint32_t lo = next_word();
        unsigned long long u = ((unsigned long long)hi << 32) | lo;
        return (double)(u >> 11) * (1.0 / 9007199254740992.0);
    }
    // Lemire's nearly-divisionless unbiased bounded integer: rejection part
    LDO_HDN unsigned long long uniform_int_reject(unsigned long long mm, uint32_t n) {
        uint32_t t = (0u - n) % n;
#pragma unroll 1
        while ((uint32_t)mm < t) mm = (unsigned long long)next_word() * n;
        return mm;
    }
    LDO_HDC int uniform_int(int lo, int hi) {
        if (LDO_TAPE_MODE(RNG())) return tape_int(lo, hi);
        uint32_t n = (uint32_t)(hi - lo) + 1u;
        unsigned long long mm = (unsigned long long)next_word() * n;
        if ((uint32_t)mm < n) mm = uniform_int_reject(mm, n);
        return lo + (int)(mm >> 32);
    }

    // ---- order parameters and biases ----
    LDO_HDS int calc_op(int i) const {
        const OpDef& o = OB().ops[i];
        const SysState<K>* s = sys.S();
        switch (o.type) {
        case OP_NUM_STAPLES: return s->num_staples;
        case OP_NUM_STAPLES_TYPE: return s->type_count[o.arg];
        case OP_STAPLE_TYPE_FULLY_BOUND: {
            // order_params.cpp:250-266
#pragma unroll 1
            for (int w = 1; w < s->n_chains; w++) {
                int c = s->order[w];
                if (s->chain_type[c] != o.arg) continue;
                bool full = true;
                int base = sys.chain_base(c);
#pragma unroll 1
                for (int k = 0; k < s->chain_len[c]; k++) {
                    if (s->dom[base + k].state != ST_BOUND) {
                        full = false;
                        break;
                    }
                }
                if (full) return 1;
            }
            return 0;
        }
        case OP_NUM_BOUND_DOMAIN_PAIRS: return s->num_fully_bound_pairs;
        case OP_NUM_MISBOUND_DOMAIN_PAIRS: return s->num_bound_pairs - s->num_fully_bound_pairs;
        case OP_NUM_STACKED_PAIRS: return s->num_stacked_pairs;
        case OP_NUM_LINEAR_HELICES: return 0; // never modified in the reference (App. A5)
        case OP_NUM_STACKED_JUNCTS: return 0;
        case OP_SUM: {
            // SumOrderParam::calc_param (order_params.cpp:143-161): undefined when a term is; keeps its value then
            int sum = 0;
            BS()->op_undefined &= ~(1u << i);
#pragma unroll 1
            for (int k = 0; k < o.n_sum; k++) {
                if ((BS()->op_undefined >> o.sum_idx[k]) & 1u) {
                    BS()->op_undefined |= 1u << i;
                    return BS()->op_val[i];
                }
                sum += BS()->op_val[o.sum_idx[k]];
            }
            return sum;
        }
        case OP_DIST:
        case OP_ADJACENT_SITE: {
            // calc_param (order_params.cpp:34-48, 84-104): defined when both domains are assigned; m_param
            // keeps its previous value otherwise
            const DomRec& a = s->dom[o.arg];
            const DomRec& b = s->dom[o.arg2];
            if (a.state == ST_UNASSIGNED || b.state == ST_UNASSIGNED) {
                BS()->op_undefined |= 1u << i;
                return BS()->op_val[i];
            }
            BS()->op_undefined &= ~(1u << i);
            int dist = abssum(rec_pos(b) - rec_pos(a));
            return o.type == OP_DIST ? dist : (dist == 1 ? 1 : 0);
        }
        }
        return 0;
    }
    // SystemOrderParams::update_move_params (order_params.cpp:595-601)
    LDO_HDC void update_move_params() {
#pragma unroll 1
        for (int i = 0; i < OB().n_ops; i++) {
            if (OB().ops[i].per_domain) continue; // updated with every domain placement (System::pd_update)
            BS()->op_val[i] = calc_op(i);
        }
    }
    LDO_HD double grid_lookup(int b) const {
        int off = BS()->grid_off[b];
        if (off < 0) return 0;
        const BiasDef& bd = OB().biases[b];
        int idx = 0;
#pragma unroll 1
        for (int k = 0; k < bd.n_ops; k++) {
            if ((BS()->op_undefined >> bd.op_idx[k]) & 1u) return 0; // bias_functions.cpp:258-262
            int v = BS()->op_val[bd.op_idx[k]] - BS()->grid_lo[b][k];
            if (v < 0 || v >= BS()->grid_n[b][k]) return 0;
            idx = idx * BS()->grid_n[b][k] + v;
        }
        double g = GV()[off + idx];
        return g == g ? g : 0; // NaN marks a point absent from the grid (m_off_grid_bias = 0)
    }
    LDO_HD double calc_bias_fn(int b) const {
        const BiasDef& bd = OB().biases[b];
        if (bd.type == BIAS_GRID) return grid_lookup(b);
        if ((BS()->op_undefined >> bd.op_idx[0]) & 1u) return 0; // update_bias: no bias while undefined (bias_functions.cpp:124-134)
        int param = BS()->op_val[bd.op_idx[0]];
        int lo = BS()->win_min[b], hi = BS()->win_max[b];
        if (bd.type == BIAS_LINEAR_STEP_WELL) {
            // bias_functions.cpp:139-151
            if (param < lo) return bd.slope * (lo - param - 1) + bd.min_bias;
            if (param > hi) return bd.slope * (param - hi - 1) + bd.min_bias;
            return bd.well_bias;
        }
        // bias_functions.cpp:194-204
        if (param < lo || param > hi) return bd.outside_bias;
        return bd.well_bias;
    }
    // SystemBiases::calc_move (bias_functions.cpp:475-487)
    LDO_HDN double calc_move_bias() {
        double diff = 0;
#pragma unroll 1
        for (int b = 0; b < OB().n_biases; b++) {
            double prev = BS()->bias_val[b];
            double nb = calc_bias_fn(b);
            BS()->bias_val[b] = nb;
            diff += nb - prev;
        }
        BS()->move_update_bias += diff;
        return diff * CTL().bias_mult;
    }
    // SystemBiases::get_total_bias (bias_functions.cpp:459-461); m_domain_update_bias is always 0
    LDO_HD double total_bias() const { return BS()->move_update_bias * CTL().bias_mult; }

    // ---- MCMovetype shared helpers (movetypes.cpp:98-158) ----
    LDO_HDC int select_random_domain() {
        int idx = uniform_int(0, sys.S()->num_domains - 1);
        return sys.domain_by_flat_index(idx);
    }
    LDO_HDC bool test_acceptance(double p_ratio) {
        double p_accept = fmin(1.0, p_ratio) * M()->modifier;
        if (p_accept == 1) return true;
        return p_accept > uniform_real();
    }
    LDO_HD void reset_internal() {
        M()->n_modified = 0;
        M()->n_assigned = 0;
        M()->added_chain = -1;
        M()->rejected = 0;
        M()->modifier = 1;
    }
    LDO_HDC void push_assigned(int dd) {
        if (M()->n_assigned >= MoveScratch<K>::A) {
            sys.fail(LDO_ERR_CAPACITY, 1);
            return;
        }
        M()->assigned[M()->n_assigned++] = (short)dd;
    }
    LDO_HDC void push_modified(int dd) {
        if (M()->n_modified >= K::LV) {
            sys.fail(LDO_ERR_CAPACITY, 2);
            return;
        }
        M()->modified[M()->n_modified++] = (short)dd;
    }

    // MCMovetype::reset_origami (movetypes.cpp:53-85)
    LDO_HDN void reset_origami() {
#pragma unroll 1
        for (int k = 0; k < M()->n_assigned; k++) sys.unassign_domain(M()->assigned[k]);
        if (M()->added_chain >= 0) {
            sys.delete_chain(M()->added_chain);
            sys.S()->current_c_i -= 1;
        }
#pragma unroll 1
        for (int k = 0; k < M()->n_modified; k++) {
            int dd = M()->modified[k];
            const DomRec& r = C()->prev[dd];
            sys.set_checked_domain_config(dd, rec_pos(r), r.ore);
        }
        sys.S()->constraints_violated = 0;
    }

    // staple_is_connector / scan_for_scaffold_domain (movetypes.cpp:160-232); chains tracked by slot
    LDO_HDN bool scan_for_scaffold_domain(int start, uint8_t* participating) {
        // iterative depth-first walk; frame = (entry domain, next index in its chain)
        short (*st)[2] = C()->scan_stack;
        int sp = 0;
        st[0][0] = (short)start;
        st[0][1] = 0;
        participating[sys.chain(start)] = 1;
#pragma unroll 1
        while (sp >= 0) {
            int dom = st[sp][0];
            int c = sys.chain(dom);
            int base = sys.chain_base(c);
            int len = sys.S()->chain_len[c];
            bool descended = false;
#pragma unroll 1
            while (st[sp][1] < len) {
                int cur = base + st[sp][1];
                st[sp][1]++;
                if (cur == dom) continue;
                int b = sys.S()->bound[cur];
                if (b < 0) continue;
                if (sys.chain(b) == c) continue;
                if (sys.chain(b) == 0) return true;
                if (participating[sys.chain(b)]) continue;
                if (sp + 1 >= K::C) {
                    sys.fail(LDO_ERR_CAPACITY, 3);
                    return true;
                }
                participating[sys.chain(b)] = 1;
                sp++;
                st[sp][0] = (short)b;
                st[sp][1] = 0;
                descended = true;
                break;
            }
            if (!descended) sp--;
        }
        return false;
    }
    LDO_HDN bool staple_is_connector(int c) {
        int base = sys.chain_base(c);
#pragma unroll 1
        for (int k = 0; k < sys.S()->chain_len[c]; k++) {
            int dd = base + k;
            if (sys.S()->dom[dd].state != ST_UNBOUND) {
                int b = sys.S()->bound[dd];
                if (sys.chain(b) == 0) continue;
#pragma unroll 1
                for (int q = 0; q < K::C; q++) C()->net_chain[q] = 0;
                C()->net_chain[c] = 1;
                if (!scan_for_scaffold_domain(b, C()->net_chain)) return true;
            }
        }
        return false;
    }
    LDO_HD int num_bound_staple_domains(int c) const {
        int n = 0;
        int base = sys.chain_base(c);
#pragma unroll 1
        for (int k = 0; k < sys.S()->chain_len[c]; k++) {
            int st = sys.S()->dom[base + k].state;
            if (st == ST_BOUND || st == ST_MISBOUND) n++;
        }
        return n;
    }
    LDO_HD bool staple_has_bound_domain(int c) const {
        int base = sys.chain_base(c);
#pragma unroll 1
        for (int k = 0; k < sys.S()->chain_len[c]; k++) {
            if (sys.S()->dom[base + k].state == ST_BOUND) return true;
        }
        return false;
    }
    // find_bound_domains (movetypes.cpp:258-277): count, and k-th pair (new domain, old domain)
    LDO_HD int count_bound_to_other_chains(int c) const {
        int n = 0;
        int base = sys.chain_base(c);
#pragma unroll 1
        for (int k = 0; k < sys.S()->chain_len[c]; k++) {
            int b = sys.S()->bound[base + k];
            if (b >= 0 && sys.chain(b) != c) n++;
        }
        return n;
    }
    LDO_HD int kth_bound_to_other_chains(int c, int kth) const {
        int base = sys.chain_base(c);
#pragma unroll 1
        for (int k = 0; k < sys.S()->chain_len[c]; k++) {
            int b = sys.S()->bound[base + k];
            if (b >= 0 && sys.chain(b) != c) {
                if (kth == 0) return base + k;
                kth--;
            }
        }
        return -1;
    }

    // ---- OrientationRotation (orientation_movetype.cpp:30-65) ----
    LDO_HDN bool move_orientation_rotation() {
        bool accepted = false;
        int dd = select_random_domain();
        int o_new = uniform_int(0, 5);
        int st = sys.S()->dom[dd].state;
        if (st == ST_BOUND || st == ST_MISBOUND) {
            double de = 0;
            int o_old = sys.S()->dom[dd].ore;
            int b = sys.S()->bound[dd];
            de += sys.unassign_domain(b);
            // set_domain_orientation: domain is now unbound (origami_system.cpp:543-551)
            sys.S()->dom[dd].ore = (int8_t)o_new;
            V3 p = rec_pos(sys.S()->dom[dd]);
            de += sys.set_domain_config(b, p, o_new ^ 1);
            if (!sys.S()->constraints_violated) accepted = test_acceptance(exp(-de));
            if (!accepted) {
                sys.unassign_domain(b);
                sys.S()->dom[dd].ore = (int8_t)o_old;
                sys.set_checked_domain_config(b, p, o_old < 6 ? (o_old ^ 1) : o_old);
            }
        }
        else {
            sys.S()->dom[dd].ore = (int8_t)o_new;
            accepted = true;
        }
        return accepted;
    }

    // ---- growth helpers (movetypes.cpp:322-383, met_movetypes.cpp:50-95) ----
    LDO_HD double set_growth_point(int d_new, int d_old) {
        const DomRec& ro = sys.S()->dom[d_old];
        int o_new = ro.ore < 6 ? (ro.ore ^ 1) : ro.ore;
        double de = sys.set_domain_config(d_new, rec_pos(ro), o_new);
        if (sys.S()->constraints_violated) M()->rejected = 1;
        else push_assigned(d_new);
        return de;
    }
    // MetMCMovetype::grow_chain over domains base+from, base+from+step, ... (count domains incl. the first)
    LDO_HDN void met_grow_chain(int first, int stepdir, int count) {
#pragma unroll 1
        for (int i = 1; i < count; i++) {
            int dd = first + stepdir * i;
            int prev = first + stepdir * (i - 1);
            V3 p = rec_pos(sys.S()->dom[prev]) + ore_vec(uniform_int(0, 5));
            int o = uniform_int(0, 5);
            W()->delta_e += sys.set_domain_config(dd, p, o);
            if (sys.S()->constraints_violated) {
                M()->rejected = 1;
                break;
            }
            push_assigned(dd);
        }
    }
    // RegrowthMCMovetype::grow_staple (movetypes.cpp:343-370)
    LDO_HD void met_grow_staple(int c, int d_i) {
        int base = sys.chain_base(c);
        int len = sys.S()->chain_len[c];
        if (len - d_i > 1) met_grow_chain(base + d_i, +1, len - d_i);
        if (M()->rejected) return;
        if (d_i + 1 > 1) met_grow_chain(base + d_i, -1, d_i + 1);
    }
    LDO_HD void met_unassign_domains(int c) {
        int base = sys.chain_base(c);
#pragma unroll 1
        for (int k = 0; k < sys.S()->chain_len[c]; k++) {
            int dd = base + k;
            C()->prev[dd] = sys.S()->dom[dd];
            push_modified(dd);
            W()->delta_e += sys.unassign_domain(dd);
        }
    }
    LDO_HD void met_add_external_bias() {
        update_move_params();
        W()->delta_e += calc_move_bias();
    }

    // ---- MetStapleExchange (met_movetypes.cpp:192-403) ----
    LDO_HD bool exchange_accept(double pratio, int type, bool staple_bound, const MoveDef& md) {
        if (staple_bound) {
            int mi = md.exchange_mults_off + type - 1;
            M()->modifier *= md.adaptive_exchange ? STATS()->exchange_mults[mi] : MS().exchange_mults[mi];
            if (M()->modifier * fmin(1.0, pratio) > 1) {
                if (md.adaptive_exchange) {
                    // a multiplier that produces a probability above one is cut tenfold and the move rejected
                    if (LDO_LANE == 0) STATS()->exchange_mults[mi] /= 10;
                    LDO_SYNCWARP();
                    return false;
                }
                if (MS().allow_nonsensical_ps) return true;
                sys.fail(LDO_ERR_NONSENSICAL_P, type);
                return false;
            }
        }
        return test_acceptance(pratio);
    }
    LDO_HDN bool move_staple_exchange(const MoveDef& md) {
        SysState<K>* s = sys.S();
        W()->delta_e = 0;
        int insertion_sites = s->num_domains;
        if (uniform_real() < 0.5) {
            // insert_staple (:303-359)
            int type = uniform_int(1, sys.SC().n_types - 1);
            W()->trk_a = type;
            W()->trk_b = 0;
            if (s->num_staples == sys.SC().max_total_staples) return false;
            if (s->type_count[type] == sys.SC().max_type_staples) return false;
            int c = sys.add_chain(type);
            if (c < 0) return false;
            if (sys.SC().apply_mean_field_cor) W()->delta_e += log(6.0);
            W()->delta_e += sys.TT().init_energy;
            M()->added_chain = c;
            // select_new_growthpoint (movetypes.cpp:372-383)
            int len = s->chain_len[c];
            int g_new = sys.chain_base(c) + uniform_int(0, len - 1);
            int g_old = select_random_domain();
#pragma unroll 1
            while (sys.chain(g_old) == c && s->status == LDO_OK) g_old = select_random_domain();
            W()->delta_e += set_growth_point(g_new, g_old);
            if (M()->rejected) return false;
            met_grow_staple(c, s->dindex[g_new]);
            if (M()->rejected) return false;
            bool staple_bound = staple_has_bound_domain(c);
            int num_bd = num_bound_staple_domains(c);
            // staple_insertion_accepted (:212-253)
            met_add_external_bias();
            double boltz = exp(-W()->delta_e);
            int Ni_new = s->type_count[type];
            double pratio = (double)len / 6.0 / Ni_new * boltz;
            pratio *= insertion_sites * sys.SC().staple_M;
            pratio /= num_bd;
            return exchange_accept(pratio, type, staple_bound, md);
        }
        // delete_staple (:361-403)
        int type = uniform_int(1, sys.SC().n_types - 1);
        W()->trk_a = type;
        W()->trk_b = 1;
        int n_of_type = s->type_count[type];
        if (n_of_type == 0) {
            M()->rejected = 1;
            return false;
        }
        int c = sys.staple_of_type(type, uniform_int(0, n_of_type - 1));
        if (staple_is_connector(c)) return false;
        bool staple_bound = staple_has_bound_domain(c);
        int num_bd = num_bound_staple_domains(c);
        int len = s->chain_len[c];
        met_unassign_domains(c);
        if (sys.SC().apply_mean_field_cor) W()->delta_e -= log(6.0);
        W()->delta_e -= sys.TT().init_energy;
        // staple_deletion_accepted (:255-301)
        s->num_staples--;
        met_add_external_bias();
        s->num_staples++;
        double boltz = exp(-W()->delta_e);
        int Ni = s->type_count[type];
        double pratio = Ni * 6.0 / (double)len * boltz;
        pratio /= sys.SC().staple_M * (insertion_sites - len);
        pratio *= num_bd;
        bool accepted = exchange_accept(pratio, type, staple_bound, md);
        if (accepted) sys.delete_chain(c);
        return accepted;
    }

    // ---- MetStapleRegrowth (met_movetypes.cpp:470-513) ----
    LDO_HDN bool move_met_staple_regrowth() {
        SysState<K>* s = sys.S();
        W()->delta_e = 0;
        if (s->num_staples == 0) {
            W()->trk_b = 1; // m_tracker.no_staples = true, never reset by this movetype (met_movetypes.cpp:474-477)
            return false;
        }
        int c = s->order[uniform_int(1, s->num_staples)];
        W()->trk_a = s->chain_type[c];
        if (staple_is_connector(c)) return false;
        int n_bd = count_bound_to_other_chains(c);
        if (n_bd == 0) {
            sys.fail(LDO_ERR_UNBOUND_STAPLE, c);
            return false;
        }
        int g_new = kth_bound_to_other_chains(c, uniform_int(0, n_bd - 1));
        int g_old = s->bound[g_new];
        met_unassign_domains(c);
        W()->delta_e += set_growth_point(g_new, g_old);
        if (M()->rejected) return false;
        met_grow_staple(c, s->dindex[g_new]);
        if (M()->rejected) return false;
        met_add_external_bias();
        int new_num_bd = num_bound_staple_domains(c);
        double pratio = exp(-W()->delta_e) * n_bd / new_num_bd;
        return test_acceptance(pratio);
    }

    // ---- CB staple regrowth (cb_movetypes.cpp:44-404) ----
    // Weights of the six neighbour sites of p_prev for `dom`, evaluated one site per lane
    // (CBMCMovetype::calc_biases, cb_movetypes.cpp:58-102). Results: site_kind 0 = skipped,
    // 1 = empty site (orientation drawn later), 2 = binds the unbound occupant with orientation site_o.
    LDO_HDN void cb_site_weights(V3 p_prev, int dom) {
        // lattice lookups and misbinding weights, one site per lane; a site holding the unbound complement of
        // `dom` is left pending (kind 3)
#pragma unroll 1
        for (int k = LDO_LANE; k < 6; k += LDO_NLANES) {
            V3 p = p_prev + ore_vec(k);
            int j = sys.occupant(p);
            int kind = 0, o = ORE_ZERO;
            double w = 0;
            if (j < 0) {
                kind = 1;
                w = 6 * exp(-0.0);
            }
            else if (sys.S()->dom[j].state == ST_UNBOUND) {
                int oj = sys.S()->dom[j].ore;
                o = oj < 6 ? (oj ^ 1) : oj;
                if (sys.ident(dom) == -sys.ident(j)) {
                    kind = 3;
                }
                else {
                    DeltaConfig dc = sys.eval_misbind(dom, j, o);
                    if (!dc.violated) {
                        kind = 2;
                        w = exp(-dc.e);
                    }
                }
            }
            M()->site_kind[k] = (int8_t)kind;
            M()->site_o[k] = (int8_t)o;
            M()->site_w[k] = w;
        }
        LDO_SYNCWARP();
        // binding evaluations of the pending sites, by the whole warp (System::eval_place is warp-uniform)
#pragma unroll 1
        for (int k = 0; k < 6; k++) {
            if (M()->site_kind[k] != 3) continue;
            int ns, partner;
            DeltaConfig dc = sys.eval_place(dom, p_prev + ore_vec(k), M()->site_o[k], &ns, &partner);
            LDO_SYNCWARP();
            M()->site_kind[k] = (int8_t)(dc.violated ? 0 : 2);
            M()->site_w[k] = dc.violated ? 0.0 : exp(-dc.e);
            LDO_SYNCWARP();
        }
    }
    // select_and_set_config for CBStapleRegrowth (cb_movetypes.cpp:104-160, 363-387)
    LDO_HDS void cb_select_and_set_config(int dom, int prev_dom, bool regrow_old, DD& bias) {
        V3 p_prev = rec_pos(sys.S()->dom[prev_dom]);
        cb_site_weights(p_prev, dom);
        sys.S()->constraints_violated = 0;
        double ros = 0;
#pragma unroll 1
        for (int k = 0; k < 6; k++) {
            if (M()->site_kind[k] != 0) ros += M()->site_w[k];
        }
        if (ros == 0) {
            M()->rejected = 1;
            return;
        }
        bias = dd_mul(bias, ros);
        if (!regrow_old) {
            double cum = 0;
            double r = uniform_real();
            V3 p_new = v3(0, 0, 0);
            int o_new = ORE_ZERO;
#pragma unroll 1
            for (int k = 0; k < 6; k++) {
                if (M()->site_kind[k] == 0) continue;
                cum += M()->site_w[k] / ros;
                if (r < cum) {
                    p_new = p_prev + ore_vec(k);
                    o_new = M()->site_kind[k] == 2 ? M()->site_o[k] : ORE_ZERO;
                    break;
                }
            }
            if (o_new == ORE_ZERO) o_new = uniform_int(0, 5);
            sys.set_checked_domain_config(dom, p_new, o_new);
        }
        else {
            const DomRec& r = C()->oldc[dom];
            sys.set_checked_domain_config(dom, rec_pos(r), r.ore);
        }
        push_assigned(dom);
    }
    LDO_HD void cb_grow_chain(int first, int stepdir, int count, bool regrow_old, DD& bias) {
#pragma unroll 1
        for (int i = 1; i < count; i++) {
            cb_select_and_set_config(first + stepdir * i, first + stepdir * (i - 1), regrow_old, bias);
            if (M()->rejected) break;
        }
    }
    LDO_HD void cb_set_growthpoint_and_grow_staple(int g_new, int g_old, int c, bool regrow_old, DD& bias) {
        if (regrow_old) {
            // set_old_growth_point (cb_movetypes.cpp:168-180)
            double de = sys.set_checked_domain_config(g_new, rec_pos(sys.S()->dom[g_old]), C()->oldc[g_new].ore);
            bias = dd_mul(bias, exp(-de));
            push_assigned(g_new);
        }
        else {
            double de = set_growth_point(g_new, g_old);
            bias = dd_mul(bias, exp(-de));
        }
        if (!M()->rejected) {
            int base = sys.chain_base(c);
            int len = sys.S()->chain_len[c];
            int d_i = sys.S()->dindex[g_new];
            if (len - d_i > 1) cb_grow_chain(base + d_i, +1, len - d_i, regrow_old, bias);
            if (M()->rejected) return;
            if (d_i + 1 > 1) cb_grow_chain(base + d_i, -1, d_i + 1, regrow_old, bias);
        }
    }
    LDO_HD void cb_unassign_domains(int c) {
        int base = sys.chain_base(c);
#pragma unroll 1
        for (int k = 0; k < sys.S()->chain_len[c]; k++) {
            int dd = base + k;
            C()->prev[dd] = sys.S()->dom[dd];
            push_modified(dd);
            sys.unassign_domain(dd);
        }
    }
    LDO_HDN bool move_cb_staple_regrowth() {
        SysState<K>* s = sys.S();
        if (s->num_staples == 0) {
            W()->trk_b = 1;
            return false;
        }
        W()->trk_b = 0; // (cb_movetypes.cpp:296-303)
        int c = s->order[uniform_int(1, s->num_staples)];
        W()->trk_a = s->chain_type[c];
        if (staple_is_connector(c)) return false;
        DD bias = dd_from(1.0);
        int n_bd = count_bound_to_other_chains(c);
        if (n_bd == 0) {
            sys.fail(LDO_ERR_UNBOUND_STAPLE, c);
            return false;
        }
        // bound (new, old) pairs are fixed before anything is unassigned
        short bd_new[8], bd_old[8];
        bool small = n_bd <= 8;
        if (small) {
#pragma unroll 1
            for (int k = 0; k < n_bd; k++) {
                bd_new[k] = (short)kth_bound_to_other_chains(c, k);
                bd_old[k] = s->bound[bd_new[k]];
            }
        }
        else {
            sys.fail(LDO_ERR_CAPACITY, 4);
            return false;
        }
        bias = dd_mul(bias, (double)n_bd);
        int gi = uniform_int(0, n_bd - 1);
        cb_unassign_domains(c);
        cb_set_growthpoint_and_grow_staple(bd_new[gi], bd_old[gi], c, false, bias);
        if (M()->rejected) return false;
        bias = dd_div(bias, (double)num_bound_staple_domains(c));
        // add_external_bias (cb_movetypes.cpp:52-56)
        update_move_params();
        bias = dd_mul(bias, exp(-calc_move_bias()));
        // setup_for_regrow_old (cb_movetypes.cpp:237-247)
        DD new_bias = bias;
        double new_modifier = M()->modifier;
        bias = dd_from(1.0);
        M()->n_modified = 0;
        M()->n_assigned = 0;
        {
            int base = sys.chain_base(c);
#pragma unroll 1
            for (int k = 0; k < s->chain_len[c]; k++) C()->oldc[base + k] = C()->prev[base + k];
        }
        gi = uniform_int(0, n_bd - 1);
        cb_unassign_domains(c);
        cb_set_growthpoint_and_grow_staple(bd_new[gi], bd_old[gi], c, true, bias);
        M()->modifier = new_modifier;
        // test_cb_acceptance (cb_movetypes.cpp:182-198)
        double ratio = dd_ratio_min1(new_bias, bias);
        if (test_acceptance(ratio)) {
            reset_origami();
            update_move_params();
            calc_move_bias();
            return true;
        }
        M()->n_modified = 0;
        M()->n_assigned = 0;
        return false;
    }

    // ---- Constraintpoints (top_constraint_points.cpp:163-601) ----
    LDO_HDN void cp_reset() {
#pragma unroll 1
        for (int k = 0; k < K::D; k++) {
            M()->seg_of[k] = -1;
            M()->stem_gp[k] = -1;
            M()->gp_stem[k] = -1;
            M()->inactive[k] = -1;
            M()->stem_seg0[k] = -1;
            C()->in_sel[k] = 0;
        }
#pragma unroll 1
        for (int k = 0; k < K::C; k++) {
            C()->checked_chain[k] = 0;
            C()->regrow_chain[k] = 0;
        }
#pragma unroll 1
        for (int k = 0; k < MoveScratch<K>::S; k++) M()->scaf_dir[k] = 0;
        M()->n_ep = 0;
        C()->n_ep0 = 0;
        M()->n_erased = 0;
        M()->n_regrow = 0;
        C()->n_sel = 0;
    }
    LDO_HD int cp_dir_of(int chain, int seg) const {
        if (chain == 0) return M()->scaf_dir[seg];
        return seg == 0 ? 1 : -1;
    }
    LDO_HD int cp_get_dir(int dd) const {
        int seg = M()->seg_of[dd] < 0 ? 0 : M()->seg_of[dd];
        return cp_dir_of(sys.chain(dd), seg);
    }
    LDO_HDN void cp_add_active_endpoint_seg(int dd, V3 p, int seg) {
        if (M()->n_ep >= MoveScratch<K>::E) {
            sys.fail(LDO_ERR_CAPACITY, 5);
            return;
        }
        int e = M()->n_ep++;
        M()->ep_chain[e] = (short)sys.chain(dd);
        M()->ep_seg[e] = (int8_t)seg;
        M()->ep_d[e] = (short)sys.dindex(dd);
        M()->ep_pos[e] = p.k;
    }
    LDO_HD void cp_add_active_endpoint(int dd, V3 p) { cp_add_active_endpoint_seg(dd, p, M()->seg_of[dd]); }
    LDO_HD void cp_erase_ep(int e) {
#pragma unroll 1
        for (int k = e; k + 1 < M()->n_ep; k++) {
            M()->ep_chain[k] = M()->ep_chain[k + 1];
            M()->ep_seg[k] = M()->ep_seg[k + 1];
            M()->ep_d[k] = M()->ep_d[k + 1];
            M()->ep_pos[k] = M()->ep_pos[k + 1];
        }
        M()->n_ep--;
    }
    LDO_HDN void cp_save_initial() {
        C()->n_ep0 = M()->n_ep;
#pragma unroll 1
        for (int k = 0; k < M()->n_ep; k++) {
            C()->ep0_chain[k] = M()->ep_chain[k];
            C()->ep0_seg[k] = M()->ep_seg[k];
            C()->ep0_d[k] = M()->ep_d[k];
            C()->ep0_pos[k] = M()->ep_pos[k];
        }
    }
    LDO_HDN void cp_reset_active_endpoints() {
        M()->n_ep = C()->n_ep0;
#pragma unroll 1
        for (int k = 0; k < C()->n_ep0; k++) {
            M()->ep_chain[k] = C()->ep0_chain[k];
            M()->ep_seg[k] = C()->ep0_seg[k];
            M()->ep_d[k] = C()->ep0_d[k];
            M()->ep_pos[k] = C()->ep0_pos[k];
        }
    }
    // remove_active_endpoint (:299-320): erased positions are kept in m_erased_endpoints
    LDO_HDN void cp_remove_active_endpoint(int dd) {
        M()->n_erased = 0;
        int c = sys.chain(dd), seg = M()->seg_of[dd], di_ = sys.dindex(dd);
        int k = 0;
#pragma unroll 1
        while (k < M()->n_ep) {
            if (M()->ep_chain[k] == c && M()->ep_seg[k] == seg && M()->ep_d[k] == di_) {
                if (M()->n_erased < 8) {
                    M()->erased_pos[M()->n_erased] = M()->ep_pos[k];
                    M()->n_erased++;
                }
                else {
                    sys.fail(LDO_ERR_CAPACITY, 6);
                }
                cp_erase_ep(k);
            }
            else {
                k++;
            }
        }
    }
    // remove_activated_endpoint (:322-338)
    LDO_HDS void cp_remove_activated_endpoint(int dd) {
        int ed = M()->inactive[dd];
        if (ed < 0) return;
        int c = sys.chain(ed), seg = M()->seg_of[ed], di_ = sys.dindex(ed);
#pragma unroll 1
        for (int k = 0; k < M()->n_ep; k++) {
            if (M()->ep_chain[k] == c && M()->ep_seg[k] == seg && M()->ep_d[k] == di_) {
                cp_erase_ep(k);
                break;
            }
        }
    }
    // update_endpoints (:340-349)
    LDO_HDN void cp_update_endpoints(int dd) {
        cp_remove_active_endpoint(dd);
        int ed = M()->inactive[dd];
        if (ed >= 0) cp_add_active_endpoint(ed, rec_pos(sys.S()->dom[dd]));
    }
    // endpoint_reached (:359-372)
    LDO_HDN bool cp_endpoint_reached(int dd, V3 p, const EpOverlay* ov = nullptr) const {
        int c = sys.chain(dd), seg = M()->seg_of[dd], di_ = sys.dindex(dd);
        if (ov && ov->add_chain == c && ov->add_seg == seg && ov->add_d == di_ && ov->add_pos == p) return true;
        if (ov && ov->rm_chain == c && ov->rm_seg == seg && ov->rm_d == di_) return false;
#pragma unroll 1
        for (int k = 0; k < M()->n_ep; k++) {
            if (M()->ep_chain[k] == c && M()->ep_seg[k] == seg && M()->ep_d[k] == di_ && M()->ep_pos[k] == p.k) {
                return true;
            }
        }
        return false;
    }
    // calc_remaining_steps (:571-601)
    LDO_HD int cp_remaining_steps(int end_d_i, int dd, int dir_) const {
        int dm = sys.dindex(dd);
        if (sys.SC().cyclic && sys.chain(dd) == 0) {
            int n = sys.S()->chain_len[0];
            if (dir_ > 0 && end_d_i < dm) return n + end_d_i - dm;
            if (dir_ < 0 && end_d_i > dm) return dm + n - end_d_i;
            if (dir_ == 0 || end_d_i == dm) return 0;
            return abs(end_d_i - dm);
        }
        return abs(end_d_i - dm);
    }
    // num_walks(...) == 0 (ideal_random_walk.cpp:35-40)
    LDO_HD static bool no_walks(V3 a, V3 b, int steps) {
        int dr = abssum(b - a);
        return dr > steps || (steps - dr) % 2 != 0;
    }
    // walks_remain (:397-445)
#ifdef LDO_X_WALKS_MERGE
    LDO_HDS
#else
    LDO_HDN
#endif
    bool cp_walks_remain_seg(int c, int seg, int dd, V3 p, const EpOverlay* ov) const {
        LDO_COUNT(6);
        int dir_ = cp_dir_of(c, seg);
        bool rm = ov && ov->rm_chain == c && ov->rm_seg == seg;
#pragma unroll 1
        for (int k = 0; k < M()->n_ep; k++) {
            if (M()->ep_chain[k] != c || M()->ep_seg[k] != seg) continue;
            if (rm && M()->ep_d[k] == ov->rm_d) continue;
            int steps = cp_remaining_steps(M()->ep_d[k], dd, dir_);
            V3 ep;
            ep.k = M()->ep_pos[k];
            if (no_walks(p, ep, steps)) return false;
        }
        if (ov && ov->add_chain == c && ov->add_seg == seg) {
            int steps = cp_remaining_steps(ov->add_d, dd, dir_);
            if (no_walks(p, ov->add_pos, steps)) return false;
        }
        return true;
    }
    LDO_HDN bool cp_walks_remain(int dd, V3 p, const EpOverlay* ov = nullptr) const {
        int c = sys.chain(dd);
#ifdef LDO_X_WALKS_MERGE // experiment twin (profiles/ab_r2.txt): one call site of the segment test, merged in
        int seg = M()->seg_of[dd], n_segs = 1;
        if (M()->stem_gp[dd] >= 0) {
            seg = M()->stem_seg0[dd];
            if (seg < 0) return true;
            n_segs = 2;
        }
#pragma unroll 1
        for (int k = 0; k < n_segs; k++) {
            if (!cp_walks_remain_seg(c, seg + k, dd, p, ov)) return false;
        }
        return true;
#else
        if (M()->stem_gp[dd] >= 0) {
            int s0 = M()->stem_seg0[dd];
            if (s0 < 0) return true; // m_stemd_to_segs[domain] default-constructs to an empty list
            if (!cp_walks_remain_seg(c, s0, dd, p, ov)) return false;
            return cp_walks_remain_seg(c, s0 + 1, dd, p, ov);
        }
        return cp_walks_remain_seg(c, M()->seg_of[dd], dd, p, ov);
#endif
    }

    // StapleNetwork::scan_network (top_constraint_points.cpp:36-161), iterative.
    // Returns whether the network is externally bound.
    LDO_HDN bool net_scan(int start) {
        SysState<K>* s = sys.S();
#pragma unroll 1
        for (int k = 0; k < K::C; k++) C()->net_chain[k] = 0;
        C()->net_chain[0] = 1;
        C()->n_pot_gps = 0;
        C()->n_pot_iaes = 0;
        C()->n_pot_ds = 0;
        bool external = false;
        short (*st)[2] = C()->scan_stack;
        int sp = 0;
        // enter frame
        st[0][0] = (short)start;
        st[0][1] = 0;
        C()->net_chain[sys.chain(start)] = 1;
        C()->net_growth_idx[sys.chain(start)] = (short)sys.dindex(start);
        C()->pot_ds[C()->n_pot_ds++] = (short)start;
#pragma unroll 1
        while (sp >= 0) {
            int g = st[sp][0];
            int ci = sys.chain(g);
            int base = sys.chain_base(ci);
            int len = s->chain_len[ci];
            int gi = sys.dindex(g);
            bool descended = false;
#pragma unroll 1
            while (st[sp][1] < len - 1) {
                int k = st[sp][1]++;
                // make_staple_stack order: 3' of the growth domain, then 5' (:135-161)
                int idx = (k < len - 1 - gi) ? (gi + 1 + k) : (gi - 1 - (k - (len - 1 - gi)));
                int dd = base + idx;
                C()->pot_ds[C()->n_pot_ds++] = (short)dd;
                int bd = s->bound[dd];
                if (bd < 0 || sys.chain(bd) == ci) continue;
                int bd_ci = sys.chain(bd);
                if (C()->net_chain[bd_ci]) {
                    bool ext = false;
                    if (bd_ci == 0) ext = !C()->in_sel[bd];
                    if (!ext) {
                        // add_potential_inactive_endpoint (:155-161)
                        bool bd_in_ds = false;
#pragma unroll 1
                        for (int q = 0; q < C()->n_pot_ds; q++) {
                            if (C()->pot_ds[q] == bd) {
                                bd_in_ds = true;
                                break;
                            }
                        }
                        int e = C()->n_pot_iaes++;
                        if (bd_in_ds) {
                            C()->pot_iaes[e][0] = (short)bd;
                            C()->pot_iaes[e][1] = (short)dd;
                        }
                        else {
                            C()->pot_iaes[e][0] = (short)dd;
                            C()->pot_iaes[e][1] = (short)bd;
                        }
                    }
                    if (!external && ext) external = true;
                }
                else {
                    int e = C()->n_pot_gps++;
                    C()->pot_gps[e][0] = (short)dd;
                    C()->pot_gps[e][1] = (short)bd;
                    if (sp + 1 >= K::C) {
                        sys.fail(LDO_ERR_CAPACITY, 7);
                        return true;
                    }
                    sp++;
                    st[sp][0] = (short)bd;
                    st[sp][1] = 0;
                    C()->net_chain[bd_ci] = 1;
                    C()->net_growth_idx[bd_ci] = (short)sys.dindex(bd);
                    C()->pot_ds[C()->n_pot_ds++] = (short)bd;
                    descended = true;
                    break;
                }
            }
            if (!descended) sp--;
        }
        return external;
    }

    // find_growthpoints_endpoints (:454-494)
    LDO_HDN void cp_find_growthpoints_endpoints(const short* doms, int n, int seg) {
        SysState<K>* s = sys.S();
#pragma unroll 1
        for (int q = 0; q < n; q++) {
            int dd = doms[q];
            if (M()->n_regrow >= K::LV) {
                sys.fail(LDO_ERR_CAPACITY, 8);
                return;
            }
            M()->regrow[M()->n_regrow++] = (short)dd;
            M()->seg_of[dd] = (int8_t)seg;
            int bd = s->bound[dd];
            if (bd < 0 || sys.chain(bd) == sys.chain(dd) || C()->checked_chain[sys.chain(bd)]) continue;
            bool external = net_scan(bd);
            int e = C()->n_pot_gps++;
            C()->pot_gps[e][0] = (short)dd;
            C()->pot_gps[e][1] = (short)bd;
            if (external) {
                // add_active_endpoints_on_scaffold (:540-559)
#pragma unroll 1
                for (int k = 0; k < C()->n_pot_gps; k++) {
                    int gd = C()->pot_gps[k][0];
                    if (sys.chain(gd) == 0) cp_add_active_endpoint_seg(gd, rec_pos(s->dom[gd]), seg);
                }
#pragma unroll 1
                for (int k = 0; k < C()->n_pot_iaes; k++) {
                    int second = C()->pot_iaes[k][1];
                    if (sys.chain(second) == 0) {
                        cp_add_active_endpoint_seg(C()->pot_iaes[k][0], rec_pos(s->dom[second]), seg);
                    }
                }
            }
            else {
#pragma unroll 1
                for (int k = 0; k < C()->n_pot_gps; k++) {
                    M()->gp_stem[C()->pot_gps[k][0]] = C()->pot_gps[k][1];
                    M()->stem_gp[C()->pot_gps[k][1]] = C()->pot_gps[k][0];
                }
#pragma unroll 1
                for (int k = 0; k < C()->n_pot_iaes; k++) M()->inactive[C()->pot_iaes[k][0]] = C()->pot_iaes[k][1];
                // add_regrowth_staples (:517-530)
#pragma unroll 1
                for (int k = 1; k < K::C; k++) {
                    if (C()->net_chain[k]) C()->regrow_chain[k] = 1;
                }
#pragma unroll 1
                for (int k = 0; k < C()->n_pot_ds; k++) {
                    int pd = C()->pot_ds[k];
                    if (M()->n_regrow >= K::LV) {
                        sys.fail(LDO_ERR_CAPACITY, 8);
                        return;
                    }
                    M()->regrow[M()->n_regrow++] = (short)pd;
                    // add_staple_to_segs_maps: unordered_map::insert keeps existing entries (:561-564)
                    if (M()->seg_of[pd] < 0) {
                        M()->seg_of[pd] = (sys.dindex(pd) >= C()->net_growth_idx[sys.chain(pd)]) ? 0 : 1;
                    }
                }
            }
#pragma unroll 1
            for (int k = 0; k < K::C; k++) {
                if (C()->net_chain[k]) C()->checked_chain[k] = 1;
            }
        }
    }

    // ---- CT selection (movetypes.cpp:469-719) ----
    // select_indices on the whole scaffold, seg 0 (min_length 2 for the scaffold regrowth moves, 1 for
    // the central segment of the linker moves)
    LDO_HDN void ct_select_indices(const MoveDef& md, int min_length = 2) {
        SysState<K>* s = sys.S();
        int n = s->chain_len[0];
#pragma unroll 1
        for (;;) {
            if (s->status != LDO_OK) return;
            int max_length = n < md.max_regrowth ? n : md.max_regrowth;
            int sel_length = uniform_int(min_length, max_length);
            int start_i = uniform_int(0, n - 1);
            W()->dir = uniform_int(0, 1);
            if (W()->dir == 0) W()->dir = -1;
            // forward part
            short* buf = C()->seg_dom; // scratch: forward list then backward list
            int nf = 0, nb = 0;
            int cur = sys.chain_base(0) + start_i;
#pragma unroll 1
            while (cur >= 0 && nf != sel_length) {
                buf[nf++] = (short)cur;
                cur = sys.step(cur, W()->dir);
            }
            int back = sys.step(sys.chain_base(0) + start_i, -W()->dir);
#pragma unroll 1
            while (back >= 0 && nf + nb != sel_length) {
                buf[K::D + 1 + nb] = (short)back;
                nb++;
                back = sys.step(back, -W()->dir);
            }
            if (nf + nb < min_length) continue;
            C()->n_sel = 0;
#pragma unroll 1
            for (int k = nb - 1; k >= 0; k--) C()->sel_scaf[C()->n_sel++] = buf[K::D + 1 + k];
#pragma unroll 1
            for (int k = 0; k < nf; k++) C()->sel_scaf[C()->n_sel++] = buf[k];
            if (cur >= 0) cp_add_active_endpoint_seg(cur, rec_pos(s->dom[cur]), 0);
            return;
        }
    }

    // check_for_stemds (movetypes.cpp:702-719); stems are queued in C()->stem_queue[qh..qt)
    LDO_HDN void ct_check_for_stemds(int cur, int& qt) {
        SysState<K>* s = sys.S();
        if (s->dom[cur].state != ST_BOUND) return;
        int bd = s->bound[cur];
#pragma unroll 1
        for (int sd = -1; sd <= 1; sd += 2) {
            int nd = sys.step(bd, sd);
            if (nd >= 0 && s->dom[nd].state == ST_BOUND) {
                int bn = s->bound[nd];
                if (sys.chain(bn) == 0) {
                    if (qt >= 4 * K::D + 8) {
                        sys.fail(LDO_ERR_CAPACITY, 9);
                        return;
                    }
                    C()->stem_queue[qt++] = (short)bn;
                }
            }
        }
    }
    // fill_seg (movetypes.cpp:666-700). seg contents are appended at C()->seg_dom[seg_n...]; returns
    // whether max_length was reached. `seg_size` counts elements already in the segment.
    LDO_HDN bool ct_fill_seg(int start_d, int max_length, int seg_max, int dir_, int& n_domains, int& qt, int& seg_n, int seg_size) {
        int cur = start_d;
        ct_check_for_stemds(cur, qt);
        int next = start_d;
#pragma unroll 1
        while (seg_size != seg_max && next >= 0) {
            next = sys.step(cur, dir_);
            if (next < 0) break;
            int next_next = sys.step(next, dir_);
            if (next_next >= 0 && C()->in_sel[next_next]) break;
            C()->seg_dom[seg_n++] = (short)next;
            seg_size++;
            C()->in_sel[next] = 1;
            n_domains++;
            if (n_domains == max_length) return true;
            cur = next;
            ct_check_for_stemds(cur, qt);
        }
        return false;
    }

    // select_noncontig_segs (movetypes.cpp:512-648). Segments are laid out in C()->seg_dom with
    // C()->seg_start; dirs in M()->scaf_dir; stems in C()->stems. Returns the number of segments.
    LDO_HDN int ct_select_noncontig_segs(const MoveDef& md, int& n_stems) {
        SysState<K>* s = sys.S();
        const int SMAX = MoveScratch<K>::S;
        int n = s->chain_len[0];
        int max_length = uniform_int(2, md.max_regrowth);
        int seg_max = uniform_int(2, md.max_seg_regrowth + 1);
        int start_d = sys.chain_base(0) + uniform_int(0, n - 1);
        int dir_ = uniform_int(0, 1);
        if (dir_ == 0) dir_ = -1;
        if (sys.step(start_d, dir_) < 0) dir_ *= -1;
        int n_domains = 0, qh = 0, qt = 0, seg_n = 0, n_segs = 0;
        n_stems = 0;
        // paired_empty[k] marks stems whose pair of segments was pushed as {[stem], []} (:556-563)
        // first segment
        C()->seg_start[0] = 0;
        C()->seg_dom[seg_n++] = (short)start_d;
        C()->in_sel[start_d] = 1;
        n_domains++;
        M()->scaf_dir[0] = (int8_t)dir_;
        bool max_reached = ct_fill_seg(start_d, max_length, seg_max, dir_, n_domains, qt, seg_n, 1);
        n_segs = 1;
        C()->seg_start[1] = (short)seg_n;
        // paired segment bookkeeping: for stem k, its two segments are 1+2k and 2+2k; pair_first[k]
        // records where the *paired_segs* (without the stem prefix) begin/end for endpoint search.
#pragma unroll 1
        while (!max_reached && qh != qt) {
            if (s->status != LDO_OK) return n_segs;
            int stemd_ = C()->stem_queue[qh++];
            if (C()->in_sel[stemd_]) continue;
            bool adjacent = false;
#pragma unroll 1
            for (int td = -1; td <= 1; td += 2) {
                int nd = sys.step(stemd_, td);
                if (nd >= 0 && C()->in_sel[nd]) {
                    adjacent = true;
                    break;
                }
            }
            if (adjacent) continue;
            if (n_segs + 2 >= SMAX) {
                sys.fail(LDO_ERR_CAPACITY, 10);
                return n_segs;
            }
            C()->in_sel[stemd_] = 1;
            n_domains++;
            C()->stems[n_stems++] = (short)stemd_;
            if (n_domains == max_length) {
                max_reached = true;
                // segs += [[stem], []], dirs += [1, -1]
                C()->seg_dom[seg_n++] = (short)stemd_;
                C()->seg_start[n_segs + 1] = (short)seg_n;
                C()->seg_start[n_segs + 2] = (short)seg_n;
                M()->scaf_dir[n_segs] = 1;
                M()->scaf_dir[n_segs + 1] = -1;
                n_segs += 2;
                break;
            }
            int dir1 = uniform_int(0, 1);
            if (dir1 == 0) dir1 = -1;
            M()->scaf_dir[n_segs] = (int8_t)dir1;
            M()->scaf_dir[n_segs + 1] = (int8_t)(-dir1);
            bool cur_seg_nonempty = true; // cur_seg = [stem]
#pragma unroll 1
            for (int i = 0; i < 2; i++) {
                int dsel = i == 0 ? dir1 : -dir1;
                int smax = uniform_int(0, md.max_seg_regrowth);
                if (i == 0) max_length++; // sic (movetypes.cpp:592-594)
                // segment i: optional stem prefix, then the filled domains
                if (cur_seg_nonempty) C()->seg_dom[seg_n++] = (short)stemd_;
                max_reached = ct_fill_seg(stemd_, max_length, smax, dsel, n_domains, qt, seg_n, 0);
                C()->seg_start[n_segs + i + 1] = (short)seg_n;
                if (max_reached) {
                    if (i == 0) C()->seg_start[n_segs + 2] = (short)seg_n; // second segment stays empty
                    break;
                }
                cur_seg_nonempty = false;
            }
            n_segs += 2;
        }
        return n_segs;
    }

    // ---- CTRG (rg_movetypes.cpp) ----
    LDO_HDN void eq_push_erased() {
        // m_erased_endpoints_q.push_back(get_erased_endpoints())
        if (M()->eq_depth >= MoveScratch<K>::A || M()->eq_npos + M()->n_erased > MoveScratch<K>::E) {
            sys.fail(LDO_ERR_CAPACITY, 11);
            return;
        }
        M()->eq_start[M()->eq_depth++] = (short)M()->eq_npos;
#pragma unroll 1
        for (int k = 0; k < M()->n_erased; k++) {
            M()->eq_pos[M()->eq_npos] = M()->erased_pos[k];
            M()->eq_npos++;
        }
    }
    // restore_endpoints (rg:288-295)
    LDO_HDN void rg_restore_endpoints() {
        cp_remove_activated_endpoint(W()->d);
        if (M()->eq_depth <= 0) {
            sys.fail(LDO_ERR_INTERNAL, 1);
            return;
        }
        int start = M()->eq_start[--M()->eq_depth];
#pragma unroll 1
        for (int k = start; k < M()->eq_npos; k++) {
            cp_add_active_endpoint(W()->d, V3{M()->eq_pos[k]});
        }
        M()->eq_npos = start;
    }
    // unassign_and_save_domains() (rg:67-87)
    LDO_HDN double rg_unassign_and_save_domains() {
        double de = 0;
        C()->prev[M()->regrow[0]] = sys.S()->dom[M()->regrow[0]];
#pragma unroll 1
        for (int k = 1; k < M()->n_regrow; k++) {
            int dd = M()->regrow[k];
            C()->prev[dd] = sys.S()->dom[dd];
            push_modified(dd);
            de += sys.unassign_domain(dd);
        }
        M()->eq_depth = 0;
        M()->eq_npos = 0;
        return de;
    }
    LDO_HD void rg_unassign_domains() {
#pragma unroll 1
        for (int k = 1; k < M()->n_regrow; k++) sys.unassign_domain(M()->regrow[k]);
        M()->eq_depth = 0;
        M()->eq_npos = 0;
    }
    // set_config (rg:233-244)
    LDO_HDN double rg_set_config(int dd, V3 p, int o) {
        double de = sys.set_checked_domain_config(dd, p, o);
        push_assigned(dd);
        cp_update_endpoints(W()->d);
        eq_push_erased();
        return de;
    }
    LDO_HD static unsigned long long all_cis() { return (1ull << 36) - 1; }
    // prepare_for_growth (rg:246-262)
    LDO_HDN void rg_prepare_for_growth() {
        W()->di++;
        W()->d = M()->regrow[W()->di];
        W()->stemd = M()->stem_gp[W()->d] >= 0;
        W()->dir = cp_get_dir(W()->d);
        W()->c_attempts = 0;
        if (W()->stemd) {
            W()->d_max_c_attempts = 1;
            W()->avail = 0;
            W()->ref_d = M()->stem_gp[W()->d];
        }
        else {
            W()->d_max_c_attempts = W()->max_c_attempts;
            W()->avail = all_cis();
            W()->ref_d = sys.step(W()->d, -W()->dir);
            rg_acquire_slot();
        }
    }
    // Chooses and fills the trial-probability cache of the current domain: the feeler memo when this is
    // the first feeler level of calc_weights and the parent sits on an empty site (the feeler's results
    // do not depend on the parent's orientation then), else the level's own slot.
    LDO_HD void rg_acquire_slot() {
        bool use_memo = W()->di == W()->memo_level && W()->memo_key >= 0;
        if (use_memo) {
            // the memo is keyed by the parent's site only: unusable when the current domain could bind
            // to the (unbound) parent itself, because that depends on the parent's orientation
            V3 dq = rec_pos(sys.S()->dom[M()->regrow[W()->di - 1]]) - rec_pos(sys.S()->dom[W()->ref_d]);
            if (abssum(dq) == 1) use_memo = false;
        }
        if (use_memo) {
            W()->cur_slot = LDO_RG_OWN_SLOTS + W()->memo_key;
            if (!((W()->memo_mask >> W()->memo_key) & 1)) {
                rg_compute_slot(W()->cur_slot);
                W()->memo_mask |= 1 << W()->memo_key;
            }
        }
        else {
            W()->cur_slot = W()->di & (LDO_RG_OWN_SLOTS - 1);
            rg_compute_slot(W()->cur_slot);
        }
    }
    // Open probability data of domain `dom` at site r (calc_p_config_open, rg:315-343, for the six
    // orientations of the site at once), in two steps. rg_site_lookup is read-only and runs on one lane per
    // site: an empty site is finished (kind 1), so is a site holding an unbound non-complementary domain
    // (misbinding needs no geometry); a site holding the unbound complement is left pending (kind 3, with the
    // orientation that binds it). rg_site_bind finishes a pending site and is executed by the whole warp, because
    // System::eval_place enters the candidate pair into the domain records while the potential is evaluated.
    // What a binding with energy change dc to domain j contributes (rg:326-343)
    LDO_HDC void rg_site_finish(int dom, bool dom_is_stem, V3 r, int j, const DeltaConfig& dc, const EpOverlay* ov, int& kind, double& pv) {
        kind = 0;
        pv = 0;
        if (!dc.violated && cp_walks_remain(dom, r, ov)) {
            bool same_chain = sys.chain(j) == sys.chain(dom);
            if (same_chain || dom_is_stem || cp_endpoint_reached(dom, r, ov)) {
                kind = 2;
                pv = fmin(1.0, exp(-dc.e));
            }
        }
    }
    LDO_HDN void rg_site_lookup(int dom, bool dom_is_stem, V3 r, const EpOverlay* ov, int& kind, int& o, double& pv) {
        LDO_COUNT(7);
        kind = 0;
        o = ORE_ZERO;
        pv = 0;
        int j = sys.occupant(r);
        if (j < 0) {
            kind = 1;
            pv = cp_walks_remain(dom, r, ov) ? 1.0 : 0.0;
        }
        else if (sys.S()->dom[j].state == ST_UNBOUND && sys.S()->dom[j].ore < 6) {
            o = sys.S()->dom[j].ore ^ 1;
            if (sys.ident(dom) == -sys.ident(j)) {
                kind = 3; // the complement: geometry matters, left to rg_site_bind
            }
            else {
                DeltaConfig dc = sys.eval_misbind(dom, j, o);
                rg_site_finish(dom, dom_is_stem, r, j, dc, ov, kind, pv);
            }
        }
    }
    LDO_HDN void rg_site_bind(int dom, bool dom_is_stem, V3 r, int o, const EpOverlay* ov, int& kind, double& pv) {
        kind = 0;
        pv = 0;
        int ns, j;
        DeltaConfig dc = sys.eval_place(dom, r, o, &ns, &j);
        if (j >= 0) rg_site_finish(dom, dom_is_stem, r, j, dc, ov, kind, pv);
    }
    // Evaluates the six neighbour sites of the reference domain for the current domain: lookups one site per
    // lane, then the pending bindings in turn.
    LDO_HDN void rg_compute_slot(int slot) {
        LDO_COUNT(8);
        RgSlot& sl = M()->slots[slot];
        V3 refp = rec_pos(sys.S()->dom[W()->ref_d]);
#pragma unroll 1
        for (int k = LDO_LANE; k < 6; k += LDO_NLANES) {
            int kind, o;
            double pv;
            rg_site_lookup(W()->d, W()->stemd != 0, refp + ore_vec(k), nullptr, kind, o, pv);
            sl.kind[k] = (uint8_t)kind;
            sl.ore[k] = (int8_t)o;
            sl.p[k] = pv;
        }
        LDO_SYNCWARP();
#pragma unroll 1
        for (int k = 0; k < 6; k++) {
            if (sl.kind[k] != 3) continue;
            int kind;
            double pv;
            rg_site_bind(W()->d, W()->stemd != 0, refp + ore_vec(k), sl.ore[k], nullptr, kind, pv);
            LDO_SYNCWARP();
            sl.kind[k] = (uint8_t)kind;
            sl.p[k] = pv;
            LDO_SYNCWARP();
        }
    }
    // Feeler slots of the NEXT domain for every empty parent site of the current domain, all 36
    // (parent site, feeler site) pairs spread over the lanes and evaluated read-only: the parent is not
    // placed; its only effects on the feeler (it is unbound on an empty site) are the endpoint updates
    // carried by an EpOverlay. Fills the memo slots and memo_mask. `fd` / `fref`: feeler domain and its
    // reference domain (the parent itself, or an already placed domain).
    LDO_HDS void rg_fill_feeler_memo(int fd, int fref) {
        LDO_COUNT(9);
        const RgSlot& own = M()->slots[W()->cur_slot];
        V3 refp = rec_pos(sys.S()->dom[W()->ref_d]);
        bool fref_is_parent = fref == W()->d;
        V3 frefp = fref_is_parent ? v3(0, 0, 0) : rec_pos(sys.S()->dom[fref]);
        int mask = 0;
#pragma unroll 1
        for (int pc = 0; pc < 6; pc++) {
            if (own.kind[pc] != 1) continue;
            // a feeler that can reach the parent's own site would bind to the parent: orientation dependent
            if (!fref_is_parent && abssum(refp + ore_vec(pc) - frefp) == 1) continue;
            mask |= 1 << pc;
        }
        EpOverlay ov;
        ov.rm_chain = sys.chain(W()->d);
        ov.rm_seg = M()->seg_of[W()->d];
        ov.rm_d = sys.dindex(W()->d);
        int ie = M()->inactive[W()->d];
        ov.add_chain = -1;
        ov.add_seg = 0;
        ov.add_d = 0;
        if (ie >= 0) {
            ov.add_chain = sys.chain(ie);
            ov.add_seg = M()->seg_of[ie];
            ov.add_d = sys.dindex(ie);
        }
#pragma unroll 1
        for (int t = LDO_LANE; t < 36; t += LDO_NLANES) {
            int pc = t / 6, k = t - 6 * pc;
            if (!((mask >> pc) & 1)) continue;
            V3 q = refp + ore_vec(pc);
            ov.add_pos = q;
            V3 r = (fref_is_parent ? q : frefp) + ore_vec(k);
            int kind, o;
            double pv;
            rg_site_lookup(fd, false, r, &ov, kind, o, pv);
            RgSlot& sl = M()->slots[LDO_RG_OWN_SLOTS + pc];
            sl.kind[k] = (uint8_t)kind;
            sl.ore[k] = (int8_t)o;
            sl.p[k] = pv;
        }
        LDO_SYNCWARP();
        // pending bindings of the feeler, by the whole warp
#pragma unroll 1
        for (int pc = 0; pc < 6; pc++) {
            if (!((mask >> pc) & 1)) continue;
            RgSlot& sl = M()->slots[LDO_RG_OWN_SLOTS + pc];
            V3 q = refp + ore_vec(pc);
#pragma unroll 1
            for (int k = 0; k < 6; k++) {
                if (sl.kind[k] != 3) continue;
                ov.add_pos = q;
                V3 r = (fref_is_parent ? q : frefp) + ore_vec(k);
                int kind;
                double pv;
                rg_site_bind(fd, false, r, sl.ore[k], &ov, kind, pv);
                LDO_SYNCWARP();
                sl.kind[k] = (uint8_t)kind;
                sl.p[k] = pv;
                LDO_SYNCWARP();
            }
        }
        W()->memo_mask = mask;
    }
    // prepare_for_regrowth (rg:264-286)
    LDO_HDN double rg_prepare_for_regrowth() {
        W()->di--;
        W()->d = M()->regrow[W()->di];
        W()->dir = cp_get_dir(W()->d);
        double de;
#ifndef LDO_NO_SLOT_CACHE
        if (W()->in_growth) {
            // the placement is taken back in the environment it was made in: minus its recorded changes
            sys.S()->weight_pass = 1;
            sys.unassign_domain(W()->d);
            sys.S()->weight_pass = 0;
            de = -C()->set_de[W()->di];
            sys.S()->energy += de;
            sys.S()->num_stacked_pairs -= C()->set_sp[W()->di];
        }
        else
#endif
        {
            de = sys.unassign_domain(W()->d);
        }
        if (M()->n_assigned > 0) M()->n_assigned--;
        rg_restore_endpoints();
        W()->stemd = M()->stem_gp[W()->d] >= 0;
        if (W()->stemd) {
            W()->d_max_c_attempts = 1;
            W()->c_attempts = 1;
            W()->avail = 0;
            W()->ref_d = M()->stem_gp[W()->d];
        }
        else {
            W()->d_max_c_attempts = W()->max_c_attempts;
            W()->c_attempts = M()->c_attempts_q[W()->di];
            W()->avail = M()->avail_q[W()->di];
            W()->ref_d = sys.step(W()->d, -W()->dir);
            W()->cur_slot = W()->di & (LDO_RG_OWN_SLOTS - 1);
            // a recoil returns to a level whose lower levels have not changed since its slot was computed
            if (!rg_load_cached_slot(W()->in_growth != 0)) {
                LDO_COUNT(14);
                rg_compute_slot(W()->cur_slot);
            }
        }
        return de;
    }
    // select_trial_config + calc_p_config_open (rg:297-343): draws the k-th remaining entry of the
    // ordered configuration list and returns its open probability from the current slot
    LDO_HDN double rg_trial(V3& p, int& o) {
        if (W()->stemd) {
            const DomRec& r = sys.S()->dom[W()->ref_d];
            p = rec_pos(r);
            o = r.ore < 6 ? (r.ore ^ 1) : r.ore;
            W()->last_pc = -1;
            W()->last_kind = 0;
            return rg_calc_p_config_open(p, o);
        }
        int n_avail = popc36(W()->avail);
        int ci = uniform_int(0, n_avail - 1);
        int i = nth_set_bit36(W()->avail, ci);
        W()->avail &= ~(1ull << i);
        // m_all_configs = all_pairs(vectors): position-major, orientation-minor (utility.hpp:167-177)
        int pc = i / 6;
        o = i - 6 * pc;
        p = ore_vec(pc) + rec_pos(sys.S()->dom[W()->ref_d]);
        const RgSlot& sl = M()->slots[W()->cur_slot];
        int kind = sl.kind[pc];
        W()->last_pc = pc;
        W()->last_kind = kind;
        double pv = 0;
        if (kind == 1) pv = sl.p[pc];
        else if (kind == 2 && o == sl.ore[pc]) pv = sl.p[pc];
        return pv;
    }
    // calc_p_config_open (rg:315-343)
    LDO_HDN double rg_calc_p_config_open(V3 p, int o) {
        double de = sys.check_domain_constraints(W()->d, p, o);
        if (sys.S()->constraints_violated) {
            sys.S()->constraints_violated = 0;
            return 0;
        }
        if (!cp_walks_remain(W()->d, p)) return 0;
        int j = sys.occupant(p);
        if (j >= 0 && sys.S()->dom[j].state == ST_UNBOUND) {
            bool same_chain = sys.chain(j) == sys.chain(W()->d);
            bool endpoint = cp_endpoint_reached(W()->d, p);
            if (!(same_chain || endpoint || W()->stemd)) return 0;
        }
        return fmin(1.0, exp(-de));
    }
    // test_config_open (rg:345-361)
    LDO_HDC bool rg_test_config_open(double p) {
        p = fmin(1.0, p);
        if (p == 1) return true;
        return p > uniform_real();
    }
    // Open probability of trial configuration ci according to the current slot (what rg_trial returns)
    LDO_HD double rg_slot_p(const RgSlot& sl, int ci) const {
        int pc = ci / 6, o = ci - 6 * pc;
        int kind = sl.kind[pc];
        if (kind == 1) return sl.p[pc];
        if (kind == 2 && o == sl.ore[pc]) return sl.p[pc];
        return 0;
    }
    // The trial loop of recoil_regrow (rg:193-198) in production (Philox) mode with exhaustive trials
    // (max_c_attempts == 36, so the attempts left always equal the configurations left). The loop tries
    // the untried configurations in a uniformly random order until one opens; its outcome is
    //   * "none opens, all tried" without any draw when no untried configuration has p > 0 (the common
    //     dead end on a crowded lattice, which the serial loop pays up to 36 iterations for);
    //   * otherwise the first open configuration in a random order: every untried configuration gets an
    //     i.i.d. random key (its place in the order) and an independent Bernoulli(p) outcome, both from
    //     a private Philox block of its lane; the winner is the open configuration with the smallest
    //     key, c_attempts advances by the number of keys not above it and those configurations leave
    //     the untried set. Same distribution as the serial loop, different stream consumption.
    // When most untried configurations can open the serial loop ends after a draw or two and is kept.
    // Returns whether a configuration opened.
    LDO_HDS bool rg_select_open_config(V3& p, int& o, double& p_c_open) {
        const RgSlot& sl = M()->slots[W()->cur_slot];
        unsigned long long rem = W()->avail;
        int n_rem = popc36(rem);
        // untried configurations that can open: 6 per empty site with a walk left, 1 per bindable site
        int n_pos = 0;
#pragma unroll 1
        for (int pc = 0; pc < 6; pc++) {
            if (sl.kind[pc] == 0 || !(sl.p[pc] > 0)) continue;
            unsigned site = (unsigned)(rem >> (6 * pc)) & 63u;
            if (sl.kind[pc] == 1) n_pos += popc36(site);
            else n_pos += (site >> sl.ore[pc]) & 1u;
        }
        if (n_pos == 0) {
            W()->c_attempts += n_rem;
            W()->avail = 0;
            W()->last_pc = -1;
            W()->last_kind = 0;
            return false;
        }
        if (4 * n_pos >= n_rem) {
            bool c_open = false;
#pragma unroll 1
            while (!c_open && W()->c_attempts != W()->d_max_c_attempts) {
                W()->c_attempts++;
                p_c_open = rg_trial(p, o);
                c_open = rg_test_config_open(p_c_open);
            }
            return c_open;
        }
        // lane L owns configurations L and L + 32; words 0,1 / 2,3 of its block: key and Bernoulli draw
        const int PER = (36 + LDO_NLANES - 1) / LDO_NLANES;
        uint32_t key[PER];
        unsigned long long ctr = RNG()->counter;
        uint32_t best = 0xffffffffu; // smallest key of an open configuration of this lane, low 6 bits = ci
        int slot_i = 0;
        const Rng* g = RNG();
#if defined(__CUDA_ARCH__)
        uint32_t w[4];
        philox4x32_10(g->key0, g->key1, g->subseq, 0x54520000u + (uint32_t)LDO_LANE, ctr, w);
#endif
#pragma unroll 1
        for (int ci = LDO_LANE; ci < 36; ci += LDO_NLANES, slot_i++) {
            key[slot_i] = 0xffffffffu;
            if (!((rem >> ci) & 1ull)) continue;
#if !defined(__CUDA_ARCH__)
            uint32_t w[4];
            philox4x32_10(g->key0, g->key1, g->subseq, 0x54520000u + (uint32_t)(ci & 31), ctr, w);
#endif
            uint32_t k = ((ci < 32 ? w[0] : w[2]) & ~63u) | (uint32_t)ci;
            if (k == 0xffffffffu) k -= 64u;
            key[slot_i] = k;
            double pv = rg_slot_p(sl, ci);
            double u = (double)(ci < 32 ? w[1] : w[3]) * (1.0 / 4294967296.0);
            if (pv > 0 && (pv == 1.0 || pv > u) && k < best) best = k;
        }
#if defined(__CUDA_ARCH__)
        best = __reduce_min_sync(0xffffffffu, best);
#endif
        // configurations tried up to and including the winner (all of them when none opened)
        int tried = 0;
        unsigned lo_mask = 0, hi_mask = 0;
        slot_i = 0;
#pragma unroll 1
        for (int ci = LDO_LANE; ci < 36; ci += LDO_NLANES, slot_i++) {
            if (key[slot_i] == 0xffffffffu || key[slot_i] > best) continue;
            tried++;
            if (ci < 32) lo_mask |= 1u << ci;
            else hi_mask |= 1u << (ci - 32);
        }
#if defined(__CUDA_ARCH__)
        tried = __reduce_add_sync(0xffffffffu, tried);
        lo_mask = __reduce_or_sync(0xffffffffu, lo_mask);
        hi_mask = __reduce_or_sync(0xffffffffu, hi_mask);
#endif
        LDO_SYNCWARP();
        RNG()->counter = ctr + 1;
        W()->c_attempts += tried;
        W()->avail = rem & ~(((unsigned long long)hi_mask << 32) | lo_mask);
        LDO_SYNCWARP();
        if (best == 0xffffffffu) {
            W()->last_pc = -1;
            W()->last_kind = 0;
            return false;
        }
        int ci = (int)(best & 63u);
        int pc = ci / 6;
        o = ci - 6 * pc;
        p = ore_vec(pc) + rec_pos(sys.S()->dom[W()->ref_d]);
        W()->last_pc = pc;
        W()->last_kind = sl.kind[pc];
        p_c_open = sl.p[pc];
        return true;
    }
    // Keeps the current level's slot for calc_weights (ColdScratch::slot_cache); lanes copy the 64 bytes together
    LDO_HD void rg_save_slot() {
#ifndef LDO_NO_SLOT_CACHE // A/B knob (profiles/ab_r2.txt)
        if (W()->stemd || sys.SC().cyclic) {
            C()->slot_cached[W()->di] = 0;
            return;
        }
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&M()->slots[W()->cur_slot]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&C()->slot_cache[W()->di]);
#pragma unroll 1
        for (int k = LDO_LANE; k < (int)(sizeof(RgSlot) / 4); k += LDO_NLANES) dst[k] = src[k];
        C()->slot_cached[W()->di] = 1;
        LDO_SYNCWARP();
#endif
    }
    LDO_HD bool rg_load_cached_slot(bool allowed) {
        if (!allowed || !C()->slot_cached[W()->di]) return false;
        LDO_COUNT(12);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&C()->slot_cache[W()->di]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&M()->slots[W()->cur_slot]);
        LDO_SYNCWARP();
#pragma unroll 1
        for (int k = LDO_LANE; k < (int)(sizeof(RgSlot) / 4); k += LDO_NLANES) dst[k] = src[k];
        LDO_SYNCWARP();
        return true;
    }
    // recoil_regrow (rg:177-231)
    LDO_HDN double rg_recoil_regrow() {
        double de = 0;
        W()->di = 0;
        W()->d = M()->regrow[0];
        W()->dir = cp_get_dir(W()->d);
        M()->c_opens[0] = 1;
        W()->in_growth = 1;
        rg_prepare_for_growth();
        int recoils = 0;
#pragma unroll 1
        for (;;) {
            if (sys.S()->status != LDO_OK) {
                M()->rejected = 1;
                break;
            }
            V3 p = v3(0, 0, 0);
            int o = 0;
            double p_c_open = 0;
            bool c_open = false;
            if (!LDO_SERIAL_DRAWS() && !W()->stemd && W()->d_max_c_attempts == 36) {
                c_open = rg_select_open_config(p, o, p_c_open);
            }
            else {
#pragma unroll 1
                while (!c_open && W()->c_attempts != W()->d_max_c_attempts) {
                    W()->c_attempts++;
                    p_c_open = rg_trial(p, o);
                    c_open = rg_test_config_open(p_c_open);
                }
            }
            if (c_open) {
                if (recoils != 0) recoils--;
                rg_save_slot();
                {
                    int sp0 = sys.S()->num_stacked_pairs;
                    double dset = rg_set_config(W()->d, p, o);
                    C()->set_de[W()->di] = dset;
                    C()->set_sp[W()->di] = (short)(sys.S()->num_stacked_pairs - sp0);
                    de += dset;
                }
                M()->c_attempts_q[W()->di] = (uint8_t)W()->c_attempts;
                M()->avail_q[W()->di] = W()->avail;
                M()->c_opens[W()->di] = p_c_open;
                if (W()->di == M()->n_regrow - 1) break;
                rg_prepare_for_growth();
            }
            else {
                if (recoils == W()->max_recoils || W()->di == 1) {
                    M()->rejected = 1;
                    break;
                }
                recoils++;
                de += rg_prepare_for_regrowth();
            }
        }
        W()->in_growth = 0;
        return de;
    }
    // test_config_avail (rg:422-480)
    LDO_HDS bool rg_test_config_avail() {
        LDO_COUNT(13);
        int feels = 0;
        if (feels == W()->max_recoils || W()->di == M()->n_regrow - 1) return true;
        rg_prepare_for_growth();
        bool c_avail = false;
        if (!LDO_SERIAL_DRAWS() && W()->max_recoils == 1 && !W()->stemd && W()->d_max_c_attempts == 36) {
            // Philox mode, one exhaustive feeler level: "some configuration opens" has probability
            // 1 - prod(1 - p) whatever the trial order, so one draw replaces up to 36 trials
            const RgSlot& sl = M()->slots[W()->cur_slot];
            double none = 1.0;
#pragma unroll 1
            for (int k = 0; k < 6; k++) {
                if (sl.kind[k] != 0) none *= 1.0 - sl.p[k];
            }
            c_avail = rg_test_config_open(1.0 - none);
            W()->di--;
            W()->d = M()->regrow[W()->di];
            return c_avail;
        }
#pragma unroll 1
        for (;;) {
            if (sys.S()->status != LDO_OK) break;
            V3 p = v3(0, 0, 0);
            int o = 0;
            bool c_open = false;
            double p_c_open;
#pragma unroll 1
            while (!c_open && W()->c_attempts != W()->d_max_c_attempts) {
                W()->c_attempts++;
                p_c_open = rg_trial(p, o);
                c_open = rg_test_config_open(p_c_open);
            }
            if (c_open) {
                feels++;
                if (feels == W()->max_recoils || W()->di == M()->n_regrow - 1) {
                    feels--;
                    c_avail = true;
                    break;
                }
                rg_set_config(W()->d, p, o);
                M()->c_attempts_q[W()->di] = (uint8_t)W()->c_attempts;
                M()->avail_q[W()->di] = W()->avail;
                rg_prepare_for_growth();
            }
            else {
                if (feels == 0) {
                    c_avail = false;
                    break;
                }
                feels--;
                rg_prepare_for_regrowth();
            }
        }
#pragma unroll 1
        while (feels != 0) {
            W()->di--;
            feels--;
            W()->d = M()->regrow[W()->di];
            sys.unassign_domain(W()->d);
            // The reference leaves these feelers in m_assigned_domains (an unbounded vector); the entries
            // are dropped here to bound the list. Equivalent: unassigning twice is a no-op
            // (origami_system.cpp:723-727) and every feeler is also a modified domain.
            if (M()->n_assigned > 0) M()->n_assigned--;
            rg_restore_endpoints();
        }
        W()->di--;
        W()->d = M()->regrow[W()->di];
        return c_avail;
    }
    // calc_weights (rg:363-417)
    LDO_HDN void rg_calc_weights() {
        W()->di = 0;
        W()->d = M()->regrow[0];
#pragma unroll 1
        while (W()->di != M()->n_regrow - 1) {
            if (sys.S()->status != LDO_OK) return;
            W()->di++;
            W()->d = M()->regrow[W()->di];
            W()->stemd = M()->stem_gp[W()->d] >= 0;
            W()->dir = cp_get_dir(W()->d);
            int avail_cs = 1;
            if (!W()->stemd) {
                W()->ref_d = sys.step(W()->d, -W()->dir);
                int catt = C()->c_attempts_wq[W()->di];
                W()->avail = C()->avail_wq[W()->di];
                W()->memo_level = -1;
                W()->memo_mask = 0;
                W()->cur_slot = W()->di & (LDO_RG_OWN_SLOTS - 1);
                if (catt != W()->max_c_attempts && !rg_load_cached_slot(W()->slot_cache_on != 0)) rg_compute_slot(W()->cur_slot);
                // With one feeler level (max_num_recoils == 1) an open trial configuration on an EMPTY
                // site only needs "does the next domain have an open configuration"; that depends on the
                // site, not on the orientation, so after the first orientation of a site has been
                // examined for real the remaining ones replay the feeler's draws on the memoised slot
                // without touching the lattice. (The reference sets and unsets the domain each time,
                // rg:378-402; the net state change is nil.)
                bool last_level = W()->di == M()->n_regrow - 1;
                bool feeler_simple = false;
                V3 feeler_refp = v3(0, 0, 0);
                bool feeler_ref_is_parent = false;
                if (!last_level && W()->max_recoils == 1 && catt != W()->max_c_attempts) {
                    int fd = M()->regrow[W()->di + 1];
                    if (M()->stem_gp[fd] < 0) {
                        int fref = sys.step(fd, -cp_get_dir(fd));
                        feeler_simple = true;
                        feeler_ref_is_parent = fref == W()->d;
                        if (!feeler_ref_is_parent) feeler_refp = rec_pos(sys.S()->dom[fref]);
                        rg_fill_feeler_memo(fd, fref);
                    }
                }
                // Production (Philox) mode with exhaustive trials: every remaining configuration is
                // examined whatever the order, and its contribution is an independent Bernoulli variable,
                // so the configurations are spread over the lanes, each with its own Philox block, and the
                // count is reduced across the warp. Replay (tape) mode keeps the reference's serial order.
                if (!LDO_SERIAL_DRAWS() && W()->max_c_attempts == 36 && W()->max_recoils == 1 && catt != W()->max_c_attempts &&
                    (last_level || feeler_simple)) {
                    avail_cs += rg_count_avail_parallel(last_level);
                    catt = W()->max_c_attempts;
                }
#pragma unroll 1
                while (catt != W()->max_c_attempts) {
                    catt++;
                    V3 p;
                    int o;
                    double p_c_open = rg_trial(p, o);
                    if (rg_test_config_open(p_c_open)) {
                        if (last_level && W()->max_recoils >= 1) {
                            avail_cs += 1;
                            continue;
                        }
                        if (feeler_simple && W()->last_kind == 1 && ((W()->memo_mask >> W()->last_pc) & 1) &&
                            (feeler_ref_is_parent || abssum(p - feeler_refp) != 1)) {
                            avail_cs += rg_feeler_from_slot(LDO_RG_OWN_SLOTS + W()->last_pc) ? 1 : 0;
                            continue;
                        }
                        avail_cs += rg_feeler_general(p, o) ? 1 : 0;
                    }
                }
            }
            W()->weight *= avail_cs / M()->c_opens[W()->di - 1];
            const DomRec& r = C()->prev[W()->d];
            sys.set_checked_domain_config(W()->d, rec_pos(r), r.ore);
            cp_update_endpoints(W()->d);
            eq_push_erased();
        }
        W()->weight /= M()->c_opens[W()->di];
    }
    // One open trial configuration of calc_weights examined the reference's way: place the domain, grow
    // feelers, take it back (rg:384-400)
    LDO_HDN bool rg_feeler_general(V3 p, int o) {
        LDO_COUNT(10);
        sys.set_checked_domain_config(W()->d, p, o);
        cp_update_endpoints(W()->d);
        eq_push_erased();
        int dir_s = W()->dir, ref_s = W()->ref_d, stem_s = W()->stemd, slot_s = W()->cur_slot;
        unsigned long long avail_s = W()->avail;
        W()->memo_level = W()->di + 1;
        W()->memo_key = W()->last_kind == 1 ? W()->last_pc : -1;
        bool c_avail = rg_test_config_avail();
        W()->memo_key = -1;
        W()->avail = avail_s;
        W()->stemd = stem_s;
        W()->ref_d = ref_s;
        W()->dir = dir_s;
        W()->cur_slot = slot_s;
        sys.unassign_domain(W()->d);
        rg_restore_endpoints();
        return c_avail;
    }
    // Lane-parallel count of the available configurations among the remaining ones (Philox mode,
    // max_c_attempts == 36, one feeler level). A configuration is available when it is open
    // (probability p) and, unless it is the last level, the next domain finds an open configuration
    // among all 36 of its own: probability 1 - prod(1 - p') over the feeler slot, independent of the
    // order in which the reference would have tried them.
    LDO_HDS int rg_count_avail_parallel(bool last_level) {
        const RgSlot& own = M()->slots[W()->cur_slot];
        // feeler availability probability per parent site (warp-uniform)
        double pav[6];
#pragma unroll 1
        for (int pc = 0; pc < 6; pc++) {
            pav[pc] = -1; // not memoised: needs the general path
            if (last_level || !((W()->memo_mask >> pc) & 1)) continue;
            const RgSlot& sl = M()->slots[LDO_RG_OWN_SLOTS + pc];
            double none = 1.0;
#pragma unroll 1
            for (int k = 0; k < 6; k++) {
                if (sl.kind[k] != 0) none *= 1.0 - sl.p[k]; // kind 1: six orientations share p (0 or 1)
            }
            pav[pc] = 1.0 - none;
        }
        unsigned long long rem = W()->avail;
        unsigned long long ctr = RNG()->counter;
        int count = 0;
        unsigned lo_mask = 0, hi_mask = 0; // configurations left to the general path
#pragma unroll 1
        for (int ci = LDO_LANE; ci < 36; ci += LDO_NLANES) {
            if (!((rem >> ci) & 1ull)) continue;
            int pc = ci / 6, o = ci - 6 * pc;
            int kind = own.kind[pc];
            double pv = 0;
            if (kind == 1) pv = own.p[pc];
            else if (kind == 2 && o == own.ore[pc]) pv = own.p[pc];
            if (pv == 0) continue;
            // private Philox block of this configuration: stream word tagged with the configuration index
            const Rng* g = RNG();
            uint32_t w[4];
            philox4x32_10(g->key0, g->key1, g->subseq, 0x52470000u + (uint32_t)ci, ctr, w);
            double ua = (double)((((unsigned long long)w[0] << 32) | w[1]) >> 11) * (1.0 / 9007199254740992.0);
            double ub = (double)((((unsigned long long)w[2] << 32) | w[3]) >> 11) * (1.0 / 9007199254740992.0);
            if (!(pv == 1.0 || pv > ua)) continue;
            if (last_level) {
                count++;
            }
            else if (kind == 1 && pav[pc] >= 0) {
                if (pav[pc] == 1.0 || pav[pc] > ub) count++;
            }
            else if (ci < 32) {
                lo_mask |= 1u << ci;
            }
            else {
                hi_mask |= 1u << (ci - 32);
            }
        }
#if defined(__CUDA_ARCH__)
        count = __reduce_add_sync(0xffffffffu, count);
        lo_mask = __reduce_or_sync(0xffffffffu, lo_mask);
        hi_mask = __reduce_or_sync(0xffffffffu, hi_mask);
#endif
        // the private blocks live on their own stream words: buffered words of the main stream stay valid
        LDO_SYNCWARP();
        RNG()->counter = ctr + 1;
        LDO_SYNCWARP();
        // configurations that bind the parent (at most one orientation per site) go through the real
        // place / feel / take-back path, serially
        unsigned long long todo = ((unsigned long long)hi_mask << 32) | lo_mask;
#pragma unroll 1
        while (todo) {
            int ci = nth_set_bit36(todo, 0);
            todo &= todo - 1;
            int pc = ci / 6, o = ci - 6 * pc;
            W()->last_pc = pc;
            W()->last_kind = own.kind[pc];
            V3 p = ore_vec(pc) + rec_pos(sys.S()->dom[W()->ref_d]);
            count += rg_feeler_general(p, o) ? 1 : 0;
        }
        W()->avail = 0;
        return count;
    }
    // test_config_avail (rg:422-480) for a single feeler level, replayed on an already computed slot:
    // same draws as the general path, no lattice updates
    LDO_HDS bool rg_feeler_from_slot(int slot) {
        const RgSlot& sl = M()->slots[slot];
        unsigned long long av = all_cis();
#pragma unroll 1
        for (int catt = 0; catt != W()->max_c_attempts; catt++) {
            int ci = uniform_int(0, popc36(av) - 1);
            int i = nth_set_bit36(av, ci);
            av &= ~(1ull << i);
            int pc = i / 6;
            int o = i - 6 * pc;
            int kind = sl.kind[pc];
            double pv = 0;
            if (kind == 1) pv = sl.p[pc];
            else if (kind == 2 && o == sl.ore[pc]) pv = sl.p[pc];
            if (rg_test_config_open(pv)) return true;
        }
        return false;
    }
    // calc_old_c_opens (rg:482-513)
    LDO_HDN void rg_calc_old_c_opens() {
        W()->di = 0;
        M()->c_opens[0] = 1;
#pragma unroll 1
        while (W()->di != M()->n_regrow - 1) {
            W()->di++;
            W()->d = M()->regrow[W()->di];
            M()->c_attempts_q[W()->di] = 1;
            const DomRec r = C()->oldc[W()->d];
            V3 p = rec_pos(r);
            W()->stemd = M()->stem_gp[W()->d] >= 0;
            C()->slot_cached[W()->di] = 0;
            bool from_slot = false;
            if (!W()->stemd) {
                W()->dir = cp_get_dir(W()->d);
                W()->ref_d = sys.step(W()->d, -W()->dir);
                V3 rel0 = p - rec_pos(sys.S()->dom[W()->ref_d]);
                int pc0 = ore_code(rel0);
                if (W()->slot_cache_on && pc0 < 6 && r.ore >= 0 && r.ore < 6 && !sys.SC().cyclic) {
                    // the six sites of this level evaluated once, for the open probability of the old configuration here
                    // and for calc_weights afterwards (only when both run with the same active endpoints: the caller says)
                    W()->cur_slot = W()->di & (LDO_RG_OWN_SLOTS - 1);
                    rg_compute_slot(W()->cur_slot);
                    M()->c_opens[W()->di] = rg_slot_p(M()->slots[W()->cur_slot], pc0 * 6 + r.ore);
                    rg_save_slot();
                    from_slot = true;
                }
            }
            if (!from_slot) M()->c_opens[W()->di] = rg_calc_p_config_open(p, r.ore);
            rg_set_config(W()->d, p, r.ore);
            if (W()->stemd) {
                M()->avail_q[W()->di] = 0;
            }
            else {
                V3 rel = p - rec_pos(C()->oldc[W()->ref_d]);
                int pc = ore_code(rel);
                int ci = pc * 6 + r.ore;
                unsigned long long a = all_cis();
                if (pc < 6 && r.ore >= 0 && r.ore < 6) a &= ~(1ull << ci);
                else sys.fail(LDO_ERR_INTERNAL, 2); // m_config_to_i.at(c) would throw
                M()->avail_q[W()->di] = a;
            }
        }
    }
    LDO_HD void rg_copy_queues_to_wq() {
#pragma unroll 1
        for (int k = 0; k < M()->n_regrow; k++) {
            C()->c_attempts_wq[k] = M()->c_attempts_q[k];
            C()->avail_wq[k] = M()->avail_q[k];
        }
    }

    // Shared tail of the two CTRG scaffold moves (rg:636-689, 810-851). `whole_cyclic`: the
    // endpoint-removal guards evaluated by the caller (App. A2 keeps the contiguous variant's quirk).
    LDO_HDN bool rg_regrow_and_test(bool remove_first_a, bool remove_first_b, int first_dom) {
        // The weight passes below tear the regrown domains down and put them back five times (new configuration,
        // old configuration, new configuration again on acceptance). The reference evaluates the binding potential
        // for every one of those placements and removals (set_checked_domain_config / unassign_domain,
        // rg_movetypes.cpp:363-417, 482-533) although the energy changes cancel and nobody reads them: only the
        // lattice, the domain records and the pair counters matter to the trial evaluations in between. Here the
        // passes run with SysState::weight_pass set - no potential evaluation for placements and removals - and the
        // running energy and the stacked-pair count are restored from snapshots: those of the old configuration on
        // rejection, those of the new one on acceptance. Same decisions, same lattice; the running energy differs
        // from the reference's by the rounding of terms that cancel (it agrees to 1e-12, as everywhere).
        const double e_old = sys.S()->energy;
        const int sp_old = sys.S()->num_stacked_pairs;
        W()->slot_cache_on = 0;
        W()->in_growth = 0;
        W()->delta_e += rg_unassign_and_save_domains();
        W()->delta_e += rg_recoil_regrow();
        if (M()->rejected) return false;
        // excluded staples: the reference hard-codes zero of them (simulation.cpp:410,527)
        update_move_params();
        W()->delta_e += calc_move_bias();
        const double e_new = sys.S()->energy;
        const int sp_new = sys.S()->num_stacked_pairs;
#ifndef LDO_NO_WEIGHT_PASS // A/B knob (profiles/ab_r2.txt): evaluate the potential in the weight passes like the reference
        sys.S()->weight_pass = 1;
#endif

        // new-configuration weights (setup_for_calc_new_weights, rg:147-153)
        rg_copy_queues_to_wq();
#pragma unroll 1
        for (int k = 0; k < M()->n_regrow; k++) C()->oldc[M()->regrow[k]] = C()->prev[M()->regrow[k]];
        M()->n_modified = 0;
        rg_unassign_and_save_domains();
        cp_reset_active_endpoints();
        if (remove_first_a) cp_remove_active_endpoint(first_dom);
        // the growth above ran from the endpoints setup_constraints left (the first domain's removed, rg:134-145); this
        // pass starts from the same set when its own guard removes it too
        W()->slot_cache_on = remove_first_a ? 1 : 0;
        rg_calc_weights();

        // old-configuration weights
        rg_unassign_domains();
        cp_reset_active_endpoints();
        if (remove_first_b) cp_remove_active_endpoint(first_dom);
        // calc_old_c_opens and the weight pass after it see the same endpoints only when their guards agree (they do not
        // for the contiguous move on a linear scaffold, App. A2)
        W()->slot_cache_on = remove_first_a == remove_first_b ? 1 : 0;
        rg_calc_old_c_opens();
        // setup_for_calc_old_weights (rg:155-163)
        W()->weight_new = W()->weight;
        W()->weight = 1;
        rg_copy_queues_to_wq();
#pragma unroll 1
        for (int k = 0; k < M()->n_regrow; k++) C()->newc[M()->regrow[k]] = C()->prev[M()->regrow[k]];
        M()->n_modified = 0;
        rg_unassign_and_save_domains();
        cp_reset_active_endpoints();
        if (remove_first_a) cp_remove_active_endpoint(first_dom);
        rg_calc_weights();
        W()->slot_cache_on = 0;

        // test_rg_acceptance (rg:515-533)
        LDO_COUNT(15);
        double ratio = W()->weight_new / W()->weight * exp(-W()->delta_e);
        bool accepted = test_acceptance(ratio);
        if (accepted) {
#pragma unroll 1
            for (int k = 0; k < M()->n_regrow; k++) C()->prev[M()->regrow[k]] = C()->newc[M()->regrow[k]];
            reset_origami();
        }
        sys.S()->weight_pass = 0;
        sys.S()->energy = accepted ? e_new : e_old;
        sys.S()->num_stacked_pairs = accepted ? sp_new : sp_old;
        if (accepted) return true;
        M()->n_modified = 0;
        M()->n_assigned = 0;
        return false;
    }

    LDO_HD void rg_reset(const MoveDef& md) {
        cp_reset();
        W()->delta_e = 0;
        W()->weight = 1;
        W()->weight_new = 1;
        W()->max_recoils = md.max_num_recoils;
        W()->max_c_attempts = md.max_c_attempts;
        W()->slot_cache_on = 0;
        W()->memo_level = -1;
        W()->memo_key = -1;
        W()->memo_mask = 0;
        W()->cur_slot = 0;
        M()->eq_depth = 0;
        M()->eq_npos = 0;
    }

    // CTRGScaffoldRegrowthMCMovetype::internal_attempt_move (rg:629-689)
    LDO_HDN bool move_ctrg_scaffold(const MoveDef& md) {
        rg_reset(md);
        ct_select_indices(md);
        if (sys.S()->status != LDO_OK) return false;
        W()->trk_a = C()->n_sel;
        // setup_constraints (rg:134-145)
#pragma unroll 1
        for (int k = 0; k < C()->n_sel; k++) C()->in_sel[C()->sel_scaf[k]] = 1;
        M()->scaf_dir[0] = (int8_t)W()->dir;
        cp_find_growthpoints_endpoints(C()->sel_scaf, C()->n_sel, 0);
        cp_save_initial();
        int n_scaf = sys.S()->chain_len[0];
        bool cyc = sys.SC().cyclic != 0;
        if (!(cyc && C()->n_sel == n_scaf)) cp_remove_active_endpoint(C()->sel_scaf[0]);
        bool guard_a = !(cyc && M()->n_regrow == n_scaf);
        bool guard_b = !cyc && M()->n_regrow == n_scaf; // sic (rg:671, App. A2)
        return rg_regrow_and_test(guard_a, guard_b, M()->regrow[0]);
    }

    // CTRGJumpScaffoldRegrowthMCMovetype::internal_attempt_move (rg:791-851)
    // What select_noncontig_segs registers (movetypes.cpp:618-647) followed by
    // calculate_constraintpoints(segs, dirs, excluded) (top_constraint_points.cpp:223-248)
    LDO_HDN void rg_register_jump_constraints(int n_segs, int n_stems) {
        SysState<K>* s = sys.S();
        // endpoints registered by select_noncontig_segs (movetypes.cpp:618-647)
        {
            int last = C()->seg_dom[C()->seg_start[1] - 1];
            int nd = sys.step(last, M()->scaf_dir[0]);
            if (nd >= 0) cp_add_active_endpoint_seg(nd, rec_pos(s->dom[nd]), 0);
        }
#pragma unroll 1
        for (int k = 0; k < n_stems; k++) {
            int stem = C()->stems[k];
            int gp = s->bound[stem];
            M()->gp_stem[gp] = (short)stem;
            M()->stem_gp[stem] = (short)gp;
            int seg_i = 1 + 2 * k;
            M()->stem_seg0[stem] = (int8_t)seg_i;
#pragma unroll 1
            for (int q = 0; q < 2; q++) {
                int sg = seg_i + q;
                int a = C()->seg_start[sg], b = C()->seg_start[sg + 1];
                // paired_segs hold the filled domains only (no stem prefix)
                int last = stem;
                if (b > a && !(b - a == 1 && C()->seg_dom[a] == stem)) last = C()->seg_dom[b - 1];
                int nd = sys.step(last, M()->scaf_dir[sg]);
                if (nd >= 0) cp_add_active_endpoint_seg(nd, rec_pos(s->dom[nd]), sg);
            }
        }
        // calculate_constraintpoints(segs, dirs, excluded) (top_constraint_points.cpp:223-248)
#pragma unroll 1
        for (int sg = 0; sg < n_segs; sg++) {
            int a = C()->seg_start[sg], b = C()->seg_start[sg + 1];
            if (b > a) cp_find_growthpoints_endpoints(C()->seg_dom + a, b - a, sg);
            // m_domain_to_dir is only filled for non-empty segments (:234-244); an empty one reads as
            // direction 0, which matters to calc_remaining_steps on cyclic scaffolds (:585-587)
            else M()->scaf_dir[sg] = 0;
        }
        cp_save_initial();
    }
    LDO_HDN bool move_ctrg_jump_scaffold(const MoveDef& md) {
        SysState<K>* s = sys.S();
        rg_reset(md);
        int n_stems = 0;
        int n_segs = ct_select_noncontig_segs(md, n_stems);
        if (s->status != LDO_OK) return false;
        W()->trk_a = 0;
        rg_register_jump_constraints(n_segs, n_stems);
        int first = C()->seg_dom[0];
        cp_remove_active_endpoint(first);
#pragma unroll 1
        for (int k = 0; k < n_stems; k++) {
            int stem = C()->stems[k];
            int gp = s->bound[stem];
            M()->gp_stem[gp] = (short)stem;
            M()->stem_gp[stem] = (short)gp;
        }
        return rg_regrow_and_test(true, true, first);
    }

    // ---- CTCB scaffold regrowth (cb_movetypes.cpp:463-983) ----
    // CTCBRegrowthMCMovetype::calc_bias (:486-539) on the six site weights of cb_site_weights
    LDO_HDN void ctcb_select_and_set_config(int dom, int prev_dom, bool regrow_old, DD& bias) {
        V3 p_prev = rec_pos(sys.S()->dom[prev_dom]);
        cb_site_weights(p_prev, dom);
        sys.S()->constraints_violated = 0;
        DD sum = dd_from(0.0);
        for (int k = 0; k < 6; k++) {
            if (M()->site_kind[k] == 0) continue;
            V3 cur = p_prev + ore_vec(k);
            double w = M()->site_w[k];
            if (M()->site_kind[k] == 2) {
                int j = sys.occupant(cur);
                bool same_chain = sys.chain(j) == sys.chain(dom);
                if (!(same_chain || cp_endpoint_reached(dom, cur))) w = 0;
            }
            if (!cp_walks_remain(dom, cur)) w = 0;
            M()->site_w[k] = w;
            sum = dd_add(sum, w);
        }
        if (sum.hi == 0) {
            M()->rejected = 1;
            return;
        }
        bias = dd_mul_dd(bias, sum);
        if (!regrow_old) {
            double cum = 0;
            double r = uniform_real();
            V3 p_new = v3(0, 0, 0);
            int o_new = ORE_ZERO;
            for (int k = 0; k < 6; k++) {
                if (M()->site_kind[k] == 0) continue;
                cum += dd_quot(M()->site_w[k], sum);
                if (r < cum) {
                    p_new = p_prev + ore_vec(k);
                    o_new = M()->site_kind[k] == 2 ? M()->site_o[k] : ORE_ZERO;
                    break;
                }
            }
            if (o_new == ORE_ZERO) o_new = uniform_int(0, 5);
            sys.set_checked_domain_config(dom, p_new, o_new);
        }
        else {
            const DomRec& r = C()->oldc[dom];
            sys.set_checked_domain_config(dom, rec_pos(r), r.ore);
        }
        push_assigned(dom);
    }
    // Explicit work stack of chain pieces to grow: (first domain, +-1, count, next index), replacing
    // the recursion grow_chain -> grow_staple_and_update_endpoints -> grow_staple -> grow_chain
    // (cb_movetypes.cpp:463-484, 541-562; movetypes.cpp:343-370)
    LDO_HD bool ctcb_push(int first, int stepdir, int count, int& sp) {
        if (count <= 1) return true;
        if (sp >= K::LV) {
            sys.fail(LDO_ERR_CAPACITY, 12);
            return false;
        }
        C()->work[sp][0] = (short)first;
        C()->work[sp][1] = (short)stepdir;
        C()->work[sp][2] = (short)count;
        C()->work[sp][3] = 1;
        sp++;
        return true;
    }
    // grow_staple_and_update_endpoints (:541-562): pushes the two pieces of the staple grown from growth_d_old
    LDO_HDN void ctcb_grow_staple_from(int growth_d_old, bool regrow_old, DD& bias, int& sp) {
        int g_new = M()->gp_stem[growth_d_old];
        int c = sys.chain(g_new);
        if (c == 0) return; // "HACK TO PREVENT ACCIDENTLY GROWING SCAFFOLD SEGMENTS" (:547-550)
        if (regrow_old) {
            double de = sys.set_checked_domain_config(g_new, rec_pos(sys.S()->dom[growth_d_old]), C()->oldc[g_new].ore);
            bias = dd_mul(bias, exp(-de));
            push_assigned(g_new);
        }
        else {
            double de = set_growth_point(g_new, growth_d_old);
            bias = dd_mul(bias, exp(-de));
        }
        if (M()->rejected) return;
        cp_update_endpoints(g_new);
        int base = sys.chain_base(c), len = sys.S()->chain_len[c], d_i = sys.dindex(g_new);
        // 3' piece is grown first: push the 5' piece below it
        ctcb_push(base + d_i, -1, d_i + 1, sp);
        ctcb_push(base + d_i, +1, len - d_i, sp);
    }
    // CTCBRegrowthMCMovetype::grow_chain over an explicit list of domains (scaffold segments)
    LDO_HDN void ctcb_run_stack(bool regrow_old, DD& bias, int& sp, int floor) {
#pragma unroll 1
        while (sp > floor && !M()->rejected && sys.S()->status == LDO_OK) {
            short* w = C()->work[sp - 1];
            int i = w[3];
            if (i >= w[2]) {
                sp--;
                continue;
            }
            w[3] = (short)(i + 1);
            int dom = w[0] + w[1] * i, prev = w[0] + w[1] * (i - 1);
            ctcb_select_and_set_config(dom, prev, regrow_old, bias);
            if (M()->rejected) break;
            cp_update_endpoints(dom);
            if (M()->gp_stem[dom] >= 0) ctcb_grow_staple_from(dom, regrow_old, bias, sp);
        }
    }
    LDO_HDN void ctcb_grow_list(const short* doms, int n, bool regrow_old, DD& bias) {
        int sp = 0;
#pragma unroll 1
        for (int i = 1; i < n && !M()->rejected && sys.S()->status == LDO_OK; i++) {
            ctcb_select_and_set_config(doms[i], doms[i - 1], regrow_old, bias);
            if (M()->rejected) break;
            cp_update_endpoints(doms[i]);
            if (M()->gp_stem[doms[i]] >= 0) {
                ctcb_grow_staple_from(doms[i], regrow_old, bias, sp);
                ctcb_run_stack(regrow_old, bias, sp, 0);
            }
        }
    }
    // staples_to_be_regrown is a std::set<int> of unique chain indices: ascending order (:167-169)
    LDO_HDN void ctcb_unassign(const short* doms, int n_doms, bool with_staples) {
#pragma unroll 1
        for (int k = 0; k < n_doms; k++) {
            int dd = doms[k];
            C()->prev[dd] = sys.S()->dom[dd];
            push_modified(dd);
            sys.unassign_domain(dd);
        }
        if (!with_staples) return;
        int last_uid = -1;
#pragma unroll 1
        for (;;) {
            int best = -1, best_uid = 0x7fffffff;
#pragma unroll 1
            for (int c = 1; c < K::C; c++) {
                if (!C()->regrow_chain[c]) continue;
                int uid = sys.S()->chain_uid[c];
                if (uid > last_uid && uid < best_uid) {
                    best = c;
                    best_uid = uid;
                }
            }
            if (best < 0) break;
            last_uid = best_uid;
            cb_unassign_domains(best);
        }
    }
    LDO_HDN void ctcb_setup_regrow_old(DD& bias, DD& new_bias) {
        // setup_for_regrow_old (cb_movetypes.cpp:237-247)
        new_bias = bias;
        bias = dd_from(1.0);
        M()->n_modified = 0;
        M()->n_assigned = 0;
#pragma unroll 1
        for (int k = 0; k < K::D; k++) C()->oldc[k] = C()->prev[k];
        cp_reset_active_endpoints();
    }
    LDO_HDN bool ctcb_finish(DD new_bias, DD bias) {
        M()->modifier = 1;
        double ratio = dd_ratio_min1(new_bias, bias);
        if (test_acceptance(ratio)) {
            reset_origami();
            update_move_params();
            calc_move_bias();
            return true;
        }
        M()->n_modified = 0;
        M()->n_assigned = 0;
        return false;
    }
    // CTCBScaffoldRegrowthMCMovetype::internal_attempt_move (:663-748)
    LDO_HDN bool move_ctcb_scaffold(const MoveDef& md) {
        cp_reset();
        ct_select_indices(md);
        if (sys.S()->status != LDO_OK) return false;
#pragma unroll 1
        for (int k = 0; k < C()->n_sel; k++) C()->in_sel[C()->sel_scaf[k]] = 1;
        M()->scaf_dir[0] = (int8_t)W()->dir;
        cp_find_growthpoints_endpoints(C()->sel_scaf, C()->n_sel, 0);
        cp_save_initial();
        bool whole = sys.SC().cyclic != 0 && C()->n_sel == sys.S()->chain_len[0];
        if (!whole) cp_remove_active_endpoint(C()->sel_scaf[0]);
        if (TRK()) {
            W()->trk_a = C()->n_sel;
            W()->trk_b = num_regrowth_staples();
        }
        DD bias = dd_from(1.0), new_bias = dd_from(1.0);
        for (int pass = 0; pass < 2; pass++) {
            bool regrow_old = pass == 1;
            if (regrow_old) {
                ctcb_setup_regrow_old(bias, new_bias);
                if (!whole) cp_remove_active_endpoint(C()->sel_scaf[0]);
            }
            ctcb_unassign(C()->sel_scaf + 1, C()->n_sel - 1, true);
            if (M()->gp_stem[C()->sel_scaf[0]] >= 0) {
                int sp = 0;
                ctcb_grow_staple_from(C()->sel_scaf[0], regrow_old, bias, sp);
                ctcb_run_stack(regrow_old, bias, sp, 0);
                if (M()->rejected && !regrow_old) return false;
            }
            ctcb_grow_list(C()->sel_scaf, C()->n_sel, regrow_old, bias);
            if (!regrow_old) {
                if (M()->rejected) return false;
                update_move_params();
                bias = dd_mul(bias, exp(-calc_move_bias()));
            }
        }
        return ctcb_finish(new_bias, bias);
    }
    // CTCBJumpScaffoldRegrowthMCMovetype::internal_attempt_move (:872-966)
    LDO_HDN bool move_ctcb_jump_scaffold(const MoveDef& md) {
        SysState<K>* s = sys.S();
        cp_reset();
        int n_stems = 0;
        int n_segs = ct_select_noncontig_segs(md, n_stems);
        if (s->status != LDO_OK) return false;
        rg_register_jump_constraints(n_segs, n_stems);
        int first = C()->seg_dom[0];
        cp_remove_active_endpoint(first);
        if (TRK()) {
            W()->trk_a = 0;
            W()->trk_b = num_regrowth_staples();
        }
        DD bias = dd_from(1.0), new_bias = dd_from(1.0);
        for (int pass = 0; pass < 2; pass++) {
            bool regrow_old = pass == 1;
            if (regrow_old) {
                ctcb_setup_regrow_old(bias, new_bias);
                cp_remove_active_endpoint(first);
            }
            ctcb_unassign(M()->regrow + 1, M()->n_regrow - 1, false);
            if (M()->gp_stem[first] >= 0) {
                int sp = 0;
                ctcb_grow_staple_from(first, regrow_old, bias, sp);
                ctcb_run_stack(regrow_old, bias, sp, 0);
                if (M()->rejected) return false;
            }
            ctcb_grow_list(C()->seg_dom + C()->seg_start[0], C()->seg_start[1] - C()->seg_start[0], regrow_old, bias);
            if (M()->rejected) return false;
#pragma unroll 1
            for (int k = 0; k < n_stems; k++) {
                int stem = C()->stems[k];
                // set_first_seg_domain (:973-983)
                int d_old = M()->stem_gp[stem];
                if (!cp_walks_remain(stem, rec_pos(s->dom[d_old]))) {
                    M()->rejected = 1;
                    return false;
                }
                double de = set_growth_point(stem, d_old);
                bias = dd_mul(bias, exp(-de));
                if (M()->rejected) return false;
                for (int q = 0; q < 2; q++) {
                    int sg = 1 + 2 * k + q;
                    int a = C()->seg_start[sg], b = C()->seg_start[sg + 1];
                    // paired segment with the stem prepended (:905-907); segment q == 0 already carries it
                    short tmp[2];
                    const short* list = C()->seg_dom + a;
                    int n = b - a;
                    if (n == 0 || C()->seg_dom[a] != stem) {
                        // build [stem] + seg in the scratch tail of seg_dom
                        short* buf = C()->seg_dom + 2 * K::D + 1 - (n + 1);
                        buf[0] = (short)stem;
                        for (int t = 0; t < n; t++) buf[1 + t] = C()->seg_dom[a + t];
                        list = buf;
                        n = n + 1;
                    }
                    (void)tmp;
                    ctcb_grow_list(list, n, regrow_old, bias);
                    if (M()->rejected) return false;
                }
            }
            if (!regrow_old) {
                update_move_params();
                bias = dd_mul(bias, exp(-calc_move_bias()));
            }
        }
        return ctcb_finish(new_bias, bias);
    }


    // ---- Linker regrowth: rigid transformation of a scaffold segment with its staples, followed by
    // regrowth of the two linkers that join it to the rest (transform_movetypes.cpp:14-1213) ----

    // scan_for_external_scaffold_domain (:262-314), iterative; `participating` starts as {scaffold}
    LDO_HDN bool lk_scan_external(int start) {
        short (*st)[2] = C()->scan_stack;
        int sp = 0;
        st[0][0] = (short)start;
        st[0][1] = 0;
        C()->net_chain[sys.chain(start)] = 1;
#pragma unroll 1
        while (sp >= 0) {
            int dom = st[sp][0];
            int c = sys.chain(dom);
            int base = sys.chain_base(c);
            int len = sys.S()->chain_len[c];
            bool descended = false;
#pragma unroll 1
            while (st[sp][1] < len) {
                int cur = base + st[sp][1];
                st[sp][1]++;
                if (cur == dom) continue;
                int b = sys.S()->bound[cur];
                if (b < 0) continue;
                if (sys.chain(b) == c) continue;
                if (sys.chain(b) == 0) {
                    if (!(C()->lk_mark[b] & 1)) return true;
                    continue;
                }
                if (C()->net_chain[sys.chain(b)]) continue;
                if (sp + 1 >= K::C) {
                    sys.fail(LDO_ERR_CAPACITY, 13);
                    return true;
                }
                C()->net_chain[sys.chain(b)] = 1;
                sp++;
                st[sp][0] = (short)b;
                st[sp][1] = 0;
                descended = true;
                break;
            }
            if (!descended) sp--;
        }
        return false;
    }
    // domains_bound_externally (:238-260) on the central segment C()->sel_scaf
    LDO_HDN bool lk_bound_externally() {
#pragma unroll 1
        for (int k = 0; k < K::D; k++) C()->lk_mark[k] = 0;
#pragma unroll 1
        for (int k = 0; k < C()->n_sel; k++) C()->lk_mark[C()->sel_scaf[k]] = 1;
#pragma unroll 1
        for (int k = 0; k < C()->n_sel; k++) {
            int dd = C()->sel_scaf[k];
            if (sys.S()->dom[dd].state == ST_UNBOUND) continue;
            int b = sys.S()->bound[dd];
            if (b < 0 || sys.chain(b) == 0) continue;
#pragma unroll 1
            for (int q = 0; q < K::C; q++) C()->net_chain[q] = 0;
            C()->net_chain[0] = 1;
            if (lk_scan_external(b)) return true;
        }
        return false;
    }
    // setup_fixed_end_biases (:210-236): terminal constraint points of the two linkers, then
    // calculate_constraintpoints(linkers without their first domain, {-dir, dir}, {})
    LDO_HDN void lk_setup_fixed_end_biases() {
        SysState<K>* s = sys.S();
        int dir_ = W()->dir;
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            int last = C()->lnk[q][C()->n_lnk[q] - 1];
            int ep = sys.step(last, q == 0 ? -dir_ : dir_);
            C()->lnk_ep[q] = (short)ep;
            if (ep >= 0) cp_add_active_endpoint_seg(ep, rec_pos(s->dom[ep]), q);
        }
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
#pragma unroll 1
            for (int k = 1; k < C()->n_lnk[q]; k++) C()->in_sel[C()->lnk[q][k]] = 1;
        }
        // the CTRG variant regrows from the end of the central segment: m_regrow_ds.insert(begin, linker1.front())
        M()->regrow[0] = C()->lnk[0][0];
        M()->n_regrow = 1;
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            if (C()->n_lnk[q] > 1) {
                M()->scaf_dir[q] = (int8_t)(q == 0 ? -dir_ : dir_);
                cp_find_growthpoints_endpoints(C()->lnk[q] + 1, C()->n_lnk[q] - 1, q);
            }
            else {
                M()->scaf_dir[q] = 0; // m_domain_to_dir has no entry for an empty segment
            }
        }
        cp_save_initial();
    }
    // LinkerRegrowthMCMovetype::select_and_setup_segments (:156-208)
    LDO_HDN void lk_select_and_setup(const MoveDef& md) {
        bool externally_bound = true;
        int attempts = 0;
#pragma unroll 1
        while (externally_bound && attempts != 10) {
            if (sys.S()->status != LDO_OK) return;
            ct_select_indices(md, 1);
            int dir_ = W()->dir;
            int front = C()->sel_scaf[0], back = C()->sel_scaf[C()->n_sel - 1];
            int l1 = uniform_int(2, md.max_linker_length + 1);
            int n = 0;
            int ld = front;
            int stop = sys.step(back, 2 * dir_);
#pragma unroll 1
            while (n != l1 && ld >= 0 && ld != stop) {
                C()->lnk[0][n++] = (short)ld;
                ld = sys.step(ld, -dir_);
            }
            C()->n_lnk[0] = n;
            if (sys.SC().cyclic && n == 1) {
                M()->rejected = 1;
                return;
            }
            int l2 = uniform_int(2, md.max_linker_length + 1);
            n = 0;
            ld = back;
            stop = sys.step(C()->lnk[0][C()->n_lnk[0] - 1], -dir_);
#pragma unroll 1
            while (n != l2 && ld >= 0 && ld != stop) {
                C()->lnk[1][n++] = (short)ld;
                ld = sys.step(ld, dir_);
            }
            C()->n_lnk[1] = n;
            externally_bound = lk_bound_externally();
            attempts++;
        }
        if (externally_bound) {
            M()->rejected = 1;
            return;
        }
        lk_setup_fixed_end_biases();
    }
    // ClusteredLinkerRegrowthMCMovetype::linear/cyclic_select_central_segment (:671-775): the maximal run
    // of bound scaffold domains around (or next to) a random kernel domain, into C()->sel_scaf
    LDO_HDN void lk_select_cluster() {
        SysState<K>* s = sys.S();
        bool cyc = sys.SC().cyclic != 0;
        int n = s->chain_len[0];
        W()->dir = 1;
        short* seg = C()->sel_scaf;
        int m = 0;
        int kd = sys.chain_base(0) + uniform_int(1, n - 2);
        bool started = false;
        if (s->dom[kd].state == ST_BOUND) {
            started = true;
            seg[m++] = (short)kd;
            int p = sys.step(kd, -1);
#pragma unroll 1
            while ((cyc || p >= 0) && s->dom[p].state == ST_BOUND && m != n - 1) {
                seg[m++] = (short)p;
                p = sys.step(p, -1);
            }
        }
        if (m == n - 1) {
            C()->n_sel = 0;
            return;
        }
#pragma unroll 1
        for (int a = 0, b = m - 1; a < b; a++, b--) {
            short t = seg[a];
            seg[a] = seg[b];
            seg[b] = t;
        }
        int nd = sys.step(kd, 1);
        bool ended = false;
#pragma unroll 1
        while (!ended && (cyc ? nd != kd : nd >= 0)) {
            bool bnd = s->dom[nd].state == ST_BOUND;
            if (bnd) {
                started = true;
                if (m < K::D) seg[m++] = (short)nd;
            }
            else if (started) {
                ended = true;
            }
            nd = sys.step(nd, 1);
        }
        if (!cyc) {
            nd = sys.step(kd, -1);
            if (!started) {
#pragma unroll 1
                while (!ended && nd >= 0) {
                    bool bnd = s->dom[nd].state == ST_BOUND;
                    if (bnd) {
                        started = true;
                        seg[m++] = (short)nd;
                    }
                    else if (started) {
                        ended = true;
                    }
                    nd = sys.step(nd, -1);
                }
#pragma unroll 1
                for (int a = 0, b = m - 1; a < b; a++, b--) {
                    short t = seg[a];
                    seg[a] = seg[b];
                    seg[b] = t;
                }
            }
        }
        C()->n_sel = m;
    }
    // ClusteredLinkerRegrowthMCMovetype::select_and_setup_segments (:607-669)
    LDO_HDN void lk_select_and_setup_clustered(const MoveDef& md) {
        SysState<K>* s = sys.S();
        bool externally_bound = true;
        int attempts = 0;
#pragma unroll 1
        while (externally_bound && attempts != 10) {
            if (s->status != LDO_OK) return;
            lk_select_cluster();
            if (C()->n_sel == 0) {
                M()->rejected = 1;
                return;
            }
            externally_bound = lk_bound_externally();
            attempts++;
        }
        if (externally_bound) {
            M()->rejected = 1;
            return;
        }
        int front = C()->sel_scaf[0], back = C()->sel_scaf[C()->n_sel - 1];
        int n = 0;
        C()->lnk[0][n++] = (short)front;
        int p = sys.step(front, -1);
        if (p != back && p != sys.step(back, 1)) {
            int stop = sys.step(back, 2);
#pragma unroll 1
            while (p >= 0 && s->dom[p].state != ST_BOUND && n != md.max_linker_length && p != stop) {
                C()->lnk[0][n++] = (short)p;
                p = sys.step(p, -1);
            }
        }
        C()->n_lnk[0] = n;
        if (n == 1) {
            M()->rejected = 1;
            return;
        }
        n = 0;
        C()->lnk[1][n++] = (short)back;
        int nd = sys.step(back, 1);
        if (nd != C()->lnk[0][C()->n_lnk[0] - 1]) {
            int stop = sys.step(C()->lnk[0][C()->n_lnk[0] - 1], -1);
#pragma unroll 1
            while (nd >= 0 && s->dom[nd].state != ST_BOUND && n != md.max_linker_length && nd != stop) {
                C()->lnk[1][n++] = (short)nd;
                nd = sys.step(nd, 1);
            }
        }
        C()->n_lnk[1] = n;
        lk_setup_fixed_end_biases();
    }
    // central_domains = central_segment + the domains of find_staples(central_segment) in ascending
    // unique chain index (movetypes.cpp:179-189; transform_movetypes.cpp:871-878)
    LDO_HDN void lk_find_central_domains() {
        SysState<K>* s = sys.S();
#pragma unroll 1
        for (int q = 0; q < K::C; q++) C()->cen_chain[q] = 0;
        int m = 0;
#pragma unroll 1
        for (int k = 0; k < C()->n_sel; k++) {
            int dd = C()->sel_scaf[k];
            C()->cen_dom[m++] = (short)dd;
            int b = s->bound[dd];
            if (b >= 0 && sys.chain(b) != 0) scan_for_scaffold_domain(b, C()->cen_chain);
        }
        int last_uid = -1;
#pragma unroll 1
        for (;;) {
            int best = -1, best_uid = 0x7fffffff;
#pragma unroll 1
            for (int c = 1; c < K::C; c++) {
                if (!C()->cen_chain[c]) continue;
                int uid = s->chain_uid[c];
                if (uid > last_uid && uid < best_uid) {
                    best = c;
                    best_uid = uid;
                }
            }
            if (best < 0) break;
            last_uid = best_uid;
            int base = sys.chain_base(best);
#pragma unroll 1
            for (int k = 0; k < s->chain_len[best]; k++) {
                if (m >= K::D) {
                    sys.fail(LDO_ERR_CAPACITY, 14);
                    return;
                }
                C()->cen_dom[m++] = (short)(base + k);
            }
        }
        C()->n_cen = m;
#pragma unroll 1
        for (int k = 0; k < m; k++) C()->lk_mark[C()->cen_dom[k]] |= 2;
    }
    // CBMCMovetype::unassign_domains / CTRGRegrowthMCMovetype::unassign_and_save_domains(domains) on the
    // central domains (cb_movetypes.cpp:200-213, rg_movetypes.cpp:109-123)
    LDO_HDN void lk_unassign_central() {
#pragma unroll 1
        for (int k = 0; k < C()->n_cen; k++) {
            int dd = C()->cen_dom[k];
            C()->prev[dd] = sys.S()->dom[dd];
            push_modified(dd);
            sys.unassign_domain(dd);
        }
    }
    // reset_segment (:145-154)
    LDO_HD void lk_reset_segment(int last_di) {
#pragma unroll 1
        for (int di = 0; di < last_di; di++) {
            sys.unassign_domain(C()->cen_dom[di]);
            if (M()->n_assigned > 0) M()->n_assigned--;
        }
    }
    // apply_transformation (:401-465): rotation about `center` by `turns` quarter turns around basis
    // vector `axis_i`, then translation by `disp`, of the configuration saved in m_prev_pos / m_prev_ore
    LDO_HDN double lk_apply_transformation(V3 disp, V3 center, int axis_i, int turns) {
        SysState<K>* s = sys.S();
        V3 axis = ore_vec(2 * axis_i);
        double bfactor = 1;
        double delta_e = 0;
#pragma unroll 1
        for (int di = 0; di < C()->n_cen; di++) {
            int dd = C()->cen_dom[di];
            const DomRec& r = C()->prev[dd];
            V3 pos = rotate_about(rec_pos(r), center, axis, turns) + disp;
            int o = ore_code(rotate_about(ore_vec(r.ore), v3(0, 0, 0), axis, turns));
            if (!in_coord_range(pos)) {
                sys.fail(LDO_ERR_COORD_RANGE, dd);
                return 0;
            }
            int j = sys.occupant(pos);
            if (j >= 0) {
                if (s->dom[j].state != ST_UNBOUND) {
                    lk_reset_segment(di);
                    bfactor = 0;
                    break;
                }
                bool scaffold_misbinding = sys.chain(dd) == 0 && sys.chain(j) == 0;
                bool new_binding_pair = !(C()->lk_mark[j] & 2);
                if (!scaffold_misbinding && new_binding_pair) {
                    lk_reset_segment(di);
                    bfactor = 0;
                    break;
                }
                else if (scaffold_misbinding && new_binding_pair) {
                    sys.check_domain_constraints(dd, pos, o);
                    if (s->constraints_violated) {
                        bfactor = 0;
                        s->constraints_violated = 0;
                        lk_reset_segment(di);
                        break;
                    }
                }
            }
            delta_e += sys.set_checked_domain_config(dd, pos, o);
            push_assigned(dd);
        }
        if (bfactor != 0) bfactor = exp(-delta_e);
        return bfactor;
    }
    // steps_less_than_distance (:524-545)
    LDO_HD bool lk_steps_less_than_distance() const {
        const SysState<K>* s = sys.S();
        int dist[2];
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            int ep = C()->lnk_ep[q];
            dist[q] = ep >= 0 ? abssum(rec_pos(s->dom[C()->lnk[q][0]]) - rec_pos(s->dom[ep])) : 0;
        }
        return !(dist[0] <= C()->n_lnk[0] && dist[1] <= C()->n_lnk[1]);
    }
    // One trial transformation (:329-366, repeated verbatim at :474-506): draws centre, axis, turns and
    // displacement, applies it and takes it back; returns its Boltzmann factor, 0 when it cannot be used
    LDO_HDN double lk_trial_transformation(const MoveDef& md, V3& center, int& axis_i, int& turns, V3& disp) {
        int center_di = uniform_int(0, C()->n_sel - 1);
        center = rec_pos(C()->prev[C()->sel_scaf[center_di]]);
        axis_i = uniform_int(0, 2);
        turns = uniform_int(0, md.max_turns);
        int dx = uniform_int(-md.max_disp, md.max_disp);
        int dy = uniform_int(-md.max_disp, md.max_disp);
        int dz = uniform_int(-md.max_disp, md.max_disp);
        disp = v3(dx, dy, dz);
        double bfactor = lk_apply_transformation(disp, center, axis_i, turns);
        if (bfactor != 0) {
            if (lk_steps_less_than_distance()) bfactor = 0;
            lk_reset_segment(C()->n_cen);
        }
        return bfactor;
    }
    // transform_segment (:316-399): modified Rosenbluth choice among m_k trial transformations
    LDO_HDN double lk_transform_segment(const MoveDef& md) {
        C()->n_tf = 0;
#pragma unroll 1
        for (int k_i = 0; k_i != md.num_transforms; k_i++) {
            if (sys.S()->status != LDO_OK) return 0;
            V3 center, disp;
            int axis_i, turns;
            // a transformation that fits but leaves too few linker steps is dropped (:358-366) ...
            double applied = lk_trial_transformation(md, center, axis_i, turns, disp);
            if (applied != 0) {
                int t = C()->n_tf++;
                C()->tf_center[t] = center.k;
                C()->tf_disp[t] = disp.k;
                C()->tf_axis[t] = (int8_t)axis_i;
                C()->tf_turns[t] = (int8_t)turns;
                C()->tf_bfactor[t] = applied;
            }
        }
        double bias = 0;
#pragma unroll 1
        for (int t = 0; t < C()->n_tf; t++) bias += C()->tf_bfactor[t];
        if (bias == 0) {
            M()->rejected = 1;
        }
        else {
            double random_real = bias * uniform_real();
            double cum = 0;
            int sel = 0;
#pragma unroll 1
            for (;;) {
                cum += C()->tf_bfactor[sel];
                if (random_real < cum || sel == C()->n_tf - 1) break; // the reference runs off the end here
                sel++;
            }
            V3 center, disp;
            center.k = C()->tf_center[sel];
            disp.k = C()->tf_disp[sel];
            lk_apply_transformation(disp, center, C()->tf_axis[sel], C()->tf_turns[sel]);
#ifndef LDO_NO_LINKER_TRACKERS
            if (TRK()) trk_lk(4, vx(disp) + vy(disp) + vz(disp), C()->tf_turns[sel]);
#endif
        }
        return bias;
    }
    // revert_transformation (:467-522): m_k - 1 further trial transformations (of the configuration
    // m_prev_pos holds at that point), then the old configuration, whose domains enter one by one (sic)
    LDO_HDN double lk_revert_transformation(const MoveDef& md) {
        double bias = 0;
#pragma unroll 1
        for (int k_i = 0; k_i != md.num_transforms - 1; k_i++) {
            if (sys.S()->status != LDO_OK) return 1;
            V3 center, disp;
            int axis_i, turns;
            bias += lk_trial_transformation(md, center, axis_i, turns, disp);
        }
#pragma unroll 1
        for (int k = 0; k < C()->n_cen; k++) {
            int dd = C()->cen_dom[k];
            const DomRec& r = C()->oldc[dd];
            double de = sys.set_checked_domain_config(dd, rec_pos(r), r.ore);
            bias += exp(-de);
            push_assigned(dd);
        }
        return bias;
    }
    // CTCBLinkerRegrowthMCMovetype::internal_attempt_move (:859-938); the clustered variant differs in
    // the selection only
    LDO_HDN bool move_ctcb_linker(const MoveDef& md, bool clustered) {
        cp_reset();
        if (clustered) lk_select_and_setup_clustered(md);
        else lk_select_and_setup(md);
        if (M()->rejected || sys.S()->status != LDO_OK) return false;
        lk_find_central_domains();
        trk_lk_segments();
        DD bias = dd_from(1.0), new_bias = dd_from(1.0);
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            bool regrow_old = pass == 1;
            if (regrow_old) {
                update_move_params();
                bias = dd_mul(bias, exp(-calc_move_bias()));
                ctcb_setup_regrow_old(bias, new_bias);
            }
            ctcb_unassign(C()->lnk[0] + 1, C()->n_lnk[0] - 1, false);
            ctcb_unassign(C()->lnk[1] + 1, C()->n_lnk[1] - 1, true);
            lk_unassign_central();
            bias = dd_mul(bias, regrow_old ? lk_revert_transformation(md) : lk_transform_segment(md));
            if (sys.S()->status != LDO_OK) return false;
            if (M()->rejected && !regrow_old) return false;
#pragma unroll 1
            for (int q = 0; q < 2; q++) {
                ctcb_grow_list(C()->lnk[q], C()->n_lnk[q], regrow_old, bias);
                if (M()->rejected && !regrow_old) return false;
            }
        }
        return ctcb_finish(new_bias, bias);
    }
    // CTRGLinkerRegrowthMCMovetype::internal_attempt_move (:1135-1207)
    LDO_HDN bool move_ctrg_linker(const MoveDef& md) {
        rg_reset(md);
        lk_select_and_setup(md);
        if (M()->rejected || sys.S()->status != LDO_OK) return false;
        lk_find_central_domains();
        trk_lk_segments();
        if (M()->n_regrow < 2) {
            // both linkers empty: the reference indexes m_regrow_ds[1] past its end here
            sys.fail(LDO_ERR_INTERNAL, 3);
            return false;
        }
        W()->delta_e += rg_unassign_and_save_domains();
        lk_unassign_central();
        W()->weight *= lk_transform_segment(md);
        if (M()->rejected || sys.S()->status != LDO_OK) return false;
        W()->delta_e += rg_recoil_regrow();
        if (M()->rejected) return false;
        update_move_params();
        W()->delta_e += calc_move_bias();

        // new-configuration weights
        rg_copy_queues_to_wq();
#pragma unroll 1
        for (int k = 0; k < K::D; k++) C()->oldc[k] = C()->prev[k];
        M()->n_modified = 0;
        rg_unassign_and_save_domains();
        cp_reset_active_endpoints();
        rg_calc_weights();

        // old-configuration weights
        rg_unassign_domains();
        lk_unassign_central();
        W()->weight /= lk_revert_transformation(md);
        cp_reset_active_endpoints();
        rg_calc_old_c_opens();
        W()->weight_new = W()->weight;
        W()->weight = 1;
        rg_copy_queues_to_wq();
#pragma unroll 1
        for (int k = 0; k < K::D; k++) C()->newc[k] = C()->prev[k];
        M()->n_modified = 0;
        rg_unassign_and_save_domains();
#pragma unroll 1
        for (int k = 0; k < C()->n_cen; k++) push_modified(C()->cen_dom[k]);
        cp_reset_active_endpoints();
        rg_calc_weights();

        double ratio = W()->weight_new / W()->weight * exp(-W()->delta_e);
        if (test_acceptance(ratio)) {
#pragma unroll 1
            for (int k = 0; k < K::D; k++) C()->prev[k] = C()->newc[k];
            reset_origami();
            return true;
        }
        M()->n_modified = 0;
        M()->n_assigned = 0;
        return false;
    }

    // ---- one Monte Carlo step (simulation.cpp:568-596, 655-665) ----
    LDO_HD int select_movetype() {
        double prob = uniform_real();
        int i;
#pragma unroll 1
        for (i = 0; i < MS().n; i++) {
            if (prob < MS().mt[i].cum_prob) break;
        }
        if (i >= MS().n) i = MS().n - 1; // reference reads out of bounds here; freqs sum to 1
        return i;
    }
    LDO_HD bool attempt(int i) {
        const MoveDef& md = MS().mt[i];
        reset_internal();
        STATS()->attempts[i]++;
        if (TRK()) {
            W()->trk_a = TRK()->sticky_a[i];
            W()->trk_b = TRK()->sticky_b[i];
        }
        bool accepted = false;
        switch (md.type) {
        case MT_ORIENTATION_ROTATION: accepted = move_orientation_rotation(); break;
        case MT_MET_STAPLE_EXCHANGE: accepted = move_staple_exchange(md); break;
        case MT_MET_STAPLE_REGROWTH: accepted = move_met_staple_regrowth(); break;
        case MT_CB_STAPLE_REGROWTH: accepted = move_cb_staple_regrowth(); break;
        case MT_CTCB_SCAFFOLD_REGROWTH: accepted = move_ctcb_scaffold(md); break;
        case MT_CTCB_JUMP_SCAFFOLD_REGROWTH: accepted = move_ctcb_jump_scaffold(md); break;
        case MT_CTRG_SCAFFOLD_REGROWTH: accepted = move_ctrg_scaffold(md); break;
        case MT_CTRG_JUMP_SCAFFOLD_REGROWTH: accepted = move_ctrg_jump_scaffold(md); break;
        case MT_CTCB_LINKER_REGROWTH: accepted = move_ctcb_linker(md, false); break;
        case MT_CTCB_CLUSTERED_LINKER_REGROWTH: accepted = move_ctcb_linker(md, true); break;
        case MT_CTRG_LINKER_REGROWTH: accepted = move_ctrg_linker(md); break;
        default: sys.fail(LDO_ERR_INTERNAL, 100 + md.type); break;
        }
        if (sys.S()->status != LDO_OK) return false;
        STATS()->accepts[i] += accepted ? 1 : 0;
        if (TRK()) track(i, md.type, accepted);
        return accepted;
    }
    // add_tracker (movetypes.hpp:327-339) with the fields the move left in Work::trk_a / trk_b
    LDO_HDN void track(int i, int type, bool accepted) {
        TrackStats* t = TRK();
        int a = W()->trk_a, b = W()->trk_b;
        int fa = -1, fb = -1, va = 0, vb = 0;
        switch (type) {
        case MT_MET_STAPLE_EXCHANGE: // a = staple type, b = 1 for a deletion
        case MT_MET_STAPLE_REGROWTH: // a = staple type, b = no_staples
        case MT_CB_STAPLE_REGROWTH:
            fa = b ? 1 : 0;
            va = a;
            break;
        case MT_CTRG_SCAFFOLD_REGROWTH:
        case MT_CTRG_JUMP_SCAFFOLD_REGROWTH: // a = number of scaffold domains
            fa = 0;
            va = a;
            break;
        case MT_CTCB_SCAFFOLD_REGROWTH:
        case MT_CTCB_JUMP_SCAFFOLD_REGROWTH: // a = number of scaffold domains, b = number of staples
            fa = 0;
            va = a;
            fb = 1;
            vb = b;
            break;
#ifndef LDO_NO_LINKER_TRACKERS // A/B knob (profiles/ab_r2.txt): the linker movetypes' trackers compiled out
        case MT_CTCB_LINKER_REGROWTH:
        case MT_CTCB_CLUSTERED_LINKER_REGROWTH:
        case MT_CTRG_LINKER_REGROWTH: // the six fields the move set (trk_lk), the others as its last move left them
            if (LDO_LANE == 0) {
                int* f = t->lk_sticky[i];
#pragma unroll 1
                for (int k = 0; k < 6; k++)
                    if (t->lk_set >> k & 1) f[k] = t->lk_now[k];
                t->lk_set = 0;
#pragma unroll 1
                for (int tb = 0; tb < 3; tb++) {
                    unsigned key = tb == 0 ? trk_lk_key(i, 0, f[0], f[2]) : (tb == 1 ? trk_lk_key(i, 1, f[1], f[3]) : trk_lk_key(i, 2, f[4], f[5]));
                    int e = 0;
#pragma unroll 1
                    while (e < t->lk_n && t->lk_key[e] != key) e++;
                    if (e == t->lk_n) {
                        if (e == LDO_TRK_LK_CAP) {
                            t->lk_dropped++;
                            continue;
                        }
                        t->lk_key[e] = key;
                        t->lk_cnt[e][0] = t->lk_cnt[e][1] = 0;
                        t->lk_n = e + 1;
                    }
                    t->lk_cnt[e][0] += 1;
                    t->lk_cnt[e][1] += accepted ? 1 : 0;
                }
            }
            break;
#endif
        default: break;
        }
        if (LDO_LANE == 0) {
            t->sticky_a[i] = a;
            t->sticky_b[i] = b;
            if (fa >= 0) {
                va = va < 0 ? 0 : (va >= LDO_TRK_BINS ? LDO_TRK_BINS - 1 : va);
                t->cnt[i][fa][va][0] += 1;
                t->cnt[i][fa][va][1] += accepted ? 1 : 0;
            }
            if (fb >= 0) {
                vb = vb < 0 ? 0 : (vb >= LDO_TRK_BINS ? LDO_TRK_BINS - 1 : vb);
                t->cnt[i][fb][vb][0] += 1;
                t->cnt[i][fb][vb][1] += accepted ? 1 : 0;
            }
        }
        LDO_SYNCWARP();
    }
    // m_tracker fields of the linker movetypes (transform_movetypes.cpp:393-394, 879-882, 1155-1158)
#ifdef LDO_NO_LINKER_TRACKERS
    LDO_HD void trk_lk(int, int, int) {}
    LDO_HD void trk_lk_segments() {}
#else
    LDO_HDN void trk_lk(int field, int a, int b) {
        TrackStats* t = TRK();
        if (t && LDO_LANE == 0) {
            t->lk_now[field] = a;
            t->lk_now[field + 1] = b;
            t->lk_set |= 3u << field;
        }
    }
#endif
#ifndef LDO_NO_LINKER_TRACKERS
    LDO_HDN void trk_lk_segments() {
        if (!TRK()) return;
        int cs = 0;
#pragma unroll 1
        for (int c = 1; c < K::C; c++) cs += C()->cen_chain[c] ? 1 : 0;
        trk_lk(0, C()->n_lnk[0] + C()->n_lnk[1], num_regrowth_staples());
        trk_lk(2, C()->n_sel, cs);
    }
#endif
    LDO_HDN int num_regrowth_staples() const {
        int n = 0;
#pragma unroll 1
        for (int c = 1; c < K::C; c++) n += C()->regrow_chain[c] ? 1 : 0;
        return n;
    }
    LDO_HD bool mc_step() {
        int i = select_movetype();
        W()->e_start = sys.S()->energy;
        W()->sp_start = sys.S()->num_stacked_pairs;
        bool accepted = attempt(i);
        if (sys.S()->status != LDO_OK) return false;
        if (!accepted) {
#ifndef LDO_NO_FAST_REVERT // A/B knob (profiles/ab_r2.txt)
            if (!LDO_SERIAL_DRAWS()) {
                // Production mode: the old configuration is put back without evaluating the potential and the energy
                // and stacked-pair count return to their values before the move - exactly, where the reference's
                // domain-by-domain sum (followed under replay) returns them up to rounding.
                sys.S()->weight_pass = 1;
                reset_origami();
                sys.S()->weight_pass = 0;
                sys.S()->energy = W()->e_start;
                sys.S()->num_stacked_pairs = W()->sp_start;
            }
            else
#endif
            reset_origami();
            update_move_params();
            calc_move_bias();
        }
        return accepted;
    }
};

} // namespace ldo
