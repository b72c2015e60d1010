/* ldo_b200.h — C-ABI of the B200-native replica-batched Monte Carlo engine for the
 * LatticeDNAOrigami model (drop-in for the MC hot path of cumberworth/LatticeDNAOrigami).
 *
 * The reference has no plugin/FFI seam for this path: it is one statically linked C++ binary whose
 * internal seam is GCMCSimulation::simulate() (include/LatticeDNAOrigami/simulation.hpp:75-81,
 * src/simulation.cpp:568-653), MCMovetype::attempt_move()/reset_origami()
 * (include/LatticeDNAOrigami/movetypes.hpp:75-78) and the OrigamiSystem public methods
 * (include/LatticeDNAOrigami/origami_system.hpp:99-165). Each entry point below names the reference
 * interface it replaces. Conventions: plain C structs, caller-owned host buffers, int return codes
 * (0 = ok, negative = error; see ldo_last_error), no exceptions cross the ABI, a handle is not
 * thread-safe (one host thread per handle / GPU).
 *
 * One engine handle owns the replicas resident on one GPU; every replica is advanced by one warp.
 */
#ifndef LDO_B200_H
#define LDO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldo_engine ldo_engine;

/* ---- enumerations (values are part of the ABI) ------------------------------------------- */
enum { LDO_DOMAIN_HALFTURN = 0, LDO_DOMAIN_THREEQUARTERTURN = 1 }; /* origami_system.cpp:415-422 */
enum { LDO_MISBIND_OPPOSING = 0, LDO_MISBIND_DISALLOWED = 1 };     /* origami_potential.cpp:978-987 */

enum { /* movetype "type" strings of the moveset JSON (simulation.cpp:283-318) */
    LDO_MT_ORIENTATION_ROTATION = 0,
    LDO_MT_MET_STAPLE_EXCHANGE = 1,
    LDO_MT_MET_STAPLE_REGROWTH = 2,
    LDO_MT_CB_STAPLE_REGROWTH = 3,
    LDO_MT_CTCB_SCAFFOLD_REGROWTH = 4,
    LDO_MT_CTCB_JUMP_SCAFFOLD_REGROWTH = 5,
    LDO_MT_CTRG_SCAFFOLD_REGROWTH = 6,
    LDO_MT_CTRG_JUMP_SCAFFOLD_REGROWTH = 7,
    LDO_MT_CTCB_LINKER_REGROWTH = 8,           /* transform_movetypes.cpp:859-938 */
    LDO_MT_CTCB_CLUSTERED_LINKER_REGROWTH = 9, /* same move, ClusteredLinkerRegrowthMCMovetype selection (:607-775) */
    LDO_MT_CTRG_LINKER_REGROWTH = 10           /* transform_movetypes.cpp:1135-1207 */
};

enum { /* order parameter "type" strings (order_params.cpp:471-545) */
    LDO_OP_NUM_STAPLES = 0,
    LDO_OP_NUM_STAPLES_TYPE = 1,
    LDO_OP_STAPLE_TYPE_FULLY_BOUND = 2,
    LDO_OP_NUM_BOUND_DOMAIN_PAIRS = 3,
    LDO_OP_NUM_MISBOUND_DOMAIN_PAIRS = 4,
    LDO_OP_NUM_STACKED_PAIRS = 5,
    LDO_OP_NUM_LINEAR_HELICES = 6,
    LDO_OP_NUM_STACKED_JUNCTS = 7,
    LDO_OP_SUM = 8,
    LDO_OP_DIST = 9,          /* "Dist", update_per_domain = false, scaffold domains (order_params.cpp:34-48) */
    LDO_OP_ADJACENT_SITE = 10 /* "AdjacentSite", likewise (order_params.cpp:84-104) */
};

enum { /* bias function "type" strings (bias_functions.cpp:366-413) */
    LDO_BIAS_LINEAR_STEP_WELL = 0,
    LDO_BIAS_SQUARE_WELL = 1,
    LDO_BIAS_GRID = 2
};

enum { /* replica exchange variants (ptmc_simulation.cpp:595-680) */
    LDO_PT_T = 0,   /* t_parallel_tempering   : temperature                        */
    LDO_PT_UT = 1,  /* ut_parallel_tempering  : temperature + staple chem. pot.    */
    LDO_PT_HUT = 2, /* hut_parallel_tempering : + bias multiplier                  */
    LDO_PT_ST = 3,  /* st_parallel_tempering  : stacking multiplier                */
    LDO_PT_2D = 4   /* 2d_parallel_tempering  : temperature x stacking multiplier (ldo_exchange_pt_2d) */
};

/* ---- descriptors -------------------------------------------------------------------------- */

/* System topology and potential selectors: what origami::setup_origami (origami_system.cpp:991-1031)
 * hands to the OrigamiSystem / OrigamiPotential constructors (origami_system.cpp:41-71,
 * origami_potential.cpp:952-1013), minus sequences (those only feed the host-side table builder). */
typedef struct {
    int n_types;           /* chain identities, scaffold (identity 0) included */
    const int* type_len;   /* [n_types] domains per chain identity */
    const int* idents;     /* flattened domain identities, type-major (m_identities) */
    int cyclic;
    int domain_type;       /* LDO_DOMAIN_* */
    int misbinding_pot;    /* LDO_MISBIND_* */
    int apply_mean_field_cor;
    int max_total_staples; /* parser.cpp: max_total_staples */
    int max_type_staples;
    int max_staple_size;
    double staple_M;       /* reduced fugacity (origami_system.cpp:58) */
    double stacking_ene;   /* Constant stacking potential, kb K (origami_potential.cpp:1219-1221) */
} ldo_system_desc;

typedef struct {
    int type;            /* LDO_MT_* */
    double freq;         /* "freq" of the moveset JSON; cumulated as simulation.cpp:233-237 */
    int max_regrowth;    /* simulation.cpp:411,530 */
    int max_seg_regrowth;
    int max_num_recoils; /* simulation.cpp:528 */
    int max_c_attempts;  /* simulation.cpp:529 */
    int max_disp;          /* linker/transform moves, simulation.cpp:453-458 */
    int max_turns;
    int max_linker_length;
    int num_transforms;
    int adaptive_exchange;
    int n_exchange_mults;
    const double* exchange_mults; /* simulation.cpp:349-352 */
} ldo_movetype_desc;

typedef struct {
    int type;    /* LDO_OP_* */
    int staple;  /* "staple" option of NumStaplesType / StapleTypeFullyBound */
    int chain1, domain1, chain2, domain2; /* Dist / AdjacentSite (order_params.cpp:496-511); chains must be 0 */
    int n_sum;   /* Sum: indices of earlier order parameters */
    const int* sum_ops;
    int update_per_domain; /* Dist / AdjacentSite: "update_per_domain" (order_params.cpp:505-511): recomputed whenever one of
                              its domains is unassigned or placed (OrigamiSystemWithBias, origami_system.cpp:923-945) instead
                              of once per move */
} ldo_order_param_desc;

typedef struct {
    int type;   /* LDO_BIAS_* */
    int n_ops;  /* 1 for the wells, 1..3 for Grid */
    const int* ops; /* indices into the order-parameter list */
    int min_op, max_op;
    double well_bias, min_bias, slope, outside_bias;
} ldo_bias_desc;

/* One draw of the value-level replay tape; layout shared with the oracle. */
typedef struct {
    int kind; /* 0 = uniform_real, 1 = uniform_int */
    int lo, hi, ival;
    double real;
} ldo_tape_draw;

/* ---- lifetime ----------------------------------------------------------------------------- */

/* Replaces: OrigamiSystem construction for n_replicas independent systems on CUDA device `device`. */
int ldo_engine_create(const ldo_system_desc* desc, int n_replicas, int device, ldo_engine** out);
void ldo_engine_destroy(ldo_engine* e);
const char* ldo_last_error(const ldo_engine* e); /* pass NULL for errors of ldo_engine_create */
int ldo_num_replicas(const ldo_engine* e);

/* ---- potential tables ---------------------------------------------------------------------- */

/* Replaces: OrigamiPotential::calc_energies / update_temp table cache (origami_potential.cpp:1019-1100).
 * n_temps tables; each hyb_* array is [n_temps][(2 n_ident + 1)^2] indexed (a + n_ident)*(2 n_ident + 1) + (b + n_ident);
 * init is [n_temps][3] = init energy, enthalpy, entropy (origami_potential.cpp:1060-1063). */
int ldo_set_temperature_tables(ldo_engine* e, int n_temps, int n_ident, const double* temps,
                               const double* hyb_energy, const double* hyb_enthalpy,
                               const double* hyb_entropy, const double* init);

/* ---- moveset, order parameters, biases ----------------------------------------------------- */

/* Replaces: GCMCSimulation::construct_movetypes (simulation.cpp:267-320). */
int ldo_set_moveset(ldo_engine* e, int n, const ldo_movetype_desc* movetypes, int allow_nonsensical_ps);
/* Exchange multipliers of staple-exchange movetype `movetype` as they stand, out[R][n_staple_types]: the movetype file's
 * values, or - with adaptive_exchange - each replica's own (a multiplier that made an acceptance probability exceed one
 * is divided by ten and the move rejected, met_movetypes.cpp:228-234, 275-282; the reference keeps them in the movetype
 * object of each process, here they are per-replica device state and part of a checkpoint). */
int ldo_get_exchange_mults(ldo_engine* e, int movetype, double* out);

/* Production (Philox) mode only. on != 0: the recoil-growth moves draw in the reference's serial trial order
 * (rg_movetypes.cpp:193-198, 378-402, 435-440) instead of re-associating the draws to lanes - the branches a
 * replay tape runs, driven by Philox. Same ensemble; slower. Used to validate the lane-parallel branches
 * against the serial ones on the same device (tests/test_production_parity.py). Default off. */
int ldo_set_reference_draw_order(ldo_engine* e, int on);
/* Replaces: the choice of OrigamiSystemWithBias over OrigamiSystem by `domain_update_biases_present`
 * (origami::setup_origami, origami_system.cpp:1006-1028). Call before ldo_set_order_params. Without it order
 * parameters marked update_per_domain keep their initial values, as in the reference. */
int ldo_set_domain_update_biases(ldo_engine* e, int present);
/* Replaces: SystemOrderParams::setup_ops (order_params.cpp:471-577), both kinds. */
int ldo_set_order_params(ldo_engine* e, int n, const ldo_order_param_desc* ops);
/* Replaces: SystemBiases::setup_biases (bias_functions.cpp:334-430). Every bias is evaluated once per move: the
 * reference's dependency test never registers one as per-domain (see System::pd_update in csrc/ldo_core.cuh). */
int ldo_set_biases(ldo_engine* e, int n, const ldo_bias_desc* biases);
/* Replaces: MWUSGCMCSimulation window override of a well bias (us_simulation.cpp:503-516). */
int ldo_set_window(ldo_engine* e, int replica, int bias, int min_op, int max_op);
/* Replaces: GridBiasFunction::replace_biases (bias_functions.cpp:238-241). Dense box of
 * prod(n[k]) values, row-major; NaN marks a point that is not on the grid. */
int ldo_set_grid_bias(ldo_engine* e, int replica, int bias, const int* lo, const int* n, const double* values);
/* Per-step visit histogram over a grid bias' box (USGCMCSimulation::update_internal,
 * us_simulation.cpp:262-266). Counts accumulate until read with clear != 0. */
int ldo_get_grid_visits(ldo_engine* e, int replica, int bias, long long* counts, int clear);

/* ---- control variables, seeds, tapes -------------------------------------------------------- */

/* Replaces: OrigamiSystem::update_temp / update_staple_us / SystemBiases::update_bias_mult
 * (origami_system.cpp:618-628, ptmc_simulation.cpp:651-680). The running energy of every touched
 * replica is rebuilt (update_energy, origami_system.cpp:808-826). */
int ldo_set_control(ldo_engine* e, int first, int count, const int* temp_idx,
                    const double* staple_u_mult, const double* bias_mult, const double* stacking_mult);
int ldo_get_control(ldo_engine* e, int first, int count, int* temp_idx,
                    double* staple_u_mult, double* bias_mult, double* stacking_mult);
/* Replaces: RandomGens seeding (simulation.cpp:200-203): Philox4x32-10 key = seed, subsequence =
 * first_subsequence + replica. */
int ldo_seed(ldo_engine* e, unsigned long long seed, unsigned int first_subsequence);
/* Same with an explicit Philox subsequence per replica ([n_replicas]); used to give a replica the same
 * stream whichever GPU it lives on. */
int ldo_seed_subsequences(ldo_engine* e, unsigned long long seed, const unsigned int* subsequences);
/* Replaces: RandomEngineStateOutputFile::write / RandomEngineStateInputFile::read_state + the stream extraction of
 * simulation.cpp:204-212 (files.cpp:220-246, 781-793), which carry the mt19937_64 state as decimal text. Here the state
 * of a replica's Philox stream is ldo_rng_state_words() numbers: key (2 words), subsequence, stream, 64-bit draw counter,
 * number of buffered words, buffered words. words is [count][ldo_rng_state_words()]. Setting the state of replica 0 also
 * re-keys the engine's exchange stream with that state's key. */
int ldo_rng_state_words(void);
int ldo_get_rng_state(ldo_engine* e, int first, int count, unsigned long long* words);
int ldo_set_rng_state(ldo_engine* e, int first, int count, const unsigned long long* words);
/* Replay mode: serve the replica's draws from a tape (n = 0 detaches). */
int ldo_attach_tape(ldo_engine* e, int replica, const ldo_tape_draw* draws, long long n);
int ldo_tape_position(ldo_engine* e, int replica, long long* pos);

/* ---- configuration -------------------------------------------------------------------------- */

/* Replaces: OrigamiSystem::set_config / set_all_domains(Chains) (origami_system.cpp:327-341, 588-616).
 * Chains in the reference's wire format (struct Chain, origami_system.hpp:46-62): working order,
 * scaffold first; pos / ore are 3 ints per domain. replica = -1 sets every replica. */
int ldo_set_state(ldo_engine* e, int replica, int n_chains, const int* chain_index,
                  const int* chain_ident, const int* chain_len, const int* pos, const int* ore);
/* OrigamiSystem::set_config on a live system (origami_system.cpp:327-341): the chains of `replica` are replaced like
 * ldo_set_state does, but the stored order parameters and bias values are NOT re-evaluated - the reference's set_config
 * does not touch them, so they describe the previous configuration until the next move updates them (what the per-window
 * restart of the multi-window umbrella-sampling drivers does, us_simulation.cpp:246-250, 531-537). */
int ldo_replace_config(ldo_engine* e, int replica, int n_chains, const int* chain_index, const int* chain_ident,
                       const int* chain_len, const int* pos, const int* ore);

/* Replaces: OrigamiSystem::chains() (origami_system.cpp:173-191). Buffers sized by ldo_state_capacity. */
int ldo_state_capacity(const ldo_engine* e, int* max_chains, int* max_domains);
int ldo_get_state(ldo_engine* e, int replica, int* n_chains, int* chain_index, int* chain_ident,
                  int* chain_len, int* pos, int* ore, int* state, int* bound);

/* ---- stepping -------------------------------------------------------------------------------- */

/* Replaces: GCMCSimulation::simulate (simulation.cpp:568-653) for every replica: n_steps attempted
 * moves each, centring every centering_freq steps and check_all_constraints every
 * constraint_check_freq steps (0 = never), step counter continuing from the previous call. */
/* Sets the step counter of every replica (the `step` the centring / constraint-check frequencies refer to). The
 * annealing driver needs it: AnnealingGCMCSimulation::run adds simulate()'s return value - the LAST step number
 * plus one - to its step counter (annealing_simulation.cpp:47, simulation.cpp:574,652), so the step numbers jump
 * between temperatures. */
int ldo_set_step(ldo_engine* e, long long step);
int ldo_run(ldo_engine* e, long long n_steps, int centering_freq, int centering_domain,
            int constraint_check_freq);
/* Per-replica status: 0 ok, else the LDO_ERR_* code mirroring the reference's exception sites. */
int ldo_get_status(ldo_engine* e, int* status, int* detail);
/* Same as ldo_run but neither synchronises nor reads status back (for timing loops). */
int ldo_run_async(ldo_engine* e, long long n_steps, int centering_freq, int centering_domain,
                  int constraint_check_freq);
int ldo_synchronize(ldo_engine* e);
/* The CUDA stream every kernel of this engine is launched on (cudaStream_t as void*). */
void* ldo_stream(ldo_engine* e);

/* Build marker of the library: "cuda sm_100a" for the product. Callers that must not run on anything else
 * (the Python binding) check it. */
const char* ldo_build_info(void);
/* Number of CUDA kernels this engine has launched so far. */
long long ldo_launch_count(const ldo_engine* e);
/* Bytes of one replica's persistent state in HBM (what a run launch loads and stores once). */
unsigned long ldo_state_bytes(const ldo_engine* e);

/* ---- checkpoint / resume ------------------------------------------------------------------------ */

/* Device-side get_state / set_state for restart: `count` replicas starting at `first`
 * as `count` contiguous opaque blobs of ldo_checkpoint_size() bytes each, blob i at host + i * size
 * (configuration, occupancy table, counters, running energy, RNG counter, control variables, bias state,
 * move statistics and - when a Grid bias is configured - the replica's grid-bias values and visit
 * histogram). A blob is self-contained: any sub-range of a saved buffer can be loaded, into any replica
 * index of an engine created with the same system descriptor, moveset, order parameters and biases (after
 * ldo_exchange_windows relabelled grid ownership, load whole window ladders). Tapes are not part of a
 * checkpoint: a loaded replica has none attached. ldo_checkpoint_size() depends on whether a Grid bias is
 * configured, so query it after ldo_set_biases. The host buffer may be pinned; ldo_checkpoint_load is
 * asynchronous on the engine's stream. */
unsigned long ldo_checkpoint_size(const ldo_engine* e);
int ldo_checkpoint_save(ldo_engine* e, int first, int count, void* host);
int ldo_checkpoint_load(ldo_engine* e, int first, int count, const void* host);

/* ---- observables ------------------------------------------------------------------------------ */

/* [n_replicas][5]: total energy, hybridization enthalpy, entropy, stacking energy, external bias
 * (.ene columns, files.cpp:707-720; update_enthalpy_and_entropy origami_system.cpp:204-246). */
int ldo_get_energies(ldo_engine* e, double* out);
/* [n_replicas][9]: staples, domains, bound pairs, fully bound pairs, self-bound pairs, misbound
 * pairs, stacked pairs, unassigned domains, current unique chain index. */
int ldo_get_counters(ldo_engine* e, int* out);
/* [n_replicas][n_types-1] staples per identity (OrigamiSystem::get_staple_counts). */
int ldo_get_staple_counts(ldo_engine* e, int* out);
/* [n_replicas][n_ops] (SystemOrderParams, order_params.cpp:587-593). */
int ldo_get_order_params(ldo_engine* e, int* out);
/* [n_replicas][n_movetypes] attempts / accepts (MovetypeTracking, movetypes.hpp:49-52; .moves). */
int ldo_get_move_stats(ldo_engine* e, long long* attempts, long long* accepts);
/* Typed move trackers (movetypes.hpp:327-339, utility.hpp:100-147): the breakdown the reference's .moves summary gives
 * for each movetype - staple moves by staple type (exchange: insertions and deletions apart; regrowth: attempts with and
 * without staples in the system apart), scaffold regrowth by the number of scaffold domains (CTCB: and of staples).
 * Off by default (25 KB per replica, one extra store per move); ldo_sim_run enables them when it writes output files.
 * counts[n_movetypes][2 fields][LDO_TRACKER_BINS values][attempts, accepts]; field / value per movetype type:
 *   MetStapleExchange   field 0 insertions, 1 deletions; value = staple type
 *   Met/CBStapleRegrowth field 0 staples present, 1 no staples in the system; value = staple type (as last set)
 *   CTRG(Jump)Scaffold  field 0; value = scaffold domains selected (0 for the jump move, as in the reference)
 *   CTCB(Jump)Scaffold  field 0 as above; field 1 value = staples regrown with the segment
 * sticky[n_movetypes][2]: the tracker fields as the last move of each type left them (they persist between moves). */
#define LDO_TRACKER_BINS 64
int ldo_enable_move_trackers(ldo_engine* e, int on);
int ldo_get_move_trackers(ldo_engine* e, int replica, int* sticky, unsigned int* counts);
/* Replaces: LinkerRegrowthMCMovetype::m_tracker / m_tracking (transform_movetypes.hpp:104-105, utility.hpp:136-144) and
 * the three tables of its write_log_summary (transform_movetypes.cpp:64-139). sticky[n_movetypes][6]: linker domains,
 * linker staples, central domains, central staples, sum of the displacement, turns, as the last move of each type left
 * them. entries[n_entries][6] (room for LDO_LINKER_TRACKER_CAP): movetype, table (0 = linker / central domains,
 * 1 = linker / central staples, 2 = displacement sum / turns), the pair of values, attempts, accepts. dropped: updates
 * lost because the list was full. */
#define LDO_LINKER_TRACKER_CAP 1024
int ldo_get_linker_trackers(ldo_engine* e, int replica, int* sticky, int* n_entries, int* entries, int* dropped);

/* Diagnostics (no reference counterpart): [n_replicas][3] = start ns, end ns (device global timer) and SM id
 * of every replica's warp in the last ldo_run launch; used by profiles/ to measure load imbalance. */
int ldo_get_run_timing(ldo_engine* e, long long* out);
/* Full energy of every replica recomputed from scratch on device without touching the state
 * (the same pass check_all_constraints relies on): [n_replicas] energies, [n_replicas] stacked pairs. */
int ldo_recompute_energies(ldo_engine* e, double* energy, int* stacked_pairs);
/* Whole-state passes (origami_system.cpp:267-325, 553-571). */
int ldo_check_all_constraints(ldo_engine* e);
int ldo_center(ldo_engine* e, int centering_domain);

/* ---- replica exchange ------------------------------------------------------------------------- */

/* Replaces: PTGCMCSimulation::attempt_exchange for the 1-D variants (ptmc_simulation.cpp:360-412)
 * over `n_ladders` independent ladders of `ladder_len` control-variable slots, sharded over `n_ranks`
 * GPUs: the slots of EVERY ladder are dealt in serpentine order (ranks 0 1 .. G-1, G-1 .. 1 0, 0 1 ..): replica
 * k of ladder l lives on rank (k/G even ? k%G : G-1-k%G) at local index l*S + k/G (S = ladder_len / n_ranks), so
 * nearly every neighbour pair straddles two GPUs and the temperature-dependent cost of a move is the same on all
 * of them. Decisions
 * are taken on device from a Philox stream shared by all ranks (identical on every rank, no
 * communication); accepted swaps relabel control variables (temperature table index and multipliers),
 * configurations never move. `dependent` is the all-gathered, rank-major [n_ranks][R][4 + n_staple_types]
 * array of (enthalpy, bias, stacking, the replica's own staple chemical-potential multiplier, staple counts...)
 * as produced by ldo_exchange_collect on every rank - what slave_send ships (ptmc_simulation.cpp:163-175; the
 * reference sends m_staple_us = reduced_u * T * multiplier, here the multiplier travels and T is the slot's); pass NULL when n_ranks == 1, or after an NCCL all-gather wrote the engine's own receive buffer
 * (ldo_exchange_buffers). slot_to_replica is the reference's m_q_to_repi per ladder (the .swp row);
 * attempts / accepts are [n_ladders][ladder_len - 1]. */
/* Control-variable ladder (m_control_qs, ptmc_simulation.cpp:341-346): temperature table index and
 * multipliers of every slot; NULL multipliers mean 1. */
int ldo_set_exchange_ladder(ldo_engine* e, int ladder_len, const int* temp_idx, const double* staple_u_mult,
                            const double* bias_mult, const double* stacking_mult);
int ldo_exchange_collect(ldo_engine* e, double* dependent_local);
int ldo_exchange_pt(ldo_engine* e, int variant, long long swap_i, int n_ladders, int ladder_len,
                    int rank, int n_ranks, const double* dependent,
                    int* slot_to_replica, long long* attempts, long long* accepts);
/* Replay mode of the exchange (parity tests): the uniform reals the reference's master drew in its test_acceptance
 * calls (ptmc_simulation.cpp:255-271; none is drawn when p == 1), in order, with round_offsets[k] = index of the first
 * draw of exchange round k + 1 (round_offsets[n_rounds] = n). While a tape is set, ldo_exchange_pt / ldo_exchange_pt_2d
 * (one ladder only) serve the draws of round swap_i from it instead of Philox; n_rounds = 0 detaches. A swap
 * probability that rounds to exactly 1 in one code and to 1 - 1e-16 in the other (running energies agree to 1e-12,
 * not bitwise) makes one of them draw and not the other: ldo_exchange_tape_status counts the draws the engine wanted
 * beyond a round's supply (taken as accepted) and the draws of finished rounds it left unused. */
int ldo_set_exchange_tape(ldo_engine* e, const double* reals, long long n, const long long* round_offsets, long long n_rounds);
int ldo_exchange_tape_status(ldo_engine* e, long long* missing, long long* unused);
/* The same round without any host synchronisation, for drivers that keep the whole exchange on the engine's stream
 * (ldo_sim_exchange_round, include/ldo_host.h): ldo_exchange_collect_async enqueues the collection of the exchange
 * records into the send buffer of ldo_exchange_buffers; the caller enqueues its all-gather into the receive buffer on
 * ldo_stream() (not needed when n_ranks == 1); ldo_exchange_pt_async enqueues the swap decisions on the
 * DEVICE-RESIDENT slot -> replica map and counters (ldo_exchange_state_set uploads them once - n_slots = n_ladders *
 * ladder_len, n_counters as ldo_exchange_pt / ldo_exchange_pt_2d size attempts / accepts - and
 * ldo_exchange_state_get reads them back, synchronising) followed by the energy rebuild. v2_dim is used by LDO_PT_2D. */
int ldo_exchange_collect_async(ldo_engine* e);
int ldo_exchange_state_set(ldo_engine* e, int n_slots, int n_counters, const int* slot_to_replica, const long long* attempts,
                           const long long* accepts);
int ldo_exchange_state_get(ldo_engine* e, int n_slots, int n_counters, int* slot_to_replica, long long* attempts, long long* accepts);
int ldo_exchange_pt_async(ldo_engine* e, int variant, int v2_dim, long long swap_i, int n_ladders, int ladder_len, int rank, int n_ranks);
/* Reduced staple chemical potentials ln(staple_M) - (2 L - 1) ln 6 per staple type (m_reduced_staple_us,
 * origami_system.cpp:965-990) as the exchange uses them; returns the number of staple types. */
int ldo_get_reduced_staple_u(ldo_engine* e, double* out);
/* The swap probability the exchange kernels use, evaluated on the host (same inline function): replaces
 * PTGCMCSimulation::calc_acceptance_p (ptmc_simulation.cpp:275-313). dependent1/2 = {enthalpy, bias, stacking,
 * (ignored: the explicit staple_u_mult arguments are used), staple counts[n_staple_types]} of the two replicas; reduced_staple_u as origami_system.cpp:965-990. */
double ldo_exchange_acceptance_p(int n_staple_types, const double* reduced_staple_u, double temp1, double temp2,
                                 double staple_u_mult1, double staple_u_mult2, double stacking_mult1, double stacking_mult2,
                                 const double* dependent1, const double* dependent2);
/* Replaces: TwoDPTGCMCSimulation::attempt_exchange (ptmc_simulation.cpp:495-560): the slots of a ladder form a
 * [v1_dim temperatures][v2_dim stacking multipliers] grid, slot (i, j) = i * v2_dim + j as the reference
 * numbers its ranks (:454-471); round swap_i tests the pair set swap_i % 4 (T direction even / stacking
 * direction even / T odd / stacking odd, ptmc_simulation.hpp:158-164). The ladder set with
 * ldo_set_exchange_ladder has v1_dim * v2_dim slots. attempts / accepts are [n_ladders][2][v1_dim][v2_dim]
 * (m_attempt_count / m_swap_count, first index = direction). Sharding as ldo_exchange_pt. */
int ldo_exchange_pt_2d(ldo_engine* e, long long swap_i, int n_ladders, int v1_dim, int v2_dim,
                       int rank, int n_ranks, const double* dependent,
                       int* slot_to_replica, long long* attempts, long long* accepts);
/* Replaces: PTMWUSGCMCSimulation::attempt_exchange (us_simulation.cpp:770-864) over n_ladders independent
 * ladders of n_windows umbrella windows (replica l * n_windows + k starts in window k). Neighbouring
 * windows swap when both current grid points lie inside both windows, with probability
 * min(1, exp((b1(p1) - b2(p1)) + (b2(p2) - b1(p2)))) on the Grid bias `grid_bias`. An accepted swap
 * exchanges the window-specific state of the two replicas (limits of the `window_biases` well biases,
 * grid-bias values and visit histogram); the reference ships the two configurations instead, which is
 * the same thing relabelled. window_to_replica is m_win_to_configi per ladder (the .swp row). Ladders are
 * independent, so multi-GPU runs shard whole ladders and need no collective. */
int ldo_exchange_windows(ldo_engine* e, long long swap_i, int n_ladders, int n_windows, int grid_bias,
                         int n_window_biases, const int* window_biases, int* window_to_replica,
                         long long* attempts, long long* accepts);
/* Device pointers for the NCCL path: local send buffer / full receive buffer of the dependent
 * quantities ([n][4 + n_staple_types] doubles), so the all-gather runs device-to-device. */
int ldo_exchange_buffers(ldo_engine* e, int n_global, void** send_dev, void** recv_dev, int* doubles_per_replica);

/* ---- exact enumeration of small systems --------------------------------------------
 * One call enumerates every conformation of ONE growthpoint set of ONE staple set - what
 * ConformationalEnumerator::enumerate() does (enumerate.cpp:218-258, 378-664) - on all replica slots of the engine at once
 * (each slot is a worker that takes a share of the recursion tree; replica states are not modified). Staple sets and
 * growthpoint sets (StapleEnumerator / GrowthpointEnumerator, enumerate.cpp:813-1133) are enumerated by the caller:
 * ldo_sim_run does it for simulation_type=enumerate. Chains are numbered 0 (scaffold) and 1 + k (k-th staple of
 * staple_type[]). Returned per state (the values of the order parameters out_ops[], sorted): the sum over its leaves of
 * exp(-energy - bias) x multiplier, WITHOUT the staple-set prefix (reduced fugacity and orientation factors,
 * enumerate.cpp:264,276), which the caller applies; sums[] = {partition sum, sum of energy x weight, sum of bias x weight,
 * number of configurations} in the same convention. */
typedef struct ldo_enum_job {
    int n_staples;
    const int* staple_type;      /* [n_staples] staple identities (1-based types) */
    int n_stack;
    const int* stack_chain;      /* [n_stack] domains in the order create_domains_stack pops them (enumerate.cpp:591-637) */
    const int* stack_d;
    int n_growthpoints;          /* m_growthpoints: old domain -> new (staple) domain */
    const int* gp_old_chain;
    const int* gp_old_d;
    const int* gp_new_chain;
    const int* gp_new_d;
    int n_ident;                 /* identities run over -n_ident .. n_ident */
    const int* ident_unassigned; /* [2 n_ident + 1] m_identities_to_num_unassigned at the start, index identity + n_ident */
    int overcount;               /* 0 MaxTwoDomainOvercountCalculator, 1 MisbindingOnlyOvercountCalculator (enumerate.cpp:83-159) */
    int n_out_ops;
    const int* out_ops;          /* [n_out_ops <= 6] indices of the order parameters that label a state (ops_to_output) */
    int split_depth;             /* levels of the recursion dealt out as prefixes; <= 0: chosen by the engine */
    int staples_only;            /* StapleConformationalEnumerator (enumerate.cpp:666-811): the scaffold keeps the configuration */
    const int* scaffold_pos;     /* below ([n_scaffold][3] positions and orientation vectors), the stack holds staple domains only */
    const int* scaffold_ore;
} ldo_enum_job;
int ldo_enumerate_conformations(ldo_engine* e, const ldo_enum_job* job, int max_keys, int* n_keys, int* keys /* [max_keys][n_out_ops] */,
                                double* weights /* [max_keys] */, double* sums /* [4] */, long long* n_leaves /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif /* LDO_B200_H */
