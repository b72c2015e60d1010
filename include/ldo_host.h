/* ldo_host.h — C-ABI of the C++ host that keeps the reference's input/output surface for the MC
 * hot path: it reads the reference's .inp parameter file and JSON side files, builds the fp64 energy
 * tables, configures one ldo_engine (include/ldo_b200.h) and runs the reference's simulation drivers
 * on it. Replaces, at the driver level:
 *   main()                                    apps/main.cpp:19-117
 *   origami::setup_origami                    src/origami_system.cpp:991-1031
 *   GCMCSimulation constructor                src/simulation.cpp:182-265
 *   ConstantTGCMCSimulation::run              include/LatticeDNAOrigami/constant_temp_simulation.hpp:31
 *   AnnealingGCMCSimulation::run              src/annealing_simulation.cpp:38-49
 *   PTGCMCSimulation::run (1-D variants)      src/ptmc_simulation.cpp:106-150
 *   output files                              src/files.cpp:519-793, src/simulation.cpp:47-147
 */
#ifndef LDO_HOST_H
#define LDO_HOST_H

#include "ldo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldo_sim ldo_sim;

/* Error text of the last failed ldo_sim_* / ldo_host_* call on this thread. */
const char* ldo_host_last_error(void);

/* Reads `inp_path` (parser.cpp:477-479 format) and builds an engine with `n_replicas` replicas on CUDA
 * device `device`. `rank` / `n_ranks` place the engine inside a multi-GPU ensemble (0 / 1 for a single
 * GPU): independent replicas are simply numbered rank * n_replicas + r; for the replica-exchange types
 * every rank holds num_reps / n_ranks slots of every ladder, dealt in serpentine order (see ldo_exchange_pt), i.e.
 * n_replicas / (num_reps / n_ranks) ladders. Every replica starts from the configuration of the system
 * file (or restart_traj_file / restart_step). With n_ranks > 1 random_seed must be set (all ranks must
 * share the exchange stream). Returns NULL on error. */
ldo_sim* ldo_sim_create(const char* inp_path, int n_replicas, int device, int rank, int n_ranks);
void ldo_sim_destroy(ldo_sim* s);
ldo_engine* ldo_sim_engine(ldo_sim* s);

/* Multi-GPU replica exchange (replaces: the Boost.MPI communicator PTGCMCSimulation owns,
 * include/LatticeDNAOrigami/ptmc_simulation.hpp:45-49, and its slave_send / master_receive traffic,
 * src/ptmc_simulation.cpp:163-253): one ldo_sim per GPU, one host thread (or process) per ldo_sim, one NCCL
 * communicator across them. Rank 0 obtains a 128-byte unique id (ncclGetUniqueId) and hands it to every rank by
 * whatever channel the caller has (threads of one process: memory; processes: a file, MPI, torch.distributed);
 * every rank then calls ldo_sim_comm_init, which is collective. NCCL is bound at run time (libnccl.so.2) - a
 * single-GPU run never touches it. */
int ldo_comm_unique_id(void* id_out_128_bytes);
int ldo_sim_comm_init(ldo_sim* s, const void* unique_id_128_bytes);

/* Runs the driver selected by simulation_type (constant_temp, annealing, t_/ut_/hut_/st_/2d_parallel_tempering,
 * umbrella_sampling, mw_/ptmw_umbrella_sampling, enumerate) to completion, writing the reference's output files for every
 * replica (`<filebase>-<replica>.*` when there is more than one). With n_ranks > 1 (exchange types, after
 * ldo_sim_comm_init) every rank calls it; rank 0 writes the .swp file. Returns 0 or -1. */
int ldo_sim_run(ldo_sim* s);

/* One whole replica-exchange round (ptmc_simulation.cpp:113-141) enqueued on the engine's stream without a host
 * synchronisation: exchange_interval moves, collection of the exchange records, ncclAllGather over the ranks (when
 * n_ranks > 1), swap decisions on the device-resident map, energy rebuild. The host waits only when an output file is
 * due. Returns 0, 1 when max_duration was hit, -1 on error. ldo_sim_exchange_state reads the map back. */
int ldo_sim_exchange_round(ldo_sim* s, long long swap_i);

/* One replica-exchange round (ptmc_simulation.cpp:113-141): `exchange_interval` MC steps on every local
 * replica, then collection of the dependent quantities. After the caller has all-gathered them (or when
 * this engine holds every replica) ldo_sim_exchange_apply takes the swap decisions. */
int ldo_sim_exchange_advance(ldo_sim* s);
int ldo_sim_exchange_apply(ldo_sim* s, long long swap_i, const double* dependent_all);
/* m_q_to_repi of every ladder ([n_ladders][num_reps]) and the per-pair counters. */
int ldo_sim_exchange_state(ldo_sim* s, int* slot_to_replica, long long* attempts, long long* accepts);

/* simulation_type=enumerate (enumerate.cpp:19-81): ldo_sim_run enumerates the staple sets and growthpoint sets on the
 * host as the reference does and every conformation of each on the device (ldo_enumerate_conformations: all replica slots
 * of the simulation are workers), writes <filebase>.weights and prints the reference's summary. After the run:
 * out = {number of configurations (with multiplicities), average energy, average bias, conformations visited}. */
int ldo_sim_enumeration_summary(ldo_sim* s, double* out);

/* Introspection used by tests and tools */
int ldo_sim_num_temps(ldo_sim* s);
int ldo_sim_num_order_params(ldo_sim* s);
const char* ldo_sim_order_param_tag(ldo_sim* s, int i);
int ldo_sim_num_movetypes(ldo_sim* s);
const char* ldo_sim_movetype_label(ldo_sim* s, int i);
int ldo_sim_num_staple_types(ldo_sim* s);
long long ldo_sim_step(ldo_sim* s);

/* Host-side energy-table builder exposed for parity tests (nearest_neighbour.cpp, origami_potential.cpp:1057-1221).
 * out = {hyb energy, hyb enthalpy, hyb entropy} of identity pair (a, b) at table `temp_idx`; returns -1 if absent. */
int ldo_sim_pair_energies(ldo_sim* s, int temp_idx, int a, int b, double* out);
int ldo_sim_init_energies(ldo_sim* s, int temp_idx, double* out);
/* GPU-free table build for the system named by a parameter file: sizes follow *n_ident ((2n+1)^2 per
 * array; call with NULL arrays first to query it). `present` flags the tabulated identity pairs. */
int ldo_host_energy_tables(const char* inp_path, double temp, int* n_ident, double* hyb_energy, double* hyb_enthalpy,
                           double* hyb_entropy, char* present, double* init);
/* Value the parameter-file reader assigns to `key` (defaults included), as text; for parser parity tests. */
int ldo_host_inp_value(const char* inp_path, const char* key, char* out, int outlen);
int ldo_host_nn_unitless_thermo(const char* seq, double temp, double cation_M, double* out);
int ldo_host_longest_contig_complement(const char* a, const char* b, char* out, int outlen);
/* ideal_random_walk.cpp:14-73 reduced to the only property the path consumes (num_walks == 0). */
int ldo_host_no_walks(const int* start, const int* end, int steps);

#ifdef __cplusplus
}
#endif
#endif /* LDO_HOST_H */
