// ldo_host.hpp — C++ host side that keeps the reference's input surface for the MC hot path:
// the key=value .inp parameter file, the system / moveset / order-parameter / bias JSON files and
// the windows file, plus the nearest-neighbour table builder that produces the fp64 energy tables
// the device consumes. Mirrors (reference file:line):
//   InputParameters                          include/LatticeDNAOrigami/parser.hpp:17-139, src/parser.cpp:19-515
//   OrigamiInputFile                         src/files.cpp:22-127
//   OrigamiMovetypeFile                      src/files.cpp:264-320
//   OrigamiOrderParamsFile / BiasFunctions   src/files.cpp:376-470
//   nearestNeighbour::*                      src/nearest_neighbour.cpp:19-177, include/.../nearest_neighbour.hpp:21-98
//   OrigamiPotential::calc_energies          src/origami_potential.cpp:1057-1221
#pragma once

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ldo_b200.h"
#include "ldo_json.hpp"

namespace ldohost {

struct FileError: std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct NotImplemented: std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct SimulationMisuse: std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct OrigamiMisuse: std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- parameter file ---------------------------------------------------------------------------
class InputParameters {
  public:
    InputParameters(); // defaults of parser.cpp:33-451
    explicit InputParameters(std::string const& inp_filename);
    void set(std::string const& key, std::string const& value); // one key=value assignment
    void finalize(); // process_custom_types (parser.cpp:484-515)

    // System input parameters
    std::string m_origami_input_filename;
    std::string m_domain_type;
    std::string m_binding_pot;
    std::string m_misbinding_pot;
    std::string m_stacking_pot;
    std::string m_hybridization_pot;
    double m_temp;
    double m_staple_M;
    double m_cation_M;
    double m_staple_u_mult;
    bool m_constant_staple_M;
    double m_stacking_ene;
    double m_binding_h, m_binding_s, m_misbinding_h, m_misbinding_s;
    bool m_apply_mean_field_cor;
    int m_min_total_staples, m_max_total_staples, m_max_type_staples, m_max_staple_size;
    std::vector<int> m_excluded_staples;
    bool m_domain_update_biases_present;
    std::string m_ops_filename;
    std::string m_bias_funcs_filename;
    double m_bias_funcs_mult;
    std::string m_energy_filebase;
    std::string m_simulation_type;

    // General simulation parameters
    int m_random_seed;
    std::string m_movetype_filename;
    bool m_read_num_walks;
    std::string m_num_walks_filename;
    bool m_restart_from_config;
    std::string m_restart_traj_file;
    std::vector<std::string> m_restart_traj_files;
    std::string m_restart_traj_filebase;
    std::string m_restart_traj_postfix;
    bool m_restart_us_iter;
    std::string m_restart_us_filebase;
    int m_restart_step;
    std::vector<int> m_restart_steps;
    bool m_restart_from_swap;
    bool m_read_rand_engine_state;
    std::string m_rand_engine_state_file;
    std::string m_vmd_file_dir;
    int m_logging_freq, m_centering_freq, m_centering_domain, m_constraint_check_freq;
    bool m_allow_nonsensical_ps;
    double m_max_duration;

    long long m_ct_steps;
    bool m_enumerate_staples_only;
    double m_max_temp, m_min_temp, m_temp_interval;
    long long m_steps_per_temp;

    std::vector<double> m_temps;
    int m_num_reps;
    int m_exchange_interval;
    long long m_swaps;
    double m_max_pt_dur;
    std::vector<double> m_bias_mults, m_stacking_mults, m_chem_pot_mults;
    std::string m_restart_swap_file;

    std::string m_us_grid_bias_tag;
    int m_max_num_iters;
    double m_max_D_bias;
    long long m_equil_steps, m_max_equil_dur, m_iter_steps, m_iter_swaps, m_max_iter_dur;
    double m_max_rel_P_diff;
    bool m_read_biases;
    std::string m_biases_file, m_biases_filebase;
    bool m_multi_window;
    std::string m_windows_file;

    std::string m_output_filebase;
    int m_configs_output_freq, m_vtf_output_freq;
    bool m_vcf_per_domain;
    int m_counts_output_freq, m_times_output_freq, m_energies_output_freq;
    std::vector<std::string> m_ops_to_output;
    int m_order_params_output_freq, m_rand_engine_state_output_freq, m_vmd_pipe_freq;
    bool m_create_vmd_instance;

  private:
    std::map<std::string, std::string> m_raw; // list-valued options kept as strings until finalize()
};

// ---- system file --------------------------------------------------------------------------------
struct Chain { // origami_system.hpp:46-62
    int index;
    int identity;
    std::vector<int> positions; // 3 ints per domain
    std::vector<int> orientations;
};
using Chains = std::vector<Chain>;

struct OrigamiInputFile { // files.cpp:22-127
    explicit OrigamiInputFile(std::string const& filename);
    std::vector<std::vector<int>> identities;
    std::vector<std::vector<std::string>> sequences;
    std::vector<double> enthalpies, entropies;
    Chains chains;
    bool cyclic {false};
};

// .trj restart reader (files.cpp:129-218)
Chains read_trj_config(std::string const& filename, int step);

// ---- nearest-neighbour thermodynamics (nearest_neighbour.cpp) --------------------------------------
struct ThermoOfHybrid {
    double enthalpy;
    double entropy;
};
std::string calc_comp_seq(std::string const& seq);
bool seq_is_palindromic(std::string const& seq);
std::vector<std::string> find_longest_contig_complement(std::string const& seq_i, std::string const& seq_j);
ThermoOfHybrid calc_hybridization_H_and_S(std::string const& seq, double cation_M);
ThermoOfHybrid calc_unitless_hybridization_thermo(std::string const& seq, double temp, double cation_M);
double calc_unitless_hybridization_energy(std::string const& seq, double temp, double cation_M);
ThermoOfHybrid calc_unitless_init_thermo(double temp);

// Dense tables for one temperature (origami_potential.cpp:1057-1221)
struct EnergyTables {
    int n_ident {0};
    double temp {0};
    double init_energy {0}, init_enthalpy {0}, init_entropy {0};
    std::vector<double> hyb_energy, hyb_enthalpy, hyb_entropy; // (2n+1)^2, absent pairs = 0
    std::vector<char> present;
    size_t index(int a, int b) const { return static_cast<size_t>(a + n_ident) * (2 * n_ident + 1) + (b + n_ident); }
};
EnergyTables calc_energy_tables(OrigamiInputFile const& sys, InputParameters const& params, double temp);

// ---- moveset / order parameters / biases ------------------------------------------------------------
struct MovetypeSpec {
    std::string type, label;
    double freq {0};
    ldo_movetype_desc desc {};
    std::vector<double> exchange_mults;
};
std::vector<MovetypeSpec> read_movetype_file(std::string const& filename);

struct OrderParamSpec {
    std::string type, label, tag;
    int level {0};
    int staple {0};
    int chain1 {0}, domain1 {0}, chain2 {0}, domain2 {0}; // Dist / AdjacentSite
    bool update_per_domain {false}; // Dist / AdjacentSite: updated with every domain placement (order_params.cpp:505-511)
    std::vector<int> sum_ops;
};
// Returned in the reference's evaluation order (level-major, file order within a level)
std::vector<OrderParamSpec> read_order_params_file(std::string const& filename);

struct BiasSpec {
    std::string type, label, tag;
    int level {0};
    std::vector<int> ops;
    int min_op {0}, max_op {0};
    double well_bias {0}, min_bias {0}, slope {0}, outside_bias {0};
};
std::vector<BiasSpec> read_bias_functions_file(std::string const& filename, std::vector<OrderParamSpec> const& ops);

struct WindowsFile { // us_simulation.cpp:555-600
    std::string bias_tag;
    std::vector<std::vector<int>> mins, maxs;
};
WindowsFile read_windows_file(std::string const& filename);

double fraction_to_double(std::string const& s); // utility.cpp:266-281

} // namespace ldohost
