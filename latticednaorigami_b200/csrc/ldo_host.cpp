// ldo_host.cpp — see ldo_host.hpp. Host-side input surface + energy-table builder.

#include "ldo_host.hpp"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>

namespace ldohost {

namespace {

std::string trim(std::string const& s) {
    size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}

std::string read_text_file(std::string const& filename, std::string const& what) {
    std::ifstream f {filename, std::ifstream::binary};
    if (!f) throw FileError {what + " " + filename + " does not exist"};
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

Json read_json_file(std::string const& filename, std::string const& what) {
    std::string text {read_text_file(filename, what)};
    try {
        return Json::parse(text);
    } catch (std::exception const& e) {
        throw FileError {what + " " + filename + " is not well formed:\n" + e.what()};
    }
}

bool parse_bool(std::string const& key, std::string v) {
    std::transform(v.begin(), v.end(), v.begin(), [](unsigned char c) { return std::tolower(c); });
    if (v == "true" || v == "1" || v == "yes" || v == "on") return true;
    if (v == "false" || v == "0" || v == "no" || v == "off") return false;
    throw FileError {"the argument ('" + v + "') for option '" + key + "' is invalid"};
}

template <class T>
T parse_number(std::string const& key, std::string const& v) {
    std::istringstream is {v};
    T out;
    is >> out;
    if (is.fail()) throw FileError {"the argument ('" + v + "') for option '" + key + "' is invalid"};
    return out;
}

std::vector<std::string> split_ws(std::string const& s) {
    std::istringstream is {s};
    std::vector<std::string> out;
    std::string tok;
    while (is >> tok) out.push_back(tok);
    return out;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// InputParameters (parser.cpp:19-515)
// ---------------------------------------------------------------------------------------------

InputParameters::InputParameters():
        m_origami_input_filename {},
        m_domain_type {"Halfturn"},
        m_binding_pot {"FourBody"},
        m_misbinding_pot {"Opposing"},
        m_stacking_pot {"Constant"},
        m_hybridization_pot {"NearestNeighbour"},
        m_temp {300},
        m_staple_M {1},
        m_cation_M {1},
        m_staple_u_mult {1},
        m_constant_staple_M {true},
        m_stacking_ene {1},
        m_binding_h {0},
        m_binding_s {0},
        m_misbinding_h {0},
        m_misbinding_s {0},
        m_apply_mean_field_cor {false},
        m_min_total_staples {0},
        m_max_total_staples {999},
        m_max_type_staples {999},
        m_max_staple_size {2},
        m_domain_update_biases_present {false},
        m_ops_filename {},
        m_bias_funcs_filename {},
        m_bias_funcs_mult {1},
        m_energy_filebase {},
        m_simulation_type {"constant_temp"},
        m_random_seed {-1},
        m_movetype_filename {},
        m_read_num_walks {false},
        m_num_walks_filename {},
        m_restart_from_config {false},
        m_restart_traj_file {},
        m_restart_traj_filebase {},
        m_restart_traj_postfix {".trj"},
        m_restart_us_iter {false},
        m_restart_us_filebase {},
        m_restart_step {0},
        m_restart_from_swap {false},
        m_read_rand_engine_state {false},
        m_rand_engine_state_file {},
        m_vmd_file_dir {},
        m_logging_freq {0},
        m_centering_freq {0},
        m_centering_domain {0},
        m_constraint_check_freq {0},
        m_allow_nonsensical_ps {false},
        m_max_duration {10e9},
        m_ct_steps {0},
        m_enumerate_staples_only {false},
        m_max_temp {400},
        m_min_temp {300},
        m_temp_interval {1},
        m_steps_per_temp {0},
        m_num_reps {1},
        m_exchange_interval {0},
        m_swaps {0},
        m_max_pt_dur {10e9},
        m_restart_swap_file {},
        m_us_grid_bias_tag {},
        m_max_num_iters {0},
        m_max_D_bias {0},
        m_equil_steps {0},
        m_max_equil_dur {0},
        m_iter_steps {0},
        m_iter_swaps {0},
        m_max_iter_dur {0},
        m_max_rel_P_diff {0.1},
        m_read_biases {false},
        m_biases_file {},
        m_biases_filebase {},
        m_multi_window {false},
        m_windows_file {},
        m_output_filebase {},
        m_configs_output_freq {0},
        m_vtf_output_freq {0},
        m_vcf_per_domain {false},
        m_counts_output_freq {0},
        m_times_output_freq {0},
        m_energies_output_freq {0},
        m_order_params_output_freq {0},
        m_rand_engine_state_output_freq {0},
        m_vmd_pipe_freq {0},
        m_create_vmd_instance {false} {}

InputParameters::InputParameters(std::string const& inp_filename): InputParameters() {
    std::ifstream f {inp_filename};
    if (!f) throw FileError {"Input parameter file " + inp_filename + " does not exist"};
    std::string line;
    while (std::getline(f, line)) {
        auto hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        line = trim(line);
        if (line.empty()) continue;
        auto eq = line.find('=');
        if (eq == std::string::npos) throw FileError {"the options configuration file contains an invalid line '" + line + "'"};
        set(trim(line.substr(0, eq)), trim(line.substr(eq + 1)));
    }
    finalize();
}

void InputParameters::set(std::string const& key, std::string const& v) {
    // First explicit value wins, as boost::program_options::store does
    if (m_raw.count(key)) return;
    m_raw[key] = v;
#define LDO_OPT_S(name, field) \
    if (key == name) {         \
        field = v;             \
        return;                \
    }
#define LDO_OPT_B(name, field)      \
    if (key == name) {              \
        field = parse_bool(key, v); \
        return;                     \
    }
#define LDO_OPT_N(name, field, T)        \
    if (key == name) {                   \
        field = parse_number<T>(key, v); \
        return;                          \
    }
    LDO_OPT_S("origami_input_filename", m_origami_input_filename)
    LDO_OPT_S("domain_type", m_domain_type)
    LDO_OPT_S("binding_pot", m_binding_pot)
    LDO_OPT_S("misbinding_pot", m_misbinding_pot)
    LDO_OPT_S("stacking_pot", m_stacking_pot)
    LDO_OPT_S("hybridization_pot", m_hybridization_pot)
    LDO_OPT_B("apply_mean_field_cor", m_apply_mean_field_cor)
    LDO_OPT_N("temp", m_temp, double)
    LDO_OPT_N("staple_M", m_staple_M, double)
    LDO_OPT_N("cation_M", m_cation_M, double)
    LDO_OPT_N("staple_u_mult", m_staple_u_mult, double)
    LDO_OPT_N("stacking_ene", m_stacking_ene, double)
    LDO_OPT_N("binding_h", m_binding_h, double)
    LDO_OPT_N("binding_s", m_binding_s, double)
    LDO_OPT_N("misbinding_h", m_misbinding_h, double)
    LDO_OPT_N("misbinding_s", m_misbinding_s, double)
    LDO_OPT_N("min_total_staples", m_min_total_staples, int)
    LDO_OPT_N("max_total_staples", m_max_total_staples, int)
    LDO_OPT_N("max_type_staples", m_max_type_staples, int)
    LDO_OPT_N("max_staple_size", m_max_staple_size, int)
    LDO_OPT_B("domain_update_biases_present", m_domain_update_biases_present)
    LDO_OPT_S("order_parameter_file", m_ops_filename)
    LDO_OPT_S("bias_functions_file", m_bias_funcs_filename)
    LDO_OPT_N("bias_functions_mult", m_bias_funcs_mult, double)
    LDO_OPT_S("energy_filebase", m_energy_filebase)
    LDO_OPT_S("simulation_type", m_simulation_type)
    LDO_OPT_B("enumerate_staples_only", m_enumerate_staples_only)
    LDO_OPT_N("random_seed", m_random_seed, int)
    LDO_OPT_S("movetype_file", m_movetype_filename)
    LDO_OPT_B("read_num_walks", m_read_num_walks)
    LDO_OPT_S("num_walks_filename", m_num_walks_filename)
    LDO_OPT_B("restart_from_config", m_restart_from_config)
    LDO_OPT_S("restart_traj_file", m_restart_traj_file)
    LDO_OPT_S("restart_traj_filebase", m_restart_traj_filebase)
    LDO_OPT_S("restart_traj_postfix", m_restart_traj_postfix)
    LDO_OPT_B("restart_from_swap", m_restart_from_swap)
    LDO_OPT_B("restart_us_iter", m_restart_us_iter)
    LDO_OPT_S("restart_us_filebase", m_restart_us_filebase)
    LDO_OPT_N("restart_step", m_restart_step, int)
    LDO_OPT_B("read_rand_engine_state", m_read_rand_engine_state)
    LDO_OPT_S("rand_engine_state_file", m_rand_engine_state_file)
    LDO_OPT_S("vmd_file_dir", m_vmd_file_dir)
    LDO_OPT_N("centering_freq", m_centering_freq, int)
    LDO_OPT_N("centering_domain", m_centering_domain, int)
    LDO_OPT_N("constraint_check_freq", m_constraint_check_freq, int)
    LDO_OPT_B("allow_nonsensical_ps", m_allow_nonsensical_ps)
    LDO_OPT_N("max_duration", m_max_duration, double)
    LDO_OPT_N("ct_steps", m_ct_steps, long long)
    LDO_OPT_B("constant_staple_M", m_constant_staple_M)
    LDO_OPT_N("max_temp", m_max_temp, double)
    LDO_OPT_N("min_temp", m_min_temp, double)
    LDO_OPT_N("temp_interval", m_temp_interval, double)
    LDO_OPT_N("steps_per_temp", m_steps_per_temp, long long)
    LDO_OPT_N("num_reps", m_num_reps, int)
    LDO_OPT_N("swaps", m_swaps, long long)
    LDO_OPT_N("max_pt_dur", m_max_pt_dur, double)
    LDO_OPT_N("exchange_interval", m_exchange_interval, int)
    LDO_OPT_S("restart_swap_file", m_restart_swap_file)
    LDO_OPT_S("us_grid_bias_tag", m_us_grid_bias_tag)
    LDO_OPT_N("max_num_iters", m_max_num_iters, int)
    LDO_OPT_N("max_D_bias", m_max_D_bias, double)
    LDO_OPT_N("equil_steps", m_equil_steps, long long)
    LDO_OPT_N("max_equil_dur", m_max_equil_dur, long long)
    LDO_OPT_N("iter_steps", m_iter_steps, long long)
    LDO_OPT_N("iter_swaps", m_iter_swaps, long long)
    LDO_OPT_N("max_iter_dur", m_max_iter_dur, long long)
    LDO_OPT_N("max_rel_P_diff", m_max_rel_P_diff, double)
    LDO_OPT_B("read_biases", m_read_biases)
    LDO_OPT_S("biases_file", m_biases_file)
    LDO_OPT_S("biases_filebase", m_biases_filebase)
    LDO_OPT_B("multi_window", m_multi_window)
    LDO_OPT_S("windows_file", m_windows_file)
    LDO_OPT_S("output_filebase", m_output_filebase)
    LDO_OPT_N("logging_freq", m_logging_freq, int)
    LDO_OPT_N("configs_output_freq", m_configs_output_freq, int)
    LDO_OPT_N("vtf_output_freq", m_vtf_output_freq, int)
    LDO_OPT_B("vcf_per_domain", m_vcf_per_domain)
    LDO_OPT_N("counts_output_freq", m_counts_output_freq, int)
    LDO_OPT_N("times_output_freq", m_times_output_freq, int)
    LDO_OPT_N("energies_output_freq", m_energies_output_freq, int)
    LDO_OPT_N("order_params_output_freq", m_order_params_output_freq, int)
    LDO_OPT_N("rand_engine_state_output_freq", m_rand_engine_state_output_freq, int)
    LDO_OPT_N("vmd_pipe_freq", m_vmd_pipe_freq, int)
    LDO_OPT_B("create_vmd_instance", m_create_vmd_instance)
#undef LDO_OPT_S
#undef LDO_OPT_B
#undef LDO_OPT_N
    // list-valued options are post-processed in finalize()
    static const char* const list_options[] {
            "excluded_staples", "restart_traj_files", "restart_steps", "temps", "chem_pot_mults",
            "bias_mults", "stacking_mults", "ops_to_output"};
    for (auto name: list_options)
        if (key == name) return;
    m_raw.erase(key);
    throw FileError {"unrecognised option '" + key + "'"};
}

void InputParameters::finalize() {
    auto doubles = [&](std::string const& key, std::vector<double>& out) {
        auto it = m_raw.find(key);
        if (it == m_raw.end()) return;
        out.clear();
        for (auto const& tok: split_ws(it->second)) out.push_back(std::stod(tok));
    };
    doubles("temps", m_temps);
    doubles("chem_pot_mults", m_chem_pot_mults);
    doubles("bias_mults", m_bias_mults);
    doubles("stacking_mults", m_stacking_mults);
    if (m_raw.count("excluded_staples")) {
        m_excluded_staples.clear();
        for (auto const& tok: split_ws(m_raw["excluded_staples"])) m_excluded_staples.push_back(std::stoi(tok));
    }
    if (m_raw.count("restart_traj_files")) m_restart_traj_files = split_ws(m_raw["restart_traj_files"]);
    if (m_raw.count("restart_steps")) {
        m_restart_steps.clear();
        for (auto const& tok: split_ws(m_raw["restart_steps"])) m_restart_steps.push_back(std::stoi(tok));
    }
    if (m_raw.count("ops_to_output") && m_raw["ops_to_output"] != "") m_ops_to_output = split_ws(m_raw["ops_to_output"]);
}

// ---------------------------------------------------------------------------------------------
// System file (files.cpp:36-115)
// ---------------------------------------------------------------------------------------------

OrigamiInputFile::OrigamiInputFile(std::string const& filename) {
    Json root {read_json_file(filename, "Origami input file")};
    Json const& o = root["origami"];
    if (o.has("sequences")) {
        for (size_t i {0}; i != o["sequences"].size(); i++) {
            sequences.push_back({});
            for (size_t j {0}; j != o["sequences"][i].size(); j++) sequences[i].push_back(o["sequences"][i][j].as_string());
        }
    }
    if (o.has("enthalpies")) {
        for (size_t i {0}; i != o["enthalpies"].size(); i++) enthalpies.push_back(o["enthalpies"][i].as_double());
    }
    if (o.has("entropies")) {
        for (size_t i {0}; i != o["entropies"].size(); i++) entropies.push_back(o["entropies"][i].as_double());
    }
    for (size_t i {0}; i != o["identities"].size(); i++) {
        identities.push_back({});
        for (size_t j {0}; j != o["identities"][i].size(); j++) identities[i].push_back(o["identities"][i][j].as_int());
    }
    Json const& cfg = o["configurations"][0]["chains"];
    for (size_t i {0}; i != cfg.size(); i++) {
        Chain c {};
        c.index = cfg[i]["index"].as_int();
        c.identity = cfg[i]["identity"].as_int();
        for (size_t j {0}; j != cfg[i]["positions"].size(); j++) {
            for (size_t k {0}; k != 3; k++) {
                c.positions.push_back(cfg[i]["positions"][j][k].as_int());
                c.orientations.push_back(cfg[i]["orientations"][j][k].as_int());
            }
        }
        chains.push_back(c);
    }
    cyclic = o["cyclic"].as_bool();
    if (identities.empty()) throw FileError {"Origami input file " + filename + " has no identities"};
}

// .trj reader (files.cpp:129-218): step records separated by blank lines
Chains read_trj_config(std::string const& filename, int step) {
    std::ifstream f {filename};
    if (!f) throw FileError {"Trajectory input file " + filename + " does not exist"};
    std::string line;
    for (int i {0}; i != step; i++) {
        for (;;) {
            if (!std::getline(f, line)) {
                throw FileError {"Step " + std::to_string(step) + " not found in trajectory input file" + filename};
            }
            if (line.empty()) break;
        }
    }
    std::getline(f, line); // step number
    Chains chains {};
    for (;;) {
        std::string ident_line;
        if (!std::getline(f, ident_line) || trim(ident_line).empty()) break;
        std::istringstream ils {ident_line};
        Chain c {};
        ils >> c.index >> c.identity;
        std::string pos_line, ore_line;
        std::getline(f, pos_line);
        std::getline(f, ore_line);
        std::istringstream pls {pos_line}, ols {ore_line};
        int v;
        while (pls >> v) c.positions.push_back(v);
        while (ols >> v) c.orientations.push_back(v);
        chains.push_back(c);
    }
    return chains;
}

// ---------------------------------------------------------------------------------------------
// Nearest-neighbour model (nearest_neighbour.hpp:21-98, nearest_neighbour.cpp:19-177)
// ---------------------------------------------------------------------------------------------

namespace {
const double R_gas {8.3144598}; // J/K/mol
const double J_Per_Cal {4.184};

struct NNEntry {
    const char* key;
    double enthalpy; // kcal/mol (SantaLucia 2004)
    double entropy; // kcal/mol/K
};
const NNEntry NN_TABLE[] {
        {"AA/TT", -7.6, -0.0213}, {"TT/AA", -7.6, -0.0213}, {"AT/TA", -7.2, -0.0204}, {"TA/AT", -7.2, -0.0213},
        {"CA/GT", -8.5, -0.0227}, {"TG/AC", -8.5, -0.0227}, {"GT/CA", -8.4, -0.0224}, {"AC/TG", -8.4, -0.0224},
        {"CT/GA", -7.8, -0.0210}, {"AG/TC", -7.8, -0.0210}, {"GA/CT", -8.2, -0.0222}, {"TC/AG", -8.2, -0.0222},
        {"CG/GC", -10.6, -0.0272}, {"GC/CG", -9.8, -0.0244}, {"GG/CC", -8.0, -0.0199}, {"CC/GG", -8.0, -0.0199}};
const double NN_INIT_H {0.2}, NN_INIT_S {-0.0057};
const double NN_TERMINAL_AT_H {2.2}, NN_TERMINAL_AT_S {0.0069};
const double NN_SYMMETRY_S {-0.0014};

NNEntry const& nn_lookup(std::string const& key) {
    for (auto const& e: NN_TABLE)
        if (key == e.key) return e;
    throw OrigamiMisuse {"no nearest-neighbour parameters for " + key};
}

char comp_base(char b) {
    switch (b) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'G': return 'C';
    case 'C': return 'G';
    }
    throw OrigamiMisuse {std::string("not a DNA base: ") + b};
}
} // namespace

std::string calc_comp_seq(std::string const& seq) {
    std::string out;
    for (char b: seq) out.push_back(comp_base(b));
    return out;
}

bool seq_is_palindromic(std::string const& seq) {
    std::string rc {calc_comp_seq(seq)};
    std::reverse(rc.begin(), rc.end());
    return rc == seq;
}

std::vector<std::string> find_longest_contig_complement(std::string const& seq_i, std::string const& seq_j) {
    std::string seq_three, seq_five;
    if (seq_i.size() <= seq_j.size()) {
        seq_three = seq_j;
        seq_five = seq_i;
    }
    else {
        seq_three = seq_i;
        seq_five = seq_j;
    }
    std::reverse(seq_three.begin(), seq_three.end());
    seq_three = calc_comp_seq(seq_three);
    std::vector<std::string> comp_seqs {};
    for (size_t len {seq_three.size()}; len != 0; len--) {
        for (size_t start {0}; start != seq_three.size() - len + 1; start++) {
            std::string sub {seq_three.substr(start, len)};
            size_t at {seq_five.find(sub)};
            while (at != std::string::npos) {
                comp_seqs.push_back(sub);
                at = seq_five.find(sub, at + 1);
            }
        }
        if (!comp_seqs.empty()) return comp_seqs;
    }
    return comp_seqs;
}

ThermoOfHybrid calc_hybridization_H_and_S(std::string const& seq, double cation_M) {
    std::string comp {calc_comp_seq(seq)};
    double DS_sym {seq_is_palindromic(seq) ? NN_SYMMETRY_S : 0};
    double DH_stack {0}, DS_stack {0};
    for (size_t i {0}; i + 1 < seq.size(); i++) {
        std::string key {seq.substr(i, 2) + "/" + comp.substr(i, 2)};
        NNEntry const& e {nn_lookup(key)};
        DH_stack += e.enthalpy;
        DS_stack += e.entropy;
    }
    int terminal_at {0};
    if (seq.front() == 'A' || seq.front() == 'T') terminal_at++;
    if (seq.back() == 'A' || seq.back() == 'T') terminal_at++;
    double DH_at {NN_TERMINAL_AT_H * terminal_at};
    double DS_at {NN_TERMINAL_AT_S * terminal_at};
    double DH {NN_INIT_H + DH_stack + DH_at};
    double DS {NN_INIT_S + DS_sym + DS_stack + DS_at};
    DS += 0.368 * seq.size() * std::log(cation_M) / 1000;
    return {DH, DS};
}

ThermoOfHybrid calc_unitless_hybridization_thermo(std::string const& seq, double temp, double cation_M) {
    ThermoOfHybrid t {calc_hybridization_H_and_S(seq, cation_M)};
    t.enthalpy = t.enthalpy * J_Per_Cal * 1000 / R_gas / temp;
    t.entropy = t.entropy * J_Per_Cal * 1000 / R_gas;
    return t;
}

double calc_unitless_hybridization_energy(std::string const& seq, double temp, double cation_M) {
    ThermoOfHybrid t {calc_unitless_hybridization_thermo(seq, temp, cation_M)};
    return t.enthalpy - t.entropy;
}

ThermoOfHybrid calc_unitless_init_thermo(double temp) {
    ThermoOfHybrid t {NN_INIT_H, NN_INIT_S};
    t.enthalpy = t.enthalpy * J_Per_Cal * 1000 / R_gas / temp;
    t.entropy = t.entropy * J_Per_Cal * 1000 / R_gas;
    return t;
}

// ---------------------------------------------------------------------------------------------
// Energy tables (origami_potential.cpp:1057-1221)
// ---------------------------------------------------------------------------------------------

EnergyTables calc_energy_tables(OrigamiInputFile const& sys, InputParameters const& params, double temp) {
    EnergyTables t {};
    t.temp = temp;
    int n {0};
    for (auto const& chain: sys.identities)
        for (int id: chain) n = std::max(n, std::abs(id));
    t.n_ident = n;
    size_t sz {static_cast<size_t>(2 * n + 1) * (2 * n + 1)};
    t.hyb_energy.assign(sz, 0);
    t.hyb_enthalpy.assign(sz, 0);
    t.hyb_entropy.assign(sz, 0);
    t.present.assign(sz, 0);
    ThermoOfHybrid init {calc_unitless_init_thermo(temp)};
    t.init_enthalpy = init.enthalpy;
    t.init_entropy = init.entropy;
    t.init_energy = init.enthalpy - init.entropy;

    std::string const& pot {params.m_hybridization_pot};
    if (pot != "NearestNeighbour" && pot != "Uniform" && pot != "Specified") {
        throw NotImplemented {pot + ": No such hybridization potential"};
    }
    if (pot == "NearestNeighbour" && sys.sequences.size() != sys.identities.size()) {
        throw FileError {"NearestNeighbour hybridization needs a sequence for every domain"};
    }
    for (size_t ci {0}; ci != sys.identities.size(); ci++) {
        for (size_t cj {0}; cj != sys.identities.size(); cj++) {
            for (size_t di {0}; di != sys.identities[ci].size(); di++) {
                int a {sys.identities[ci][di]};
                for (size_t dj {0}; dj != sys.identities[cj].size(); dj++) {
                    int b {sys.identities[cj][dj]};
                    double H {0}, S {0};
                    if (pot == "NearestNeighbour") {
                        // calc_hybridization_energy(seq_i, seq_j, key) (:1102-1162)
                        std::string const& seq_i {sys.sequences[ci][di]};
                        std::string const& seq_j {sys.sequences[cj][dj]};
                        std::vector<std::string> comps {find_longest_contig_complement(seq_i, seq_j)};
                        if (!comps.empty()) {
                            int N {0};
                            for (auto const& cs: comps) {
                                ThermoOfHybrid th {calc_unitless_hybridization_thermo(cs, temp, params.m_cation_M)};
                                H += th.enthalpy - t.init_enthalpy;
                                S += th.entropy - t.init_entropy;
                                N++;
                            }
                            H /= N;
                            S /= N;
                            S += std::log(6);
                            if (a == -b) {
                                if (params.m_apply_mean_field_cor) S += 3 * std::log(6);
                                if (comps[0].size() != seq_i.size() && seq_i.size() == seq_j.size()) {
                                    throw OrigamiMisuse {"Sequences that should be complementary are not: \n" + seq_i + "\n" + seq_j};
                                }
                            }
                            else if (comps[0].size() == seq_i.size() && seq_i.size() == seq_j.size()) {
                                throw OrigamiMisuse {"Sequences that should not be complementary are: \n" + seq_i + "\n" + seq_j};
                            }
                        }
                    }
                    else {
                        // Uniform (:1164-1184) / Specified (:1186-1208)
                        if (a == -b) {
                            if (pot == "Uniform") {
                                H = params.m_binding_h / temp;
                                S = params.m_binding_s;
                            }
                            else {
                                size_t k {static_cast<size_t>(std::abs(a) - 1)};
                                if (k >= sys.enthalpies.size() || k >= sys.entropies.size()) {
                                    throw FileError {"Specified hybridization needs enthalpies and entropies per domain pair"};
                                }
                                H = sys.enthalpies[k] / temp;
                                S = sys.entropies[k];
                            }
                            if (params.m_apply_mean_field_cor) S += 3 * std::log(6);
                        }
                        else {
                            H = params.m_misbinding_h / temp;
                            S = params.m_misbinding_s;
                        }
                        S += std::log(6);
                    }
                    size_t k {t.index(a, b)};
                    t.hyb_enthalpy[k] = H;
                    t.hyb_entropy[k] = S;
                    t.hyb_energy[k] = H - S;
                    t.present[k] = 1;
                }
            }
        }
    }
    return t;
}

// ---------------------------------------------------------------------------------------------
// Moveset / order parameters / biases / windows
// ---------------------------------------------------------------------------------------------

double fraction_to_double(std::string const& s) {
    auto slash = s.find('/');
    if (slash == std::string::npos) return std::stod(s);
    return std::stod(s.substr(0, slash)) / std::stod(s.substr(slash + 1));
}

std::vector<MovetypeSpec> read_movetype_file(std::string const& filename) {
    Json root {read_json_file(filename, "Move type file")};
    Json const& mts = root["origami"]["movetypes"];
    std::vector<MovetypeSpec> out {};
    for (size_t i {0}; i != mts.size(); i++) {
        Json const& j = mts[i];
        MovetypeSpec m {};
        m.type = j["type"].as_string();
        m.label = j["label"].as_string();
        m.freq = fraction_to_double(j["freq"].as_string());
        m.desc.freq = m.freq;
        // simulation.cpp:283-318 (type dispatch) and :340-566 (per-type options)
        if (m.type == "OrientationRotation") {
            m.desc.type = LDO_MT_ORIENTATION_ROTATION;
        }
        else if (m.type == "MetStapleExchange") {
            m.desc.type = LDO_MT_MET_STAPLE_EXCHANGE;
            for (size_t k {0}; k != j["exchange_mults"].size(); k++) m.exchange_mults.push_back(j["exchange_mults"][k].as_double());
            m.desc.adaptive_exchange = j["adaptive_exchange"].as_bool() ? 1 : 0;
        }
        else if (m.type == "MetStapleRegrowth") {
            m.desc.type = LDO_MT_MET_STAPLE_REGROWTH;
        }
        else if (m.type == "CBStapleRegrowth") {
            m.desc.type = LDO_MT_CB_STAPLE_REGROWTH;
        }
        else if (m.type == "CTCBScaffoldRegrowth" || m.type == "CTCBJumpScaffoldRegrowth") {
            m.desc.type = m.type == "CTCBScaffoldRegrowth" ? LDO_MT_CTCB_SCAFFOLD_REGROWTH : LDO_MT_CTCB_JUMP_SCAFFOLD_REGROWTH;
            m.desc.max_regrowth = j["max_regrowth"].as_int();
            m.desc.max_seg_regrowth = j["max_seg_regrowth"].as_int();
        }
        else if (m.type == "CTRGScaffoldRegrowth" || m.type == "CTRGJumpScaffoldRegrowth") {
            m.desc.type = m.type == "CTRGScaffoldRegrowth" ? LDO_MT_CTRG_SCAFFOLD_REGROWTH : LDO_MT_CTRG_JUMP_SCAFFOLD_REGROWTH;
            m.desc.max_num_recoils = j["max_num_recoils"].as_int();
            m.desc.max_c_attempts = j["max_c_attempts"].as_int();
            m.desc.max_regrowth = j["max_regrowth"].as_int();
            m.desc.max_seg_regrowth = j["max_seg_regrowth"].as_int();
        }
        else if (m.type == "CTCBLinkerRegrowth" || m.type == "CTCBClusteredLinkerRegrowth" || m.type == "CTRGLinkerRegrowth") {
            // setup_scaffold_transform_movetype (simulation.cpp:444-513)
            m.desc.type = m.type == "CTCBLinkerRegrowth" ? LDO_MT_CTCB_LINKER_REGROWTH
                    : m.type == "CTCBClusteredLinkerRegrowth" ? LDO_MT_CTCB_CLUSTERED_LINKER_REGROWTH : LDO_MT_CTRG_LINKER_REGROWTH;
            m.desc.max_disp = j["max_disp"].as_int();
            m.desc.max_turns = j["max_turns"].as_int();
            m.desc.max_regrowth = j["max_regrowth"].as_int();
            m.desc.max_linker_length = j["max_linker_length"].as_int();
            m.desc.num_transforms = j["num_transforms"].as_int();
            if (m.type == "CTRGLinkerRegrowth") {
                m.desc.max_num_recoils = j["max_num_recoils"].as_int();
                m.desc.max_c_attempts = j["max_c_attempts"].as_int();
            }
        }
        else if (m.type == "CTRGClusteredLinkerRegrowth") {
            // accepted by the reference's type dispatch (simulation.cpp:302) but never constructed
            // (setup_scaffold_transform_movetype has no case for it and returns a null movetype)
            throw NotImplemented {m.type + ": the reference does not construct this movetype either"};
        }
        else {
            throw SimulationMisuse {m.type + ": no such movetype"};
        }
        out.push_back(m);
    }
    for (auto& m: out) {
        m.desc.n_exchange_mults = static_cast<int>(m.exchange_mults.size());
        m.desc.exchange_mults = m.exchange_mults.empty() ? nullptr : m.exchange_mults.data();
    }
    return out;
}

std::vector<OrderParamSpec> read_order_params_file(std::string const& filename) {
    Json root {read_json_file(filename, "Order parameter file")};
    Json const& jops = root["origami"]["order_params"];
    std::vector<OrderParamSpec> file_order {};
    int max_level {0};
    for (size_t i {0}; i != jops.size(); i++) {
        OrderParamSpec op {};
        op.type = jops[i]["type"].as_string();
        op.label = jops[i]["label"].as_string();
        op.tag = jops[i]["tag"].as_string();
        op.level = jops[i]["level"].as_int();
        op.staple = jops[i]["staple"].as_int();
        if (op.type == "Dist" || op.type == "AdjacentSite") {
            // order_params.cpp:496-516; update_per_domain selects the per-domain kind (OrigamiSystemWithBias)
            op.update_per_domain = jops[i]["update_per_domain"].as_bool();
            op.chain1 = jops[i]["chain1"].as_int();
            op.domain1 = jops[i]["domain1"].as_int();
            op.chain2 = jops[i]["chain2"].as_int();
            op.domain2 = jops[i]["domain2"].as_int();
        }
        max_level = std::max(max_level, op.level);
        file_order.push_back(op);
    }
    // level-major order (files.cpp:406-417); Sum terms refer to tags of earlier entries
    std::vector<OrderParamSpec> out {};
    std::vector<size_t> src {};
    for (int level {0}; level != max_level + 1; level++) {
        for (size_t i {0}; i != file_order.size(); i++) {
            if (file_order[i].level == level) {
                out.push_back(file_order[i]);
                src.push_back(i);
            }
        }
    }
    // The reference's setup loop carries `update_per_domain` over from one entry of a level to the next
    // (order_params.cpp:485-511: the flag is set by Dist / AdjacentSite / Sum entries and read by all): a counter-type
    // parameter that follows a per-domain one in its level is registered nowhere and is never updated again. Refused.
    for (size_t k {0}; k != out.size(); k++) {
        bool dist_like {out[k].type == "Dist" || out[k].type == "AdjacentSite"};
        if (!dist_like && out[k].type != "Sum") {
            for (size_t q {k}; q-- > 0 && out[q].level == out[k].level;) {
                bool q_dist {out[q].type == "Dist" || out[q].type == "AdjacentSite"};
                if (q_dist || out[q].type == "Sum") {
                    if (out[q].update_per_domain) {
                        throw SimulationMisuse {out[k].tag + ": follows a per-domain order parameter in its level; the reference never updates such a parameter"};
                    }
                    break;
                }
            }
        }
    }
    for (size_t k {0}; k != out.size(); k++) {
        if (out[k].type == "Sum") {
            Json const& terms = jops[src[k]]["ops"];
            for (size_t t {0}; t != terms.size(); t++) {
                std::string tag {terms[t].as_string()};
                int found {-1};
                for (size_t q {0}; q != k; q++)
                    if (out[q].tag == tag) found = static_cast<int>(q);
                if (found < 0) throw SimulationMisuse {"Sum order parameter refers to unknown tag " + tag};
                out[k].sum_ops.push_back(found);
            }
        }
    }
    return out;
}

std::vector<BiasSpec> read_bias_functions_file(std::string const& filename, std::vector<OrderParamSpec> const& ops) {
    Json root {read_json_file(filename, "Bias functions file")};
    Json const& jb = root["origami"]["bias_functions"];
    std::vector<BiasSpec> file_order {};
    int max_level {0};
    for (size_t i {0}; i != jb.size(); i++) {
        BiasSpec b {};
        b.type = jb[i]["type"].as_string();
        b.label = jb[i]["label"].as_string();
        b.tag = jb[i]["tag"].as_string();
        b.level = jb[i]["level"].as_int();
        for (size_t k {0}; k != jb[i]["ops"].size(); k++) {
            std::string tag {jb[i]["ops"][k].as_string()};
            int found {-1};
            for (size_t q {0}; q != ops.size(); q++)
                if (ops[q].tag == tag) found = static_cast<int>(q);
            if (found < 0) throw SimulationMisuse {"bias function refers to unknown order parameter " + tag};
            b.ops.push_back(found);
        }
        b.min_op = jb[i]["min_op"].as_int();
        b.max_op = jb[i]["max_op"].as_int();
        b.well_bias = jb[i]["well_bias"].as_double();
        b.min_bias = jb[i]["min_bias"].as_double();
        b.slope = jb[i]["slope"].as_double();
        b.outside_bias = jb[i]["outside_bias"].as_double();
        if (b.type != "LinearStepWell" && b.type != "SquareWell" && b.type != "Grid") {
            // a LinearStep entry always throws in the reference as well (App. A4)
            throw SimulationMisuse {b.type + "; no such bias function type"};
        }
        max_level = std::max(max_level, b.level);
        file_order.push_back(b);
    }
    std::vector<BiasSpec> out {};
    for (int level {0}; level != max_level + 1; level++)
        for (auto const& b: file_order)
            if (b.level == level) out.push_back(b);
    return out;
}

WindowsFile read_windows_file(std::string const& filename) {
    std::ifstream f {filename};
    if (!f) throw FileError {"Windows file " + filename + " does not exist"};
    WindowsFile w {};
    std::string line;
    std::getline(f, line);
    auto tags = split_ws(line);
    if (tags.empty()) throw FileError {"Windows file " + filename + " has no bias tag"};
    w.bias_tag = tags[0];
    while (std::getline(f, line)) {
        if (trim(line).empty()) continue;
        auto comma = line.find(',');
        if (comma == std::string::npos) throw FileError {"Windows file " + filename + ": expected 'min, max'"};
        std::vector<int> lo, hi;
        for (auto const& tok: split_ws(line.substr(0, comma))) lo.push_back(std::stoi(tok));
        for (auto const& tok: split_ws(line.substr(comma + 1))) hi.push_back(std::stoi(tok));
        w.mins.push_back(lo);
        w.maxs.push_back(hi);
    }
    return w;
}

} // namespace ldohost
