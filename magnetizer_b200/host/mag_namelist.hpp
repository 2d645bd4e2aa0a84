// mag_namelist.hpp -- reader of the reference's parameter file for the dynamo path.
//
// The reference reads seven Fortran namelists from one file (read_global_parameters,
// source/global_input_parameters.f90:356-374; example: example/example_global_parameters.in).
// This parser understands the subset of the namelist syntax those files use: `&group ... /`,
// `!` comments, `key = value` with logical (.true./.false./T/F), integer, real (d/e
// exponents) and quoted-string values, and comma-separated lists that may continue on the
// following lines (output_quantities_list).  Keys are case-insensitive.  Values given in
// the file are parsed as doubles -- the float rounding of the reference's default
// literals only applies to parameters the file leaves untouched, which is
// what mag_params_default provides.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/magnetizer_b200.h"

namespace maghost {

struct Namelist {
  // group -> key -> list of raw value tokens (strings keep their content without quotes)
  std::map<std::string, std::map<std::string, std::vector<std::string>>> groups;

  static std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
  }

  static Namelist parse(const std::string& text) {
    Namelist nl;
    std::string group, key;
    size_t i = 0;
    const size_t n = text.size();
    auto skip_ws = [&]() {
      while (i < n) {
        if (text[i] == '!') { while (i < n && text[i] != '\n') i++; }
        else if (std::isspace((unsigned char)text[i]) || text[i] == ',') i++;
        else break;
      }
    };
    while (true) {
      skip_ws();
      if (i >= n) break;
      const char c = text[i];
      if (c == '&') {
        size_t j = ++i;
        while (j < n && (std::isalnum((unsigned char)text[j]) || text[j] == '_')) j++;
        group = lower(text.substr(i, j - i));
        nl.groups[group];
        key.clear();
        i = j;
      } else if (c == '/') {
        group.clear(); key.clear(); i++;
      } else if (c == '\'' || c == '"') {
        size_t j = i + 1;
        while (j < n && text[j] != c) j++;
        if (group.empty() || key.empty()) throw std::runtime_error("namelist: string value outside an assignment");
        nl.groups[group][key].push_back(text.substr(i + 1, j - i - 1));
        i = j + 1;
      } else {
        size_t j = i;
        while (j < n && !std::isspace((unsigned char)text[j]) && text[j] != ',' && text[j] != '=' && text[j] != '!' && text[j] != '/') j++;
        std::string tok = text.substr(i, j - i);
        i = j;
        size_t k = i;
        while (k < n && (text[k] == ' ' || text[k] == '\t')) k++;
        if (k < n && text[k] == '=') {  // a new key
          if (group.empty()) throw std::runtime_error("namelist: assignment outside a group: " + tok);
          key = lower(tok);
          nl.groups[group][key].clear();
          i = k + 1;
        } else {
          if (group.empty() || key.empty()) throw std::runtime_error("namelist: stray token " + tok);
          nl.groups[group][key].push_back(tok);
        }
      }
    }
    return nl;
  }
  static Namelist parse_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open parameter file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
  }

  const std::vector<std::string>* find(const std::string& key) const {
    for (const auto& g : groups) {
      auto it = g.second.find(lower(key));
      if (it != g.second.end() && !it->second.empty()) return &it->second;
    }
    return nullptr;
  }
  static double to_double(std::string s) {
    for (auto& ch : s) if (ch == 'd' || ch == 'D') ch = 'e';
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str()) throw std::runtime_error("namelist: not a number: " + s);
    return v;
  }
  static bool to_bool(const std::string& s0) {
    const std::string s = lower(s0);
    if (s == ".true." || s == "t" || s == "true" || s == ".t.") return true;
    if (s == ".false." || s == "f" || s == "false" || s == ".f.") return false;
    throw std::runtime_error("namelist: not a logical: " + s0);
  }
  bool get(const std::string& key, double* v) const { auto p = find(key); if (!p) return false; *v = to_double((*p)[0]); return true; }
  bool get(const std::string& key, int32_t* v) const {
    auto p = find(key);
    if (!p) return false;
    const std::string s = lower((*p)[0]);
    if (s.find("true") != std::string::npos || s.find("false") != std::string::npos || s == "t" || s == "f") *v = to_bool(s) ? 1 : 0;
    else *v = (int32_t)to_double(s);
    return true;
  }
  bool get(const std::string& key, std::string* v) const { auto p = find(key); if (!p) return false; *v = (*p)[0]; return true; }
};

// I/O and run-control settings the driver needs (global_input_parameters.f90:30-91)
struct RunConfig {
  mag_params_t params;
  std::string model_name = "Magnetized SAM";
  std::string input_file_name = "magnetized_galaxies_input.hdf5", output_file_name = "magnetized_galaxies_output.hdf5";
  std::vector<std::string> output_quantities_list;
  int info = 1;
  // io_parameters: dataset storage and checkpoint cadence (IO_hdf5.f90:940-981, distributor.f90:299-336)
  int p_IO_chunking = 1, p_IO_number_of_galaxies_in_chunks = 1000, p_IO_compression = 1, p_IO_compression_level = 6;
  int p_ncheckpoint = 5000;
  double p_max_walltime = -1;   // seconds; < 0: no limit (run_parameters, :85)
  std::vector<std::string> warnings;   // unknown keys, keys without effect on this path
};

// Every key of the reference's seven namelists that is NOT a member of mag_params_t, with what
// a non-default value means here.  A switch the CUDA path does not implement must keep its
// default: anything else would run default physics under a non-default label (the effective
// parameters are stamped into the output file), so it is an error, not a warning.
struct OtherKey { const char* name; const char* dflt; int kind; };   // kind 0: must equal dflt, 1: no effect on this path
static const OtherKey kOtherKeys[] = {
    // dynamo_parameters
    {"alp_ceiling", ".true.", 0}, {"alp_squared", ".false.", 0}, {"krause", ".true.", 0}, {"advect", ".true.", 0},
    {"turb_dif", ".true.", 0},
    // grid_parameters
    {"p_rescale_field_for_shrinking_disks", ".true.", 0}, {"p_rescale_field_for_expanding_disks", ".false.", 0},
    {"p_scale_back_f_array", ".true.", 0}, {"p_interp_method", "simple", 0},
    // ISM_and_disk_parameters
    {"p_check_hydro_solution", ".false.", 1}, {"p_use_fixed_turbulent_to_scaleheight_ratio", ".false.", 0},
    {"p_turbulent_to_scaleheight_ratio", "0.25", 1}, {"p_truncates_within_rreg", ".false.", 0},
    {"p_extra_pressure_outputs", ".false.", 0}, {"p_p2_workaround", ".false.", 0}, {"p_sech2_profile", ".false.", 0},
    {"p_pdm_numerical", ".false.", 0}, {"p_pbulge_numerical", ".false.", 0},
    // run_parameters / io_parameters
    {"p_onesnaphotdebugmode", ".false.", 0}, {"output_everything", ".false.", 0}, {"p_io_separate_output", ".true.", 0},
    {"p_master_works_too", ".false.", 1}, {"p_master_skip", "2", 1},   // MPI farm details: replaced by the GPU partition
    // observables_parameters: another executable (Observables.f90)
    {"p_obs_cr_alpha", "", 1}, {"p_obs_dust_alpha", "", 1}, {"p_obs_ignore_small_scale", "", 1}, {"p_obs_scale_with_z", "", 1},
    {"p_obs_redshift_indices", "", 1}, {"p_obs_use_all_redshifts", "", 1},
};

inline bool same_value(const std::string& a0, const std::string& b0) {
  const std::string a = Namelist::lower(a0), b = Namelist::lower(b0);
  if (a == b) return true;
  try { return Namelist::to_bool(a) == Namelist::to_bool(b); } catch (...) {}
  try { return Namelist::to_double(a) == Namelist::to_double(b); } catch (...) {}
  return false;
}

inline RunConfig load_run_config(const std::string& path) {
  RunConfig rc;
  mag_params_default(&rc.params);
  const Namelist nl = Namelist::parse_file(path);
  mag_params_t& p = rc.params;
  std::map<std::string, bool> known;
  // every mag_params_t member has the name of the reference's namelist variable
#define MAG_NL(field) (known[Namelist::lower(#field)] = true, nl.get(#field, &p.field))
  MAG_NL(nsteps_0); MAG_NL(p_nsteps_max); MAG_NL(p_MAX_FAILS); MAG_NL(p_random_seed); MAG_NL(p_no_magnetic_fields_test_run);
  MAG_NL(p_courant_v); MAG_NL(p_courant_eta); MAG_NL(p_nx_ref); MAG_NL(p_nx_MAX); MAG_NL(p_rmax_over_rdisk);
  MAG_NL(frac_seed); MAG_NL(C_alp); MAG_NL(C_floor); MAG_NL(p_floor_kappa); MAG_NL(fmag); MAG_NL(R_kappa);
  MAG_NL(p_ISM_sound_speed_km_s); MAG_NL(p_ISM_kappa); MAG_NL(p_ISM_xi); MAG_NL(p_ISM_gamma); MAG_NL(p_ISM_turbulent_length);
  MAG_NL(p_stellarHeightToRadiusScale); MAG_NL(p_molecularHeightToRadiusScale);
  MAG_NL(p_gasScaleRadiusToStellarScaleRadius_ratio); MAG_NL(p_Rmol_alpha); MAG_NL(p_Rmol_P0); MAG_NL(p_rreg_to_rdisk);
  MAG_NL(p_minimum_density); MAG_NL(p_rdisk_min); MAG_NL(Mgas_disk_min); MAG_NL(rmin_over_rmax);
  MAG_NL(p_ignore_satellite_DM_haloes); MAG_NL(p_use_Pdm); MAG_NL(p_use_Pbulge);
  MAG_NL(p_variable_timesteps); MAG_NL(p_simplified_pressure); MAG_NL(p_halo_contraction); MAG_NL(p_enable_P2);
  MAG_NL(Dyn_quench); MAG_NL(Alg_quench); MAG_NL(Damp); MAG_NL(lFloor); MAG_NL(p_space_varying_floor);
  MAG_NL(p_time_varying_floor); MAG_NL(p_limit_turbulent_scale); MAG_NL(p_allow_positive_shears);
  MAG_NL(p_neumann_boundary_condition_rmax); MAG_NL(p_use_fixed_physical_grid); MAG_NL(p_nn_seed); MAG_NL(p_r_seed_decay);
  MAG_NL(p_outflow_Lsn); MAG_NL(p_outflow_fOB); MAG_NL(p_outflow_etaSN); MAG_NL(p_tOB); MAG_NL(p_N_SN1OB);
  MAG_NL(p_outflow_hot_gas_density); MAG_NL(p_outflow_nu0); MAG_NL(p_outflow_alphahot); MAG_NL(p_outflow_Vhot);
#undef MAG_NL
  std::string s;
  if (nl.get("p_seed_choice", &s)) {
    static const char* seeds[] = {"random", "fraction", "decaying"};
    p.seed_choice = -1;
    for (int k = 0; k < 3; k++) if (Namelist::lower(s) == seeds[k]) p.seed_choice = k;
    if (p.seed_choice < 0)   // 'floor', or a name init_seed_field stops on (seed_field.f90:58-60)
      throw std::runtime_error("parameter file sets p_seed_choice = " + s + ": this seed prescription is not implemented by the CUDA path");
  }
  if (nl.get("p_outflow_type", &s)) {
    static const char* names[] = {"no_outflow", "vturb", "wind", "superbubble_simple", "superbubble"};
    p.outflow_type = -1;   // unknown prescription: mag_create rejects it (outflow.f90:92-94 stops)
    for (int k = 0; k < 5; k++) if (Namelist::lower(s) == names[k]) p.outflow_type = k;
  }
  nl.get("model_name", &rc.model_name);
  nl.get("input_file_name", &rc.input_file_name);
  nl.get("output_file_name", &rc.output_file_name);
  nl.get("info", &rc.info);
  nl.get("p_IO_chunking", &rc.p_IO_chunking);
  nl.get("p_IO_number_of_galaxies_in_chunks", &rc.p_IO_number_of_galaxies_in_chunks);
  nl.get("p_IO_compression", &rc.p_IO_compression);
  nl.get("p_IO_compression_level", &rc.p_IO_compression_level);
  nl.get("p_ncheckpoint", &rc.p_ncheckpoint);
  nl.get("p_max_walltime", &rc.p_max_walltime);
  if (auto q = nl.find("output_quantities_list")) rc.output_quantities_list = *q;
  for (const char* k : {"p_seed_choice", "p_outflow_type", "model_name", "input_file_name", "output_file_name", "info", "p_io_chunking",
                        "p_io_number_of_galaxies_in_chunks", "p_io_compression", "p_io_compression_level", "p_ncheckpoint",
                        "p_max_walltime", "output_quantities_list"})
    known[Namelist::lower(k)] = true;
  // keys outside mag_params_t: unimplemented switches must keep their defaults
  for (const OtherKey& ok : kOtherKeys) {
    known[ok.name] = true;
    const auto* v = nl.find(ok.name);
    if (!v) continue;
    if (ok.kind == 0 && !same_value((*v)[0], ok.dflt))
      throw std::runtime_error(std::string("parameter file sets ") + ok.name + " = " + (*v)[0] +
                               ": this switch is not implemented by the CUDA path and must keep its default (" + ok.dflt + ")");
    if (ok.kind == 1) rc.warnings.push_back(std::string(ok.name) + " has no effect on the dynamo path");
  }
  for (const auto& g : nl.groups)
    for (const auto& kv : g.second)
      if (!known.count(kv.first)) rc.warnings.push_back("unknown key " + kv.first + " in &" + g.first + " ignored");
  return rc;
}

}  // namespace maghost
