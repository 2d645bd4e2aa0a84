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

// I/O settings of io_parameters the driver needs (global_input_parameters.f90:30-62)
struct RunConfig {
  mag_params_t params;
  std::string model_name = "Magnetizer run";
  std::string input_file_name, output_file_name;
  std::vector<std::string> output_quantities_list;
  int info = 1;
};

inline RunConfig load_run_config(const std::string& path) {
  RunConfig rc;
  mag_params_default(&rc.params);
  const Namelist nl = Namelist::parse_file(path);
  mag_params_t& p = rc.params;
  // every mag_params_t member has the name of the reference's namelist variable
#define MAG_NL(field) nl.get(#field, &p.field)
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
#undef MAG_NL
  std::string s;
  if (nl.get("p_seed_choice", &s)) p.seed_choice = (Namelist::lower(s) == "random") ? 0 : 1;
  if (nl.get("p_outflow_type", &s)) p.outflow_type = (Namelist::lower(s) == "no_outflow") ? 0 : 1;
  nl.get("model_name", &rc.model_name);
  nl.get("input_file_name", &rc.input_file_name);
  nl.get("output_file_name", &rc.output_file_name);
  nl.get("info", &rc.info);
  if (auto q = nl.find("output_quantities_list")) rc.output_quantities_list = *q;
  return rc;
}

}  // namespace maghost
