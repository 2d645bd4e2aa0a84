// magnetizer_main.cpp -- C++ host driver of the B200 dynamo path.
//
// Mirrors the reference's main program for this operator:
//   program magnetizer (source/Magnetizer.f90:18-38)
//     jobs_prepare            -> read the parameter file (namelists) and the input file
//     jobs_reads_galaxy_list  -> optional 1-based galaxy ids on the command line (single-galaxy mode)
//     jobs_distribute(dynamo_run, write_output, ...)  (source/distributor.f90:180)
//                             -> checkpoint cycles of p_ncheckpoint galaxies (:253-337), each solved by
//                                mag_partition_solve (static + work-stealing chunks over the GPUs) and
//                                followed by a flush of the output file; STOP file / p_max_walltime end
//                                the run gracefully after a cycle (:299-312)
//     write_output            -> unit conversion table of source/output.f90:63-201, Log/* datasets
// The reference cannot be built in this image (no Fortran/MPI/HDF5/GSL toolchain), so the
// host side above the C ABI is C++; INTEGRATION.md shows the ISO_C_BINDING shim that lets the
// unchanged Fortran driver call the same entry points.
//
// Input: the reference's HDF5 input file (own minimal reader), or --raw-input <file> (a flat binary
// catalogue, tools/make_raw_catalogue.py: the synthetic catalogues of BASELINE configs 3-5).
// Output: the reference's HDF5 layout (mag_hdf5_writer.hpp): every dataset has one row per galaxy of
// the INPUT FILE -- also in single-galaxy mode --, chunked + deflate, fill value -1d50 where nothing
// was written; the file is the checkpoint: an existing one is extended, galaxies whose Log/completed
// flag is set are skipped (IO_start_galaxy, IO_hdf5.f90:244-266).
//
//   Magnetizer_b200 <parameters.in> [gal_id ...] [--devices 0,1,...] [--chunk N] [--out file.hdf5]
//                   [--restart] [--stop-after N] [--dry-run] [--raw-input file] [--no-overlap]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include <cuda_runtime.h>

#include "../../include/magnetizer_b200.h"
#include "mag_hdf5_reader.hpp"
#include "mag_hdf5_writer.hpp"
#include "mag_namelist.hpp"

namespace {

const char* const kProfileNames[MAG_NPROFILE] = {"Br", "Bp", "alp_m", "Bzmod", "h", "Omega", "Omega_b", "Omega_d", "Omega_h",
                                                 "Shear", "l", "v", "etat", "tau", "alp_k", "alp", "Uz", "Ur", "n", "Beq", "rkpc"};
const char* const kScalarNames[MAG_NSCALAR] = {"t_Gyr", "dt", "rmax", "Bmax", "Bmax_idx", "Bavg", "Beavg"};

// write_output's unit factors (output.f90:63-201 with constants.f90:40-74)
double unit_factor(const std::string& q) {
  const double km_kpc = 3.08567758e16, mp_g = 1.67262158e-24;
  const double t0_kpcskm = 1.0 / (1.0 / 3 * 0.1 * 10.0), t0_s = t0_kpcskm * km_kpc, h0_km = km_kpc, h0_cm = h0_km * 1e5;
  const double vel = h0_km / t0_s, n0_cm3 = (1.0 / 1e6) * (1.0 / 1e6) / (h0_cm * h0_cm) * (t0_s * t0_s) / mp_g;
  if (q == "alp_m" || q == "alp_k" || q == "alp" || q == "v" || q == "Uz" || q == "Omega" || q == "Shear" || q == "etat") return vel;
  if (q == "h" || q == "l") return 1e3;
  if (q == "n") return n0_cm3;
  return 1.0;
}

// 'Units' and 'Description' attributes of write_output (output.f90:37-201)
struct QuantityDoc { const char *name, *units, *description; };
const QuantityDoc kQuantityDocs[] = {
    {"r", "kpc", "Radius"}, {"rkpc", "kpc", "Radius"}, {"Br", "microgauss", ""}, {"Bp", "microgauss", ""}, {"alp_m", "km/s", ""},
    {"Bzmod", "microgauss", ""}, {"h", "pc", "Disk scaleheight"}, {"Omega", "km/s/kpc", "Angular velocity"},
    {"Shear", "km/s/kpc", "Shear"}, {"l", "pc", "Turbulent length"}, {"v", "km/s", "Turbulent velocity"},
    {"etat", "kpc km/s", ""}, {"tau", "", ""}, {"alp_k", "km/s", ""}, {"Uz", "km/s", "Vertical velocity (outflow)"},
    {"n", "cm^-3", "Number density of diffuse gas"}, {"Beq", "microgauss", "Equipartition field"},
    {"rmax", "kpc", "Radius of maximum |B|"}, {"Bmax", "microgauss", "Maximum B"}, {"Bmax_idx", "", "Index of maximum |B|"},
    {"alp", "km/s", ""}, {"dt", "", ""}, {"Omega_h", "km/s/kpc", "Angular velocity of the halo"},
    {"Omega_b", "km/s/kpc", "Angular velocity of the bulge"}, {"Omega_d", "km/s/kpc", "Angular velocity of the disc"},
    {"Bavg", "microgauss", "Surface average magnetic field"}, {"Beavg", "microgauss", "\"Energy averaged\" magnetic field"},
    {"runtime", "s", "Running time of the galaxy"}, {"nsteps", "", "Number of timesteps"}, {"status", "", "Status codes."}};
maghost::Hdf5Writer::TextAttrs quantity_attrs(const std::string& q) {
  maghost::Hdf5Writer::TextAttrs a;
  for (const auto& d : kQuantityDocs)
    if (q == d.name) {
      if (d.units[0]) a.push_back({"Units", d.units});
      if (d.description[0]) a.push_back({"Description", d.description});
    }
  return a;
}

// the namelists as gfortran prints them into one record (IO_hdf5.f90:191-216; parsed by
// python/magnetizer/parameters.py): "&NAME  KEY=value, KEY=value, /"
std::string nml_text(const char* name, const std::vector<std::pair<std::string, std::string>>& kv) {
  std::string s = std::string("&") + name + " ";
  for (const auto& p : kv) {
    std::string k = p.first;
    for (auto& c : k) c = (char)std::toupper((unsigned char)c);
    s += " " + k + "=" + p.second + ",";
  }
  return s + " /";
}
std::string fnum(double v) { char b[40]; std::snprintf(b, sizeof b, "%.17g", v); return b; }
std::string fint(long v) { return std::to_string(v); }
std::string fbool(int v) { return v ? "T" : "F"; }

void write_parameter_attributes(maghost::Hdf5Writer& w, const maghost::RunConfig& cfg) {
  const mag_params_t& p = cfg.params;
  w.add_root_attribute("Model name", cfg.model_name);
  w.add_root_attribute("run_parameters", nml_text("RUN_PARAMETERS", {{"model_name", "\"" + cfg.model_name + "\""}, {"info", fint(cfg.info)},
      {"nsteps_0", fint(p.nsteps_0)}, {"p_courant_v", fnum(p.p_courant_v)}, {"p_courant_eta", fnum(p.p_courant_eta)},
      {"p_variable_timesteps", fbool(p.p_variable_timesteps)}, {"p_nsteps_max", fint(p.p_nsteps_max)},
      {"p_no_magnetic_fields_test_run", fbool(p.p_no_magnetic_fields_test_run)}, {"p_MAX_FAILS", fint(p.p_MAX_FAILS)},
      {"p_random_seed", fint(p.p_random_seed)}}));
  w.add_root_attribute("io_parameters", nml_text("IO_PARAMETERS", {{"output_file_name", "\"" + cfg.output_file_name + "\""},
      {"input_file_name", "\"" + cfg.input_file_name + "\""}, {"p_IO_separate_output", "T"}, {"p_IO_chunking", fbool(cfg.p_IO_chunking)},
      {"p_IO_number_of_galaxies_in_chunks", fint(cfg.p_IO_number_of_galaxies_in_chunks)}, {"p_IO_compression", fbool(cfg.p_IO_compression)},
      {"p_IO_compression_level", fint(cfg.p_IO_compression_level)}, {"p_ncheckpoint", fint(cfg.p_ncheckpoint)}}));
  w.add_root_attribute("grid_parameters", nml_text("GRID_PARAMETERS", {{"p_nx_ref", fint(p.p_nx_ref)}, {"p_nx_MAX", fint(p.p_nx_MAX)},
      {"p_rmax_over_rdisk", fnum(p.p_rmax_over_rdisk)}, {"p_use_fixed_physical_grid", fbool(p.p_use_fixed_physical_grid)}}));
  w.add_root_attribute("dynamo_parameters", nml_text("DYNAMO_PARAMETERS", {{"Dyn_quench", fbool(p.Dyn_quench)}, {"Alg_quench", fbool(p.Alg_quench)},
      {"lFloor", fbool(p.lFloor)}, {"Damp", fbool(p.Damp)}, {"frac_seed", fnum(p.frac_seed)},
      {"p_seed_choice", std::string("\"") + (p.seed_choice == MAG_SEED_FRACTION ? "fraction" : p.seed_choice == MAG_SEED_DECAYING ? "decaying" : "random") + "\""},
      {"p_r_seed_decay", fnum(p.p_r_seed_decay)}, {"p_nn_seed", fint(p.p_nn_seed)},
      {"p_neumann_boundary_condition_rmax", fbool(p.p_neumann_boundary_condition_rmax)},
      {"C_alp", fnum(p.C_alp)}, {"C_floor", fnum(p.C_floor)}, {"p_floor_kappa", fnum(p.p_floor_kappa)},
      {"p_space_varying_floor", fbool(p.p_space_varying_floor)}, {"p_time_varying_floor", fbool(p.p_time_varying_floor)},
      {"fmag", fnum(p.fmag)}, {"R_kappa", fnum(p.R_kappa)}}));
  w.add_root_attribute("ISM_and_disk_parameters", nml_text("ISM_AND_DISK_PARAMETERS", {
      {"p_ISM_sound_speed_km_s", fnum(p.p_ISM_sound_speed_km_s)}, {"p_ISM_kappa", fnum(p.p_ISM_kappa)}, {"p_ISM_xi", fnum(p.p_ISM_xi)},
      {"p_ISM_gamma", fnum(p.p_ISM_gamma)}, {"p_ISM_turbulent_length", fnum(p.p_ISM_turbulent_length)},
      {"p_limit_turbulent_scale", fbool(p.p_limit_turbulent_scale)},
      {"p_stellarHeightToRadiusScale", fnum(p.p_stellarHeightToRadiusScale)},
      {"p_molecularHeightToRadiusScale", fnum(p.p_molecularHeightToRadiusScale)},
      {"p_gasScaleRadiusToStellarScaleRadius_ratio", fnum(p.p_gasScaleRadiusToStellarScaleRadius_ratio)},
      {"p_Rmol_alpha", fnum(p.p_Rmol_alpha)}, {"p_Rmol_P0", fnum(p.p_Rmol_P0)},
      {"p_allow_positive_shears", fbool(p.p_allow_positive_shears)}, {"p_simplified_pressure", fbool(p.p_simplified_pressure)},
      {"p_rreg_to_rdisk", fnum(p.p_rreg_to_rdisk)}, {"p_rdisk_min", fnum(p.p_rdisk_min)}, {"Mgas_disk_min", fnum(p.Mgas_disk_min)},
      {"rmin_over_rmax", fnum(p.rmin_over_rmax)}, {"p_ignore_satellite_DM_haloes", fbool(p.p_ignore_satellite_DM_haloes)},
      {"p_enable_P2", fbool(p.p_enable_P2)}, {"p_minimum_density", fnum(p.p_minimum_density)}, {"p_use_Pdm", fbool(p.p_use_Pdm)},
      {"p_use_Pbulge", fbool(p.p_use_Pbulge)}, {"p_halo_contraction", fbool(p.p_halo_contraction)}}));
  static const char* const kOutflow[] = {"no_outflow", "vturb", "wind", "superbubble_simple", "superbubble"};
  w.add_root_attribute("outflow_parameters", nml_text("OUTFLOW_PARAMETERS", {
      {"p_outflow_type", std::string("\"") + kOutflow[p.outflow_type >= 0 && p.outflow_type < 5 ? p.outflow_type : 0] + "\""},
      {"p_outflow_Lsn", fnum(p.p_outflow_Lsn)}, {"p_outflow_fOB", fnum(p.p_outflow_fOB)}, {"p_outflow_etaSN", fnum(p.p_outflow_etaSN)},
      {"p_tOB", fnum(p.p_tOB)}, {"p_N_SN1OB", fnum(p.p_N_SN1OB)}, {"p_outflow_hot_gas_density", fnum(p.p_outflow_hot_gas_density)},
      {"p_outflow_nu0", fnum(p.p_outflow_nu0)}, {"p_outflow_alphahot", fnum(p.p_outflow_alphahot)}, {"p_outflow_Vhot", fnum(p.p_outflow_Vhot)}}));
  w.add_root_attribute("File creation", "magnetizer_b200 host driver");
}

}  // namespace


namespace {

// flat binary catalogue (tools/make_raw_catalogue.py): "MAGRAW1\0", int64 ngal, int64 nz, t[nz], inputs[13][ngal][nz]
maghost::MagnetizerInput load_raw_input(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  char magic[8];
  int64_t dims[2];
  maghost::MagnetizerInput in;
  bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "MAGRAW1", 8) == 0 && std::fread(dims, 8, 2, f) == 2;
  if (ok) {
    in.ngal = (int)dims[0]; in.nz = (int)dims[1];
    in.t_gyr.resize(in.nz);
    in.inputs.resize((size_t)13 * in.ngal * in.nz);
    ok = std::fread(in.t_gyr.data(), 8, in.nz, f) == (size_t)in.nz && std::fread(in.inputs.data(), 8, in.inputs.size(), f) == in.inputs.size();
  }
  std::fclose(f);
  if (!ok) throw std::runtime_error(path + ": not a raw catalogue");
  return in;
}

struct HostBuf {   // page-locked host memory every GPU can store into (mag_host_alloc)
  void* p = nullptr;
  size_t bytes = 0;
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) mag_host_free(p);
    p = mag_host_alloc(n);
    bytes = p ? n : 0;
    if (!p) throw std::runtime_error("cannot allocate page-locked host memory");
  }
  ~HostBuf() { if (p) mag_host_free(p); }
};

double seconds_since(const std::chrono::steady_clock::time_point& t0) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <parameters.in> [gal_id ...] [--devices 0,1,...] [--chunk N] [--out file.hdf5]\n"
                         "          [--restart] [--stop-after N] [--dry-run] [--raw-input file] [--no-overlap]\n", argv[0]);
    return 2;
  }
  try {
    const auto t_start = std::chrono::steady_clock::now();
    maghost::RunConfig cfg = maghost::load_run_config(argv[1]);
    for (const auto& wmsg : cfg.warnings) std::fprintf(stderr, "warning: %s\n", wmsg.c_str());
    std::vector<int32_t> list, devices;
    int chunk = 0;
    std::string out_path, raw_input;
    bool dry_run = false, restart = false, overlap = true;
    long stop_after = -1;
    for (int i = 2; i < argc; i++) {
      const std::string a = argv[i];
      if (a == "--dry-run") { dry_run = true; continue; }
      if (a == "--restart") { restart = true; continue; }
      if (a == "--no-overlap") { overlap = false; continue; }
      if (a == "--stop-after" && i + 1 < argc) { stop_after = std::atol(argv[++i]); continue; }
      if (a == "--raw-input" && i + 1 < argc) { raw_input = argv[++i]; continue; }
      if (a == "--devices" && i + 1 < argc) {
        for (char* tok = std::strtok(argv[++i], ","); tok; tok = std::strtok(nullptr, ",")) devices.push_back(std::atoi(tok));
      } else if (a == "--chunk" && i + 1 < argc) chunk = std::atoi(argv[++i]);
      else if (a == "--out" && i + 1 < argc) out_path = argv[++i];
      else list.push_back(std::atoi(a.c_str()));
    }
    const maghost::MagnetizerInput in = raw_input.empty() ? maghost::load_magnetizer_input(cfg.input_file_name) : load_raw_input(raw_input);
    // jobs_reads_galaxy_list: ids are 1-based rows of the input file; the output has one row per galaxy of the FILE
    const int ngals = in.ngal, nz = in.nz, nr = cfg.params.p_nx_ref;
    if (list.empty()) for (int g = 1; g <= ngals; g++) list.push_back(g);
    std::vector<char> selected(ngals, 0);
    for (int g : list) {
      if (g < 1 || g > ngals) throw std::runtime_error("galaxy id out of range: " + std::to_string(g));
      selected[g - 1] = 1;
    }
    // requested quantities (output_quantities_list; default: everything)
    std::vector<int32_t> pq, sq;
    std::vector<std::string> names = cfg.output_quantities_list;
    if (names.empty()) { for (auto n : kProfileNames) names.push_back(n); for (auto n : kScalarNames) names.push_back(n); }
    bool have_rkpc = false;
    for (const auto& q : names) {
      bool found = false;
      for (int i = 0; i < MAG_NPROFILE; i++) if (q == kProfileNames[i]) { pq.push_back(i); found = true; if (i == MAG_Q_RKPC) have_rkpc = true; }
      for (int i = 0; i < MAG_NSCALAR; i++) if (q == kScalarNames[i]) { sq.push_back(i); found = true; }
      if (!found) std::fprintf(stderr, "warning: output quantity '%s' is not produced by this path; skipped\n", q.c_str());
    }
    if (!have_rkpc) pq.push_back(MAG_Q_RKPC);  // Output/r is always written (output.f90:56-60)
    // ---- resume: the output file doubles as checkpoint (distributor.f90:150-156).  Galaxies whose
    // Log/completed flag is set are not solved again (IO_start_galaxy, IO_hdf5.f90:244-266);
    // --restart ignores an existing file, --stop-after N ends this invocation after N galaxies.
    if (out_path.empty()) out_path = cfg.output_file_name;
    std::unique_ptr<maghost::H5Reader> prev;
    std::vector<char> completed(ngals, 0);
    if (!restart) {
      if (FILE* tf = std::fopen(out_path.c_str(), "rb")) {
        std::fclose(tf);
        prev.reset(new maghost::H5Reader(out_path));
        const std::vector<uint64_t> shp = prev->exists("Output/r") ? prev->shape("Output/r") : std::vector<uint64_t>();
        if (shp != std::vector<uint64_t>{(uint64_t)ngals, (uint64_t)nr, (uint64_t)nz})
          throw std::runtime_error("cannot resume: " + out_path + " was written for another input file or grid (use --restart)");
        const maghost::H5Reader::Data comp = prev->read("Log/completed");
        for (int g = 0; g < ngals; g++) completed[g] = comp.v[g] > 0.5;   // never-written rows hold the fill value
      }
    }
    std::vector<int> todo;
    size_t n_done_before = 0;
    for (int g = 0; g < ngals; g++) {
      if (!selected[g]) continue;
      if (completed[g]) n_done_before++; else todo.push_back(g);
    }
    const size_t n_open = todo.size();
    if (stop_after >= 0 && todo.size() > (size_t)stop_after) todo.resize((size_t)stop_after);
    if (dry_run) {  // host-side logic only (parameter file, input file, galaxy list, quantity list): no device needed
      std::printf("resume %s: %zu of %zu galaxies completed, %zu to solve now\n", prev ? "yes" : "no", n_done_before, list.size(), todo.size());
      std::printf("dry-run ngal %d nz %d list %zu p_courant_v %.17g p_nx_ref %d p_random_seed %d\n", ngals, nz, list.size(),
                  cfg.params.p_courant_v, cfg.params.p_nx_ref, cfg.params.p_random_seed);
      for (int c = 0; c < 13; c++) {
        double sum = 0;
        for (int g = 0; g < ngals; g++) if (selected[g]) for (int iz = 0; iz < nz; iz++) sum += in.inputs[((size_t)c * ngals + g) * nz + iz];
        std::printf("column %s sum %.17g\n", maghost::kInputColumns[c], sum);
      }
      std::printf("t %.17g .. %.17g\nprofiles", in.t_gyr.front(), in.t_gyr.back());
      for (int q : pq) std::printf(" %s", kProfileNames[q]);
      std::printf("\nscalars");
      for (int q : sq) std::printf(" %s", kScalarNames[q]);
      std::printf("\n");
      return 0;
    }
    if (devices.empty()) {
      int nd = 0;
      if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) throw std::runtime_error("no CUDA device (there is no CPU fallback)");
      for (int d = 0; d < nd; d++) devices.push_back(d);
    }
    mag_partition_t* part = nullptr;
    if (mag_partition_create(&cfg.params, (int)devices.size(), devices.data(), &part) != MAG_OK)
      throw std::runtime_error(std::string("cannot create the solver: ") + mag_partition_last_error());

    // ---- the output file (IO_start, IO_hdf5.f90:114-241): new, or the existing one extended in place
    maghost::Hdf5Writer::Options wopt;
    wopt.chunking = cfg.p_IO_chunking != 0; wopt.gals_per_chunk = cfg.p_IO_number_of_galaxies_in_chunks;
    wopt.compression = cfg.p_IO_compression != 0; wopt.level = cfg.p_IO_compression_level;
    maghost::Hdf5Writer w(out_path, wopt, /*extend=*/prev != nullptr);
    w.add_group("Output");
    w.add_group("Log");
    write_parameter_attributes(w, cfg);
    struct Col { std::string path; int h; };
    std::vector<Col> prof_cols, scal_cols;
    int h_r = -1;
    const std::vector<uint64_t> d3{(uint64_t)ngals, (uint64_t)nr, (uint64_t)nz}, d2{(uint64_t)ngals, (uint64_t)nz}, d1{(uint64_t)ngals, 1};
    for (int q : pq) {
      const std::string name = kProfileNames[q];
      if (name != "rkpc" || have_rkpc) prof_cols.push_back({"Output/" + name, w.create_dataset("Output", name, maghost::Hdf5Writer::F64, d3, quantity_attrs(name))});
      else prof_cols.push_back({"", -1});
      if (name == "rkpc") h_r = w.create_dataset("Output", "r", maghost::Hdf5Writer::F64, d3, quantity_attrs("r"));
    }
    for (int q : sq) scal_cols.push_back({std::string("Output/") + kScalarNames[q], w.create_dataset("Output", kScalarNames[q], maghost::Hdf5Writer::F64, d2, quantity_attrs(kScalarNames[q]))});
    const int h_status = w.create_dataset("Log", "status", maghost::Hdf5Writer::CHAR1, d2, quantity_attrs("status"), false);
    const int h_nsteps = w.create_dataset("Log", "nsteps", maghost::Hdf5Writer::F64, d1, quantity_attrs("nsteps"));
    const int h_completed = w.create_dataset("Log", "completed", maghost::Hdf5Writer::F64, d1);
    const int h_runtime = w.create_dataset("Log", "runtime", maghost::Hdf5Writer::F64, d1, quantity_attrs("runtime"));
    std::vector<std::pair<std::string, int>> all_sets;
    for (auto& c : prof_cols) if (c.h >= 0) all_sets.push_back({c.path, c.h});
    all_sets.push_back({"Output/r", h_r});
    for (auto& c : scal_cols) all_sets.push_back({c.path, c.h});
    all_sets.push_back({"Log/status", h_status}); all_sets.push_back({"Log/nsteps", h_nsteps});
    all_sets.push_back({"Log/completed", h_completed}); all_sets.push_back({"Log/runtime", h_runtime});
    if (prev)
      for (auto& kv : all_sets) {
        if (!prev->exists(kv.first)) continue;   // a quantity the earlier invocation did not ask for
        w.adopt(kv.second, prev->chunks(kv.first));
        w.adopt_contiguous(kv.second, prev->contiguous_address(kv.first));
      }
    const size_t cs = (size_t)w.gals_per_block(h_r), nblocks = (size_t)w.n_blocks(h_r);

    std::printf("Magnetizer (B200 path): '%s'  %d galaxies x %d snapshots in the input, %zu selected (%zu already completed, %zu to solve), %zu GPU(s)\n",
                cfg.model_name.c_str(), ngals, nz, list.size(), n_done_before, todo.size(), devices.size());
    (void)n_open;
    // ---- checkpoint cycles (distributor.f90:253-337): blocks of galaxies are accumulated until at least
    // p_ncheckpoint galaxies are to be solved; solve, write, flush; then look for the STOP file / the walltime
    std::vector<char> is_todo(ngals, 0);
    for (int g : todo) is_todo[g] = 1;
    // The results of a cycle are written (unit conversion, deflate, flush) by a writer thread while the next cycle is
    // packed and solved: two sets of page-locked buffers, unless --no-overlap or two sets would not fit the host memory.
    struct Cycle {
      HostBuf b_in, b_prof, b_scal, b_stat, b_nsteps, b_err, b_ps, b_flags;
      std::vector<int> sel;            // catalogue indices solved in this cycle, ascending
      size_t kb0 = 0, kb1 = 0, solved_after = 0;
      double cyc_s = 0;
      long long cyc_stages = 0;
    } cyc[2];
    std::vector<int32_t> done((size_t)devices.size()), stolen((size_t)devices.size()), done_tot(devices.size(), 0), stolen_tot(devices.size(), 0);
    std::map<char, long> counts;
    double solve_s = 0, write_s = 0, stage_s = 0, wait_s = 0;
    const double setup_s = seconds_since(t_start);   // namelist, input file, output file, contexts
    long long total_stages = 0;
    size_t solved = 0, cycles = 0;
    bool stopped = false;
    const size_t ncheck = (size_t)std::max(1, cfg.p_ncheckpoint);
    {
      const double cycle_bytes = (double)std::min(ncheck + cs, todo.size()) * nz * (8.0 * (13 + pq.size() * (double)nr + sq.size()) + 1.0);
      const double host_bytes = (double)sysconf(_SC_PHYS_PAGES) * (double)sysconf(_SC_PAGESIZE);
      if (overlap && 2.0 * cycle_bytes > 0.4 * host_bytes) {
        overlap = false;
        if (cfg.info > 0) std::printf("  (two result sets of %.1f GB do not fit the host memory: cycles are written before the next one is solved)\n", cycle_bytes / 1e9);
      }
    }

    // -- write_output (output.f90:23-236): unit conversion, (ngals, nx, nz) layout, Log group; block by block
    auto write_cycle = [&](const Cycle& C) {
      const std::vector<int>& sel = C.sel;
      const int nsel = (int)sel.size();
      const double* profiles = (const double*)C.b_prof.p;
      const double* scalars = (const double*)C.b_scal.p;
      const char* status = (const char*)C.b_stat.p;
      const int32_t *nsteps = (const int32_t*)C.b_nsteps.p, *error = (const int32_t*)C.b_err.p;
      const int64_t* stages = (const int64_t*)C.b_ps.p;
      const double cyc_s = C.cyc_s;
      const long long cyc_stages = C.cyc_stages;
      const auto t1 = std::chrono::steady_clock::now();
      std::vector<double> blk3, blk2(cs * (size_t)nz), blk1(cs);
      std::vector<char> blkc(cs * (size_t)nz);
      std::vector<maghost::Hdf5Writer::ChunkJob> jobs;   // profile chunks of the whole cycle
      size_t k_lo = 0;
      for (size_t kb = C.kb0; kb < C.kb1; kb++) {
        const size_t g0 = kb * cs, nb = std::min(cs, (size_t)ngals - g0);
        size_t k_hi = k_lo;
        while (k_hi < (size_t)nsel && (size_t)sel[k_hi] < g0 + nb) k_hi++;
        if (k_hi == k_lo) continue;                                   // nothing solved in this block: its chunks stay as they are
        bool merge = false;                                           // rows an earlier invocation completed are kept
        for (size_t g = g0; g < g0 + nb; g++) if (completed[g]) merge = true;
        auto base3 = [&](const std::string& path) {
          blk3.resize(cs * (size_t)nr * nz);
          if (merge && prev && prev->exists(path)) { const maghost::H5Reader::Data d = prev->read_rows(path, g0, nb); std::copy(d.v.begin(), d.v.end(), blk3.begin()); }
          else std::fill(blk3.begin(), blk3.begin() + nb * (size_t)nr * nz, maghost::Hdf5Writer::kFill);
        };
        auto base2 = [&](const std::string& path, std::vector<double>& b, size_t row) {
          if (merge && prev && prev->exists(path)) { const maghost::H5Reader::Data d = prev->read_rows(path, g0, nb); std::copy(d.v.begin(), d.v.end(), b.begin()); }
          else std::fill(b.begin(), b.begin() + nb * row, maghost::Hdf5Writer::kFill);
        };
        for (size_t q = 0; q < pq.size(); q++) {
          const std::string name = kProfileNames[pq[q]];
          const double fac = unit_factor(name);
          const double* src = profiles + q * (size_t)nsel * nz * nr;
          if (wopt.chunking && !merge) {
            // a chunk holds all radii of ONE snapshot of the block's galaxies (IO_hdf5.f90:940-952), and the solver's rows are
            // [galaxy][snapshot][radius]: every chunk is filled from contiguous rows, on the writer's threads
            const int* selp = sel.data();
            for (int iz = 0; iz < nz; iz++) {
              maghost::Hdf5Writer::ChunkJob job;
              job.kb = kb; job.il = (uint64_t)iz;
              job.fill = [=](char* raw, uint64_t) {
                double* dst = (double*)raw;
                for (size_t k = k_lo; k < k_hi; k++) {
                  const double* row = src + (k * nz + iz) * nr;
                  double* o = dst + ((size_t)selp[k] - g0) * nr;
                  for (int i = 0; i < nr; i++) o[i] = (row[i] == MAG_INVALID) ? row[i] : row[i] * fac;
                }
              };
              if (prof_cols[q].h >= 0) { job.h = prof_cols[q].h; jobs.push_back(job); }
              if (name == "rkpc") { job.h = h_r; jobs.push_back(job); }
            }
            continue;
          }
          base3(name == "rkpc" && !have_rkpc ? "Output/r" : "Output/" + name);
          for (size_t k = k_lo; k < k_hi; k++) {
            double* dst = &blk3[((size_t)sel[k] - g0) * nr * nz];
            for (int iz = 0; iz < nz; iz++)
              for (int i = 0; i < nr; i++) {
                const double v = src[(k * nz + iz) * nr + i];
                dst[(size_t)i * nz + iz] = (v == MAG_INVALID) ? v : v * fac;
              }
          }
          if (prof_cols[q].h >= 0) w.write_block(prof_cols[q].h, kb, blk3.data());
          if (name == "rkpc") w.write_block(h_r, kb, blk3.data());
        }
        for (size_t q = 0; q < sq.size(); q++) {
          base2(scal_cols[q].path, blk2, nz);
          for (size_t k = k_lo; k < k_hi; k++)
            std::memcpy(&blk2[((size_t)sel[k] - g0) * nz], &scalars[(q * (size_t)nsel + k) * nz], sizeof(double) * nz);
          w.write_block(scal_cols[q].h, kb, blk2.data());
        }
        if (merge && prev) { const maghost::H5Reader::Data d = prev->read_rows("Log/status", g0, nb); for (size_t i = 0; i < nb * (size_t)nz; i++) blkc[i] = (char)(int)d.v[i]; }
        else std::fill(blkc.begin(), blkc.begin() + nb * (size_t)nz, '\0');
        for (size_t k = k_lo; k < k_hi; k++) {
          std::memcpy(&blkc[((size_t)sel[k] - g0) * nz], &status[k * (size_t)nz], nz);
          for (int iz = 0; iz < nz; iz++) counts[status[k * (size_t)nz + iz]]++;
        }
        w.write_block(h_status, kb, blkc.data());
        // Log datasets are (ngals, 1) in the reference's file (Fortran (1, ngals), output.f90:37-54)
        base2("Log/nsteps", blk1, 1);
        for (size_t k = k_lo; k < k_hi; k++) blk1[(size_t)sel[k] - g0] = nsteps[k];
        w.write_block(h_nsteps, kb, blk1.data());
        base2("Log/completed", blk1, 1);
        for (size_t k = k_lo; k < k_hi; k++) blk1[(size_t)sel[k] - g0] = error[k] ? 0.0 : 1.0;  // only written when runtime > 0 (distributor.f90:363)
        w.write_block(h_completed, kb, blk1.data());
        base2("Log/runtime", blk1, 1);
        for (size_t k = k_lo; k < k_hi; k++) blk1[(size_t)sel[k] - g0] = cyc_stages > 0 ? cyc_s * (double)stages[k] / (double)cyc_stages : 0.0;
        w.write_block(h_runtime, kb, blk1.data());
        k_lo = k_hi;
      }
      if (!jobs.empty()) w.write_chunks(jobs);
      w.flush();                                                      // IO_flush every checkpoint cycle (distributor.f90:336)
      const double ws = seconds_since(t1);
      write_s += ws;
      cycles++;
      if (cfg.info > 1) std::printf("  cycle %zu: %d galaxies solved in %.3f s, written and flushed in %.3f s (%zu of %zu done)\n", cycles, nsel, cyc_s, ws, C.solved_after, todo.size());
    };
    std::thread writer;
    std::exception_ptr writer_error;
    auto join_writer = [&] {
      if (!writer.joinable()) return;
      const auto tw = std::chrono::steady_clock::now();
      writer.join();
      wait_s += seconds_since(tw);
      if (writer_error) std::rethrow_exception(writer_error);
    };
    struct JoinGuard { std::thread& t; ~JoinGuard() { if (t.joinable()) t.join(); } } join_guard{writer};   // an exception on the solve side

    int cur = 0;
    for (size_t kb0 = 0; kb0 < nblocks && !stopped;) {
      Cycle& C = cyc[cur];             // its previous results (two cycles ago) are on disk: the writer of the last cycle started after that join
      std::vector<int>& sel = C.sel;
      sel.clear();
      size_t kb1 = kb0;
      while (kb1 < nblocks && sel.size() < ncheck) {
        for (size_t g = kb1 * cs; g < std::min((kb1 + 1) * cs, (size_t)ngals); g++) if (is_todo[g]) sel.push_back((int)g);
        kb1++;
      }
      if (sel.empty()) { kb0 = kb1; continue; }
      const int nsel = (int)sel.size();
      // -- solve the cycle's galaxies: packed inputs in, packed results out (page-locked: the kernels store into them)
      const auto ts = std::chrono::steady_clock::now();
      C.b_in.ensure(sizeof(double) * 13 * nsel * nz);
      C.b_prof.ensure(sizeof(double) * pq.size() * nsel * nz * nr); C.b_scal.ensure(sizeof(double) * std::max<size_t>(1, sq.size()) * nsel * nz);
      C.b_stat.ensure((size_t)nsel * nz); C.b_nsteps.ensure(4 * (size_t)nsel); C.b_err.ensure(4 * (size_t)nsel); C.b_ps.ensure(8 * (size_t)nsel); C.b_flags.ensure(4 * (size_t)nsel);
      double* sin = (double*)C.b_in.p;
      std::vector<int32_t> sel_ids(nsel);
      for (int k = 0; k < nsel; k++) sel_ids[k] = sel[k] + 1;
      for (int c = 0; c < 13; c++)
        for (int k = 0; k < nsel; k++)
          std::memcpy(&sin[((size_t)c * nsel + k) * nz], &in.inputs[((size_t)c * ngals + sel[k]) * nz], sizeof(double) * nz);
      mag_outputs_t out{(double*)C.b_prof.p, sq.empty() ? nullptr : (double*)C.b_scal.p, (char*)C.b_stat.p, (int32_t*)C.b_nsteps.p, (int32_t*)C.b_err.p,
                        (int64_t*)C.b_ps.p, (uint32_t*)C.b_flags.p};
      stage_s += seconds_since(ts);
      const auto t0 = std::chrono::steady_clock::now();
      const int rc = mag_partition_solve(part, nsel, sel_ids.data(), nz, in.t_gyr.data(), sin, (int)pq.size(), pq.data(), (int)sq.size(), sq.data(),
                                         &out, chunk, done.data(), stolen.data());
      if (rc != MAG_OK) throw std::runtime_error(std::string("solve failed (") + std::to_string(rc) + "): " + mag_partition_last_error());
      C.cyc_s = seconds_since(t0);
      solve_s += C.cyc_s;
      C.cyc_stages = 0;
      const int64_t* stages = (const int64_t*)C.b_ps.p;
      for (int k = 0; k < nsel; k++) C.cyc_stages += stages[k];
      total_stages += C.cyc_stages;
      for (size_t d = 0; d < devices.size(); d++) { done_tot[d] += done[d]; stolen_tot[d] += stolen[d]; }
      solved += (size_t)nsel;
      C.kb0 = kb0; C.kb1 = kb1; C.solved_after = solved;
      join_writer();                   // the previous cycle is on disk
      writer = std::thread([&, pc = &C] { try { write_cycle(*pc); } catch (...) { writer_error = std::current_exception(); } });
      if (!overlap) join_writer();
      else cur ^= 1;
      kb0 = kb1;
      // graceful stop (distributor.f90:299-312): a file named STOP in the working directory, or the walltime
      if (FILE* sf = std::fopen("STOP", "rb")) { std::fclose(sf); stopped = true; std::printf("  STOP file found: exiting after this cycle\n"); }
      if (cfg.p_max_walltime > 0 && seconds_since(t_start) > cfg.p_max_walltime) { stopped = true; std::printf("  p_max_walltime reached: exiting after this cycle\n"); }
    }
    join_writer();
    w.close();
    prev.reset();
    mag_partition_destroy(part);
    std::printf("  solved %zu galaxies in %.3f s  (%.1f galaxies/s, %.3e point-stages/s); output written in %.3f s (%.1f MB, %zu flush(es))\n",
                solved, solve_s, solve_s > 0 ? solved / solve_s : 0.0, solve_s > 0 ? total_stages / solve_s : 0.0, write_s,
                (double)w.bytes_written() / 1e6, cycles);
    std::printf("  wall %.3f s: setup (namelist, input, contexts) %.3f, page-locked buffers + input packing %.3f, solve %.3f, write_output + flush %.3f "
                "(%s; the solve side waited %.3f s for the writer)\n",
                seconds_since(t_start), setup_s, stage_s, solve_s, write_s, overlap ? "on a writer thread beside the next cycle's solve" : "between the solves", wait_s);
    for (size_t d = 0; d < devices.size(); d++)
      std::printf("  GPU %d: %d chunks (%d stolen)\n", devices[d], done_tot[d], stolen_tot[d]);
    std::printf("  status codes:");
    for (auto& kv : counts) std::printf(" '%c' x%ld", kv.first, kv.second);
    std::printf("\n  wrote %s\n", out_path.c_str());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "Magnetizer_b200: %s\n", e.what());
    return 1;
  }
}
