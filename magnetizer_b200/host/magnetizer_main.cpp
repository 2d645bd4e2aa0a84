// magnetizer_main.cpp -- C++ host driver of the B200 dynamo path.
//
// Mirrors the reference's main program for this operator:
//   program magnetizer (source/Magnetizer.f90:18-38)
//     jobs_prepare            -> read the parameter file (namelists) and the input file
//     jobs_reads_galaxy_list  -> optional 1-based galaxy ids on the command line (single-galaxy mode)
//     jobs_distribute(dynamo_run, write_output, ...)  (source/distributor.f90:180)
//                             -> mag_solve_partitioned: static + work-stealing chunks over the GPUs
//     write_output            -> unit conversion table of source/output.f90:63-201, Log/* datasets
// The reference cannot be built in this image (no Fortran/MPI/HDF5/GSL toolchain), so the
// host side above the C ABI is C++; INTEGRATION.md shows the ISO_C_BINDING shim that lets the
// unchanged Fortran driver call the same entry points.
//
// Input: the reference's HDF5 input file (own minimal reader).  Output: the reference's
// dataset names and shapes (Output/<q> [ngal][nx][nz], scalars [ngal][nz], Log/status ...),
// written either as an HDF5 file in the reference's layout (mag_hdf5_writer.hpp: /Output, /Log,
// units/description attributes, the namelists as root attributes; the default) or, when the
// output name ends in .npz, as an uncompressed .npz container (readable with numpy.load).
//
//   Magnetizer_b200 <parameters.in> [gal_id ...] [--devices 0,1,...] [--chunk N] [--out file.npz]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <zlib.h>

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

// One output sink for both containers: HDF5 (reference layout) or .npz
struct OutputFile {
  std::unique_ptr<maghost::Hdf5Writer> h5;
  std::unique_ptr<struct NpzWriter> npz;
  void add_f64(const std::string& group, const std::string& name, const std::vector<size_t>& shape, const double* data);
  void add_status(const std::vector<size_t>& shape, const char* data);
  void add_i32(const std::string& group, const std::string& name, const std::vector<size_t>& shape, const int32_t* data);
  void close();
};

// ---- uncompressed .npz (zip of .npy files, "stored" method)
struct NpzWriter {
  FILE* f;
  struct Entry { std::string name; uint32_t crc, size, offset; };
  std::vector<Entry> entries;
  explicit NpzWriter(const std::string& path) : f(std::fopen(path.c_str(), "wb")) {}
  bool ok() const { return f != nullptr; }
  static void le16(std::string& s, uint16_t v) { s.push_back((char)(v & 255)); s.push_back((char)(v >> 8)); }
  static void le32(std::string& s, uint32_t v) { for (int i = 0; i < 4; i++) s.push_back((char)((v >> (8 * i)) & 255)); }
  void add(const std::string& key, const char* descr, const std::vector<size_t>& shape, const void* data, size_t nbytes) {
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
    for (size_t i = 0; i < shape.size(); i++) hdr += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
    hdr += "), }";
    while ((10 + hdr.size() + 1) % 64 != 0) hdr.push_back(' ');
    hdr.push_back('\n');
    std::string npy("\x93NUMPY\x01\x00", 8);
    le16(npy, (uint16_t)hdr.size());
    npy += hdr;
    const std::string name = key + ".npy";
    uint32_t crc = crc32(0L, (const Bytef*)npy.data(), (uInt)npy.size());
    crc = crc32(crc, (const Bytef*)data, (uInt)nbytes);
    const uint32_t size = (uint32_t)(npy.size() + nbytes);
    Entry e{name, crc, size, (uint32_t)std::ftell(f)};
    std::string lh;
    le32(lh, 0x04034b50); le16(lh, 20); le16(lh, 0); le16(lh, 0); le16(lh, 0); le16(lh, 0x21);
    le32(lh, crc); le32(lh, size); le32(lh, size); le16(lh, (uint16_t)name.size()); le16(lh, 0);
    lh += name;
    std::fwrite(lh.data(), 1, lh.size(), f);
    std::fwrite(npy.data(), 1, npy.size(), f);
    std::fwrite(data, 1, nbytes, f);
    entries.push_back(e);
  }
  void close() {
    const uint32_t cd_off = (uint32_t)std::ftell(f);
    std::string cd;
    for (const auto& e : entries) {
      le32(cd, 0x02014b50); le16(cd, 20); le16(cd, 20); le16(cd, 0); le16(cd, 0); le16(cd, 0); le16(cd, 0x21);
      le32(cd, e.crc); le32(cd, e.size); le32(cd, e.size); le16(cd, (uint16_t)e.name.size()); le16(cd, 0); le16(cd, 0);
      le16(cd, 0); le16(cd, 0); le32(cd, 0); le32(cd, e.offset);
      cd += e.name;
    }
    std::string end;
    le32(end, 0x06054b50); le16(end, 0); le16(end, 0); le16(end, (uint16_t)entries.size()); le16(end, (uint16_t)entries.size());
    le32(end, (uint32_t)cd.size()); le32(end, cd_off); le16(end, 0);
    std::fwrite(cd.data(), 1, cd.size(), f);
    std::fwrite(end.data(), 1, end.size(), f);
    std::fclose(f);
    f = nullptr;
  }
};

void OutputFile::add_f64(const std::string& group, const std::string& name, const std::vector<size_t>& shape, const double* data) {
  size_t n = 1;
  for (size_t d : shape) n *= d;
  if (h5) h5->add_dataset(group, name, maghost::Hdf5Writer::F64, std::vector<uint64_t>(shape.begin(), shape.end()), data, quantity_attrs(name));
  else npz->add(group + "/" + name, "<f8", shape, data, n * 8);
}
void OutputFile::add_status(const std::vector<size_t>& shape, const char* data) {
  if (h5) h5->add_dataset("Log", "status", maghost::Hdf5Writer::CHAR1, std::vector<uint64_t>(shape.begin(), shape.end()), data, quantity_attrs("status"));
  else npz->add("Log/status", "|S1", shape, data, shape[0] * shape[1]);
}
void OutputFile::add_i32(const std::string& group, const std::string& name, const std::vector<size_t>& shape, const int32_t* data) {
  size_t n = 1;
  for (size_t d : shape) n *= d;
  if (h5) h5->add_dataset(group, name, maghost::Hdf5Writer::I32, std::vector<uint64_t>(shape.begin(), shape.end()), data);
  else npz->add(group + "/" + name, "<i4", shape, data, n * 4);
}
void OutputFile::close() {
  if (h5) h5->close();
  if (npz) npz->close();
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
      {"input_file_name", "\"" + cfg.input_file_name + "\""}, {"p_IO_separate_output", "T"}, {"p_IO_chunking", "F"}, {"p_IO_compression", "F"}}));
  w.add_root_attribute("grid_parameters", nml_text("GRID_PARAMETERS", {{"p_nx_ref", fint(p.p_nx_ref)}, {"p_nx_MAX", fint(p.p_nx_MAX)},
      {"p_rmax_over_rdisk", fnum(p.p_rmax_over_rdisk)}}));
  w.add_root_attribute("dynamo_parameters", nml_text("DYNAMO_PARAMETERS", {{"Dyn_quench", fbool(p.Dyn_quench)}, {"Alg_quench", fbool(p.Alg_quench)},
      {"lFloor", fbool(p.lFloor)}, {"Damp", fbool(p.Damp)}, {"frac_seed", fnum(p.frac_seed)}, {"p_seed_choice", "\"random\""},
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
  w.add_root_attribute("outflow_parameters", nml_text("OUTFLOW_PARAMETERS", {{"p_outflow_type", "\"no_outflow\""}}));
  w.add_root_attribute("File creation", "magnetizer_b200 host driver");
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <parameters.in> [gal_id ...] [--devices 0,1,...] [--chunk N] [--out file.hdf5|file.npz]\n"
                         "          [--restart] [--stop-after N] [--dry-run]\n", argv[0]);
    return 2;
  }
  try {
    maghost::RunConfig cfg = maghost::load_run_config(argv[1]);
    std::vector<int32_t> list, devices;
    int chunk = 2048;
    std::string out_path;
    bool dry_run = false, restart = false;
    long stop_after = -1;
    for (int i = 2; i < argc; i++) {
      const std::string a = argv[i];
      if (a == "--dry-run") { dry_run = true; continue; }
      if (a == "--restart") { restart = true; continue; }
      if (a == "--stop-after" && i + 1 < argc) { stop_after = std::atol(argv[++i]); continue; }
      if (a == "--devices" && i + 1 < argc) {
        for (char* tok = std::strtok(argv[++i], ","); tok; tok = std::strtok(nullptr, ",")) devices.push_back(std::atoi(tok));
      } else if (a == "--chunk" && i + 1 < argc) chunk = std::atoi(argv[++i]);
      else if (a == "--out" && i + 1 < argc) out_path = argv[++i];
      else list.push_back(std::atoi(a.c_str()));
    }
    if (cfg.input_file_name.empty()) throw std::runtime_error("input_file_name missing in io_parameters");
    const maghost::MagnetizerInput in = maghost::load_magnetizer_input(cfg.input_file_name);
    // jobs_reads_galaxy_list: ids are 1-based rows of the input file
    if (list.empty()) for (int g = 1; g <= in.ngal; g++) list.push_back(g);
    for (int g : list) if (g < 1 || g > in.ngal) throw std::runtime_error("galaxy id out of range: " + std::to_string(g));
    const int ngal = (int)list.size(), nz = in.nz, nr = cfg.params.p_nx_ref;
    std::vector<double> inputs((size_t)13 * ngal * nz);
    for (int c = 0; c < 13; c++)
      for (int k = 0; k < ngal; k++)
        std::memcpy(&inputs[((size_t)c * ngal + k) * nz], &in.inputs[((size_t)c * in.ngal + (list[k] - 1)) * nz], sizeof(double) * nz);
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
    // --restart ignores an existing file, --stop-after N ends this invocation after N galaxies
    // (the graceful stop of the STOP file / p_max_walltime, distributor.f90:299-312).
    if (out_path.empty()) out_path = cfg.output_file_name.empty() ? "magnetizer_b200_output.hdf5" : cfg.output_file_name;
    const bool as_npz = out_path.size() > 4 && out_path.compare(out_path.size() - 4, 4, ".npz") == 0;
    std::unique_ptr<maghost::H5Reader> prev;
    std::vector<char> completed(ngal, 0);
    if (!restart && !as_npz) {
      if (FILE* t = std::fopen(out_path.c_str(), "rb")) {
        std::fclose(t);
        prev.reset(new maghost::H5Reader(out_path));
        const maghost::H5Reader::Data ids = prev->read("Log/gal_id"), comp = prev->read("Log/completed");
        bool same = ids.v.size() == (size_t)ngal && comp.v.size() == (size_t)ngal;
        for (int g = 0; same && g < ngal; g++) same = (int32_t)ids.v[g] == list[g];
        if (!same) throw std::runtime_error("cannot resume: " + out_path + " holds another galaxy list (use --restart)");
        for (int g = 0; g < ngal; g++) completed[g] = comp.v[g] > 0.5;
      }
    }
    std::vector<int> todo;
    for (int g = 0; g < ngal; g++) if (!completed[g]) todo.push_back(g);
    const size_t n_open = todo.size();
    if (stop_after >= 0 && todo.size() > (size_t)stop_after) todo.resize((size_t)stop_after);
    if (dry_run) std::printf("resume %s: %zu of %d galaxies completed, %zu to solve now\n", prev ? "yes" : "no", ngal - n_open, ngal, todo.size());
    if (dry_run) {  // host-side logic only (parameter file, input file, galaxy list, quantity list): no device needed
      std::printf("dry-run ngal %d nz %d list %zu p_courant_v %.17g p_nx_ref %d p_random_seed %d\n", in.ngal, nz, list.size(),
                  cfg.params.p_courant_v, cfg.params.p_nx_ref, cfg.params.p_random_seed);
      for (int c = 0; c < 13; c++) {
        double sum = 0;
        for (size_t k = 0; k < (size_t)ngal * nz; k++) sum += inputs[(size_t)c * ngal * nz + k];
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
    const int nsel = (int)todo.size();
    std::vector<int32_t> sel_ids(nsel);
    std::vector<double> sel_inputs((size_t)13 * nsel * nz);
    for (int k = 0; k < nsel; k++) sel_ids[k] = list[todo[k]];
    for (int c = 0; c < 13; c++)
      for (int k = 0; k < nsel; k++)
        std::memcpy(&sel_inputs[((size_t)c * nsel + k) * nz], &inputs[((size_t)c * ngal + todo[k]) * nz], sizeof(double) * nz);
    std::vector<double> profiles((size_t)pq.size() * nsel * nz * nr), scalars((size_t)sq.size() * nsel * nz);
    std::vector<char> status((size_t)nsel * nz);
    std::vector<int32_t> nsteps(nsel), error(nsel), done(devices.size()), stolen(devices.size());
    std::vector<int64_t> stages(nsel);
    std::vector<uint32_t> flags(nsel);
    mag_outputs_t out{profiles.data(), scalars.empty() ? nullptr : scalars.data(), status.data(), nsteps.data(), error.data(),
                      stages.data(), flags.data()};
    std::printf("Magnetizer (B200 path): '%s'  %d galaxies x %d snapshots (%zu already completed, %d to solve), %zu GPU(s), chunk %d\n",
                cfg.model_name.c_str(), ngal, nz, ngal - n_open, nsel, devices.size(), chunk);
    const auto t0 = std::chrono::steady_clock::now();
    if (nsel > 0) {
      const int rc = mag_solve_partitioned(&cfg.params, (int)devices.size(), devices.data(), nsel, sel_ids.data(), nz, in.t_gyr.data(),
                                           sel_inputs.data(), (int)pq.size(), pq.data(), (int)sq.size(), sq.data(), &out, chunk,
                                           done.data(), stolen.data());
      if (rc != MAG_OK) throw std::runtime_error(std::string("solve failed (") + std::to_string(rc) + "): " + mag_partition_last_error());
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // ---- write_output (output.f90:23-236): unit conversion, [ngal][nx][nz] layout, Log group.
    // Galaxies completed by an earlier invocation keep their rows, galaxies not solved yet hold
    // "never written" markers and completed = 0.
    const std::string tmp_path = out_path + ".tmp";
    OutputFile w;
    if (as_npz) {
      w.npz.reset(new NpzWriter(tmp_path));
      if (!w.npz->ok()) throw std::runtime_error("cannot write " + tmp_path);
    } else {
      w.h5.reset(new maghost::Hdf5Writer(tmp_path));
      w.h5->add_group("Output");
      w.h5->add_group("Log");
      write_parameter_attributes(*w.h5, cfg);
    }
    auto previous = [&](const std::string& path, size_t n) {   // dataset of the resumed file or "never written"
      if (!prev) return std::vector<double>(n, MAG_INVALID);
      maghost::H5Reader::Data d = prev->read(path);
      if (d.v.size() != n) throw std::runtime_error("cannot resume: " + path + " has another shape (use --restart)");
      return d.v;
    };
    long long total_stages = 0;
    for (auto s : stages) total_stages += s;
    for (size_t q = 0; q < pq.size(); q++) {
      const std::string name = kProfileNames[pq[q]];
      const double fac = unit_factor(name);
      const double* src = &profiles[q * (size_t)nsel * nz * nr];
      std::vector<double> full = previous("Output/" + std::string(name == "rkpc" && !have_rkpc ? "r" : name), (size_t)ngal * nr * nz);
      for (int k = 0; k < nsel; k++)
        for (int iz = 0; iz < nz; iz++)
          for (int i = 0; i < nr; i++) {
            const double v = src[((size_t)k * nz + iz) * nr + i];
            full[((size_t)todo[k] * nr + i) * nz + iz] = (v == MAG_INVALID) ? v : v * fac;
          }
      if (name != "rkpc" || have_rkpc) w.add_f64("Output", name, {(size_t)ngal, (size_t)nr, (size_t)nz}, full.data());
      if (name == "rkpc") w.add_f64("Output", "r", {(size_t)ngal, (size_t)nr, (size_t)nz}, full.data());
    }
    for (size_t q = 0; q < sq.size(); q++) {
      std::vector<double> full = previous(std::string("Output/") + kScalarNames[sq[q]], (size_t)ngal * nz);
      for (int k = 0; k < nsel; k++)
        std::memcpy(&full[(size_t)todo[k] * nz], &scalars[(q * (size_t)nsel + k) * nz], sizeof(double) * nz);
      w.add_f64("Output", kScalarNames[sq[q]], {(size_t)ngal, (size_t)nz}, full.data());
    }
    std::vector<char> status_full((size_t)ngal * nz, '-');
    if (prev) {
      const std::vector<double> st = previous("Log/status", (size_t)ngal * nz);
      for (size_t i = 0; i < st.size(); i++) status_full[i] = (char)(int)st[i];
    }
    for (int k = 0; k < nsel; k++) std::memcpy(&status_full[(size_t)todo[k] * nz], &status[(size_t)k * nz], nz);
    w.add_status({(size_t)ngal, (size_t)nz}, status_full.data());
    std::vector<double> log_n = prev ? previous("Log/nsteps", ngal) : std::vector<double>(ngal, 0.0);
    std::vector<double> log_c = prev ? previous("Log/completed", ngal) : std::vector<double>(ngal, 0.0);
    std::vector<double> log_rt = prev ? previous("Log/runtime", ngal) : std::vector<double>(ngal, 0.0);
    std::map<char, long> counts;
    for (int k = 0; k < nsel; k++) {
      const int g = todo[k];
      log_n[g] = nsteps[k];
      log_c[g] = error[k] ? 0.0 : 1.0;  // Log/completed is only written when runtime > 0 (distributor.f90:363)
      log_rt[g] = total_stages > 0 ? secs * (double)stages[k] / (double)total_stages : 0.0;
    }
    for (char c : status_full) counts[c]++;
    // Log datasets are (ngals, 1) in the reference's file (Fortran (1, ngals), output.f90:37-54)
    const std::vector<size_t> log_shape = as_npz ? std::vector<size_t>{(size_t)ngal} : std::vector<size_t>{(size_t)ngal, 1};
    w.add_f64("Log", "nsteps", log_shape, log_n.data());
    w.add_f64("Log", "completed", log_shape, log_c.data());
    w.add_f64("Log", "runtime", log_shape, log_rt.data());
    w.add_i32("Log", "gal_id", {(size_t)ngal}, list.data());
    if (as_npz) w.add_f64("Input", "t", {(size_t)nz}, in.t_gyr.data());
    w.close();
    prev.reset();
    if (std::rename(tmp_path.c_str(), out_path.c_str()) != 0) throw std::runtime_error("cannot rename " + tmp_path + " to " + out_path);
    std::printf("  solved in %.3f s  (%.1f galaxies/s, %.3e point-stages/s)\n", secs, nsel / secs, total_stages / secs);
    for (size_t d = 0; d < devices.size(); d++)
      std::printf("  GPU %d: %d chunks (%d stolen)\n", devices[d], done[d], stolen[d]);
    std::printf("  status codes:");
    for (auto& kv : counts) std::printf(" '%c' x%ld", kv.first, kv.second);
    std::printf("\n  wrote %s\n", out_path.c_str());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "Magnetizer_b200: %s\n", e.what());
    return 1;
  }
}
