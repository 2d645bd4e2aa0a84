// Test helper (tests/test_hdf5_writer.py): writes a small Magnetizer-style output file with
// mag_hdf5_writer.hpp and reads two datasets back with the C++ reader of the host driver.
//   hdf5_writer_probe <out.hdf5> <ngal> <nr> <nz>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../magnetizer_b200/host/mag_hdf5_reader.hpp"
#include "../magnetizer_b200/host/mag_hdf5_writer.hpp"

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const int ngal = std::atoi(argv[2]), nr = std::atoi(argv[3]), nz = std::atoi(argv[4]);
  try {
    {
      maghost::Hdf5Writer w(argv[1]);
      w.add_group("Output");
      w.add_group("Log");
      w.add_root_attribute("Model name", "probe model");
      w.add_root_attribute("run_parameters", "&RUN_PARAMETERS  NSTEPS_0=1000, P_COURANT_V=0.09000000357627869, P_VARIABLE_TIMESTEPS=T, /");
      std::vector<double> prof((size_t)ngal * nr * nz), scal((size_t)ngal * nz), one(ngal);
      std::vector<char> status((size_t)ngal * nz);
      std::vector<int32_t> ids(ngal);
      // more datasets than one would fit in a default-sized symbol-table node, names out of order
      const char* names[] = {"r", "Br", "Bp", "alp_m", "Bzmod", "h", "Omega", "Shear", "l", "v", "etat", "tau", "alp_k", "Uz",
                             "n", "Beq", "alp", "Omega_h", "Omega_b", "Omega_d", "rkpc"};
      int q = 0;
      for (const char* nm : names) {
        for (size_t i = 0; i < prof.size(); i++) prof[i] = 1000.0 * q + 0.5 * (double)i;
        w.add_dataset("Output", nm, maghost::Hdf5Writer::F64, {(uint64_t)ngal, (uint64_t)nr, (uint64_t)nz}, prof.data(),
                      {{"Units", "microgauss"}, {"Description", std::string("profile ") + nm}});
        q++;
      }
      for (size_t i = 0; i < scal.size(); i++) scal[i] = -99999.0 + (double)i;
      w.add_dataset("Output", "Bmax", maghost::Hdf5Writer::F64, {(uint64_t)ngal, (uint64_t)nz}, scal.data(), {{"Units", "microgauss"}});
      for (size_t i = 0; i < status.size(); i++) status[i] = "Mm-i"[i % 4];
      w.add_dataset("Log", "status", maghost::Hdf5Writer::CHAR1, {(uint64_t)ngal, (uint64_t)nz}, status.data(), {{"Description", "Status codes."}});
      for (int g = 0; g < ngal; g++) { one[g] = g + 0.25; ids[g] = g + 1; }
      w.add_dataset("Log", "runtime", maghost::Hdf5Writer::F64, {(uint64_t)ngal, 1}, one.data(), {{"Units", "s"}});
      w.add_dataset("Log", "gal_id", maghost::Hdf5Writer::I32, {(uint64_t)ngal}, ids.data());
      for (int g = 0; g < ngal; g++) one[g] = (g % 3 == 0) ? 1.0 : 0.0;
      w.add_dataset("Log", "completed", maghost::Hdf5Writer::F64, {(uint64_t)ngal, 1}, one.data());
      w.close();
    }
    maghost::H5Reader r(argv[1]);
    const maghost::H5Reader::Data a = r.read("Output/Omega"), b = r.read("Log/runtime"), c = r.read("Log/gal_id");
    std::printf("Omega shape %llu %llu %llu first %.17g last %.17g\n", (unsigned long long)a.shape[0], (unsigned long long)a.shape[1],
                (unsigned long long)a.shape[2], a.v.front(), a.v.back());
    std::printf("runtime n %zu last %.17g\n", b.v.size(), b.v.back());
    std::printf("gal_id n %zu last %.17g\n", c.v.size(), c.v.back());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "probe failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
