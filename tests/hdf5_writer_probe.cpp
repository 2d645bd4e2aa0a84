// Test helper (tests/test_hdf5_writer.py): writes a Magnetizer-style output file with
// mag_hdf5_writer.hpp the way the host driver does -- datasets sized to ALL galaxies, written in
// blocks of `cs` galaxies, flushed in between, a second invocation extending the flushed file --
// and reads parts of it back with the C++ reader of the driver.
//   hdf5_writer_probe write1 <out.hdf5> <ngal> <nr> <nz> <gals_per_chunk> <mode>
//       blocks 0 and 2, flush, block 3 WITHOUT flush, then exit without closing (a killed run)
//   hdf5_writer_probe resume <out.hdf5> <ngal> <nr> <nz> <gals_per_chunk> <mode>
//       extends the file: adopts its chunks, writes every block whose galaxies are not completed
//   hdf5_writer_probe rows <out.hdf5> <dataset> <g0> <n>      prints rows through H5Reader::read_rows
//     mode: chunked | raw (chunked, no deflate) | contiguous
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../magnetizer_b200/host/mag_hdf5_reader.hpp"
#include "../magnetizer_b200/host/mag_hdf5_writer.hpp"

using maghost::Hdf5Writer;

static double pval(int q, size_t g, size_t i, size_t iz, size_t nr, size_t nz) { return 1000.0 * q + 0.5 * (double)((g * nr + i) * nz + iz); }

int main(int argc, char** argv) {
  if (argc >= 6 && std::string(argv[1]) == "rows") {
    try {
      maghost::H5Reader r(argv[2]);
      const maghost::H5Reader::Data d = r.read_rows(argv[3], std::strtoull(argv[4], nullptr, 10), std::strtoull(argv[5], nullptr, 10));
      std::printf("shape");
      for (auto s : d.shape) std::printf(" %llu", (unsigned long long)s);
      std::printf("\n");
      for (double v : d.v) std::printf("%.17g\n", v);
      return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "probe failed: %s\n", e.what()); return 1; }
  }
  if (argc < 8) return 2;
  const std::string path = argv[2], mode = argv[7];
  const size_t ngal = std::atoi(argv[3]), nr = std::atoi(argv[4]), nz = std::atoi(argv[5]);
  Hdf5Writer::Options opt;
  opt.gals_per_chunk = std::atoi(argv[6]);
  opt.chunking = mode != "contiguous";
  opt.compression = mode == "chunked";
  opt.threads = 3;
  const char* names[] = {"r", "Br", "Bp", "alp_m", "Bzmod", "h", "Omega", "Shear", "l", "v", "etat", "tau", "alp_k", "Uz",
                         "n", "Beq", "alp", "Omega_h", "Omega_b", "Omega_d", "rkpc"};
  const int nq = 21;
  try {
    std::vector<int> hp(nq);
    int h_bmax, h_status, h_runtime, h_completed, h_gal;
    auto declare = [&](Hdf5Writer& w) {
      w.add_group("Output");
      w.add_group("Log");
      w.add_root_attribute("Model name", "probe model");
      w.add_root_attribute("run_parameters", "&RUN_PARAMETERS  NSTEPS_0=1000, P_COURANT_V=0.09000000357627869, P_VARIABLE_TIMESTEPS=T, /");
      for (int q = 0; q < nq; q++)
        hp[q] = w.create_dataset("Output", names[q], Hdf5Writer::F64, {ngal, nr, nz}, {{"Units", "microgauss"}, {"Description", std::string("profile ") + names[q]}});
      h_bmax = w.create_dataset("Output", "Bmax", Hdf5Writer::F64, {ngal, nz}, {{"Units", "microgauss"}});
      h_status = w.create_dataset("Log", "status", Hdf5Writer::CHAR1, {ngal, nz}, {{"Description", "Status codes."}}, false);
      h_runtime = w.create_dataset("Log", "runtime", Hdf5Writer::F64, {ngal, 1}, {{"Units", "s"}});
      h_completed = w.create_dataset("Log", "completed", Hdf5Writer::F64, {ngal, 1});
      h_gal = w.create_dataset("Log", "gal_id", Hdf5Writer::I32, {ngal}, {}, false);
    };
    auto write_block = [&](Hdf5Writer& w, size_t kb) {
      const size_t cs = w.gals_per_block(hp[0]), g0 = kb * cs, nb = std::min(cs, ngal - g0);
      std::vector<double> prof(nb * nr * nz), scal(nb * nz), one(nb);
      std::vector<char> status(nb * nz);
      std::vector<int32_t> ids(nb);
      for (int q = 0; q < nq; q++) {
        for (size_t g = 0; g < nb; g++) for (size_t i = 0; i < nr; i++) for (size_t iz = 0; iz < nz; iz++)
          prof[(g * nr + i) * nz + iz] = pval(q, g0 + g, i, iz, nr, nz);
        w.write_block(hp[q], kb, prof.data());
      }
      for (size_t g = 0; g < nb; g++) for (size_t iz = 0; iz < nz; iz++) {
        scal[g * nz + iz] = -99999.0 + (double)((g0 + g) * nz + iz);
        status[g * nz + iz] = "Mm-i"[((g0 + g) * nz + iz) % 4];
      }
      w.write_block(h_bmax, kb, scal.data());
      w.write_block(h_status, kb, status.data());
      for (size_t g = 0; g < nb; g++) { one[g] = (double)(g0 + g) + 0.25; ids[g] = (int32_t)(g0 + g + 1); }
      w.write_block(h_runtime, kb, one.data());
      w.write_block(h_gal, kb, ids.data());
      for (size_t g = 0; g < nb; g++) one[g] = 1.0;
      w.write_block(h_completed, kb, one.data());
    };
    const std::string cmd = argv[1];
    if (cmd == "write1") {
      Hdf5Writer w(path, opt);
      declare(w);
      const size_t nblocks = w.n_blocks(hp[0]);
      write_block(w, 0);
      if (nblocks > 2) write_block(w, 2);
      w.flush();
      if (nblocks > 3) write_block(w, 3);   // never flushed: must not be visible
      std::printf("blocks %zu bytes %llu\n", nblocks, (unsigned long long)w.bytes_written());
      std::fflush(stdout);
      std::_Exit(0);                         // no close(): the process dies here
    }
    if (cmd == "resume") {
      std::vector<char> done;
      std::vector<std::pair<std::string, std::vector<maghost::H5ChunkRec>>> old;
      std::vector<uint64_t> old_addr;
      {
        maghost::H5Reader r(path);
        const maghost::H5Reader::Data c = r.read("Log/completed");
        for (double v : c.v) done.push_back(v > 0.5);
        for (int q = 0; q < nq; q++) old.push_back({std::string("Output/") + names[q], r.chunks(std::string("Output/") + names[q])});
        for (const char* nm : {"Output/Bmax", "Log/status", "Log/runtime", "Log/completed", "Log/gal_id"}) old.push_back({nm, r.chunks(nm)});
        for (auto& kv : old) old_addr.push_back(r.contiguous_address(kv.first));
      }
      Hdf5Writer w(path, opt, /*extend=*/true);
      declare(w);
      for (auto& kv : old) {
        const size_t slash = kv.first.find('/');
        const int h = w.find_dataset(kv.first.substr(0, slash), kv.first.substr(slash + 1));
        if (h < 0) throw std::runtime_error("dataset vanished: " + kv.first);
        w.adopt(h, kv.second);
        w.adopt_contiguous(h, old_addr[&kv - &old[0]]);
      }
      const size_t cs = w.gals_per_block(hp[0]);
      std::printf("resume: completed");
      for (size_t kb = 0; kb < w.n_blocks(hp[0]); kb++) std::printf(" %d", (int)done[kb * cs]);
      std::printf("\n");
      for (size_t kb = 0; kb < w.n_blocks(hp[0]); kb++) if (!done[kb * cs]) write_block(w, kb);
      w.close();
      return 0;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "probe failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
