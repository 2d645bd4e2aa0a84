// mag_hdf5_writer.hpp -- native HDF5 writer for the Magnetizer OUTPUT file in the reference's
// layout (no libhdf5 in this environment).
//
// What the reference's output file is (source/IO_hdf5.f90, source/output.f90; consumer
// python/magnetizer/MagnetizerRun.py:117-150, python/magnetizer/parameters.py):
//   /Output/<quantity>  float64 (gals_number, nx_ref, nz) profiles;  (gals_number, nz) scalars
//   /Log/status         1-character strings (gals_number, nz);  /Log/{runtime,nsteps,completed} float64 (gals_number, 1)
//   every dataset has gals_number rows -- ALL galaxies of the input file, also in single-galaxy
//   mode (IO_hdf5.f90:463,551) -- is chunked (1000 galaxies x all radii x ONE snapshot,
//   IO_hdf5.f90:940-952) and deflate-compressed (level 6, :964); doubles have the fill value
//   -1d50 (:44,979), which is what a reader sees for galaxies that were never written
//   dataset attributes  'Units', 'Description'  (arrays of one fixed-length string)
//   root attributes     'Model name', '<namelist>_parameters' (the namelist as one line of text), 'File creation'
//   the file doubles as checkpoint: it is flushed every p_ncheckpoint galaxies (IO_hdf5.f90:721-731)
//
// Encoding: the oldest, most widely readable one (mirror image of mag_hdf5_reader.hpp):
//   superblock version 0, 8-byte offsets and lengths
//   groups   = version-1 object header + symbol-table message -> B-tree v1 node (type 0) ->
//              one symbol-table node (SNOD) + local heap with the link names
//   datasets = version-1 object header: dataspace (v1), datatype (v1), fill value (v2), filter
//              pipeline (v1, deflate), CHUNKED layout (v3) -> B-tree v1 (type 1: raw data chunks;
//              leaf and internal nodes of 2K = 64 entries, sibling pointers, right-most key one
//              chunk beyond the last), fixed-length string attributes (attribute message v1)
//   (p_IO_chunking = .false.: CONTIGUOUS layout, every element written, no filter)
//
// The file is LOG-STRUCTURED: chunk data is appended as blocks of galaxies are written; flush()
// appends a complete, fresh set of metadata (chunk B-trees, object headers, groups) and then
// rewrites the 96-byte superblock at offset 0 to point at it.  After every flush the file is a
// valid HDF5 file holding everything written so far (older metadata is dead space), so a run
// that is stopped or killed between flushes resumes from the last one -- the reference's
// checkpoint semantics -- and an existing file is extended in place (adopt()) instead of being
// copied.  Only metadata and one block of one dataset are ever held in memory: the 1e6-galaxy
// gather of BASELINE config 4 streams.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <zlib.h>

namespace maghost {

struct H5ChunkRec {                 // one stored chunk of a dataset
  std::vector<uint64_t> offs;       // element offsets of the chunk's corner, one per dataset dimension
  uint64_t addr = 0;
  uint32_t size = 0, mask = 0;
};

class Hdf5Writer {
 public:
  enum Type { F64, I32, CHAR1 };
  typedef std::vector<std::pair<std::string, std::string>> TextAttrs;
  struct Options {
    bool chunking = true;           // p_IO_chunking
    int gals_per_chunk = 1000;      // p_IO_number_of_galaxies_in_chunks
    bool compression = true;        // p_IO_compression (requires chunking)
    int level = 6;                  // p_IO_compression_level
    int threads = 0;                // deflate threads (0: hardware concurrency, at most 32)
  };
  static constexpr double kFill = -1e50;   // INVALID, IO_hdf5.f90:44

  // Creates `path`, or -- extend = true and the file exists -- opens it for appending (the caller
  // then hands the existing chunk lists over with adopt()).
  Hdf5Writer(const std::string& path, const Options& opt, bool extend = false) : opt_(opt), pos_(0) {
    if (!opt_.chunking) opt_.compression = false;
    if (opt_.gals_per_chunk < 1) opt_.gals_per_chunk = 1;
    f_ = extend ? std::fopen(path.c_str(), "r+b") : nullptr;
    if (f_) {
      std::fseek(f_, 0, SEEK_END);
      pos_ = (uint64_t)std::ftell(f_);
      if (pos_ < kDataStart) { std::fclose(f_); f_ = nullptr; }
    }
    if (!f_) {
      f_ = std::fopen(path.c_str(), "w+b");
      if (!f_) throw std::runtime_error("cannot create " + path);
      pos_ = 0;
      put(std::string(kDataStart, '\0'));   // the superblock (96 bytes) is written by flush()
    }
  }
  ~Hdf5Writer() { if (f_) std::fclose(f_); }

  void add_group(const std::string& name) {
    for (auto& g : groups_) if (g.name == name) return;
    groups_.push_back(Group{name, {}});
  }

  // dims in C order (slowest first); dims[0] is the number of galaxies.  Returns a handle.
  int create_dataset(const std::string& group, const std::string& name, Type t, const std::vector<uint64_t>& dims,
                     const TextAttrs& attrs = TextAttrs(), bool has_fill = true, double fill = kFill) {
    Group* g = nullptr;
    for (auto& x : groups_) if (x.name == group) g = &x;
    if (!g) throw std::runtime_error("unknown group " + group);
    if (dims.empty() || dims.size() > 3) throw std::runtime_error("dataset rank must be 1..3: " + name);
    Dset d;
    d.name = name; d.type = t; d.dims = dims; d.attrs = attrs; d.has_fill = has_fill && t == F64; d.fill = fill;
    d.row = 1;
    for (size_t k = 1; k < dims.size(); k++) d.row *= dims[k];
    d.cs = std::min<uint64_t>(dims[0], (uint64_t)opt_.gals_per_chunk);
    if (d.cs == 0) d.cs = 1;
    // chunk = (min(ngals, gals_per_chunk), dims[1], 1): all radii of ONE snapshot (IO_hdf5.f90:940-952)
    d.cdims.assign(dims.size(), 1);
    d.cdims[0] = d.cs;
    if (dims.size() == 3) d.cdims[1] = dims[1];
    if (!opt_.chunking) {   // contiguous: allocate and fill now
      align8();
      d.addr = pos_;
      const uint64_t n = dims[0] * d.row, es = elem_size(t);
      std::vector<char> blk((size_t)std::min<uint64_t>(n, 1u << 16) * es);
      fill_buffer(d, blk.data(), blk.size() / es);
      for (uint64_t done = 0; done < n;) {
        const uint64_t m = std::min<uint64_t>(n - done, blk.size() / es);
        put_raw(blk.data(), m * es);
        done += m;
      }
    }
    dsets_.push_back(d);
    g->members.push_back((int)dsets_.size() - 1);
    return (int)dsets_.size() - 1;
  }
  int find_dataset(const std::string& group, const std::string& name) const {
    for (auto& g : groups_) if (g.name == group) for (int h : g.members) if (dsets_[h].name == name) return h;
    return -1;
  }
  uint64_t gals_per_block(int h) const { return dsets_[h].cs; }
  uint64_t n_blocks(int h) const { return (dsets_[h].dims[0] + dsets_[h].cs - 1) / dsets_[h].cs; }

  // Existing chunks of this dataset in a file opened with extend = true (mag_hdf5_reader.hpp lists them).
  void adopt(int h, const std::vector<H5ChunkRec>& recs) {
    for (const auto& r : recs) dsets_[h].chunks[r.offs] = r;
  }
  // ... and the storage of a contiguous one (p_IO_chunking = .false.): blocks are rewritten in place
  void adopt_contiguous(int h, uint64_t addr) { if (!opt_.chunking && addr != kUndef) dsets_[h].addr = addr; }

  // Block kb = galaxies [kb*cs, min((kb+1)*cs, ngals)) of dataset h; `data` holds them in the dataset's
  // own (C) layout, [nb][dims[1]][dims[2]].  Writing a block again replaces it.
  void write_block(int h, uint64_t kb, const void* data) {
    Dset& d = dsets_[h];
    const uint64_t g0 = kb * d.cs;
    if (g0 >= d.dims[0]) throw std::runtime_error("block out of range: " + d.name);
    const uint64_t nb = std::min<uint64_t>(d.cs, d.dims[0] - g0), es = elem_size(d.type);
    if (!opt_.chunking) {
      std::fseek(f_, (long)(d.addr + g0 * d.row * es), SEEK_SET);
      if (std::fwrite(data, 1, (size_t)(nb * d.row * es), f_) != nb * d.row * es) throw std::runtime_error("write failed");
      std::fseek(f_, (long)pos_, SEEK_SET);
      return;
    }
    // one chunk per index of the LAST dimension (rank 3: per snapshot; rank 2: per column; rank 1: one)
    const size_t rank = d.dims.size();
    const uint64_t nlast = rank == 1 ? 1 : d.dims[rank - 1];
    const uint64_t mid = rank == 3 ? d.dims[1] : 1;
    std::vector<ChunkJob> jobs((size_t)nlast);
    for (uint64_t il = 0; il < nlast; il++) {
      jobs[(size_t)il].h = h; jobs[(size_t)il].kb = kb; jobs[(size_t)il].il = il;
      jobs[(size_t)il].fill = [=](char* raw, uint64_t) {
        const char* src = (const char*)data;
        for (uint64_t g = 0; g < nb; g++)
          for (uint64_t i = 0; i < mid; i++)
            std::memcpy(raw + (g * mid + i) * es, src + ((g * mid + i) * nlast + il) * es, (size_t)es);
      };
    }
    write_chunks(jobs);
  }

  // One chunk of a chunked dataset: block kb (galaxies [kb*cs, ...)), index `il` of the last dimension.
  // `fill` receives the chunk's buffer -- [cs][dims[1]] elements for a rank-3 dataset, [cs] otherwise --
  // already holding the fill value, and the number of galaxies of the block that exist; it stores the
  // rows it has.  Called on a worker thread.
  struct ChunkJob {
    int h = -1;
    uint64_t kb = 0, il = 0;
    std::function<void(char* raw, uint64_t nb)> fill;
  };
  // Builds (fill + deflate) the chunks on the writer's threads and appends them in the order of `jobs`;
  // at most a window of encoded chunks is held in memory.  Writing a chunk again replaces it.
  void write_chunks(const std::vector<ChunkJob>& jobs) {
    if (!opt_.chunking) throw std::runtime_error("write_chunks needs a chunked file");
    const size_t n = jobs.size();
    if (n == 0) return;
    uint64_t raw_total = 0;
    for (const ChunkJob& j : jobs) {
      if (j.h < 0 || j.h >= (int)dsets_.size()) throw std::runtime_error("bad dataset handle");
      const Dset& d = dsets_[j.h];
      const uint64_t nlast = d.dims.size() == 1 ? 1 : d.dims.back();
      if (j.kb * d.cs >= d.dims[0] || j.il >= nlast) throw std::runtime_error("chunk out of range: " + d.name);
      raw_total += d.cs * (d.dims.size() == 3 ? d.dims[1] : 1) * elem_size(d.type);
    }
    auto encode = [&](const ChunkJob& j, std::vector<char>& raw, std::string& out) {
      const Dset& d = dsets_[j.h];
      const uint64_t es = elem_size(d.type), celems = d.cs * (d.dims.size() == 3 ? d.dims[1] : 1);
      raw.resize((size_t)(celems * es));
      fill_buffer(d, raw.data(), celems);   // rows of an edge chunk beyond the last galaxy hold the fill value
      j.fill(raw.data(), std::min<uint64_t>(d.cs, d.dims[0] - j.kb * d.cs));
      if (opt_.compression) {
        uLongf cap = compressBound((uLong)raw.size());
        out.assign((size_t)cap, '\0');
        if (compress2((Bytef*)&out[0], &cap, (const Bytef*)raw.data(), (uLong)raw.size(), opt_.level) != Z_OK) throw std::runtime_error("deflate failed");
        out.resize((size_t)cap);
      } else {
        out.assign(raw.data(), raw.size());
      }
    };
    auto append = [&](const ChunkJob& j, const std::string& z) {
      Dset& d = dsets_[j.h];
      const size_t rank = d.dims.size();
      H5ChunkRec r;
      r.offs.assign(rank, 0);
      r.offs[0] = j.kb * d.cs;
      if (rank >= 2) r.offs[rank - 1] = j.il;
      r.addr = pos_; r.size = (uint32_t)z.size(); r.mask = 0;
      put_raw(z.data(), z.size());
      d.chunks[r.offs] = r;
    };
    int nt = opt_.threads > 0 ? opt_.threads : (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    if ((size_t)nt > n) nt = (int)n;
    if (nt <= 1 || raw_total < (1u << 20)) {          // small: not worth a thread
      std::vector<char> raw; std::string z;
      for (const ChunkJob& j : jobs) { encode(j, raw, z); append(j, z); }
      return;
    }
    std::vector<std::string> enc(n);
    std::vector<char> ready(n, 0);
    std::mutex mu;
    std::condition_variable cv_ready, cv_room;
    size_t next = 0, written = 0;
    const size_t window = 4 * (size_t)nt;
    std::string err;
    auto worker = [&] {
      std::vector<char> raw;
      for (;;) {
        size_t i;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv_room.wait(lk, [&] { return next >= n || next < written + window || !err.empty(); });
          if (next >= n || !err.empty()) return;
          i = next++;
        }
        std::string e;
        try { encode(jobs[i], raw, enc[i]); } catch (const std::exception& ex) { e = ex.what(); if (e.empty()) e = "chunk encoding failed"; }
        {
          std::lock_guard<std::mutex> lk(mu);
          if (!e.empty() && err.empty()) err = e;
          ready[i] = 1;
        }
        cv_ready.notify_all();
        if (!e.empty()) { cv_room.notify_all(); return; }
      }
    };
    std::vector<std::thread> th;
    for (int k = 0; k < nt; k++) th.emplace_back(worker);
    std::string werr;
    for (size_t i = 0; i < n; i++) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_ready.wait(lk, [&] { return ready[i] || !err.empty(); });
        if (!err.empty()) break;
      }
      try { append(jobs[i], enc[i]); } catch (const std::exception& ex) { werr = ex.what(); }
      std::string().swap(enc[i]);
      {
        std::lock_guard<std::mutex> lk(mu);
        written = i + 1;
        if (!werr.empty() && err.empty()) err = werr;
      }
      cv_room.notify_all();
      if (!werr.empty()) break;
    }
    for (auto& t : th) t.join();
    if (!err.empty()) throw std::runtime_error(err);
  }

  void add_root_attribute(const std::string& name, const std::string& text) {
    for (auto& a : root_attrs_) if (a.first == name) { a.second = text; return; }
    root_attrs_.push_back({name, text});
  }

  // Appends a complete set of metadata and points the superblock at it: the file is valid and
  // holds everything written so far (H5Fflush, IO_hdf5.f90:721-731).
  void flush() {
    if (!f_) return;
    std::vector<Link> root_links;
    for (auto& g : groups_) {
      std::vector<Link> links;
      for (int h : g.members) {
        Dset& d = dsets_[h];
        const uint64_t bt = opt_.chunking ? write_chunk_btree(d) : 0;
        align8();
        d.ohdr = pos_;
        put(dataset_header(d, bt));
        links.push_back(Link{d.name, d.ohdr, 0, 0, 0});
      }
      uint64_t bt, hp;
      const uint64_t oh = write_group(links, TextAttrs(), &bt, &hp);
      root_links.push_back(Link{g.name, oh, 1, bt, hp});
    }
    uint64_t rbt, rhp;
    const uint64_t root = write_group(root_links, root_attrs_, &rbt, &rhp);
    align8();
    const uint64_t eof = pos_;
    // ---- superblock, version 0
    std::string sb("\x89HDF\r\n\x1a\n", 8);
    const unsigned char ver[8] = {0, 0, 0, 0, 0, 8, 8, 0};   // superblock, free space, root entry, -, shared header, offsets, lengths, -
    sb.append((const char*)ver, 8);
    u16(sb, kLeafK); u16(sb, kInternalK);
    u32(sb, 0);                                      // file consistency flags
    u64(sb, 0); u64(sb, kUndef); u64(sb, eof); u64(sb, kUndef);   // base, free-space info, end of file, driver info
    u64(sb, 0); u64(sb, root); u32(sb, 1); u32(sb, 0); u64(sb, rbt); u64(sb, rhp);   // root symbol-table entry
    std::fflush(f_);                                 // data and metadata first, then the pointer to them
    std::fseek(f_, 0, SEEK_SET);
    if (std::fwrite(sb.data(), 1, sb.size(), f_) != sb.size()) throw std::runtime_error("write failed");
    std::fflush(f_);
    std::fseek(f_, (long)pos_, SEEK_SET);
  }

  void close() {
    if (!f_) return;
    flush();
    std::fclose(f_);
    f_ = nullptr;
  }
  uint64_t bytes_written() const { return pos_; }

 private:
  static constexpr uint64_t kUndef = 0xFFFFFFFFFFFFFFFFull;
  static constexpr unsigned kLeafK = 32, kInternalK = 16;   // group symbol tables: up to 64 links in one node
  static constexpr unsigned kChunkK = 32;                   // chunk B-trees: libhdf5's default for version-0 superblocks
  static constexpr size_t kDataStart = 2048;
  struct Dset {
    std::string name; Type type; std::vector<uint64_t> dims, cdims; TextAttrs attrs;
    bool has_fill = false; double fill = 0;
    uint64_t row = 1, cs = 1, addr = 0, ohdr = 0;
    std::map<std::vector<uint64_t>, H5ChunkRec> chunks;     // ordered like the B-tree keys (lexicographic offsets)
  };
  struct Group { std::string name; std::vector<int> members; };
  struct Link { std::string name; uint64_t ohdr; uint32_t cache; uint64_t btree, heap; };

  Options opt_;
  FILE* f_;
  uint64_t pos_;
  std::vector<Group> groups_;
  std::vector<Dset> dsets_;
  TextAttrs root_attrs_;

  static uint64_t elem_size(Type t) { return t == F64 ? 8 : (t == I32 ? 4 : 1); }
  static void fill_buffer(const Dset& d, char* p, uint64_t n) {
    if (d.type == F64) { const double v = d.has_fill ? d.fill : 0.0; for (uint64_t i = 0; i < n; i++) std::memcpy(p + 8 * i, &v, 8); }
    else std::memset(p, 0, (size_t)(n * elem_size(d.type)));
  }
  static void u8(std::string& s, unsigned v) { s.push_back((char)(v & 255)); }
  static void u16(std::string& s, unsigned v) { u8(s, v); u8(s, v >> 8); }
  static void u32(std::string& s, uint32_t v) { for (int i = 0; i < 4; i++) u8(s, v >> (8 * i)); }
  static void u64(std::string& s, uint64_t v) { for (int i = 0; i < 8; i++) u8(s, (unsigned)(v >> (8 * i))); }
  static void pad8(std::string& s) { while (s.size() % 8) s.push_back('\0'); }
  void put_raw(const void* p, size_t n) {
    if (n && std::fwrite(p, 1, n, f_) != n) throw std::runtime_error("write failed");
    pos_ += n;
  }
  void put(const std::string& s) { put_raw(s.data(), s.size()); }
  void align8() { if (pos_ % 8) put(std::string(8 - pos_ % 8, '\0')); }

  // ---- header messages (bodies)
  static std::string dataspace(const std::vector<uint64_t>& dims) {
    std::string s;
    u8(s, 1); u8(s, (unsigned)dims.size()); u8(s, 0); u8(s, 0); u32(s, 0);   // version 1, rank, flags, reserved
    for (uint64_t d : dims) u64(s, d);
    return s;
  }
  static std::string datatype(Type t, uint32_t strlen_ = 1) {
    std::string s;
    if (t == F64) {   // IEEE little-endian double
      u8(s, 0x11); u8(s, 0x20); u8(s, 0x3f); u8(s, 0x00); u32(s, 8);
      u16(s, 0); u16(s, 64); u8(s, 52); u8(s, 11); u8(s, 0); u8(s, 52); u32(s, 1023);
    } else if (t == I32) {   // little-endian two's complement
      u8(s, 0x10); u8(s, 0x08); u8(s, 0); u8(s, 0); u32(s, 4);
      u16(s, 0); u16(s, 32);
    } else {          // fixed-length ASCII string, space padded (a Fortran character)
      u8(s, 0x13); u8(s, 0x02); u8(s, 0); u8(s, 0); u32(s, strlen_);
    }
    return s;
  }
  static std::string message(unsigned type, unsigned flags, std::string body) {
    pad8(body);
    std::string s;
    u16(s, type); u16(s, (unsigned)body.size()); u8(s, flags); u8(s, 0); u8(s, 0); u8(s, 0);
    return s + body;
  }
  static std::string attribute(const std::string& name, const std::string& text) {
    const std::string value = text.empty() ? std::string(" ") : text;
    std::string dt = datatype(CHAR1, (uint32_t)value.size()), ds = dataspace({1});
    std::string s;
    u8(s, 1); u8(s, 0); u16(s, (unsigned)name.size() + 1); u16(s, (unsigned)dt.size()); u16(s, (unsigned)ds.size());
    std::string nm = name; nm.push_back('\0'); pad8(nm);
    pad8(dt); pad8(ds);
    s += nm + dt + ds + value;
    if (s.size() + 8 > 65000) throw std::runtime_error("attribute too long: " + name);
    return message(0x000C, 0, s);
  }
  static std::string object_header(const std::vector<std::string>& msgs) {
    std::string blob;
    for (const auto& m : msgs) blob += m;
    std::string s;
    u8(s, 1); u8(s, 0); u16(s, (unsigned)msgs.size()); u32(s, 1); u32(s, (uint32_t)blob.size()); u32(s, 0);   // 16-byte prefix
    return s + blob;
  }
  std::string dataset_header(const Dset& d, uint64_t btree) const {
    std::vector<std::string> m;
    m.push_back(message(0x0001, 0, dataspace(d.dims)));
    m.push_back(message(0x0003, 1, datatype(d.type)));
    {  // fill value, version 2: allocation time (3 incremental for chunks, 2 late), write time 2 (if set), defined?
      std::string fv;
      u8(fv, 2); u8(fv, opt_.chunking ? 3 : 2); u8(fv, 2); u8(fv, d.has_fill ? 1 : 0);
      if (d.has_fill) { u32(fv, 8); uint64_t bits; std::memcpy(&bits, &d.fill, 8); u64(fv, bits); }
      m.push_back(message(0x0005, 0, fv));
    }
    if (opt_.chunking) {
      if (opt_.compression) {   // filter pipeline, version 1: deflate (id 1), optional, one client value = level
        std::string fp;
        u8(fp, 1); u8(fp, 1); u16(fp, 0); u32(fp, 0);
        u16(fp, 1); u16(fp, 8); u16(fp, 1); u16(fp, 1);
        fp.append("deflate\0", 8);
        u32(fp, (uint32_t)opt_.level); u32(fp, 0);   // client data + padding to an even count
        m.push_back(message(0x000B, 0, fp));
      }
      std::string l;   // layout version 3, chunked: dimensionality rank+1, B-tree address, chunk dims + element size
      u8(l, 3); u8(l, 2); u8(l, (unsigned)d.dims.size() + 1); u64(l, d.chunks.empty() ? kUndef : btree);
      for (uint64_t c : d.cdims) u32(l, (uint32_t)c);
      u32(l, (uint32_t)elem_size(d.type));
      m.push_back(message(0x0008, 0, l));
    } else {
      const uint64_t nbytes = d.dims[0] * d.row * elem_size(d.type);
      std::string l; u8(l, 3); u8(l, 1); u64(l, nbytes ? d.addr : kUndef); u64(l, nbytes);
      m.push_back(message(0x0008, 0, l));
    }
    for (const auto& a : d.attrs) m.push_back(attribute(a.first, a.second));
    return object_header(m);
  }

  // ---- version-1 B-tree of raw data chunks (node type 1), built bottom-up; returns the root's address
  uint64_t write_chunk_btree(const Dset& d) {
    if (d.chunks.empty()) return kUndef;
    const size_t rank = d.dims.size();
    const uint64_t keysz = 8 + 8 * (rank + 1);
    const uint64_t nodesz = 24 + 2 * kChunkK * (keysz + 8) + keysz;
    struct Ent { uint32_t size, mask; std::vector<uint64_t> offs; uint64_t child; };
    std::vector<Ent> level;
    for (const auto& kv : d.chunks) level.push_back(Ent{kv.second.size, kv.second.mask, kv.second.offs, kv.second.addr});
    std::vector<uint64_t> beyond = level.back().offs;      // right-most key: one chunk beyond the last in every dimension
    for (size_t k = 0; k < rank; k++) beyond[k] += d.cdims[k];
    auto key = [&](std::string& s, uint32_t size, uint32_t mask, const std::vector<uint64_t>& offs) {
      u32(s, size); u32(s, mask);
      for (uint64_t o : offs) u64(s, o);
      u64(s, 0);                                            // the element-size dimension
    };
    for (unsigned lev = 0;; lev++) {
      const size_t nnodes = (level.size() + 2 * kChunkK - 1) / (2 * kChunkK);
      align8();
      const uint64_t base = pos_;
      std::vector<Ent> up;
      for (size_t n = 0; n < nnodes; n++) {
        const size_t a = n * 2 * kChunkK, b = std::min(level.size(), a + 2 * kChunkK);
        std::string s("TREE", 4);
        u8(s, 1); u8(s, lev); u16(s, (unsigned)(b - a));
        u64(s, n == 0 ? kUndef : base + (n - 1) * nodesz);
        u64(s, n + 1 == nnodes ? kUndef : base + (n + 1) * nodesz);
        for (size_t i = a; i < b; i++) { key(s, level[i].size, level[i].mask, level[i].offs); u64(s, level[i].child); }
        if (b < level.size()) key(s, level[b].size, level[b].mask, level[b].offs);   // = left key of the next node's first child
        else key(s, 0, 0, beyond);
        s.resize((size_t)nodesz, '\0');
        put(s);
        up.push_back(Ent{level[a].size, level[a].mask, level[a].offs, base + n * nodesz});
      }
      if (nnodes == 1) return base;
      level.swap(up);
    }
  }

  // writes heap, symbol-table node, B-tree node and the group's object header; returns its address
  uint64_t write_group(std::vector<Link> links, const TextAttrs& attrs, uint64_t* btree_out, uint64_t* heap_out) {
    if (links.size() > 2 * kLeafK) throw std::runtime_error("too many links in one group");
    std::sort(links.begin(), links.end(), [](const Link& a, const Link& b) { return std::strcmp(a.name.c_str(), b.name.c_str()) < 0; });
    // local heap data segment: "" at offset 0, then the names, then one free block
    std::string seg(8, '\0');
    std::vector<uint64_t> name_off;
    for (const auto& l : links) {
      name_off.push_back(seg.size());
      seg += l.name; seg.push_back('\0'); pad8(seg);
    }
    const uint64_t free_off = seg.size();
    u64(seg, 1); u64(seg, 16);   // free block: next = H5HL_FREE_NULL, size 16
    align8();
    const uint64_t heap = pos_;
    std::string hp("HEAP", 4);
    u8(hp, 0); u8(hp, 0); u8(hp, 0); u8(hp, 0);
    u64(hp, seg.size()); u64(hp, free_off); u64(hp, heap + 32);
    put(hp + seg);
    // symbol-table node
    align8();
    const uint64_t snod = pos_;
    std::string sn("SNOD", 4);
    u8(sn, 1); u8(sn, 0); u16(sn, (unsigned)links.size());
    for (size_t i = 0; i < 2 * kLeafK; i++) {
      if (i < links.size()) {
        const Link& l = links[i];
        u64(sn, name_off[i]); u64(sn, l.ohdr); u32(sn, l.cache); u32(sn, 0);
        if (l.cache == 1) { u64(sn, l.btree); u64(sn, l.heap); } else { u64(sn, 0); u64(sn, 0); }
      } else {
        sn.append(40, '\0');
      }
    }
    put(sn);
    // B-tree node (type 0 = group, leaf level, one child)
    align8();
    const uint64_t bt = pos_;
    std::string tr("TREE", 4);
    u8(tr, 0); u8(tr, 0); u16(tr, links.empty() ? 0 : 1); u64(tr, kUndef); u64(tr, kUndef);
    u64(tr, 0);                                   // key 0: the empty string
    u64(tr, snod);                                // child 0
    u64(tr, links.empty() ? 0 : name_off.back()); // key 1: the largest name of child 0
    tr.resize(24 + (2 * kInternalK + 1) * 8 + 2 * kInternalK * 8, '\0');
    put(tr);
    // object header
    std::vector<std::string> m;
    { std::string st; u64(st, bt); u64(st, heap); m.push_back(message(0x0011, 0, st)); }
    for (const auto& a : attrs) m.push_back(attribute(a.first, a.second));
    align8();
    const uint64_t oh = pos_;
    put(object_header(m));
    *btree_out = bt; *heap_out = heap;
    return oh;
  }
};

}  // namespace maghost
