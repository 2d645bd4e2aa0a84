// mag_hdf5_reader.hpp -- minimal HDF5 reader for the Magnetizer input file (no libhdf5 in
// the image).  Handles what the reference's input files use (written with h5py defaults by
// python/prepare_input.py:93-150 of the reference, read on the Fortran side by
// IO_read_dataset_scalar, source/IO_hdf5.f90:365-460): superblock v0, v1 object headers,
// symbol-table groups (v1 B-tree + local heap), contiguous or chunked (v1 chunk B-tree)
// datasets of f32/f64/integers, deflate and shuffle filters, fill values.  Values are returned as
// double.  The file is memory-mapped, so that the driver can look into an output file of any size
// when it resumes a run (Log/completed, the chunk lists of the datasets it extends).
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "mag_hdf5_writer.hpp"   // H5ChunkRec

namespace maghost {

// read-only memory map of a whole file
class FileMap {
 public:
  explicit FileMap(const std::string& path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); throw std::runtime_error("cannot stat " + path); }
    size_ = (size_t)st.st_size;
    if (size_ > 0) {
      void* p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p == MAP_FAILED) { ::close(fd); throw std::runtime_error("cannot map " + path); }
      data_ = (const char*)p;
    }
    ::close(fd);
  }
  ~FileMap() { if (data_) munmap((void*)data_, size_); }
  FileMap(const FileMap&) = delete;
  FileMap& operator=(const FileMap&) = delete;
  size_t size() const { return size_; }
  const char* data() const { return data_; }
  const char& operator[](size_t i) const { return data_[i]; }
 private:
  const char* data_ = nullptr;
  size_t size_ = 0;
};

class H5Reader {
 public:
  explicit H5Reader(const std::string& path) : buf_(path) {
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (buf_.size() < 96 || std::memcmp(buf_.data(), sig, 8) != 0) throw std::runtime_error(path + ": not an HDF5 file");
    if (buf_[8] != 0) throw std::runtime_error("HDF5 superblock version not supported");
    if (buf_[13] != 8 || buf_[14] != 8) throw std::runtime_error("only 8-byte offsets/lengths supported");
    root_ohdr_ = u64(56 + 8);  // root symbol-table entry at 56: name offset (8), object header address (8)
  }

  struct Data {
    std::vector<uint64_t> shape;
    std::vector<double> v;
  };

  // dataset by path, e.g. "Input/r_disk"
  Data read(const std::string& path) const { return read_dataset(locate(path), 0, UNDEF); }
  // rows [g0, g0+n) of the first dimension only (the chunks that overlap them are decoded)
  Data read_rows(const std::string& path, uint64_t g0, uint64_t n) const { return read_dataset(locate(path), g0, n); }
  bool exists(const std::string& path) const {
    try { locate(path); return true; } catch (const std::exception&) { return false; }
  }
  std::vector<uint64_t> shape(const std::string& path) const { return locate(path).shape; }
  // file address of a contiguous dataset's storage (UNDEF for chunked ones)
  uint64_t contiguous_address(const std::string& path) const { const Node n = locate(path); return n.layout == 1 ? n.addr : UNDEF; }
  // the stored chunks of a chunked dataset (empty for contiguous ones): what a writer that extends the file adopts
  std::vector<H5ChunkRec> chunks(const std::string& path) const {
    const Node n = locate(path);
    std::vector<H5ChunkRec> out;
    if (n.layout == 2 && n.addr != UNDEF) list_chunks(n, n.addr, &out);
    return out;
  }

 private:
  struct Node;
  Node locate(const std::string& path) const {
    uint64_t oh = root_ohdr_;
    size_t p = 0;
    while (p < path.size()) {
      size_t q = path.find('/', p);
      if (q == std::string::npos) q = path.size();
      if (q > p) {
        const Node n = open(oh);
        if (!n.is_group) throw std::runtime_error(path + ": not a group on the way");
        auto it = n.children.find(path.substr(p, q - p));
        if (it == n.children.end()) throw std::runtime_error("no such HDF5 object: " + path);
        oh = it->second;
      }
      p = q + 1;
    }
    const Node n = open(oh);
    if (n.is_group) throw std::runtime_error(path + " is a group");
    return n;
  }

  FileMap buf_;
  uint64_t root_ohdr_ = 0;
  static constexpr uint64_t UNDEF = 0xFFFFFFFFFFFFFFFFull;

  uint8_t u8(uint64_t o) const { check(o, 1); return (uint8_t)buf_[o]; }
  uint16_t u16(uint64_t o) const { check(o, 2); uint16_t v; std::memcpy(&v, &buf_[o], 2); return v; }
  uint32_t u32(uint64_t o) const { check(o, 4); uint32_t v; std::memcpy(&v, &buf_[o], 4); return v; }
  uint64_t u64(uint64_t o) const { check(o, 8); uint64_t v; std::memcpy(&v, &buf_[o], 8); return v; }
  void check(uint64_t o, uint64_t n) const { if (o + n > buf_.size()) throw std::runtime_error("HDF5: read past end of file"); }

  struct Node {
    bool is_group = false;
    std::map<std::string, uint64_t> children;  // name -> object header address
    std::vector<uint64_t> shape;
    int dt_class = -1, dt_size = 0;
    bool dt_signed = true;
    int layout = 0;  // 1 contiguous, 2 chunked
    uint64_t addr = UNDEF, size = 0;
    std::vector<uint32_t> cdims;
    std::vector<int> filters;
    std::vector<uint32_t> shuffle_elem;
    bool has_fill = false;
    double fill = 0.0;
  };

  Node open(uint64_t ohdr) const {
    if (u8(ohdr) != 1) throw std::runtime_error("HDF5 object header version not supported");
    const int nmsg = u16(ohdr + 2);
    std::vector<std::pair<uint64_t, uint64_t>> blocks = {{ohdr + 16, u32(ohdr + 8)}};
    Node n;
    int seen = 0;
    for (size_t b = 0; b < blocks.size() && seen < nmsg; b++) {
      uint64_t off = blocks[b].first;
      const uint64_t end = off + blocks[b].second;
      while (off + 8 <= end && seen < nmsg) {
        const int type = u16(off), msize = u16(off + 2);
        const uint64_t body = off + 8;
        seen++;
        switch (type) {
          case 0x10: blocks.push_back({u64(body), u64(body + 8)}); break;
          case 0x01: {  // dataspace
            const int ver = u8(body), rank = u8(body + 1);
            const uint64_t p = body + (ver == 1 ? 8 : 4);
            for (int i = 0; i < rank; i++) n.shape.push_back(u64(p + 8 * i));
            break;
          }
          case 0x03:  // datatype
            n.dt_class = u8(body) & 0x0F; n.dt_size = (int)u32(body + 4); n.dt_signed = (u8(body + 1) & 0x08) != 0;
            if (n.dt_class == 1 && (u8(body + 1) & 1)) throw std::runtime_error("big-endian floats not supported");
            break;
          case 0x08: {  // layout v3
            if (u8(body) != 3) throw std::runtime_error("HDF5 layout version not supported");
            n.layout = u8(body + 1);
            if (n.layout == 1) { n.addr = u64(body + 2); n.size = u64(body + 10); }
            else if (n.layout == 2) {
              const int nd1 = u8(body + 2);
              n.addr = u64(body + 3);
              for (int i = 0; i + 1 < nd1; i++) n.cdims.push_back(u32(body + 11 + 4 * i));
            } else throw std::runtime_error("compact datasets not supported");
            break;
          }
          case 0x05: {  // fill value, versions 1-2 (version 3 packs the fields into flag bits)
            const int ver = u8(body);
            if (ver == 1 || ver == 2) {
              const bool defined = ver == 1 ? true : u8(body + 3) != 0;
              if (defined) {
                const uint32_t sz = u32(body + 4);
                if (sz == 8) { n.has_fill = true; std::memcpy(&n.fill, &buf_[body + 8], 8); }
              }
            } else if (ver == 3) {
              const int flags = u8(body + 1);
              if ((flags & 0x20) && u32(body + 2) == 8) { n.has_fill = true; std::memcpy(&n.fill, &buf_[body + 6], 8); }
            }
            break;
          }
          case 0x0B: {  // filter pipeline
            const int ver = u8(body), nf = u8(body + 1);
            uint64_t p = body + (ver == 1 ? 8 : 2);
            for (int i = 0; i < nf; i++) {
              const int fid = u16(p), nlen = u16(p + 2), ncd = u16(p + 6);
              p += 8;
              if (ver == 1) p += (nlen + 7) & ~7; else if (fid >= 256) p += nlen;
              n.filters.push_back(fid);
              n.shuffle_elem.push_back(ncd > 0 ? u32(p) : 0);
              p += 4 * ncd;
              if (ver == 1 && (ncd % 2)) p += 4;
            }
            break;
          }
          case 0x11: {  // symbol table: B-tree address, local heap address
            n.is_group = true;
            const uint64_t heap_data = u64(u64(body + 8) + 24);
            walk_group(u64(body), heap_data, &n.children);
            break;
          }
          default: break;
        }
        off = body + msize;
      }
    }
    return n;
  }

  void walk_group(uint64_t addr, uint64_t heap_data, std::map<std::string, uint64_t>* out) const {
    check(addr, 8);
    if (std::memcmp(&buf_[addr], "TREE", 4) == 0) {
      const int nent = u16(addr + 6);
      uint64_t p = addr + 24 + 8;
      for (int i = 0; i < nent; i++, p += 16) walk_group(u64(p), heap_data, out);
    } else if (std::memcmp(&buf_[addr], "SNOD", 4) == 0) {
      const int nsym = u16(addr + 6);
      for (int i = 0; i < nsym; i++) {
        const uint64_t e = addr + 8 + 40 * (uint64_t)i;
        const uint64_t s = heap_data + u64(e);
        check(s, 1);
        (*out)[std::string(&buf_[s])] = u64(e + 8);
      }
    } else {
      throw std::runtime_error("HDF5: bad group node");
    }
  }

  double elem(const Node& n, const char* p) const {
    if (n.dt_class == 1) {
      if (n.dt_size == 8) { double v; std::memcpy(&v, p, 8); return v; }
      if (n.dt_size == 4) { float v; std::memcpy(&v, p, 4); return (double)v; }
    } else if (n.dt_class == 0) {
      if (n.dt_size == 8) { int64_t v; std::memcpy(&v, p, 8); return n.dt_signed ? (double)v : (double)(uint64_t)v; }
      if (n.dt_size == 4) { int32_t v; std::memcpy(&v, p, 4); return n.dt_signed ? (double)v : (double)(uint32_t)v; }
      if (n.dt_size == 2) { int16_t v; std::memcpy(&v, p, 2); return n.dt_signed ? (double)v : (double)(uint16_t)v; }
      if (n.dt_size == 1) return n.dt_signed ? (double)*(const int8_t*)p : (double)*(const uint8_t*)p;
    } else if (n.dt_class == 3 && n.dt_size == 1) {
      return (double)*(const uint8_t*)p;   // one-character strings (Log/status): the character code
    }
    throw std::runtime_error("HDF5 datatype not supported");
  }

  void list_chunks(const Node& n, uint64_t addr, std::vector<H5ChunkRec>* out) const {
    check(addr, 24);
    if (std::memcmp(&buf_[addr], "TREE", 4) != 0) throw std::runtime_error("HDF5: bad chunk B-tree");
    const int level = u8(addr + 5), nent = u16(addr + 6);
    const size_t nd = n.shape.size();
    const uint64_t keysz = 8 + 8 * (nd + 1);
    uint64_t p = addr + 24;
    for (int i = 0; i < nent; i++, p += keysz + 8) {
      const uint64_t child = u64(p + keysz);
      if (level > 0) { list_chunks(n, child, out); continue; }
      H5ChunkRec r;
      r.size = u32(p); r.mask = u32(p + 4); r.addr = child;
      for (size_t k = 0; k < nd; k++) r.offs.push_back(u64(p + 8 + 8 * k));
      out->push_back(r);
    }
  }

  // decodes the chunks that overlap rows [g0, g0+rows) of the first dimension into d (whose first row is g0)
  void chunk_walk(const Node& n, uint64_t addr, Data* d, uint64_t g0, uint64_t rows) const {
    check(addr, 24);
    if (std::memcmp(&buf_[addr], "TREE", 4) != 0) throw std::runtime_error("HDF5: bad chunk B-tree");
    const int level = u8(addr + 5), nent = u16(addr + 6);
    const size_t nd = n.shape.size();
    const uint64_t keysz = 8 + 8 * (nd + 1);
    uint64_t p = addr + 24;
    for (int i = 0; i < nent; i++, p += keysz + 8) {
      const uint32_t csize = u32(p), mask = u32(p + 4);
      const uint64_t child = u64(p + keysz);
      if (level > 0) {
        // children are ordered by their first key: skip those that end before g0 or start beyond the range
        const uint64_t first = u64(p + 8), next = u64(p + keysz + 8 + 8);
        if (first >= g0 + rows) break;
        if (i + 1 < nent && next + n.cdims[0] <= g0) continue;   // every chunk of this child ends before row g0
        chunk_walk(n, child, d, g0, rows);
        continue;
      }
      std::vector<uint64_t> offs(nd);
      for (size_t k = 0; k < nd; k++) offs[k] = u64(p + 8 + 8 * k);
      if (offs[0] >= g0 + rows || offs[0] + n.cdims[0] <= g0) continue;
      check(child, csize);
      std::vector<char> raw(buf_.data() + child, buf_.data() + child + csize);
      size_t nelem = 1;
      for (auto c : n.cdims) nelem *= c;
      for (int fi = (int)n.filters.size() - 1; fi >= 0; fi--) {
        if (mask & (1u << fi)) continue;
        if (n.filters[fi] == 1) {
          std::vector<char> outb(nelem * n.dt_size);
          uLongf dl = outb.size();
          if (uncompress((Bytef*)outb.data(), &dl, (const Bytef*)raw.data(), raw.size()) != Z_OK) throw std::runtime_error("HDF5: inflate failed");
          outb.resize(dl);
          raw.swap(outb);
        } else if (n.filters[fi] == 2) {
          const size_t es = n.shuffle_elem[fi] ? n.shuffle_elem[fi] : (size_t)n.dt_size, ne = raw.size() / es;
          std::vector<char> outb(raw.size());
          for (size_t b = 0; b < es; b++) for (size_t e = 0; e < ne; e++) outb[e * es + b] = raw[b * ne + e];
          raw.swap(outb);
        } else if (n.filters[fi] == 3) {
          raw.resize(raw.size() - 4);
        } else {
          throw std::runtime_error("HDF5 filter not supported");
        }
      }
      // scatter the chunk (row-major) into the dataset
      std::vector<uint64_t> idx(nd, 0);
      for (size_t e = 0; e < nelem; e++) {
        bool inside = true;
        uint64_t lin = 0;
        for (size_t k = 0; k < nd; k++) {
          uint64_t g = offs[k] + idx[k];
          if (g >= n.shape[k]) { inside = false; break; }
          if (k == 0) {
            if (g < g0 || g >= g0 + rows) { inside = false; break; }
            g -= g0;
            lin = g;
          } else {
            lin = lin * n.shape[k] + g;
          }
        }
        if (inside) d->v[lin] = elem(n, raw.data() + e * n.dt_size);
        for (int k = (int)nd - 1; k >= 0; k--) { if (++idx[k] < n.cdims[k]) break; idx[k] = 0; }
      }
    }
  }

  Data read_dataset(const Node& n, uint64_t g0, uint64_t rows) const {
    Data d;
    d.shape = n.shape;
    if (n.shape.empty()) throw std::runtime_error("HDF5: scalar dataspace not supported");
    if (rows == UNDEF) { g0 = 0; rows = n.shape[0]; }
    if (g0 + rows > n.shape[0]) throw std::runtime_error("HDF5: rows out of range");
    d.shape[0] = rows;
    size_t rowsz = 1;
    for (size_t k = 1; k < n.shape.size(); k++) rowsz *= n.shape[k];
    const size_t total = rows * rowsz;
    d.v.assign(total, n.has_fill ? n.fill : 0.0);   // never-written chunks read as the fill value
    if (n.addr == UNDEF) return d;
    if (n.layout == 1) {
      check(n.addr + g0 * rowsz * n.dt_size, total * n.dt_size);
      for (size_t i = 0; i < total; i++) d.v[i] = elem(n, &buf_[n.addr + (g0 * rowsz + i) * n.dt_size]);
    } else if (n.layout == 2) {
      chunk_walk(n, n.addr, &d, g0, rows);
    } else {
      throw std::runtime_error("HDF5 layout not supported");
    }
    return d;
  }
};

// the 13 per-galaxy time series in the order of load_input_parameters (input_parameters.f90:158-170)
static const char* const kInputColumns[13] = {"r_disk", "v_disk", "r_bulge", "v_bulge", "r_halo", "v_halo", "nfw_cs1",
                                             "Mgas_disk", "Mstars_disk", "SFR", "Mhalo", "Mstars_bulge", "central"};

struct MagnetizerInput {
  int ngal = 0, nz = 0;
  std::vector<double> t_gyr;   // [nz]
  std::vector<double> inputs;  // [13][ngal][nz]
};

inline MagnetizerInput load_magnetizer_input(const std::string& path) {
  H5Reader f(path);
  MagnetizerInput in;
  H5Reader::Data t = f.read("Input/t");
  in.t_gyr = t.v;
  in.nz = (int)t.v.size();
  for (int c = 0; c < 13; c++) {
    H5Reader::Data d = f.read(std::string("Input/") + kInputColumns[c]);
    if (d.shape.size() != 2 || (int)d.shape[1] != in.nz) throw std::runtime_error("unexpected shape of Input/" + std::string(kInputColumns[c]));
    if (c == 0) { in.ngal = (int)d.shape[0]; in.inputs.resize((size_t)13 * in.ngal * in.nz); }
    std::memcpy(&in.inputs[(size_t)c * in.ngal * in.nz], d.v.data(), sizeof(double) * d.v.size());
  }
  return in;
}

}  // namespace maghost
