// mag_hdf5_writer.hpp -- minimal native HDF5 writer for the Magnetizer OUTPUT file (no libhdf5
// in this environment).
//
// Produces the subset of the HDF5 file format the reference's output uses, in its oldest and
// most widely readable encoding (the mirror image of mag_hdf5_reader.hpp):
//   superblock version 0, 8-byte offsets and lengths
//   groups  = version-1 object header + symbol-table message -> B-tree v1 node (type 0) ->
//             one symbol-table node (SNOD) + local heap with the link names
//   datasets = version-1 object header with dataspace (v1), datatype (v1), fill value (v2),
//             CONTIGUOUS layout (v3) and fixed-length string attributes (attribute message v1)
// What the reference writes and this writer does not: chunking + deflate (IO_hdf5.f90:940-981;
// a storage choice that readers do not see) and the -1d50 fill value (every element of every
// dataset is written here, "never computed" entries hold -99999 like tsDataObj.f90:5).
//
// File contract followed (source/IO_hdf5.f90, source/output.f90, consumer
// python/magnetizer/MagnetizerRun.py:117-150, python/magnetizer/parameters.py):
//   /Output/<quantity>  float64 (ngals, nx_ref, nz)  profiles;  (ngals, nz) scalars
//   /Log/status         1-character strings (ngals, nz);  /Log/{runtime,nsteps,completed} float64 (ngals, 1)
//   dataset attributes  'Units', 'Description'  (arrays of one fixed-length string)
//   root attributes     'Model name', '<namelist>_parameters' (the namelist as one line of text),
//                       'File creation'
// Data is streamed to the file as it is added (only metadata is kept in memory), so the
// 1e6-galaxy gather of BASELINE config 4 does not need the whole output in RAM.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace maghost {

class Hdf5Writer {
 public:
  enum Type { F64, I32, CHAR1 };
  typedef std::vector<std::pair<std::string, std::string>> TextAttrs;

  explicit Hdf5Writer(const std::string& path) : f_(std::fopen(path.c_str(), "wb")), pos_(0) {
    if (!f_) throw std::runtime_error("cannot create " + path);
    std::string zeros(kDataStart, '\0');   // superblock (96 bytes) is written by close()
    put(zeros);
  }
  ~Hdf5Writer() { if (f_) std::fclose(f_); }

  void add_group(const std::string& name) { groups_.push_back(Group{name, {}}); }

  // dims in C order (slowest first), data contiguous
  void add_dataset(const std::string& group, const std::string& name, Type t, const std::vector<uint64_t>& dims, const void* data,
                   const TextAttrs& attrs = TextAttrs()) {
    Group* g = nullptr;
    for (auto& x : groups_) if (x.name == group) g = &x;
    if (!g) throw std::runtime_error("unknown group " + group);
    uint64_t n = 1;
    for (uint64_t d : dims) n *= d;
    const uint64_t nbytes = n * elem_size(t);
    align8();
    Dset d{name, t, dims, pos_, nbytes, attrs, 0};
    if (nbytes && std::fwrite(data, 1, nbytes, f_) != nbytes) throw std::runtime_error("write failed");
    pos_ += nbytes;
    g->dsets.push_back(d);
  }

  void add_root_attribute(const std::string& name, const std::string& text) { root_attrs_.push_back({name, text}); }

  // writes all metadata and the superblock
  void close() {
    if (!f_) return;
    std::vector<Link> root_links;
    for (auto& g : groups_) {
      std::vector<Link> links;
      for (auto& d : g.dsets) {
        align8();
        d.ohdr = pos_;
        put(dataset_header(d));
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
    std::fseek(f_, 0, SEEK_SET);
    if (std::fwrite(sb.data(), 1, sb.size(), f_) != sb.size()) throw std::runtime_error("write failed");
    std::fclose(f_);
    f_ = nullptr;
  }

 private:
  static constexpr uint64_t kUndef = 0xFFFFFFFFFFFFFFFFull;
  static constexpr unsigned kLeafK = 32, kInternalK = 16;   // up to 64 links per group in one symbol-table node
  static constexpr size_t kDataStart = 2048;
  struct Dset { std::string name; Type type; std::vector<uint64_t> dims; uint64_t addr, nbytes; TextAttrs attrs; uint64_t ohdr; };
  struct Group { std::string name; std::vector<Dset> dsets; };
  struct Link { std::string name; uint64_t ohdr; uint32_t cache; uint64_t btree, heap; };

  FILE* f_;
  uint64_t pos_;
  std::vector<Group> groups_;
  TextAttrs root_attrs_;

  static uint64_t elem_size(Type t) { return t == F64 ? 8 : (t == I32 ? 4 : 1); }
  static void u8(std::string& s, unsigned v) { s.push_back((char)(v & 255)); }
  static void u16(std::string& s, unsigned v) { u8(s, v); u8(s, v >> 8); }
  static void u32(std::string& s, uint32_t v) { for (int i = 0; i < 4; i++) u8(s, v >> (8 * i)); }
  static void u64(std::string& s, uint64_t v) { for (int i = 0; i < 8; i++) u8(s, (unsigned)(v >> (8 * i))); }
  static void pad8(std::string& s) { while (s.size() % 8) s.push_back('\0'); }
  void put(const std::string& s) {
    if (std::fwrite(s.data(), 1, s.size(), f_) != s.size()) throw std::runtime_error("write failed");
    pos_ += s.size();
  }
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
  std::string dataset_header(const Dset& d) const {
    std::vector<std::string> m;
    m.push_back(message(0x0001, 0, dataspace(d.dims)));
    m.push_back(message(0x0003, 1, datatype(d.type)));
    { std::string fv; u8(fv, 2); u8(fv, 2); u8(fv, 2); u8(fv, 0); m.push_back(message(0x0005, 0, fv)); }   // fill value v2: late alloc, undefined
    { std::string l; u8(l, 3); u8(l, 1); u64(l, d.nbytes ? d.addr : kUndef); u64(l, d.nbytes); m.push_back(message(0x0008, 0, l)); }
    for (const auto& a : d.attrs) m.push_back(attribute(a.first, a.second));
    return object_header(m);
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
