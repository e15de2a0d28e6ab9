// h5min.hpp — self-contained HDF5 writer for the momentum-space loop file of Loop_Mugiq::writeLoopsHDF5_Mom
// (/root/reference/lib/loop_mugiq.cpp:530-656: /mom_%+d_%+d_%+d/<disp tag>/<GammaName>/loop, dataset [T][2] native
// double or float).  No HDF5 library exists in this image; the file is produced directly from the HDF5 File Format
// Specification in the dialect H5Fcreate writes by default: superblock version 0, symbol-table groups (v1 B-tree
// "TREE" -> "SNOD" + local heap "HEAP"), version-1 object headers, contiguous little-endian IEEE datasets.
// Same algorithm and byte-for-byte the same output as mugiq_b200/h5min.py (tests/test_host_mirror.py compares them);
// that module's header lists what is and is not verified.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace h5min {

constexpr uint64_t kUndef = 0xFFFFFFFFFFFFFFFFull;
constexpr int kInternalK = 16;  // library default

struct Dataset {
  std::vector<uint64_t> dims;
  int itemsize = 8;  // 8 = double, 4 = float
  std::vector<char> raw;
};

struct Node {
  std::map<std::string, std::unique_ptr<Node>> groups;  // std::map orders by bytes, as the B-tree requires
  std::map<std::string, Dataset> datasets;
  size_t entries() const { return groups.size() + datasets.size(); }
};

class File {
  Node root_;
  std::vector<char> buf_;
  int leafK_ = 4;

  template <typename T> static void put(std::vector<char> &b, T v) {
    const char *p = reinterpret_cast<const char *>(&v);
    b.insert(b.end(), p, p + sizeof(T));
  }
  static void pad8(std::vector<char> &b) { b.resize((b.size() + 7) & ~size_t(7), 0); }
  uint64_t alloc(const std::vector<char> &data) {
    const uint64_t off = buf_.size();
    buf_.insert(buf_.end(), data.begin(), data.end());
    pad8(buf_);
    return off;
  }
  static std::vector<char> message(uint16_t type, std::vector<char> data, uint8_t flags) {
    pad8(data);
    std::vector<char> m;
    put<uint16_t>(m, type);
    put<uint16_t>(m, (uint16_t)data.size());
    put<uint8_t>(m, flags);
    m.resize(m.size() + 3, 0);
    m.insert(m.end(), data.begin(), data.end());
    return m;
  }
  static std::vector<char> objectHeader(const std::vector<std::vector<char>> &msgs) {
    std::vector<char> body;
    for (const auto &m : msgs) body.insert(body.end(), m.begin(), m.end());
    std::vector<char> h;
    put<uint8_t>(h, 1);
    put<uint8_t>(h, 0);
    put<uint16_t>(h, (uint16_t)msgs.size());
    put<uint32_t>(h, 1);
    put<uint32_t>(h, (uint32_t)body.size());
    h.resize(h.size() + 4, 0);
    h.insert(h.end(), body.begin(), body.end());
    return h;
  }
  uint64_t writeDataset(const Dataset &d) {
    const uint64_t raw = d.raw.empty() ? kUndef : alloc(d.raw);
    std::vector<char> space;
    put<uint8_t>(space, 1);
    put<uint8_t>(space, (uint8_t)d.dims.size());
    put<uint8_t>(space, 0);
    space.resize(8, 0);
    for (uint64_t x : d.dims) put<uint64_t>(space, x);
    std::vector<char> type;
    const bool dbl = d.itemsize == 8;
    put<uint8_t>(type, 0x11);
    put<uint8_t>(type, 0x20);
    put<uint8_t>(type, dbl ? 63 : 31);
    put<uint8_t>(type, 0);
    put<uint32_t>(type, (uint32_t)d.itemsize);
    put<uint16_t>(type, 0);
    put<uint16_t>(type, dbl ? 64 : 32);
    put<uint8_t>(type, dbl ? 52 : 23);
    put<uint8_t>(type, dbl ? 11 : 8);
    put<uint8_t>(type, 0);
    put<uint8_t>(type, dbl ? 52 : 23);
    put<uint32_t>(type, dbl ? 1023 : 127);
    std::vector<char> fill{2, 2, 2, 0};
    std::vector<char> layout;
    put<uint8_t>(layout, 3);
    put<uint8_t>(layout, 1);
    put<uint64_t>(layout, raw);
    put<uint64_t>(layout, (uint64_t)d.raw.size());
    return alloc(objectHeader({message(0x0001, space, 0), message(0x0003, type, 1), message(0x0005, fill, 1), message(0x0008, layout, 1)}));
  }
  struct GroupAddr {
    uint64_t oh, tree, heap;
  };
  GroupAddr writeGroup(const Node &n) {
    struct Entry {
      std::string name;
      uint64_t oh;
      uint32_t cache;
      uint64_t s0, s1;
    };
    // merge the two sorted maps into one list sorted by name
    std::vector<Entry> entries;
    auto g = n.groups.begin();
    auto d = n.datasets.begin();
    while (g != n.groups.end() || d != n.datasets.end()) {
      const bool takeGroup = d == n.datasets.end() || (g != n.groups.end() && g->first < d->first);
      if (takeGroup) {
        const GroupAddr a = writeGroup(*g->second);
        entries.push_back({g->first, a.oh, 1, a.tree, a.heap});
        ++g;
      } else {
        entries.push_back({d->first, writeDataset(d->second), 0, 0, 0});
        ++d;
      }
    }
    std::vector<char> heap(8, 0);
    std::vector<uint64_t> offs;
    for (const Entry &e : entries) {
      offs.push_back(heap.size());
      heap.insert(heap.end(), e.name.begin(), e.name.end());
      heap.push_back(0);
      pad8(heap);
    }
    const uint64_t freeOff = heap.size();
    put<uint64_t>(heap, 1);   // next free block: none (H5HL_FREE_NULL)
    put<uint64_t>(heap, 16);  // size of this free block
    const uint64_t heapData = alloc(heap);
    std::vector<char> hh{'H', 'E', 'A', 'P', 0, 0, 0, 0};
    put<uint64_t>(hh, (uint64_t)heap.size());
    put<uint64_t>(hh, freeOff);
    put<uint64_t>(hh, heapData);
    const uint64_t heapAddr = alloc(hh);
    std::vector<char> snod{'S', 'N', 'O', 'D', 1, 0};
    put<uint16_t>(snod, (uint16_t)entries.size());
    for (size_t i = 0; i < entries.size(); i++) {
      put<uint64_t>(snod, offs[i]);
      put<uint64_t>(snod, entries[i].oh);
      put<uint32_t>(snod, entries[i].cache);
      put<uint32_t>(snod, 0);
      put<uint64_t>(snod, entries[i].s0);
      put<uint64_t>(snod, entries[i].s1);
    }
    snod.resize(8 + (size_t)40 * 2 * leafK_, 0);
    const uint64_t snodAddr = alloc(snod);
    std::vector<char> tree{'T', 'R', 'E', 'E', 0, 0};
    put<uint16_t>(tree, entries.empty() ? 0 : 1);
    put<uint64_t>(tree, kUndef);
    put<uint64_t>(tree, kUndef);
    put<uint64_t>(tree, 0);
    if (!entries.empty()) {
      put<uint64_t>(tree, snodAddr);
      put<uint64_t>(tree, offs.back());
    }
    tree.resize(24 + (size_t)(2 * kInternalK + 1) * 8 + (size_t)2 * kInternalK * 8, 0);
    const uint64_t treeAddr = alloc(tree);
    std::vector<char> st;
    put<uint64_t>(st, treeAddr);
    put<uint64_t>(st, heapAddr);
    return {alloc(objectHeader({message(0x0011, st, 1)})), treeAddr, heapAddr};
  }
  static size_t maxEntries(const Node &n) {
    size_t m = n.entries();
    for (const auto &g : n.groups) m = std::max(m, maxEntries(*g.second));
    return m;
  }

 public:
  // path = "/a/b/name"; data = T x 2 values of `itemsize` bytes each
  void addDataset(const std::string &path, const std::vector<uint64_t> &dims, int itemsize, const void *data, size_t nbytes) {
    Node *node = &root_;
    size_t pos = 0;
    std::string last;
    while (pos < path.size()) {
      while (pos < path.size() && path[pos] == '/') pos++;
      size_t end = path.find('/', pos);
      if (end == std::string::npos) end = path.size();
      if (end == pos) break;
      const std::string part = path.substr(pos, end - pos);
      pos = end;
      size_t rest = pos;
      while (rest < path.size() && path[rest] == '/') rest++;
      if (rest >= path.size()) {
        last = part;
        break;
      }
      if (node->datasets.count(part)) throw std::runtime_error(path + ": " + part + " is a dataset");
      auto &child = node->groups[part];
      if (!child) child.reset(new Node);
      node = child.get();
    }
    if (last.empty()) throw std::runtime_error("empty dataset path");
    if (node->datasets.count(last) || node->groups.count(last)) throw std::runtime_error(path + ": duplicate");
    Dataset d;
    d.dims = dims;
    d.itemsize = itemsize;
    d.raw.assign(static_cast<const char *>(data), static_cast<const char *>(data) + nbytes);
    node->datasets.emplace(last, std::move(d));
  }
  std::vector<char> dump() {
    leafK_ = (int)std::max<size_t>(4, (maxEntries(root_) + 1) / 2);
    if (leafK_ > 0xFFFF) throw std::runtime_error("too many links in one group");
    buf_.assign(96, 0);
    const GroupAddr r = writeGroup(root_);
    std::vector<char> sb{'\x89', 'H', 'D', 'F', '\r', '\n', '\x1a', '\n', 0, 0, 0, 0, 0, 8, 8, 0};
    put<uint16_t>(sb, (uint16_t)leafK_);
    put<uint16_t>(sb, (uint16_t)kInternalK);
    put<uint32_t>(sb, 0);
    put<uint64_t>(sb, 0);
    put<uint64_t>(sb, kUndef);
    put<uint64_t>(sb, (uint64_t)buf_.size());
    put<uint64_t>(sb, kUndef);
    put<uint64_t>(sb, 0);
    put<uint64_t>(sb, r.oh);
    put<uint32_t>(sb, 1);
    put<uint32_t>(sb, 0);
    put<uint64_t>(sb, r.tree);
    put<uint64_t>(sb, r.heap);
    std::memcpy(buf_.data(), sb.data(), 96);
    return buf_;
  }
  bool write(const std::string &filename) {
    const std::vector<char> b = dump();
    std::ofstream out(filename, std::ios::binary);
    if (!out) return false;
    out.write(b.data(), (std::streamsize)b.size());
    return (bool)out;
  }
};

}  // namespace h5min
