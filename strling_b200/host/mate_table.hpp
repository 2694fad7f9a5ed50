// The mate table of `strling extract` (Cache.tbl, extract.nim:89-91) and the qname hash that picks a record's replay shard and its
// slot in the table.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "tread.hpp"

namespace strling {

inline uint64_t hash_name(const char *s, size_t n) {  // 8 bytes at a time, multiply-xorshift mixing
  uint64_t h = 0x9e3779b97f4a7c15ull ^ (n * 0xff51afd7ed558ccdull);
  while (n >= 8) {
    uint64_t w;
    std::memcpy(&w, s, 8);
    h = (h ^ w) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    s += 8;
    n -= 8;
  }
  uint64_t w = 0;
  std::memcpy(&w, s, n);
  h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 29;
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 32;
  return h;
}

// Cache.tbl (extract.nim:89-91): qname -> the first-seen mate.  Open addressing over indices into an entry pool
// (linear probing, backward-shift deletion); names of up to 54 bytes live inside the entry, so the common insert / take
// pair allocates nothing.
class MateTable {
 public:
  struct Entry {
    uint64_t hash;
    TreadCore t;
    uint32_t name_len;
    char name[54];
    std::string long_name;
    const char *name_ptr() const { return name_len <= sizeof(name) ? name : long_name.data(); }
  };
  MateTable() { slots_.assign(1024, 0); }
  size_t size() const { return live_; }
  // slot of the key or SIZE_MAX
  size_t find(const char *name, uint32_t len, uint64_t h) const {
    const size_t mask = slots_.size() - 1;
    for (size_t s = (size_t)h & mask;; s = (s + 1) & mask) {
      const uint32_t e = slots_[s];
      if (!e) return SIZE_MAX;
      const Entry &en = pool_[e - 1];
      if (en.hash == h && en.name_len == len && std::memcmp(en.name_ptr(), name, len) == 0) return s;
    }
  }
  Entry &at(size_t slot) { return pool_[slots_[slot] - 1]; }
  void insert(const char *name, uint32_t len, uint64_t h, const TreadCore &t) {  // the key must be absent
    if ((live_ + 1) * 2 > slots_.size()) grow();
    uint32_t idx;
    if (!free_.empty()) { idx = free_.back(); free_.pop_back(); }
    else { pool_.emplace_back(); idx = (uint32_t)pool_.size() - 1; }
    Entry &en = pool_[idx];
    en.hash = h;
    en.t = t;
    en.name_len = len;
    if (len <= sizeof(en.name)) std::memcpy(en.name, name, len);
    else en.long_name.assign(name, len);
    place(idx);
    live_++;
  }
  void erase(size_t slot) {
    const size_t mask = slots_.size() - 1;
    const uint32_t idx = slots_[slot] - 1;
    pool_[idx].long_name.clear();
    free_.push_back(idx);
    live_--;
    // backward shift: pull later members of the probe run into the hole while that shortens their probe distance
    size_t hole = slot;
    for (size_t s = (slot + 1) & mask;; s = (s + 1) & mask) {
      const uint32_t e = slots_[s];
      if (!e) break;
      const size_t home = (size_t)pool_[e - 1].hash & mask;
      if (((s - home) & mask) >= ((s - hole) & mask)) { slots_[hole] = e; hole = s; }
    }
    slots_[hole] = 0;
  }

 private:
  void place(uint32_t idx) {
    const size_t mask = slots_.size() - 1;
    size_t s = (size_t)pool_[idx].hash & mask;
    while (slots_[s]) s = (s + 1) & mask;
    slots_[s] = idx + 1;
  }
  void grow() {
    std::vector<uint32_t> old;
    old.swap(slots_);
    slots_.assign(old.size() * 2, 0);
    for (uint32_t e : old)
      if (e) place(e - 1);
  }
  std::vector<uint32_t> slots_;
  std::vector<Entry> pool_;
  std::vector<uint32_t> free_;
  size_t live_ = 0;
};

}  // namespace strling
