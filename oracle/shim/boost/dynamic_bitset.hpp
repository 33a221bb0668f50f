// Shim: the subset of boost::dynamic_bitset<> used by engine/db/index/nsg/nsg.cpp
// (ctor (n, 0), operator[] lvalue, reset()).
#pragma once
#include <cstddef>
#include <algorithm>
#include <vector>
namespace boost {
template <typename B = unsigned long>
class dynamic_bitset {
  std::vector<unsigned char> v_;
 public:
  dynamic_bitset() {}
  dynamic_bitset(size_t n, unsigned long) : v_(n, 0) {}
  unsigned char& operator[](size_t i) { return v_[i]; }
  bool operator[](size_t i) const { return v_[i] != 0; }
  void reset() { std::fill(v_.begin(), v_.end(), 0); }
  size_t size() const { return v_.size(); }
};
}  // namespace boost
