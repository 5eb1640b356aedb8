// Test scaffolding (see oracle/shim/Eigen/Core): the part of boost::dynamic_bitset<> the reference's Individuals class
// uses (ref: ASMC_SRC/SRC/HASHING/Individuals.hpp:32-62) — words of at most 64 bits.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
namespace boost
{
template <class Block = unsigned long> class dynamic_bitset
{
  unsigned long bits = 0ul;
  std::size_t n = 0;

public:
  dynamic_bitset() = default;
  explicit dynamic_bitset(std::size_t size, unsigned long value = 0ul) : bits(value), n(size)
  {
    if (size > 64) {
      throw std::length_error("dynamic_bitset shim: at most 64 bits");
    }
  }
  dynamic_bitset& reset()
  {
    bits = 0ul;
    return *this;
  }
  dynamic_bitset& set(std::size_t i)
  {
    bits |= 1ul << i;
    return *this;
  }
  bool test(std::size_t i) const { return (bits >> i) & 1ul; }
  unsigned long to_ulong() const { return bits; }
  std::size_t size() const { return n; }
};
template <class B> void to_string(const dynamic_bitset<B>& b, std::string& out)
{
  out.assign(b.size(), '0');
  for (std::size_t i = 0; i < b.size(); ++i) {
    if (b.test(i)) {
      out[b.size() - 1 - i] = '1';
    }
  }
}
}  // namespace boost
