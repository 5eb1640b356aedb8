// fastsmc_b200 host layer — small numeric and id helpers with the reference's names and semantics
// (ref: ASMC_SRC/SRC/HmmUtils.hpp, HmmUtils.cpp:65-94,153-217; ASMC_SRC/SRC/HASHING/Utils.cpp:22-34).
// Known-answer tests: tests/test_host_layer.py (the reference's ASMC_SRC/TESTS/test_hmm_utils.cpp cases).
#pragma once

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace asmc
{

// Distance key of the transition tables: two significant digits above the 1e-10 grid, all in float.
inline float roundMorgans(const float value, const int precision, const float min)
{
  if (value <= min) {
    return min;
  }
  const float correction = 10.f - static_cast<float>(precision);
  const float l10 = std::max<float>(0.f, floorf(log10f(value)) + correction);
  const float factor = powf(10.f, 10.f - l10);
  return roundf(value * factor) / factor;
}

inline int roundPhysical(const int value, const int precision)
{
  if (value <= 1) {
    return 1;
  }
  const int l10 = std::max<int>(0, static_cast<int>(floor(log10(value))) - precision);
  const int factor = static_cast<int>(pow(10, l10));
  return static_cast<int>(round(value / static_cast<double>(factor))) * factor;
}

// First site of the decode window: walk left from `from` until cmDist centimorgans have been covered.
inline unsigned getFromPosition(const std::vector<float>& geneticPositions, unsigned from, const float cmDist = 0.5f)
{
  float covered = 0.f;
  while (covered < cmDist && from > 0u) {
    --from;
    covered += (geneticPositions[from + 1u] - geneticPositions[from]) * 100.f;
  }
  return from;
}

// One past the last site of the decode window.
inline unsigned getToPosition(const std::vector<float>& geneticPositions, unsigned to, const float cmDist = 0.5f)
{
  float covered = 0.f;
  while (covered < cmDist && to + 1u < geneticPositions.size()) {
    ++to;
    covered += (geneticPositions[to] - geneticPositions[to - 1u]) * 100.f;
  }
  return std::min<unsigned>(to + 1u, static_cast<unsigned>(geneticPositions.size()));
}

// Genetic length in cM of the hashing words [w1, w2] (inclusive), clamped to the last site.
inline double cmBetween(const int w1, const int w2, const std::vector<float>& geneticPositions, const int wordSize)
{
  const size_t start = static_cast<size_t>(wordSize) * w1;
  const size_t end = std::min<size_t>(static_cast<size_t>(wordSize) * w2 + wordSize - 1, geneticPositions.size() - 1);
  return 100.0 * (geneticPositions[end] - geneticPositions[start]);
}

inline std::pair<unsigned long, unsigned long> hapToDipId(const unsigned long hapId)
{
  return {hapId / 2ul, 1ul + (hapId % 2ul)};
}

inline unsigned long dipToHapId(const unsigned long ind, const unsigned long hap)
{
  return 2ul * ind + hap - 1ul;
}

inline std::string indPlusHapToCombinedId(const std::string& indId, const unsigned long hap)
{
  if (indId.empty() || !(hap == 1ul || hap == 2ul)) {
    throw std::runtime_error("Expected an individual ID and either 1 or 2, but got " + indId + " and " +
                             std::to_string(hap) + "\n");
  }
  return indId + "#" + std::to_string(hap);
}

inline std::pair<std::string, unsigned long> combinedIdToIndPlusHap(const std::string& combinedId)
{
  const size_t n = combinedId.length();
  if (n < 3 || !(combinedId.compare(n - 2, 2, "#1") == 0 || combinedId.compare(n - 2, 2, "#2") == 0)) {
    throw std::runtime_error("Expected combined ID in form <id>#1 OR <id>#2, but got " + combinedId + "\n");
  }
  return {combinedId.substr(0, n - 2), combinedId.back() == '1' ? 1ul : 2ul};
}

inline unsigned long getIndIdxFromIdString(const std::vector<std::string>& idStrings, const std::string& idString)
{
  const auto it = std::find(idStrings.begin(), idStrings.end(), idString);
  if (it == idStrings.end()) {
    throw std::runtime_error("The ID string " + idString + " is not in the list of IDs\n");
  }
  return static_cast<unsigned long>(std::distance(idStrings.begin(), it));
}

}  // namespace asmc
