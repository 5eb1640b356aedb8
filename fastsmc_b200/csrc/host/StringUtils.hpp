// fastsmc_b200 host layer — string helpers (ref: ASMC_SRC/SRC/StringUtils.hpp, StringUtils.cpp:36-44).
#pragma once

#include <algorithm>
#include <sstream>
#include <string>
#include <vector>

namespace StringUtils
{

inline std::string toLower(std::string s)
{
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
  return s;
}

// parse through long double and narrow, as the reference does
inline float stof(const std::string& s)
{
  return static_cast<float>(std::stold(s));
}
inline double stod(const std::string& s)
{
  return static_cast<double>(std::stold(s));
}

inline std::vector<std::string> tokenizeMultipleDelimiters(const std::string& s)
{
  std::vector<std::string> out;
  std::istringstream ss(s);
  std::string t;
  while (ss >> t) {
    out.push_back(t);
  }
  return out;
}

}  // namespace StringUtils
