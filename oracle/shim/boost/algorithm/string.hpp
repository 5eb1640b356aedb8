// Test scaffolding (see oracle/shim/Eigen/Core): boost::algorithm::to_lower as used by the reference.
#pragma once
#include <cctype>
#include <string>
namespace boost
{
namespace algorithm
{
inline void to_lower(std::string& s)
{
  for (char& ch : s) {
    ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
  }
}
}  // namespace algorithm
using algorithm::to_lower;
}  // namespace boost
