// fastsmc_b200 host layer — gzip-transparent line reader / writer over zlib.
// Replaces the reference's boost::iostreams wrappers (ref: ASMC_SRC/SRC/FileUtils.hpp, FileUtils.cpp).
#pragma once

#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <zlib.h>

namespace FileUtils
{

inline bool fileExists(const std::string& path)
{
  std::ifstream f(path);
  return f.good();
}

// Reads text lines from a plain or gzip file (zlib's gz* API passes plain files through).
class LineReader
{
  gzFile mFile = nullptr;
  std::vector<char> mChunk;

public:
  explicit LineReader(const std::string& path) : mChunk(1 << 16)
  {
    mFile = gzopen(path.c_str(), "rb");
    if (!mFile) {
      throw std::runtime_error("ERROR: could not open " + path);
    }
    gzbuffer(mFile, 1u << 20);
  }
  LineReader(const LineReader&) = delete;
  LineReader& operator=(const LineReader&) = delete;
  ~LineReader()
  {
    if (mFile) {
      gzclose(mFile);
    }
  }
  // false at end of file; strips the trailing newline
  bool next(std::string& line)
  {
    line.clear();
    for (;;) {
      if (!gzgets(mFile, mChunk.data(), static_cast<int>(mChunk.size()))) {
        return !line.empty();
      }
      const size_t n = std::strlen(mChunk.data());
      line.append(mChunk.data(), n);
      if (n && line.back() == '\n') {
        line.pop_back();
        if (!line.empty() && line.back() == '\r') {
          line.pop_back();
        }
        return true;
      }
    }
  }
};

// First of root+ext that exists, or "".
inline std::string firstExisting(const std::string& root, std::initializer_list<const char*> exts)
{
  for (const char* e : exts) {
    if (fileExists(root + e)) {
      return root + e;
    }
  }
  return "";
}

}  // namespace FileUtils
