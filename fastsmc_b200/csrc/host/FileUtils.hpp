// fastsmc_b200 host layer — gzip-transparent line reader / writer over zlib.
// Replaces the reference's boost::iostreams wrappers (ref: ASMC_SRC/SRC/FileUtils.hpp, FileUtils.cpp).
#pragma once

#include <condition_variable>
#include <cstring>
#include <deque>
#include <fstream>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

namespace FileUtils
{

inline bool fileExists(const std::string& path)
{
  std::ifstream f(path);
  return f.good();
}

// Reads text lines from a plain or gzip file (zlib's gz* API passes plain files through).  Decompression runs on
// its own thread, a few blocks ahead of the parser: inflating a large .hap.gz and turning its text into packed bits
// are the two halves of the read time, and they overlap.
class LineReader
{
  static constexpr size_t kBlock = size_t{4} << 20;
  static constexpr size_t kDepth = 4;
  gzFile mFile = nullptr;
  std::thread mThread;
  std::mutex mMutex;
  std::condition_variable mCv;
  std::deque<std::vector<char>> mFull;
  bool mEof = false, mStop = false;
  std::vector<char> mCur;
  size_t mPos = 0;

  void produce()
  {
    for (;;) {
      std::vector<char> block(kBlock);
      const int n = gzread(mFile, block.data(), static_cast<unsigned>(kBlock));
      std::unique_lock<std::mutex> lock(mMutex);
      if (n <= 0) {
        mEof = true;
        mCv.notify_all();
        return;
      }
      block.resize(static_cast<size_t>(n));
      mCv.wait(lock, [this] { return mFull.size() < kDepth || mStop; });
      if (mStop) {
        return;
      }
      mFull.push_back(std::move(block));
      mCv.notify_all();
    }
  }
  bool refill()
  {
    std::unique_lock<std::mutex> lock(mMutex);
    mCv.wait(lock, [this] { return !mFull.empty() || mEof; });
    if (mFull.empty()) {
      return false;
    }
    mCur = std::move(mFull.front());
    mFull.pop_front();
    mPos = 0;
    mCv.notify_all();
    return true;
  }

public:
  explicit LineReader(const std::string& path)
  {
    mFile = gzopen(path.c_str(), "rb");
    if (!mFile) {
      throw std::runtime_error("ERROR: could not open " + path);
    }
    gzbuffer(mFile, 1u << 20);
    mThread = std::thread([this] { produce(); });
  }
  LineReader(const LineReader&) = delete;
  LineReader& operator=(const LineReader&) = delete;
  ~LineReader()
  {
    {
      std::lock_guard<std::mutex> lock(mMutex);
      mStop = true;
    }
    mCv.notify_all();
    if (mThread.joinable()) {
      mThread.join();
    }
    if (mFile) {
      gzclose(mFile);
    }
  }
  // false at end of file; strips the trailing newline
  bool next(std::string& line)
  {
    line.clear();
    for (;;) {
      if (mPos >= mCur.size() && !refill()) {
        return !line.empty();
      }
      const char* b = mCur.data() + mPos;
      const size_t left = mCur.size() - mPos;
      const char* e = static_cast<const char*>(std::memchr(b, '\n', left));
      if (!e) {
        line.append(b, left);
        mPos = mCur.size();
        continue;
      }
      line.append(b, static_cast<size_t>(e - b));
      mPos += static_cast<size_t>(e - b) + 1;
      if (!line.empty() && line.back() == '\r') {
        line.pop_back();
      }
      return true;
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
