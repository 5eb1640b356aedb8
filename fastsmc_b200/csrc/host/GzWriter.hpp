// fastsmc_b200 host layer — multi-threaded gzip writer for the IBD record stream.
// The reference writes records one gzwrite at a time on its only thread (ref: ASMC_SRC/SRC/HMM.cpp:1110-1177); at
// GPU decode rates that is the bottleneck of a run, so records are buffered and compressed in blocks on all host
// cores.  Each block becomes one gzip member; concatenated members are a valid gzip file and gunzip to the same
// byte stream the reference produces.
#pragma once

#include <algorithm>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

class GzWriter
{
  FILE* mFile = nullptr;
  std::string mBuf;
  size_t mFlushBytes;
  size_t mBlockBytes;
  int mLevel;
  unsigned mThreads;

  static void compressBlock(const char* src, size_t n, int level, std::string& out)
  {
    z_stream zs{};
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
      throw std::runtime_error("deflateInit2 failed");
    }
    out.resize(deflateBound(&zs, n) + 32);
    zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(src));
    zs.avail_in = static_cast<uInt>(n);
    zs.next_out = reinterpret_cast<Bytef*>(&out[0]);
    zs.avail_out = static_cast<uInt>(out.size());
    const int rc = deflate(&zs, Z_FINISH);
    const size_t produced = zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) {
      throw std::runtime_error("deflate failed");
    }
    out.resize(produced);
  }

public:
  explicit GzWriter(const std::string& path, int level = Z_DEFAULT_COMPRESSION, size_t blockBytes = 1u << 20,
                    size_t flushBytes = 32u << 20)
      : mFlushBytes(flushBytes), mBlockBytes(blockBytes), mLevel(level),
        mThreads(std::max(1u, std::thread::hardware_concurrency()))
  {
    mFile = std::fopen(path.c_str(), "wb");
    if (!mFile) {
      throw std::runtime_error("ERROR: could not open " + path + " for writing");
    }
    mBuf.reserve(flushBytes + (1u << 16));
  }
  GzWriter(const GzWriter&) = delete;
  GzWriter& operator=(const GzWriter&) = delete;
  ~GzWriter()
  {
    try {
      close();
    } catch (...) {
    }
  }

  void write(const void* p, size_t n)
  {
    mBuf.append(static_cast<const char*>(p), n);
    if (mBuf.size() >= mFlushBytes) {
      flush();
    }
  }
  void write(const std::string& s) { write(s.data(), s.size()); }

  void flush()
  {
    if (mBuf.empty() || !mFile) {
      return;
    }
    const size_t nBlocks = (mBuf.size() + mBlockBytes - 1) / mBlockBytes;
    std::vector<std::string> out(nBlocks);
    const unsigned nThreads = static_cast<unsigned>(std::min<size_t>(mThreads, nBlocks));
    auto work = [&](const unsigned t) {
      for (size_t b = t; b < nBlocks; b += nThreads) {
        const size_t lo = b * mBlockBytes, hi = std::min(mBuf.size(), lo + mBlockBytes);
        compressBlock(mBuf.data() + lo, hi - lo, mLevel, out[b]);
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nThreads; ++t) {
      pool.emplace_back(work, t);
    }
    work(0);
    for (auto& th : pool) {
      th.join();
    }
    for (const auto& o : out) {
      if (std::fwrite(o.data(), 1, o.size(), mFile) != o.size()) {
        throw std::runtime_error("ERROR: short write to output file");
      }
    }
    mBuf.clear();
  }

  void close()
  {
    if (!mFile) {
      return;
    }
    if (mBuf.empty() && std::ftell(mFile) == 0) {
      // an empty gzip member, so that the file is a valid (empty) gzip stream like the reference's
      std::string o;
      compressBlock("", 0, mLevel, o);
      std::fwrite(o.data(), 1, o.size(), mFile);
    }
    flush();
    std::fclose(mFile);
    mFile = nullptr;
  }
};
