// fastsmc_b200 host layer — ordered, multi-threaded gzip output pipeline for the IBD record stream.
//
// The reference formats and gzwrite()s one record at a time on its only thread (ref: ASMC_SRC/SRC/HMM.cpp:1110-1177).
// At GPU decode rates that is the bottleneck of a whole run, so here the decoder thread only hands over
// *blocks* of segment records; a pool of workers turns each block into text (or packed binary), deflates it into its
// own gzip member, and the members are written to the file in submission order.  Concatenated members are a valid
// gzip file and gunzip to exactly the byte stream the reference produces.
#pragma once

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <exception>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

class OutputPipeline
{
public:
  /// produce(out) appends the uncompressed bytes of one block to `out`; it runs on a worker thread.
  using Producer = std::function<void(std::string&)>;

  explicit OutputPipeline(const std::string& path, int level = 1, unsigned threads = 0)
      : mLevel(level)
  {
    mFile = std::fopen(path.c_str(), "wb");
    if (!mFile) {
      throw std::runtime_error("ERROR: could not open " + path + " for writing");
    }
    if (threads == 0) {
      threads = std::max(1u, std::thread::hardware_concurrency());
    }
    mMaxQueued = 4 * static_cast<size_t>(threads);
    for (unsigned t = 0; t < threads; ++t) {
      mWorkers.emplace_back([this] { workerLoop(); });
    }
  }
  OutputPipeline(const OutputPipeline&) = delete;
  OutputPipeline& operator=(const OutputPipeline&) = delete;
  ~OutputPipeline()
  {
    try {
      close();
    } catch (...) {
    }
  }

  /// Small in-order writes from the submitting thread (file header); flushed as a block before the next submit().
  void write(const void* p, const size_t n)
  {
    mInline.append(static_cast<const char*>(p), n);
    if (mInline.size() >= (size_t{1} << 20)) {
      flushInline();
    }
  }

  /// Queues one block.  Blocks reach the file in the order of the submit() calls.  Applies back-pressure when the
  /// workers fall behind.
  void submit(Producer produce)
  {
    flushInline();
    enqueue(std::move(produce));
  }

  /// Waits for every queued block, writes the trailer and closes the file.  Rethrows a worker's exception.
  void close()
  {
    if (!mFile) {
      return;
    }
    flushInline();
    {
      std::unique_lock<std::mutex> lock(mMutex);
      mDone.wait(lock, [this] { return mNextToWrite == mNextSeq || mError; });
      mStop = true;
    }
    mWork.notify_all();
    for (auto& w : mWorkers) {
      w.join();
    }
    mWorkers.clear();
    if (!mError && mBytesWritten == 0) {
      // an empty gzip member, so that the file is a valid (empty) gzip stream like the reference's
      std::string o;
      deflateBlock("", 0, mLevel, o);
      std::fwrite(o.data(), 1, o.size(), mFile);
    }
    const bool closeFailed = std::fclose(mFile) != 0;  // a late write error (disk full) surfaces here
    mFile = nullptr;
    if (closeFailed && !mError) {
      mError = std::make_exception_ptr(std::runtime_error("OutputPipeline: error closing the output file (write failed)"));
    }
    if (mError) {
      std::rethrow_exception(mError);
    }
  }

  size_t compressedBytes() const { return mBytesWritten; }

private:
  struct Task {
    size_t seq;
    Producer produce;
  };

  FILE* mFile = nullptr;
  int mLevel;
  std::string mInline;
  std::vector<std::thread> mWorkers;
  std::mutex mMutex;
  std::condition_variable mWork, mSpace, mDone;
  std::deque<Task> mQueue;
  std::map<size_t, std::string> mFinished;  // compressed members waiting for their turn
  size_t mNextSeq = 0, mNextToWrite = 0, mMaxQueued = 8, mBytesWritten = 0;
  bool mStop = false;
  std::exception_ptr mError;

  static void deflateBlock(const char* src, const size_t n, const int level, std::string& out)
  {
    z_stream zs{};
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
      throw std::runtime_error("deflateInit2 failed");
    }
    out.resize(deflateBound(&zs, static_cast<uLong>(n)) + 32);
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

  void flushInline()
  {
    if (mInline.empty()) {
      return;
    }
    auto bytes = std::make_shared<std::string>(std::move(mInline));
    mInline.clear();
    enqueue([bytes](std::string& out) { out.append(*bytes); });
  }

  void enqueue(Producer produce)
  {
    std::unique_lock<std::mutex> lock(mMutex);
    mSpace.wait(lock, [this] { return mNextSeq - mNextToWrite < mMaxQueued || mError; });
    if (mError) {
      std::rethrow_exception(mError);
    }
    mQueue.push_back(Task{mNextSeq++, std::move(produce)});
    lock.unlock();
    mWork.notify_one();
  }

  void workerLoop()
  {
    std::string raw, packed;
    for (;;) {
      Task task;
      {
        std::unique_lock<std::mutex> lock(mMutex);
        mWork.wait(lock, [this] { return mStop || !mQueue.empty(); });
        if (mQueue.empty()) {
          return;
        }
        task = std::move(mQueue.front());
        mQueue.pop_front();
      }
      std::exception_ptr err;
      try {
        raw.clear();
        task.produce(raw);
        packed.clear();
        // zlib takes 32-bit lengths: split very large blocks into several members
        constexpr size_t kMaxMember = size_t{1} << 30;
        std::string member;
        for (size_t lo = 0; lo < raw.size(); lo += kMaxMember) {
          deflateBlock(raw.data() + lo, std::min(kMaxMember, raw.size() - lo), mLevel, member);
          packed.append(member);
        }
      } catch (...) {
        err = std::current_exception();
      }
      std::unique_lock<std::mutex> lock(mMutex);
      if (err && !mError) {
        mError = err;
      }
      mFinished.emplace(task.seq, err ? std::string() : packed);
      // whoever completes the next block in line writes it (and any successors that are ready)
      while (!mFinished.empty() && mFinished.begin()->first == mNextToWrite) {
        std::string bytes = std::move(mFinished.begin()->second);
        mFinished.erase(mFinished.begin());
        if (!mError && !bytes.empty()) {
          if (std::fwrite(bytes.data(), 1, bytes.size(), mFile) != bytes.size()) {
            mError = std::make_exception_ptr(std::runtime_error("ERROR: short write to output file"));
          }
          mBytesWritten += bytes.size();
        }
        ++mNextToWrite;
      }
      lock.unlock();
      mSpace.notify_all();
      mDone.notify_all();
    }
  }
};
