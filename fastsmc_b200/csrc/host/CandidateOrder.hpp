// fastsmc_b200 host layer — puts the GPU's match intervals into the order in which the reference hands candidates
// to the HMM.
//
// Why this exists: the reference decodes candidates in batches of `batchSize` consecutive decodeFromHashing calls, and
// every pair of a batch is decoded and scanned over the union window of the batch (ref: ASMC_SRC/SRC/HMM.cpp:561-565,
// 1199-1204), so segment boundaries depend on the call order.  That order is the iteration order of two
// boost::unordered_map instances (ref: HASHING/SeedHash.hpp:34,80; HASHING/ExtendHash.hpp:29,88-97,112-115; boost 1.75
// per the reference's vcpkg manifest).  The match intervals themselves are order-free and come from the GPU
// (fsmc_seed); this file only replays the two maps' node order over them:
//   * per word, new pairs enter the extend map in seed-bucket order, then (a, b) ascending within a bucket;
//   * after each word the extend map is scanned in list order and intervals that ended more than `gap` words ago
//     are handed to the HMM (if long enough) and erased; at the end everything left is flushed in list order.
// Node order rules of boost <= 1.79 with an integer key (identity hash, prime bucket counts, max load factor 1):
// see NodeOrderMap below.
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <new>
#include <stdexcept>
#include <thread>
#include <utility>
#include <vector>

#include <sys/mman.h>

#include "../../../include/fastsmc_b200.h"
#include "SeedMapOrder.hpp"

namespace candidate_order
{

// Replays the reference's emission order.
//   intervals : ALL match intervals of the job (FSMC_SEED_ALL_INTERVALS), any order
//   rawWord   : rawWord(h, w) = the 64-SNP word of RAW alleles of local haplotype h (the SeedHash key)
//   longEnough: the Match::print length test for an interval
//   emit      : called with the interval index for every candidate, in reference order
template <class WordFn, class LengthFn, class EmitFn>
void replayReferenceOrder(const std::vector<fsmc_match>& intervals, const uint32_t numHaps, const int numWords,
                          const int gap, WordFn&& rawWord, LengthFn&& longEnough, EmitFn&& emit)
{
  // intervals by start word
  std::vector<std::vector<int64_t>> startingAt(numWords);
  for (int64_t i = 0; i < static_cast<int64_t>(intervals.size()); ++i) {
    startingAt[intervals[i].startWord].push_back(i);
  }
  NodeOrderMap seeds, extend;
  std::vector<int64_t> rankOfNode;
  struct Creation {
    int64_t rank;
    uint32_t a, b;
    int64_t index;
  };
  std::vector<Creation> created;
  for (int w = 0; w < numWords; ++w) {
    // SeedHash::insertIndividuals for every haplotype in index order (ref: FastSMC.cpp:204-206)
    for (uint32_t h = 0; h < numHaps; ++h) {
      bool isNew;
      seeds.insert(rawWord(h, w), 0, isNew);
    }
    if (!startingAt[w].empty()) {
      // rank of each distinct word in the seed map's iteration order
      std::vector<std::pair<uint64_t, int64_t>> rank;
      rank.reserve(seeds.size());
      int64_t r = 0;
      for (int n = seeds.first(); n != NodeOrderMap::kEnd; n = seeds.next(n)) {
        rank.emplace_back(seeds.key(n), r++);
      }
      std::sort(rank.begin(), rank.end());
      created.clear();
      for (const int64_t i : startingAt[w]) {
        const fsmc_match& m = intervals[i];
        const uint64_t k = rawWord(m.hapA, w);
        const auto it = std::lower_bound(rank.begin(), rank.end(), std::make_pair(k, int64_t{0}));
        created.push_back(Creation{it->second, m.hapA, m.hapB, i});
      }
      // within a bucket the reference enumerates i < ii over haplotypes in insertion (= index) order
      std::sort(created.begin(), created.end(), [](const Creation& x, const Creation& y) {
        return x.rank != y.rank ? x.rank < y.rank : (x.a != y.a ? x.a < y.a : x.b < y.b);
      });
      for (const Creation& c : created) {
        bool isNew;
        extend.insert(static_cast<uint64_t>(c.a) * numHaps + c.b, c.index, isNew);
      }
    }
    seeds.clear();
    // ExtendHash::clearPairsPriorTo(w - gap) (ref: HASHING/ExtendHash.hpp:85-98)
    for (int n = extend.first(); n != NodeOrderMap::kEnd;) {
      const int64_t i = extend.payload(n);
      if (intervals[i].endWord < w - gap) {
        if (longEnough(intervals[i])) {
          emit(i);
        }
        n = extend.erase(n);
      } else {
        n = extend.next(n);
      }
    }
  }
  // ExtendHash::clearAllPairs (ref: HASHING/ExtendHash.hpp:110-116)
  for (int n = extend.first(); n != NodeOrderMap::kEnd;) {
    const int64_t i = extend.payload(n);
    if (longEnough(intervals[i])) {
      emit(i);
    }
    n = extend.erase(n);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same order, computed without walking linked lists (replayReferenceOrder above is kept as the executable
// specification; tests compare the two).
//
// At biobank density most intervals are one or two words long and never become candidates, but every one of them
// is a node of the reference's extend map, and the reference (and its literal replay) walks the whole map after
// every word: ~10^9 dependent pointer loads at 10 000 samples x 50 000 SNPs (125 s).  The node order of a
// boost <= 1.79 table is a function of the insertion history alone:
//   * the list is a sequence of bucket groups; a node for an EMPTY bucket starts a new group at the front of the
//     list, a node for a non-empty bucket goes to the front of its group; erasing keeps the order;
//   * a rehash walks the list once: groups re-form in the order their first node is met, and inside a group the
//     nodes end up in reverse order of the walk.
// So every node gets a pair of integers (G, W), both "minus the time of the event that placed it", such that list
// order == ascending (G, W): G is the key of its bucket group (valid while the bucket stays non-empty, hence for
// the node's whole life), W its place in the group.  Which intervals leave the map after word w is known directly
// (end word == w - gap - 1), so a flush sorts just the candidates among them by (G, W).  Per word, the creation order
// (seed-map iteration order, then (a, b)) is independent of other words once the seed map's bucket count at the
// start of the word is known, and is computed on all host threads.
// ---------------------------------------------------------------------------------------------------------------
// Sort on all host threads (sample sort): splitters from a sorted sample cut the key range into one part per thread;
// every thread classifies its slice of the input, the parts are gathered by a counting scatter and sorted
// independently.  No merge rounds, whose last ones would run on one or two threads.
template <class T, class Less>
void parallelSort(std::vector<T>& v, Less less, unsigned threads = 0, std::vector<T>* reusable = nullptr)
{
  if (threads == 0) {
    threads = std::max(1u, std::thread::hardware_concurrency());
  }
  const size_t n = v.size();
  const size_t parts = std::min<size_t>({threads, n / (size_t{1} << 16), size_t{255}});  // part ids are bytes
  if (parts < 2) {
    std::sort(v.begin(), v.end(), less);
    return;
  }
  auto runAll = [&](const size_t jobs, auto&& body) {
    std::vector<std::thread> pool;
    for (size_t j = 1; j < jobs; ++j) {
      pool.emplace_back(body, j);
    }
    body(size_t{0});
    for (auto& th : pool) {
      th.join();
    }
  };
  // splitters: every (sample / parts)-th element of a sorted, evenly spaced sample
  const size_t sampleSize = std::min<size_t>(n, parts * 256);
  std::vector<T> sample(sampleSize);
  for (size_t i = 0; i < sampleSize; ++i) {
    sample[i] = v[n / sampleSize * i];
  }
  std::sort(sample.begin(), sample.end(), less);
  std::vector<T> splitter(parts - 1);
  for (size_t p = 1; p < parts; ++p) {
    splitter[p - 1] = sample[sampleSize * p / parts];
  }
  auto partOf = [&](const T& x) {  // number of splitters <= x
    return static_cast<size_t>(std::upper_bound(splitter.begin(), splitter.end(), x, less) - splitter.begin());
  };
  auto sliceBound = [&](const size_t t) { return n * t / parts; };
  std::vector<std::vector<size_t>> count(parts, std::vector<size_t>(parts, 0));
  std::vector<uint8_t> partId(n);
  static_assert(sizeof(uint8_t) == 1, "");
  runAll(parts, [&](const size_t t) {
    for (size_t i = sliceBound(t); i < sliceBound(t + 1); ++i) {
      const size_t p = partOf(v[i]);
      partId[i] = static_cast<uint8_t>(p);
      ++count[t][p];
    }
  });
  std::vector<size_t> partBegin(parts + 1, 0);
  {
    size_t run = 0;
    for (size_t p = 0; p < parts; ++p) {
      partBegin[p] = run;
      for (size_t t = 0; t < parts; ++t) {
        const size_t c = count[t][p];
        count[t][p] = run;  // where slice t writes its first element of part p
        run += c;
      }
    }
    partBegin[parts] = run;
  }
  std::vector<T> local;
  std::vector<T>& out = reusable ? *reusable : local;  // fresh memory is slow to fault in: callers that sort repeatedly keep it
  out.resize(n);
  runAll(parts, [&](const size_t t) {
    std::vector<size_t>& cursor = count[t];
    for (size_t i = sliceBound(t); i < sliceBound(t + 1); ++i) {
      out[cursor[partId[i]]++] = v[i];
    }
  });
  runAll(parts, [&](const size_t p) { std::sort(out.begin() + partBegin[p], out.begin() + partBegin[p + 1], less); });
  v.swap(out);
}

// Array of trivially constructible elements that is NOT zero-filled: every element is written before it is read, and
// the first touch of a multi-GB buffer is better done by the threads that fill it than by one zeroing pass.
template <class T> class RawArray
{
public:
  RawArray() = default;
  explicit RawArray(const size_t n) { reset(n); }
  void reset(const size_t n)
  {
    mData.reset(new T[std::max<size_t>(n, 1)]);  // default-initialised: no fill for plain types
    mSize = n;
  }
  size_t size() const { return mSize; }
  T* data() { return mData.get(); }
  const T* data() const { return mData.get(); }
  T& operator[](const size_t i) { return mData[i]; }
  const T& operator[](const size_t i) const { return mData[i]; }

private:
  std::unique_ptr<T[]> mData;
  size_t mSize = 0;
};

// Zero-filled array for a table that is hit at random: backed by transparent huge pages where the system allows it
// (a 25 MB table on 4 KB pages misses the TLB on nearly every access, and page walks are slow under virtualisation).
template <class T> class HugeArray
{
public:
  explicit HugeArray(const size_t count) { reset(count); }
  HugeArray(const HugeArray&) = delete;
  HugeArray& operator=(const HugeArray&) = delete;
  ~HugeArray() { release(); }
  void reset(const size_t count)
  {
    release();
    constexpr size_t kHuge = size_t{2} << 20;
    mBytes = (std::max<size_t>(count, 1) * sizeof(T) + kHuge - 1) / kHuge * kHuge;
    void* p = mmap(nullptr, mBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
      throw std::bad_alloc();
    }
    madvise(p, mBytes, MADV_HUGEPAGE);  // advisory; anonymous mappings are zero-filled either way
    mData = static_cast<T*>(p);
  }
  T& operator[](const size_t i) { return mData[i]; }
  const T& operator[](const size_t i) const { return mData[i]; }

private:
  void release()
  {
    if (mData) {
      munmap(mData, mBytes);
      mData = nullptr;
    }
  }
  T* mData = nullptr;
  size_t mBytes = 0;
};

// Stable counting sort of the items 0..n-1 by key(i) in [0, numKeys): out[...] = item indices grouped by key, ascending
// inside a group; begin[k] = first position of key k.  The item range is cut into one slice per thread, every slice
// counts and scatters on its own.
template <class KeyFn>
void groupByKey(const int64_t n, const int numKeys, unsigned threads, KeyFn&& key, std::vector<int64_t>& begin,
                RawArray<int64_t>& out)
{
  if (threads == 0) {
    threads = std::max(1u, std::thread::hardware_concurrency());
  }
  const int T = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(threads, n / (1 << 16) + 1)));
  std::vector<std::vector<int64_t>> hist(static_cast<size_t>(T), std::vector<int64_t>(static_cast<size_t>(numKeys), 0));
  auto slice = [&](const int t) { return std::make_pair(n * t / T, n * (t + 1) / T); };
  auto run = [&](auto&& body) {
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) {
      pool.emplace_back(body, t);
    }
    body(0);
    for (auto& th : pool) {
      th.join();
    }
  };
  run([&](const int t) {
    const auto [lo, hi] = slice(t);
    for (int64_t i = lo; i < hi; ++i) {
      ++hist[static_cast<size_t>(t)][static_cast<size_t>(key(i))];
    }
  });
  begin.assign(static_cast<size_t>(numKeys) + 1, 0);
  int64_t runTotal = 0;
  for (int k = 0; k < numKeys; ++k) {
    begin[static_cast<size_t>(k)] = runTotal;
    for (int t = 0; t < T; ++t) {
      const int64_t c = hist[static_cast<size_t>(t)][static_cast<size_t>(k)];
      hist[static_cast<size_t>(t)][static_cast<size_t>(k)] = runTotal;  // now: where slice t writes its first item of key k
      runTotal += c;
    }
  }
  begin[static_cast<size_t>(numKeys)] = runTotal;
  out.reset(static_cast<size_t>(n));
  run([&](const int t) {
    const auto [lo, hi] = slice(t);
    std::vector<int64_t>& cursor = hist[static_cast<size_t>(t)];
    for (int64_t i = lo; i < hi; ++i) {
      out[static_cast<size_t>(cursor[static_cast<size_t>(key(i))]++)] = i;
    }
  });
}

template <class Intervals, class WordFn, class LengthFn, class EmitFn>
void replayReferenceOrderFast(const Intervals& intervals, const uint32_t numHaps, const int numWords,
                              const int gap, WordFn&& rawWord, LengthFn&& longEnough, EmitFn&& emit,
                              const unsigned threads = 0)
{
  const int64_t n = static_cast<int64_t>(intervals.size());
  if (numWords <= 0) {
    return;
  }
  const bool trace = std::getenv("FSMC_TRACE") != nullptr;  // development: host time of the phases
  auto clock = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tPhase = clock();
  auto lap = [&](const char* what) {
    if (trace) {
      const double t = clock();
      std::fprintf(stderr, "replayReferenceOrderFast: %-28s %.3f s\n", what, t - tPhase);
      tPhase = t;
    }
  };
  // ---- intervals grouped by start word ---------------------------------------------------------------------------
  std::vector<int64_t> startBegin, endBegin;
  RawArray<int64_t> byStart;
  groupByKey(n, numWords, threads, [&](const int64_t i) { return intervals[static_cast<size_t>(i)].startWord; }, startBegin,
             byStart);
  if (trace) {
    long long sumLen = 0, longOnes = 0, maxLen = 0;
    for (int64_t i = 0; i < n; ++i) {
      const long long len = intervals[static_cast<size_t>(i)].endWord - intervals[static_cast<size_t>(i)].startWord + 1;
      sumLen += len;
      longOnes += len > 50;
      maxLen = std::max(maxLen, len);
    }
    std::fprintf(stderr, "replayReferenceOrderFast: %lld intervals, mean length %.2f words, %lld longer than 50 words, longest %lld; "
                 "starts in word 0: %lld, in word %d: %lld\n",
                 static_cast<long long>(n), n ? static_cast<double>(sumLen) / n : 0.0, longOnes, maxLen,
                 static_cast<long long>(startBegin[1] - startBegin[0]), numWords / 2,
                 static_cast<long long>(startBegin[numWords / 2 + 1] - startBegin[numWords / 2]));
  }
  lap("group by start word");
  // ---- phase 1: creation order of each word's new intervals -----------------------------------------------------
  // bucket count of the seed map at the start of every word: it only grows, by the number of distinct keys
  std::vector<size_t> distinct(static_cast<size_t>(numWords), 0);
  parallelForWords(numWords, threads, [&](const int w) {
    std::vector<uint64_t> keys(numHaps);
    for (uint32_t h = 0; h < numHaps; ++h) {
      keys[h] = rawWord(h, w);
    }
    std::sort(keys.begin(), keys.end());
    distinct[w] = static_cast<size_t>(std::unique(keys.begin(), keys.end()) - keys.begin());
  });
  std::vector<size_t> seedBuckets(static_cast<size_t>(numWords), 17);
  {
    size_t buckets = 17;
    for (int w = 0; w < numWords; ++w) {
      seedBuckets[w] = buckets;
      buckets = NodeOrderMap::bucketsAfter(buckets, distinct[w]);
    }
  }
  RawArray<fsmc_match> ordered(static_cast<size_t>(n));
  lap("seed-map bucket counts");
  // A low-complexity word (a stretch of rare SNPs where most haplotypes are identical) starts millions of intervals
  // at once: such words are taken one at a time with the sort itself on all threads, the others one word per thread.
  int64_t kBigWord = int64_t{1} << 21;
  if (const char* e = std::getenv("FSMC_BIG_WORD")) {  // tests: exercise the whole-machine path on small data
    kBigWord = std::max<int64_t>(1, std::atoll(e));
  }
  auto creationOrderOfWord = [&](const int w, const bool wholeMachine) {
    const int64_t lo = startBegin[w], hi = startBegin[w + 1];
    if (lo == hi || ((hi - lo >= kBigWord) != wholeMachine)) {
      return;
    }
    // node of every haplotype's word in the seed map, then the node's place in the map's iteration order
    NodeOrderMap seeds(seedBuckets[w]);
    std::vector<int> nodeOfHap(numHaps);
    for (uint32_t h = 0; h < numHaps; ++h) {
      bool isNew;
      nodeOfHap[h] = seeds.insert(rawWord(h, w), 0, isNew);
    }
    std::vector<int64_t> rankOfNode(numHaps, 0);
    int64_t r = 0;
    for (int nd = seeds.first(); nd != NodeOrderMap::kEnd; nd = seeds.next(nd)) {
      rankOfNode[static_cast<size_t>(nd)] = r++;
    }
    struct Creation {
      int64_t rank;
      uint32_t a, b;
      int64_t index;
      int32_t endWord;
    };
    std::vector<Creation> created;
    created.reserve(static_cast<size_t>(hi - lo));
    for (int64_t q = lo; q < hi; ++q) {
      const int64_t i = byStart[static_cast<size_t>(q)];
      const fsmc_match& m = intervals[static_cast<size_t>(i)];
      created.push_back(Creation{rankOfNode[static_cast<size_t>(nodeOfHap[m.hapA])], m.hapA, m.hapB, i, m.endWord});
    }
    auto less = [](const Creation& x, const Creation& y) {
      return x.rank != y.rank ? x.rank < y.rank : (x.a != y.a ? x.a < y.a : x.b < y.b);
    };
    if (wholeMachine) {
      parallelSort(created, less, threads);
    } else {
      std::sort(created.begin(), created.end(), less);
    }
    // From here on an interval is named by its creation rank q (its position in byStart): everything the sequential
    // phase touches is then laid out in the order it is visited.
    for (int64_t q = lo; q < hi; ++q) {
      const Creation& c = created[static_cast<size_t>(q - lo)];
      byStart[static_cast<size_t>(q)] = c.index;
      ordered[static_cast<size_t>(q)] = fsmc_match{c.a, c.b, w, c.endWord};
    }
  };
  parallelForWords(numWords, threads, [&](const int w) { creationOrderOfWord(w, false); });
  for (int w = 0; w < numWords; ++w) {
    creationOrderOfWord(w, true);
  }
  lap("creation order per word");
  // byEnd lists the creation ranks by end word, ascending
  RawArray<int64_t> byEnd;
  groupByKey(n, numWords, threads, [&](const int64_t q) { return ordered[static_cast<size_t>(q)].endWord; }, endBegin, byEnd);
  lap("reorder, group by end word");
  // ---- phase 2: the extend map's node order as (G, W) keys ------------------------------------------------------
  // G lives in the bucket table (valid while the bucket is non-empty, so it can be read whenever one of the bucket's
  // nodes is still in the map); W and the bucket index are kept per node.
  RawArray<int64_t> nodeW(static_cast<size_t>(n));
  RawArray<uint32_t> nodeBucket(static_cast<size_t>(n));
  // 8 bytes per bucket (the table reaches 2 x 10^8 buckets at biobank density): the group key as a positive magnitude
  // in the upper 40 bits, the number of the bucket's nodes in the map in the lower 24
  struct BucketState {
    uint64_t bits;
    int64_t G() const { return -static_cast<int64_t>(bits >> 24); }  // valid while live() > 0
    void setG(const int64_t g) { bits = (static_cast<uint64_t>(-g) << 24) | (bits & 0xffffffull); }
    uint64_t live() const { return bits & 0xffffffull; }
    void addNode()
    {
      if (live() == 0xffffffull) {
        throw std::runtime_error("replayReferenceOrderFast: more than 2^24 nodes in one bucket");
      }
      ++bits;
    }
    void dropNode() { --bits; }
  };
  size_t buckets = 17, count = 0;
  HugeArray<BucketState> bucket(buckets);
  int64_t tick = 1;
  constexpr int64_t kAhead = 24;  // software prefetch distance: the bucket table (tens of MB) is hit at random
  auto pairKey = [&](const int64_t q) {
    return static_cast<uint64_t>(ordered[static_cast<size_t>(q)].hapA) * numHaps + ordered[static_cast<size_t>(q)].hapB;
  };
  // key % buckets without a hardware divide: floor(key * ceil(2^64 / d) / 2^64) is the quotient or one more
  uint64_t magic = 0;
  auto setBuckets = [&](const size_t d) {
    buckets = d;
    magic = static_cast<uint64_t>((static_cast<unsigned __int128>(1) << 64) / d) + 1;
  };
  auto bucketOf = [&](const uint64_t key) {
    const uint64_t quot = static_cast<uint64_t>((static_cast<unsigned __int128>(key) * magic) >> 64);
    int64_t r = static_cast<int64_t>(key - quot * buckets);
    if (r < 0) {
      r += static_cast<int64_t>(buckets);
    }
    return static_cast<size_t>(r);
  };
  setBuckets(17);
  struct Placed {
    int64_t G, W, q;
    bool operator<(const Placed& o) const { return G != o.G ? G < o.G : W < o.W; }
  };
  auto placeOf = [&](const int64_t q) {
    return Placed{bucket[nodeBucket[static_cast<size_t>(q)]].G(), nodeW[static_cast<size_t>(q)], q};
  };
  std::vector<Placed> scratch;
  // nodes alive while word w's intervals are being inserted: created so far (rank < upTo), end word >= w - gap - 1
  double tRehash = 0;
  int64_t numRehash = 0, rehashScanned = 0;
  size_t maxLive = 0;
  // A rehash walks every live node in list order.  With 10^8 live nodes that is the expensive part of the replay, so
  // each step runs on all threads: collecting the live nodes (slices of the creation ranks), sorting them, and
  // re-keying, where every thread owns a contiguous range of the NEW buckets and handles the nodes that fall into it
  // (a bucket's group key is set by its first node in walk order; each thread meets its nodes in walk order).
  const unsigned nThreads = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
  auto runThreads = [&](const unsigned jobs, auto&& body) {
    std::vector<std::thread> pool;
    for (unsigned j = 1; j < jobs; ++j) {
      pool.emplace_back(body, j);
    }
    body(0u);
    for (auto& th : pool) {
      th.join();
    }
  };
  std::vector<uint32_t> newBucketOf;
  std::vector<Placed> sortBuffer;
  auto rehash = [&](const size_t newBuckets, const int w, const int64_t upTo) {
    const double r0 = trace ? clock() : 0;
    ++numRehash;
    rehashScanned += upTo;
    const int minEnd = w - gap - 1;
    const unsigned T = upTo < std::min<int64_t>(int64_t{1} << 18, kBigWord) ? 1u : nThreads;
    {
      // two passes (count, then write in place): no per-thread vectors to grow and copy
      std::vector<int64_t> found(T + 1, 0);
      runThreads(T, [&](const unsigned t) {
        int64_t c = 0;
        for (int64_t q = upTo * t / T; q < upTo * (t + 1) / T; ++q) {
          c += ordered[static_cast<size_t>(q)].endWord >= minEnd;
        }
        found[t + 1] = c;
      });
      for (unsigned t = 0; t < T; ++t) {
        found[t + 1] += found[t];
      }
      scratch.resize(static_cast<size_t>(found[T]));
      runThreads(T, [&](const unsigned t) {
        int64_t at = found[t];
        for (int64_t q = upTo * t / T; q < upTo * (t + 1) / T; ++q) {
          if (ordered[static_cast<size_t>(q)].endWord >= minEnd) {
            scratch[static_cast<size_t>(at++)] = placeOf(q);
          }
        }
      });
    }
    parallelSort(scratch, std::less<Placed>(), threads, &sortBuffer);
    setBuckets(newBuckets);
    bucket.reset(buckets);
    const int64_t N = static_cast<int64_t>(scratch.size());
    newBucketOf.resize(static_cast<size_t>(N));
    runThreads(T, [&](const unsigned t) {
      for (int64_t e = N * t / T; e < N * (t + 1) / T; ++e) {
        newBucketOf[static_cast<size_t>(e)] = static_cast<uint32_t>(bucketOf(pairKey(scratch[static_cast<size_t>(e)].q)));
      }
    });
    const int64_t tick0 = tick;
    runThreads(T, [&](const unsigned t) {
      const size_t bLo = buckets * t / T, bHi = buckets * (t + 1) / T;
      for (int64_t e = 0; e < N; ++e) {
        const size_t b = newBucketOf[static_cast<size_t>(e)];
        if (b < bLo || b >= bHi) {
          continue;
        }
        const int64_t q = scratch[static_cast<size_t>(e)].q;
        if (bucket[b].live() == 0) {
          bucket[b].setG(-(tick0 + N - e));  // groups in the order their first node is met
        }
        bucket[b].addNode();
        nodeBucket[static_cast<size_t>(q)] = static_cast<uint32_t>(b);
        nodeW[static_cast<size_t>(q)] = -(tick0 + e);  // inside a group: reverse order of the walk
      }
    });
    tick += N + 1;
    tRehash += trace ? clock() - r0 : 0;
  };
  std::vector<Placed> leaving;
  auto flushSet = [&] {
    std::sort(leaving.begin(), leaving.end());
    for (const Placed& p : leaving) {
      emit(byStart[static_cast<size_t>(p.q)]);
    }
    leaving.clear();
  };
  double tIns = 0, tErase = 0, tFlush = 0;
  for (int w = 0; w < numWords; ++w) {
    const double c0 = trace ? clock() : 0;
    for (int64_t q = startBegin[w]; q < startBegin[w + 1]; ++q) {
      if (count + 1 > buckets) {  // max load factor 1.0
        const size_t want = NodeOrderMap::growTo(count);
        if (want != buckets) {
          rehash(want, w, q);
        }
      }
      if (q + kAhead < n) {
        __builtin_prefetch(&bucket[bucketOf(pairKey(q + kAhead))], 1);
      }
      const size_t b = bucketOf(pairKey(q));
      BucketState& st = bucket[b];
      if (st.live() == 0) {
        st.setG(-tick);  // a new group goes to the front of the list
      }
      st.addNode();
      ++count;
      nodeBucket[static_cast<size_t>(q)] = static_cast<uint32_t>(b);
      nodeW[static_cast<size_t>(q)] = -tick;  // front of its group
      ++tick;
    }
    maxLive = std::max(maxLive, count);
    const double c1 = trace ? clock() : 0;
    tIns += c1 - c0;
    // ExtendHash::clearPairsPriorTo(w - gap): exactly the intervals that ended at word w - gap - 1 leave now
    const int e = w - gap - 1;
    if (e >= 0) {
      for (int64_t k = endBegin[e]; k < endBegin[e + 1]; ++k) {
        const int64_t q = byEnd[static_cast<size_t>(k)];
        if (k + kAhead < endBegin[e + 1]) {
          __builtin_prefetch(&bucket[nodeBucket[static_cast<size_t>(byEnd[static_cast<size_t>(k + kAhead)])]], 1);
        }
        BucketState& st = bucket[nodeBucket[static_cast<size_t>(q)]];
        if (longEnough(ordered[static_cast<size_t>(q)])) {
          leaving.push_back(Placed{st.G(), nodeW[static_cast<size_t>(q)], q});  // G stays readable until the next insert
        }
        st.dropNode();
        --count;
      }
      const double c2 = trace ? clock() : 0;
      tErase += c2 - c1;
      flushSet();
      tFlush += trace ? clock() - c2 : 0;
    }
  }
  if (trace) {
    std::fprintf(stderr, "replayReferenceOrderFast:   inserts %.3f s (of which %lld rehashes over %lld nodes: %.3f s), erases %.3f s, "
                 "sorted emission %.3f s; most nodes in the map %zu, buckets %zu\n",
                 tIns, static_cast<long long>(numRehash), static_cast<long long>(rehashScanned), tRehash, tErase, tFlush, maxLive, buckets);
  }
  // ExtendHash::clearAllPairs: everything still in the map, in list order
  for (int e = std::max(0, numWords - gap - 1); e < numWords; ++e) {
    for (int64_t k = endBegin[e]; k < endBegin[e + 1]; ++k) {
      const int64_t q = byEnd[static_cast<size_t>(k)];
      if (longEnough(ordered[static_cast<size_t>(q)])) {
        leaving.push_back(placeOf(q));
      }
    }
  }
  flushSet();
  lap("node keys, flushes, emission");
}

}  // namespace candidate_order
