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
#include <cstdint>
#include <thread>
#include <utility>
#include <vector>

#include "../../../include/fastsmc_b200.h"

namespace candidate_order
{

// Singly linked node list + per-bucket "node before my first node" pointers, as boost::unordered's
// table implementation keeps them.  Only what the replay needs: insert, erase-while-iterating, clear.
class NodeOrderMap
{
public:
  static constexpr int kEnd = -1;

  NodeOrderMap() = default;
  /// a map whose bucket array was already grown to `buckets` by earlier use (unordered_map::clear keeps it)
  explicit NodeOrderMap(const size_t buckets) { allocateBuckets(buckets); }
  size_t bucketCount() const { return mBuckets; }
  /// bucket count after inserting `distinct` new keys into an EMPTY map that currently has `buckets` buckets
  static size_t bucketsAfter(size_t buckets, const size_t distinct)
  {
    for (size_t count = 0; count < distinct; ++count) {
      if (count + 1 > buckets) {
        buckets = primeAtLeast(std::max(count + 1, count + (count >> 1)) + 1);
      }
    }
    return buckets;
  }
  static size_t growTo(const size_t count) { return primeAtLeast(std::max(count + 1, count + (count >> 1)) + 1); }

  int insert(const uint64_t key, const int64_t payload, bool& isNew)
  {
    if (!mBefore.empty()) {
      const size_t b = key % mBuckets;
      if (mBefore[b] != kNone) {
        for (int n = follower(mBefore[b]); n != kEnd && mNodes[n].bucket == b; n = mNodes[n].next) {
          if (mNodes[n].key == key) {
            isNew = false;
            return n;
          }
        }
      }
    }
    isNew = true;
    if (mBefore.empty()) {
      allocateBuckets(std::max(mBuckets, primeAtLeast(mCount + 2)));
    } else if (mCount + 1 > mBuckets) {  // max load factor 1.0
      const size_t want = primeAtLeast(std::max(mCount + 1, mCount + (mCount >> 1)) + 1);
      if (want != mBuckets) {
        rebucket(want);
      }
    }
    int id;
    if (!mSpare.empty()) {
      id = mSpare.back();
      mSpare.pop_back();
    } else {
      id = static_cast<int>(mNodes.size());
      mNodes.emplace_back();
    }
    const size_t b = key % mBuckets;
    Node& node = mNodes[id];
    node.key = key;
    node.payload = payload;
    node.bucket = b;
    if (mBefore[b] == kNone) {
      // first node of an empty bucket goes to the front of the whole list
      if (mFirst != kEnd) {
        mBefore[mNodes[mFirst].bucket] = id;
      }
      mBefore[b] = kFront;
      node.next = mFirst;
      mFirst = id;
    } else {
      node.next = follower(mBefore[b]);
      follower(mBefore[b]) = id;
    }
    ++mCount;
    return id;
  }

  // unlink node n; returns the node after it
  int erase(const int n)
  {
    const size_t b = mNodes[n].bucket;
    int prev = mBefore[b];
    while (follower(prev) != n) {
      prev = follower(prev);
    }
    const int after = mNodes[n].next;
    follower(prev) = after;
    --mCount;
    bool bucketContinues = false;
    if (after != kEnd) {
      if (mNodes[after].bucket == b) {
        bucketContinues = true;
      } else {
        mBefore[mNodes[after].bucket] = prev;
      }
    }
    if (!bucketContinues && mBefore[b] == prev) {
      mBefore[b] = kNone;
    }
    mSpare.push_back(n);
    return after;
  }

  // frees the nodes, keeps the bucket array at its grown size (as unordered_map::clear does)
  void clear()
  {
    if (mCount == 0) {
      return;
    }
    std::fill(mBefore.begin(), mBefore.end(), kNone);
    mNodes.clear();
    mSpare.clear();
    mFirst = kEnd;
    mCount = 0;
  }

  int first() const { return mFirst; }
  int next(const int n) const { return mNodes[n].next; }
  uint64_t key(const int n) const { return mNodes[n].key; }
  int64_t payload(const int n) const { return mNodes[n].payload; }
  size_t size() const { return mCount; }

private:
  static constexpr int kFront = -2, kNone = -3;
  struct Node {
    uint64_t key = 0;
    int64_t payload = 0;
    size_t bucket = 0;
    int next = kEnd;
  };
  std::vector<Node> mNodes;
  std::vector<int> mSpare;
  std::vector<int> mBefore;
  size_t mBuckets = 17;  // default-constructed map: next prime >= 11, allocated on first insert
  size_t mCount = 0;
  int mFirst = kEnd;

  int& follower(const int p) { return p == kFront ? mFirst : mNodes[p].next; }

  static size_t primeAtLeast(const size_t n)
  {
    static const size_t primes[] = {17ul,         29ul,         37ul,        53ul,        67ul,        79ul,
                                    97ul,         131ul,        193ul,       257ul,       389ul,       521ul,
                                    769ul,        1031ul,       1543ul,      2053ul,      3079ul,      6151ul,
                                    12289ul,      24593ul,      49157ul,     98317ul,     196613ul,    393241ul,
                                    786433ul,     1572869ul,    3145739ul,   6291469ul,   12582917ul,  25165843ul,
                                    50331653ul,   100663319ul,  201326611ul, 402653189ul, 805306457ul, 1610612741ul,
                                    3221225473ul, 4294967291ul};
    for (const size_t p : primes) {
      if (p >= n) {
        return p;
      }
    }
    return primes[sizeof(primes) / sizeof(primes[0]) - 1];
  }
  void allocateBuckets(const size_t count)
  {
    mBuckets = count;
    mBefore.assign(count, kNone);
  }
  // walk the list once; a node whose new bucket is still empty stays in place, others are spliced to the
  // front of their bucket's run
  void rebucket(const size_t count)
  {
    allocateBuckets(count);
    int prev = kFront;
    while (follower(prev) != kEnd) {
      const int n = follower(prev);
      const size_t b = mNodes[n].key % mBuckets;
      mNodes[n].bucket = b;
      if (mBefore[b] == kNone) {
        mBefore[b] = prev;
        prev = n;
      } else {
        const int after = mNodes[n].next;
        mNodes[n].next = follower(mBefore[b]);
        follower(mBefore[b]) = n;
        follower(prev) = after;
      }
    }
  }
};

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
template <class Fn> void parallelForWords(const int numWords, unsigned threads, Fn&& fn)
{
  if (threads == 0) {
    threads = std::max(1u, std::thread::hardware_concurrency());
  }
  threads = std::min<unsigned>(threads, static_cast<unsigned>(std::max(1, numWords)));
  std::atomic<int> next{0};
  auto work = [&] {
    for (int w = next++; w < numWords; w = next++) {
      fn(w);
    }
  };
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < threads; ++t) {
    pool.emplace_back(work);
  }
  work();
  for (auto& th : pool) {
    th.join();
  }
}

template <class WordFn, class LengthFn, class EmitFn>
void replayReferenceOrderFast(const std::vector<fsmc_match>& intervals, const uint32_t numHaps, const int numWords,
                              const int gap, WordFn&& rawWord, LengthFn&& longEnough, EmitFn&& emit,
                              const unsigned threads = 0)
{
  const int64_t n = static_cast<int64_t>(intervals.size());
  if (numWords <= 0) {
    return;
  }
  // ---- intervals grouped by start word and by end word (counting sorts) ------------------------------------
  std::vector<int64_t> startBegin(static_cast<size_t>(numWords) + 1, 0), endBegin(static_cast<size_t>(numWords) + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    ++startBegin[static_cast<size_t>(intervals[i].startWord) + 1];
    ++endBegin[static_cast<size_t>(intervals[i].endWord) + 1];
  }
  for (int w = 0; w < numWords; ++w) {
    startBegin[w + 1] += startBegin[w];
    endBegin[w + 1] += endBegin[w];
  }
  std::vector<int64_t> byStart(static_cast<size_t>(n)), byEnd(static_cast<size_t>(n));
  {
    std::vector<int64_t> cs(startBegin.begin(), startBegin.end() - 1), ce(endBegin.begin(), endBegin.end() - 1);
    for (int64_t i = 0; i < n; ++i) {
      byStart[static_cast<size_t>(cs[intervals[i].startWord]++)] = i;
      byEnd[static_cast<size_t>(ce[intervals[i].endWord]++)] = i;
    }
  }

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
  parallelForWords(numWords, threads, [&](const int w) {
    const int64_t lo = startBegin[w], hi = startBegin[w + 1];
    if (lo == hi) {
      return;
    }
    NodeOrderMap seeds(seedBuckets[w]);
    for (uint32_t h = 0; h < numHaps; ++h) {
      bool isNew;
      seeds.insert(rawWord(h, w), 0, isNew);
    }
    std::vector<std::pair<uint64_t, int64_t>> rank;
    rank.reserve(seeds.size());
    int64_t r = 0;
    for (int nd = seeds.first(); nd != NodeOrderMap::kEnd; nd = seeds.next(nd)) {
      rank.emplace_back(seeds.key(nd), r++);
    }
    std::sort(rank.begin(), rank.end());
    struct Creation {
      int64_t rank;
      uint32_t a, b;
      int64_t index;
    };
    std::vector<Creation> created;
    created.reserve(static_cast<size_t>(hi - lo));
    for (int64_t q = lo; q < hi; ++q) {
      const int64_t i = byStart[static_cast<size_t>(q)];
      const fsmc_match& m = intervals[static_cast<size_t>(i)];
      const auto it = std::lower_bound(rank.begin(), rank.end(), std::make_pair(rawWord(m.hapA, w), int64_t{0}));
      created.push_back(Creation{it->second, m.hapA, m.hapB, i});
    }
    std::sort(created.begin(), created.end(), [](const Creation& x, const Creation& y) {
      return x.rank != y.rank ? x.rank < y.rank : (x.a != y.a ? x.a < y.a : x.b < y.b);
    });
    for (int64_t q = lo; q < hi; ++q) {
      byStart[static_cast<size_t>(q)] = created[static_cast<size_t>(q - lo)].index;
    }
  });

  // ---- phase 2: the extend map's node order as (G, W) keys ------------------------------------------------------
  std::vector<int64_t> nodeG(static_cast<size_t>(n)), nodeW(static_cast<size_t>(n));
  std::vector<uint32_t> nodeBucket(static_cast<size_t>(n));
  size_t buckets = 17, count = 0;
  std::vector<int32_t> live(buckets, 0);
  std::vector<int64_t> bucketG(buckets, 0);
  int64_t tick = 1;
  auto pairKey = [&](const int64_t i) {
    return static_cast<uint64_t>(intervals[static_cast<size_t>(i)].hapA) * numHaps + intervals[static_cast<size_t>(i)].hapB;
  };
  auto byListOrder = [&](const int64_t x, const int64_t y) {
    return nodeG[static_cast<size_t>(x)] != nodeG[static_cast<size_t>(y)] ? nodeG[static_cast<size_t>(x)] < nodeG[static_cast<size_t>(y)]
                                                                           : nodeW[static_cast<size_t>(x)] < nodeW[static_cast<size_t>(y)];
  };
  std::vector<int64_t> scratch;
  // nodes alive while word w's intervals are being inserted: inserted so far, end word >= w - gap - 1
  auto rehash = [&](const size_t newBuckets, const int w, const int64_t insertedOfW) {
    scratch.clear();
    const int minEnd = w - gap - 1;
    for (int v = 0; v <= w; ++v) {
      const int64_t hi = v < w ? startBegin[v + 1] : startBegin[v] + insertedOfW;
      for (int64_t q = startBegin[v]; q < hi; ++q) {
        const int64_t i = byStart[static_cast<size_t>(q)];
        if (intervals[static_cast<size_t>(i)].endWord >= minEnd) {
          scratch.push_back(i);
        }
      }
    }
    std::sort(scratch.begin(), scratch.end(), byListOrder);
    buckets = newBuckets;
    live.assign(buckets, 0);
    bucketG.assign(buckets, 0);
    const int64_t N = static_cast<int64_t>(scratch.size());
    for (int64_t e = 0; e < N; ++e) {
      const int64_t i = scratch[static_cast<size_t>(e)];
      const size_t b = static_cast<size_t>(pairKey(i) % buckets);
      if (live[b] == 0) {
        bucketG[b] = -(tick + N - e);  // groups in the order their first node is met
      }
      ++live[b];
      nodeBucket[static_cast<size_t>(i)] = static_cast<uint32_t>(b);
      nodeG[static_cast<size_t>(i)] = bucketG[b];
      nodeW[static_cast<size_t>(i)] = -(tick + e);  // inside a group: reverse order of the walk
    }
    tick += N + 1;
  };
  auto flushSet = [&](std::vector<int64_t>& leaving) {
    std::sort(leaving.begin(), leaving.end(), byListOrder);
    for (const int64_t i : leaving) {
      emit(i);
    }
    leaving.clear();
  };
  std::vector<int64_t> leaving;
  for (int w = 0; w < numWords; ++w) {
    for (int64_t q = startBegin[w]; q < startBegin[w + 1]; ++q) {
      const int64_t i = byStart[static_cast<size_t>(q)];
      if (count + 1 > buckets) {  // max load factor 1.0
        const size_t want = NodeOrderMap::growTo(count);
        if (want != buckets) {
          rehash(want, w, q - startBegin[w]);
        }
      }
      const size_t b = static_cast<size_t>(pairKey(i) % buckets);
      if (live[b] == 0) {
        bucketG[b] = -tick;  // a new group goes to the front of the list
      }
      ++live[b];
      ++count;
      nodeBucket[static_cast<size_t>(i)] = static_cast<uint32_t>(b);
      nodeG[static_cast<size_t>(i)] = bucketG[b];
      nodeW[static_cast<size_t>(i)] = -tick;  // front of its group
      ++tick;
    }
    // ExtendHash::clearPairsPriorTo(w - gap): exactly the intervals that ended at word w - gap - 1 leave now
    const int e = w - gap - 1;
    if (e >= 0) {
      for (int64_t q = endBegin[e]; q < endBegin[e + 1]; ++q) {
        const int64_t i = byEnd[static_cast<size_t>(q)];
        --live[nodeBucket[static_cast<size_t>(i)]];
        --count;
        if (longEnough(intervals[static_cast<size_t>(i)])) {
          leaving.push_back(i);
        }
      }
      flushSet(leaving);
    }
  }
  // ExtendHash::clearAllPairs: everything still in the map, in list order
  for (int e = std::max(0, numWords - gap - 1); e < numWords; ++e) {
    for (int64_t q = endBegin[e]; q < endBegin[e + 1]; ++q) {
      const int64_t i = byEnd[static_cast<size_t>(q)];
      if (longEnough(intervals[static_cast<size_t>(i)])) {
        leaving.push_back(i);
      }
    }
  }
  flushSet(leaving);
}

}  // namespace candidate_order
