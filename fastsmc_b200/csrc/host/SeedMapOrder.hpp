// fastsmc_b200 — node order of the reference's hash maps (boost::unordered_map <= 1.79, integer keys): the literal
// linked-list model, and the seed map's per-word iteration ranks.  Header-only and free of CUDA / OS specifics: it is
// used by the host layer (CandidateOrder.hpp) and inside libfastsmc_b200 (fsmc_seed with FSMC_SEED_REFERENCE_ORDER).
#pragma once

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <thread>
#include <vector>

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

// Seed-map iteration rank of every haplotype's word group, for every word: rank[w * numHaps + h] = position, in the
// iteration order of the reference's SeedHash map after word w's insertIndividuals calls (ref: FastSMC.cpp:204-206,
// HASHING/SeedHash.hpp:41-45, 80), of the bucket that holds haplotype h.  Within a word the reference visits buckets in
// this order and enumerates a bucket's pairs (i < ii) in haplotype order, so (rank, a, b) is the creation order of the
// word's new match intervals.  The seed map keeps its grown bucket count from word to word (unordered_map::clear), which
// is the only coupling between words; the words are then independent and run on all host threads.
//
// max_seeds > 0 (ref: HASHING/SeedHash.hpp:56-69, 85-93): an oversized bucket is not enumerated itself; its haplotypes go, in
// bucket order, into a FRESH map keyed by the next word, whose buckets are visited recursively in that map's iteration
// order.  The rank is then the position of the haplotype's final (nested) bucket in this depth-first visit.
template <class WordFn>
std::vector<uint32_t> seedGroupRanks(const uint32_t numHaps, const int numWords, WordFn&& rawWord, const unsigned threads = 0,
                                     const int maxSeeds = 0, const int readAhead = 0)
{
  std::vector<uint32_t> rank(static_cast<size_t>(std::max(numWords, 0)) * numHaps);
  if (numWords <= 0) {
    return rank;
  }
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
  size_t buckets = 17;
  for (int w = 0; w < numWords; ++w) {
    seedBuckets[w] = buckets;
    buckets = NodeOrderMap::bucketsAfter(buckets, distinct[w]);
  }
  parallelForWords(numWords, threads, [&](const int w) {
    NodeOrderMap seeds(seedBuckets[w]);
    std::vector<int> nodeOfHap(numHaps);
    for (uint32_t h = 0; h < numHaps; ++h) {
      bool isNew;
      nodeOfHap[h] = seeds.insert(rawWord(h, w), 0, isNew);
    }
    uint32_t* out = rank.data() + static_cast<size_t>(w) * numHaps;
    uint32_t r = 0;
    if (maxSeeds > 0) {
      // members of every bucket in insertion (= haplotype) order, then the depth-first visit
      std::vector<std::vector<uint32_t>> members(numHaps);
      for (uint32_t h = 0; h < numHaps; ++h) {
        members[static_cast<size_t>(nodeOfHap[h])].push_back(h);
      }
      const int readWords = std::min(numWords, w + readAhead);
      struct Visit {
        static void run(const std::vector<uint32_t>& bucket, const int level, const int readWords, const int maxSeeds,
                        WordFn& rawWord, uint32_t* out, uint32_t& r)
        {
          if (bucket.size() > static_cast<size_t>(maxSeeds) && level + 1 < readWords) {
            NodeOrderMap sub;
            std::vector<std::vector<uint32_t>> subMembers;
            for (const uint32_t h : bucket) {
              bool isNew;
              const int nd = sub.insert(rawWord(h, level + 1), 0, isNew);
              if (static_cast<size_t>(nd) >= subMembers.size()) {
                subMembers.resize(static_cast<size_t>(nd) + 1);
              }
              subMembers[static_cast<size_t>(nd)].push_back(h);
            }
            for (int nd = sub.first(); nd != NodeOrderMap::kEnd; nd = sub.next(nd)) {
              run(subMembers[static_cast<size_t>(nd)], level + 1, readWords, maxSeeds, rawWord, out, r);
            }
            return;
          }
          for (const uint32_t h : bucket) {
            out[h] = r;
          }
          ++r;
        }
      };
      for (int nd = seeds.first(); nd != NodeOrderMap::kEnd; nd = seeds.next(nd)) {
        Visit::run(members[static_cast<size_t>(nd)], w, readWords, maxSeeds, rawWord, out, r);
      }
    } else {
      std::vector<uint32_t> rankOfNode(numHaps, 0);
      for (int nd = seeds.first(); nd != NodeOrderMap::kEnd; nd = seeds.next(nd)) {
        rankOfNode[static_cast<size_t>(nd)] = r++;
      }
      for (uint32_t h = 0; h < numHaps; ++h) {
        out[h] = rankOfNode[static_cast<size_t>(nodeOfHap[h])];
      }
    }
  });
  return rank;
}

}  // namespace candidate_order
