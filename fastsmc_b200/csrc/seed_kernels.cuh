// fastsmc_b200 — sm_100a kernels for GERMLINE-style candidate seeding (FastSMC's hashing step).
//
// Reference semantics (ASMC_SRC/SRC/FastSMC.cpp:144-229, HASHING/SeedHash.hpp, HASHING/ExtendHash.hpp): per 64-SNP
// word, haplotypes with identical words are grouped; every pair of a group "matches" at that word; a pair's matches
// are merged into intervals that tolerate `gap` non-matching words.  The reference keeps two hash maps alive across
// words and touches every matching pair at every word.  Here the interval structure is computed order-free:
//
//   transposeWordsKernel : [hap][word] packed haplotypes -> [word][hap] keys (coalesced for the per-word passes)
//   per word w (words are independent: a launch handles a batch of words, blockIdx.y = word within the batch, each
//   word with its own slice of the scratch tables, so that small sample counts are not bound by launch latency):
//     groupInsertKernel  : open-addressing hash on the 64-bit word itself (the reference's hash is the identity on
//                          the word, so equal key <=> identical word: no false positives); a slot is owned by the
//                          first haplotype that claims it; every haplotype gets (slot, rank within slot)
//     groupCompactKernel : slots with >= 2 members become groups; pair count n(n-1)/2 per group
//     groupScanKernel    : exclusive scans of group sizes and pair counts (one CTA; #groups <= H/2)
//     groupScatterKernel : members of each group, contiguous
//     pairExtendKernel   : persistent CTAs walk the word's pair space in chunks.  For each pair (a<b) matching at w:
//                          job filter on global ids, then the START test (no match in the gap+1 preceding words, read
//                          from the hap-major matrix).  Lanes that hold a start are served one at a time by the whole
//                          warp: 32 lanes compare 32 consecutive words of the two haplotypes per step (two coalesced
//                          256-byte reads) and a ballot finds where the run of > gap misses begins.  The interval is
//                          length-filtered in genetic distance and appended to the output.
//
// A pair's interval is discovered exactly once (at its start word), so the output is the set of the reference's
// Match objects at flush time; their order is fixed afterwards (canonical sort, or the host's replay of the
// reference's hash-map iteration order).
#pragma once

#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fastsmc_b200.h"

namespace fsmc
{

struct SeedArgs {
  const uint64_t* haps;      // [H][wordsPerHap] folded alleles
  long long wordsPerHap;
  const uint64_t* keysT;     // [W][H]
  uint32_t H;
  int W;
  int L;                     // sites
  int gap;
  float minLengthCm;
  const float* genPos;       // [L]
  const uint32_t* globalId;  // [H]
  uint32_t loI, hiI, loJ, hiJ;
  int lastJob, aboveDiag;
  unsigned flags;
  // per-word scratch; word j of a batch uses slice j of every table (strides C, H, maxGroups, maxGroups + 1)
  uint32_t* owner;           // [C] hap+1 owning the slot, 0 = empty
  uint32_t* slotCount;       // [C]
  uint32_t* slotGroup;       // [C]
  uint32_t C;                // power of two
  uint32_t* slotOf;          // [H]
  uint32_t* rankOf;          // [H]
  uint32_t* groupSize;       // [maxGroups]
  uint32_t* groupMemberBase; // [maxGroups]
  unsigned long long* groupPairBase;  // [maxGroups + 1]  (exclusive scan, last = total)
  uint32_t* members;         // [H]
  uint32_t maxGroups;
  unsigned long long* wordCounters;  // [word in batch][4]: [0] numGroups [1] pairCursor [2] totalPairs
  unsigned long long* counters;      // whole job: [3] matchCount [4] pairVisits [5] numStarts
  int wordBase;              // first word of the batch
  int wordsInBatch;
  unsigned long long* batchChunkBase;  // [wordsInBatch + 1] exclusive scan of the words' chunk counts, then [+1] cursor
  fsmc_match* out;
  long long capacity;
  const unsigned char* lowComplexity;  // [W] 1 = low-complexity word (DecodingParams::skip), nullptr = none
  // max_seeds > 0 (nestedKeysKernels below): haps / keysT hold REGISTRATION keys instead of the haplotype words, and a
  // registration at word x moves the interval's end to x + depth (bits 32.. of the key), at most maxDepth words ahead
  int nested;
  int maxDepth;
};

// The tables of word j of the batch.
__device__ __forceinline__ SeedArgs wordSlice(const SeedArgs& a, const uint32_t j)
{
  SeedArgs v = a;
  v.owner += static_cast<size_t>(j) * a.C;
  v.slotCount += static_cast<size_t>(j) * a.C;
  v.slotGroup += static_cast<size_t>(j) * a.C;
  v.slotOf += static_cast<size_t>(j) * a.H;
  v.rankOf += static_cast<size_t>(j) * a.H;
  v.members += static_cast<size_t>(j) * a.H;
  v.groupSize += static_cast<size_t>(j) * a.maxGroups;
  v.groupMemberBase += static_cast<size_t>(j) * a.maxGroups;
  v.groupPairBase += static_cast<size_t>(j) * (a.maxGroups + 1);
  v.wordCounters += static_cast<size_t>(j) * 4;
  return v;
}

__device__ __forceinline__ uint32_t mixKey(uint64_t k)
{
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return static_cast<uint32_t>(k);
}

// [H][wph] -> [W][H] through a 32x33 shared tile: reads coalesced along words, writes coalesced along haplotypes
__global__ void transposeWordsKernel(const uint64_t* __restrict__ haps, const long long wph, const uint32_t H,
                                     const int W, uint64_t* __restrict__ keysT)
{
  __shared__ uint64_t tile[32][33];
  const uint32_t h0 = blockIdx.x * 32u;
  const int w0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const uint32_t h = h0 + r;
    const int w = w0 + threadIdx.x;
    tile[r][threadIdx.x] = (h < H && w < W) ? haps[static_cast<size_t>(h) * wph + w] : 0ull;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int w = w0 + r;
    const uint32_t h = h0 + threadIdx.x;
    if (h < H && w < W) {
      keysT[static_cast<size_t>(w) * H + h] = tile[threadIdx.x][r];
    }
  }
}

__global__ void groupInsertKernel(const SeedArgs args)
{
  const SeedArgs a = wordSlice(args, blockIdx.y);
  const int w = args.wordBase + static_cast<int>(blockIdx.y);
  const uint64_t* keys = a.keysT + static_cast<size_t>(w) * a.H;
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < a.H; h += gridDim.x * blockDim.x) {
    const uint64_t k = keys[h];
    uint32_t slot = mixKey(k) & (a.C - 1);
    for (;;) {
      uint32_t o = a.owner[slot];
      if (o == 0) {
        o = atomicCAS(&a.owner[slot], 0u, h + 1u);
        if (o == 0) {
          o = h + 1u;
        }
      }
      if (o == h + 1u || keys[o - 1u] == k) {
        break;
      }
      slot = (slot + 1u) & (a.C - 1);
    }
    a.slotOf[h] = slot;
    a.rankOf[h] = atomicAdd(&a.slotCount[slot], 1u);
  }
}

__global__ void groupCompactKernel(const SeedArgs args)
{
  const SeedArgs a = wordSlice(args, blockIdx.y);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < a.C; s += gridDim.x * blockDim.x) {
    const uint32_t n = a.slotCount[s];
    if (n >= 2u) {
      const uint32_t g = static_cast<uint32_t>(atomicAdd(&a.wordCounters[0], 1ull));
      a.slotGroup[s] = g;
      a.groupSize[g] = n;
    }
  }
}

// One CTA: exclusive scans over the groups (member offsets, pair offsets).  Each thread owns a contiguous run.
__global__ void groupScanKernel(const SeedArgs args)
{
  const SeedArgs a = wordSlice(args, blockIdx.x);
  __shared__ unsigned long long partMembers[1024], partPairs[1024];
  const uint32_t G = static_cast<uint32_t>(a.wordCounters[0]);
  const uint32_t T = blockDim.x, t = threadIdx.x;
  const uint32_t per = (G + T - 1) / T;
  const uint32_t lo = min(G, t * per), hi = min(G, lo + per);
  unsigned long long m = 0, p = 0;
  for (uint32_t g = lo; g < hi; ++g) {
    const unsigned long long n = a.groupSize[g];
    m += n;
    p += n * (n - 1ull) / 2ull;
  }
  partMembers[t] = m;
  partPairs[t] = p;
  __syncthreads();
  if (t == 0) {
    unsigned long long rm = 0, rp = 0;
    for (uint32_t i = 0; i < T; ++i) {
      const unsigned long long xm = partMembers[i], xp = partPairs[i];
      partMembers[i] = rm;
      partPairs[i] = rp;
      rm += xm;
      rp += xp;
    }
    a.groupPairBase[G] = rp;
    a.wordCounters[2] = rp;
    a.wordCounters[1] = 0ull;
    atomicAdd(&a.counters[4], rp);
  }
  __syncthreads();
  m = partMembers[t];
  p = partPairs[t];
  for (uint32_t g = lo; g < hi; ++g) {
    const unsigned long long n = a.groupSize[g];
    a.groupMemberBase[g] = static_cast<uint32_t>(m);
    a.groupPairBase[g] = p;
    m += n;
    p += n * (n - 1ull) / 2ull;
  }
}

__global__ void groupScatterKernel(const SeedArgs args)
{
  const SeedArgs a = wordSlice(args, blockIdx.y);
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < a.H; h += gridDim.x * blockDim.x) {
    const uint32_t s = a.slotOf[h];
    if (a.slotCount[s] >= 2u) {
      a.members[a.groupMemberBase[a.slotGroup[s]] + a.rankOf[h]] = h;
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// max_seeds (ref: HASHING/SeedHash.hpp:56-69, 85-93): a bucket of word c with more than max_seeds haplotypes is not
// enumerated; its members are re-hashed on word c+1 (recursively, while the next word is inside the read-ahead buffer:
// c+j+1 < min(W, c + readAhead)), and the pairs of the final, nested bucket are extended to word c+j, not c.  Order-free
// form: every (word c, haplotype h) gets a REGISTRATION KEY = (depth j, representative haplotype of the nested bucket at
// that depth).  Two haplotypes are paired by the reference at word c iff their registration keys at c are equal, and
// the pair's interval end moves to c + j.  With these keys in place of the words, grouping, start test and extension
// are the kernels of the plain case.
//
// Level 0 groups word c by the word itself (groupInsertKernel) and records each haplotype's representative rep0[c][h]
// (the slot owner: unique per distinct word).  Level j >= 1 groups the haplotypes of word c that are not final yet by the
// exact 64-bit key (representative at level j-1, rep0[c+j][h]).
// -------------------------------------------------------------------------------------------------------------------
struct NestArgs {
  uint32_t H;
  int W;
  uint32_t C;
  int maxSeeds, readAhead;
  int level;
  int wordBase;             // first word of the batch; blockIdx.y = word within the batch
  uint32_t* owner;          // [batch][C]
  uint32_t* slotCount;      // [batch][C]
  uint32_t* slotOf;         // [batch][H]
  uint32_t* rep0;           // [W][H]
  uint32_t* cur;            // [W][H] representative at the current depth
  unsigned char* depth;     // [W][H] bit 7 = final
  unsigned long long* pending;  // number of (word, haplotype) entries that go one level deeper
  uint64_t* regKeyT;        // [W][H]
};

__device__ __forceinline__ bool nestGoesDeeper(const NestArgs& a, const uint32_t n, const int w, const int level)
{
  const int readWords = min(a.W, w + a.readAhead);  // ref: FastSMC.cpp:188-199 (GLOBAL_READ_WORDS while word w is current)
  return n > static_cast<uint32_t>(a.maxSeeds) && w + level + 1 < readWords;
}

// after groupInsertKernel on the raw words of the batch
__global__ void nestLevel0Kernel(const NestArgs a)
{
  const int w = a.wordBase + static_cast<int>(blockIdx.y);
  const size_t j = blockIdx.y;
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < a.H; h += gridDim.x * blockDim.x) {
    const uint32_t s = a.slotOf[j * a.H + h];
    const uint32_t rep = a.owner[j * a.C + s] - 1u;
    const bool deeper = nestGoesDeeper(a, a.slotCount[j * a.C + s], w, 0);
    const size_t o = static_cast<size_t>(w) * a.H + h;
    a.rep0[o] = rep;
    a.cur[o] = rep;
    a.depth[o] = deeper ? 0 : 0x80;
    if (deeper) {
      atomicAdd(a.pending, 1ull);
    }
  }
}

__device__ __forceinline__ uint64_t nestKey(const NestArgs& a, const int w, const uint32_t h)
{
  return (static_cast<uint64_t>(a.cur[static_cast<size_t>(w) * a.H + h]) << 32) |
         a.rep0[static_cast<size_t>(w + a.level) * a.H + h];
}

__global__ void nestInsertKernel(const NestArgs a)
{
  const int w = a.wordBase + static_cast<int>(blockIdx.y);
  const size_t j = blockIdx.y;
  uint32_t* owner = a.owner + j * a.C;
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < a.H; h += gridDim.x * blockDim.x) {
    if (a.depth[static_cast<size_t>(w) * a.H + h] & 0x80) {
      continue;
    }
    const uint64_t k = nestKey(a, w, h);
    uint32_t slot = mixKey(k) & (a.C - 1);
    for (;;) {
      uint32_t o = owner[slot];
      if (o == 0) {
        o = atomicCAS(&owner[slot], 0u, h + 1u);
        if (o == 0) {
          o = h + 1u;
        }
      }
      if (o == h + 1u || nestKey(a, w, o - 1u) == k) {
        break;
      }
      slot = (slot + 1u) & (a.C - 1);
    }
    a.slotOf[j * a.H + h] = slot;
    atomicAdd(&a.slotCount[j * a.C + slot], 1u);
  }
}

// reads owner / slotCount only (no haplotype's `cur` is read here, so they can be updated in place)
__global__ void nestFinishKernel(const NestArgs a)
{
  const int w = a.wordBase + static_cast<int>(blockIdx.y);
  const size_t j = blockIdx.y;
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < a.H; h += gridDim.x * blockDim.x) {
    const size_t o = static_cast<size_t>(w) * a.H + h;
    if (a.depth[o] & 0x80) {
      continue;
    }
    const uint32_t s = a.slotOf[j * a.H + h];
    const bool deeper = nestGoesDeeper(a, a.slotCount[j * a.C + s], w, a.level);
    a.cur[o] = a.owner[j * a.C + s] - 1u;
    a.depth[o] = static_cast<unsigned char>(a.level | (deeper ? 0 : 0x80));
    if (deeper) {
      atomicAdd(a.pending, 1ull);
    }
  }
}

__global__ void nestEmitKernel(const NestArgs a)
{
  const size_t total = static_cast<size_t>(a.W) * a.H;
  for (size_t o = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
    a.regKeyT[o] = (static_cast<uint64_t>(a.depth[o] & 0x7f) << 32) | a.cur[o];
  }
}

// [W][H] -> [H][W] (the walks of pairExtendKernel read one haplotype's consecutive words)
__global__ void transposeKeysBackKernel(const uint64_t* __restrict__ keysT, const uint32_t H, const int W, uint64_t* __restrict__ keys)
{
  __shared__ uint64_t tile[32][33];
  const uint32_t h0 = blockIdx.x * 32u;
  const int w0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int w = w0 + r;
    const uint32_t h = h0 + threadIdx.x;
    tile[r][threadIdx.x] = (h < H && w < W) ? keysT[static_cast<size_t>(w) * H + h] : 0ull;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const uint32_t h = h0 + r;
    const int w = w0 + threadIdx.x;
    if (h < H && w < W) {
      keys[static_cast<size_t>(h) * W + w] = tile[threadIdx.x][r];
    }
  }
}

constexpr int kPairChunk = 8;        // consecutive pair indices per thread
constexpr int kLaneWords = 8;        // words a lane extends its own interval before the warp takes over
constexpr int kPairBlockThreads = 256;
constexpr unsigned long long kPairsPerChunk = static_cast<unsigned long long>(kPairChunk) * kPairBlockThreads;

// One warp: the pair space of every word of the batch is cut into chunks of kPairsPerChunk pairs; the chunks of all
// words form one queue (exclusive scan of the chunk counts), so that pairExtendKernel's CTAs balance across words.
__global__ void batchChunksKernel(const SeedArgs a)
{
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int j = 0; j < a.wordsInBatch; ++j) {
      a.batchChunkBase[j] = run;
      if (a.lowComplexity && a.lowComplexity[a.wordBase + j]) {
        continue;  // no pair is seeded at a low-complexity word (ref: FastSMC.cpp:212-219)
      }
      run += (a.wordCounters[static_cast<size_t>(j) * 4 + 2] + kPairsPerChunk - 1) / kPairsPerChunk;
    }
    a.batchChunkBase[a.wordsInBatch] = run;
    a.batchChunkBase[a.wordsInBatch + 1] = 0ull;  // queue head
  }
}

// job filter on global haplotype ids, gi > gj (ref: HASHING/SeedHash.hpp:99-129)
__device__ __forceinline__ bool pairInJob(const SeedArgs& a, const uint32_t hi, const uint32_t lo)
{
  const uint32_t gi = a.globalId[hi], gj = a.globalId[lo];
  if (a.lastJob) {
    return gi >= a.loI && gj >= a.loJ && gj < a.loJ + (gi - a.loI);
  }
  if (gi >= a.loI && gi < a.hiI && gj >= a.loJ && gj < a.hiJ) {
    return a.aboveDiag ? (gj < a.loJ + (gi - a.loI)) : (gj >= a.loJ + (gi - a.loI));
  }
  return false;
}

// NESTED (max_seeds > 0): A/B rows hold registration keys; a registration at word x proposes the end x + depth(key).
// The reference's bookkeeping per pair (ExtendHash.hpp:61-106): a registration sets end = max(end, x + depth); a
// low-complexity word sets end = x for whatever is alive; after a seeded word c an interval with end < c - gap is flushed.
__device__ __forceinline__ int keyDepth(const uint64_t k)
{
  return static_cast<int>(k >> 32);
}

template <bool NESTED> __global__ void __launch_bounds__(kPairBlockThreads) pairExtendKernel(const SeedArgs args)
{
  const unsigned lane = threadIdx.x & 31u;
  __shared__ unsigned long long blockChunk;
  __shared__ unsigned long long chunkBase[65];
  for (int j = threadIdx.x; j <= args.wordsInBatch; j += blockDim.x) {
    chunkBase[j] = args.batchChunkBase[j];
  }
  __syncthreads();
  const unsigned long long totalChunks = chunkBase[args.wordsInBatch];
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      blockChunk = atomicAdd(&args.batchChunkBase[args.wordsInBatch + 1], 1ull);
    }
    __syncthreads();
    const unsigned long long chunk = blockChunk;
    if (chunk >= totalChunks) {
      break;
    }
    // word of this chunk: last j with chunkBase[j] <= chunk (at most 64 words per batch)
    int j = 0;
    for (int step = 32; step > 0; step >>= 1) {
      if (j + step < args.wordsInBatch && chunkBase[j + step] <= chunk) {
        j += step;
      }
    }
    while (j + 1 < args.wordsInBatch && chunkBase[j + 1] <= chunk) {
      ++j;
    }
    const SeedArgs a = wordSlice(args, static_cast<uint32_t>(j));
    const int w = args.wordBase + j;
    const unsigned long long total = a.wordCounters[2];
    const uint32_t G = static_cast<uint32_t>(a.wordCounters[0]);
    const unsigned long long base = (chunk - chunkBase[j]) * kPairsPerChunk;
    unsigned long long q = base + static_cast<unsigned long long>(threadIdx.x) * kPairChunk;
    // locate the group of pair index q: last g with groupPairBase[g] <= q
    uint32_t g = 0, n = 0, memberBase = 0, i = 0, ii = 0;
    unsigned long long left = 0;  // pairs of this thread's run still to visit
    if (q < total) {
      uint32_t lo = 0, hi = G;
      while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.groupPairBase[mid] <= q) {
          lo = mid;
        } else {
          hi = mid;
        }
      }
      g = lo;
      const unsigned long long r = q - a.groupPairBase[g];
      // r = i(i-1)/2 + ii with ii < i
      i = static_cast<uint32_t>((1.0 + sqrt(1.0 + 8.0 * static_cast<double>(r))) * 0.5);
      while (static_cast<unsigned long long>(i) * (i - 1ull) / 2ull > r) {
        --i;
      }
      while (static_cast<unsigned long long>(i + 1ull) * i / 2ull <= r) {
        ++i;
      }
      ii = static_cast<uint32_t>(r - static_cast<unsigned long long>(i) * (i - 1ull) / 2ull);
      n = a.groupSize[g];
      memberBase = a.groupMemberBase[g];
      left = min(static_cast<unsigned long long>(kPairChunk), total - q);
    }
    for (int step = 0; step < kPairChunk; ++step) {
      bool isStart = false;
      uint32_t hLo = 0, hHi = 0;
      if (left > 0) {
        const uint32_t x = a.members[memberBase + i], y = a.members[memberBase + ii];
        hLo = min(x, y);
        hHi = max(x, y);
        if (pairInJob(a, hHi, hLo)) {
          isStart = true;
          const uint64_t* A = a.haps + static_cast<size_t>(hLo) * a.wordsPerHap;
          const uint64_t* B = a.haps + static_cast<size_t>(hHi) * a.wordsPerHap;
          if constexpr (!NESTED) {
            // alive before w?  Walk back: gap+1 misses in a row mean the earlier interval was flushed; a low-complexity
            // word extends whatever is alive without being a match itself, so it restarts the count of misses
            // (ref: ExtendHash.hpp:85-106)
            int missesBack = 0;
            for (int x = w - 1; x >= 0 && missesBack <= a.gap; --x) {
              if (a.lowComplexity && a.lowComplexity[x]) {
                missesBack = 0;
              } else if (__ldg(A + x) == __ldg(B + x)) {
                isStart = false;
                break;
              } else {
                ++missesBack;
              }
            }
          } else {
            // The same question when ends run ahead of the registrations.  Walking back, `need` is the smallest end an
            // interval must have after word x to be alive before w (kAny: being alive is enough): a seeded word x without
            // a sufficient registration raises it to x - gap; a low-complexity word resets every end to x, so it either
            // satisfies the need (alive before it is then enough) or rules the past out.  Nothing before x can help
            // once need > x - 1 + maxDepth.
            constexpr int kAny = INT_MIN;
            int need = kAny;
            for (int x = w - 1; x >= 0; --x) {
              if (a.lowComplexity && a.lowComplexity[x]) {
                if (x >= need) {
                  need = kAny;
                  continue;
                }
                break;
              }
              const uint64_t ka = __ldg(A + x);
              if (ka == __ldg(B + x) && x + keyDepth(ka) >= need) {
                isStart = false;
                break;
              }
              need = max(need, x - a.gap);
              if (need > x - 1 + a.maxDepth) {
                break;
              }
            }
          }
        }
        // advance to the next pair of the run
        --left;
        if (++ii == i) {
          ii = 0;
          if (++i == n && left > 0) {
            ++g;
            n = a.groupSize[g];
            memberBase = a.groupMemberBase[g];
            i = 1;
          }
        }
      }
      // ---- extension.  Most intervals are one or two words long, so every lane first walks its own interval for up
      // to kLaneWords words (32 independent load streams per warp instead of one interval at a time); the few that
      // are still open after that are finished by the whole warp, 32 words per step.
      int end = w, misses = 0, pos = w + 1;
      bool open = isStart;
      if (isStart) {
        const uint64_t* A = a.haps + static_cast<size_t>(hLo) * a.wordsPerHap;
        const uint64_t* B = a.haps + static_cast<size_t>(hHi) * a.wordsPerHap;
        if constexpr (NESTED) {
          end = w + keyDepth(__ldg(A + w));
        }
        for (int k = 0; k < kLaneWords && open && pos < a.W; ++k, ++pos) {
          if constexpr (!NESTED) {
            if ((a.lowComplexity && a.lowComplexity[pos]) || __ldg(A + pos) == __ldg(B + pos)) {
              end = pos;
              misses = 0;
            } else if (++misses > a.gap) {
              open = false;
            }
          } else {
            if (a.lowComplexity && a.lowComplexity[pos]) {
              end = pos;
            } else {
              const uint64_t ka = __ldg(A + pos);
              if (ka == __ldg(B + pos)) {
                end = max(end, pos + keyDepth(ka));
              }
              if (end < pos - a.gap) {
                open = false;
              }
            }
          }
        }
        if (pos >= a.W) {
          open = false;  // ran into the end of the chromosome
        }
      }
      unsigned pending = __ballot_sync(0xffffffffu, isStart && open);
      while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1u;
        const uint32_t sLo = __shfl_sync(0xffffffffu, hLo, src), sHi = __shfl_sync(0xffffffffu, hHi, src);
        int sEnd = __shfl_sync(0xffffffffu, end, src), sMisses = __shfl_sync(0xffffffffu, misses, src);
        const int sPos = __shfl_sync(0xffffffffu, pos, src);
        const uint64_t* A = a.haps + static_cast<size_t>(sLo) * a.wordsPerHap;
        const uint64_t* B = a.haps + static_cast<size_t>(sHi) * a.wordsPerHap;
        bool sOpen = true;
        for (int p0 = sPos; p0 < a.W && sOpen; p0 += 32) {
          const int x = p0 + static_cast<int>(lane);
          const int valid = min(32, a.W - p0);
          if constexpr (!NESTED) {
            const bool eq = x < a.W && ((a.lowComplexity && a.lowComplexity[x]) || __ldg(A + x) == __ldg(B + x));
            const unsigned m = __ballot_sync(0xffffffffu, eq);
            for (int b = 0; b < valid; ++b) {  // warp-uniform walk over the 32 comparison bits
              if ((m >> b) & 1u) {
                sEnd = p0 + b;
                sMisses = 0;
              } else if (++sMisses > a.gap) {
                sOpen = false;
                break;
              }
            }
          } else {
            // per word: -1 = nothing, -2 = low-complexity word, else the end a registration proposes
            int ev = -1;
            if (x < a.W) {
              if (a.lowComplexity && a.lowComplexity[x]) {
                ev = -2;
              } else {
                const uint64_t ka = __ldg(A + x);
                if (ka == __ldg(B + x)) {
                  ev = x + keyDepth(ka);
                }
              }
            }
            for (int b = 0; b < valid; ++b) {
              const int e = __shfl_sync(0xffffffffu, ev, b);
              if (e == -2) {
                sEnd = p0 + b;
                continue;
              }
              sEnd = max(sEnd, e);
              if (sEnd < p0 + b - a.gap) {
                sOpen = false;
                break;
              }
            }
          }
        }
        if (static_cast<int>(lane) == src) {
          end = sEnd;
        }
      }
      // ---- length filter and output, one atomic per warp (ref: HASHING/Utils.cpp:22-34, HASHING/Match.hpp:46-51)
      bool keep = false;
      if (isStart) {
        const int sEnd = min(64 * end + 63, a.L - 1);
        const float d = a.genPos[sEnd] - a.genPos[64 * w];
        keep = (a.flags & FSMC_SEED_ALL_INTERVALS) || (100.0 * static_cast<double>(d) >= static_cast<double>(a.minLengthCm));
      }
      const unsigned startMask = __ballot_sync(0xffffffffu, isStart), keepMask = __ballot_sync(0xffffffffu, keep);
      if (startMask) {
        unsigned long long base = 0;
        if (lane == 0) {
          atomicAdd(&a.counters[5], static_cast<unsigned long long>(__popc(startMask)));
          if (keepMask) {
            base = atomicAdd(&a.counters[3], static_cast<unsigned long long>(__popc(keepMask)));
          }
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
          const unsigned long long idx = base + __popc(keepMask & ((1u << lane) - 1u));
          if (static_cast<long long>(idx) < a.capacity) {
            a.out[idx] = fsmc_match{hLo, hHi, w, end};
          }
        }
      }
    }
  }
}

}  // namespace fsmc
