// fastsmc_b200 — see seed_order.h.
#include "seed_order.h"

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

namespace fsmc
{

namespace
{

constexpr uint32_t kLongFlag = 0x80000000u;  // in Node::end: the interval passes the length filter (is a candidate)

struct Node {     // per interval, in creation order
  int32_t start;  // start word
  uint32_t end;   // end word | kLongFlag
};

__device__ __forceinline__ long long gridStart()
{
  return blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
}
__device__ __forceinline__ long long gridStep()
{
  return static_cast<long long>(gridDim.x) * blockDim.x;
}

// mode 0: one mixed-radix key ((start * H + rank) * H + a) * H + b; mode 1: a * H + b; mode 2: start * H + rank
__global__ void creationKeyKernel(const fsmc_match* __restrict__ iv, const uint32_t* __restrict__ index, const long long n,
                                  const uint64_t H, const uint32_t* __restrict__ rank, const int mode,
                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
  for (long long i = gridStart(); i < n; i += gridStep()) {
    const uint32_t src = index ? index[i] : static_cast<uint32_t>(i);
    const fsmc_match m = iv[src];
    const uint64_t r = rank[static_cast<size_t>(m.startWord) * H + m.hapA];
    const uint64_t pair = static_cast<uint64_t>(m.hapA) * H + m.hapB;
    const uint64_t head = static_cast<uint64_t>(m.startWord) * H + r;
    keys[i] = mode == 0 ? head * H * H + pair : (mode == 1 ? pair : head);
    vals[i] = src;
  }
}

// creation order q -> pair key, (start, end | long flag); boundaries of the start words; histogram of end words
__global__ void gatherNodesKernel(const fsmc_match* __restrict__ iv, const uint32_t* __restrict__ sortedIndex, const long long n,
                                  const uint64_t H, const int L, const float* __restrict__ genPos, const float minLengthCm,
                                  uint64_t* __restrict__ pairKey, Node* __restrict__ node, long long* __restrict__ startBegin,
                                  long long* __restrict__ startEnd, unsigned long long* __restrict__ endCount,
                                  unsigned long long* __restrict__ numLong)
{
  unsigned long long mine = 0;
  for (long long q0 = blockIdx.x * static_cast<long long>(blockDim.x); q0 < n; q0 += gridStep()) {
    const long long q = q0 + threadIdx.x;
    const bool live = q < n;
    fsmc_match m{0, 0, -1, -1};
    if (live) {
      m = iv[sortedIndex[q]];
      // ref: HASHING/Utils.cpp:22-34, HASHING/Match.hpp:46-51 — float difference, double product
      const float d = genPos[min(64 * m.endWord + 63, L - 1)] - genPos[64 * m.startWord];
      const bool isLong = 100.0 * static_cast<double>(d) >= static_cast<double>(minLengthCm);
      mine += isLong;
      pairKey[q] = static_cast<uint64_t>(m.hapA) * H + m.hapB;
      node[q] = Node{m.startWord, static_cast<uint32_t>(m.endWord) | (isLong ? kLongFlag : 0u)};
      const int prev = q > 0 ? iv[sortedIndex[q - 1]].startWord : -1;
      const int next = q + 1 < n ? iv[sortedIndex[q + 1]].startWord : -1;
      if (prev != m.startWord) {
        startBegin[m.startWord] = q;
      }
      if (next != m.startWord) {
        startEnd[m.startWord] = q + 1;
      }
    }
    // one atomic per distinct end word of the warp
    const unsigned active = __ballot_sync(0xffffffffu, live);
    if (live) {
      const unsigned same = __match_any_sync(active, m.endWord);
      if ((__ffs(same) - 1) == static_cast<int>(threadIdx.x & 31)) {
        atomicAdd(&endCount[m.endWord], static_cast<unsigned long long>(__popc(same)));
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mine += __shfl_down_sync(0xffffffffu, mine, o);
  }
  if ((threadIdx.x & 31) == 0 && mine) {
    atomicAdd(numLong, mine);
  }
}

struct LiveAt {  // nodes created before the rehash that are still in the map: end word >= minEnd
  const Node* node;
  int minEnd;
  __device__ bool operator()(const uint32_t q) const { return static_cast<int>(node[q].end & ~kLongFlag) >= minEnd; }
};

// list order = descending (g, w): ascending sort of the complemented key
__global__ void walkKeyKernel(const uint32_t* __restrict__ ids, const long long N, const uint32_t* __restrict__ g,
                              const uint32_t* __restrict__ w, uint64_t* __restrict__ keys)
{
  for (long long i = gridStart(); i < N; i += gridStep()) {
    const uint32_t q = ids[i];
    keys[i] = ~((static_cast<uint64_t>(g[q]) << 32) | w[q]);
  }
}

__global__ void rehashFirstKernel(const uint32_t* __restrict__ walk, const long long N, const uint64_t* __restrict__ pairKey,
                                  const uint64_t B, uint32_t* __restrict__ bucketFirst)
{
  for (long long i = gridStart(); i < N; i += gridStep()) {
    atomicMin(&bucketFirst[pairKey[walk[i]] % B], static_cast<uint32_t>(i));
  }
}

// groups re-form in the order their first node is met; inside a group the nodes end up in reverse walk order
__global__ void rehashAssignKernel(const uint32_t* __restrict__ walk, const long long N, const uint64_t* __restrict__ pairKey,
                                   const uint64_t B, const uint32_t* __restrict__ bucketFirst, uint32_t* __restrict__ g,
                                   uint32_t* __restrict__ w)
{
  for (long long i = gridStart(); i < N; i += gridStep()) {
    const uint32_t q = walk[i];
    g[q] = static_cast<uint32_t>(N) - bucketFirst[pairKey[q] % B];
    w[q] = static_cast<uint32_t>(i);
  }
}

// (bucket, isNew, creation rank): carried nodes of a bucket first, then the epoch's new nodes in creation order
__global__ void epochKeyKernel(const uint32_t* __restrict__ carried, const long long N, const long long q0, const long long M,
                               const uint64_t* __restrict__ pairKey, const uint64_t B, uint64_t* __restrict__ keys)
{
  for (long long i = gridStart(); i < N + M; i += gridStep()) {
    const bool isNew = i >= N;
    const uint32_t q = isNew ? static_cast<uint32_t>(q0 + (i - N)) : carried[i];
    keys[i] = ((pairKey[q] % B) << 32) | (isNew ? 0x80000000ull : 0ull) | q;
  }
}

struct CandidateList {
  uint64_t* key;    // ~(g << 32 | w)
  uint32_t* phase;  // word after whose insertions the node leaves the map; numWords = the final flush
  uint32_t* q;
  unsigned long long* count;
};

// One thread per bucket: the bucket's nodes in time order.  `until` = last word during whose insertions the bucket still
// holds one of the nodes seen so far (a node leaves after word end + gap + 1, ref: HASHING/ExtendHash.hpp:85-98).
__global__ void bucketWalkKernel(const uint64_t* __restrict__ keys, const long long total, const long long q0,
                                 const uint32_t base, const int epoch, const int gap, const int numWords,
                                 const Node* __restrict__ node, uint32_t* __restrict__ g, uint32_t* __restrict__ w,
                                 const int* __restrict__ phaseEpoch, const CandidateList cand)
{
  for (long long i = gridStart(); i < total; i += gridStep()) {
    const uint64_t bucket = keys[i] >> 32;
    if (i > 0 && (keys[i - 1] >> 32) == bucket) {
      continue;  // not the first node of its bucket
    }
    uint32_t curG = 0;
    int until = -1;
    for (long long j = i; j < total; ++j) {
      const uint64_t k = keys[j];
      if ((k >> 32) != bucket) {
        break;
      }
      const uint32_t q = static_cast<uint32_t>(k) & 0x7fffffffu;
      const Node nd = node[q];
      const int flush = min(static_cast<int>(nd.end & ~kLongFlag) + gap + 1, numWords);
      uint32_t myG, myW;
      if (!(k & 0x80000000ull)) {  // carried over the rehash: the bucket's carried nodes share one group
        myG = g[q];
        myW = w[q];
        curG = myG;
        until = max(until, flush);
      } else {
        myW = base + static_cast<uint32_t>(q - q0);  // tick of the insertion
        if (until >= nd.start) {
          myG = curG;  // non-empty bucket: front of its group
          until = max(until, flush);
        } else {
          myG = myW;   // empty bucket: a new group at the front of the list
          curG = myW;
          until = flush;
        }
        g[q] = myG;
        w[q] = myW;
      }
      if ((nd.end & kLongFlag) && phaseEpoch[flush] == epoch) {
        const unsigned long long at = atomicAdd(cand.count, 1ull);
        cand.key[at] = ~((static_cast<uint64_t>(myG) << 32) | myW);
        cand.phase[at] = static_cast<uint32_t>(flush);
        cand.q[at] = q;
      }
    }
  }
}

__global__ void iotaKernel(uint32_t* __restrict__ v, const long long n)
{
  for (long long i = gridStart(); i < n; i += gridStep()) {
    v[i] = static_cast<uint32_t>(i);
  }
}
__global__ void gatherU32Kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ index, const long long n,
                                uint32_t* __restrict__ dst)
{
  for (long long i = gridStart(); i < n; i += gridStep()) {
    dst[i] = src[index[i]];
  }
}
__global__ void emitKernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ candQ, const long long n,
                           const uint64_t* __restrict__ pairKey, const Node* __restrict__ node, const uint64_t H,
                           fsmc_match* __restrict__ out)
{
  for (long long i = gridStart(); i < n; i += gridStep()) {
    const uint32_t q = candQ[order[i]];
    const uint64_t pk = pairKey[q];
    out[i] = fsmc_match{static_cast<uint32_t>(pk / H), static_cast<uint32_t>(pk % H), node[q].start,
                        static_cast<int32_t>(node[q].end & ~kLongFlag)};
  }
}

int bitsFor(unsigned long long maxValue)
{
  int bits = 1;
  while (bits < 64 && (maxValue >> bits) != 0ull) {
    ++bits;
  }
  return bits;
}

// boost's prime bucket counts (boost/unordered/detail/implementation.hpp, prime_list) and growth rule
unsigned long long primeAtLeast(const unsigned long long n)
{
  static const unsigned long long primes[] = {
      17ull,       29ull,       37ull,        53ull,        67ull,        79ull,        97ull,         131ull,        193ull,       257ull,
      389ull,      521ull,      769ull,       1031ull,      1543ull,      2053ull,      3079ull,       6151ull,       12289ull,     24593ull,
      49157ull,    98317ull,    196613ull,    393241ull,    786433ull,    1572869ull,   3145739ull,    6291469ull,    12582917ull,  25165843ull,
      50331653ull, 100663319ull, 201326611ull, 402653189ull, 805306457ull, 1610612741ull, 3221225473ull, 4294967291ull};
  for (const unsigned long long p : primes) {
    if (p >= n) {
      return p;
    }
  }
  return 4294967291ull;
}
unsigned long long growTo(const unsigned long long count)
{
  return primeAtLeast(std::max(count + 1, count + (count >> 1)) + 1);
}

#define ORDER_CUDA(expr)       \
  do {                         \
    const cudaError_t e_ = (expr); \
    if (e_ != cudaSuccess) {   \
      return e_;               \
    }                          \
  } while (0)

}  // namespace

CandidateOrderer::~CandidateOrderer()
{
  release();
}

void CandidateOrderer::release()
{
  for (Buf* b : {&mKeysA, &mKeysB, &mValsA, &mValsB, &mPairKey, &mSe, &mG, &mW, &mBucketFirst, &mCarried, &mCandKey, &mCandKeyB,
                 &mCandQ, &mCandQB, &mCandPhase, &mCandPhaseB, &mOut, &mTemp, &mCounts, &mPhaseEpoch}) {
    cudaFree(b->p);
    b->p = nullptr;
    b->bytes = 0;
  }
}

cudaError_t CandidateOrderer::ensure(Buf& b, const size_t bytes)
{
  if (bytes <= b.bytes && b.p) {
    return cudaSuccess;
  }
  cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  const size_t want = std::max<size_t>(bytes + bytes / 8, 256);
  const cudaError_t e = cudaMalloc(&b.p, want);
  if (e == cudaSuccess) {
    b.bytes = want;
  }
  return e;
}

cudaError_t CandidateOrderer::order(const fsmc_match* intervals, const long long n, const uint32_t numHaps, const int numWords,
                                    const int sites, const int gap, const float minLengthCm, const float* genPos,
                                    const uint32_t* rank, cudaStream_t stream, const fsmc_match** out, long long* count,
                                    OrderStats* stats)
{
  *out = nullptr;
  *count = 0;
  if (stats) {
    *stats = OrderStats{};
    stats->intervals = n;
  }
  if (n <= 0 || numWords <= 0) {
    return cudaSuccess;
  }
  if (n >= 0x7fffffffll) {
    return cudaErrorInvalidValue;  // creation ranks are carried in 31 bits
  }
  const uint64_t H = numHaps;
  const int W = numWords;
  const size_t N = static_cast<size_t>(n);
  const int threads = 256;
  auto blocksFor = [&](const long long items) { return static_cast<int>(std::max<long long>(1, std::min<long long>((items + threads - 1) / threads, 148 * 16))); };
  auto temp = [&](const size_t bytes) { return ensure(mTemp, bytes); };

  ORDER_CUDA(ensure(mKeysA, N * 8));
  ORDER_CUDA(ensure(mKeysB, N * 8));
  ORDER_CUDA(ensure(mValsA, N * 4));
  ORDER_CUDA(ensure(mValsB, N * 4));
  ORDER_CUDA(ensure(mPairKey, N * 8));
  ORDER_CUDA(ensure(mSe, N * sizeof(Node)));
  ORDER_CUDA(ensure(mG, N * 4));
  ORDER_CUDA(ensure(mW, N * 4));
  // counters: [0, W) startBegin, [W, 2W) startEnd (long long); then [2W, 3W) endCount, [3W] numLong, [3W+1] candidates, [3W+2] selected
  const size_t nCounters = 3 * static_cast<size_t>(W) + 4;
  ORDER_CUDA(ensure(mCounts, nCounters * 8));
  ORDER_CUDA(ensure(mPhaseEpoch, (static_cast<size_t>(W) + 1) * sizeof(int)));
  uint64_t* keysA = static_cast<uint64_t*>(mKeysA.p);
  uint64_t* keysB = static_cast<uint64_t*>(mKeysB.p);
  uint32_t* valsA = static_cast<uint32_t*>(mValsA.p);
  uint32_t* valsB = static_cast<uint32_t*>(mValsB.p);
  uint64_t* pairKey = static_cast<uint64_t*>(mPairKey.p);
  Node* node = static_cast<Node*>(mSe.p);
  uint32_t* g = static_cast<uint32_t*>(mG.p);
  uint32_t* w = static_cast<uint32_t*>(mW.p);
  long long* startBegin = static_cast<long long*>(mCounts.p);
  long long* startEnd = startBegin + W;
  unsigned long long* endCount = reinterpret_cast<unsigned long long*>(startEnd + W);
  unsigned long long* numLong = endCount + W;
  unsigned long long* candCount = numLong + 1;
  int* numSelected = reinterpret_cast<int*>(candCount + 1);

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  ORDER_CUDA(cudaEventCreate(&ev0));
  ORDER_CUDA(cudaEventCreate(&ev1));
  struct EventGuard {
    cudaEvent_t a, b;
    ~EventGuard()
    {
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  } guard{ev0, ev1};
  ORDER_CUDA(cudaEventRecord(ev0, stream));

  // ---- creation order: (start word, seed-group rank of a, a, b) ------------------------------------------------------
  {
    // the mixed-radix key fits 64 bits when W * H^3 does; otherwise two stable passes (pair, then start word / rank)
    const long double span = static_cast<long double>(W) * H * H * H;
    const bool onePass = span < 1.8e19L && std::getenv("FSMC_ORDER_TWO_PASS") == nullptr;
    size_t bytes = 0;
    ORDER_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keysA, keysB, valsA, valsB, static_cast<int>(n), 0, 64, stream));
    ORDER_CUDA(temp(bytes));
    if (onePass) {
      creationKeyKernel<<<blocksFor(n), threads, 0, stream>>>(intervals, nullptr, n, H, rank, 0, keysA, valsA);
      const int bits = bitsFor(static_cast<unsigned long long>(W) * H * H * H);
      bytes = mTemp.bytes;
      ORDER_CUDA(cub::DeviceRadixSort::SortPairs(mTemp.p, bytes, keysA, keysB, valsA, valsB, static_cast<int>(n), 0, bits, stream));
    } else {
      creationKeyKernel<<<blocksFor(n), threads, 0, stream>>>(intervals, nullptr, n, H, rank, 1, keysA, valsA);
      bytes = mTemp.bytes;
      ORDER_CUDA(cub::DeviceRadixSort::SortPairs(mTemp.p, bytes, keysA, keysB, valsA, valsB, static_cast<int>(n), 0,
                                                 bitsFor(H * H), stream));
      creationKeyKernel<<<blocksFor(n), threads, 0, stream>>>(intervals, valsB, n, H, rank, 2, keysA, valsA);
      bytes = mTemp.bytes;
      ORDER_CUDA(cub::DeviceRadixSort::SortPairs(mTemp.p, bytes, keysA, keysB, valsA, valsB, static_cast<int>(n), 0,
                                                 bitsFor(static_cast<unsigned long long>(W) * H), stream));
    }
  }
  ORDER_CUDA(cudaMemsetAsync(mCounts.p, 0, nCounters * 8, stream));
  gatherNodesKernel<<<blocksFor(n), threads, 0, stream>>>(intervals, valsB, n, H, sites, genPos, minLengthCm, pairKey, node,
                                                          startBegin, startEnd, endCount, numLong);
  ORDER_CUDA(cudaGetLastError());
  std::vector<long long> hostCounts(nCounters);
  ORDER_CUDA(cudaMemcpyAsync(hostCounts.data(), mCounts.p, nCounters * 8, cudaMemcpyDeviceToHost, stream));
  ORDER_CUDA(cudaStreamSynchronize(stream));
  const long long numCandidates = hostCounts[3 * static_cast<size_t>(W)];

  // ---- rehash schedule from the counts alone (ref: boost reserve_for_insert) -----------------------------
  std::vector<long long> begin(static_cast<size_t>(W) + 1, 0);  // creation ranks of word w: [begin[w], begin[w + 1])
  {
    long long run = 0;
    for (int x = 0; x < W; ++x) {
      begin[x] = run;
      run += hostCounts[static_cast<size_t>(W) + x] - hostCounts[x];  // startEnd - startBegin (both 0 for an empty word)
    }
    begin[W] = run;
    if (run != n) {
      return cudaErrorUnknown;
    }
  }
  struct Epoch {
    long long q0;
    unsigned long long buckets;
  };
  std::vector<Epoch> epochs{{0, 17ull}};
  long long maxLive = 0;
  {
    long long live = 0;
    unsigned long long B = 17;
    for (int x = 0; x < W; ++x) {
      long long q = begin[x];
      const long long qEnd = begin[x + 1];
      while (q < qEnd) {
        const long long room = static_cast<long long>(B) - live;  // insertions that fit before size + 1 > max load
        if (q + room >= qEnd) {
          live += qEnd - q;
          q = qEnd;
        } else {
          q += room;
          live += room;
          B = growTo(static_cast<unsigned long long>(live));
          epochs.push_back(Epoch{q, B});
        }
      }
      maxLive = std::max(maxLive, live);
      if (x - gap - 1 >= 0) {
        live -= hostCounts[2 * static_cast<size_t>(W) + (x - gap - 1)];
      }
    }
  }
  // epoch in which the flush after word f happens: the last rehash at a creation rank < begin[f + 1]
  std::vector<int> phaseEpoch(static_cast<size_t>(W) + 1, 0);
  {
    size_t k = 0;
    for (int f = 0; f < W; ++f) {
      while (k + 1 < epochs.size() && epochs[k + 1].q0 < begin[f + 1]) {
        ++k;
      }
      phaseEpoch[f] = static_cast<int>(k);
    }
    phaseEpoch[W] = static_cast<int>(epochs.size()) - 1;
  }
  ORDER_CUDA(cudaMemcpyAsync(mPhaseEpoch.p, phaseEpoch.data(), phaseEpoch.size() * sizeof(int), cudaMemcpyHostToDevice, stream));

  const size_t C = static_cast<size_t>(std::max<long long>(numCandidates, 1));
  ORDER_CUDA(ensure(mCandKey, C * 8));
  ORDER_CUDA(ensure(mCandKeyB, C * 8));
  ORDER_CUDA(ensure(mCandQ, C * 4));
  ORDER_CUDA(ensure(mCandQB, C * 4));
  ORDER_CUDA(ensure(mCandPhase, C * 4));
  ORDER_CUDA(ensure(mCandPhaseB, C * 4));
  ORDER_CUDA(ensure(mOut, C * sizeof(fsmc_match)));
  ORDER_CUDA(ensure(mCarried, static_cast<size_t>(std::max<long long>(maxLive, 1)) * 4));
  ORDER_CUDA(ensure(mBucketFirst, static_cast<size_t>(epochs.back().buckets) * 4));
  uint32_t* carried = static_cast<uint32_t*>(mCarried.p);
  uint32_t* bucketFirst = static_cast<uint32_t*>(mBucketFirst.p);
  const CandidateList cand{static_cast<uint64_t*>(mCandKey.p), static_cast<uint32_t*>(mCandPhase.p),
                           static_cast<uint32_t*>(mCandQ.p), candCount};

  // ---- one pass per epoch ----------------------------------------------------------------------------------------------
  for (size_t k = 0; k < epochs.size(); ++k) {
    const long long q0 = epochs[k].q0;
    const long long q1 = k + 1 < epochs.size() ? epochs[k + 1].q0 : n;
    const unsigned long long B = epochs[k].buckets;
    long long live = 0;
    if (k > 0) {
      // the word in which the rehash happens: begin[x] <= q0 < begin[x + 1]
      const int x = static_cast<int>(std::upper_bound(begin.begin(), begin.end(), q0) - begin.begin()) - 1;
      size_t bytes = 0;
      const LiveAt pred{node, x - gap - 1};
      ORDER_CUDA(cub::DeviceSelect::If(nullptr, bytes, thrust::counting_iterator<uint32_t>(0), carried, numSelected,
                                       static_cast<int>(q0), pred, stream));
      ORDER_CUDA(temp(bytes));
      bytes = mTemp.bytes;
      ORDER_CUDA(cub::DeviceSelect::If(mTemp.p, bytes, thrust::counting_iterator<uint32_t>(0), carried, numSelected,
                                       static_cast<int>(q0), pred, stream));
      int selected = 0;
      ORDER_CUDA(cudaMemcpyAsync(&selected, numSelected, sizeof(int), cudaMemcpyDeviceToHost, stream));
      ORDER_CUDA(cudaStreamSynchronize(stream));
      live = selected;
      if (live > maxLive) {
        return cudaErrorUnknown;
      }
      if (live > 0) {
        // list order of the live nodes, then the re-keying of the rehash
        walkKeyKernel<<<blocksFor(live), threads, 0, stream>>>(carried, live, g, w, keysA);
        bytes = 0;
        ORDER_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keysA, keysB, carried, valsA, static_cast<int>(live), 0, 64, stream));
        ORDER_CUDA(temp(bytes));
        bytes = mTemp.bytes;
        ORDER_CUDA(cub::DeviceRadixSort::SortPairs(mTemp.p, bytes, keysA, keysB, carried, valsA, static_cast<int>(live), 0, 64, stream));
        ORDER_CUDA(cudaMemsetAsync(bucketFirst, 0xff, static_cast<size_t>(B) * 4, stream));
        rehashFirstKernel<<<blocksFor(live), threads, 0, stream>>>(valsA, live, pairKey, B, bucketFirst);
        rehashAssignKernel<<<blocksFor(live), threads, 0, stream>>>(valsA, live, pairKey, B, bucketFirst, g, w);
      }
    }
    const long long M = q1 - q0;
    const long long total = live + M;
    if (total == 0) {
      continue;
    }
    if (total > n) {
      return cudaErrorUnknown;
    }
    const uint32_t base = static_cast<uint32_t>(live) + 1u;
    epochKeyKernel<<<blocksFor(total), threads, 0, stream>>>(k > 0 ? valsA : carried, live, q0, M, pairKey, B, keysA);
    size_t bytes = 0;
    const int keyBits = 32 + bitsFor(B);
    ORDER_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keysA, keysB, static_cast<int>(total), 0, keyBits, stream));
    ORDER_CUDA(temp(bytes));
    bytes = mTemp.bytes;
    ORDER_CUDA(cub::DeviceRadixSort::SortKeys(mTemp.p, bytes, keysA, keysB, static_cast<int>(total), 0, keyBits, stream));
    bucketWalkKernel<<<blocksFor(total), threads, 0, stream>>>(keysB, total, q0, base, static_cast<int>(k), gap, W, node, g, w,
                                                              static_cast<const int*>(mPhaseEpoch.p), cand);
    ORDER_CUDA(cudaGetLastError());
  }

  // ---- emission order: (flush word, list order) -----------------------------------------------------------------------
  unsigned long long found = 0;
  ORDER_CUDA(cudaMemcpyAsync(&found, candCount, sizeof found, cudaMemcpyDeviceToHost, stream));
  ORDER_CUDA(cudaStreamSynchronize(stream));
  if (static_cast<long long>(found) != numCandidates) {
    return cudaErrorUnknown;  // every candidate leaves the map exactly once
  }
  if (numCandidates > 0) {
    uint32_t* idxA = static_cast<uint32_t*>(mCandQB.p);      // indices into the candidate list
    uint32_t* idxB = static_cast<uint32_t*>(mCandPhaseB.p);
    iotaKernel<<<blocksFor(numCandidates), threads, 0, stream>>>(idxA, numCandidates);
    size_t bytes = 0;
    ORDER_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, cand.key, static_cast<uint64_t*>(mCandKeyB.p), idxA, idxB,
                                               static_cast<int>(numCandidates), 0, 64, stream));
    ORDER_CUDA(temp(bytes));
    bytes = mTemp.bytes;
    ORDER_CUDA(cub::DeviceRadixSort::SortPairs(mTemp.p, bytes, cand.key, static_cast<uint64_t*>(mCandKeyB.p), idxA, idxB,
                                               static_cast<int>(numCandidates), 0, 64, stream));
    // stable second pass by flush word
    uint32_t* phaseSorted = reinterpret_cast<uint32_t*>(keysA);
    uint32_t* phaseOut = reinterpret_cast<uint32_t*>(keysB);
    gatherU32Kernel<<<blocksFor(numCandidates), threads, 0, stream>>>(cand.phase, idxB, numCandidates, phaseSorted);
    bytes = 0;
    ORDER_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, phaseSorted, phaseOut, idxB, idxA, static_cast<int>(numCandidates), 0,
                                               bitsFor(static_cast<unsigned long long>(W)), stream));
    ORDER_CUDA(temp(bytes));
    bytes = mTemp.bytes;
    ORDER_CUDA(cub::DeviceRadixSort::SortPairs(mTemp.p, bytes, phaseSorted, phaseOut, idxB, idxA, static_cast<int>(numCandidates), 0,
                                               bitsFor(static_cast<unsigned long long>(W)), stream));
    emitKernel<<<blocksFor(numCandidates), threads, 0, stream>>>(idxA, cand.q, numCandidates, pairKey, node, H,
                                                                 static_cast<fsmc_match*>(mOut.p));
    ORDER_CUDA(cudaGetLastError());
  }
  ORDER_CUDA(cudaEventRecord(ev1, stream));
  ORDER_CUDA(cudaStreamSynchronize(stream));
  *out = static_cast<const fsmc_match*>(mOut.p);
  *count = numCandidates;
  if (stats) {
    stats->candidates = numCandidates;
    stats->maxLive = maxLive;
    stats->epochs = static_cast<int>(epochs.size());
    cudaEventElapsedTime(&stats->deviceMs, ev0, ev1);
  }
  return cudaSuccess;
}

}  // namespace fsmc
