// fastsmc_b200 — all-state age estimates without the beta round trip ("sparse refinement").
//
// FastSMC's regression flags (noConditionalAgeEstimates) ask, for every IBD segment, for the per-state sums of the
// posterior over the segment's sites (ref: HMM.cpp:1191-1197, 1209-1218, 1087-1107).  Only the sites INSIDE segments need
// all S states; which sites those are follows from the states below the IBD time threshold alone, which the narrow kernel
// computes without streaming beta through HBM.  With hashing off (every pair decoded over the whole chromosome) a few
// percent of the pair-sites lie inside segments, so instead of materialising beta everywhere (8*S bytes per pair-site):
//
//   pass 1  decodeNarrowKernel<SPARSE>: segments are called as usual; the backward sweep leaves a full beta vector at the
//           last site of every block of 2^ckptShift sites, the forward sweep records one ITEM per (run, block) together
//           with alpha at the block's first site;
//   sort    items by block (radix sort), so that 32 items that share a block share the coefficient rows;
//   pass 2  refineKernel: one warp per 32 items of a block; both recurrences are re-run over the block from the two
//           checkpoints (beta rows of the block go through a per-warp slab that stays in L2, with the bulk-copy rings of
//           decodeFastKernel) and the posterior's per-state sums over the item's sites are written per item;
//   pass 3  finalizeSegmentsKernel: per segment, the sums of its chain of items -> posterior mean and MAP of the segment.
//
// HBM traffic per pair-site: 2 bits of genotypes, the narrow record (16 B written + read) and 8*S/2^ckptShift of
// checkpoints, instead of 8*S.  The kernel is bound by the FP32 pipe like the narrow kernel.
#pragma once

#include "decode_fast.cuh"

namespace fsmc
{

constexpr int kRefineDepth = 2;

// sort key of an item = the first site of its block (blocks are counted from the window's first site; items whose
// blocks start at the same site share the coefficient rows); slots beyond the item count get the largest key
__global__ void itemKeysKernel(const SparseItem* __restrict__ items, const unsigned long long* __restrict__ itemCount,
                               const long long capacity, const int* __restrict__ tileFrom, const int ckptShift,
                               uint32_t* __restrict__ keys, uint32_t* __restrict__ index)
{
  const long long n = static_cast<long long>(min(*itemCount, static_cast<unsigned long long>(capacity)));
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < capacity;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    keys[i] = i < n ? static_cast<uint32_t>(tileFrom[items[i].pair >> 5] + (items[i].block << ckptShift)) : 0xffffffffu;
    index[i] = static_cast<uint32_t>(i);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// refineKernel: lane = item.  The 32 items of a warp share the first site of their block; the warp sweeps the block (to
// the last site any of its pairs' windows reaches in it).  A lane joins the backward sweep at the last site of its pair's
// window in the block (beta checkpoint); the forward sweep starts for all lanes at the block's first site, from the item's
// alpha of the site before (or from the prior when the block is the first of the window), and every lane accumulates the
// posterior over its item's own sites.  Lanes outside their range compute on zeros / garbage, which is never looked at.
// ---------------------------------------------------------------------------------------------------------------------
template <int S_T, int DEPTH, int RESCALE, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) refineKernel(const FastModel fm, const DecodeArgs args)
{
  constexpr int S = S_T;
  constexpr int SQ = (S + 3) / 4;
  constexpr int Spad = SQ * 4;
  constexpr uint32_t kBetaBytes = SQ * 32 * 16;
  constexpr uint32_t kCoefBytes = kRowArrays * Spad * 4;
  constexpr size_t kBetaFloats = static_cast<size_t>(SQ) * 32 * 4;
  constexpr int kWarps = THREADS / 32;
  constexpr size_t kWarpBytes = static_cast<size_t>(DEPTH) * (kBetaBytes + kCoefBytes);

  extern __shared__ __align__(128) unsigned char smemRaw[];
  const DeviceModel m = fm.base;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  unsigned char* mine = smemRaw + static_cast<size_t>(warp) * kWarpBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemRaw + static_cast<size_t>(kWarps) * kWarpBytes) + warp * 2 * DEPTH;
  auto betaSlot = [&](const int i) { return reinterpret_cast<float4*>(mine + static_cast<size_t>(i) * kBetaBytes); };
  auto coefSlot = [&](const int i) {
    return reinterpret_cast<const float*>(mine + static_cast<size_t>(DEPTH) * kBetaBytes + static_cast<size_t>(i) * kCoefBytes);
  };
  if (lane == 0) {
    for (int i = 0; i < 2 * DEPTH; ++i) {
      mbarInit(&bars[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t betaParity = 0, coefParity = 0;

  const long long warpGlobal = static_cast<long long>(blockIdx.x) * kWarps + warp;
  float* slab = args.scratch + warpGlobal * args.scratchPerWarp;  // beta rows of the block being refined
  const int ckShift = args.ckptShift;
  const int C = 1 << ckShift;
  const long long numItems = static_cast<long long>(min(*args.itemCount, static_cast<unsigned long long>(args.itemCapacity)));
  const long long numTasks = (numItems + 31) / 32;

  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) {
      t = atomicAdd(args.refineCounter, 1ull);
    }
    t = __shfl_sync(kFull, t, 0);
    if (static_cast<long long>(t) >= numTasks) {
      break;
    }
    const long long at = static_cast<long long>(t) * 32 + lane;
    const bool valid = at < numItems;
    const uint32_t itemIndex = valid ? args.itemOrder[at] : 0u;
    SparseItem item{0u, -1, 0, -1, -1};
    int tFrom = 0, tTo = 0;
    long long slot0 = 0;
    PairBits bits;
    bits.a = m.haps;
    bits.b = m.haps;
    if (valid) {
      item = args.items[itemIndex];
      const uint32_t tile = item.pair >> 5;
      tFrom = args.tileFrom[tile];
      tTo = args.tileTo[tile];
      slot0 = args.tileCkptBase[tile];
      bits.a = m.haps + static_cast<size_t>(args.hapA[item.pair]) * m.wordsPerHap;
      bits.b = m.haps + static_cast<size_t>(args.hapB[item.pair]) * m.wordsPerHap;
    }
    const int myStart = valid ? tFrom + (item.block << ckShift) : -1;  // first site of the item's block
    unsigned todo = __ballot_sync(kFull, valid);
    while (todo) {  // the 32 items are sorted by block start: one round, two where the warp straddles a boundary
      const int lo = __shfl_sync(kFull, myStart, __ffs(todo) - 1);
      const bool on = valid && myStart == lo;
      todo &= ~__ballot_sync(kFull, on);
      // last site of the pair's window inside the block, and the furthest over the lanes
      const int e0 = on ? min(tTo - 1, lo + C - 1) : -1;
      int hi = e0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        hi = max(hi, __shfl_xor_sync(kFull, hi, o));
      }
      const int len = hi - lo + 1;
      const bool firstBlock = item.block == 0;  // alpha at the window's first site comes from the prior
      const float* rowBase = m.siteRows + static_cast<size_t>(lo) * kRowArrays * Spad;
      bits.word = -1;

      float a[S], c[S], acc[S];
#pragma unroll
      for (int k = 0; k < S; ++k) {
        a[k] = 0.f;
        acc[k] = 0.f;
      }
      auto loadVector = [&](float (&v)[S], const float4* src, const int stride) {
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          const float4 x = src[static_cast<size_t>(q) * stride];
          v[4 * q] = x.x;
          if (4 * q + 1 < S) v[4 * q + 1] = x.y;
          if (4 * q + 2 < S) v[4 * q + 2] = x.z;
          if (4 * q + 3 < S) v[4 * q + 3] = x.w;
        }
      };
      // beta at the last site of the pair's window in this block: the checkpoint of pass 1, column of the pair's lane
      auto loadBeta = [&](float (&v)[S]) {
        loadVector(v, reinterpret_cast<const float4*>(args.ckptBeta + static_cast<size_t>(slot0 + item.block) * kBetaFloats) + (item.pair & 31u), 32);
      };
      auto storeRow = [&](const float (&v)[S], const int p, const int bslot) {
        if (lane == 0) {
          bulkWaitRead<DEPTH - 1>();
        }
        __syncwarp();
        float4* out = betaSlot(bslot);
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          out[q * 32 + lane] = make_float4(v[4 * q], 4 * q + 1 < S ? v[4 * q + 1] : 0.f, 4 * q + 2 < S ? v[4 * q + 2] : 0.f,
                                           4 * q + 3 < S ? v[4 * q + 3] : 0.f);
        }
        fenceProxyAsync();
        __syncwarp();
        if (lane == 0) {
          bulkStore(slab + static_cast<size_t>(p) * kBetaFloats, out, kBetaBytes);
          bulkCommit();
        }
      };

      // ---- backward over [lo, hi] ----------------------------------------------------------------------------------
      {
        auto prefetchCoef = [&](const int j) {
          if (lane == 0) {
            uint64_t* bar = &bars[DEPTH + j % DEPTH];
            mbarExpectTx(bar, kCoefBytes);
            bulkLoad(const_cast<float*>(coefSlot(j % DEPTH)), rowBase + static_cast<size_t>(len - 1 - j) * kRowArrays * Spad,
                     kCoefBytes, bar);
          }
        };
        const int steps = len - 1;
        for (int j = 0; j < DEPTH && j < steps; ++j) {
          prefetchCoef(j);
        }
        if (on && e0 == hi) {
          loadBeta(a);
        }
        storeRow(a, len - 1, 0);
        auto step = [&](const int j, float (&x)[S], float (&y)[S]) {
          const int p = len - 2 - j;
          const int slot = j % DEPTH;
          const int cls = bits.cls(lo + p + 1);
          mbarWait(&bars[DEPTH + slot], (coefParity >> slot) & 1u);
          coefParity ^= 1u << slot;
          backwardStep<S>(x, y, coefSlot(slot), cls);
          __syncwarp();
          if (j + DEPTH < steps) {
            prefetchCoef(j + DEPTH);
          }
          if ((p & (RESCALE - 1)) == 0) {
            scaleStates<S>(y, 1.0f / sumStates<S>(y));
          }
          if (on && e0 == lo + p) {
            loadBeta(y);  // this lane's window ends here: it joins the sweep
          }
          storeRow(y, p, (j + 1) % DEPTH);
        };
        int j = 0;
        for (; j + 1 < steps; j += 2) {
          step(j, a, c);
          step(j + 1, c, a);
        }
        if (j < steps) {
          step(j, a, c);
        }
        if (lane == 0) {
          bulkWaitAll<0>();
        }
        __syncwarp();
      }

      // ---- forward over [lo, hi] + per-state sums over the item's sites -------------------------------------------------
      {
        auto prefetch = [&](const int p) {
          if (lane == 0) {
            const int slot = p % DEPTH;
            mbarExpectTx(&bars[slot], kBetaBytes);
            bulkLoad(betaSlot(slot), slab + static_cast<size_t>(p) * kBetaFloats, kBetaBytes, &bars[slot]);
            mbarExpectTx(&bars[DEPTH + slot], kCoefBytes);
            bulkLoad(const_cast<float*>(coefSlot(slot)), rowBase + static_cast<size_t>(p) * kRowArrays * Spad, kCoefBytes,
                     &bars[DEPTH + slot]);
          }
        };
        for (int p = 0; p < DEPTH && p < len; ++p) {
          prefetch(p);
        }
        const float4* alphaSrc = reinterpret_cast<const float4*>(args.itemAlpha + static_cast<size_t>(itemIndex) * Spad);
        auto consume = [&](const int p, const float (&v)[S], float (&w)[S]) {
          const int site = lo + p;
          const int slot = p % DEPTH;
          mbarWait(&bars[slot], (betaParity >> slot) & 1u);
          betaParity ^= 1u << slot;
          const bool mineSite = on && site >= item.first && site <= item.last;
          if (__any_sync(kFull, mineSite)) {  // sites between a block's first site and the item's are only swept
            const float4* B4 = betaSlot(slot);
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int q = 0; q < SQ; ++q) {
              const float4 b4 = B4[q * 32 + lane];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int k = 4 * q + i;
                if (k < S) {
                  w[k] = v[k] * f4(b4, i);
                }
              }
              q0 += w[4 * q];
              if (4 * q + 1 < S) q1 += w[4 * q + 1];
              if (4 * q + 2 < S) q2 += w[4 * q + 2];
              if (4 * q + 3 < S) q3 += w[4 * q + 3];
            }
            const float r = mineSite ? 1.0f / ((q0 + q1) + (q2 + q3)) : 0.f;  // ref HMM.cpp:681-685
#pragma unroll
            for (int k = 0; k < S; ++k) {
              acc[k] = fmaf(mineSite ? w[k] : 0.f, r, acc[k]);
            }
          }
          __syncwarp();  // both slots of this position are drained
          if (p + DEPTH < len) {
            prefetch(p + DEPTH);
          }
        };
        // p = 0, the block's first site: one step from the parked alpha of the site before, or prior * emission when the
        // block opens the window (ref HMM.cpp:736-743)
        {
          const int cls = bits.cls(lo);
          mbarWait(&bars[DEPTH], coefParity & 1u);
          coefParity ^= 1u;
#pragma unroll
          for (int k = 0; k < S; ++k) {
            c[k] = 0.f;
          }
          if (on && !firstBlock) {
            loadVector(c, alphaSrc, 1);
          }
          forwardStep<S>(fm.colRatios, c, a, coefSlot(0), cls);
          if (on && firstBlock) {
            const float4* E = reinterpret_cast<const float4*>(coefSlot(0) + cls * Spad);
#pragma unroll
            for (int q = 0; q < SQ; ++q) {
              const float4 e4 = E[q];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int k = 4 * q + i;
                if (k < S) {
                  a[k] = fm.prior[k] * f4(e4, i);
                }
              }
            }
          }
          consume(0, a, c);
        }
        auto step = [&](const int p, float (&x)[S], float (&y)[S]) {
          const int slot = p % DEPTH;
          const int cls = bits.cls(lo + p);
          mbarWait(&bars[DEPTH + slot], (coefParity >> slot) & 1u);
          coefParity ^= 1u << slot;
          const float total = forwardStep<S>(fm.colRatios, x, y, coefSlot(slot), cls);
          if ((p & (RESCALE - 1)) == 0) {
            scaleStates<S>(y, 1.0f / total);
          }
          consume(p, y, x);
        };
        int p = 1;
        for (; p + 1 < len; p += 2) {
          step(p, a, c);
          step(p + 1, c, a);
        }
        if (p < len) {
          step(p, a, c);
        }
      }
      if (on) {
        float4* dst = reinterpret_cast<float4*>(args.itemSums + static_cast<size_t>(itemIndex) * Spad);
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          dst[q] = make_float4(acc[4 * q], 4 * q + 1 < S ? acc[4 * q + 1] : 0.f, 4 * q + 2 < S ? acc[4 * q + 2] : 0.f,
                               4 * q + 3 < S ? acc[4 * q + 3] : 0.f);
        }
      }
      __syncwarp();
    }
  }
}

// One warp per segment: the sums of its chain of items (pieces in descending site order; lane l adds up states l, l+32,
// l+64: 288-byte rows read coalesced), then posterior mean and MAP as HMM::getPosteriorMean / getMAP do
// (ref: HMM.cpp:1087-1107; the first maximum wins a tie).
__global__ void finalizeSegmentsKernel(const DeviceModel m, fsmc_segment* __restrict__ segments,
                                       const unsigned long long* __restrict__ segmentCount, const long long segmentCapacity,
                                       const SparseItem* __restrict__ items, const float* __restrict__ itemSums,
                                       const unsigned long long* __restrict__ itemCount, const long long itemCapacity)
{
  const long long n = static_cast<long long>(min(*segmentCount, static_cast<unsigned long long>(segmentCapacity)));
  if (*itemCount > static_cast<unsigned long long>(itemCapacity)) {
    return;  // the host re-runs the request with a larger item buffer
  }
  const int S = m.ageThreshold;
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  constexpr int kRounds = (kMaxParamStates + 31) / 32;
  for (long long i = warp; i < n; i += warps) {
    const int last = segments[i].mapState;
    if (last < 0) {
      if (lane == 0) {
        segments[i].mapState = -1;
      }
      continue;
    }
    float x[kRounds];
#pragma unroll
    for (int j = 0; j < kRounds; ++j) {
      x[j] = 0.f;
    }
    for (int it = last; it >= 0; it = items[it].prev) {
      const float* row = itemSums + static_cast<size_t>(it) * m.Spad;
#pragma unroll
      for (int j = 0; j < kRounds; ++j) {
        const int k = lane + 32 * j;
        if (k < S) {
          x[j] += row[k];
        }
      }
    }
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < kRounds; ++j) {
      tot += x[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot += __shfl_xor_sync(kFull, tot, o);
    }
    const float norm = 1.f / tot;
    float mean = 0.f, bestRatio = -1.f;
    int best = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kRounds; ++j) {
      const int k = lane + 32 * j;
      if (k < S) {
        mean += (norm * x[j]) * __ldg(m.expTimes + k);
        const float r = x[j] / __ldg(m.prior + k);
        if (r > bestRatio) {  // ascending k within the lane: the first maximum is kept
          bestRatio = r;
          best = k;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mean += __shfl_xor_sync(kFull, mean, o);
      const float otherRatio = __shfl_xor_sync(kFull, bestRatio, o);
      const int otherBest = __shfl_xor_sync(kFull, best, o);
      if (otherRatio > bestRatio || (otherRatio == bestRatio && otherBest < best)) {
        bestRatio = otherRatio;
        best = otherBest;
      }
    }
    if (lane == 0) {
      segments[i].postMean = mean;
      segments[i].mapState = best;
      segments[i].mapTime = __ldg(m.expTimes + best);
    }
  }
}

}  // namespace fsmc
