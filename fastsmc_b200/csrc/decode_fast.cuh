// fastsmc_b200 — the production decode kernel for sm_100a (FMA arithmetic; results within 1e-4 of the reference).
//
// Same two sweeps per 32-pair tile as decode_kernels.cuh (backward sweep streams beta to a per-warp HBM slab, forward
// sweep keeps alpha on chip and consumes the posterior at once), rebuilt around the three things the first kernel
// stalled on (profiles/r1_v1_*: 24 % issue-slot utilisation, 3.3 warps stalled on the long scoreboard per issue):
//
//  1. every HBM access of the sweeps is a bulk asynchronous copy (cp.async.bulk, the non-tensor TMA path) through a
//     per-warp shared-memory ring completed on mbarriers: one elected lane issues one 9 KB copy per site instead of
//     69 scalar loads/stores per lane, and the copy for site p+DEPTH is in flight while site p is computed.  beta rows
//     are laid out [state/4][lane][4] so that a lane moves four states per 128-bit shared-memory access;
//  2. the per-site coefficient row (three emission classes, D, B, U, RR: 2 KB) is prefetched into the same ring, so
//     the recurrences read their coefficients with warp-broadcast LDS.128 instead of waiting on L2;
//  3. the per-site rescaling to sum 1 (ref: HmmUtils.cpp:102-151, two instructions per state per sweep) is done every
//     fourth site only: the posterior alpha*beta/sum(alpha*beta) is invariant to the scale of alpha and of beta, so
//     scaling exists only to keep fp32 in range.  The forward sweep gets its normaliser for free from the suffix sums.
//     The combine step is a dot product (one FMA per state); per-state posteriors are formed only for lanes inside an
//     IBD segment or when per-site summaries are requested.
//
// State vectors stay in registers (fully unrolled linear-time recurrences, ref: HMM.cpp:787-879, 943-1041); the small
// site-independent vectors (columnRatios, expectedTimes) are kernel parameters, i.e. constant-bank operands.
#pragma once

#include "decode_kernels.cuh"

namespace fsmc
{

constexpr int kMaxParamStates = 160;

struct FastModel {
  DeviceModel base;
  float colRatios[kMaxParamStates];
  float expTimes[kMaxParamStates];
  float prior[kMaxParamStates];
};

// ---- bulk async copy / mbarrier primitives (PTX ISA: cp.async.bulk, mbarrier) -----------------------------------
__device__ __forceinline__ uint32_t smemPtr(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbarInit(uint64_t* bar, const uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemPtr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, const uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemPtr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, const uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smemPtr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, const uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemPtr(smemDst)),
               "l"(gmemSrc), "r"(bytes), "r"(smemPtr(bar))
               : "memory");
}
__device__ __forceinline__ void bulkStore(void* gmemDst, const void* smemSrc, const uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmemDst), "r"(smemPtr(smemSrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulkCommit()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulkWaitRead()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulkWaitAll()
{
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fenceProxyAsync()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float f4(const float4& v, const int j)
{
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}

// -------------------------------------------------------------------------------------------------------------------
// S_T   : states (compile time), DEPTH: ring depth, RESCALE: sites between rescalings (power of two)
// -------------------------------------------------------------------------------------------------------------------
// ACC   : keep per-state segment accumulators (FSMC_SEG_AGE) in registers
template <int S_T, int DEPTH, int RESCALE, bool ACC, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) decodeFastKernel(const FastModel fm, const DecodeArgs args)
{
  constexpr int S = S_T;
  constexpr int SQ = (S + 3) / 4;
  constexpr int Spad = SQ * 4;
  constexpr uint32_t kBetaBytes = SQ * 32 * 16;              // one site of beta for 32 lanes
  constexpr uint32_t kCoefBytes = kRowArrays * Spad * 4;     // one site's coefficient row
  constexpr size_t kBetaFloats = static_cast<size_t>(SQ) * 32 * 4;
  constexpr int kWarps = THREADS / 32;
  constexpr uint32_t kAccBytes = 0;  // the accumulators live in registers
  constexpr size_t kWarpBytes = static_cast<size_t>(DEPTH) * (kBetaBytes + kCoefBytes) + kAccBytes;

  extern __shared__ __align__(128) unsigned char smemRaw[];
  const DeviceModel m = fm.base;  // a copy: taking the address of a kernel parameter would move all of fm to local memory
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  unsigned char* mine = smemRaw + static_cast<size_t>(warp) * kWarpBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemRaw + static_cast<size_t>(kWarps) * kWarpBytes) +
                   warp * 2 * DEPTH;  // [0,DEPTH): beta slot full, [DEPTH,2*DEPTH): coefficient slot full
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
  uint32_t betaParity = 0, coefParity = 0;  // bit i = parity the next wait on slot i expects

  const unsigned flags = args.flags;
  const bool wantSeg = flags & FSMC_CALL_SEGMENTS;
  const bool wantAge = ACC && (flags & FSMC_SEG_AGE) && wantSeg;
  const bool wantSite = flags & (FSMC_SITE_MEAN | FSMC_SITE_MAP);
  const int sT = m.stateThreshold;
  const int nAcc = m.ageThreshold;
  const long long warpGlobal = static_cast<long long>(blockIdx.x) * kWarps + warp;
  float* slab = args.scratch + warpGlobal * args.scratchPerWarp;                      // beta rows of this warp


  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) {
      t = atomicAdd(args.tileCounter, 1ull);
    }
    t = __shfl_sync(kFull, t, 0);
    if (static_cast<long long>(t) >= args.numTiles) {
      break;
    }
    const int tile = args.order ? args.order[t] : static_cast<int>(t);
    const int nPairs = args.tilePairs[tile];
    const int from = args.tileFrom[tile];
    const int len = args.tileTo[tile] - from;
    const int scanFrom = wantSeg ? args.tileScanFrom[tile] : 0;
    const int scanTo = wantSeg ? args.tileScanTo[tile] : 0;
    const bool laneActive = lane < nPairs;
    const int srcLane = laneActive ? lane : nPairs - 1;
    const uint32_t pair = static_cast<uint32_t>(tile) * 32u + static_cast<uint32_t>(lane);
    PairBits bits;
    bits.a = m.haps + static_cast<size_t>(args.hapA[static_cast<size_t>(tile) * 32 + srcLane]) * m.wordsPerHap;
    bits.b = m.haps + static_cast<size_t>(args.hapB[static_cast<size_t>(tile) * 32 + srcLane]) * m.wordsPerHap;
    const float* rowBase = m.siteRows + static_cast<size_t>(from) * kRowArrays * Spad;  // row of window position 0

    float a[S], c[S];
    float acc[ACC ? S : 1];  // per-state posterior sums of the lane's current IBD run (FSMC_SEG_AGE)
#pragma unroll
    for (int k = 0; k < (ACC ? S : 1); ++k) {
      acc[k] = 0.f;
    }

    // =============================================================================================================
    // sweep 1: backward (ref: HMM.cpp:882-1041).  Step j = 0 .. len-2 handles p = len-2-j with the row of p+1.
    // =============================================================================================================
    {
      auto prefetchCoef = [&](const int j) {  // row of window position len-1-j into slot j % DEPTH
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
      // beta at the last site: all ones (any positive scale is equivalent)
      {
        float4* out = betaSlot(0);
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          out[q * 32 + lane] = make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int k = 0; k < S; ++k) {
          a[k] = 1.f;
        }
        fenceProxyAsync();
        __syncwarp();
        if (lane == 0) {
          bulkStore(slab + static_cast<size_t>(len - 1) * kBetaFloats, out, kBetaBytes);
          bulkCommit();
        }
      }
      for (int j = 0; j < steps; ++j) {
        const int p = len - 2 - j;
        const int slot = j % DEPTH;
        const int cls = bits.cls(from + p + 1);
        mbarWait(&bars[DEPTH + slot], (coefParity >> slot) & 1u);
        coefParity ^= 1u << slot;
        const float* row = coefSlot(slot);
        const float4* E = reinterpret_cast<const float4*>(row + cls * Spad);
        const float4* Dr = reinterpret_cast<const float4*>(row + 3 * Spad);
        const float4* Br = reinterpret_cast<const float4*>(row + 4 * Spad);
        const float4* Ur = reinterpret_cast<const float4*>(row + 5 * Spad);
        const float4* Rr = reinterpret_cast<const float4*>(row + 6 * Spad);
        // vec = beta(p+1) * emission(p+1), in place in a
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          const float4 e4 = E[q];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = 4 * q + i;
            if (k < S) {
              a[k] *= f4(e4, i);
            }
          }
        }
        // BU[k] = U[k] vec[k+1] + RR[k] BU[k+1]
        {
          float bu = 0.f;
#pragma unroll
          for (int q = SQ - 1; q >= 0; --q) {
            const float4 u4 = Ur[q];
            const float4 r4 = Rr[q];
#pragma unroll
            for (int i = 3; i >= 0; --i) {
              const int k = 4 * q + i;
              if (k == S - 1) {
                c[k] = 0.f;
              } else if (k < S - 1) {
                bu = fmaf(f4(r4, i), bu, f4(u4, i) * a[k + 1]);
                c[k] = bu;
              }
            }
          }
        }
        // beta(p)[k] = BL + D[k] vec[k] + BU[k], BL += B[k-1] vec[k-1]; written over vec in a
        {
          float bl = 0.f, bPrev = 0.f, vPrev = 0.f;
#pragma unroll
          for (int q = 0; q < SQ; ++q) {
            const float4 d4 = Dr[q];
            const float4 b4 = Br[q];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = 4 * q + i;
              if (k < S) {
                const float v = a[k];
                if (k) {
                  bl = fmaf(bPrev, vPrev, bl);
                }
                a[k] = fmaf(f4(d4, i), v, bl) + c[k];
                bPrev = f4(b4, i);
                vPrev = v;
              }
            }
          }
        }
        __syncwarp();  // every lane is done with the coefficient slot
        if (j + DEPTH < steps) {
          prefetchCoef(j + DEPTH);
        }
        if ((p & (RESCALE - 1)) == 0) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int k = 0; k < S; k += 4) {
            s0 += a[k];
            if (k + 1 < S) s1 += a[k + 1];
            if (k + 2 < S) s2 += a[k + 2];
            if (k + 3 < S) s3 += a[k + 3];
          }
          const float sc = 1.0f / ((s0 + s1) + (s2 + s3));
#pragma unroll
          for (int k = 0; k < S; ++k) {
            a[k] *= sc;
          }
        }
        // stage the row and hand it to the copy engine
        const int bslot = (j + 1) % DEPTH;
        if (lane == 0) {
          bulkWaitRead<DEPTH - 1>();  // the copy that last read this staging slot has drained it
        }
        __syncwarp();
        float4* out = betaSlot(bslot);
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          out[q * 32 + lane] = make_float4(a[4 * q], 4 * q + 1 < S ? a[4 * q + 1] : 0.f, 4 * q + 2 < S ? a[4 * q + 2] : 0.f,
                                           4 * q + 3 < S ? a[4 * q + 3] : 0.f);
        }
        fenceProxyAsync();
        __syncwarp();
        if (lane == 0) {
          bulkStore(slab + static_cast<size_t>(p) * kBetaFloats, out, kBetaBytes);
          bulkCommit();
        }
      }
      if (lane == 0) {
        bulkWaitAll<0>();  // all beta rows are in global memory before the forward sweep reads them back
      }
      __syncwarp();
    }

    // =============================================================================================================
    // sweep 2: forward + fused consumers (ref: HMM.cpp:725-879, 669-692, 1179-1357, 1378-1409)
    // =============================================================================================================
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
      CallerState cs;
      for (int p = 0; p < len; ++p) {
        const int site = from + p;
        const int slot = p % DEPTH;
        const int cls = bits.cls(site);
        mbarWait(&bars[DEPTH + slot], (coefParity >> slot) & 1u);
        coefParity ^= 1u << slot;
        const float* row = coefSlot(slot);
        const float4* E = reinterpret_cast<const float4*>(row + cls * Spad);
        if (p == 0) {
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
        } else {
          const float4* Dr = reinterpret_cast<const float4*>(row + 3 * Spad);
          const float4* Br = reinterpret_cast<const float4*>(row + 4 * Spad);
          const float4* Ur = reinterpret_cast<const float4*>(row + 5 * Spad);
          // alphaC[k] = sum_{j>=k} alpha(p-1)[j]; its head is the normaliser of alpha(p-1), for free
          {
            float run = 0.f;
#pragma unroll
            for (int k = S - 1; k >= 0; --k) {
              run = (k == S - 1) ? a[k] : run + a[k];
              c[k] = run;
            }
          }
          const bool rescale = (p & (RESCALE - 1)) == 0;
          const float sc = rescale ? 1.0f / c[0] : 1.0f;
          float au = 0.f, uPrev = 0.f, crPrev = 0.f, aPrev = 0.f;
#pragma unroll
          for (int q = 0; q < SQ; ++q) {
            const float4 e4 = E[q];
            const float4 d4 = Dr[q];
            const float4 b4 = Br[q];
            const float4 u4 = Ur[q];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = 4 * q + i;
              if (k < S) {
                const float ak = a[k];
                if (k) {
                  au = fmaf(crPrev, au, uPrev * aPrev);
                }
                float term = fmaf(f4(d4, i), ak, au);
                if (k < S - 1) {
                  term = fmaf(f4(b4, i), c[k + 1], term);
                }
                a[k] = f4(e4, i) * term;
                uPrev = f4(u4, i);
                crPrev = fm.colRatios[k];
                aPrev = ak;
              }
            }
          }
          if (rescale) {
#pragma unroll
            for (int k = 0; k < S; ++k) {
              a[k] *= sc;
            }
          }
        }

        // ---- combine: q[k] = alpha[k] beta[k] (ref: HMM.cpp:672-680); the products stay in c for the consumers
        mbarWait(&bars[slot], (betaParity >> slot) & 1u);
        betaParity ^= 1u << slot;
        const float4* B4 = betaSlot(slot);
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, ibdRaw = 0.f;
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          const float4 b4 = B4[q * 32 + lane];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = 4 * q + i;
            if (k < S) {
              c[k] = a[k] * f4(b4, i);
            }
          }
          q0 += c[4 * q];
          if (4 * q + 1 < S) q1 += c[4 * q + 1];
          if (4 * q + 2 < S) q2 += c[4 * q + 2];
          if (4 * q + 3 < S) q3 += c[4 * q + 3];
        }
        const float r = 1.0f / ((q0 + q1) + (q2 + q3));

        if (wantSite) {
          float mean = 0.f, best = 0.f;
          int arg = 0;
#pragma unroll
          for (int k = 0; k < S; ++k) {
            const float post = c[k];
            mean = fmaf(post, fm.expTimes[k], mean);
            if (best < post) {
              best = post;
              arg = k;
            }
          }
          if (laneActive) {
            if (flags & FSMC_SITE_MEAN) {
              args.siteMean[static_cast<size_t>(pair) * args.siteStride + p] = mean * r;
            }
            if (flags & FSMC_SITE_MAP) {
              args.siteMap[static_cast<size_t>(pair) * args.siteStride + p] = arg;
            }
          }
        }

        const bool inScan = wantSeg && site >= scanFrom && site < scanTo;
        if ((flags & FSMC_SITE_IBD) || inScan) {
          forStatesBelow<S_T>(sT, [&](const int k) { ibdRaw += c[k]; });
          const float ibd = ibdRaw * r;
          if ((flags & FSMC_SITE_IBD) && laneActive) {
            args.siteIbd[static_cast<size_t>(pair) * args.siteStride + p] = ibd;
          }
          if (inScan) {
            int now = -1;
            if (ibd >= m.thr[0]) {
              now = 0;
            } else if (ibd >= m.thr[1]) {
              now = 1;
            } else if (ibd >= m.thr[2]) {
              now = 2;
            } else if (ibd >= m.thr[3]) {
              now = 3;
            }
            if (!laneActive) {
              now = -1;
            }
            const bool changed = now != cs.level;
            const bool ending = changed && cs.level >= 0;  // the run that ended at site-1 is written now
            const bool closing = now >= 0 && site == scanTo - 1;
            if constexpr (ACC) {
              if (wantAge) {
                if (__any_sync(kFull, ending)) {
                  // rare: park the per-state sums (through site-1) in this site's drained beta slot for emitSegment
                  float* park = reinterpret_cast<float*>(betaSlot(slot)) + lane;
#pragma unroll
                  for (int k = 0; k < S; ++k) {
                    park[k * 32] = acc[k];
                  }
                  __syncwarp();
                  if (ending) {
                    emitSegment<false>(m, args, pair, cs.start, site - 1, cs.prob, cs.level, park, true);
                  }
                  __syncwarp();
                }
                if (__any_sync(kFull, now >= 0)) {
                  // per-state sums of the current run: restart on a new run, accumulate otherwise
                  // (ref: HMM.cpp:1209-1218,1229,1257,1284,1311)
                  const float rr = now >= 0 ? r : 0.f;
                  const float keep = changed ? 0.f : 1.f;
                  forStatesBelow<S_T>(nAcc, [&](const int k) { acc[k] = fmaf(c[k], rr, keep * acc[k]); });
                }
                if (__any_sync(kFull, closing)) {
                  float* park = reinterpret_cast<float*>(betaSlot(slot)) + lane;
#pragma unroll
                  for (int k = 0; k < S; ++k) {
                    park[k * 32] = acc[k];
                  }
                  __syncwarp();
                  if (closing) {
                    emitSegment<false>(m, args, pair, cs.start, site, changed ? ibd : cs.prob + ibd, now, park, true);
                  }
                  __syncwarp();
                }
              }
            }
            if (!wantAge) {
              if (ending) {
                emitSegment<false>(m, args, pair, cs.start, site - 1, cs.prob, cs.level, nullptr, false);
              }
              if (closing) {
                emitSegment<false>(m, args, pair, cs.start, site, changed ? ibd : cs.prob + ibd, now, nullptr, false);
              }
            }
            if (now >= 0) {
              if (changed) {
                cs.start = site;
                cs.prob = ibd;
              } else {
                cs.prob += ibd;
              }
              if (closing) {
                cs.prob = 0.f;
              }
            } else {
              cs.prob = 0.f;
            }
            cs.level = now;
          }
        }
        __syncwarp();  // both slots of this position are drained (the beta slot doubles as the parking area above)
        if (p + DEPTH < len) {
          prefetch(p + DEPTH);
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace fsmc
