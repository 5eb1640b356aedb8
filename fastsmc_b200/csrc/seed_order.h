// fastsmc_b200 — the reference's candidate order, computed on the device.
//
// The reference hands candidates to the HMM in the iteration order of two boost::unordered_map instances
// (ref: ASMC_SRC/SRC/HASHING/SeedHash.hpp:34,80; HASHING/ExtendHash.hpp:29,85-116; boost 1.75), and batch composition
// — hence decode windows and segment boundaries — depends on that order (ref: HMM.cpp:561-565, 1199-1204).  EVERY
// match interval of the job is a node of the extend map, candidates or not (2 x 10^8 nodes for 2 x 10^7 candidates at
// 10 000 samples x 50 000 SNPs), so the round-1 host replay (host/CandidateOrder.hpp) was the longest stage of a run.
//
// The node order of a boost <= 1.79 table is a function of the insertion history, which is known up front:
//   * creation order of a word's new nodes = (iteration rank of haplotype a's word group in the seed map, a, b):
//     ONE radix sort of all intervals by (start word, rank, a, b);
//   * the map's size after every insertion follows from the per-word start / end counts alone, so the rehash schedule
//     (creation rank at which the bucket array grows, new bucket count) is a host loop over the words;
//   * between two rehashes (an EPOCH) buckets are independent: sort the epoch's nodes by (bucket, creation rank) and
//     walk each bucket with one thread.  A node joins the bucket's group if the bucket is non-empty when it is inserted
//     (some earlier node of the bucket leaves the map later), else it founds a new group at the front of the list;
//   * a rehash re-keys the live nodes: sort them by list order, new bucket = key mod new count, groups re-form in the
//     order their first node is met (atomicMin per bucket), inside a group in reverse walk order.
// Every node thus gets (g, w) with list order == descending (g, w); nodes leave the map after word end+gap+1, and a
// flush emits the leaving candidates in list order: one final sort by (flush word, g, w).
// tests/probes/order_epochs_proto.py is the numpy statement of the same algorithm, checked against the literal replay.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/fastsmc_b200.h"

namespace fsmc
{

struct OrderStats {
  long long intervals = 0, candidates = 0, maxLive = 0;
  int epochs = 0;
  float deviceMs = 0.f;
};

class CandidateOrderer
{
public:
  CandidateOrderer() = default;
  CandidateOrderer(const CandidateOrderer&) = delete;
  CandidateOrderer& operator=(const CandidateOrderer&) = delete;
  ~CandidateOrderer();

  // intervals : device array of all n match intervals of the job, any order (pairExtendKernel, FSMC_SEED_ALL_INTERVALS)
  // rank      : DEVICE array [numWords][numHaps], seed-map iteration rank of each haplotype's word group
  // genPos    : device array [sites]
  // On return *out points at the candidates (length >= minLengthCm) in the reference's decodeFromHashing call order
  // (device memory owned by the orderer, valid until the next call), *count = their number.
  cudaError_t order(const fsmc_match* intervals, long long n, uint32_t numHaps, int numWords, int sites, int gap,
                    float minLengthCm, const float* genPos, const uint32_t* rank, cudaStream_t stream,
                    const fsmc_match** out, long long* count, OrderStats* stats);

private:
  struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
  };
  cudaError_t ensure(Buf& b, size_t bytes);
  void release();
  Buf mKeysA, mKeysB, mValsA, mValsB, mPairKey, mSe, mG, mW, mBucketFirst, mCarried, mCandKey, mCandKeyB, mCandQ, mCandQB,
      mCandPhase, mCandPhaseB, mOut, mTemp, mCounts, mPhaseEpoch;
};

}  // namespace fsmc
