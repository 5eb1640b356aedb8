// fastsmc_b200 — selection of the state-split decode kernels (decode_split.cuh).  The instantiations live in their own
// translation units (split_69.cu, split_159.cu) so that the library's sources compile in parallel.
#pragma once

#include <cstddef>

#include "decode_fast.cuh"

namespace fsmc
{

typedef void (*FastKernelFn)(const FastModel, const DecodeArgs);

struct SplitChoice {
  FastKernelFn fn = nullptr;
  int warps = 0;          // warps per tile == warps per CTA
  size_t smemBytes = 0;   // dynamic shared memory per CTA
  int recordQuads = 0;    // 0 = wide (full beta rows in HBM)
  int Spad = 0;
  bool acc = false;       // per-state segment accumulators in registers
};

// recordQuads: 0 for the wide kernel, else ceil((stateThreshold + 1) / 4).  Returns fn == nullptr when there is no
// instantiation for the request.
SplitChoice splitKernel69(int recordQuads, bool acc);
SplitChoice splitKernel159(int recordQuads, bool acc);

// Lane-split kernels (decode_lane.cuh, lane_159.cu) at 159 states: records of 1 to 4 quads (FastSMC_exe's default flags),
// or recordQuads = 0: full beta rows (per-site mean / MAP, age estimates over all states).
// It reads the per-site laneAux table of the model (buildLaneAux159).
SplitChoice laneKernel159(int recordQuads);
size_t laneAuxFloats159();
void buildLaneAux159(int L, const float* rows, float* aux, int blocks, cudaStream_t st);

}  // namespace fsmc
