// fastsmc_b200 — lane-split decode kernel (decode_lane.cuh) for the 159-state decoding quantities (FASTSMC_EXAMPLE).
#include <cstdlib>

#include "decode_lane.cuh"
#include "split_select.h"

namespace fsmc
{

constexpr int kLaneGroup = 4, kLaneDepth = 2;

template <int S, int RQ, int MINB> static SplitChoice makeLane()
{
  return SplitChoice{decodeLaneKernel<S, RQ, kLaneGroup, kLaneDepth, MINB>, kLaneQuarters,
                     LaneSmem<S, RQ, kLaneGroup, kLaneDepth>::kTotal, RQ, LaneGeom<S>::Spad, false};
}

SplitChoice laneKernel159(const int recordQuads)
{
  static const bool four = [] { const char* e = std::getenv("FSMC_LANE"); return e && *e == '4'; }();  // A/B: 4 CTAs per SM (spills)
  switch (recordQuads) {
  case 1:
    return four ? makeLane<159, 1, 4>() : makeLane<159, 1, 3>();
  case 2:
    return four ? makeLane<159, 2, 4>() : makeLane<159, 2, 3>();
  default:
    return {};
  }
}

size_t laneAuxFloats159()
{
  return LaneGeom<159>::kAuxFloats;
}

void buildLaneAux159(const int L, const float* rows, float* aux, const int blocks, cudaStream_t st)
{
  buildLaneAuxKernel<<<blocks, 256, 0, st>>>(LaneGeom<159>::Spad, L, rows, aux);
}

}  // namespace fsmc
