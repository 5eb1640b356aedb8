// fastsmc_b200 — state-split decode kernels for the 159-state decoding quantities (FASTSMC_EXAMPLE).
#include "decode_split.cuh"
#include "split_select.h"

namespace fsmc
{

template <int S, int NW, int RQ, bool ACC, int GRP, int DEPTH, int MINB> static SplitChoice make()
{
  return SplitChoice{decodeSplitKernel<S, NW, RQ, ACC, GRP, DEPTH, MINB>, NW, SplitSmem<S, NW, RQ, GRP, DEPTH>::kTotal, RQ,
                     SplitGeom<S, NW>::Spad, ACC};
}

SplitChoice splitKernel159(const int recordQuads, const bool acc)
{
  switch (recordQuads) {
  case 0:
    return acc ? make<159, 4, 0, true, 1, 2, 3>() : make<159, 4, 0, false, 1, 2, 4>();
  case 1:
    return make<159, 4, 1, false, 4, 2, 4>();
  case 2:
    return make<159, 4, 2, false, 4, 2, 4>();
  default:
    return {};
  }
}

}  // namespace fsmc
