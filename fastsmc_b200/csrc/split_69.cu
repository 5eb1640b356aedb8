// fastsmc_b200 — state-split decode kernels for the 69-state decoding quantities (30-100-2000, UKBB-style).
#include "decode_split.cuh"
#include "split_select.h"

namespace fsmc
{

template <int S, int NW, int RQ, bool ACC, int GRP, int DEPTH, int MINB> static SplitChoice make()
{
  return SplitChoice{decodeSplitKernel<S, NW, RQ, ACC, GRP, DEPTH, MINB>, NW, SplitSmem<S, NW, RQ, GRP, DEPTH>::kTotal, RQ,
                     SplitGeom<S, NW>::Spad, ACC};
}

SplitChoice splitKernel69(const int recordQuads, const bool acc)
{
  switch (recordQuads) {
  case 0:
    return acc ? make<69, 2, 0, true, 1, 2, 8>() : make<69, 2, 0, false, 1, 2, 8>();
  case 1:
    return make<69, 2, 1, false, 4, 2, 8>();
  case 2:
    return make<69, 2, 2, false, 4, 2, 8>();
  default:
    return {};
  }
}

}  // namespace fsmc
