// fastsmc_b200 host layer — see ASMC.hpp.
#include "ASMC.hpp"

#include <iostream>
#include <stdexcept>

#include "HmmUtils.hpp"

ASMC::ASMC::ASMC(DecodingParams params) : mParams{std::move(params)}, mData{mParams}, mHmm{mData, mParams} {}

// ref: ASMC.cpp:28-50 — array mode, CSFS on, posterior sums + per-pair means + MAP requested
ASMC::ASMC::ASMC(const std::string& inFileRoot, const std::string& decodingQuantFile, const std::string& outFileRoot)
    : mParams{inFileRoot, decodingQuantFile, outFileRoot.empty() ? inFileRoot : outFileRoot,
              1,          1,                 "array",
              false,      true,              false,
              false,      0.f,               false,
              true,       false,             "",
              false,      true,              true},
      mData{mParams}, mHmm{mData, mParams}
{
}

DecodingReturnValues ASMC::ASMC::decodeAllInJob()
{
  std::cout << "Decoding job " << mParams.jobInd << " of " << mParams.jobs << "\n\n";
  mHmm.decodeAll(mParams.jobs, mParams.jobInd);
  return mHmm.getDecodingReturnValues();
}

// ref: ASMC.cpp:80-100
void ASMC::ASMC::decodePairs(const std::vector<unsigned long>& hapIndicesA, const std::vector<unsigned long>& hapIndicesB,
                             bool perPairPosteriors, bool sumOfPosteriors, bool perPairPosteriorMeans, bool perPairMAPs)
{
  if (hapIndicesA.empty() || hapIndicesA.size() != hapIndicesB.size()) {
    throw std::runtime_error("Vector of A indices (" + std::to_string(hapIndicesA.size()) +
                             ") must be the same size as vector of B indices (" + std::to_string(hapIndicesB.size()) +
                             ").\n");
  }
  mHmm.getDecodePairsReturnStruct().initialise(hapIndicesA, hapIndicesB, mData.sites,
                                               mHmm.getDecodingQuantities().states, perPairPosteriors, sumOfPosteriors,
                                               perPairPosteriorMeans, perPairMAPs);
  mHmm.setStorePerPairPosteriorMean(perPairPosteriorMeans);
  mHmm.setStorePerPairMap(perPairMAPs);
  mHmm.setStorePerPairPosterior(perPairPosteriors);
  mHmm.setStoreSumOfPosterior(sumOfPosteriors);
  mHmm.decodeHapPairs(hapIndicesA, hapIndicesB);
  mHmm.finishDecoding();
  mHmm.getDecodePairsReturnStruct().finaliseCalculations();
}

// ref: ASMC.cpp:102-128 — ids of the form "<IID>#1" / "<IID>#2"
void ASMC::ASMC::decodePairs(const std::vector<std::string>& hapIdsA, const std::vector<std::string>& hapIdsB,
                             bool perPairPosteriors, bool sumOfPosteriors, bool perPairPosteriorMeans, bool perPairMAPs)
{
  if (hapIdsA.size() != hapIdsB.size()) {
    throw std::runtime_error("Vector of A IDs (" + std::to_string(hapIdsA.size()) +
                             ") must be the same size as vector of B IDs (" + std::to_string(hapIdsB.size()) + ").\n");
  }
  std::vector<unsigned long> a(hapIdsA.size()), b(hapIdsB.size());
  for (size_t i = 0; i < hapIdsA.size(); ++i) {
    const auto [idA, hapA] = asmc::combinedIdToIndPlusHap(hapIdsA[i]);
    const auto [idB, hapB] = asmc::combinedIdToIndPlusHap(hapIdsB[i]);
    a[i] = asmc::dipToHapId(asmc::getIndIdxFromIdString(mData.IIDList, idA), hapA);
    b[i] = asmc::dipToHapId(asmc::getIndIdxFromIdString(mData.IIDList, idB), hapB);
  }
  decodePairs(a, b, perPairPosteriors, sumOfPosteriors, perPairPosteriorMeans, perPairMAPs);
}

DecodePairsReturnStruct ASMC::ASMC::getCopyOfResults()
{
  return mHmm.getDecodePairsReturnStruct();
}

const DecodePairsReturnStruct& ASMC::ASMC::getRefOfResults()
{
  return mHmm.getDecodePairsReturnStruct();
}
