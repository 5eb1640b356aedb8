// fastsmc_b200 host layer — result container of ASMC::decodePairs, same member names as the reference's
// struct (ref: ASMC_SRC/SRC/DecodePairsReturnStruct.hpp:20-127) with plain row-major matrices instead of Eigen
// arrays (Eigen is not a dependency of this build).
#pragma once

#include <cstddef>
#include <string>
#include <tuple>
#include <vector>

template <class T> struct RowMajorMatrix {
  long nRows = 0, nCols = 0;
  std::vector<T> values;
  void resize(long r, long c)
  {
    nRows = r;
    nCols = c;
    values.assign(static_cast<size_t>(r) * c, T{});
  }
  void setZero() { std::fill(values.begin(), values.end(), T{}); }
  long rows() const { return nRows; }
  long cols() const { return nCols; }
  T& operator()(long r, long c) { return values[static_cast<size_t>(r) * nCols + c]; }
  const T& operator()(long r, long c) const { return values[static_cast<size_t>(r) * nCols + c]; }
  T* row(long r) { return values.data() + static_cast<size_t>(r) * nCols; }
  const T* row(long r) const { return values.data() + static_cast<size_t>(r) * nCols; }
};

struct DecodePairsReturnStruct {
private:
  bool m_storeFullPosteriors = false;
  bool m_storeSumOfPosteriors = false;
  bool m_storePerPairPosteriors = false;
  bool m_storePerPairMAPs = false;
  std::size_t numWritten = 0ul;

public:
  void initialise(const std::vector<unsigned long>& individualsA, const std::vector<unsigned long>& individualsB,
                  long int numSites, long int numStates, bool _fullPosteriors = false, bool _sumOfPosteriors = false,
                  bool _perPairPosteriors = false, bool _perPairMAPs = false)
  {
    (void)individualsB;
    numWritten = 0ul;
    const long n = static_cast<long>(individualsA.size());
    m_storeFullPosteriors = _fullPosteriors;
    m_storeSumOfPosteriors = _sumOfPosteriors;
    m_storePerPairPosteriors = _perPairPosteriors;
    m_storePerPairMAPs = _perPairMAPs;
    perPairIndices.resize(n);
    perPairPosteriors.clear();
    if (m_storeFullPosteriors) {
      perPairPosteriors.resize(n);
      for (auto& m : perPairPosteriors) {
        m.resize(numStates, numSites);
      }
    }
    if (m_storeSumOfPosteriors) {
      sumOfPosteriors.resize(numStates, numSites);
    }
    if (m_storePerPairPosteriors) {
      perPairPosteriorMeans.resize(n, numSites);
      minPosteriorMeans.assign(numSites, 0.f);
      argminPosteriorMeans.assign(numSites, 0);
    }
    if (m_storePerPairMAPs) {
      perPairMAPs.resize(n, numSites);
      minMAPs.assign(numSites, 0);
      argminMAPs.assign(numSites, 0);
    }
  }

  /// (hap index A, "id#hap" A, hap index B, "id#hap" B) per decoded pair
  std::vector<std::tuple<unsigned long, std::string, unsigned long, std::string>> perPairIndices;
  /// per pair: states x sites, posterior weighted by the expected time of the state (as the reference stores it,
  /// ref: HMM.cpp:1382-1386)
  std::vector<RowMajorMatrix<float>> perPairPosteriors;
  RowMajorMatrix<float> sumOfPosteriors;        // states x sites
  RowMajorMatrix<float> perPairPosteriorMeans;  // pairs x sites
  std::vector<float> minPosteriorMeans;
  std::vector<int> argminPosteriorMeans;
  RowMajorMatrix<int> perPairMAPs;  // pairs x sites
  std::vector<int> minMAPs;
  std::vector<int> argminMAPs;

  void incrementNumWritten() { numWritten += 1; }
  std::size_t getNumWritten() const { return numWritten; }

  /// column-wise min / argmin over pairs (ref: DecodePairsReturnStruct.hpp:105-118); first minimum wins
  void finaliseCalculations()
  {
    for (long s = 0; s < perPairPosteriorMeans.cols(); ++s) {
      long arg = 0;
      float best = perPairPosteriorMeans(0, s);
      for (long r = 1; r < perPairPosteriorMeans.rows(); ++r) {
        if (perPairPosteriorMeans(r, s) < best) {
          best = perPairPosteriorMeans(r, s);
          arg = r;
        }
      }
      minPosteriorMeans[s] = best;
      argminPosteriorMeans[s] = static_cast<int>(arg);
    }
    for (long s = 0; s < perPairMAPs.cols(); ++s) {
      long arg = 0;
      int best = perPairMAPs(0, s);
      for (long r = 1; r < perPairMAPs.rows(); ++r) {
        if (perPairMAPs(r, s) < best) {
          best = perPairMAPs(r, s);
          arg = r;
        }
      }
      minMAPs[s] = best;
      argminMAPs[s] = static_cast<int>(arg);
    }
  }
};
