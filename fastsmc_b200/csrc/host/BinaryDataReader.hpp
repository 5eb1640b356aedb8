// fastsmc_b200 host layer — reader of the binary IBD output (.bibd.gz), same interface as the reference
// (ref: ASMC_SRC/SRC/BinaryDataReader.hpp:18-185).  File layout (ref: HMM.cpp:383-401, 1147-1175):
//   header : bool hasLength, bool hasPosteriorMean, bool hasMAP, int chromosome, unsigned nIds,
//            nIds x { unsigned len, famId bytes, unsigned len, iid bytes }
//   record : unsigned ind1, u8 hap1, unsigned ind2, u8 hap2, int bpStart, int bpEnd,
//            [float cM], float score, [float posteriorMean], [float MAP]
#pragma once

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include <zlib.h>

struct IbdPairDataLine {
  std::string ind1FamId = "0_00";
  std::string ind1Id = "0_00";
  int ind1Hap = -1;
  std::string ind2FamId = "0_00";
  std::string ind2Id = "0_00";
  int ind2Hap = -1;
  int chromosome = -1;
  int ibdStart = -1;
  int ibdEnd = -1;
  float lengthInCentimorgans = -1.f;
  float ibdScore = -1.f;
  float postEst = -1.f;
  float mapEst = -1.f;

  /// the text line FastSMC writes for this record: tab separated, floats with 7 significant digits
  std::string toString() const
  {
    auto g7 = [](float v) {
      char b[48];
      std::snprintf(b, sizeof b, "%.7g", static_cast<double>(v));
      return std::string(b);
    };
    std::string s = ind1FamId + '\t' + ind1Id + '\t' + std::to_string(ind1Hap) + '\t' + ind2FamId + '\t' + ind2Id +
                    '\t' + std::to_string(ind2Hap) + '\t' + std::to_string(chromosome) + '\t' +
                    std::to_string(ibdStart) + '\t' + std::to_string(ibdEnd);
    if (lengthInCentimorgans != -1.f) {
      s += '\t' + g7(lengthInCentimorgans);
    }
    s += '\t' + g7(ibdScore);
    if (postEst != -1.f) {
      s += '\t' + g7(postEst);
    }
    if (mapEst != -1.f) {
      s += '\t' + g7(mapEst);
    }
    return s;
  }
};

class BinaryDataReader
{
  gzFile mFile = nullptr;
  bool mHasLength = false, mHasPosterior = false, mHasMap = false;
  int mChromosome = -1;
  std::vector<std::string> mFamIds, mIIds;
  unsigned mNextInd1 = 0;
  bool mMoreLinesInFile = true;

  template <class T> bool get(T& v) { return gzread(mFile, &v, sizeof(T)) == static_cast<int>(sizeof(T)); }
  std::string getString()
  {
    unsigned n = 0;
    get(n);
    std::string s(n, '\0');
    if (n) {
      gzread(mFile, &s[0], n);
    }
    return s;
  }
  void peek() { mMoreLinesInFile = get(mNextInd1); }

public:
  explicit BinaryDataReader(const std::string& binaryFile)
  {
    mFile = gzopen(binaryFile.c_str(), "rb");
    if (!mFile) {
      throw std::runtime_error("ERROR: could not open " + binaryFile);
    }
    unsigned nIds = 0;
    get(mHasLength);
    get(mHasPosterior);
    get(mHasMap);
    get(mChromosome);
    get(nIds);
    for (unsigned i = 0; i < nIds; ++i) {
      mFamIds.push_back(getString());
      mIIds.push_back(getString());
    }
    peek();
  }
  BinaryDataReader(const BinaryDataReader&) = delete;
  BinaryDataReader& operator=(const BinaryDataReader&) = delete;
  ~BinaryDataReader()
  {
    if (mFile) {
      gzclose(mFile);
    }
  }

  IbdPairDataLine getNextLine()
  {
    IbdPairDataLine line;
    const unsigned ind1 = mNextInd1;
    unsigned ind2 = 0;
    std::uint_least8_t hap1 = 0, hap2 = 0;
    get(hap1);
    get(ind2);
    get(hap2);
    get(line.ibdStart);
    get(line.ibdEnd);
    if (mHasLength) {
      get(line.lengthInCentimorgans);
    }
    get(line.ibdScore);
    if (mHasPosterior) {
      get(line.postEst);
    }
    if (mHasMap) {
      get(line.mapEst);
    }
    line.ind1Hap = hap1;
    line.ind2Hap = hap2;
    line.chromosome = mChromosome;
    line.ind1FamId = mFamIds.at(ind1);
    line.ind1Id = mIIds.at(ind1);
    line.ind2FamId = mFamIds.at(ind2);
    line.ind2Id = mIIds.at(ind2);
    peek();
    return line;
  }

  bool moreLinesInFile() const { return mMoreLinesInFile; }
};
