// fastsmc_b200 host layer — input data of one job: samples, haplotypes, genetic map.
// Public field names follow the reference's Data class (ref: ASMC_SRC/SRC/Data.hpp:33-75).  The layout is
// different: haplotypes are kept bit-packed ([haplotype][site/64] uint64, minor-allele folded), which is what
// the GPU kernels read, instead of two std::vector<bool> per Individual.
#pragma once

#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <map>
#include <functional>

#include "DecodingParams.hpp"

// One diploid sample, materialised on demand from the packed matrix (ref: ASMC_SRC/SRC/Individual.hpp).
struct Individual {
  std::vector<bool> genotype1;
  std::vector<bool> genotype2;
  explicit Individual(int numOfSites = 0) : genotype1(numOfSites), genotype2(numOfSites) {}
  void setGenotype(int_least8_t hap, int pos, bool val)
  {
    (hap == 1 ? genotype1 : genotype2).at(pos) = val;
  }
};

class Data
{
public:
  std::vector<std::string> FamIDList = {};
  std::vector<std::string> IIDList = {};
  std::vector<std::string> famAndIndNameList = {};

  unsigned long sampleSize = 0ul;         // diploid samples in the file (all jobs)
  unsigned long haploidSampleSize = 0ul;  // 2 * sampleSize
  int sites = 0;
  bool decodingUsesCSFS = false;
  bool mJobbing = false;
  bool mUseKnownSeed = false;
  bool foldToMinorAlleles = false;
  std::vector<float> geneticPositions = {};
  std::vector<int> physicalPositions = {};
  std::vector<bool> siteWasFlippedDuringFolding = {};
  std::vector<float> recRateAtMarker = {};

  // FastSMC job geometry (ref: Data.cpp:62-80)
  int chrNumber = 0;
  unsigned int windowSize = 0u;
  unsigned int w_i = 0u;
  unsigned int w_j = 0u;
  bool is_j_above_diag = false;
  int jobs = 1, jobInd = 1;

  // ---- B200 layout -------------------------------------------------------------------------------
  // Loaded haplotypes (the job's subset): haplotype h = individual h/2, hap 1 + h%2.
  // bit (s % 64) of hapBits[h * wordsPerHap + s / 64] = folded allele (1 = minor) at site s.
  std::vector<uint64_t> hapBits;
  long wordsPerHap = 0;
  // flipMask[s / 64] bit (s % 64) = site s was flipped by folding; raw allele = folded ^ flip.  The seeding
  // path hashes RAW alleles (ref: FastSMC.cpp:176-186).
  std::vector<uint64_t> flipMask;
  // global haplotype index (2 * line in .samples + hap) of each loaded haplotype (ref: FastSMC.cpp:97-103)
  std::vector<uint32_t> globalHapId;
  // whole-file allele counts per site (ref: Data.cpp:498-503)
  std::vector<int> totalSamplesCount;
  std::vector<int> derivedAlleleCounts;

  Data() = default;
  explicit Data(const DecodingParams& params);

  // In-memory construction for synthetic workloads: `rawAlleles` is [numHaps][sites] bytes (0/1) for ALL samples of
  // the data set; the job subset, folding and counts are derived exactly as when reading files.
  static Data fromArrays(const DecodingParams& params, const std::vector<std::string>& famIds,
                         const std::vector<std::string>& iids, const uint8_t* rawAlleles, long numHaps, int numSites,
                         const std::vector<int>& physPos, const std::vector<double>& cM, int chr);

  /// The Data of job params.jobInd of params.jobs, cut out of `whole` (a Data that loaded every sample, i.e. read
  /// with jobs = jobInd = 1): the same object Data(params) would read from the files, without reading them again.
  static Data forJob(const Data& whole, const DecodingParams& params);
  /// An object that depends on the WHOLE data set only, not on the job's sample subset (e.g. the emission / transition
  /// tables of HMM): built once by `make` and shared by the jobs cut out of one data set (forJob), like the
  /// undistinguished counts.  Builders of the same data set are serialised; `make` may call calculateUndistinguishedCounts.
  std::shared_ptr<const void> sharedObject(const std::string& key, const std::function<std::shared_ptr<const void>()>& make) const;

  /// Packed-matrix cache of a FastSMC-mode data set read with jobs = 1 (DecodingParams::hapBitCache): false when there is
  /// no valid cache for these files and options.
  static bool loadBitCache(const DecodingParams& params, Data& whole);
  void writeBitCache(const DecodingParams& params) const;
  static std::string bitCachePath(const std::string& inFileRoot);

  static int countHapLines(std::string inFileRoot);
  static int countSamplesLines(std::string inFileRoot);

  // Hypergeometric draws through std::rand -> std::mt19937 -> std::shuffle in the reference's call order
  // (ref: Data.cpp:144-160, 567-599).
  std::vector<std::vector<int>> calculateUndistinguishedCounts(int numCsfsSamples) const;

  unsigned long numLoadedIndividuals() const { return famAndIndNameList.size(); }
  unsigned long numLoadedHaplotypes() const { return 2ul * famAndIndNameList.size(); }
  bool allele(unsigned long hap, long site) const
  {
    return (hapBits[hap * wordsPerHap + (site >> 6)] >> (site & 63)) & 1ull;
  }
  Individual individual(unsigned long i) const;
  // whether line `n` of the samples file belongs to this job (ref: Data.cpp:251-262)
  bool readSample(unsigned linesProcessed) const;

private:
  void readFromFiles(const DecodingParams& params);
  void setJobGeometry(const DecodingParams& params);
  void readSamplesList(const std::string& inFileRoot);
  void readHapsFastSMC(const std::string& inFileRoot, const std::vector<std::pair<unsigned long, double>>& geneticMap);
  void readHapsAsmc(const std::string& inFileRoot);
  void readMapAsmc(const std::string& inFileRoot);
  static std::vector<std::pair<unsigned long, double>> readMapFastSMC(const std::string& inFileRoot);
  void allocate();
  void addSite(int pos, const char* alleles /* 2 chars per haplotype: ' ' + '0'/'1' */, unsigned long nHapsInFile,
               bool subset);
  void flushSiteWord();
  void finishSites(int numSites);
  // calculateUndistinguishedCounts depends on the WHOLE file's allele counts, the seed and the CSFS sample count only, not on
  // the job's sample subset: the jobs cut out of one data set (forJob) share the result instead of re-drawing it
  // (3 x sites shuffles of 2N shorts, the dominant per-job host cost at 64 jobs per data set).
  struct UndistinguishedCache {
    std::mutex lock;
    int csfsSamples = -1;
    bool knownSeed = false;
    std::vector<std::vector<int>> counts;
    // further objects that depend on the whole data set only (HMM's model tables): see sharedObject
    std::mutex objectsLock;
    std::map<std::string, std::shared_ptr<const void>> objects;
  };
  mutable std::shared_ptr<UndistinguishedCache> mUndistinguished = std::make_shared<UndistinguishedCache>();
  std::vector<std::vector<int>> drawUndistinguishedCounts(int numCsfsSamples) const;
  std::vector<uint64_t> mWordBuf;    // bits of the current 64-site word, one entry per loaded haplotype
  std::vector<uint64_t> mWordMajor;  // completed words, [word][hap]; transposed into hapBits by finishSites
  long mWordBufIndex = -1;
  void addMarker(int pos, unsigned long bp, const std::vector<std::pair<unsigned long, double>>& gmap, unsigned& cur_g);
};
