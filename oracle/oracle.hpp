// TEST INFRASTRUCTURE ONLY — CPU oracle for the FastSMC IBD hot path.
//
// This is a from-scratch CPU restatement of the reference algorithm (PalamaraLab/FastSMC,
// ASMC_SRC/SRC).  It exists so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs can check and time the CUDA product against it.
// Nothing under fastsmc_b200/ may include, link or call this code.
//
// Parity pin: the oracle is validated against the reference's own golden outputs
// FILES/FASTSMC_EXAMPLE/regression_output.ibd.gz (G1, hashing) and
// regression_output_no_hashing.ibd.gz (G2) with the parameters of
// ASMC_SRC/TESTS/test_fastsmc_regression.cpp:32-160 (see tests/test_oracle_golden.py), and
// against the known-answer tests in ASMC_SRC/TESTS/test_hmm_utils.cpp / test_hashing.cpp.
//
// All "ref:" citations are paths relative to /root/reference/ASMC_SRC/SRC.
#pragma once

#include <array>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace fo
{

// ref: DecodingParams.hpp:27-123 (only the fields the IBD path reads)
struct Params {
  std::string inFileRoot;
  std::string decodingQuantFile;
  std::string outFileRoot;
  int jobs = 1;
  int jobInd = 1;
  bool foldData = true;
  bool usingCSFS = true;
  float skipCSFSdistance = 0.f;
  int batchSize = 32;
  float skip = 0.f;
  int gap = 1;
  int max_seeds = 0;
  float min_m = 1.f;
  bool hashing = true;
  bool FastSMC = true;
  bool BIN_OUT = false;
  bool useKnownSeed = true;
  bool outputIbdSegmentLength = true;
  int hashingWordSize = 64;
  int constReadAhead = 10;
  int time = 100;
  bool noConditionalAgeEstimates = false;
  bool doPerPairPosteriorMean = false;
  bool doPerPairMAP = false;
  bool withinOnly = false;
  // 0 = this C++ library's std::shuffle, 1 = Lemire written out, 2 = libstdc++ <= 10 divide-and-reject
  int shuffleFlavor = 0;
  // false = arithmetic of the reference's NO_SSE build (the parity target); true = arithmetic of its
  // SSE/AVX builds: RCPPS approximate reciprocal in the combine step (ref: HMM.cpp:704-709) and the
  // SIMD association in the backward step (ref: HMM.cpp:1036-1037).  The golden files came from an AVX build.
  bool simdFlavor = false;
};

// ref: DecodingQuantities.hpp:45-70, DecodingQuantities.cpp:60-346
struct Quantities {
  int states = 0;
  int csfsSamples = 0;
  std::vector<float> initialStateProb, expectedTimes, discretization, timeVector, columnRatios;
  std::vector<std::vector<float>> classicEmission, compressedEmission;  // [2][S]
  // distance-keyed transition rows; key is the exact float parsed from the file
  std::unordered_map<float, std::vector<float>> D, B, U, RR;
  std::vector<std::vector<std::vector<float>>> csfs, foldedCsfs, ascCsfs, foldedAscCsfs;  // [undist][dist][S]
  void load(const std::string& file);
};

// One haplotype pair submitted to the HMM.  ref: HMM.hpp PairObservations, HMM.cpp:129-145
struct PairObs {
  int aHap = 1;        // 1 or 2   (reference iHap)
  unsigned aInd = 0;   // individual index within the job subset (reference iInd)
  int bHap = 1;        // reference jHap
  unsigned bInd = 0;   // reference jInd
};

// A reference batch: the pairs decoded together, their per-slot match ranges and the windows
// derived from them.  ref: HMM.cpp:555-636
struct Batch {
  std::vector<PairObs> pairs;            // actual pairs (no padding)
  unsigned scanFrom = 0, scanTo = 0;     // startBatch / endBatch (segment scan range)
  unsigned from = 0, to = 0;             // decode window [from, to)
};

// One emitted IBD segment record, before formatting.  ref: HMM.cpp:1110-1177
struct Segment {
  uint32_t batch = 0;
  uint32_t lane = 0;
  PairObs obs;
  unsigned posStart = 0, posEnd = 0;
  float prob = 0.f;        // sum of per-site IBD probabilities over the segment
  float postMean = 0.f;    // valid if doPerPairPosteriorMean
  float map = 0.f;         // valid if doPerPairMAP
  int mapState = -1;
};

struct Candidate {
  uint32_t hapA, hapB, from, to;  // arguments of HMM::decodeFromHashing, in call order
};

class Oracle
{
public:
  explicit Oracle(const Params& p, bool asmcMode = false);

  // ---- data (ref: Data.hpp:33-75)
  Params params;
  Quantities dq;
  int sites = 0;
  int sampleSizeTotal = 0;  // diploid samples in the file
  std::vector<std::string> famId, iid;
  std::vector<std::vector<uint8_t>> hap;  // [2*nInd][sites], minor-allele folded if foldData
  std::vector<float> genPos;
  std::vector<int> physPos;
  std::vector<float> recRate;
  std::vector<int> derivedCount, totalCount;
  std::vector<uint8_t> flipped;
  int chrNumber = 0;
  unsigned windowSize = 0, w_i = 0, w_j = 0;
  bool aboveDiag = false;
  std::vector<std::array<int, 3>> undistinguished;

  // ---- model (ref: HMM.cpp:65-127,159-256)
  std::vector<float> e1, e0m1, e2m0;  // [sites][S]
  unsigned stateThreshold = 0, ageThreshold = 0;
  float probabilityThreshold = 0.f;

  // ---- hot path
  // posterior out: [(to-from)][S][nLanes]
  void decodeBatch(const std::vector<PairObs>& pairs, unsigned from, unsigned to, std::vector<float>& posterior) const;
  void callSegments(const Batch& b, uint32_t batchIdx, const std::vector<float>& posterior,
                    std::vector<Segment>& out) const;
  // per-site posterior mean / MAP (ref: HMM.cpp:1360-1410); outputs [nLanes][to-from]
  void perSiteSummary(const std::vector<float>& posterior, unsigned nLanes, unsigned len, float* mean, int* map) const;

  // ---- drivers
  // all-pairs enumeration of HMM::decodeAll (ref: HMM.cpp:283-381)
  std::vector<PairObs> enumerateAllPairs() const;
  // GERMLINE-style seeding (ref: FastSMC.cpp:41-238, HASHING/*) → decodeFromHashing call stream
  std::vector<Candidate> seedCandidates() const;
  // group a pair stream into reference batches
  std::vector<Batch> makeBatches(const std::vector<PairObs>& pairs, const std::vector<Candidate>* cands) const;
  // whole FastSMC::run.  Writes <out>.<jobInd>.<jobs>.FastSMC.{ibd,bibd}.gz unless outPath given.
  // Returns number of records.  Keeps the candidates/batches/segments of the run for inspection.
  long run(const std::string& outPath = "", int threads = 0);
  std::vector<Candidate> lastCandidates;
  std::vector<Batch> lastBatches;
  std::vector<Segment> lastSegments;
  double lastDecodeSeconds = 0.0;   // time inside decodeBatch+callSegments (wall, all threads)
  double lastPairSites = 0.0;

  std::string formatText(const Segment& s) const;  // ref: HMM.cpp:1116-1141

  unsigned hapIndex(const PairObs& o, bool second) const
  {
    return second ? 2u * o.bInd + (o.bHap - 1) : 2u * o.aInd + (o.aHap - 1);
  }
  bool inJob(unsigned sampleLine) const;

private:
  void loadSamples();
  void loadMapAndHapsFastSMC();
  void loadHapsAndMapAsmc();
  void computeUndistinguished();
  void prepareEmissions();
};

// helpers with reference KATs (ref: HmmUtils.cpp:65-94,153-177; HASHING/Utils.cpp:22-34)
float roundMorgans(float value, int precision, float minv);
int roundPhysical(int value, int precision);
unsigned getFromPosition(const std::vector<float>& gen, unsigned from, float cmDist = 0.5f);
unsigned getToPosition(const std::vector<float>& gen, unsigned to, float cmDist = 0.5f);
double cmBetween(int w1, int w2, const std::vector<float>& gen, int wordSize);
float parseFloat(const std::string& s);  // ref: StringUtils.cpp:36-39

}  // namespace fo
