// fastsmc_b200 host layer — decoding options.
// Same field names, defaults, constructors and validation rules as the reference's DecodingParams
// (ref: ASMC_SRC/SRC/DecodingParams.hpp:27-123, DecodingParams.cpp:31-558); the command-line parser is
// hand-rolled (no boost::program_options) but accepts the reference's option names and defaults,
// including the fact that its bool_switch options default to *on* (SURVEY F8).
#pragma once

#include <string>

enum class DecodingMode { sequenceFolded, arrayFolded, sequence, array };
enum class DecodingModeOverall { sequence, array };

class DecodingParams
{
  bool fastSmcInvokedWithProgramOptions = false;

public:
  std::string inFileRoot;
  std::string decodingQuantFile;
  std::string outFileRoot;
  int jobs = 1;
  int jobInd = 1;
  std::string decodingModeString = "array";
  DecodingModeOverall decodingModeOverall = DecodingModeOverall::array;
  DecodingMode decodingMode = DecodingMode::arrayFolded;
  bool decodingSequence = false;
  bool foldData = false;
  bool usingCSFS = false;
  bool compress = false;
  bool useAncestral = false;
  float skipCSFSdistance{};
  bool noBatches = false;

  // FastSMC
  int batchSize = 64;
  int recallThreshold = 3;
  float skip = 0.f;
  int gap = 1;
  int max_seeds = 0;
  float min_maf = 0;
  float min_m = 1;
  bool hashing = false;
  bool FastSMC = false;
  bool BIN_OUT = false;
  bool useKnownSeed = false;
  bool outputIbdSegmentLength = false;
  int hashingWordSize = 64;
  int constReadAhead = 10;
  bool haploid = true;
  int time = 100;

  // tasks
  bool noConditionalAgeEstimates = false;
  bool doPosteriorSums = false;
  bool doPerPairPosteriorMean = false;
  bool doPerPairMAP = false;
  std::string expectedCoalTimesFile;
  bool withinOnly = false;
  bool doMajorMinorPosteriorSums = false;

  // B200 build only: CUDA device this job runs on, and whether the kernels must reproduce the
  // reference's unfused NO_SSE arithmetic bit for bit (FSMC_EXACT) instead of using FMA.
  int device = 0;
  bool exactArithmetic = false;
  // B200 build only: emit hashing candidates in the reference's boost::unordered_map iteration
  // order (needed for record-for-record identical output).  false = canonical order
  // (flush word, then pair key), which is cheaper at scale.
  bool referenceCandidateOrder = true;
  /// Input codec: keep the packed haplotype matrix of the whole data set next to the haps file
  /// (<haps file>.fsmcbits) and load that instead of inflating and parsing the text again, as long as the .samples /
  /// .map / haps files have not changed.  Written by the first FastSMC-mode read that has it switched on; also switched
  /// on by the environment variable FSMC_HAP_CACHE=1.
  bool hapBitCache = false;
  // B200 build only: the IBD file is written as concatenated gzip members compressed on `outputThreads` host
  // threads (0 = all cores) at this zlib level; gunzip yields the reference's byte stream at any level.
  int outputCompressionLevel = 1;
  int outputThreads = 0;

  bool processOptions();
  bool processCommandLineArgs(int argc, char* argv[]);
  bool processCommandLineArgsFastSMC(int argc, char* argv[]);
  bool validateParamsFastSMC();

  DecodingParams();
  explicit DecodingParams(std::string _inFileRoot, std::string _decodingQuantFile = "", std::string _outFileRoot = "",
                          int _jobs = 1, int _jobInd = 1, std::string _decodingModeString = "array",
                          bool _decodingSequence = false, bool _usingCSFS = true, bool _compress = false,
                          bool _useAncestral = false, float _skipCSFSdistance = 0.f, bool _noBatches = false,
                          bool _doPosteriorSums = false, bool _doPerPairPosteriorMean = false,
                          std::string _expectedCoalTimesFile = "", bool _withinOnly = false,
                          bool _doMajorMinorPosteriorSums = false, bool _doPerPairMAP = false);
  // FastSMC defaults (ref: DecodingParams.cpp:58-76)
  DecodingParams(std::string _inFileRoot, std::string _decodingQuantFile, std::string _outFileRoot, bool _fastSMC);

  // When true (the default, as in the reference) validation prints the options banner to stdout.
  bool verbose = true;
};
