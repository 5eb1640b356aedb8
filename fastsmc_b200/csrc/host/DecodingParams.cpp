// fastsmc_b200 host layer — see DecodingParams.hpp.
#include "DecodingParams.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <stdexcept>
#include <vector>

namespace
{

std::string toLower(std::string s)
{
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
  return s;
}

// Resolves decodingModeString/useAncestral into the mode enums (ref: DecodingParams.cpp:327-351, 499-526).
bool resolveMode(DecodingParams& p)
{
  p.decodingModeString = toLower(p.decodingModeString);
  if (p.decodingModeString == "sequence") {
    p.decodingModeOverall = DecodingModeOverall::sequence;
    p.decodingSequence = true;
    p.decodingMode = p.useAncestral ? DecodingMode::sequence : DecodingMode::sequenceFolded;
  } else if (p.decodingModeString == "array") {
    p.decodingModeOverall = DecodingModeOverall::array;
    p.decodingSequence = false;
    p.decodingMode = p.useAncestral ? DecodingMode::array : DecodingMode::arrayFolded;
  } else {
    return false;
  }
  p.foldData = !p.useAncestral;
  return true;
}

// Minimal "--name value" / "--flag" parser with the option table passed in.
struct OptionTable {
  std::map<std::string, std::string*> strings;
  std::map<std::string, int*> ints;
  std::map<std::string, float*> floats;
  std::map<std::string, bool*> switches;  // bool_switch: presence sets true
};

bool parseArgs(const int argc, char* argv[], OptionTable& t, const std::vector<std::string>& required)
{
  std::map<std::string, bool> seen;
  std::vector<std::string> bad;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.rfind("--", 0) != 0) {
      bad.push_back(a);
      continue;
    }
    std::string name = a.substr(2), value;
    bool hasValue = false;
    const size_t eq = name.find('=');
    if (eq != std::string::npos) {
      value = name.substr(eq + 1);
      name = name.substr(0, eq);
      hasValue = true;
    }
    if (t.switches.count(name)) {
      *t.switches[name] = true;
      seen[name] = true;
      continue;
    }
    const bool known = t.strings.count(name) || t.ints.count(name) || t.floats.count(name);
    if (!known) {
      std::cerr << "ERROR: unrecognised option '--" << name << "'" << std::endl;
      return false;
    }
    if (!hasValue) {
      if (i + 1 >= argc) {
        std::cerr << "ERROR: the required argument for option '--" << name << "' is missing" << std::endl;
        return false;
      }
      value = argv[++i];
    }
    try {
      if (t.strings.count(name)) {
        *t.strings[name] = value;
      } else if (t.ints.count(name)) {
        *t.ints[name] = std::stoi(value);
      } else {
        *t.floats[name] = std::stof(value);
      }
    } catch (const std::exception&) {
      std::cerr << "ERROR: the argument ('" << value << "') for option '--" << name << "' is invalid" << std::endl;
      return false;
    }
    seen[name] = true;
  }
  if (!bad.empty()) {
    std::cerr << "ERROR: Unknown options:";
    for (const auto& b : bad) {
      std::cerr << " " << b;
    }
    std::cerr << std::endl;
    return false;
  }
  for (const auto& r : required) {
    if (!seen.count(r)) {
      std::cerr << "ERROR: the option '--" << r << "' is required but missing" << std::endl;
      return false;
    }
  }
  return true;
}

}  // namespace

DecodingParams::DecodingParams() : outFileRoot(inFileRoot), usingCSFS(true), skipCSFSdistance(0.f) {}

DecodingParams::DecodingParams(std::string _inFileRoot, std::string _decodingQuantFile, std::string _outFileRoot,
                               int _jobs, int _jobInd, std::string _decodingModeString, bool _decodingSequence,
                               bool _usingCSFS, bool _compress, bool _useAncestral, float _skipCSFSdistance,
                               bool _noBatches, bool _doPosteriorSums, bool _doPerPairPosteriorMean,
                               std::string _expectedCoalTimesFile, bool _withinOnly, bool _doMajorMinorPosteriorSums,
                               bool _doPerPairMAP)
    : inFileRoot(std::move(_inFileRoot)), decodingQuantFile(std::move(_decodingQuantFile)),
      outFileRoot(std::move(_outFileRoot)), jobs(_jobs), jobInd(_jobInd),
      decodingModeString(std::move(_decodingModeString)), decodingSequence(_decodingSequence), usingCSFS(_usingCSFS),
      compress(_compress), useAncestral(_useAncestral), skipCSFSdistance(_skipCSFSdistance), noBatches(_noBatches),
      doPosteriorSums(_doPosteriorSums), doPerPairPosteriorMean(_doPerPairPosteriorMean), doPerPairMAP(_doPerPairMAP),
      expectedCoalTimesFile(std::move(_expectedCoalTimesFile)), withinOnly(_withinOnly),
      doMajorMinorPosteriorSums(_doMajorMinorPosteriorSums)
{
  if (!processOptions()) {
    throw std::exception();  // as the reference does (ref: DecodingParams.cpp:51-53)
  }
}

DecodingParams::DecodingParams(std::string _inFileRoot, std::string _decodingQuantFile, std::string _outFileRoot,
                               bool _fastSMC)
    : inFileRoot(std::move(_inFileRoot)), decodingQuantFile(std::move(_decodingQuantFile)),
      outFileRoot(std::move(_outFileRoot)), foldData(true), usingCSFS(true), batchSize(32), min_m(1.5f), hashing(true),
      FastSMC(_fastSMC), outputIbdSegmentLength(true), time(50), noConditionalAgeEstimates(true),
      doPerPairPosteriorMean(true), doPerPairMAP(true)
{
  if (!FastSMC) {
    std::cerr << "This DecodingParams constructor sets sensible FastSMC defaults, and is only intended for use with"
                 "FastSMC. Please set the fastSMC parameter to true, or use a different constructor."
              << std::endl;
    exit(1);
  }
  validateParamsFastSMC();
}

// ref: DecodingParams.cpp:78-160 — ASMC_exe options
bool DecodingParams::processCommandLineArgs(int argc, char* argv[])
{
  jobs = 0;
  jobInd = 0;
  skipCSFSdistance = 0.f;
  OptionTable t;
  t.strings = {{"inFileRoot", &inFileRoot}, {"decodingQuantFile", &decodingQuantFile}, {"outFileRoot", &outFileRoot},
               {"mode", &decodingModeString}};
  t.ints = {{"jobs", &jobs}, {"jobInd", &jobInd}};
  t.floats = {{"skipCSFSdistance", &skipCSFSdistance}};
  t.switches = {{"compress", &compress}, {"useAncestral", &useAncestral},
                {"majorMinorPosteriorSums", &doMajorMinorPosteriorSums}, {"posteriorSums", &doPosteriorSums}};
  if (!parseArgs(argc, argv, t, {"inFileRoot"})) {
    return false;
  }
  if (compress && skipCSFSdistance == 0.f) {
    skipCSFSdistance = std::numeric_limits<float>::quiet_NaN();
  }
  return processOptions();
}

// ref: DecodingParams.cpp:164-276 — FastSMC_exe options.  The reference declares hashing, segmentLength,
// perPairMAP and perPairPosteriorMeans as bool_switch with default true, so they are always on from the CLI.
bool DecodingParams::processCommandLineArgsFastSMC(int argc, char* argv[])
{
  fastSmcInvokedWithProgramOptions = true;
  FastSMC = true;
  std::string modeShadow = "array";  // the reference parses --mode into a shadowing local
  time = 100;
  jobs = 1;
  jobInd = 1;
  batchSize = 32;
  recallThreshold = 3;
  outputIbdSegmentLength = true;
  doPerPairMAP = true;
  doPerPairPosteriorMean = true;
  hashing = true;
  min_m = 1.0f;
  skipCSFSdistance = std::numeric_limits<float>::quiet_NaN();
  OptionTable t;
  t.strings = {{"inFileRoot", &inFileRoot}, {"outFileRoot", &outFileRoot}, {"decodingQuantFile", &decodingQuantFile},
               {"mode", &modeShadow}};
  t.ints = {{"time", &time}, {"jobs", &jobs}, {"jobInd", &jobInd}, {"batchSize", &batchSize},
            {"recall", &recallThreshold}, {"gap", &gap}, {"max_seeds", &max_seeds}, {"device", &device},
            {"outputCompressionLevel", &outputCompressionLevel}, {"outputThreads", &outputThreads}};
  t.floats = {{"skipCSFSdistance", &skipCSFSdistance}, {"min_m", &min_m}, {"skip", &skip}, {"min_maf", &min_maf}};
  t.switches = {{"bin", &BIN_OUT}, {"segmentLength", &outputIbdSegmentLength}, {"perPairMAP", &doPerPairMAP},
                {"perPairPosteriorMeans", &doPerPairPosteriorMean},
                {"noConditionalAgeEstimates", &noConditionalAgeEstimates}, {"withinOnly", &withinOnly},
                {"useAncestral", &useAncestral}, {"compress", &compress}, {"noBatches", &noBatches},
                {"hashing", &hashing}, {"exactArithmetic", &exactArithmetic}};
  if (!parseArgs(argc, argv, t, {"inFileRoot", "outFileRoot"})) {
    return false;
  }
  return validateParamsFastSMC();
}

// ref: DecodingParams.cpp:278-464
bool DecodingParams::validateParamsFastSMC()
{
  const std::string del = fastSmcInvokedWithProgramOptions ? "--" : "";
  auto die = [](const std::string& msg) {
    std::cerr << msg << std::endl;
    exit(1);
  };
  if (!FastSMC) {
    die("Attempting to validate FastSMC parameters but FastSMC flag is false. Set DecodingParams::FastSMC to true?");
  }
  if (hashing) {
    if (withinOnly) {
      die(del + "hashing & " + del + "withinOnly cannot be used together. Please remove one of the two flags.");
    }
    if (time <= 0) {
      die(del + "time must be a positive integer.");
    }
  }
  if (batchSize == 0 || batchSize % 8 != 0) {
    die(del + "batchSize must be strictly positive and a multiple of 8.");
  }
  if (compress) {
    if (useAncestral) {
      die(del + "compress & " + del + "useAncestral cannot be used together. A compressed emission cannot use" +
          " ancestral allele information.");
    }
    if (!std::isnan(skipCSFSdistance)) {
      die(del + "compress & " + del + "skipCSFSdistance cannot be used together. " + del + "compress is a" +
          " shorthand for " + del + "skipCSFSdistance Infinity.");
    }
    skipCSFSdistance = std::numeric_limits<float>::infinity();
  } else if (std::isnan(skipCSFSdistance)) {
    skipCSFSdistance = 0.f;
  }
  if (skipCSFSdistance != std::numeric_limits<float>::infinity()) {
    usingCSFS = true;
  }
  if (!resolveMode(*this)) {
    die("ERROR. Unknown decoding mode: " + decodingModeString);
  }
  if (decodingQuantFile.empty()) {
    std::cout << "Setting " << del << "decodingQuantFile to " << del << "inFileRoot + .decodingQuantities.bin"
              << std::endl;
    decodingQuantFile = inFileRoot + ".decodingQuantities.bin";
  }
  if ((jobs == 0) != (jobInd == 0)) {
    die("ERROR: " + del + "jobs and " + del + "jobInd must either both be set or both be unset");
  }
  if (jobs == 0) {
    jobs = 1;
    jobInd = 1;
  }
  if (jobInd <= 0 || jobInd > jobs || jobs <= 0) {
    die("ERROR: " + del + "jobInd must be between 1 and " + del + "jobs inclusive");
  }
  // jobs must be a perfect square 1, 4, 9, ... (ref: DecodingParams.cpp:376-395)
  {
    int odd = 1, square = 1, prev = 1;
    bool ok = false;
    for (int i = 0; i < 200; ++i) {
      if (square == jobs) {
        ok = true;
        break;
      }
      if (square > jobs) {
        break;
      }
      odd += 2;
      prev = square;
      square += odd;
    }
    if (!ok) {
      die("ERROR: jobs value is incorrect. You should use either " + std::to_string(prev) + " or " +
          std::to_string(square));
    }
  }
  if (recallThreshold < 0 || recallThreshold > 3) {
    die("ERROR: " + del + "recall must be between 0 and 3. ");
  }
  if (outFileRoot.empty()) {
    outFileRoot = inFileRoot;
    if (jobs > 0) {
      outFileRoot += "." + std::to_string(jobInd) + "-" + std::to_string(jobs);
    }
  }
  if (verbose) {
    std::cout << std::boolalpha << "\n---------------------------\n        ASMC OPTIONS       \n---------------------------\n"
              << "Input will have prefix : " << inFileRoot << "\nDecoding quantities file : " << decodingQuantFile
              << "\nOutput will have prefix : " << outFileRoot << "." << jobInd << "." << jobs
              << (hashing ? ".FastSMC" : ".asmc") << (BIN_OUT ? ".bibd" : ".ibd.gz") << "\nBinary output ? " << BIN_OUT
              << "\nTime threshold to define IBD in generations : " << time << "\nUse batches ? " << !noBatches << "\n";
    if (!noBatches) {
      std::cout << "Batch size : " << batchSize << "\n";
    }
    std::cout << "Running job " << jobInd << " of " << jobs << "\nRecall level " << recallThreshold
              << "\nskipCSFSdistance is " << skipCSFSdistance << "\ncompress ? " << compress << "\nuseAncestral ? "
              << useAncestral << "\noutputIbdSegmentLength ? " << outputIbdSegmentLength
              << "\ndoPerPairPosteriorMean ? " << doPerPairPosteriorMean << "\ndoPerPairMAP ? " << doPerPairMAP
              << "\nnoConditionalAgeEstimates ? " << noConditionalAgeEstimates
              << "\nUse hashing as a preprocessing step ? " << hashing << "\n";
    if (hashing) {
      std::cout << "\n---------------------------\n      hashing OPTIONS     \n---------------------------\n"
                << "Minimum match length (in cM) : " << min_m << "\nSkipping words with (seeds/samples) less than "
                << skip << "\nMinimum minor allele frequency : " << min_maf << "\nAllowed gaps " << gap
                << "\nDynamic hash seed cutoff : " << max_seeds << "\n";
    }
    std::cout << std::noboolalpha << std::flush;
  }
  return true;
}

// ref: DecodingParams.cpp:466-558
bool DecodingParams::processOptions()
{
  if (compress) {
    if (useAncestral) {
      std::cerr << "--compress & --useAncestral cannot be used together. A compressed emission cannot use ancestral "
                   "allele information."
                << std::endl;
      exit(1);
    }
    if (!std::isnan(skipCSFSdistance)) {
      std::cerr << "--compress & --skipCSFSdistance cannot be used together. --compress is a shorthand for "
                   "--skipCSFSdistance Infinity."
                << std::endl;
      exit(1);
    }
    skipCSFSdistance = std::numeric_limits<float>::infinity();
  } else if (std::isnan(skipCSFSdistance)) {
    skipCSFSdistance = 0.f;
  }
  if (!expectedCoalTimesFile.empty()) {
    doPerPairPosteriorMean = true;
  }
  if (skipCSFSdistance != std::numeric_limits<float>::infinity()) {
    usingCSFS = true;
  }
  if (!resolveMode(*this)) {
    std::cerr << "Decoding mode should be one of {sequence, array}";
    return false;
  }
  if (decodingQuantFile.empty()) {
    std::cout << "Setting --decodingQuantFile to --inFileRoot + .decodingQuantities.bin" << std::endl;
    decodingQuantFile = inFileRoot + ".decodingQuantities.bin";
  }
  if ((jobs == 0) != (jobInd == 0)) {
    std::cerr << "ERROR: --jobs and --jobInd must either both be set or both be unset" << std::endl;
    return false;
  }
  if (jobs == 0) {
    jobs = 1;
    jobInd = 1;
  }
  if (jobInd <= 0 || jobInd > jobs) {
    std::cerr << "ERROR: --jobInd must be between 1 and --jobs inclusive" << std::endl;
    return false;
  }
  if (outFileRoot.empty()) {
    outFileRoot = inFileRoot;
    if (jobs > 0) {
      outFileRoot += "." + std::to_string(jobInd) + "-" + std::to_string(jobs);
    }
  }
  return true;
}
