// Test scaffolding — driver for the UNMODIFIED reference sources compiled against oracle/shim (oracle/Makefile, target
// `reference`).  It is the reference's own public API (DecodingParams fields, ASMC::FastSMC::run, ASMC::ASMC::decodePairs)
// behind a key=value command line, because the reference's regression tests set options (useKnownSeed, foldData, ...)
// that its FastSMC_exe command line does not expose (ref: ASMC_SRC/TESTS/test_fastsmc_regression.cpp:34-52).
//
//   ref_driver in=<root> dq=<file> out=<root> [hashing=0|1] [jobs=J jobInd=K] [time=50] [min_m=1.5] [skip=0] [gap=1]
//              [max_seeds=0] [batchSize=32] [noConditionalAgeEstimates=0|1] [bin=0|1] [useKnownSeed=1]
//              [perPairMAP=1] [perPairPosteriorMean=1] [segmentLength=1]
//   ref_driver mode=decodePairs in=... dq=... pairs=<file of "hapA hapB" lines> dump=<file>   (ASMC per-site outputs)
//
// Prints one JSON line: {"construct_s":..., "run_s":...}.  Only tests/ and bench.py's CPU-baseline legs execute it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "ASMC.hpp"
#include "DecodingParams.hpp"
#include "FastSMC.hpp"

namespace
{
double seconds(const std::chrono::steady_clock::time_point a, const std::chrono::steady_clock::time_point b)
{
  return std::chrono::duration<double>(b - a).count();
}
}  // namespace

int main(int argc, char* argv[])
{
  std::map<std::string, std::string> kv;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    const size_t eq = a.find('=');
    if (eq == std::string::npos) {
      std::cerr << "ref_driver: expected key=value, got " << a << std::endl;
      return 2;
    }
    kv[a.substr(0, eq)] = a.substr(eq + 1);
  }
  auto get = [&](const char* k, const std::string& dflt) { return kv.count(k) ? kv[k] : dflt; };
  auto geti = [&](const char* k, const int dflt) { return kv.count(k) ? std::atoi(kv[k].c_str()) : dflt; };
  auto getf = [&](const char* k, const float dflt) { return kv.count(k) ? static_cast<float>(std::atof(kv[k].c_str())) : dflt; };
  const std::string mode = get("mode", "run");

  DecodingParams params;
  params.inFileRoot = get("in", "");
  params.decodingQuantFile = get("dq", params.inFileRoot + ".decodingQuantities.gz");
  params.outFileRoot = get("out", "/tmp/ref_driver_out");
  params.decodingModeString = "array";
  params.foldData = geti("foldData", 1);
  params.usingCSFS = geti("usingCSFS", 1);
  params.batchSize = geti("batchSize", 32);
  params.recallThreshold = 3;
  params.min_m = getf("min_m", 1.5f);
  params.skip = getf("skip", 0.f);
  params.gap = geti("gap", 1);
  params.max_seeds = geti("max_seeds", 0);
  params.jobs = geti("jobs", 1);
  params.jobInd = geti("jobInd", 1);
  params.time = geti("time", 50);
  params.useKnownSeed = geti("useKnownSeed", 1);
  const auto t0 = std::chrono::steady_clock::now();

  if (mode == "run") {
    params.hashing = geti("hashing", 1);
    params.FastSMC = true;
    params.BIN_OUT = geti("bin", 0);
    params.outputIbdSegmentLength = geti("segmentLength", 1);
    params.noConditionalAgeEstimates = geti("noConditionalAgeEstimates", 1);
    params.doPerPairMAP = geti("perPairMAP", 1);
    params.doPerPairPosteriorMean = geti("perPairPosteriorMean", 1);
    if (!params.validateParamsFastSMC()) {
      return 3;
    }
    ASMC::FastSMC fastSMC(params);
    const auto t1 = std::chrono::steady_clock::now();
    fastSMC.run();
    const auto t2 = std::chrono::steady_clock::now();
    std::printf("{\"construct_s\": %.6f, \"run_s\": %.6f}\n", seconds(t0, t1), seconds(t1, t2));
    return 0;
  }
  if (mode == "decodePairs") {
    // per-site posterior mean / MAP (and optionally the full posteriors) of listed haplotype pairs, dumped as raw
    // little-endian arrays: int32 nPairs, int32 sites, int32 states, then means [nPairs][sites] float32,
    // MAPs [nPairs][sites] int32, then (full=1) posteriors [nPairs][states][sites] float32
    params.FastSMC = false;
    params.hashing = false;
    params.batchSize = geti("batchSize", 64);
    params.doPerPairPosteriorMean = true;
    params.doPerPairMAP = true;
    const bool full = geti("full", 0);
    std::vector<unsigned long> a, b;
    {
      std::ifstream in(get("pairs", ""));
      unsigned long x, y;
      while (in >> x >> y) {
        a.push_back(x);
        b.push_back(y);
      }
    }
    ASMC::ASMC asmc(params);
    const auto t1 = std::chrono::steady_clock::now();
    asmc.decodePairs(a, b, full, false, true, true);
    const auto t2 = std::chrono::steady_clock::now();
    const DecodePairsReturnStruct& r = asmc.getRefOfResults();
    std::ofstream out(get("dump", "/tmp/ref_driver_pairs.bin"), std::ios::binary);
    const int n = static_cast<int>(a.size());
    const int sites = static_cast<int>(r.perPairPosteriorMeans.cols());
    const int states = full && n ? static_cast<int>(r.perPairPosteriors.at(0).rows()) : 0;
    out.write(reinterpret_cast<const char*>(&n), 4);
    out.write(reinterpret_cast<const char*>(&sites), 4);
    out.write(reinterpret_cast<const char*>(&states), 4);
    out.write(reinterpret_cast<const char*>(r.perPairPosteriorMeans.data()), sizeof(float) * n * sites);
    out.write(reinterpret_cast<const char*>(r.perPairMAPs.data()), sizeof(int) * n * sites);
    if (full) {
      for (int i = 0; i < n; ++i) {
        out.write(reinterpret_cast<const char*>(r.perPairPosteriors.at(i).data()), sizeof(float) * states * sites);
      }
    }
    std::printf("{\"construct_s\": %.6f, \"run_s\": %.6f}\n", seconds(t0, t1), seconds(t1, t2));
    return 0;
  }
  std::cerr << "ref_driver: unknown mode " << mode << std::endl;
  return 2;
}
