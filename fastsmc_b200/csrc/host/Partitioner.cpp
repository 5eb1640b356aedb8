// fastsmc_b200 host layer — see Partitioner.hpp.
#include "Partitioner.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "FastSMC.hpp"

namespace ASMC
{

std::vector<int> jobOrder(const int jobs)
{
  // job id -> (row w_i, position in row r): r even = below-diagonal full square, r odd = triangle (ref: Data.cpp:69-79)
  std::vector<int> full, half;
  int w = 1, inRow = 1, upTo = 1;
  for (int j = 1; j <= jobs; ++j) {
    while (upTo < j) {
      ++w;
      inRow += 2;
      upTo += inRow;
    }
    const int r = inRow - (upTo - j);
    const int wj = static_cast<int>(std::ceil(static_cast<float>(r) / 2));
    // the last job also takes the remainder samples, so it goes first
    (wj == w && j != jobs ? half : full).push_back(j);
  }
  std::reverse(full.begin(), full.end());
  full.insert(full.end(), half.begin(), half.end());
  return full;
}

std::vector<int> jobsOfRank(const int jobs, const int world, const int rank)
{
  const std::vector<int> order = jobOrder(jobs);
  std::vector<int> mine;
  for (size_t i = static_cast<size_t>(rank); i < order.size(); i += static_cast<size_t>(world)) {
    mine.push_back(order[i]);
  }
  return mine;
}

std::vector<JobReport> runAllJobs(const DecodingParams& params, const std::vector<int>& devices)
{
  if (devices.empty()) {
    throw std::runtime_error("runAllJobs: no devices given");
  }
  const std::vector<int> order = jobOrder(params.jobs);
  std::vector<JobReport> reports(static_cast<size_t>(params.jobs));
  // the files are read once; every job cuts its two sample windows out of the result (Data::forJob)
  DecodingParams wholeParams = params;
  wholeParams.jobs = 1;
  wholeParams.jobInd = 1;
  const Data whole(wholeParams);
  std::atomic<size_t> next{0};
  auto worker = [&](const int device) {
    for (size_t i = next++; i < order.size(); i = next++) {
      JobReport& rep = reports[static_cast<size_t>(order[i] - 1)];
      rep.jobInd = order[i];
      rep.device = device;
      const auto t0 = std::chrono::steady_clock::now();
      try {
        DecodingParams p = params;
        p.jobInd = order[i];
        p.device = device;
        p.verbose = false;
        // reading, model preparation and decoding of different jobs overlap freely: the only process-wide state, the
        // std::rand sequence of the emission preparation, is seeded and consumed under its own lock (Data.cpp)
        auto job = std::make_unique<FastSMC>(p, Data::forJob(whole, p));
        job->run();
        const HMM::RunStats& st = job->hmm().getRunStats();
        rep.candidates = job->getSeedingStats().candidates;
        rep.pairsDecoded = st.pairsDecoded;
        rep.segments = st.segments;
        rep.pairSites = st.pairSites;
        rep.kernelMs = st.kernelMs;
        rep.seedMs = job->getSeedingStats().device.kernelMs;
      } catch (const std::exception& e) {
        rep.error = e.what();
      }
      rep.wallSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
  };
  std::vector<std::thread> pool;
  for (size_t d = 1; d < devices.size(); ++d) {
    pool.emplace_back(worker, devices[d]);
  }
  worker(devices[0]);
  for (auto& t : pool) {
    t.join();
  }
  return reports;
}

}  // namespace ASMC
