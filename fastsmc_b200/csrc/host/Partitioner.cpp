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

std::vector<JobReport> runJobs(const DecodingParams& params, const Data& whole, const std::function<int()>& nextJob,
                               const std::vector<int>& devices)
{
  if (devices.empty()) {
    throw std::runtime_error("runJobs: no devices given");
  }
  std::vector<JobReport> reports;
  std::mutex lock;  // the job source and the report list
  auto worker = [&](const int device) {
    for (;;) {
      int jobInd;
      {
        std::lock_guard<std::mutex> g(lock);
        jobInd = nextJob();
      }
      if (jobInd < 1 || jobInd > params.jobs) {
        return;
      }
      JobReport rep;
      rep.jobInd = jobInd;
      rep.device = device;
      const auto t0 = std::chrono::steady_clock::now();
      try {
        DecodingParams p = params;
        p.jobInd = jobInd;
        p.device = device;
        p.verbose = false;
        // reading, model preparation and decoding of different jobs overlap freely: the only process-wide state, the
        // std::rand sequence of the emission preparation, is seeded and consumed under its own lock (Data.cpp)
        Data cut = Data::forJob(whole, p);
        rep.cutSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        auto job = std::make_unique<FastSMC>(p, std::move(cut));
        rep.prepareSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        rep.tablesSeconds = job->hmm().getRunStats().tablesWallS;
        rep.uploadSeconds = job->hmm().getRunStats().uploadWallS;
        job->run();
        const HMM::RunStats& st = job->hmm().getRunStats();
        rep.candidates = job->getSeedingStats().candidates;
        rep.pairsDecoded = st.pairsDecoded;
        rep.segments = st.segments;
        rep.pairSites = st.pairSites;
        rep.kernelMs = st.kernelMs;
        rep.seedMs = job->getSeedingStats().device.kernelMs;
        rep.seedSeconds = job->getSeedingStats().seedWallS;
        rep.orderSeconds = job->getSeedingStats().orderWallS;
        rep.decodeSeconds = st.decodeWallS;
        rep.outputSeconds = st.outputWallS;
      } catch (const std::exception& e) {
        rep.error = e.what();
      }
      rep.wallSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      std::lock_guard<std::mutex> g(lock);
      reports.push_back(rep);
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
  std::sort(reports.begin(), reports.end(), [](const JobReport& a, const JobReport& b) { return a.jobInd < b.jobInd; });
  return reports;
}

std::vector<JobReport> runAllJobs(const DecodingParams& params, const std::vector<int>& devices)
{
  if (devices.empty()) {
    throw std::runtime_error("runAllJobs: no devices given");
  }
  const std::vector<int> order = jobOrder(params.jobs);
  // the files are read once; every job cuts its two sample windows out of the result (Data::forJob)
  DecodingParams wholeParams = params;
  wholeParams.jobs = 1;
  wholeParams.jobInd = 1;
  const Data whole(wholeParams);
  size_t next = 0;
  return runJobs(params, whole, [&] { return next < order.size() ? order[next++] : 0; }, devices);
}

}  // namespace ASMC
