// fastsmc_b200 host layer — deals the jobs of one data set to the GPUs of one box.
//
// FastSMC's only parallelism is process-level: `--jobs J --jobInd i` runs J independent processes over disjoint
// triangles of the sample-pair matrix, each writing its own output file (ref: ASMC_SRC/SRC/Data.cpp:62-80,
// cpp_example/FastSMC_example_multiple_jobs.sh:9-43; J must be a perfect square, DecodingParams.cpp:376-395).  Jobs
// share nothing, so no collective is needed: here one host thread per GPU pulls job indices from a queue and runs
// ASMC::FastSMC for each on its device; outputs are the reference's per-job files
// <outFileRoot>.<jobInd>.<jobs>.FastSMC.{ibd,bibd}.gz, identical to running the jobs one process at a time.
//
// The reference draws its emission tables through the process-global std::rand(); model preparation of
// different jobs is therefore serialised under a mutex (it is the same sequence for every job), decoding is not.
#pragma once

#include <functional>
#include <string>
#include <vector>

#include "Data.hpp"
#include "DecodingParams.hpp"

namespace ASMC
{

struct JobReport {
  int jobInd = 0;
  int device = 0;
  unsigned long candidates = 0, pairsDecoded = 0, segments = 0;
  double pairSites = 0.0;
  double kernelMs = 0.0;   // device time in the decode kernels
  double seedMs = 0.0;     // device time in the seeding kernels
  double wallSeconds = 0.0;
  // host wall time of the job's stages: cutting the job's samples out of the data set + model tables + upload,
  // fsmc_seed, candidate order, fsmc_decode calls, record formatting/compression
  double prepareSeconds = 0.0, seedSeconds = 0.0, orderSeconds = 0.0, decodeSeconds = 0.0, outputSeconds = 0.0;
  double cutSeconds = 0.0, tablesSeconds = 0.0, uploadSeconds = 0.0;  // parts of prepareSeconds
  std::string error;       // empty on success
};

// Diagonal jobs hold about half the pairs of off-diagonal ones: jobs are queued largest first.
std::vector<int> jobOrder(int jobs);

// Static split used when every rank is its own process (torchrun): the jobs of `rank` out of `world`, dealt round-robin
// over jobOrder(jobs).
std::vector<int> jobsOfRank(int jobs, int world, int rank);

// Runs the jobs that `nextJob` hands out (a job index in 1..params.jobs, or 0 when none is left; called under a lock)
// on `devices`, one host thread per entry, every job cut out of `whole` (the data set read with jobs = 1).  This is
// what a torchrun rank calls with a counter shared by all ranks as the job source.
std::vector<JobReport> runJobs(const DecodingParams& params, const Data& whole, const std::function<int()>& nextJob,
                               const std::vector<int>& devices);

// Runs jobs 1..params.jobs of the data set on `devices` (CUDA ordinals; a device may appear twice to run two host
// threads on it).  params.jobInd is ignored.  Returns one report per job, ordered by jobInd.
std::vector<JobReport> runAllJobs(const DecodingParams& params, const std::vector<int>& devices);

}  // namespace ASMC
