// FastSMC_exe — command-line driver with the reference's options (ref: ASMC_SRC/SRC/main_fastsmc.cpp:13-25).
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "FastSMC.hpp"
#include "Partitioner.hpp"

// B200 build: `--allJobs G` runs every job 1..--jobs of the data set on G GPUs of this box (one host thread per GPU)
// instead of the single job --jobInd; outputs are the same per-job files.
int main(int argc, char* argv[])
{
  int allJobsOnDevices = 0;
  std::vector<char*> args;
  for (int i = 0; i < argc; ++i) {
    if (std::string(argv[i]) == "--allJobs" && i + 1 < argc) {
      allJobsOnDevices = std::atoi(argv[++i]);
    } else {
      args.push_back(argv[i]);
    }
  }
  argc = static_cast<int>(args.size());
  argv = args.data();
  DecodingParams params;
  if (!params.processCommandLineArgsFastSMC(argc, argv)) {
    std::cerr << "Error processing command line; exiting." << std::endl;
    return 1;
  }
  try {
    if (allJobsOnDevices > 0) {
      std::vector<int> devices(allJobsOnDevices);
      for (int d = 0; d < allJobsOnDevices; ++d) {
        devices[d] = d;
      }
      int failed = 0;
      for (const auto& r : ASMC::runAllJobs(params, devices)) {
        std::cout << "job " << r.jobInd << "/" << params.jobs << " on GPU " << r.device << ": " << r.segments
                  << " segments, " << r.wallSeconds << " s" << (r.error.empty() ? "" : " ERROR " + r.error) << std::endl;
        failed += !r.error.empty();
      }
      return failed ? 1 : 0;
    }
    ASMC::FastSMC fastSMC(params);
    fastSMC.run();
  } catch (const std::exception& e) {
    std::cerr << "ERROR: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
