// FastSMC_exe — command-line driver with the reference's options (ref: ASMC_SRC/SRC/main_fastsmc.cpp:13-25).
#include <cstdlib>
#include <iostream>

#include "FastSMC.hpp"

int main(int argc, char* argv[])
{
  DecodingParams params;
  if (!params.processCommandLineArgsFastSMC(argc, argv)) {
    std::cerr << "Error processing command line; exiting." << std::endl;
    return 1;
  }
  try {
    ASMC::FastSMC fastSMC(params);
    fastSMC.run();
  } catch (const std::exception& e) {
    std::cerr << "ERROR: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
