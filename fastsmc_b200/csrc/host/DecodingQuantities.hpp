// fastsmc_b200 host layer — parser of the .decodingQuantities.gz model file.
// Same public fields as the reference class (ref: ASMC_SRC/SRC/DecodingQuantities.hpp:47-70); file grammar
// as read by ref: DecodingQuantities.cpp:60-346.
#pragma once

#include <string>
#include <unordered_map>
#include <vector>

class DecodingQuantities
{
public:
  unsigned int states = 0u;
  int CSFSSamples = 0;
  std::vector<float> initialStateProb;
  std::vector<float> expectedTimes;
  std::vector<float> discretization;
  std::vector<float> timeVector;
  std::vector<float> columnRatios;
  std::vector<std::vector<float>> classicEmissionTable;
  std::vector<std::vector<float>> compressedEmissionTable;
  // transition rows keyed by the exact float distance parsed from the file
  std::unordered_map<float, std::vector<float>> Dvectors;
  std::unordered_map<float, std::vector<float>> Bvectors;
  std::unordered_map<float, std::vector<float>> Uvectors;
  std::unordered_map<float, std::vector<float>> rowRatioVectors;
  std::unordered_map<int, std::vector<float>> homozygousEmissionMap;
  std::vector<std::vector<std::vector<float>>> CSFSmap;                  // [undistinguished][distinguished 0..2][state]
  std::vector<std::vector<std::vector<float>>> foldedCSFSmap;            // [undistinguished][0..1][state]
  std::vector<std::vector<std::vector<float>>> ascertainedCSFSmap;
  std::vector<std::vector<std::vector<float>>> foldedAscertainedCSFSmap;

  DecodingQuantities() = default;
  explicit DecodingQuantities(const std::string& fileName);
  /// The parsed contents of `fileName`, read once per process: the jobs of one data set (runJobs / runAllJobs) all use
  /// the same table, and parsing its 20-40 MB of gzipped text costs more than a small job's decoding.
  static const DecodingQuantities& cached(const std::string& fileName);

private:
  void validateDecodingQuantitiesFile(const std::string& fileName);
  void createFromGzippedText(const std::string& fileName);
};
