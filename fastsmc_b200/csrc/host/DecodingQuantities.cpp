// fastsmc_b200 host layer — see DecodingQuantities.hpp.
#include "DecodingQuantities.hpp"

#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <iostream>
#include <set>
#include <sstream>
#include <stdexcept>

#include "FileUtils.hpp"
#include "StringUtils.hpp"

DecodingQuantities::DecodingQuantities(const std::string& fileName)
{
  validateDecodingQuantitiesFile(fileName);
  std::cout << "Using precomputed decoding info from " << fileName << std::endl;
  createFromGzippedText(fileName);
}

const DecodingQuantities& DecodingQuantities::cached(const std::string& fileName)
{
  static std::mutex lock;
  static std::map<std::string, std::unique_ptr<DecodingQuantities>> table;
  std::lock_guard<std::mutex> g(lock);
  auto it = table.find(fileName);
  if (it == table.end()) {
    it = table.emplace(fileName, std::make_unique<DecodingQuantities>(fileName)).first;
  }
  return *it->second;
}

// ref: DecodingQuantities.cpp:39-58
void DecodingQuantities::validateDecodingQuantitiesFile(const std::string& fileName)
{
  if (!FileUtils::fileExists(fileName)) {
    throw std::runtime_error("ERROR: Decoding quantities file " + fileName + " does not exist.\n");
  }
  FileUtils::LineReader in(fileName);
  std::string first;
  in.next(first);
  if (first != "TransitionType") {
    throw std::runtime_error("ERROR: Decoding quantities file " + fileName +
                             " does not seem to contain the correct information.\n" +
                             "Expected file to begin with \"TransitionType\", but instead found \"" + first + "\"\n");
  }
}

namespace
{

// Numbers are parsed as long double and narrowed, exactly like the reference's StringUtils::stof
// (ref: StringUtils.cpp:36-39); a plain strtof could round differently on some inputs.
void parseFloats(const std::string& line, std::vector<float>& out, const size_t skipTokens = 0)
{
  out.clear();
  const char* p = line.c_str();
  size_t tok = 0;
  for (;;) {
    while (*p == ' ' || *p == '\t') {
      ++p;
    }
    if (!*p) {
      return;
    }
    char* end = nullptr;
    const long double v = std::strtold(p, &end);
    if (end == p) {
      throw std::runtime_error("ERROR: could not parse number in decoding quantities: " + line.substr(0, 40));
    }
    if (tok++ >= skipTokens) {
      out.push_back(static_cast<float>(v));
    }
    p = end;
  }
}

std::string firstToken(const std::string& line, std::string* second = nullptr)
{
  std::istringstream ss(line);
  std::string a, b;
  ss >> a >> b;
  if (second) {
    *second = b;
  }
  return a;
}

}  // namespace

void DecodingQuantities::createFromGzippedText(const std::string& fileName)
{
  FileUtils::LineReader in(fileName);
  std::string line, arg;
  std::vector<float> row;

  auto need = [&]() {
    if (!in.next(line)) {
      throw std::runtime_error("ERROR: unexpected end of decoding quantities file " + fileName);
    }
  };
  auto block = [&](const int rows) {
    std::vector<std::vector<float>> b;
    for (int r = 0; r < rows; ++r) {
      need();
      parseFloats(line, row);
      if (row.size() != states) {
        throw std::runtime_error("ERROR: a row of " + fileName + " does not hold one value per state");
      }
      b.push_back(row);
    }
    return b;
  };
  auto sized = [&](std::vector<std::vector<std::vector<float>>>& v, const std::string& idx) -> std::vector<std::vector<float>>& {
    const int i = std::stoi(idx);
    if (i < 0 || i >= static_cast<int>(v.size())) {
      throw std::runtime_error("ERROR: CSFS index out of range in " + fileName);
    }
    return v[i];
  };

  enum class Section { none, initialStateProb, columnRatios, rowRatios, U, B, D, homozygous } section = Section::none;

  while (in.next(line)) {
    if (line.empty()) {
      continue;
    }
    static const std::set<std::string> kHeaders = {
        "transitiontype", "sizevector", "states", "csfssamples", "timevector", "expectedtimes", "discretization",
        "classicemission", "compressedascertainedemission", "csfs", "foldedcsfs", "ascertainedcsfs",
        "foldedascertainedcsfs", "initialstateprob", "columnratios", "rowratios", "uvectors", "bvectors", "dvectors",
        "homozygousemissions"};
    const std::string head = StringUtils::toLower(firstToken(line, &arg));
    const bool isHeader = kHeaders.count(head) != 0;
    if (!isHeader) {
      // a content row of the current keyed section
      switch (section) {
      case Section::initialStateProb:  // both vectors are sized to `states` and zero-filled (ref: :299-310)
        parseFloats(line, initialStateProb);
        initialStateProb.resize(states, 0.f);
        break;
      case Section::columnRatios:
        parseFloats(line, columnRatios);
        columnRatios.resize(states, 0.f);
        break;
      case Section::rowRatios:
      case Section::U:
      case Section::B:
      case Section::D: {
        char* end = nullptr;
        const float key = static_cast<float>(std::strtold(line.c_str(), &end));
        parseFloats(line, row, 1);
        row.resize(states, 0.f);
        auto& m = section == Section::rowRatios ? rowRatioVectors
                  : section == Section::U       ? Uvectors
                  : section == Section::B       ? Bvectors
                                                : Dvectors;
        m[key] = row;
        break;
      }
      case Section::homozygous: {
        char* end = nullptr;
        const int key = static_cast<int>(std::strtol(line.c_str(), &end, 10));
        parseFloats(line, row, 1);
        homozygousEmissionMap[key] = row;
        break;
      }
      default:
        break;
      }
      continue;
    }
    if (head == "transitiontype" || head == "sizevector") {
      need();
    } else if (head == "states") {
      need();
      states = static_cast<unsigned>(std::stoi(line));
    } else if (head == "csfssamples") {
      need();
      CSFSSamples = std::stoi(line);
      const size_t n = CSFSSamples > 1 ? CSFSSamples - 1 : 0;
      CSFSmap.assign(n, {});
      foldedCSFSmap.assign(n, {});
      ascertainedCSFSmap.assign(n, {});
      foldedAscertainedCSFSmap.assign(n, {});
    } else if (head == "timevector") {
      need();
      parseFloats(line, timeVector);
    } else if (head == "expectedtimes") {
      need();
      parseFloats(line, expectedTimes);
    } else if (head == "discretization") {
      need();
      parseFloats(line, discretization);
    } else if (head == "classicemission") {
      classicEmissionTable = block(2);
    } else if (head == "compressedascertainedemission") {
      compressedEmissionTable = block(2);
    } else if (head == "csfs") {
      sized(CSFSmap, arg) = block(3);
    } else if (head == "foldedcsfs") {
      sized(foldedCSFSmap, arg) = block(2);
    } else if (head == "ascertainedcsfs") {
      sized(ascertainedCSFSmap, arg) = block(3);
    } else if (head == "foldedascertainedcsfs") {
      sized(foldedAscertainedCSFSmap, arg) = block(2);
    } else if (head == "initialstateprob") {
      section = Section::initialStateProb;
    } else if (head == "columnratios") {
      section = Section::columnRatios;
    } else if (head == "rowratios") {
      section = Section::rowRatios;
    } else if (head == "uvectors") {
      section = Section::U;
    } else if (head == "bvectors") {
      section = Section::B;
    } else if (head == "dvectors") {
      section = Section::D;
    } else if (head == "homozygousemissions") {
      section = Section::homozygous;
    }
  }
  if (states == 0 || initialStateProb.size() != states || expectedTimes.size() != states ||
      columnRatios.size() != states || discretization.size() < states) {
    throw std::runtime_error("ERROR: decoding quantities file " + fileName + " is incomplete (states " +
                             std::to_string(states) + ", initialStateProb " + std::to_string(initialStateProb.size()) +
                             ", expectedTimes " + std::to_string(expectedTimes.size()) + ", columnRatios " +
                             std::to_string(columnRatios.size()) + ", discretization " +
                             std::to_string(discretization.size()) + ").");
  }
}
