// TEST INFRASTRUCTURE ONLY — see oracle.hpp.  CPU restatement of the FastSMC IBD hot path.
// Every routine cites the reference lines it restates ("ref:" = /root/reference/ASMC_SRC/SRC/...).
#include "oracle.hpp"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>

#include <xmmintrin.h>
#include <zlib.h>

#include <atomic>
#include <thread>

namespace fo
{

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------

namespace
{

bool fileExists(const std::string& f)
{
  std::ifstream s(f);
  return s.good();
}

// Line reader that is transparent to gzip (zlib's gz* API reads plain files as-is).
class LineReader
{
  gzFile mFile = nullptr;
  std::vector<char> mBuf;

public:
  explicit LineReader(const std::string& path) : mBuf(1 << 16)
  {
    mFile = gzopen(path.c_str(), "rb");
    if (!mFile) {
      throw std::runtime_error("oracle: cannot open " + path);
    }
    gzbuffer(mFile, 1 << 20);
  }
  ~LineReader()
  {
    if (mFile) {
      gzclose(mFile);
    }
  }
  bool next(std::string& line)
  {
    line.clear();
    while (true) {
      if (!gzgets(mFile, mBuf.data(), static_cast<int>(mBuf.size()))) {
        return !line.empty();
      }
      const size_t n = std::strlen(mBuf.data());
      line.append(mBuf.data(), n);
      if (n && line.back() == '\n') {
        line.pop_back();
        return true;
      }
    }
  }
};

std::vector<std::string> splitWs(const std::string& line)
{
  std::vector<std::string> out;
  std::istringstream iss(line);
  std::string tok;
  while (iss >> tok) {
    out.push_back(tok);
  }
  return out;
}

std::string lower(std::string s)
{
  for (auto& c : s) {
    c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  }
  return s;
}

bool isSamplesHeader(const std::vector<std::string>& t)
{
  // ref: Data.cpp:234-238
  return t.size() >= 3 && ((t[0] == "ID_1" && t[1] == "ID_2" && t[2] == "missing") ||
                           (t[0] == "0" && t[1] == "0" && t[2] == "0"));
}

}  // namespace

// ref: StringUtils.cpp:36-39 — parse as long double, then narrow
float parseFloat(const std::string& s)
{
  return static_cast<float>(std::stold(s));
}

// ref: HmmUtils.cpp:65-79
float roundMorgans(const float value, const int precision, const float minv)
{
  if (value <= minv) {
    return minv;
  }
  const float correction = 10.f - static_cast<float>(precision);
  const float l10 = std::max<float>(0.f, floorf(log10f(value)) + correction);
  const float factor = powf(10.f, 10.f - l10);
  return roundf(value * factor) / factor;
}

// ref: HmmUtils.cpp:81-94
int roundPhysical(const int value, const int precision)
{
  if (value <= 1) {
    return 1;
  }
  const int l10 = std::max<int>(0, static_cast<int>(floor(log10(value))) - precision);
  const int factor = static_cast<int>(pow(10, l10));
  return static_cast<int>(round(value / static_cast<double>(factor))) * factor;
}

// ref: HmmUtils.cpp:153-164
unsigned getFromPosition(const std::vector<float>& gen, unsigned from, const float cmDist)
{
  float cum = 0.f;
  while (cum < cmDist && from > 0u) {
    --from;
    cum += (gen[from + 1u] - gen[from]) * 100.f;
  }
  return from;
}

// ref: HmmUtils.cpp:166-177
unsigned getToPosition(const std::vector<float>& gen, unsigned to, const float cmDist)
{
  float cum = 0.f;
  while (cum < cmDist && to + 1u < gen.size()) {
    ++to;
    cum += (gen[to] - gen[to - 1u]) * 100.f;
  }
  return std::min<unsigned>(to + 1u, static_cast<unsigned>(gen.size()));
}

// ref: HASHING/Utils.cpp:22-34
double cmBetween(const int w1, const int w2, const std::vector<float>& gen, const int wordSize)
{
  const size_t start = static_cast<size_t>(wordSize) * w1;
  const size_t end = std::min<size_t>(static_cast<size_t>(wordSize) * w2 + wordSize - 1, gen.size() - 1ul);
  return 100.0 * (gen[end] - gen[start]);
}

// ---------------------------------------------------------------------------------------------
// decoding quantities.  ref: DecodingQuantities.cpp:60-346
// ---------------------------------------------------------------------------------------------

void Quantities::load(const std::string& file)
{
  if (!fileExists(file)) {
    throw std::runtime_error("ERROR: Decoding quantities file " + file + " does not exist.\n");
  }
  LineReader in(file);
  std::string line;
  bool first = true;

  auto readFloats = [&](const std::string& l) {
    std::vector<float> v;
    for (const auto& t : splitWs(l)) {
      v.push_back(parseFloat(t));
    }
    return v;
  };
  auto readBlock = [&](int rows) {
    std::vector<std::vector<float>> block;
    for (int r = 0; r < rows; ++r) {
      in.next(line);
      block.push_back(readFloats(line));
      if (static_cast<int>(block.back().size()) != states) {
        throw std::runtime_error("oracle: bad row length in decoding quantities");
      }
    }
    return block;
  };

  enum Section { None, InitialStateProb, ColumnRatios, RowRatios, Uvec, Bvec, Dvec, Homozygous };
  Section section = None;

  while (in.next(line)) {
    if (first) {
      first = false;
      if (line != "TransitionType") {  // ref: DecodingQuantities.cpp:39-57
        throw std::runtime_error("ERROR: Decoding quantities file " + file +
                                 " does not seem to contain the correct information.\n");
      }
    }
    const auto tok = splitWs(line);
    if (tok.empty()) {
      continue;
    }
    const std::string head = lower(tok[0]);
    if (head == "states") {
      in.next(line);
      states = std::stoi(line);
    } else if (head == "transitiontype" || head == "sizevector") {
      in.next(line);
    } else if (head == "csfssamples") {
      in.next(line);
      csfsSamples = std::stoi(line);
      csfs.assign(csfsSamples - 1, {});
      foldedCsfs.assign(csfsSamples - 1, {});
      ascCsfs.assign(csfsSamples - 1, {});
      foldedAscCsfs.assign(csfsSamples - 1, {});
    } else if (head == "timevector") {
      in.next(line);
      timeVector = readFloats(line);
    } else if (head == "expectedtimes") {
      in.next(line);
      expectedTimes = readFloats(line);
    } else if (head == "discretization") {
      in.next(line);
      discretization = readFloats(line);
    } else if (head == "classicemission") {
      classicEmission = readBlock(2);
    } else if (head == "compressedascertainedemission") {
      compressedEmission = readBlock(2);
    } else if (head == "csfs") {
      csfs.at(std::stoi(tok.at(1))) = readBlock(3);
    } else if (head == "foldedcsfs") {
      foldedCsfs.at(std::stoi(tok.at(1))) = readBlock(2);
    } else if (head == "ascertainedcsfs") {
      ascCsfs.at(std::stoi(tok.at(1))) = readBlock(3);
    } else if (head == "foldedascertainedcsfs") {
      foldedAscCsfs.at(std::stoi(tok.at(1))) = readBlock(2);
    } else if (head == "homozygousemissions") {
      section = Homozygous;
    } else if (head == "initialstateprob") {
      section = InitialStateProb;
    } else if (head == "columnratios") {
      section = ColumnRatios;
    } else if (head == "rowratios") {
      section = RowRatios;
    } else if (head == "uvectors") {
      section = Uvec;
    } else if (head == "bvectors") {
      section = Bvec;
    } else if (head == "dvectors") {
      section = Dvec;
    } else {
      // content row of the current keyed section
      auto keyed = [&](std::unordered_map<float, std::vector<float>>& m) {
        std::vector<float> row(states, 0.f);
        for (size_t i = 1; i < tok.size(); ++i) {
          row[i - 1] = parseFloat(tok[i]);
        }
        m[parseFloat(tok[0])] = row;
      };
      switch (section) {
      case ColumnRatios:
        columnRatios.assign(states, 0.f);
        for (size_t i = 0; i < tok.size(); ++i) {
          columnRatios[i] = parseFloat(tok[i]);
        }
        break;
      case InitialStateProb:
        initialStateProb.assign(states, 0.f);
        for (size_t i = 0; i < tok.size(); ++i) {
          initialStateProb[i] = parseFloat(tok[i]);
        }
        break;
      case RowRatios:
        keyed(RR);
        break;
      case Uvec:
        keyed(U);
        break;
      case Bvec:
        keyed(B);
        break;
      case Dvec:
        keyed(D);
        break;
      default:
        break;  // homozygous emissions: sequence mode only, out of scope
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// data loading.  ref: Data.cpp:36-96
// ---------------------------------------------------------------------------------------------

Oracle::Oracle(const Params& p, const bool asmcMode) : params(p)
{
  const std::string& root = params.inFileRoot;
  // ref: Data.cpp:49-50 — count sites and samples
  {
    std::string hapFile;
    for (const char* ext : {".hap.gz", ".hap", ".haps.gz", ".haps"}) {
      if (fileExists(root + ext)) {
        hapFile = root + ext;
        break;
      }
    }
    if (hapFile.empty()) {
      throw std::runtime_error("oracle: no hap file for " + root);
    }
    LineReader in(hapFile);
    std::string line;
    while (in.next(line)) {
      ++sites;
    }
  }
  {
    LineReader in(fileExists(root + ".samples") ? root + ".samples" : root + ".sample");
    std::string line;
    while (in.next(line)) {
      if (!isSamplesHeader(splitWs(line))) {
        ++sampleSizeTotal;
      }
    }
  }

  // ref: Data.cpp:55-60
  if (params.useKnownSeed) {
    std::srand(1234u);
  } else {
    std::random_device rd;
    std::srand(rd());
  }

  // ref: Data.cpp:62-80 — job geometry (jobs/jobInd are never -1 on this path)
  {
    const double n = static_cast<double>(sampleSizeTotal);
    windowSize = static_cast<unsigned>(ceil(sqrt((2. * pow(n, 2) - n) * 2. / params.jobs)));
    if (windowSize % 2 != 0) {
      ++windowSize;
    }
    w_i = 1;
    int rowJobs = 1, cumJobs = 1;
    while (cumJobs < params.jobInd) {
      ++w_i;
      rowJobs += 2;
      cumJobs += rowJobs;
    }
    const int r = rowJobs - (cumJobs - params.jobInd);
    w_j = static_cast<unsigned>(ceil(static_cast<float>(r) / 2));
    aboveDiag = (r % 2 == 1);
  }

  loadSamples();
  hap.assign(2 * famId.size(), std::vector<uint8_t>(sites, 0));
  if (!asmcMode) {
    loadMapAndHapsFastSMC();
  } else {
    loadHapsAndMapAsmc();
  }

  dq.load(params.decodingQuantFile);
  computeUndistinguished();
  prepareEmissions();

  // ref: HMM.cpp:504-513
  stateThreshold = 0;
  while (dq.discretization[stateThreshold] < static_cast<float>(params.time) &&
         stateThreshold < static_cast<unsigned>(dq.states)) {
    ++stateThreshold;
  }
  // ref: HMM.cpp:96-105
  probabilityThreshold = 0.f;
  for (unsigned i = 0; i < stateThreshold; ++i) {
    probabilityThreshold += dq.initialStateProb.at(i);
  }
  ageThreshold = params.noConditionalAgeEstimates ? static_cast<unsigned>(dq.states) : stateThreshold;
}

// ref: Data.cpp:251-262, FastSMC.cpp:62-66
bool Oracle::inJob(const unsigned n) const
{
  return (n >= ((w_i - 1) * windowSize) / 2 && n < (w_i * windowSize) / 2) ||
         (n >= ((w_j - 1) * windowSize) / 2 && n < (w_j * windowSize) / 2) ||
         (params.jobs == params.jobInd && n >= ((w_j - 1) * windowSize) / 2);
}

// ref: Data.cpp:212-249
void Oracle::loadSamples()
{
  const std::string& root = params.inFileRoot;
  LineReader in(fileExists(root + ".samples") ? root + ".samples" : root + ".sample");
  std::string line;
  unsigned n = 0;
  while (in.next(line)) {
    const auto t = splitWs(line);
    if (isSamplesHeader(t)) {
      continue;
    }
    if (inJob(n)) {
      famId.push_back(t.at(0));
      iid.push_back(t.at(1));
    }
    ++n;
  }
}

// ref: Data.cpp:98-141 (map), 397-515 (haps), 523-565 (interpolation + addMarker)
void Oracle::loadMapAndHapsFastSMC()
{
  const std::string& root = params.inFileRoot;
  std::vector<std::pair<unsigned long, double>> gmap;
  {
    LineReader in(fileExists(root + ".map.gz") ? root + ".map.gz" : root + ".map");
    std::string line;
    while (in.next(line)) {
      std::stringstream ss(line);
      std::string f0, f1, f2;
      ss >> f0 >> f1 >> f2;
      if (f0.empty()) {
        continue;
      }
      try {
        (void)std::stoi(f0);
      } catch (const std::invalid_argument&) {
        continue;  // header row
      }
      gmap.emplace_back(std::stol(f0), std::stod(f2));
    }
  }

  std::string hapFile;
  for (const char* ext : {".hap.gz", ".hap", ".haps.gz", ".haps"}) {
    if (fileExists(root + ext)) {
      hapFile = root + ext;
      break;
    }
  }
  LineReader in(hapFile);
  std::string line;
  totalCount.assign(sites, 0);
  derivedCount.assign(sites, 0);
  flipped.assign(sites, 0);
  unsigned g = 0;
  unsigned long largestBp = 0;
  int pos = 0;
  const unsigned N = static_cast<unsigned>(sampleSizeTotal);
  while (in.next(line)) {
    std::istringstream ss(line);
    std::string chr, snp, a0, a1;
    unsigned long bp = 0;
    if (!(ss >> chr >> snp >> bp >> a0 >> a1)) {
      break;
    }
    std::string rest;
    std::getline(ss, rest);
    if (!(rest.length() == 4ul * N || rest.length() == 4ul * N + 1)) {
      throw std::runtime_error("oracle: haps line has wrong length");
    }
    if (bp <= largestBp) {
      throw std::runtime_error("oracle: haps rows must be sorted by increasing physical position");
    }
    largestBp = bp;
    if (g == 0 && pos == 0) {
      const size_t c = chr.find(':');
      chrNumber = std::stoi(c == std::string::npos ? chr : chr.substr(0, c));
      if (chrNumber <= 0 || chrNumber > 1260) {
        chrNumber = 0;
      }
    }
    // genetic position by linear interpolation in the map (ref: Data.cpp:523-547)
    {
      while (bp > gmap[g].first && g < gmap.size() - 1) {
        ++g;
      }
      double cm;
      if (bp >= gmap[g].first || g == 0) {
        cm = gmap[g].second;
      } else {
        cm = gmap[g - 1].second +
             (bp - gmap[g - 1].first) * (gmap[g].second - gmap[g - 1].second) / (gmap[g].first - gmap[g - 1].first);
      }
      genPos.push_back(cm / 100.f);
      physPos.push_back(static_cast<int>(bp));
      if (pos > 0) {
        const double gd = genPos[pos] - genPos[pos - 1];
        const unsigned long pd = physPos[pos] - physPos[pos - 1];
        const float rr = gd / pd;
        if (pos == 1) {
          recRate.push_back(rr);
        }
        recRate.push_back(rr);
      }
    }
    const int total = 2 * static_cast<int>(N);
    int da = 0;
    for (unsigned i = 0; i < 2 * N; ++i) {
      da += (rest[2 * i + 1] == '1');
    }
    const bool minorValue = params.foldData ? (da <= total - da) : true;
    flipped[pos] = !minorValue;
    unsigned local = 0;
    for (unsigned d = 0; d < N; ++d) {
      if (!inJob(d)) {
        continue;
      }
      for (unsigned h = 0; h < 2; ++h) {
        const char c = rest[2 * (2 * d + h) + 1];
        if (c != '0' && c != '1') {
          throw std::runtime_error("oracle: hap is not '0' or '1'");
        }
        hap[2 * local + h][pos] = (c == '1') ? minorValue : !minorValue;
      }
      ++local;
    }
    totalCount[pos] = total;
    derivedCount[pos] = params.foldData ? std::min(da, total - da) : da;
    ++pos;
  }
}

// ASMC (non-FastSMC) loading.  ref: Data.cpp:162-210 (PLINK-style map), 318-395 (haps)
void Oracle::loadHapsAndMapAsmc()
{
  const std::string& root = params.inFileRoot;
  std::string hapFile;
  for (const char* ext : {".hap.gz", ".hap", ".haps.gz", ".haps"}) {
    if (fileExists(root + ext)) {
      hapFile = root + ext;
      break;
    }
  }
  {
    LineReader in(hapFile);
    std::string line;
    totalCount.assign(sites, 0);
    derivedCount.assign(sites, 0);
    flipped.assign(sites, 0);
    int pos = 0;
    const unsigned nHap = 2 * static_cast<unsigned>(famId.size());
    while (in.next(line)) {
      std::istringstream ss(line);
      std::string chr, snp, a0, a1;
      unsigned long bp = 0;
      if (!(ss >> chr >> snp >> bp >> a0 >> a1)) {
        break;
      }
      std::string rest;
      std::getline(ss, rest);
      int da = 0;
      for (unsigned i = 0; i < nHap; ++i) {
        da += (rest[2 * i + 1] == '1');
      }
      const int total = static_cast<int>(nHap);
      const bool minorValue = params.foldData ? (da <= total - da) : true;
      flipped[pos] = !minorValue;
      for (unsigned i = 0; i < nHap; ++i) {
        hap[i][pos] = (rest[2 * i + 1] == '1') ? minorValue : !minorValue;
      }
      totalCount[pos] = total;
      derivedCount[pos] = params.foldData ? std::min(da, total - da) : da;
      ++pos;
    }
  }
  {
    LineReader in(fileExists(root + ".map.gz") ? root + ".map.gz" : root + ".map");
    std::string line;
    genPos.assign(sites, 0.f);
    physPos.assign(sites, 0);
    recRate.assign(sites, 0.f);
    int pos = 0;
    while (in.next(line)) {
      const auto t = splitWs(line);
      genPos[pos] = parseFloat(t.at(2)) / 100.f;
      physPos[pos] = std::stoi(t.at(3));
      if (pos > 0) {
        recRate[pos] = (genPos[pos] - genPos[pos - 1]) / (physPos[pos] - physPos[pos - 1]);
      }
      ++pos;
    }
  }
}

// libstdc++'s std::shuffle, written out so that both generations of its uniform-integer draw can be
// reproduced.  The reference draws through std::shuffle(..., std::mt19937(std::rand()))
// (ref: Data.cpp:154).  libstdc++ >= 11 downsizes a 32-bit engine draw with Lemire's multiply-shift
// (bits/uniform_int_dist.h, _S_nd); libstdc++ <= 10 — which produced the reference's golden files
// (CI: g++-9/g++-10, ref: ../../.github/workflows/ubuntu-regression.yml) — used divide-and-reject.
// flavor 0 = the C++ library this oracle is compiled with (std::shuffle itself),
// flavor 1 = Lemire written out (must equal flavor 0 on libstdc++ >= 11; checked by the tests),
// flavor 2 = divide-and-reject (golden-file era).
namespace
{
inline uint64_t drawBelow(std::mt19937& g, const uint64_t n, const int flavor)
{
  if (flavor == 2) {
    const uint64_t range = 0xFFFFFFFFull;
    const uint64_t scaling = range / n;
    const uint64_t past = n * scaling;
    uint64_t r;
    do {
      r = g();
    } while (r >= past);
    return r / scaling;
  }
  const uint32_t range = static_cast<uint32_t>(n);
  uint64_t product = static_cast<uint64_t>(g()) * range;
  uint32_t low = static_cast<uint32_t>(product);
  if (low < range) {
    const uint32_t threshold = -range % range;
    while (low < threshold) {
      product = static_cast<uint64_t>(g()) * range;
      low = static_cast<uint32_t>(product);
    }
  }
  return product >> 32;
}

void shuffleUrn(std::vector<unsigned short>& v, std::mt19937 g, const int flavor)
{
  if (flavor == 0) {
    std::shuffle(v.begin(), v.end(), g);
    return;
  }
  const uint64_t n = v.size();
  if (n == 0) {
    return;
  }
  if (0xFFFFFFFFull / n >= n) {
    uint64_t i = 1;
    if (n % 2 == 0) {
      std::swap(v[i], v[drawBelow(g, 2, flavor)]);
      ++i;
    }
    while (i != n) {
      const uint64_t swapRange = i + 1;
      const uint64_t x = drawBelow(g, swapRange * (swapRange + 1), flavor);
      std::swap(v[i], v[x / (swapRange + 1)]);
      ++i;
      std::swap(v[i], v[x % (swapRange + 1)]);
      ++i;
    }
    return;
  }
  for (uint64_t i = 1; i < n; ++i) {
    std::swap(v[i], v[drawBelow(g, i + 1, flavor)]);
  }
}
}  // namespace

// ref: Data.cpp:144-160 (sampleHypergeometric), 567-599 (calculateUndistinguishedCounts).
// The draw sequence (std::rand → std::mt19937 → shuffle) must be made in exactly this order.
void Oracle::computeUndistinguished()
{
  const int nCsfs = dq.csfsSamples;
  undistinguished.assign(sites, {0, 0, 0});
  for (int s = 0; s < sites; ++s) {
    const int derived = derivedCount[s];
    const int total = totalCount[s];
    if (params.usingCSFS && nCsfs > total) {
      throw std::runtime_error("oracle: CSFS requires more samples than the data holds");
    }
    for (int dist = 0; dist < 3; ++dist) {
      const int population = total - 2;
      const int successes = derived - dist;
      int sample;
      if (successes < 0 || successes > population) {
        sample = -1;
      } else {
        std::vector<unsigned short> urn(population, 0);
        std::fill(urn.begin(), urn.begin() + successes, 1);
        shuffleUrn(urn, std::mt19937(std::rand()), params.shuffleFlavor);
        sample = std::accumulate(urn.begin(), urn.begin() + (nCsfs - 2), 0);
      }
      if (params.foldData && (sample + dist > nCsfs / 2)) {
        sample = nCsfs - 2 - sample;
      }
      undistinguished[s][dist] = sample;
    }
  }
}

// ref: HMM.cpp:159-256 (array mode only: decodingSequence == false)
void Oracle::prepareEmissions()
{
  const int S = dq.states;
  e1.assign(static_cast<size_t>(sites) * S, 0.f);
  e0m1.assign(static_cast<size_t>(sites) * S, 0.f);
  e2m0.assign(static_cast<size_t>(sites) * S, 0.f);

  std::vector<uint8_t> useCsfs(sites, 0);
  if (params.skipCSFSdistance < std::numeric_limits<float>::infinity()) {
    useCsfs[0] = 1;
    float last = 0.f;
    for (int pos = 1; pos < sites; ++pos) {
      if (genPos[pos] - last >= params.skipCSFSdistance) {
        useCsfs[pos] = 1;
        last = genPos[pos];
      }
    }
  }
  for (int pos = 0; pos < sites; ++pos) {
    float* o1 = &e1[static_cast<size_t>(pos) * S];
    float* o0 = &e0m1[static_cast<size_t>(pos) * S];
    float* o2 = &e2m0[static_cast<size_t>(pos) * S];
    if (!useCsfs[pos]) {
      for (int k = 0; k < S; ++k) {
        o1[k] = dq.compressedEmission[1][k];
        o0[k] = dq.compressedEmission[0][k] - dq.compressedEmission[1][k];
        o2[k] = 0.f;
      }
      continue;
    }
    const int u0 = undistinguished[pos][0], u1 = undistinguished[pos][1], u2 = undistinguished[pos][2];
    if (params.foldData) {
      const auto& T = dq.foldedAscCsfs;
      for (int k = 0; k < S; ++k) {
        o1[k] = (u1 >= 0) ? T[u1][1][k] : 0.f;
        o0[k] = T[u0][0][k] - o1[k];
        o2[k] = (u2 >= 0) ? (T[u2][0][k] - T[u0][0][k]) : (0 - T[u0][0][k]);
      }
    } else {
      const auto& T = dq.ascCsfs;
      for (int k = 0; k < S; ++k) {
        o1[k] = (u1 >= 0) ? T[u1][1][k] : 0.f;
        const float em0 = (u0 >= 0) ? T[u0][0][k] : 0.f;
        o0[k] = em0 - o1[k];
        if (u2 >= 0) {
          const bool mono = (u2 == dq.csfsSamples - 2);  // monomorphic derived folds to CSFS[0][0]
          o2[k] = T[mono ? 0 : u2][mono ? 0 : 2][k] - em0;
        } else {
          o2[k] = 0 - em0;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the HMM.  ref: HMM.cpp:639-1041 (NO_SSE branches), HmmUtils.cpp:102-151
// Buffers are state-major with the pair (lane) index fastest, as in the reference; every lane is
// arithmetically independent, so decoding n lanes together equals decoding them one at a time.
// ---------------------------------------------------------------------------------------------

namespace
{

// ref: HmmUtils.cpp:102-151 (calculateScalingBatch + applyScalingBatch, NO_SSE)
inline void rescale(float* vec, const int S, const int n, float* sums)
{
  for (int v = 0; v < n; ++v) {
    sums[v] = 0.f;
  }
  for (int k = 0; k < S; ++k) {
    for (int v = 0; v < n; ++v) {
      sums[v] += vec[k * n + v];
    }
  }
  for (int v = 0; v < n; ++v) {
    sums[v] = 1.0f / sums[v];
  }
  for (int k = 0; k < S; ++k) {
    for (int v = 0; v < n; ++v) {
      vec[k * n + v] *= sums[v];
    }
  }
}

}  // namespace

void Oracle::decodeBatch(const std::vector<PairObs>& pairs, const unsigned from, const unsigned to,
                         std::vector<float>& posterior) const
{
  const int S = dq.states;
  const int n = static_cast<int>(pairs.size());
  const long len = static_cast<long>(to) - from;
  const size_t plane = static_cast<size_t>(S) * n;

  // ref: HMM.cpp:147-157,647-652 — observation indicators
  std::vector<float> isZero(static_cast<size_t>(len) * n), isTwo(static_cast<size_t>(len) * n);
  for (int v = 0; v < n; ++v) {
    const auto& a = hap[hapIndex(pairs[v], false)];
    const auto& b = hap[hapIndex(pairs[v], true)];
    for (long p = 0; p < len; ++p) {
      isZero[p * n + v] = (a[from + p] ^ b[from + p]) ? 0.0f : 1.0f;
      isTwo[p * n + v] = (a[from + p] & b[from + p]) ? 1.0f : 0.0f;
    }
  }

  posterior.resize(static_cast<size_t>(len) * plane);  // holds alpha, then the posterior (every element is written)
  // per-thread scratch, reused across batches (the reference allocates its buffers once per HMM, ref: HMM.cpp:111-116)
  static thread_local std::vector<float> beta;
  beta.resize(static_cast<size_t>(len) * plane);
  std::vector<float> scratchA(plane), scratchB(plane), lane1(n), sums(n);
  const std::vector<float>& cr = dq.columnRatios;

  // ---- forward.  ref: HMM.cpp:725-784
  {
    float* __restrict a0 = &posterior[0];
    for (int k = 0; k < S; ++k) {
      const float pi = dq.initialStateProb[k];
      const float c1 = e1[from * static_cast<size_t>(S) + k], c0 = e0m1[from * static_cast<size_t>(S) + k],
                  c2 = e2m0[from * static_cast<size_t>(S) + k];
      const float* __restrict z = &isZero[0];
      const float* __restrict t = &isTwo[0];
      float* __restrict out = a0 + static_cast<size_t>(k) * n;
#pragma GCC ivdep
      for (int v = 0; v < n; ++v) {
        out[v] = pi * (c1 + c0 * z[v] + c2 * t[v]);
      }
    }
    rescale(a0, S, n, sums.data());
  }
  float* __restrict alphaC = scratchA.data();
  float* __restrict AU = lane1.data();
  for (long p = 1; p < len; ++p) {
    const float dist = roundMorgans(genPos[from + p] - genPos[from + p - 1], 2, 1e-10f);
    const float* Bv = dq.B.at(dist).data();
    const float* Uv = dq.U.at(dist).data();
    const float* Dv = dq.D.at(dist).data();
    const float* __restrict prev = &posterior[(p - 1) * plane];
    float* __restrict next = &posterior[p * plane];
    const float* __restrict z = &isZero[p * n];
    const float* __restrict t = &isTwo[p * n];
    const float* E1 = &e1[(from + p) * static_cast<size_t>(S)];
    const float* E0 = &e0m1[(from + p) * static_cast<size_t>(S)];
    const float* E2 = &e2m0[(from + p) * static_cast<size_t>(S)];
    // ref: HMM.cpp:799-814 — suffix sums of the previous alpha
    std::memcpy(&alphaC[static_cast<size_t>(S - 1) * n], &prev[static_cast<size_t>(S - 1) * n], n * sizeof(float));
    for (int k = S - 2; k >= 0; --k) {
      float* __restrict o = alphaC + static_cast<size_t>(k) * n;
      const float* __restrict up = alphaC + static_cast<size_t>(k + 1) * n;
      const float* __restrict pk = prev + static_cast<size_t>(k) * n;
#pragma GCC ivdep
      for (int v = 0; v < n; ++v) {
        o[v] = up[v] + pk[v];
      }
    }
    // ref: HMM.cpp:816-830
    std::fill(AU, AU + n, 0.f);
    for (int k = 0; k < S; ++k) {
      const float c1 = E1[k], c0 = E0[k], c2 = E2[k], dk = Dv[k];
      const float* __restrict pk = prev + static_cast<size_t>(k) * n;
      float* __restrict o = next + static_cast<size_t>(k) * n;
      if (k) {
        const float u = Uv[k - 1], c = cr[k - 1];
        const float* __restrict pm = prev + static_cast<size_t>(k - 1) * n;
#pragma GCC ivdep
        for (int v = 0; v < n; ++v) {
          AU[v] = u * pm[v] + c * AU[v];
        }
      }
      if (k < S - 1) {
        const float bk = Bv[k];
        const float* __restrict ac = alphaC + static_cast<size_t>(k + 1) * n;
#pragma GCC ivdep
        for (int v = 0; v < n; ++v) {
          float term = AU[v] + dk * pk[v];
          term += bk * ac[v];
          o[v] = (c1 + c0 * z[v] + c2 * t[v]) * term;
        }
      } else {
#pragma GCC ivdep
        for (int v = 0; v < n; ++v) {
          const float term = AU[v] + dk * pk[v];
          o[v] = (c1 + c0 * z[v] + c2 * t[v]) * term;
        }
      }
    }
    rescale(next, S, n, sums.data());
  }

  // ---- backward.  ref: HMM.cpp:882-940
  {
    float* last = &beta[(len - 1) * plane];
    std::fill(last, last + plane, 1.0f);
    rescale(last, S, n, sums.data());
  }
  float* __restrict vec = scratchA.data();
  float* __restrict BU = scratchB.data();
  float* __restrict BL = lane1.data();
  for (long p = len - 2; p >= 0; --p) {
    const float dist = roundMorgans(genPos[from + p + 1] - genPos[from + p], 2, 1e-10f);
    const float* Bv = dq.B.at(dist).data();
    const float* Uv = dq.U.at(dist).data();
    const float* Rv = dq.RR.at(dist).data();
    const float* Dv = dq.D.at(dist).data();
    const float* __restrict lastBeta = &beta[(p + 1) * plane];
    float* __restrict cur = &beta[p * plane];
    const float* __restrict z = &isZero[(p + 1) * n];
    const float* __restrict t = &isTwo[(p + 1) * n];
    const float* E1 = &e1[(from + p + 1) * static_cast<size_t>(S)];
    const float* E0 = &e0m1[(from + p + 1) * static_cast<size_t>(S)];
    const float* E2 = &e2m0[(from + p + 1) * static_cast<size_t>(S)];
    // ref: HMM.cpp:957-964
    for (int k = 0; k < S; ++k) {
      const float c1 = E1[k], c0 = E0[k], c2 = E2[k];
      const float* __restrict lb = lastBeta + static_cast<size_t>(k) * n;
      float* __restrict o = vec + static_cast<size_t>(k) * n;
#pragma GCC ivdep
      for (int v = 0; v < n; ++v) {
        o[v] = lb[v] * (c1 + c0 * z[v] + c2 * t[v]);
      }
    }
    // ref: HMM.cpp:986-990
    std::fill(BU + static_cast<size_t>(S - 1) * n, BU + static_cast<size_t>(S) * n, 0.f);
    for (int k = S - 2; k >= 0; --k) {
      const float u = Uv[k], r = Rv[k];
      float* __restrict o = BU + static_cast<size_t>(k) * n;
      const float* __restrict up = BU + static_cast<size_t>(k + 1) * n;
      const float* __restrict vk = vec + static_cast<size_t>(k + 1) * n;
#pragma GCC ivdep
      for (int v = 0; v < n; ++v) {
        o[v] = u * vk[v] + r * up[v];
      }
    }
    // ref: HMM.cpp:1008-1016
    std::fill(BL, BL + n, 0.f);
    for (int k = 0; k < S; ++k) {
      if (k) {
        const float bk = Bv[k - 1];
        const float* __restrict vm = vec + static_cast<size_t>(k - 1) * n;
#pragma GCC ivdep
        for (int v = 0; v < n; ++v) {
          BL[v] += bk * vm[v];
        }
      }
      const float dk = Dv[k];
      const float* __restrict vk = vec + static_cast<size_t>(k) * n;
      const float* __restrict bu = BU + static_cast<size_t>(k) * n;
      float* __restrict o = cur + static_cast<size_t>(k) * n;
      // NO_SSE build: (BL + D*vec) + BU (ref: HMM.cpp:1014); SIMD builds: BL + (D*vec + BU) (ref: HMM.cpp:1036-1037)
      if (params.simdFlavor) {
#pragma GCC ivdep
        for (int v = 0; v < n; ++v) {
          o[v] = BL[v] + (dk * vk[v] + bu[v]);
        }
      } else {
#pragma GCC ivdep
        for (int v = 0; v < n; ++v) {
          o[v] = BL[v] + dk * vk[v] + bu[v];
        }
      }
    }
    rescale(cur, S, n, sums.data());
  }

  // ---- combine.  ref: HMM.cpp:669-692 (NO_SSE: exact reciprocal)
  for (long p = 0; p < len; ++p) {
    float* __restrict a = &posterior[p * plane];
    const float* __restrict b = &beta[p * plane];
    float* __restrict sm = sums.data();
    std::fill(sums.begin(), sums.end(), 0.f);
    for (int k = 0; k < S; ++k) {
      float* __restrict ak = a + static_cast<size_t>(k) * n;
      const float* __restrict bk = b + static_cast<size_t>(k) * n;
#pragma GCC ivdep
      for (int v = 0; v < n; ++v) {
        ak[v] *= bk[v];
        sm[v] += ak[v];
      }
    }
    for (int v = 0; v < n; ++v) {
      // The reference's SIMD builds normalise with the hardware approximate reciprocal
      // (ref: AvxDefinitions.hpp:36,50,64; HMM.cpp:704-709); its NO_SSE build divides exactly.
      sums[v] = params.simdFlavor ? _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(sums[v]))) : 1.0f / sums[v];
    }
    for (int k = 0; k < S; ++k) {
      float* __restrict ak = a + static_cast<size_t>(k) * n;
#pragma GCC ivdep
      for (int v = 0; v < n; ++v) {
        ak[v] *= sm[v];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// segment calling.  ref: HMM.cpp:1179-1357 (state machine), 1087-1107 (age estimates)
// The reference keeps one flag per confidence level; that is equivalent to tracking the level of
// the previous site (levels are mutually exclusive), which is how it is restated here.
// ---------------------------------------------------------------------------------------------

void Oracle::callSegments(const Batch& b, const uint32_t batchIdx, const std::vector<float>& posterior,
                          std::vector<Segment>& out) const
{
  const int S = dq.states;
  const int n = static_cast<int>(b.pairs.size());
  const size_t plane = static_cast<size_t>(S) * n;
  const bool ages = params.doPerPairPosteriorMean || params.doPerPairMAP;
  const unsigned nAcc = ages ? ageThreshold : 0u;
  const float pT = probabilityThreshold;

  for (int v = 0; v < n; ++v) {
    std::vector<float> post(nAcc, 0.f), acc(nAcc, 0.f), prevAcc(nAcc, 0.f);
    int level = -1;
    unsigned start = 0;
    float prob = 0.f;

    auto emit = [&](const unsigned s, const unsigned e, const float pr, const std::vector<float>& a) {
      Segment seg;
      seg.batch = batchIdx;
      seg.lane = static_cast<uint32_t>(v);
      seg.obs = b.pairs[v];
      seg.posStart = s;
      seg.posEnd = e;
      seg.prob = pr;
      if (params.doPerPairPosteriorMean) {  // ref: HMM.cpp:1087-1097
        const float norm = 1.f / std::accumulate(a.begin(), a.end(), 0.f);
        float mean = 0.f;
        for (size_t k = 0; k < a.size(); ++k) {
          mean += norm * a[k] * dq.expectedTimes[k];
        }
        seg.postMean = mean;
      }
      if (params.doPerPairMAP) {  // ref: HMM.cpp:1099-1107
        size_t best = 0;
        float bestRatio = a.empty() ? 0.f : a[0] / dq.initialStateProb[0];
        for (size_t k = 1; k < a.size(); ++k) {
          const float r = a[k] / dq.initialStateProb[k];
          if (bestRatio < r) {
            bestRatio = r;
            best = k;
          }
        }
        seg.mapState = static_cast<int>(best);
        seg.map = dq.expectedTimes[best];
      }
      out.push_back(seg);
    };

    for (unsigned pos = b.scanFrom; pos < b.scanTo; ++pos) {
      const float* p = &posterior[static_cast<size_t>(pos - b.from) * plane];
      float sum = 0.f;
      if (ages) {
        for (unsigned k = 0; k < nAcc; ++k) {
          const float x = p[k * n + v];
          post[k] = x;
          prevAcc[k] = acc[k];
          acc[k] += x;
          if (k < stateThreshold) {
            sum += x;
          }
        }
      } else {
        for (unsigned k = 0; k < stateThreshold; ++k) {
          sum += p[k * n + v];
        }
      }
      int now = -1;
      if (sum >= 1000 * pT) {
        now = 0;
      } else if (sum >= 100 * pT) {
        now = 1;
      } else if (sum >= 10 * pT) {
        now = 2;
      } else if (sum >= pT) {
        now = 3;
      }
      if (now != level) {
        if (level >= 0) {
          emit(start, pos - 1, prob, prevAcc);
        }
        if (now >= 0) {
          start = pos;
          acc = post;
          prob = sum;
        } else {
          prob = 0.f;
        }
      } else if (now >= 0) {
        prob += sum;
      }
      if (now >= 0 && pos == b.scanTo - 1) {
        emit(start, pos, prob, acc);
        prob = 0.f;
      }
      level = now;
    }
  }
}

// ref: HMM.cpp:1378-1410
void Oracle::perSiteSummary(const std::vector<float>& posterior, const unsigned nLanes, const unsigned len,
                            float* mean, int* map) const
{
  const int S = dq.states;
  for (unsigned v = 0; v < nLanes; ++v) {
    for (unsigned p = 0; p < len; ++p) {
      float m = 0.f, best = 0.f;
      int arg = 0;
      for (int k = 0; k < S; ++k) {
        const float x = posterior[(static_cast<size_t>(p) * S + k) * nLanes + v];
        m += x * dq.expectedTimes[k];
        if (best < x) {
          best = x;
          arg = k;
        }
      }
      if (mean) {
        mean[static_cast<size_t>(v) * len + p] = m;
      }
      if (map) {
        map[static_cast<size_t>(v) * len + p] = arg;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------------------------

// ref: HMM.cpp:311-357
std::vector<PairObs> Oracle::enumerateAllPairs() const
{
  const uint64_t N = famId.size();
  const uint64_t tot = params.withinOnly ? N : 2 * N * N - N;
  const uint64_t lo = tot * (params.jobInd - 1) / params.jobs;
  const uint64_t hi = tot * params.jobInd / params.jobs;
  std::vector<PairObs> out;
  uint64_t idx = 0;
  for (unsigned i = 0; i < N; ++i) {
    if (!params.withinOnly) {
      for (unsigned j = 0; j < i; ++j) {
        for (int iHap = 1; iHap <= 2; ++iHap) {
          for (int jHap = 1; jHap <= 2; ++jHap) {
            if (lo <= idx && idx < hi) {
              out.push_back(PairObs{jHap, j, iHap, i});
            }
            ++idx;
          }
        }
      }
    }
    if (lo <= idx && idx < hi) {
      out.push_back(PairObs{1, i, 2, i});
    }
    ++idx;
  }
  return out;
}

namespace
{

// Node-order model of boost::unordered_map (boost <= 1.79, prime bucket policy, identity hash for
// unsigned long keys).  The reference iterates SeedHash and ExtendHash in this order
// (ref: HASHING/SeedHash.hpp:34,80; HASHING/ExtendHash.hpp:29,88-97,112-115), so the order of
// decodeFromHashing calls — and with it the batch composition — depends on it.
template <class V> class NodeOrderedMap
{
  static constexpr int kSentinel = -2, kNone = -1;
  struct Node {
    uint64_t key;
    V val;
    int next;
    size_t bucket;
    bool live;
  };
  std::vector<Node> mNodes;
  std::vector<int> mFree;
  std::vector<int> mBucketPrev;  // per bucket: predecessor of its first node (kSentinel / node id / kNone)
  size_t mBucketCount = 17;
  size_t mSize = 0;
  size_t mMaxLoad = 0;
  int mHead = kNone;

  static size_t nextPrime(const size_t n)
  {
    static const size_t primes[] = {17ul,        29ul,        37ul,        53ul,        67ul,         79ul,
                                    97ul,        131ul,       193ul,       257ul,       389ul,        521ul,
                                    769ul,       1031ul,      1543ul,      2053ul,      3079ul,       6151ul,
                                    12289ul,     24593ul,     49157ul,     98317ul,     196613ul,     393241ul,
                                    786433ul,    1572869ul,   3145739ul,   6291469ul,   12582917ul,   25165843ul,
                                    50331653ul,  100663319ul, 201326611ul, 402653189ul, 805306457ul,  1610612741ul,
                                    3221225473ul, 4294967291ul};
    for (const size_t p : primes) {
      if (p >= n) {
        return p;
      }
    }
    return primes[sizeof(primes) / sizeof(primes[0]) - 1];
  }
  int& nextOf(const int p)
  {
    return p == kSentinel ? mHead : mNodes[p].next;
  }
  void createBuckets(const size_t count)
  {
    mBucketCount = count;
    mBucketPrev.assign(count, kNone);
    mMaxLoad = count;  // max load factor 1.0
  }
  void rehash(const size_t count)
  {
    createBuckets(count);
    int prev = kSentinel;
    while (nextOf(prev) != kNone) {
      const int n = nextOf(prev);
      const size_t b = mNodes[n].key % mBucketCount;
      mNodes[n].bucket = b;
      if (mBucketPrev[b] == kNone) {
        mBucketPrev[b] = prev;
        prev = n;
      } else {
        const int after = mNodes[n].next;
        mNodes[n].next = nextOf(mBucketPrev[b]);
        nextOf(mBucketPrev[b]) = n;
        nextOf(prev) = after;
      }
    }
  }

public:
  size_t size() const
  {
    return mSize;
  }
  int begin() const
  {
    return mHead;
  }
  int next(const int n) const
  {
    return mNodes[n].next;
  }
  uint64_t key(const int n) const
  {
    return mNodes[n].key;
  }
  V& value(const int n)
  {
    return mNodes[n].val;
  }

  // returns node id; `inserted` tells whether the key was new
  int insert(const uint64_t key, const V& val, bool& inserted)
  {
    if (!mBucketPrev.empty()) {
      const size_t b = key % mBucketCount;
      if (mBucketPrev[b] != kNone) {
        for (int n = nextOf(mBucketPrev[b]); n != kNone && mNodes[n].bucket == b; n = mNodes[n].next) {
          if (mNodes[n].key == key) {
            inserted = false;
            return n;
          }
        }
      }
    }
    inserted = true;
    // reserve_for_insert(size + 1)
    if (mBucketPrev.empty()) {
      createBuckets(std::max(mBucketCount, nextPrime(mSize + 1 + 1)));
    } else if (mSize + 1 > mMaxLoad) {
      const size_t want = nextPrime(std::max(mSize + 1, mSize + (mSize >> 1)) + 1);
      if (want != mBucketCount) {
        rehash(want);
      }
    }
    int id;
    if (!mFree.empty()) {
      id = mFree.back();
      mFree.pop_back();
    } else {
      id = static_cast<int>(mNodes.size());
      mNodes.push_back(Node{});
    }
    const size_t b = key % mBucketCount;
    mNodes[id] = Node{key, val, kNone, b, true};
    if (mBucketPrev[b] == kNone) {
      // empty bucket: the node becomes the head of the whole list
      if (mHead != kNone) {
        mBucketPrev[mNodes[mHead].bucket] = id;
      }
      mBucketPrev[b] = kSentinel;
      mNodes[id].next = mHead;
      mHead = id;
    } else {
      mNodes[id].next = nextOf(mBucketPrev[b]);
      nextOf(mBucketPrev[b]) = id;
    }
    ++mSize;
    return id;
  }

  // erase node n, return the id of the following node
  int erase(const int n)
  {
    const size_t b = mNodes[n].bucket;
    int prev = mBucketPrev[b];
    while (nextOf(prev) != n) {
      prev = nextOf(prev);
    }
    const int after = mNodes[n].next;
    nextOf(prev) = after;
    --mSize;
    bool sameBucketFollows = false;
    if (after != kNone) {
      const size_t b2 = mNodes[after].bucket;
      if (b2 == b) {
        sameBucketFollows = true;
      } else {
        mBucketPrev[b2] = prev;
      }
    }
    if (!sameBucketFollows && mBucketPrev[b] == prev) {
      mBucketPrev[b] = kNone;
    }
    mNodes[n].live = false;
    mFree.push_back(n);
    return after;
  }

  void clear()
  {
    if (!mSize) {
      return;
    }
    std::fill(mBucketPrev.begin(), mBucketPrev.end(), kNone);
    mNodes.clear();
    mFree.clear();
    mHead = kNone;
    mSize = 0;
  }
};

struct MatchInterval {
  int start = 0, end = 0;
};

}  // namespace

// ref: FastSMC.cpp:41-238 (word loop), HASHING/SeedHash.hpp:41-135, HASHING/ExtendHash.hpp:52-116,
//      HASHING/Match.hpp:42-57, HASHING/Individuals.hpp:45-62
std::vector<Candidate> Oracle::seedCandidates() const
{
  const int ws = params.hashingWordSize;
  if (ws != 64) {
    throw std::runtime_error("oracle: only 64-SNP words are supported");
  }
  const unsigned H = static_cast<unsigned>(hap.size());
  const int W = sites / ws;  // a trailing partial word is never hashed (ref: FastSMC.cpp:188-199)

  // raw (unfolded) alleles: '1' sets the bit (ref: FastSMC.cpp:176-186); bit b of word w = SNP 64w+b
  std::vector<uint64_t> words(static_cast<size_t>(W) * H, 0);
  for (unsigned h = 0; h < H; ++h) {
    for (int s = 0; s < W * ws; ++s) {
      const bool raw = flipped[s] ? !hap[h][s] : hap[h][s];
      if (raw) {
        words[static_cast<size_t>(s / ws) * H + h] |= (uint64_t{1} << (s % ws));
      }
    }
  }
  // global haplotype id of each local haplotype (ref: FastSMC.cpp:97-103, haploid mode)
  std::vector<unsigned> globalId;
  for (unsigned n = 0; n < static_cast<unsigned>(sampleSizeTotal); ++n) {
    if (inJob(n)) {
      globalId.push_back(2 * n);
      globalId.push_back(2 * n + 1);
    }
  }

  std::vector<Candidate> out;
  NodeOrderedMap<MatchInterval> extend;
  NodeOrderedMap<std::vector<unsigned>> seeds;
  const bool lastJob = (params.jobInd == params.jobs);
  const unsigned lo_i = (w_i - 1) * windowSize, lo_j = (w_j - 1) * windowSize;

  auto flush = [&](const int node) {  // ref: Match.hpp:42-52, ExtendHash.hpp:44-50
    const MatchInterval m = extend.value(node);
    if (cmBetween(m.start, m.end, genPos, ws) >= static_cast<double>(params.min_m)) {
      const uint64_t loc = extend.key(node);
      const unsigned second = static_cast<unsigned>(loc % H);
      const unsigned first = static_cast<unsigned>((loc - second) / H);
      out.push_back(Candidate{first, second, static_cast<uint32_t>(m.start * ws),
                              static_cast<uint32_t>(m.end * ws + ws - 1)});
    }
  };
  auto pairPasses = [&](const unsigned hi, const unsigned lo) {  // ref: SeedHash.hpp:99-129
    const unsigned gi = globalId[hi], gj = globalId[lo];
    if (lastJob) {
      return gi >= lo_i && gj >= lo_j && gj < lo_j + (gi - lo_i);
    }
    if (gi >= lo_i && gi < w_i * windowSize && gj >= lo_j && gj < w_j * windowSize) {
      return aboveDiag ? (gj < lo_j + (gi - lo_i)) : (gj >= lo_j + (gi - lo_i));
    }
    return false;
  };

  // extendAllPairs incl. the max_seeds sub-hash recursion (ref: SeedHash.hpp:56-69,85-93)
  std::function<void(NodeOrderedMap<std::vector<unsigned>>&, int, int, int)> extendAll =
      [&](NodeOrderedMap<std::vector<unsigned>>& sh, const int w, const int readWords, const int curWord) {
        for (int it = sh.begin(); it != -1; it = sh.next(it)) {
          const std::vector<unsigned>& bucket = sh.value(it);
          if (params.max_seeds != 0 && bucket.size() > static_cast<size_t>(params.max_seeds) && w + 1 < readWords) {
            NodeOrderedMap<std::vector<unsigned>> sub;
            for (const unsigned h : bucket) {
              bool ins;
              const int node = sub.insert(words[static_cast<size_t>(w + 1) * H + h], {}, ins);
              sub.value(node).push_back(h);
            }
            extendAll(sub, w + 1, readWords, curWord);
            continue;
          }
          for (size_t a = 0; a < bucket.size(); ++a) {
            for (size_t c = a + 1; c < bucket.size(); ++c) {
              const unsigned hi = std::max(bucket[a], bucket[c]);
              const unsigned lo = std::min(bucket[a], bucket[c]);
              if (!pairPasses(hi, lo)) {
                continue;
              }
              // ref: ExtendHash.hpp:61-81
              bool ins;
              const int node = extend.insert(static_cast<uint64_t>(lo) * H + hi, MatchInterval{curWord, 0}, ins);
              MatchInterval& m = extend.value(node);
              m.end = std::max(w, m.end);
            }
          }
        }
      };

  for (int cur = 0; cur < W; ++cur) {
    const int readWords = std::min(cur + params.constReadAhead, W);
    for (unsigned h = 0; h < H; ++h) {  // ref: FastSMC.cpp:204-206
      bool ins;
      const int node = seeds.insert(words[static_cast<size_t>(cur) * H + h], {}, ins);
      seeds.value(node).push_back(h);
    }
    if (static_cast<float>(seeds.size()) / static_cast<float>(H) > params.skip) {  // ref: FastSMC.cpp:212-215
      extendAll(seeds, cur, readWords, cur);
      // ref: ExtendHash.hpp:85-98
      const int limit = cur - params.gap;
      for (int it = extend.begin(); it != -1;) {
        if (extend.value(it).end < limit) {
          flush(it);
          it = extend.erase(it);
        } else {
          it = extend.next(it);
        }
      }
    } else {
      for (int it = extend.begin(); it != -1; it = extend.next(it)) {  // ref: ExtendHash.hpp:102-106
        extend.value(it).end = cur;
      }
    }
    seeds.clear();
  }
  for (int it = extend.begin(); it != -1;) {  // ref: ExtendHash.hpp:110-116
    flush(it);
    it = extend.erase(it);
  }
  return out;
}

// ref: HMM.cpp:470-502 (slot windows), 555-636 (batch windows)
std::vector<Batch> Oracle::makeBatches(const std::vector<PairObs>& pairs, const std::vector<Candidate>* cands) const
{
  std::vector<Batch> out;
  const size_t bs = static_cast<size_t>(params.batchSize);
  for (size_t i = 0; i < pairs.size(); i += bs) {
    Batch b;
    const size_t e = std::min(pairs.size(), i + bs);
    b.pairs.assign(pairs.begin() + i, pairs.begin() + e);
    if (cands) {
      unsigned lo = std::numeric_limits<unsigned>::max(), hi = 0;
      for (size_t c = i; c < e; ++c) {
        lo = std::min(lo, (*cands)[c].from);
        hi = std::max(hi, (*cands)[c].to);
      }
      b.scanFrom = lo;
      b.scanTo = hi;
    } else {
      b.scanFrom = 0;
      b.scanTo = static_cast<unsigned>(sites);
    }
    b.from = getFromPosition(genPos, b.scanFrom);
    b.to = getToPosition(genPos, b.scanTo);
    out.push_back(std::move(b));
  }
  return out;
}

// ref: HMM.cpp:1116-1141 — default ostream float formatting at 7 significant digits == "%.7g"
std::string Oracle::formatText(const Segment& s) const
{
  char buf[64];
  std::string r;
  r += famId[s.obs.aInd] + '\t' + iid[s.obs.aInd] + '\t' + std::to_string(s.obs.aHap) + '\t';
  r += famId[s.obs.bInd] + '\t' + iid[s.obs.bInd] + '\t' + std::to_string(s.obs.bHap) + '\t';
  r += std::to_string(chrNumber);
  r += '\t' + std::to_string(physPos[s.posStart]) + '\t' + std::to_string(physPos[s.posEnd]);
  if (params.outputIbdSegmentLength) {
    const float cm = 100.f * (genPos[s.posEnd] - genPos[s.posStart]);
    std::snprintf(buf, sizeof buf, "\t%.7g", static_cast<double>(cm));
    r += buf;
  }
  const double score = s.prob / static_cast<double>(s.posEnd - s.posStart + 1u);
  std::snprintf(buf, sizeof buf, "\t%.7g", score);
  r += buf;
  if (params.doPerPairPosteriorMean) {
    std::snprintf(buf, sizeof buf, "\t%.7g", static_cast<double>(s.postMean));
    r += buf;
  }
  if (params.doPerPairMAP) {
    std::snprintf(buf, sizeof buf, "\t%.7g", static_cast<double>(s.map));
    r += buf;
  }
  r += '\n';
  return r;
}

// ref: FastSMC.cpp:41-51,232 ; HMM.cpp:296-303,383-401,1147-1175
long Oracle::run(const std::string& outPath, int threads)
{
  std::vector<PairObs> pairs;
  lastCandidates.clear();
  if (params.hashing) {
    lastCandidates = seedCandidates();
    for (const auto& c : lastCandidates) {  // ref: HMM.cpp:483-486
      pairs.push_back(PairObs{c.hapA % 2 == 0 ? 1 : 2, c.hapA / 2, c.hapB % 2 == 0 ? 1 : 2, c.hapB / 2});
    }
  } else {
    pairs = enumerateAllPairs();
  }
  lastBatches = makeBatches(pairs, params.hashing ? &lastCandidates : nullptr);

  std::vector<std::vector<Segment>> perBatch(lastBatches.size());
  const auto t0 = std::chrono::steady_clock::now();
  if (threads <= 0) {
    threads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
  }
  {
    std::atomic<long> nextBatch{0};
    auto worker = [&] {
      std::vector<float> posterior;
      for (long b = nextBatch++; b < static_cast<long>(lastBatches.size()); b = nextBatch++) {
        decodeBatch(lastBatches[b].pairs, lastBatches[b].from, lastBatches[b].to, posterior);
        callSegments(lastBatches[b], static_cast<uint32_t>(b), posterior, perBatch[b]);
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
      pool.emplace_back(worker);
    }
    worker();
    for (auto& t : pool) {
      t.join();
    }
  }
  lastDecodeSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  lastPairSites = 0.0;
  for (const auto& b : lastBatches) {
    lastPairSites += static_cast<double>(b.pairs.size()) * (b.to - b.from);
  }
  lastSegments.clear();
  for (auto& v : perBatch) {
    lastSegments.insert(lastSegments.end(), v.begin(), v.end());
  }

  std::string path = outPath;
  if (path.empty()) {
    path = params.outFileRoot + "." + std::to_string(params.jobInd) + "." + std::to_string(params.jobs) +
           (params.BIN_OUT ? ".FastSMC.bibd.gz" : ".FastSMC.ibd.gz");
  }
  gzFile gz = gzopen(path.c_str(), params.BIN_OUT ? "wb" : "w");
  if (!gz) {
    throw std::runtime_error("oracle: cannot write " + path);
  }
  if (params.BIN_OUT) {
    gzwrite(gz, &params.outputIbdSegmentLength, sizeof(bool));
    gzwrite(gz, &params.doPerPairPosteriorMean, sizeof(bool));
    gzwrite(gz, &params.doPerPairMAP, sizeof(bool));
    gzwrite(gz, &chrNumber, sizeof(int));
    const unsigned nInd = static_cast<unsigned>(famId.size());
    gzwrite(gz, &nInd, sizeof(unsigned));
    for (unsigned i = 0; i < nInd; ++i) {
      unsigned len = static_cast<unsigned>(famId[i].size());
      gzwrite(gz, &len, sizeof(unsigned));
      gzwrite(gz, famId[i].c_str(), len);
      len = static_cast<unsigned>(iid[i].size());
      gzwrite(gz, &len, sizeof(unsigned));
      gzwrite(gz, iid[i].c_str(), len);
    }
  }
  for (const auto& s : lastSegments) {
    if (!params.BIN_OUT) {
      const std::string line = formatText(s);
      gzwrite(gz, line.c_str(), static_cast<unsigned>(line.size()));
    } else {
      const unsigned ind[2] = {s.obs.aInd, s.obs.bInd};
      const uint8_t hp[2] = {static_cast<uint8_t>(s.obs.aHap), static_cast<uint8_t>(s.obs.bHap)};
      const int bp[2] = {physPos[s.posStart], physPos[s.posEnd]};
      const float score = static_cast<float>(s.prob / static_cast<double>(s.posEnd - s.posStart + 1u));
      gzwrite(gz, &ind[0], sizeof(unsigned));
      gzwrite(gz, &hp[0], 1);
      gzwrite(gz, &ind[1], sizeof(unsigned));
      gzwrite(gz, &hp[1], 1);
      gzwrite(gz, &bp[0], sizeof(int));
      gzwrite(gz, &bp[1], sizeof(int));
      if (params.outputIbdSegmentLength) {
        const float cm = 100.f * (genPos[s.posEnd] - genPos[s.posStart]);
        gzwrite(gz, &cm, sizeof(float));
      }
      gzwrite(gz, &score, sizeof(float));
      if (params.doPerPairPosteriorMean) {
        gzwrite(gz, &s.postMean, sizeof(float));
      }
      if (params.doPerPairMAP) {
        gzwrite(gz, &s.map, sizeof(float));
      }
    }
  }
  gzclose(gz);
  return static_cast<long>(lastSegments.size());
}

}  // namespace fo
