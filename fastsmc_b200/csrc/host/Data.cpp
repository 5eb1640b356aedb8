// fastsmc_b200 host layer — see Data.hpp.
#include "Data.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <mutex>
#include <random>
#include <sstream>
#include <stdexcept>

#include "FileUtils.hpp"
#include "StringUtils.hpp"

namespace
{

const std::initializer_list<const char*> kHapExts = {".hap.gz", ".hap", ".haps.gz", ".haps"};

bool isSamplesHeader(const std::vector<std::string>& t)
{
  // ref: Data.cpp:234-238
  return t.size() >= 3 &&
         ((t[0] == "ID_1" && t[1] == "ID_2" && t[2] == "missing") || (t[0] == "0" && t[1] == "0" && t[2] == "0"));
}

std::string samplesFile(const std::string& root)
{
  const std::string f = FileUtils::firstExisting(root, {".samples", ".sample"});
  if (f.empty()) {
    std::cerr << "ERROR. Could not find sample file in " + root + ".sample or " + root + ".samples" << std::endl;
    exit(1);
  }
  return f;
}

std::string hapsFile(const std::string& root)
{
  const std::string f = FileUtils::firstExisting(root, kHapExts);
  if (f.empty()) {
    std::cerr << "ERROR. Could not find hap file in " + root + ".hap.gz, " + root + ".hap, " + ".haps.gz, or " + root +
                     ".haps"
              << std::endl;
    exit(1);
  }
  return f;
}

}  // namespace

int Data::countHapLines(std::string inFileRoot)
{
  FileUtils::LineReader in(hapsFile(inFileRoot));
  std::string line;
  int n = 0;
  while (in.next(line)) {
    ++n;
  }
  return n;
}

int Data::countSamplesLines(std::string inFileRoot)
{
  FileUtils::LineReader in(samplesFile(inFileRoot));
  std::string line;
  int n = 0;
  while (in.next(line)) {
    if (!isSamplesHeader(StringUtils::tokenizeMultipleDelimiters(line))) {
      ++n;
    }
  }
  return n;
}

// ref: Data.cpp:55-80
void Data::setJobGeometry(const DecodingParams& params)
{
  jobs = params.jobs;
  jobInd = params.jobInd;
  foldToMinorAlleles = params.foldData;
  decodingUsesCSFS = params.usingCSFS;
  mJobbing = (jobInd != -1) && (jobs != -1);
  // ref: Data.cpp:55-61 seeds std::rand here; the only consumer is calculateUndistinguishedCounts, which seeds and draws
  // under one lock, so that Data objects built on several host threads (one per GPU) do not share the sequence
  mUseKnownSeed = params.useKnownSeed;
  if (mJobbing) {
    const double n = static_cast<double>(sampleSize);
    windowSize = static_cast<unsigned>(std::ceil(std::sqrt((2. * n * n - n) * 2. / jobs)));
    if (windowSize % 2 != 0) {
      ++windowSize;
    }
    w_i = 1;
    int inRow = 1, upTo = 1;
    while (upTo < jobInd) {
      ++w_i;
      inRow += 2;
      upTo += inRow;
    }
    const int r = inRow - (upTo - jobInd);
    w_j = static_cast<unsigned>(std::ceil(static_cast<float>(r) / 2));
    is_j_above_diag = (r % 2 == 1);
  }
}

// ref: Data.cpp:251-262
bool Data::readSample(const unsigned n) const
{
  if (!mJobbing) {
    return true;
  }
  return (n >= ((w_i - 1) * windowSize) / 2 && n < (w_i * windowSize) / 2) ||
         (n >= ((w_j - 1) * windowSize) / 2 && n < (w_j * windowSize) / 2) ||
         (jobs == jobInd && n >= ((w_j - 1) * windowSize) / 2);
}

namespace
{
constexpr char kCacheMagic[8] = {'F', 'S', 'M', 'C', 'B', 'I', 'T', '2'};

// what the cache depends on: the three input files (size and modification time) and the options that shape the matrix
struct CacheKey {
  long long size[3], mtime[3];
  int fold, csfs;
};
bool cacheKeyOf(const DecodingParams& params, CacheKey& key)
{
  const std::string files[3] = {hapsFile(params.inFileRoot), samplesFile(params.inFileRoot),
                                FileUtils::firstExisting(params.inFileRoot, {".map.gz", ".map"})};
  for (int i = 0; i < 3; ++i) {
    struct stat st;
    if (stat(files[i].c_str(), &st) != 0) {
      return false;
    }
    key.size[i] = static_cast<long long>(st.st_size);
    key.mtime[i] = static_cast<long long>(st.st_mtime);
  }
  key.fold = params.foldData;
  key.csfs = params.usingCSFS;
  return true;
}
template <class T> void writeVec(std::FILE* f, const std::vector<T>& v)
{
  const uint64_t n = v.size();
  std::fwrite(&n, sizeof n, 1, f);
  if (n) {
    std::fwrite(v.data(), sizeof(T), n, f);
  }
}
template <class T> bool readVec(std::FILE* f, std::vector<T>& v)
{
  uint64_t n = 0;
  if (std::fread(&n, sizeof n, 1, f) != 1 || n > (uint64_t{1} << 40)) {
    return false;
  }
  v.resize(n);
  return n == 0 || std::fread(v.data(), sizeof(T), n, f) == n;
}
bool cacheWanted(const DecodingParams& params)
{
  const char* e = std::getenv("FSMC_HAP_CACHE");
  return params.FastSMC && (params.hapBitCache || (e && *e && *e != '0'));
}
}  // namespace

std::string Data::bitCachePath(const std::string& inFileRoot)
{
  return hapsFile(inFileRoot) + ".fsmcbits";
}

void Data::writeBitCache(const DecodingParams& params) const
{
  CacheKey key{};
  if (!cacheKeyOf(params, key) || numLoadedHaplotypes() != haploidSampleSize) {
    return;
  }
  const std::string path = bitCachePath(params.inFileRoot), tmp = path + ".tmp";
  std::FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) {
    return;  // read-only data directory: no cache, no error
  }
  std::fwrite(kCacheMagic, 1, sizeof kCacheMagic, f);
  std::fwrite(&key, sizeof key, 1, f);
  const long long header[4] = {static_cast<long long>(sampleSize), sites, chrNumber, wordsPerHap};
  std::fwrite(header, sizeof header, 1, f);
  writeVec(f, hapBits);
  writeVec(f, flipMask);
  writeVec(f, totalSamplesCount);
  writeVec(f, derivedAlleleCounts);
  writeVec(f, physicalPositions);
  writeVec(f, geneticPositions);
  writeVec(f, recRateAtMarker);
  const bool ok = std::fflush(f) == 0 && !std::ferror(f);
  std::fclose(f);
  if (ok) {
    std::rename(tmp.c_str(), path.c_str());
  } else {
    std::remove(tmp.c_str());
  }
}

bool Data::loadBitCache(const DecodingParams& params, Data& whole)
{
  CacheKey want{}, have{};
  if (!cacheKeyOf(params, want)) {
    return false;
  }
  std::FILE* f = std::fopen(bitCachePath(params.inFileRoot).c_str(), "rb");
  if (!f) {
    return false;
  }
  struct Closer {
    std::FILE* f;
    ~Closer() { std::fclose(f); }
  } closer{f};
  char magic[sizeof kCacheMagic];
  long long header[4];
  if (std::fread(magic, 1, sizeof magic, f) != sizeof magic || std::memcmp(magic, kCacheMagic, sizeof magic) != 0 ||
      std::fread(&have, sizeof have, 1, f) != 1 || std::memcmp(&have, &want, sizeof want) != 0 ||
      std::fread(header, sizeof header, 1, f) != 1) {
    return false;  // another format, or the input files / options changed since it was written
  }
  Data d;
  d.sampleSize = static_cast<unsigned long>(header[0]);
  d.haploidSampleSize = 2ul * d.sampleSize;
  d.sites = static_cast<int>(header[1]);
  d.chrNumber = static_cast<int>(header[2]);
  d.wordsPerHap = static_cast<long>(header[3]);
  DecodingParams wholeParams = params;
  wholeParams.jobs = 1;
  wholeParams.jobInd = 1;
  d.setJobGeometry(wholeParams);
  if (!readVec(f, d.hapBits) || !readVec(f, d.flipMask) || !readVec(f, d.totalSamplesCount) || !readVec(f, d.derivedAlleleCounts) ||
      !readVec(f, d.physicalPositions) || !readVec(f, d.geneticPositions) || !readVec(f, d.recRateAtMarker) ||
      d.hapBits.size() != d.haploidSampleSize * static_cast<size_t>(d.wordsPerHap) || static_cast<int>(d.geneticPositions.size()) != d.sites) {
    return false;
  }
  d.siteWasFlippedDuringFolding.resize(static_cast<size_t>(d.sites));
  for (int s = 0; s < d.sites; ++s) {
    d.siteWasFlippedDuringFolding[static_cast<size_t>(s)] = (d.flipMask[static_cast<size_t>(s) >> 6] >> (s & 63)) & 1ull;
  }
  d.readSamplesList(params.inFileRoot);  // ids of every sample (jobs = 1)
  if (d.famAndIndNameList.size() != d.sampleSize) {
    return false;
  }
  d.globalHapId.resize(d.haploidSampleSize);
  for (size_t h = 0; h < d.globalHapId.size(); ++h) {
    d.globalHapId[h] = static_cast<uint32_t>(h);
  }
  whole = std::move(d);
  return true;
}

Data::Data(const DecodingParams& params)
{
  if (cacheWanted(params)) {
    // the packed matrix of the whole data set, from the cache or read once and cached; a job is cut out of it
    Data whole;
    if (!loadBitCache(params, whole)) {
      DecodingParams wholeParams = params;
      wholeParams.jobs = 1;
      wholeParams.jobInd = 1;
      whole.readFromFiles(wholeParams);
      whole.writeBitCache(params);
    } else {
      std::cout << "Read data for " << whole.haploidSampleSize << " haploid samples from " << bitCachePath(params.inFileRoot) << std::endl;
    }
    const bool jobbing = params.jobInd != -1 && params.jobs != -1 && params.jobs > 1;
    *this = jobbing ? forJob(whole, params) : std::move(whole);
    if (!jobbing) {
      setJobGeometry(params);
    }
    return;
  }
  readFromFiles(params);
}

void Data::readFromFiles(const DecodingParams& params)
{
  const std::string& root = params.inFileRoot;
  // `sites` is set by the reader itself (finishSites): counting the lines first would inflate the whole .hap.gz one
  // more time, and inflation is what bounds the read (the reference opens the file four times)
  sites = 0;
  sampleSize = static_cast<unsigned long>(countSamplesLines(root));
  haploidSampleSize = sampleSize * 2ul;
  setJobGeometry(params);
  if (!params.FastSMC) {
    mJobbing = false;  // ASMC loads every sample; jobs only slice the pair list (ref: Data.cpp:86-95, HMM.cpp:319-321)
  }
  readSamplesList(root);
  allocate();
  if (params.FastSMC) {
    readHapsFastSMC(root, readMapFastSMC(root));
  } else {
    readHapsAsmc(root);
    readMapAsmc(root);
  }
}

// Every job of a data set needs the same per-site information (allele counts are over the whole file) and the rows
// of its own two sample windows; runAllJobs reads the files once and cuts the jobs out of the result.
Data Data::forJob(const Data& whole, const DecodingParams& params)
{
  if (whole.numLoadedHaplotypes() != whole.haploidSampleSize) {
    throw std::runtime_error("Data::forJob: the source must hold every sample of the data set");
  }
  Data d;
  d.sampleSize = whole.sampleSize;
  d.haploidSampleSize = whole.haploidSampleSize;
  d.sites = whole.sites;
  d.geneticPositions = whole.geneticPositions;
  d.physicalPositions = whole.physicalPositions;
  d.siteWasFlippedDuringFolding = whole.siteWasFlippedDuringFolding;
  d.recRateAtMarker = whole.recRateAtMarker;
  d.chrNumber = whole.chrNumber;
  d.wordsPerHap = whole.wordsPerHap;
  d.flipMask = whole.flipMask;
  d.totalSamplesCount = whole.totalSamplesCount;
  d.derivedAlleleCounts = whole.derivedAlleleCounts;
  d.mUndistinguished = whole.mUndistinguished;  // same whole-file counts -> same draws
  d.setJobGeometry(params);
  if (d.foldToMinorAlleles != whole.foldToMinorAlleles) {
    throw std::runtime_error("Data::forJob: folding differs between the source and the job");
  }
  for (unsigned n = 0; n < whole.sampleSize; ++n) {
    if (d.readSample(n)) {
      d.FamIDList.push_back(whole.FamIDList[n]);
      d.IIDList.push_back(whole.IIDList[n]);
      d.famAndIndNameList.push_back(whole.famAndIndNameList[n]);
      d.globalHapId.push_back(2 * n);
      d.globalHapId.push_back(2 * n + 1);
    }
  }
  d.hapBits.resize(d.globalHapId.size() * static_cast<size_t>(d.wordsPerHap));
  for (size_t l = 0; l < d.globalHapId.size(); ++l) {
    std::memcpy(d.hapBits.data() + l * d.wordsPerHap, whole.hapBits.data() + static_cast<size_t>(d.globalHapId[l]) * d.wordsPerHap,
                sizeof(uint64_t) * static_cast<size_t>(d.wordsPerHap));
  }
  return d;
}

void Data::allocate()
{
  wordsPerHap = 0;
  hapBits.clear();
  flipMask.clear();
  siteWasFlippedDuringFolding.clear();
  totalSamplesCount.clear();
  derivedAlleleCounts.clear();
  mWordMajor.clear();
  mWordBufIndex = -1;
  geneticPositions.clear();
  physicalPositions.clear();
  recRateAtMarker.clear();
}

// ref: Data.cpp:212-249
void Data::readSamplesList(const std::string& inFileRoot)
{
  FileUtils::LineReader in(samplesFile(inFileRoot));
  std::string line;
  unsigned n = 0;
  while (in.next(line)) {
    const auto t = StringUtils::tokenizeMultipleDelimiters(line);
    if (isSamplesHeader(t) || t.size() < 2) {
      continue;
    }
    if (readSample(n)) {
      FamIDList.push_back(t[0]);
      IIDList.push_back(t[1]);
      famAndIndNameList.push_back(t[0] + "\t" + t[1]);
      globalHapId.push_back(2 * n);
      globalHapId.push_back(2 * n + 1);
    }
    ++n;
  }
  std::cout << "Read data for " << famAndIndNameList.size() * 2 << " haploid samples." << std::endl;
}

// ref: Data.cpp:98-141
std::vector<std::pair<unsigned long, double>> Data::readMapFastSMC(const std::string& inFileRoot)
{
  const std::string f = FileUtils::firstExisting(inFileRoot, {".map.gz", ".map"});
  if (f.empty()) {
    std::cerr << "ERROR. Could not find hap file in " + inFileRoot + ".map.gz or " + inFileRoot + ".map" << std::endl;
    exit(1);
  }
  FileUtils::LineReader in(f);
  std::vector<std::pair<unsigned long, double>> gmap;
  std::string line;
  while (in.next(line)) {
    std::istringstream ss(line);
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
  return gmap;
}

// One site's alleles -> packed bits, counts and folding (ref: Data.cpp:449-503).  `alleles` points at the text
// after the five meta columns: for haplotype h the allele character is alleles[2*h + 1].
//
// Sites arrive in ascending order.  The bits of the 64 sites of the current word are collected in mWordBuf (one
// uint64 per loaded haplotype, contiguous) and written to the hap-major matrix once per word: setting one bit per
// site directly in hapBits touches one cache line per haplotype per SITE (10^9 scattered read-modify-writes at
// 10 000 samples x 50 000 SNPs, which was most of the read time); this way it is one per haplotype per WORD.
void Data::addSite(const int pos, const char* alleles, const unsigned long nHapsInFile, const bool subset)
{
  int derived = 0;
  unsigned bad = 0;
  for (unsigned long h = 0; h < nHapsInFile; ++h) {
    const char c = alleles[2 * h + 1];
    bad |= static_cast<unsigned>(c != '0') & static_cast<unsigned>(c != '1');
    derived += (c == '1');
  }
  if (bad) {
    std::cerr << "ERROR: hap is not '0' or '1'" << std::endl;
    exit(1);
  }
  const int total = static_cast<int>(nHapsInFile);
  const bool minorIsOne = foldToMinorAlleles ? (derived <= total - derived) : true;
  const int shift = pos & 63;
  const long w = pos >> 6;
  if (static_cast<size_t>(pos) >= totalSamplesCount.size()) {
    siteWasFlippedDuringFolding.resize(pos + 1, false);
    totalSamplesCount.resize(pos + 1, 0);
    derivedAlleleCounts.resize(pos + 1, 0);
  }
  if (static_cast<size_t>(w) >= flipMask.size()) {
    flipMask.resize(w + 1, 0ull);
  }
  siteWasFlippedDuringFolding[pos] = !minorIsOne;
  if (!minorIsOne) {
    flipMask[w] |= 1ull << shift;
  }
  if (w != mWordBufIndex) {
    flushSiteWord();
    mWordBufIndex = w;
    mWordBuf.assign(numLoadedHaplotypes(), 0ull);
  }
  const char minorChar = minorIsOne ? '1' : '0';
  uint64_t* buf = mWordBuf.data();
  if (subset) {
    const size_t n = globalHapId.size();
    const uint32_t* ids = globalHapId.data();
    for (size_t l = 0; l < n; ++l) {
      buf[l] |= static_cast<uint64_t>(alleles[2 * static_cast<size_t>(ids[l]) + 1] == minorChar) << shift;
    }
  } else {
    for (unsigned long h = 0; h < nHapsInFile; ++h) {
      buf[h] |= static_cast<uint64_t>(alleles[2 * h + 1] == minorChar) << shift;
    }
  }
  totalSamplesCount[pos] = total;
  derivedAlleleCounts[pos] = foldToMinorAlleles ? std::min(derived, total - derived) : derived;
}

// Appends the collected word of every loaded haplotype to the word-major store.  Called when the next word starts and
// by finishSites.
void Data::flushSiteWord()
{
  if (mWordBufIndex < 0) {
    return;
  }
  if (static_cast<size_t>(mWordBufIndex) * mWordBuf.size() != mWordMajor.size()) {
    std::cerr << "ERROR: sites must be added in ascending order" << std::endl;
    exit(1);
  }
  mWordMajor.insert(mWordMajor.end(), mWordBuf.begin(), mWordBuf.end());
  mWordBufIndex = -1;
}

// After the last site: fixes the site count and turns the word-major store [word][hap] into the hap-major matrix
// [hap][word] that the decoder uploads.
void Data::finishSites(const int numSites)
{
  flushSiteWord();
  sites = numSites;
  wordsPerHap = (sites + 63) / 64;
  const size_t H = numLoadedHaplotypes();
  hapBits.assign(H * wordsPerHap, 0ull);
  flipMask.resize(wordsPerHap, 0ull);
  siteWasFlippedDuringFolding.resize(sites, false);
  totalSamplesCount.resize(sites, 0);
  derivedAlleleCounts.resize(sites, 0);
  const size_t W = H ? mWordMajor.size() / H : 0;
  constexpr size_t kTile = 64;  // transpose in tiles: both sides stay within a few cache lines per row
  for (size_t w0 = 0; w0 < W; w0 += kTile) {
    const size_t w1 = std::min(W, w0 + kTile);
    for (size_t l = 0; l < H; ++l) {
      uint64_t* dst = hapBits.data() + l * wordsPerHap;
      for (size_t w = w0; w < w1; ++w) {
        dst[w] = mWordMajor[w * H + l];
      }
    }
  }
  std::vector<uint64_t>().swap(mWordMajor);
  std::vector<uint64_t>().swap(mWordBuf);
}

// ref: Data.cpp:523-565 (readGeneticMap + addMarker): linear interpolation of the map at bp
void Data::addMarker(const int pos, const unsigned long bp, const std::vector<std::pair<unsigned long, double>>& gmap,
                     unsigned& g)
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
  geneticPositions.push_back(static_cast<float>(cm / 100.f));
  physicalPositions.push_back(static_cast<int>(bp));
  if (pos > 0) {
    const double gd = geneticPositions[pos] - geneticPositions[pos - 1];
    const unsigned long pd = physicalPositions[pos] - physicalPositions[pos - 1];
    const float rate = static_cast<float>(gd / pd);
    if (pos == 1) {
      recRateAtMarker.push_back(rate);
    }
    recRateAtMarker.push_back(rate);
  }
}

namespace
{
// Splits the five meta columns off a haps line; returns the offset of the allele text.
size_t parseHapsMeta(const std::string& line, std::string& chr, unsigned long& bp)
{
  size_t p = 0;
  std::string tok[5];
  for (int f = 0; f < 5; ++f) {
    while (p < line.size() && (line[p] == ' ' || line[p] == '\t')) {
      ++p;
    }
    const size_t b = p;
    while (p < line.size() && line[p] != ' ' && line[p] != '\t') {
      ++p;
    }
    tok[f] = line.substr(b, p - b);
  }
  chr = tok[0];
  bp = std::stoul(tok[2]);
  return p;
}
}  // namespace

// ref: Data.cpp:397-515
void Data::readHapsFastSMC(const std::string& inFileRoot, const std::vector<std::pair<unsigned long, double>>& gmap)
{
  if (gmap.empty()) {
    std::cerr << "ERROR: genetic map is empty" << std::endl;
    exit(1);
  }
  FileUtils::LineReader in(hapsFile(inFileRoot));
  std::string line, chr;
  unsigned g = 0;
  unsigned long lastBp = 0;
  int pos = 0;
  while (in.next(line)) {
    if (line.empty()) {
      break;
    }
    unsigned long bp = 0;
    const size_t off = parseHapsMeta(line, chr, bp);
    const size_t rest = line.size() - off;
    if (!(rest == 4ul * sampleSize || rest == 4ul * sampleSize + 1)) {
      std::cerr << "ERROR: haps line has wrong length. Length is " << rest << ", should be 4*" << sampleSize << "."
                << std::endl;
      exit(1);
    }
    if (bp <= lastBp) {
      std::cerr << "ERROR: hap file must be sorted by increasing physical position." << std::endl;
      exit(1);
    }
    lastBp = bp;
    if (pos == 0) {
      const size_t colon = chr.find(':');
      try {
        chrNumber = std::stoi(colon == std::string::npos ? chr : chr.substr(0, colon));
      } catch (const std::exception&) {
        chrNumber = 0;
      }
      if (chrNumber <= 0 || chrNumber > 1260) {
        chrNumber = 0;
      }
    }
    addMarker(pos, bp, gmap, g);
    addSite(pos, line.c_str() + off, haploidSampleSize, true);
    ++pos;
  }
  finishSites(pos);
  std::cout << "Read " << pos << " markers" << std::endl;
}

// ref: Data.cpp:318-395 — ASMC mode: every sample is loaded, counts are over the loaded haplotypes
void Data::readHapsAsmc(const std::string& inFileRoot)
{
  FileUtils::LineReader in(hapsFile(inFileRoot));
  std::string line, chr;
  int pos = 0;
  while (in.next(line)) {
    if (line.empty()) {
      break;
    }
    unsigned long bp = 0;
    const size_t off = parseHapsMeta(line, chr, bp);
    if (line.size() - off < 2ul * haploidSampleSize) {
      std::cerr << "ERROR: haps line has wrong length." << std::endl;
      exit(1);
    }
    addSite(pos, line.c_str() + off, haploidSampleSize, false);
    ++pos;
  }
  finishSites(pos);
}

// ref: Data.cpp:162-210 — PLINK-style map: chr, snp, cM, bp
void Data::readMapAsmc(const std::string& inFileRoot)
{
  const std::string f = FileUtils::firstExisting(inFileRoot, {".map.gz", ".map"});
  if (f.empty()) {
    std::cerr << "ERROR. Could not find map file in " + inFileRoot + ".map.gz or " + inFileRoot + ".map" << std::endl;
    exit(1);
  }
  FileUtils::LineReader in(f);
  geneticPositions.assign(sites, 0.f);
  physicalPositions.assign(sites, 0);
  recRateAtMarker.assign(sites, 0.f);
  std::string line;
  int pos = 0;
  while (in.next(line) && pos < sites) {
    const auto t = StringUtils::tokenizeMultipleDelimiters(line);
    if (t.size() < 4) {
      continue;
    }
    geneticPositions[pos] = StringUtils::stof(t[2]) / 100.f;
    physicalPositions[pos] = std::stoi(t[3]);
    if (pos > 0) {
      recRateAtMarker[pos] = (geneticPositions[pos] - geneticPositions[pos - 1]) /
                             static_cast<float>(physicalPositions[pos] - physicalPositions[pos - 1]);
    }
    ++pos;
  }
}

Data Data::fromArrays(const DecodingParams& params, const std::vector<std::string>& famIds,
                      const std::vector<std::string>& iids, const uint8_t* raw, const long numHaps, const int numSites,
                      const std::vector<int>& physPos, const std::vector<double>& cM, const int chr)
{
  Data d;
  d.sites = numSites;
  d.sampleSize = static_cast<unsigned long>(numHaps / 2);
  d.haploidSampleSize = static_cast<unsigned long>(numHaps);
  d.setJobGeometry(params);
  for (unsigned n = 0; n < d.sampleSize; ++n) {
    if (d.readSample(n)) {
      d.FamIDList.push_back(famIds[n]);
      d.IIDList.push_back(iids[n]);
      d.famAndIndNameList.push_back(famIds[n] + "\t" + iids[n]);
      d.globalHapId.push_back(2 * n);
      d.globalHapId.push_back(2 * n + 1);
    }
  }
  d.allocate();
  d.chrNumber = chr;
  // the same interpolation code path as the file reader, with the data's own sites as the map
  std::vector<std::pair<unsigned long, double>> gmap(numSites);
  for (int s = 0; s < numSites; ++s) {
    gmap[s] = {static_cast<unsigned long>(physPos[s]), cM[s]};
  }
  std::string text(2 * static_cast<size_t>(numHaps), ' ');
  unsigned g = 0;
  for (int s = 0; s < numSites; ++s) {
    for (long h = 0; h < numHaps; ++h) {
      text[2 * h + 1] = raw[static_cast<size_t>(h) * numSites + s] ? '1' : '0';
    }
    d.addMarker(s, static_cast<unsigned long>(physPos[s]), gmap, g);
    d.addSite(s, text.c_str(), d.haploidSampleSize, true);
  }
  d.finishSites(numSites);
  return d;
}

Individual Data::individual(const unsigned long i) const
{
  Individual ind(sites);
  for (int s = 0; s < sites; ++s) {
    ind.genotype1[s] = allele(2 * i, s);
    ind.genotype2[s] = allele(2 * i + 1, s);
  }
  return ind;
}

// ref: Data.cpp:144-160 (sampleHypergeometric) and 567-599.  The reference seeds one std::mt19937 per draw from
// std::rand(); the sequence of std::rand() calls must match the reference's exactly, so the seeds are drawn serially,
// in the reference's order.  The shuffles themselves only depend on their seed and run on all host threads (at UK
// Biobank scale they are 3 x sites shuffles of ~10^6 shorts).
std::vector<std::vector<int>> Data::calculateUndistinguishedCounts(const int numCsfsSamples) const
{
  UndistinguishedCache& cache = *mUndistinguished;
  std::lock_guard<std::mutex> g(cache.lock);
  if (cache.csfsSamples != numCsfsSamples || cache.knownSeed != mUseKnownSeed) {
    cache.counts = drawUndistinguishedCounts(numCsfsSamples);
    cache.csfsSamples = numCsfsSamples;
    cache.knownSeed = mUseKnownSeed;
  }
  return cache.counts;
}

std::shared_ptr<const void> Data::sharedObject(const std::string& key,
                                               const std::function<std::shared_ptr<const void>()>& make) const
{
  UndistinguishedCache& cache = *mUndistinguished;
  std::lock_guard<std::mutex> g(cache.objectsLock);
  auto it = cache.objects.find(key);
  if (it == cache.objects.end()) {
    it = cache.objects.emplace(key, make()).first;
  }
  return it->second;
}

std::vector<std::vector<int>> Data::drawUndistinguishedCounts(const int numCsfsSamples) const
{
  std::vector<std::vector<int>> counts(sites, std::vector<int>(3, 0));
  struct Draw {
    int site, dist, population, successes;
    unsigned seed;
  };
  std::vector<Draw> draws;
  draws.reserve(static_cast<size_t>(sites) * 3);
  static std::mutex randMutex;  // std::rand's state is process-wide
  std::unique_lock<std::mutex> randLock(randMutex);
  if (mUseKnownSeed) {
    std::srand(1234u);
  } else {
    std::random_device rd;
    std::srand(rd());
  }
  for (int s = 0; s < sites; ++s) {
    const int total = totalSamplesCount[s];
    const int derived = derivedAlleleCounts[s];
    if (decodingUsesCSFS && numCsfsSamples > total) {
      std::cerr << "ERROR. The number of CSFS samples (" << numCsfsSamples
                << ") is larger than the number of samples in the data (" << total << ")." << std::endl;
      exit(1);
    }
    for (int dist = 0; dist < 3; ++dist) {
      const int population = total - 2;
      const int successes = derived - dist;
      if (successes >= 0 && successes <= population) {
        draws.push_back(Draw{s, dist, population, successes, static_cast<unsigned>(std::rand())});
      } else {
        counts[s][dist] = -1;
      }
    }
  }
  randLock.unlock();
  std::atomic<size_t> next{0};
  auto work = [&] {
    std::vector<unsigned short> urn;
    constexpr size_t kGrain = 16;
    for (size_t lo = next.fetch_add(kGrain); lo < draws.size(); lo = next.fetch_add(kGrain)) {
      const size_t hi = std::min(draws.size(), lo + kGrain);
      for (size_t i = lo; i < hi; ++i) {
        const Draw& d = draws[i];
        urn.assign(d.population, 0);
        std::fill(urn.begin(), urn.begin() + d.successes, 1);
        std::shuffle(urn.begin(), urn.end(), std::mt19937(d.seed));
        counts[d.site][d.dist] = std::accumulate(urn.begin(), urn.begin() + (numCsfsSamples - 2), 0);
      }
    }
  };
  const size_t cost = draws.empty() ? 0 : draws.size() * static_cast<size_t>(std::max(1, draws[0].population));
  const unsigned nThreads = cost < (size_t{1} << 24) ? 1u : std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nThreads; ++t) {
    pool.emplace_back(work);
  }
  work();
  for (auto& th : pool) {
    th.join();
  }
  for (int s = 0; s < sites; ++s) {
    for (int dist = 0; dist < 3; ++dist) {
      int sample = counts[s][dist];
      if (foldToMinorAlleles && (sample + dist > numCsfsSamples / 2)) {
        sample = numCsfsSamples - 2 - sample;
      }
      counts[s][dist] = sample;
    }
  }
  return counts;
}
