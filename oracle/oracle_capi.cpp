// TEST INFRASTRUCTURE ONLY — C entry points over the CPU oracle for ctypes (tests/, smoke(), bench.py
// cpu_baseline).  See oracle.hpp for the scope rules.
#include "oracle.hpp"

#include <atomic>
#include <xmmintrin.h>
#include <thread>

#include <cstring>
#include <exception>
#include <string>

namespace
{
thread_local std::string gError;

template <class F> int guarded(F&& f)
{
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    gError = e.what();
    return -1;
  }
}
}  // namespace

extern "C" {

struct fo_params {
  const char* inFileRoot;
  const char* decodingQuantFile;
  const char* outFileRoot;
  int jobs, jobInd;
  int foldData, usingCSFS;
  float skipCSFSdistance;
  int batchSize;
  float skip;
  int gap, max_seeds;
  float min_m;
  int hashing, FastSMC, BIN_OUT, useKnownSeed, outputIbdSegmentLength;
  int time;
  int noConditionalAgeEstimates, doPerPairPosteriorMean, doPerPairMAP, withinOnly;
  int shuffleFlavor;
  int simdFlavor;
  int asmcMode;  // load with the ASMC (PLINK map, all samples) reader instead of the FastSMC one
};

const char* fo_last_error()
{
  return gError.c_str();
}

void* fo_create(const fo_params* c)
{
  fo::Oracle* o = nullptr;
  const int rc = guarded([&] {
    fo::Params p;
    p.inFileRoot = c->inFileRoot;
    p.decodingQuantFile = c->decodingQuantFile;
    p.outFileRoot = c->outFileRoot ? c->outFileRoot : "";
    p.jobs = c->jobs;
    p.jobInd = c->jobInd;
    p.foldData = c->foldData;
    p.usingCSFS = c->usingCSFS;
    p.skipCSFSdistance = c->skipCSFSdistance;
    p.batchSize = c->batchSize;
    p.skip = c->skip;
    p.gap = c->gap;
    p.max_seeds = c->max_seeds;
    p.min_m = c->min_m;
    p.hashing = c->hashing;
    p.FastSMC = c->FastSMC;
    p.BIN_OUT = c->BIN_OUT;
    p.useKnownSeed = c->useKnownSeed;
    p.outputIbdSegmentLength = c->outputIbdSegmentLength;
    p.time = c->time;
    p.noConditionalAgeEstimates = c->noConditionalAgeEstimates;
    p.doPerPairPosteriorMean = c->doPerPairPosteriorMean;
    p.doPerPairMAP = c->doPerPairMAP;
    p.withinOnly = c->withinOnly;
    p.shuffleFlavor = c->shuffleFlavor;
    p.simdFlavor = c->simdFlavor != 0;
    o = new fo::Oracle(p, c->asmcMode != 0);
  });
  return rc == 0 ? o : nullptr;
}

void fo_destroy(void* h)
{
  delete static_cast<fo::Oracle*>(h);
}

// scalar facts: 0 sites, 1 states, 2 haplotypes in job, 3 stateThreshold, 4 ageThreshold, 5 chr,
// 6 windowSize, 7 w_i, 8 w_j, 9 aboveDiag, 10 total diploid samples, 11 csfsSamples
long fo_info(void* h, int what)
{
  auto* o = static_cast<fo::Oracle*>(h);
  switch (what) {
  case 0: return o->sites;
  case 1: return o->dq.states;
  case 2: return static_cast<long>(o->hap.size());
  case 3: return o->stateThreshold;
  case 4: return o->ageThreshold;
  case 5: return o->chrNumber;
  case 6: return o->windowSize;
  case 7: return o->w_i;
  case 8: return o->w_j;
  case 9: return o->aboveDiag;
  case 10: return o->sampleSizeTotal;
  case 11: return o->dq.csfsSamples;
  default: return -1;
  }
}

float fo_probability_threshold(void* h)
{
  return static_cast<fo::Oracle*>(h)->probabilityThreshold;
}

void fo_get_emissions(void* h, float* e1, float* e0m1, float* e2m0)
{
  auto* o = static_cast<fo::Oracle*>(h);
  std::memcpy(e1, o->e1.data(), o->e1.size() * sizeof(float));
  std::memcpy(e0m1, o->e0m1.data(), o->e0m1.size() * sizeof(float));
  std::memcpy(e2m0, o->e2m0.data(), o->e2m0.size() * sizeof(float));
}

void fo_get_undistinguished(void* h, int* out)
{
  auto* o = static_cast<fo::Oracle*>(h);
  for (int s = 0; s < o->sites; ++s) {
    for (int d = 0; d < 3; ++d) {
      out[3 * s + d] = o->undistinguished[s][d];
    }
  }
}

void fo_get_positions(void* h, float* gen, int* phys)
{
  auto* o = static_cast<fo::Oracle*>(h);
  std::memcpy(gen, o->genPos.data(), o->genPos.size() * sizeof(float));
  std::memcpy(phys, o->physPos.data(), o->physPos.size() * sizeof(int));
}

void fo_get_hap(void* h, int hap, unsigned char* out)
{
  auto* o = static_cast<fo::Oracle*>(h);
  std::memcpy(out, o->hap[hap].data(), o->hap[hap].size());
}

void fo_get_flipped(void* h, unsigned char* out)
{
  auto* o = static_cast<fo::Oracle*>(h);
  std::memcpy(out, o->flipped.data(), o->flipped.size());
}

// vectors of the model: 0 initialStateProb, 1 expectedTimes, 2 columnRatios, 3 discretization(S+1)
void fo_get_vector(void* h, int which, float* out)
{
  auto* o = static_cast<fo::Oracle*>(h);
  const std::vector<float>* v = which == 0   ? &o->dq.initialStateProb
                                : which == 1 ? &o->dq.expectedTimes
                                : which == 2 ? &o->dq.columnRatios
                                             : &o->dq.discretization;
  std::memcpy(out, v->data(), v->size() * sizeof(float));
}

// transition rows for one distance key: out = [D | B | U | RR], each S floats; returns -1 if absent
int fo_get_transition(void* h, float key, float* out)
{
  auto* o = static_cast<fo::Oracle*>(h);
  const int S = o->dq.states;
  if (!o->dq.D.count(key)) {
    return -1;
  }
  std::memcpy(out, o->dq.D.at(key).data(), S * sizeof(float));
  std::memcpy(out + S, o->dq.B.at(key).data(), S * sizeof(float));
  std::memcpy(out + 2 * S, o->dq.U.at(key).data(), S * sizeof(float));
  std::memcpy(out + 3 * S, o->dq.RR.at(key).data(), S * sizeof(float));
  return 0;
}


// Decodes the first nPairs pairs of the job's all-pairs enumeration (ref: HMM.cpp:311-357) in reference batches over the
// whole sequence, segment calling included, on `threads` host threads; returns the pair-sites processed (the CPU
// baseline of bench.py).  Enumeration of the whole job is avoided: only the needed prefix is generated.
double fo_decode_sample(void* h, long nPairs, int threads)
{
  double pairSites = -1.0;
  guarded([&] {
    auto* o = static_cast<fo::Oracle*>(h);
    std::vector<fo::PairObs> pairs;
    const unsigned N = static_cast<unsigned>(o->famId.size());
    for (unsigned i = 0; i < N && static_cast<long>(pairs.size()) < nPairs; ++i) {
      for (unsigned j = 0; j < i; ++j) {
        for (int iHap = 1; iHap <= 2; ++iHap) {
          for (int jHap = 1; jHap <= 2; ++jHap) {
            pairs.push_back(fo::PairObs{jHap, j, iHap, i});
          }
        }
      }
      pairs.push_back(fo::PairObs{1, i, 2, i});
    }
    if (static_cast<long>(pairs.size()) > nPairs) {
      pairs.resize(static_cast<size_t>(nPairs));
    }
    const std::vector<fo::Batch> batches = o->makeBatches(pairs, nullptr);
    std::atomic<long> next{0};
    std::vector<size_t> found(std::max(1, threads), 0);
    auto worker = [&](const int t) {
      std::vector<float> posterior;
      std::vector<fo::Segment> segs;
      for (long b = next++; b < static_cast<long>(batches.size()); b = next++) {
        o->decodeBatch(batches[b].pairs, batches[b].from, batches[b].to, posterior);
        segs.clear();
        o->callSegments(batches[b], static_cast<uint32_t>(b), posterior, segs);
        found[t] += segs.size();
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
      pool.emplace_back(worker, t);
    }
    worker(0);
    for (auto& t : pool) {
      t.join();
    }
    pairSites = 0.0;
    for (const auto& b : batches) {
      pairSites += static_cast<double>(b.pairs.size()) * (b.to - b.from);
    }
  });
  return pairSites;
}

// timing experiment only: flush-to-zero / denormals-are-zero on the calling thread
void fo_set_ftz(int on)
{
  unsigned csr = _mm_getcsr();
  csr = on ? (csr | 0x8040u) : (csr & ~0x8040u);
  _mm_setcsr(csr);
}

// whole FastSMC::run; returns #records or -1
long fo_run(void* h, const char* outPath, int threads)
{
  long n = -1;
  guarded([&] { n = static_cast<fo::Oracle*>(h)->run(outPath ? outPath : "", threads); });
  return n;
}

double fo_last_decode_seconds(void* h)
{
  return static_cast<fo::Oracle*>(h)->lastDecodeSeconds;
}
double fo_last_pair_sites(void* h)
{
  return static_cast<fo::Oracle*>(h)->lastPairSites;
}

long fo_num_candidates(void* h)
{
  return static_cast<long>(static_cast<fo::Oracle*>(h)->lastCandidates.size());
}
void fo_get_candidates(void* h, unsigned* out /* [n][4] */)
{
  auto* o = static_cast<fo::Oracle*>(h);
  for (size_t i = 0; i < o->lastCandidates.size(); ++i) {
    out[4 * i + 0] = o->lastCandidates[i].hapA;
    out[4 * i + 1] = o->lastCandidates[i].hapB;
    out[4 * i + 2] = o->lastCandidates[i].from;
    out[4 * i + 3] = o->lastCandidates[i].to;
  }
}
// seeding only (no decode); fills lastCandidates
long fo_seed(void* h)
{
  long n = -1;
  guarded([&] {
    auto* o = static_cast<fo::Oracle*>(h);
    o->lastCandidates = o->seedCandidates();
    n = static_cast<long>(o->lastCandidates.size());
  });
  return n;
}

long fo_num_batches(void* h)
{
  return static_cast<long>(static_cast<fo::Oracle*>(h)->lastBatches.size());
}
// per batch: nPairs, scanFrom, scanTo, from, to
void fo_get_batches(void* h, unsigned* out /* [n][5] */)
{
  auto* o = static_cast<fo::Oracle*>(h);
  for (size_t i = 0; i < o->lastBatches.size(); ++i) {
    const auto& b = o->lastBatches[i];
    out[5 * i + 0] = static_cast<unsigned>(b.pairs.size());
    out[5 * i + 1] = b.scanFrom;
    out[5 * i + 2] = b.scanTo;
    out[5 * i + 3] = b.from;
    out[5 * i + 4] = b.to;
  }
}

long fo_num_segments(void* h)
{
  return static_cast<long>(static_cast<fo::Oracle*>(h)->lastSegments.size());
}
// per segment ints: batch, lane, hapA, hapB, posStart, posEnd, mapState ; floats: prob, postMean, map
void fo_get_segments(void* h, int* ints /* [n][7] */, float* floats /* [n][3] */)
{
  auto* o = static_cast<fo::Oracle*>(h);
  for (size_t i = 0; i < o->lastSegments.size(); ++i) {
    const auto& s = o->lastSegments[i];
    ints[7 * i + 0] = static_cast<int>(s.batch);
    ints[7 * i + 1] = static_cast<int>(s.lane);
    ints[7 * i + 2] = static_cast<int>(o->hapIndex(s.obs, false));
    ints[7 * i + 3] = static_cast<int>(o->hapIndex(s.obs, true));
    ints[7 * i + 4] = static_cast<int>(s.posStart);
    ints[7 * i + 5] = static_cast<int>(s.posEnd);
    ints[7 * i + 6] = s.mapState;
    floats[3 * i + 0] = s.prob;
    floats[3 * i + 1] = s.postMean;
    floats[3 * i + 2] = s.map;
  }
}

// Decode explicit haplotype pairs over [from,to) as ONE batch; posterior out is [n][to-from][S]
// (pair-major, transposed from the internal layout for convenience).
int fo_decode_posterior(void* h, int n, const unsigned* hapA, const unsigned* hapB, unsigned from, unsigned to,
                        float* out)
{
  return guarded([&] {
    auto* o = static_cast<fo::Oracle*>(h);
    std::vector<fo::PairObs> pairs(n);
    for (int i = 0; i < n; ++i) {
      pairs[i] = fo::PairObs{static_cast<int>(hapA[i] % 2 + 1), hapA[i] / 2, static_cast<int>(hapB[i] % 2 + 1),
                             hapB[i] / 2};
    }
    std::vector<float> post;
    o->decodeBatch(pairs, from, to, post);
    const int S = o->dq.states;
    const size_t len = to - from;
    for (int v = 0; v < n; ++v) {
      for (size_t p = 0; p < len; ++p) {
        for (int k = 0; k < S; ++k) {
          out[(static_cast<size_t>(v) * len + p) * S + k] = post[(p * S + k) * n + v];
        }
      }
    }
  });
}

// Per-site posterior mean / MAP and IBD probability for explicit pairs over [from,to).
// mean, ibd: [n][to-from] floats; map: [n][to-from] ints.  Any output may be null.
int fo_decode_summary(void* h, int n, const unsigned* hapA, const unsigned* hapB, unsigned from, unsigned to,
                      float* mean, int* map, float* ibd)
{
  return guarded([&] {
    auto* o = static_cast<fo::Oracle*>(h);
    const int S = o->dq.states;
    const size_t len = to - from;
    const int chunk = 32;
    for (int base = 0; base < n; base += chunk) {
      const int m = std::min(chunk, n - base);
      std::vector<fo::PairObs> pairs(m);
      for (int i = 0; i < m; ++i) {
        const unsigned a = hapA[base + i], b = hapB[base + i];
        pairs[i] = fo::PairObs{static_cast<int>(a % 2 + 1), a / 2, static_cast<int>(b % 2 + 1), b / 2};
      }
      std::vector<float> post;
      o->decodeBatch(pairs, from, to, post);
      o->perSiteSummary(post, m, static_cast<unsigned>(len), mean ? mean + static_cast<size_t>(base) * len : nullptr,
                        map ? map + static_cast<size_t>(base) * len : nullptr);
      if (ibd) {
        for (int v = 0; v < m; ++v) {
          for (size_t p = 0; p < len; ++p) {
            float s = 0.f;
            for (unsigned k = 0; k < o->stateThreshold; ++k) {
              s += post[(p * S + k) * m + v];
            }
            ibd[(static_cast<size_t>(base) + v) * len + p] = s;
          }
        }
      }
    }
  });
}

// KAT entry points
float fo_round_morgans(float v, int precision, float minv)
{
  return fo::roundMorgans(v, precision, minv);
}
int fo_round_physical(int v, int precision)
{
  return fo::roundPhysical(v, precision);
}
unsigned fo_get_from_position(const float* gen, unsigned n, unsigned from, float cm)
{
  return fo::getFromPosition(std::vector<float>(gen, gen + n), from, cm);
}
unsigned fo_get_to_position(const float* gen, unsigned n, unsigned to, float cm)
{
  return fo::getToPosition(std::vector<float>(gen, gen + n), to, cm);
}
double fo_cm_between(int w1, int w2, const float* gen, unsigned n, int wordSize)
{
  return fo::cmBetween(w1, w2, std::vector<float>(gen, gen + n), wordSize);
}

}  // extern "C"
