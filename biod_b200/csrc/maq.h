// MAQ genotype likelihoods over pileup columns: the model's tables on the host, the per-column
// kernel's interface.
//
// Replaces, for the pileup columns the GPU has just built, MaqSnpCaller.genotypeLikelihoodInfo / makeCall / findSNPs
// (bio/std/hts/snpcallers/maq.d:388-540) over ErrorModelCoefficients.computeLikelihoods (:138-248).  The coefficient
// tables (:66-132) are a one-time setup the reference also does on the CPU, in x87 `real` arithmetic; they are computed
// here the same way (long double) and uploaded — 64 x 256 x 256 doubles of beta, 256 x 256 of lhet, 256 of fk.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

namespace biodb {

struct MaqTablesHost {
  std::vector<double> fk, beta, lhet;
  // ErrorModelCoefficients(depcorr, eta), maq.d:80-120
  MaqTablesHost(double depcorr, double eta) : fk(256), beta((size_t)64 * 256 * 256, 0.0), lhet((size_t)256 * 256, 0.0) {
    for (int n = 0; n < 256; ++n) fk[n] = pow(1.0 - depcorr, (double)n) * (1.0 - eta) + eta;
    std::vector<double> lC((size_t)256 * 256, 0.0), lG(256);
    const long double ln2 = logl(2.0L), ln10 = logl(10.0L);
    for (int n = 0; n < 256; ++n) {
      lG[n] = lgamma((double)(n + 1));
      for (int k = 0; k <= n / 2; ++k) {
        const double v = lG[n] - lG[k] - lG[n - k];                     // log C(n, k)
        lC[(size_t)n * 256 + k] = lC[(size_t)n * 256 + (n - k)] = v;
        lhet[(size_t)n << 8 | (size_t)k] = lhet[(size_t)n << 8 | (size_t)(n - k)] = v - (double)n * (double)ln2;
      }
    }
    for (int q = 1; q < 64; ++q) {
      const long double e = powl(10.0L, -(long double)q / 10.0L);       // error probability of a base of quality q
      const long double le = logl(e), le1 = logl(1.0L - e);
      for (int n = 1; n < 256; ++n) {
        long double tail = 0.0L;                                         // P(X >= k + 1), X ~ Binomial(n, e)
        for (int k = n; k >= 0; --k) {
          const long double with_k = tail + expl((long double)lC[(size_t)n * 256 + k] + k * le + (n - k) * le1);
          beta[(size_t)q << 16 | (size_t)n << 8 | (size_t)k] = (double)(-10.0L / ln10 * logl(tail / with_k));
          tail = with_k;
        }
      }
    }
  }
};

struct MaqParams {           // MaqSnpCaller's knobs (maq.d:327-380)
  float depcorr = 0.17f, eta = 0.03f, minimum_call_quality = 6.0f;
  int32_t minimum_base_quality = 13;
};

struct MaqDevTables {        // device pointers
  const double* fk;
  const double* beta;
  const double* lhet;
};

struct MaqColumns {          // per column, device
  uint8_t* gt0;              // best genotype, DiploidGenotype!Base5 code (first * 5 + second), 255 = no valid base
  uint8_t* gt1;              // second best
  float* s0;                 // their scores
  float* s1;
  uint16_t* n_valid;         // bases that passed the filters (before the cap of 255)
};

// entries of MAQ mode (written by the entries kernel): base | strand << 7 (0xFF = filtered out), min(quality, mapq)
void maq_columns(const uint64_t* col_off, const uint8_t* base_s, const uint8_t* qual_m, uint32_t n_col, MaqDevTables t,
                 MaqColumns out, cudaStream_t st);
// findSNPs' filter (maq.d:516-521): flag[c] = 1 where a call exists, differs from the reference genotype and its
// quality exceeds min_q.  ref_base may be null ('N' everywhere).
void maq_call_flags(const MaqColumns& m, const uint8_t* ref_base, uint32_t n_col, float min_q, uint32_t* flag, cudaStream_t st);
void maq_call_gather(const MaqColumns& m, const uint8_t* ref_base, const uint64_t* col_pos, uint32_t n_col, const uint32_t* flag,
                     const uint32_t* incl, uint32_t* call_col, uint64_t* call_pos, uint8_t* call_gt, uint8_t* call_ref,
                     float* call_qual, cudaStream_t st);

}  // namespace biodb
