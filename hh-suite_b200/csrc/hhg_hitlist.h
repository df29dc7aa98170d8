// hh-suite_b200/csrc/hhg_hitlist.h -- hit-list statistics of the reference, host side of the library.
//
// Replaces, for the hits the GPU path produces:
//   HitList::CalculatePvalues          /root/reference/src/hhhitlist.cpp:499-531   (per-hit EVD parameters from the
//     neural-network regression lamda_NN / mu_NN, src/hhhitlist-inl.h:14-69; Pvalue/logPvalue src/hhhit-inl.h:44-53;
//     Hit::CalcEvalScoreProbab src/hhhit.h:134-141 with CalcProbab :151-194)
//   HitList::CalculateHHblitsEvalues   src/hhhitlist.cpp:465-494
//   HitList::SortList + Hit::operator< src/hhhit.h:116-126 (score_aass ascending, then file name)
// These are O(N) double-precision formulas over exp/log of the C library.  They stay on the host ON PURPOSE: the
// reference calls glibc's exp/expf/log, and only the same library reproduces its last bits (a device exp would not);
// at 40 ns per hit they cost 4 ms per 100k hits on one core, and the multi-GPU merge only needs them for the K
// records it exchanges.  Operation types follow the reference expression by expression (which operand is float,
// which is double, where a float overload of exp is selected) and are pinned by tests/test_hitlist_cpu.py against
// the compiled reference and a committed golden file.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/hhg.h"

namespace hhg_hitlist {

// calc_hidden_output, src/hhhitlist-inl.h:14-22: float accumulation; exp(float) is the float overload (<math.h> of
// libstdc++ injects std::exp(float) into the global namespace), the logistic itself is evaluated in double
inline float hidden(const float* w, const float* bias, float Lq, float Lt, float Nq, float Nt) {
  float res = Lq * w[0] + Lt * w[1] + Nq * w[2] + Nt * w[3] + *bias;
  res = (float)(1.0 / (1.0 + (double)expf(-(res))));
  return res;
}

inline float lamda_NN(float Lq, float Lt, float Nq, float Nt) {     // :27-42
  static const float biases[] = {-0.73195f, -1.43792f, -1.18839f, -3.01141f};
  static const float weights[] = {-0.52356f, -3.37650f, 1.12984f, -0.46796f, -4.71361f, 0.14166f, 1.66807f, 0.16383f,
                                  -0.94895f, -1.24358f, -1.20293f, 0.95434f, -0.00318f, 0.53022f, -0.04914f, -0.77046f,
                                  2.45630f, 3.02905f, 2.53803f, 2.64379f};
  float lamda = 0.0f;
  for (int h = 0; h < 4; h++) lamda += hidden(weights + 4 * h, biases + h, Lq, Lt, Nq, Nt) * weights[16 + h];
  return lamda;
}

inline float mu_NN(float Lq, float Lt, float Nq, float Nt) {        // :47-69
  static const float biases[] = {-4.25264f, -3.63484f, -5.86653f, -4.78472f, -2.76356f, -2.21580f};
  static const float weights[] = {1.96172f, 1.07181f, -7.41256f, 0.26471f, 0.84643f, 1.46777f, -1.04800f, -0.51425f,
                                  1.42697f, 1.99927f, 0.64647f, 0.27834f, 1.34216f, 1.64064f, 0.35538f, -8.08311f,
                                  2.30046f, 1.31700f, -0.46435f, -0.46803f, 0.90090f, -3.53067f, 0.59212f, 1.47503f,
                                  -1.26036f, 1.52812f, 1.58413f, -1.90409f, 0.92803f, -0.66871f};
  float mu = 0.0f;
  for (int h = 0; h < 6; h++) mu += hidden(weights + 4 * h, biases + h, Lq, Lt, Nq, Nt) * weights[24 + h];
  return (float)(20.0 * mu);
}

inline double Pvalue(float x, float lamda, float mu) {               // src/hhhit-inl.h:44-47
  double h = lamda * (x - mu);
  return (h > 10) ? exp(-h) : (double(1.0) - exp(-exp(-h)));
}
inline double logPvalue(float x, float lamda, float mu) {            // :49-53
  double h = lamda * (x - mu);
  return (h > 10) ? -h : (h < -2.5) ? -exp(-exp(-h)) : log((double(1.0) - exp(-exp(-h))));
}

// Hit::CalcProbab, src/hhhit.h:151-194
inline double CalcProbab(float score_aass, int loc, int ssm, int hit_has_ss, float ssw) {
  double s = -score_aass;
  double t = 0;
  if (s > 200) return 100.0;
  if (loc) {
    if (ssm && hit_has_ss && ssw > 0) { const double a = sqrt(6000.0), b = 2.0 * 2.5, c = sqrt(0.12), d = 2.0 * 32.0; t = a * exp(-s / b) + c * exp(-s / d); }
    else { const double a = sqrt(4000.0), b = 2.0 * 2.5, c = sqrt(0.15), d = 2.0 * 34.0; t = a * exp(-s / b) + c * exp(-s / d); }
  } else {
    if (ssm > 0 && ssw > 0) { const double a = sqrt(4000.0), b = 2.0 * 3.0, c = sqrt(0.13), d = 2.0 * 34.0; t = a * exp(-s / b) + c * exp(-s / d); }
    else { const double a = sqrt(6000.0), b = 2.0 * 2.5, c = sqrt(0.10), d = 2.0 * 37.0; t = a * exp(-s / b) + c * exp(-s / d); }
  }
  return 100.0 / (1.0 + t * t);
}

}  // namespace hhg_hitlist

extern "C" {

int hhg_hitlist_pvalues(int n, const float* score, const float* score_ss, const int32_t* Lt, const float* t_neff,
                        const int32_t* hit_has_ss, int Lq, float q_neff, int N_searched, int loc, int ssm, float ssw,
                        hhg_hit_stats* out) {
  using namespace hhg_hitlist;
  if (n < 0 || (n > 0 && (!score || !score_ss || !Lt || !t_neff || !out)) || Lq < 1) return HHG_EINVAL;
  float lamda = 0.42f /* LAMDA_GLOB, src/hhdecl.h:43 */, mu = 3.0f;   // global search: fixed
  const float log1000 = (float)log(1000.0);
  if (N_searched == 0) N_searched = 1;
  for (int k = 0; k < n; ++k) {
    if (loc) {
      lamda = lamda_NN((float)(log(Lq) / log1000), (float)(log(Lt[k]) / log1000), (float)(q_neff / 10.0), (float)(t_neff[k] / 10.0));
      mu = mu_NN((float)(log(Lq) / log1000), (float)(log(Lt[k]) / log1000), (float)(q_neff / 10.0), (float)(t_neff[k] / 10.0));
    }
    hhg_hit_stats& o = out[k];
    o.logPval = logPvalue(score[k], lamda, mu);
    o.Pval = Pvalue(score[k], lamda, mu);
    // Hit::CalcEvalScoreProbab, src/hhhit.h:134-141
    o.Eval = exp(o.logPval + log(N_searched));
    o.logEval = o.logPval + log(N_searched);
    o.score_aass = (float)((o.logPval < -10.0 ? o.logPval : log(-log(1 - o.Pval))) / 0.45 -
                           fmin(lamda * score_ss[k], fmax(0.0, 0.2 * (score[k] - 8.0))) / 0.45 - 3.0);
    o.Probab = (float)CalcProbab(o.score_aass, loc, ssm, hit_has_ss ? hit_has_ss[k] : 0, ssw);
    o.lamda = lamda;
    o.mu = mu;
  }
  return HHG_OK;
}

int hhg_hitlist_hhblits_evalues(int n, hhg_hit_stats* stats, const float* t_neff, float q_neff, int dbsize, float alphaa,
                                float alphab, float alphac, double prefilter_evalue_thresh) {
  if (n < 0 || (n > 0 && (!stats || !t_neff)) || dbsize < 1) return HHG_EINVAL;
  const double log_Pcut = log(prefilter_evalue_thresh / dbsize);
  const double log_dbsize = log((double)dbsize);
  for (int k = 0; k < n; ++k) {
    double alpha = alphaa + alphab * (t_neff[k] - 1) * (1 - alphac * (q_neff - 1));   // float expression, then double
    stats[k].Eval = exp(stats[k].logPval + log_dbsize + (alpha * log_Pcut));
    stats[k].logEval = stats[k].logPval + log_dbsize + (alpha * log_Pcut);
  }
  return HHG_OK;
}

// ViterbiRunner::calculateEarlyStop, src/hhviterbirunner.cpp:213-247: the sum over one chunk's hits of 1/(1+Eval) that
// hhblits compares with chunk_size * filter_thresh (:178-188).  The reference evaluates this one in FLOAT where
// CalculateHHblitsEvalues uses double, and feeds the already normalised Neff/10 into alpha; both reproduced.
float hhg_early_stop_sum(int n, const float* score, const int32_t* Lt, const float* t_neff, int Lq, float q_neff_hmm,
                         int prefilter, int dbsize, float alphaa, float alphab, float alphac,
                         double prefilter_evalue_thresh) {
  using namespace hhg_hitlist;
  const float LOG1000 = (float)log(1000.0);                        // src/hhdecl.h:48
  float early_stop_result = 0.0f;
  for (int k = 0; k < n; ++k) {
    float q_len = (float)(log(Lq) / LOG1000);
    float hit_len = (float)(log(Lt[k]) / LOG1000);
    float q_neff = (float)(q_neff_hmm / 10.0);
    float hit_neff = (float)(t_neff[k] / 10.0);
    float lamda = lamda_NN(q_len, hit_len, q_neff, hit_neff);
    float mu = mu_NN(q_len, hit_len, q_neff, hit_neff);
    const double logPval = logPvalue(score[k], lamda, mu);
    float alpha = 0;
    float log_Pcut = (float)log(prefilter_evalue_thresh / dbsize);
    float log_dbsize = (float)log(dbsize);
    if (prefilter) alpha = alphaa + alphab * (hit_neff - 1) * (1 - alphac * (q_neff - 1));
    const double Eval = exp(logPval + log_dbsize + (alpha * log_Pcut));
    float eval_normalized = (float)(1.0 / (1.0 + Eval));
    early_stop_result += eval_normalized;
  }
  return early_stop_result;
}

int hhg_hitlist_order(int n, const hhg_hit_stats* stats, const char* const* file, int32_t* order) {
  if (n < 0 || (n > 0 && (!stats || !order))) return HHG_EINVAL;
  std::iota(order, order + n, 0);
  std::stable_sort(order, order + n, [&](int32_t a, int32_t b) {
    if (stats[a].score_aass < stats[b].score_aass) return true;
    if (stats[b].score_aass < stats[a].score_aass) return false;
    if (file) { const int c = strcmp(file[a], file[b]); if (c) return c < 0; }
    return false;
  });
  return HHG_OK;
}

}  // extern "C"
