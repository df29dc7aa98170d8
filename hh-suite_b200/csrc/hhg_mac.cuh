// hhg_mac.cuh -- MAC realignment of the reported hits, the step right after Viterbi:
//   PosteriorDecoder::realign            src/hhposteriordecoder.cpp:85-118
//   maskViterbiAlignment / excludeMAC    src/hhposteriordecoder.cpp:207-258   (cell-off band, FWD_BKW_PATHWITDH)
//   forwardAlgorithm                     src/hhforwardalgorithm.cpp:10-220    (double, row scaling)
//   backwardAlgorithm                    src/hhbackwardalgorithm.cpp:10-140   (posterior = F*B/Pforward, float store)
//   macAlgorithm                         src/hhmacalgorithm.cpp:17-160        (float DP over posteriors - mact)
//   backtraceMAC                         src/hhbacktracemac.cpp:112-210
//
// One warp per hit.  Every value is produced by the same sequence of IEEE operations as in the reference (same
// types, same association, no FMA contraction), so the posterior matrix, Pforward and the MAC path are
// bit-identical -- the parallelism comes only from what is independent in the recurrences:
//   * within a row, MM / DG / MI depend on the previous row only  -> lanes = columns
//   * GD and IM are first-order recurrences along the row          -> two lanes scan the row sequentially
//   * the row maximum (scale factors) is order-independent         -> warp reduction
//   * hits are independent                                         -> one warp each, all hits of a query concurrently
// No secondary-structure term (hit.ssm2 == 0) and no self-alignment mode.
#pragma once
#include <cfloat>
#include <cstdint>

namespace hhg {

struct MacHitOut {          // mirrors hhg_mac_hit (include/hhg.h)
  int32_t i1, i2, j1, j2, nsteps, matched_cols;
  float sum_of_probs;
  int32_t flags;
  double pforward;
  long long path_off;
};

struct MacArgs {
  int n, Lq, local;
  float mact;
  double Cshift;                       // pow(2.0, shift), host-computed (libm)
  const float* q_p;                    // [(Lq+2)*20]
  const float* q_tr;                   // [(Lq+1)*7] linear, boundary rows already reset
  const ColRec* cols;                  // prepared shard records
  const long long* rec0;               // [n] first record of the request's target
  const int* Lt;                       // [n]
  const float* t_tr;                   // linear template transitions, request r at tr_off[r], (Lt+1)*7 floats
  const long long* tr_off;
  const int* vit;                      // [n*5] i1,i2,j1,j2,nsteps
  const long long* vit_off;            // [n+1] into vit_i/vit_j (entries 0..nsteps-1 = steps 1..nsteps)
  const int* vit_i; const int* vit_j;
  const long long* excl_off;           // [n+1] or nullptr
  const int* excl_i; const int* excl_j;
  // scratch / outputs
  const long long* cell_off;           // [n] offset of the request's (Lq+1)*(Lt+1) cell block
  float* post; uint8_t* off; uint8_t* bt;
  const long long* row_off;            // [n] offset (in doubles) of 10*(Lt+3) row buffers
  double* rows;
  double* scale;                       // [n*(Lq+3)]
  MacHitOut* out;
  long long* path_off;                 // [n] into out_i/out_j/out_states/out_post
  int* out_i; int* out_j; uint8_t* out_states; float* out_post;
};

__device__ __forceinline__ float mac_dot20(const float* __restrict__ qi, const float* __restrict__ tj) {
  float s = __fmul_rn(tj[0], qi[0]);                 // ScalarProd20, src/hhhit-inl.h:117-122: left to right
#pragma unroll
  for (int a = 1; a < 20; ++a) s = __fadd_rn(s, __fmul_rn(tj[a], qi[a]));
  return s;
}

// Gather the log2 transition rows of the requested templates from the shard's column records into (Lt+1) x 7
// arrays (enum order M2M,M2I,M2D,I2M,I2I,D2M,D2D).  The host then applies HMM::Log2LinTransitionProbs
// (src/hhhmm.cpp:2305-2313: pow(2.0f, tr) = the C library's powf, whose rounding a device function cannot reproduce)
// and the boundary rows of initializeForAlignment (src/hhposteriordecoder.cpp:158-167).
__global__ void k_mac_gather_tr(int n, const ColRec* __restrict__ cols, const long long* __restrict__ rec0,
                                const int* __restrict__ Lt, const long long* __restrict__ tr_off,
                                float* __restrict__ t_tr) {
  const int r = blockIdx.y;
  if (r >= n) return;
  const int L = Lt[r];
  float* tr = t_tr + tr_off[r];
  const ColRec* c = cols + rec0[r];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= L; i += gridDim.x * blockDim.x) {
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (i >= 1 && i < L) {
      // row i: M2M,M2D,D2M,D2D,I2M sit in the record of column i+1 (index i), I2I,M2I in that of column i (index i-1)
      v[0] = c[i].m2m; v[1] = c[i - 1].m2i; v[2] = c[i].m2d; v[3] = c[i].i2m; v[4] = c[i - 1].i2i; v[5] = c[i].d2m;
      v[6] = c[i].d2d;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) tr[(size_t)i * 7 + k] = v[k];
  }
}

// Cell-off band: block per request.
__global__ void __launch_bounds__(256) k_mac_band(const MacArgs A) {
  const int r = blockIdx.x;
  const int Lq = A.Lq, Lt = A.Lt[r], W = Lt + 1;
  uint8_t* off = A.off + A.cell_off[r];
  const int i1 = A.vit[r * 5], i2 = A.vit[r * 5 + 1], j1 = A.vit[r * 5 + 2], j2 = A.vit[r * 5 + 3], ns = A.vit[r * 5 + 4];
  const long long total = (long long)(Lq + 1) * W;
  for (long long c = threadIdx.x; c < total; c += blockDim.x) {
    const int i = (int)(c / W), j = (int)(c - (long long)i * W);
    off[c] = (i >= 1 && j >= 1) ? (uint8_t)!((i < i1 && j < j1) || (i > i2 && j > j2)) : (uint8_t)0;
  }
  __syncthreads();
  const int* vi = A.vit_i + A.vit_off[r];
  const int* vj = A.vit_j + A.vit_off[r];
  // switch on the +-40 cross around every step of the Viterbi path (idempotent stores of 0: order-free)
  for (int t = threadIdx.x; t < ns * 81; t += blockDim.x) {
    const int s = t / 81, d = t - s * 81 - 40;
    const int i = vi[s], j = vj[s];
    if (i + d >= 1 && i + d <= Lq) off[(size_t)(i + d) * W + j] = 0;
    if (j + d >= 1 && j + d <= Lt) off[(size_t)i * W + (j + d)] = 0;
  }
  __syncthreads();
  if (A.excl_off) {
    const int* ei = A.excl_i + A.excl_off[r];
    const int* ej = A.excl_j + A.excl_off[r];
    const int ne = (int)(A.excl_off[r + 1] - A.excl_off[r]);
    for (int t = threadIdx.x; t < ne * 5; t += blockDim.x) {
      const int s = t / 5, d = t - s * 5 - 2;
      const int i = ei[s], j = ej[s];
      if (i + d >= 1 && i + d <= Lq) off[(size_t)(i + d) * W + j] = 1;
      if (j + d >= 1 && j + d <= Lt) off[(size_t)i * W + (j + d)] = 1;
    }
  }
}

// Forward + Pforward + Backward + MAC DP + backtrace: one warp per request.
__global__ void __launch_bounds__(32) k_mac_realign(const MacArgs A) {
  const int r = blockIdx.x, lane = threadIdx.x;
  const int Lq = A.Lq, Lt = A.Lt[r], W = Lt + 1;
  const uint8_t* off = A.off + A.cell_off[r];
  uint8_t* bt = A.bt + A.cell_off[r];
  float* post = A.post + A.cell_off[r];
  const ColRec* tcol = A.cols + A.rec0[r] - 1;            // tcol[j] = record of column j (1-based)
  const float* ttr = A.t_tr + A.tr_off[r];
  const float* qtr = A.q_tr;
  double* scale = A.scale + (size_t)r * (Lq + 3);
  const size_t RS = (size_t)Lt + 3;
  double* base = A.rows + A.row_off[r];
  double *Pm = base, *Pg = base + RS, *Pi = base + 2 * RS, *Pd = base + 3 * RS, *Px = base + 4 * RS;
  double *Cm = base + 5 * RS, *Cg = base + 6 * RS, *Ci = base + 7 * RS, *Cd = base + 8 * RS, *Cx = base + 9 * RS;
  const double Cshift = A.Cshift;
  const unsigned FULL = 0xffffffffu;
#define OFFC(i, j) off[(size_t)(i) * W + (j)]
#define QT(i, k) qtr[(size_t)(i) * 7 + (k)]
#define TT(j, k) ttr[(size_t)(j) * 7 + (k)]
#define SWAP_ROWS() do { double* t_; t_ = Pm; Pm = Cm; Cm = t_; t_ = Pg; Pg = Cg; Cg = t_; t_ = Pi; Pi = Ci; Ci = t_; \
                         t_ = Pd; Pd = Cd; Cd = t_; t_ = Px; Px = Cx; Cx = t_; } while (0)
  enum { M2M = 0, M2I = 1, M2D = 2, I2M = 3, I2I = 4, D2M = 5, D2D = 6 };

  // ------------------------------------------------------------------ Forward, row 1
  for (int j = lane; j <= Lt + 1; j += 32) { Cm[j] = Cg[j] = Ci[j] = Cd[j] = Cx[j] = 0.0; Pm[j] = Pg[j] = Pi[j] = Pd[j] = Px[j] = 0.0; }
  __syncwarp();
  for (int j = 1 + lane; j <= Lt; j += 32)
    if (!OFFC(1, j)) Cm[j] = (double)mac_dot20(A.q_p + 20, tcol[j].p) * Cshift;
  __syncwarp();
  if (lane < 2) {
    double v = 0.0;
    for (int j = 1; j <= Lt; ++j) {
      if (OFFC(1, j)) v = 0.0;
      else if (lane == 0) v = Cm[j - 1] * TT(j - 1, M2D) + v * TT(j - 1, D2D);
      else v = Cm[j - 1] * QT(1, M2I) * TT(j - 1, M2M) + v * QT(1, I2I) * TT(j - 1, M2M);
      if (lane == 0) Cg[j] = v; else Ci[j] = v;
    }
  }
  __syncwarp();
  for (int j = lane; j <= Lt; j += 32) post[(size_t)W + j] = (float)Cm[j];
  SWAP_ROWS();
  if (lane == 0) { scale[0] = scale[1] = scale[2] = 1.0; }
  double pmin = A.local ? 1.0 : 0.0, scale_prod = 1.0;
  __syncwarp();

  // ------------------------------------------------------------------ Forward, rows 2..Lq
  for (int i = 2; i <= Lq; ++i) {
    const double sc_i = scale[i];
    if (scale_prod < DBL_MIN * 100) scale_prod = 0.0; else scale_prod *= sc_i;
    const float* qi = A.q_p + (size_t)i * 20;
    const float q_m2m = QT(i - 1, M2M), q_i2m = QT(i - 1, I2M), q_d2m = QT(i - 1, D2M), q_m2d = QT(i - 1, M2D),
                q_d2d = QT(i - 1, D2D);
    double pmax = 0.0;
    for (int j = 1 + lane; j <= Lt; j += 32) {
      double mm = 0.0, dg = 0.0, mi = 0.0;
      if (!OFFC(i, j)) {
        const float pf = mac_dot20(qi, tcol[j].p);
        if (j == 1) {
          mm = scale_prod * 1.0f * pf * Cshift;
        } else {
          mm = pf * Cshift * 1.0f * sc_i *
               (pmin + Pm[j - 1] * q_m2m * TT(j - 1, M2M) + Pg[j - 1] * q_m2m * TT(j - 1, D2M) +
                Pi[j - 1] * q_i2m * TT(j - 1, M2M) + Pd[j - 1] * q_d2m * TT(j - 1, M2M) +
                Px[j - 1] * q_m2m * TT(j - 1, I2M));
          pmax = fmax(pmax, mm);
        }
        dg = sc_i * (Pm[j] * q_m2d + Pd[j] * q_d2d);
        mi = sc_i * (Pm[j] * q_m2m * TT(j, M2I) + Px[j] * q_m2m * TT(j, I2I));
      }
      Cm[j] = mm; Cd[j] = dg; Cx[j] = mi;
    }
    __syncwarp();
    if (lane < 2) {                                       // GD (lane 0) and IM (lane 1): sequential along the row
      const float q_m2i = QT(i, M2I), q_i2i = QT(i, I2I);
      double v = 0.0;
      if (lane == 0) Cg[1] = 0.0; else Ci[1] = 0.0;
      for (int j = 2; j <= Lt; ++j) {
        if (OFFC(i, j)) v = 0.0;
        else if (lane == 0) v = (Cm[j - 1] * TT(j - 1, M2D) + v * TT(j - 1, D2D));
        else v = (Cm[j - 1] * q_m2i * TT(j - 1, M2M) + v * q_i2i * TT(j - 1, M2M));
        if (lane == 0) Cg[j] = v; else Ci[j] = v;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) pmax = fmax(pmax, __shfl_xor_sync(FULL, pmax, o));
    __syncwarp();
    for (int j = lane; j <= Lt; j += 32) post[(size_t)i * W + j] = (float)(j ? Cm[j] : 0.0);
    SWAP_ROWS();
    pmin *= sc_i;
    if (pmin < DBL_MIN * 100) pmin = 0.0;
    if (lane == 0) scale[i + 1] = 1.0 / (pmax + 1.0);
    __syncwarp();
  }

  // ------------------------------------------------------------------ Pforward (sequential sum, row-major)
  double Pf = 0.0;
  if (lane == 0) {
    if (A.local) {
      Pf = 1.0;
      for (int i = 1; i <= Lq; ++i) {
        const float* row = post + (size_t)i * W;
        for (int j = 1; j <= Lt; ++j) Pf += row[j];
        Pf *= scale[i + 1];
      }
    } else {
      for (int i = 1; i < Lq; ++i) Pf = (Pf + post[(size_t)i * W + Lt] * scale[i + 1]);
      for (int j = 1; j <= Lt; ++j) Pf += post[(size_t)Lq * W + j];
      Pf *= scale[Lq + 1];
    }
  }
  Pf = __shfl_sync(FULL, Pf, 0);

  // ------------------------------------------------------------------ Backward
  const double sc_last = scale[Lq + 1];
  for (int j = lane; j <= Lt + 1; j += 32) { Pm[j] = Pg[j] = Pi[j] = Pd[j] = Px[j] = 0.0; Cm[j] = Cg[j] = Ci[j] = Cd[j] = Cx[j] = 0.0; }
  __syncwarp();
  for (int j = 1 + lane; j <= Lt; j += 32) {
    float* pp = post + (size_t)Lq * W + j;
    if (OFFC(Lq, j)) { *pp = 0.0f; Pm[j] = 0.0; }
    else { Pm[j] = sc_last; *pp = (float)(*pp * sc_last / Pf); }
  }
  scale_prod = sc_last;
  pmin = A.local ? sc_last : 0.0;
  __syncwarp();
  for (int i = Lq - 1; i >= 1; --i) {
    const double sc_n = scale[i + 1];
    scale_prod *= sc_n;
    if (scale_prod < DBL_MIN * 100) scale_prod = 0.0;
    pmin *= sc_n;
    if (pmin < DBL_MIN * 100) pmin = 0.0;
    const float* qn = A.q_p + (size_t)(i + 1) * 20;
    const float q_m2m = QT(i, M2M), q_m2i = QT(i, M2I), q_m2d = QT(i, M2D), q_i2m = QT(i, I2M), q_i2i = QT(i, I2I),
                q_d2m = QT(i, D2M), q_d2d = QT(i, D2D);
    // phase A: pmatch-dependent parts that need only the row below: pm (stored in Cd as scratch? no: own array)
    // Cx <- mi, Cd <- dg, and the partial sums of mm/gd/im that do not involve curr[j+1]
    for (int j = 1 + lane; j <= Lt; j += 32) {
      if (j == Lt) {
        float* pp = post + (size_t)i * W + Lt;
        if (OFFC(i, Lt)) { *pp = 0.0f; Cm[Lt] = 0.0; }
        else { Cm[Lt] = scale_prod; *pp = (float)(*pp * scale_prod / Pf); }
        Cg[Lt] = Ci[Lt] = Cd[Lt] = Cx[Lt] = 0.0;
      } else if (OFFC(i, j)) {
        Cm[j] = Cg[j] = Ci[j] = Cd[j] = Cx[j] = 0.0;
      } else {
        const double pmatch = Pm[j + 1] * mac_dot20(qn, tcol[j + 1].p) * 1.0f * Cshift * sc_n;
        Cd[j] = (+pmatch * q_d2m * TT(j, M2M) + Pd[j] * q_d2d * sc_n);
        Cx[j] = (+pmatch * q_m2m * TT(j, I2M) + Px[j] * q_m2m * TT(j, I2I) * sc_n);
        Cm[j] = pmatch;                                  // completed by the scan below
      }
    }
    __syncwarp();
    if (lane == 0) {                                      // right-to-left: GD, IM, then MM which needs both
      double g = 0.0, im = 0.0;                           // curr[Lt].gd = curr[Lt].im = 0
      for (int j = Lt - 1; j >= 1; --j) {
        if (OFFC(i, j)) { g = 0.0; im = 0.0; continue; }
        const double pmatch = Cm[j];
        const double mm = (+pmin + pmatch * q_m2m * TT(j, M2M) + g * TT(j, M2D) + im * q_m2i * TT(j, M2M) +
                           Pd[j] * q_m2d * sc_n + Px[j] * q_m2m * TT(j, M2I) * sc_n);
        const double g2 = (+pmatch * q_m2m * TT(j, D2M) + g * TT(j, D2D));
        const double i2 = (+pmatch * q_i2m * TT(j, M2M) + im * q_i2i * TT(j, M2M));
        Cm[j] = mm; Cg[j] = g2; Ci[j] = i2;
        g = g2; im = i2;
      }
    }
    __syncwarp();
    for (int j = 1 + lane; j <= Lt - 1; j += 32) post[(size_t)i * W + j] *= (float)(Cm[j] / Pf);
    SWAP_ROWS();
    __syncwarp();
  }

  // ------------------------------------------------------------------ MAC dynamic programming (float)
  float* Sp = reinterpret_cast<float*>(base);
  float* Sc = Sp + RS;
  float* T12 = Sc + RS;                                   // max(term1, term2) candidates per column
  float* T3 = T12 + RS;
  uint8_t* st12 = reinterpret_cast<uint8_t*>(T3 + RS);
  __syncwarp();
  for (int j = lane; j <= Lt; j += 32) Sp[j] = 0.0f;
  const float mact = A.mact;
  const double half_mact = 0.5 * mact;
  float score_MAC = -FLT_MAX;
  int mi2 = 0, mj2 = 0;
  if (lane == 0) bt[0] = 0;
  __syncwarp();
  for (int i = 1; i <= Lq; ++i) {
    for (int j = 1 + lane; j <= Lt; j += 32) {
      if (OFFC(i, j)) continue;
      const float p = post[(size_t)i * W + j];
      const float term1 = __fsub_rn(p, mact);
      const float term2 = __fsub_rn(__fadd_rn(Sp[j - 1], p), mact);
      if (term1 > term2) { T12[j] = term1; st12[j] = 0; } else { T12[j] = term2; st12[j] = 2; }
      T3[j] = (float)((double)Sp[j] - half_mact);
    }
    __syncwarp();
    if (lane == 0) {
      float left = 0.0f;                                  // S_curr[jmin-1] = 0
      Sc[0] = 0.0f;
      for (int j = 1; j <= Lt; ++j) {
        float mx; uint8_t st;
        if (OFFC(i, j)) { mx = -FLT_MIN; st = 0; }
        else {
          mx = T12[j]; st = st12[j];
          const float t3 = T3[j];
          if (t3 > mx) { mx = t3; st = 6; }               // MI
          const float t4 = (float)((double)left - half_mact);
          if (t4 > mx) { mx = t4; st = 4; }               // IM
          if (mx > score_MAC && (A.local || i == Lq)) { mi2 = i; mj2 = j; score_MAC = mx; }
        }
        Sc[j] = mx; bt[(size_t)i * W + j] = st;
        left = mx;
      }
      if (!A.local && Sc[Lt] > score_MAC) { mi2 = i; mj2 = Lt; score_MAC = Sc[Lt]; }
    }
    __syncwarp();
    { float* t_ = Sp; Sp = Sc; Sc = t_; }
  }

  // ------------------------------------------------------------------ MAC backtrace (lane 0)
  if (lane == 0) {
    for (int i = 0; i <= Lq; ++i) bt[(size_t)i * W + 1] = 0;
    for (int j = 1; j <= Lt; ++j) bt[(size_t)W + j] = 0;
    const long long po = A.path_off[r];
    int* oi = A.out_i + po; int* oj = A.out_j + po; uint8_t* os = A.out_states + po; float* op = A.out_post + po;
    int matched = 1, step = 0, i = mi2, j = mj2;
    uint8_t state = 2;
    oi[0] = 0; oj[0] = 0; os[0] = 0; op[0] = 0.f;
    if (mi2 < 1 || mj2 < 1 || bt[(size_t)i * W + j] != 2) {
      oi[0] = i; oj[0] = j;
    } else {
      while (state != 0) {
        ++step;
        os[step] = state = bt[(size_t)i * W + j];
        oi[step] = i; oj[step] = j;
        if (state == 2) { matched++; i--; j--; }
        else if (state == 4) j--;
        else if (state == 6) i--;
      }
    }
    MacHitOut o;
    o.i1 = oi[step]; o.i2 = mi2; o.j1 = oj[step]; o.j2 = mj2; o.nsteps = step; o.matched_cols = matched;
    if (step) os[step] = 2;
    float sum = 0.0f;
    for (int s = 1; s <= step; ++s) {
      if (os[s] == 2) { op[s] = post[(size_t)oi[s] * W + oj[s]]; sum = __fadd_rn(sum, op[s]); }
      else op[s] = 0.0f;
    }
    o.sum_of_probs = sum; o.flags = 0; o.pforward = Pf; o.path_off = po;
    A.out[r] = o;
  }
#undef OFFC
#undef QT
#undef TT
#undef SWAP_ROWS
}

}  // namespace hhg
