/* oracle/hh_oracle.c -- TEST INFRASTRUCTURE (the parity oracle), NOT product code.
 *
 * A scalar, per-target CPU restatement of the reference's hot path, written from the behaviour
 * documented in SURVEY.md App. A and checked bit-for-bit against the compiled reference
 * (oracle/_ref/libhhref_shim.so, AVX2 / no-FMA build = the official release flags) by
 * tests/test_oracle_vs_ref.py and against the committed fixtures in tests/golden/.
 * PARITY PINNED: yes (reference run in this container; goldens committed with their generator
 * tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product (hh-suite_b200/csrc) never links or calls it.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared  (no -mfma: every mul/add is a
 * separately rounded IEEE fp32 operation, as in the reference's -mavx2 build).
 *
 * Data layout (shared with include/hhg.h):
 *   p  : float[(L+2)*20]   p[i*20+a], columns 0..L+1 (reference HMM::p, src/hhhmm.h:153)
 *   tr : float[(L+1)*7]    tr[i*7+k], k in the reference enum order M2M,M2I,M2D,I2M,I2I,D2M,D2D
 *                          (src/hhdecl.h:68), log2 transition probabilities
 *   ss : u8[L+2]           ss[i] = ss_pred[i]*MAXCF + ss_conf[i]  (src/hhhmmsimd.cpp:133), MAXCF=11
 *   S33: float[44*44]      S33[q_ss*44 + t_ss]   (src/hhdecl.h:53-55: NSSPRED=4, MAXCF=11)
 *   bt : u8[(Lq+1)*(Lt+1)] bt[i*(Lt+1)+j], bit layout of ViterbiMatrix (src/hhviterbimatrix-inl.h)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { M2M = 0, M2I = 1, M2D = 2, I2M = 3, I2I = 4, D2M = 5, D2D = 6 };
enum { ST_STOP = 0, ST_MM = 2, ST_GD = 3, ST_IM = 4, ST_DG = 5, ST_MI = 6 };

static inline uint32_t f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }

/* Follows Viterbi::ScalarProd20Vec, src/hhviterbi.h:126-190: four partial sums over a = k mod 4,
 * accumulated in increasing a, combined as (r0+r1)+(r2+r3). */
static inline float dot20(const float* q, const float* t) {
  float r0 = t[0] * q[0], r1 = t[1] * q[1], r2 = t[2] * q[2], r3 = t[3] * q[3];
  for (int m = 1; m < 5; ++m) {
    r0 = t[4 * m + 0] * q[4 * m + 0] + r0;
    r1 = t[4 * m + 1] * q[4 * m + 1] + r1;
    r2 = t[4 * m + 2] * q[4 * m + 2] + r2;
    r3 = t[4 * m + 3] * q[4 * m + 3] + r3;
  }
  r0 = r0 + r1;
  r2 = r2 + r3;
  return r0 + r2;
}

/* Follows log2f4, src/hhutil-inl.h:509-541 (LOG_POLY_DEGREE 4 => POLY3). */
static inline float log2f4_scalar(float x) {
  uint32_t i = f2u(x);
  float e = (float)((int32_t)((i & 0x7F800000u) >> 23) - 127);
  float m = u2f((i & 0x007FFFFFu) | 0x3F800000u);
  float p = -0.107254423828329604454f;
  p = p * m + 0.688243882994381274313f;
  p = p * m + -1.75647175389045657003f;
  p = p * m + 2.61761038894603480148f;
  p = p * (m - 1.0f);
  return p + e;
}

/* The forward pass. Follows Viterbi::AlignWith[Out]CellOff[AndSS],
 * src/hhviterbialgorithm.cpp:77-497 (one lane of the SIMD batch, no padding columns).
 *   q_ss/t_ss/S33 : all non-NULL selects the *AndSS variants (PRED_PRED mode, :199-211,279)
 *   celloff       : non-NULL selects the *CellOff variants (:373-392); bytes != 0 are off
 * Returns 0. Outputs: *score,*i2,*j2 (ViterbiResult, src/hhviterbi.h:21), bt bytes (bits 0-6). */
int hho_viterbi_align(int Lq, const float* q_p, const float* q_tr, const uint8_t* q_ss,
                      int Lt, const float* t_p, const float* t_tr, const uint8_t* t_ss,
                      const float* S33, float ssw, const uint8_t* celloff, int local, float egq,
                      float egt, float shift, float* score, int* i2, int* j2, uint8_t* bt) {
  const float smin = local ? 0.0f : -FLT_MAX;
  const int use_ss = (q_ss && t_ss && S33);
  float* sMM = (float*)malloc(sizeof(float) * 5 * (size_t)(Lt + 1));
  float *sDG = sMM + (Lt + 1), *sMI = sDG + (Lt + 1), *sGD = sMI + (Lt + 1), *sIM = sGD + (Lt + 1);
  float best = -FLT_MAX;
  int bi = 0, bj = 0;
  for (int j = 0; j <= Lt; ++j) {           /* :145-153 */
    sMM[j] = (float)(-j) * egt;
    sDG[j] = sMI[j] = sGD[j] = sIM[j] = -FLT_MAX;
  }
  for (int i = 1; i <= Lq; ++i) {
    float dMM = (float)(-(i - 1)) * egq;      /* :161-165 */
    float dIM = -FLT_MAX, dMI = -FLT_MAX, dDG = -FLT_MAX, dGD = -FLT_MAX;
    sMM[0] = (float)(-i) * egq;               /* :169-173 */
    sDG[0] = sMI[0] = sGD[0] = sIM[0] = -FLT_MAX;
    const float* qt1 = q_tr + (size_t)(i - 1) * 7;
    const float* qt0 = q_tr + (size_t)i * 7;
    const float q_m2m = qt1[M2M], q_m2d = qt1[M2D], q_d2m = qt1[D2M], q_d2d = qt1[D2D],
                q_i2m = qt1[I2M], q_i2i = qt0[I2I], q_m2i = qt0[M2I];
    const int find_max = (local || i == Lq);
    const float* qp = q_p + (size_t)i * 20;
    const float* ssrow = use_ss ? S33 + (size_t)q_ss[i] * 44 : NULL;
    float mm = 0.0f;
    for (int j = 1; j <= Lt; ++j) {
      const float* tt1 = t_tr + (size_t)(j - 1) * 7;
      const float* tt0 = t_tr + (size_t)j * 7;
      const float t_m2m = tt1[M2M], t_m2d = tt1[M2D], t_d2m = tt1[D2M], t_d2d = tt1[D2D],
                  t_i2m = tt1[I2M], t_i2i = tt0[I2I], t_m2i = tt0[M2I];
      unsigned b = 0;
      float c;
      /* 5-way max with strict '>' and first-wins ties, :241-273 */
      c = (dMM + q_m2m) + t_m2m;
      if (c > smin) b = ST_MM;
      mm = (smin > c) ? smin : c;           /* _mm256_max_ps(smin, c) */
      c = (dGD + q_m2m) + t_d2m;
      if (c > mm) b = ST_GD;
      mm = (mm > c) ? mm : c;
      c = (dIM + q_i2m) + t_m2m;
      if (c > mm) b = ST_IM;
      mm = (mm > c) ? mm : c;
      c = (dDG + q_d2m) + t_m2m;
      if (c > mm) b = ST_DG;
      mm = (mm > c) ? mm : c;
      c = (dMI + q_m2m) + t_i2m;
      if (c > mm) b = ST_MI;
      mm = (mm > c) ? mm : c;
      float Si = log2f4_scalar(dot20(qp, t_p + (size_t)j * 20));   /* :277 */
      if (use_ss) Si = (ssw * ssrow[t_ss[j]]) + Si;                /* :210,279 */
      Si = Si + shift;                                             /* :281 */
      mm = mm + Si;
      const float lMM = sMM[j - 1], lGD = sGD[j - 1], lIM = sIM[j - 1];
      const float uMM = sMM[j], uDG = sDG[j], uMI = sMI[j];
      dMM = sMM[j]; dDG = sDG[j]; dMI = sMI[j]; dGD = sGD[j]; dIM = sIM[j];   /* :294-298 */
      float a1, a2, gd, im, dg, mi;
      a1 = lMM + t_m2d; a2 = lGD + t_d2d;                          /* :307-315 */
      if (a1 > a2) b |= 8;
      gd = (a1 > a2) ? a1 : a2;
      a1 = (lMM + q_m2i) + t_m2m; a2 = (lIM + q_i2i) + t_m2m;      /* :324-332 */
      if (a1 > a2) b |= 16;
      im = (a1 > a2) ? a1 : a2;
      a1 = uMM + q_m2d; a2 = uDG + q_d2d;                          /* :340-348 */
      if (a1 > a2) b |= 32;
      dg = (a1 > a2) ? a1 : a2;
      a1 = (uMM + q_m2m) + t_m2i; a2 = (uMI + q_m2m) + t_i2i;      /* :358-366 */
      if (a1 > a2) b |= 64;
      mi = (a1 > a2) ? a1 : a2;
      if (celloff && celloff[(size_t)i * (Lt + 1) + j]) {          /* :373-392 */
        mm = mm + -FLT_MAX; gd = gd + -FLT_MAX; im = im + -FLT_MAX;
        dg = dg + -FLT_MAX; mi = mi + -FLT_MAX;
      }
      sMM[j] = mm; sDG[j] = dg; sMI[j] = mi; sGD[j] = gd; sIM[j] = im;        /* :396-400 */
      if (bt) bt[(size_t)i * (Lt + 1) + j] = (uint8_t)b;                      /* :403-417 */
      if (find_max && mm > best) { best = mm; bi = i; bj = j; }               /* :423-455 */
    }
    if (!local && mm > best) { best = mm; bi = i; bj = Lt; }                  /* :462-486 */
  }
  free(sMM);
  *score = best; *i2 = bi; *j2 = bj;
  return 0;
}

/* Follows Viterbi::Backtrace, src/hhviterbi.cpp:83-160. Arrays need i2+j2+2 entries; entry 0 unused.
 * Returns nsteps; step 1 = (i2,j2). */
int hho_backtrace(int Lt, const uint8_t* bt, int i2, int j2, int* i_steps, int* j_steps,
                  uint8_t* states, int* matched_cols) {
  int step = 0, i = i2, j = j2, mc = 0;
  int state = ST_MM;
  while (state != ST_STOP) {
    ++step;
    states[step] = (uint8_t)state; i_steps[step] = i; j_steps[step] = j;
    const uint8_t c = bt[(size_t)i * (Lt + 1) + j];
    switch (state) {
      case ST_MM: ++mc; state = (i <= 1 || j <= 1) ? ST_STOP : (c & 7); --i; --j; break;
      case ST_GD: if (j <= 1) state = ST_STOP; else { if (c & 8) state = ST_MM; --j; } break;
      case ST_IM: if (j <= 1) state = ST_STOP; else { if (c & 16) state = ST_MM; --j; } break;
      case ST_DG: if (i <= 1) state = ST_STOP; else { if (c & 32) state = ST_MM; --i; } break;
      case ST_MI: if (i <= 1) state = ST_STOP; else { if (c & 64) state = ST_MM; --i; } break;
      default: state = ST_STOP; break;
    }
  }
  states[step] = ST_MM;
  *matched_cols = mc;
  return step;
}

/* Follows Viterbi::ExcludeAlignment, src/hhviterbi.cpp:61-77 (cross of half-width 40 around each
 * path step 1 <= step < nsteps). celloff: u8[(Lq+1)*(Lt+1)], set to 1. */
void hho_exclude_alignment(int Lq, int Lt, const int* i_steps, const int* j_steps, int nsteps,
                           uint8_t* celloff) {
  const int W = 40; /* VITERBI_PATH_WIDTH, src/hhdecl.h:50 */
  for (int s = 1; s < nsteps; ++s) {
    const int i = i_steps[s], j = j_steps[s];
    for (int ii = (i - W > 1 ? i - W : 1); ii <= (i + W < Lq ? i + W : Lq); ++ii)
      celloff[(size_t)ii * (Lt + 1) + j] = 1;
    for (int jj = (j - W > 1 ? j - W : 1); jj <= (j + W < Lt ? j + W : Lt); ++jj)
      celloff[(size_t)i * (Lt + 1) + jj] = 1;
  }
}

/* ------------------------------------------------------------------ prefilter */

/* Follows flog2, src/util-inl.h:83-93 (note the double-precision polynomial constants). */
static inline float flog2_scalar(float x) {
  if (x <= 0) return -128;
  uint32_t u = f2u(x);
  float e = (float)((int)((u & 0x7F800000u) >> 23) - 0x7f);
  x = u2f((u & 0x007FFFFFu) | 0x3f800000u);
  x -= 1.0;
  x *= (1.441740 + x * (-0.7077702 + x * (0.4123442 + x * (-0.1903190 + x * 0.0440047))));
  return x + e;
}
float hho_flog2(float x) { return flog2_scalar(x); }

/* Linear (un-striped) query profile. Follows Prefilter::stripe_query_profile,
 * src/hhprefilter.cpp:356-424: prof[k*Lq + pos] for k<219 from q.p[pos] (0-based pos indexing the
 * 1-based p array: SURVEY App. D-4), row 219 = ANY state = offset-1.
 *   q_p: (Lq+2)*20, q_pav: 20, lib: 219*20 linear column-state probabilities. */
void hho_prefilter_query_profile(int Lq, const float* q_p, const float* q_pav, const float* lib,
                                 int score_offset, int bit_factor, uint8_t* prof) {
  for (int k = 0; k < 219; ++k)
    for (int pos = 0; pos < Lq; ++pos) {
      float sum = 0;
      for (int a = 0; a < 20; ++a) sum += ((q_p[(size_t)pos * 20 + a] * lib[k * 20 + a]) / q_pav[a]);
      float dummy = flog2_scalar(sum) * bit_factor + score_offset + 0.5;
      uint8_t v;
      if (dummy > 255.0) v = 255; else if (dummy < 0) v = 0; else v = (uint8_t)dummy;
      prof[(size_t)k * Lq + pos] = v;
    }
  for (int pos = 0; pos < Lq; ++pos) prof[(size_t)219 * Lq + pos] = (uint8_t)(score_offset - 1);
}

/* Follows Prefilter::ungapped_sse_score, src/hhprefilter.cpp:214-275, un-striped:
 * S(i,j) = max(0, min(255, S(i-1,j-1) + prof[x_j][i]) - offset); result = max over all cells.
 * Striped padding positions (i >= Lq) hold `offset` => contribute max(0, S) which never exceeds
 * the running maximum of real cells, so they are omitted. */
int hho_ungapped_score(int Lq, const uint8_t* prof, const uint8_t* dbseq, int Lt, int offset) {
  uint8_t* prev = (uint8_t*)calloc((size_t)Lq + 1, 1);
  uint8_t* cur = (uint8_t*)calloc((size_t)Lq + 1, 1);
  int smax = 0;
  for (int j = 0; j < Lt; ++j) {
    const uint8_t* row = prof + (size_t)dbseq[j] * Lq;
    for (int i = 0; i < Lq; ++i) {
      int s = (i > 0 ? prev[i - 1] : 0) + row[i];
      if (s > 255) s = 255;
      s -= offset;
      if (s < 0) s = 0;
      cur[i] = (uint8_t)s;
      if (s > smax) smax = s;
    }
    uint8_t* t = prev; prev = cur; cur = t;
  }
  free(prev); free(cur);
  return smax;
}

/* Stage-1 length correction, src/hhprefilter.cpp:477. */
int hho_ungapped_corrected(int raw, int Lq, int Lt, int bit_factor) {
  return raw - (int)(bit_factor * (flog2_scalar((float)Lq) + flog2_scalar((float)Lt)));
}

/* Follows Prefilter::swStripedByte, src/hhprefilter.cpp:70-212, for the AVX2 build (32 byte lanes,
 * segLen = ceil(Lq/32), full-width byte shift _mm256_shift_left<1>, lib/simd/simd.h:184-187): gapped local
 * SW on unsigned saturating bytes, Farrar striping with the SWPS3-style lazy-F loop that does NOT update E
 * (so the result can depend on the striping, SURVEY App. D-5 -- hence this lane-exact emulation rather than
 * a textbook SW).  prof: LINEAR profile [220][Lq]; padding positions behave as `bias` (:392-393).
 * gap_open here is the reference's gapOpen argument (= prefilter_gap_open + prefilter_gap_extend, :456). */
int hho_sw_striped_byte(int Lq, const uint8_t* prof, const uint8_t* dbseq, int Lt, int gap_open,
                        int gap_extend, int bias) {
  enum { V = 32 };
  const int W = (Lq + V - 1) / V;
  uint8_t* buf = (uint8_t*)calloc((size_t)3 * W * V, 1);
  uint8_t *Hst = buf, *Hld = buf + (size_t)W * V, *E = buf + (size_t)2 * W * V;
  int vmax[V];
  for (int k = 0; k < V; ++k) vmax[k] = 0;
#define SUBS(a, b) ((a) > (b) ? (a) - (b) : 0)
  for (int i = 0; i < Lt; ++i) {
    int vF[V], vH[V], vMaxCol[V];
    const uint8_t* row = prof + (size_t)dbseq[i] * Lq;
    for (int k = 0; k < V; ++k) { vF[k] = 0; vMaxCol[k] = 0; }
    for (int k = V - 1; k > 0; --k) vH[k] = Hst[(size_t)(W - 1) * V + k - 1];   /* shiftl(pvHStore[segLen-1]) */
    vH[0] = 0;
    { uint8_t* t = Hld; Hld = Hst; Hst = t; }
    for (int j = 0; j < W; ++j) {
      for (int k = 0; k < V; ++k) {
        const int pos = k * W + j;
        const int p = pos < Lq ? row[pos] : bias;
        int h = vH[k] + p; if (h > 255) h = 255;
        h = SUBS(h, bias);
        int e = E[(size_t)j * V + k];
        if (e > h) h = e;
        if (vF[k] > h) h = vF[k];
        if (h > vMaxCol[k]) vMaxCol[k] = h;
        Hst[(size_t)j * V + k] = (uint8_t)h;
        h = SUBS(h, gap_open);
        e = SUBS(e, gap_extend);
        if (h > e) e = h;
        E[(size_t)j * V + k] = (uint8_t)e;
        int f = SUBS(vF[k], gap_extend);
        if (h > f) f = h;
        vF[k] = f;
        vH[k] = Hld[(size_t)j * V + k];
      }
    }
    /* lazy F */
    int j = 0;
    for (int k = V - 1; k > 0; --k) vF[k] = vF[k - 1];
    vF[0] = 0;
    for (;;) {
      int all = 1;
      for (int k = 0; k < V; ++k) {
        const int h = Hst[(size_t)j * V + k];
        if (SUBS(vF[k], SUBS(h, gap_open)) != 0) all = 0;
      }
      if (all) break;
      for (int k = 0; k < V; ++k) {
        int h = Hst[(size_t)j * V + k];
        if (vF[k] > h) h = vF[k];
        if (h > vMaxCol[k]) vMaxCol[k] = h;
        Hst[(size_t)j * V + k] = (uint8_t)h;
        vF[k] = SUBS(vF[k], gap_extend);
      }
      if (++j >= W) {
        j = 0;
        for (int k = V - 1; k > 0; --k) vF[k] = vF[k - 1];
        vF[0] = 0;
      }
    }
    for (int k = 0; k < V; ++k) if (vMaxCol[k] > vmax[k]) vmax[k] = vMaxCol[k];
  }
#undef SUBS
  int score = 0;
  for (int k = 0; k < V; ++k) if (vmax[k] > score) score = vmax[k];
  free(buf);
  return score;
}

/* Follows fpow2, src/util-inl.h:190-214. */
float hho_fpow2(float x) {
  if (x >= FLT_MAX_EXP) return FLT_MAX;
  if (x <= FLT_MIN_EXP) return 0.0f;
  float tx = (x - 0.5f) + (3 << 22);
  int lx = (int)(f2u(tx) - 0x4b400000u);
  float dx = x - (float)lx;
  x = 1.0f + dx * (0.693019f + dx * (0.241404f + dx * (0.0520749f + dx * 0.0134929f)));
  return u2f(f2u(x) + ((uint32_t)lx << 23));
}
