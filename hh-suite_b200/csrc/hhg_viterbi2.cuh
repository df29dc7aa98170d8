// hh-suite_b200/csrc/hhg_viterbi2.cuh -- forward pass with TWO target columns per query-row visit.
//
// Why: ncu + a timing experiment on k_viterbi (hhg_kernels.cuh) showed that delivering the query row
// operands to the register file is a co-bottleneck: 28 floats x 32 lanes = 3584 B per row visit = 28 cycles
// of the SM's 128 B/clk shared-memory return path, against ~105 issue slots per row on 4 schedulers
// (halving the q loads alone gave +25 %).  Here every fetched query row is used for the cells (i,j) and
// (i,j+1), so the q bytes per cell halve.  The register file cannot hold a second column's operands AND a
// one-column-ahead register prefetch of both columns next to the 80 state registers, so the next column
// pair is only prefetched into L2 (prefetch.global.L2) and loaded at the top of its iteration; with a
// column pair taking ~10k cycles per warp the exposed L2 latency is a few percent.
//
// Same arithmetic, same outputs, same work-item / hand-off-slot protocol as k_viterbi (see there).
// The plan rounds every job's Lmax up to an even number when this kernel is selected.
#pragma once
#include "hhg_kernels.cuh"

namespace hhg {

struct ColOps {                  // operands of one target column, in registers
  unsigned long long tp[10];
  float m2m, m2d, d2m, d2d, i2m, i2i, m2i;
  uint32_t ss;
};

__device__ __forceinline__ void unpack_col(ColOps& c, const float4 (&v)[7]) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    c.tp[2 * k] = pack2(v[k].x, v[k].y);
    c.tp[2 * k + 1] = pack2(v[k].z, v[k].w);
  }
  c.m2m = v[5].x; c.m2d = v[5].y; c.d2m = v[5].z; c.d2d = v[5].w;
  c.i2m = v[6].x; c.i2i = v[6].y; c.m2i = v[6].z;
  c.ss = __float_as_uint(v[6].w);
}

__device__ __forceinline__ void load_col(ColOps& c, const float4* src) {
  float4 v[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) v[k] = __ldg(src + k);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    c.tp[2 * k] = pack2(v[k].x, v[k].y);
    c.tp[2 * k + 1] = pack2(v[k].z, v[k].w);
  }
  c.m2m = v[5].x; c.m2d = v[5].y; c.d2m = v[5].z; c.d2d = v[5].w;
  c.i2m = v[6].x; c.i2i = v[6].y; c.m2i = v[6].z;
  c.ss = __float_as_uint(v[6].w);
}

struct QRow {
  float4 q[5];
  float m2m, m2d, d2m, d2d, i2m, i2i, m2i;
  uint32_t ss;
};

// One DP cell (src/hhviterbialgorithm.cpp:241-392).  d*: cell (i-1,j-1); o*: cell (i,j-1); u*: cell (i-1,j).
template <bool LOCAL, bool SS, bool CELLOFF>
__device__ __forceinline__ void dp_cell(const ColOps& t, const QRow& q, const float* s33, float ssw, float shift,
                                        unsigned long long one2, float dMM, float dGD, float dIM, float dDG,
                                        float dMI, float oMM, float oGD, float oIM, float uMM, float uDG,
                                        float uMI, bool off, float& mm, float& gd, float& im, float& dg,
                                        float& mi, uint32_t& b) {
  const float smin = LOCAL ? 0.0f : HHG_NEG;
  float c;
  c = __fadd_rn(__fadd_rn(dMM, q.m2m), t.m2m);
  b = (c > smin) ? 2u : 0u;
  mm = fmaxf(smin, c);
  c = __fadd_rn(__fadd_rn(dGD, q.m2m), t.d2m);
  b = (c > mm) ? 3u : b;
  mm = fmaxf(mm, c);
  c = __fadd_rn(__fadd_rn(dIM, q.i2m), t.m2m);
  b = (c > mm) ? 4u : b;
  mm = fmaxf(mm, c);
  c = __fadd_rn(__fadd_rn(dDG, q.d2m), t.m2m);
  b = (c > mm) ? 5u : b;
  mm = fmaxf(mm, c);
  c = __fadd_rn(__fadd_rn(dMI, q.m2m), t.i2m);
  b = (c > mm) ? 6u : b;
  mm = fmaxf(mm, c);
  float Si = log2f4_dev(dot20_dev(t.tp, q.q, one2));
  if (SS) Si = __fadd_rn(__fmul_rn(ssw, s33[q.ss * 44 + t.ss]), Si);
  Si = __fadd_rn(Si, shift);
  mm = __fadd_rn(mm, Si);
  float a1, a2;
  a1 = __fadd_rn(oMM, t.m2d); a2 = __fadd_rn(oGD, t.d2d);
  b |= (a1 > a2) ? 8u : 0u;  gd = fmaxf(a1, a2);
  a1 = __fadd_rn(__fadd_rn(oMM, q.m2i), t.m2m); a2 = __fadd_rn(__fadd_rn(oIM, q.i2i), t.m2m);
  b |= (a1 > a2) ? 16u : 0u; im = fmaxf(a1, a2);
  a1 = __fadd_rn(uMM, q.m2d); a2 = __fadd_rn(uDG, q.d2d);
  b |= (a1 > a2) ? 32u : 0u; dg = fmaxf(a1, a2);
  a1 = __fadd_rn(__fadd_rn(uMM, q.m2m), t.m2i); a2 = __fadd_rn(__fadd_rn(uMI, q.m2m), t.i2i);
  b |= (a1 > a2) ? 64u : 0u; mi = fmaxf(a1, a2);
  if (CELLOFF) {
    const float o = off ? HHG_NEG : 0.0f;
    mm = __fadd_rn(mm, o); gd = __fadd_rn(gd, o); im = __fadd_rn(im, o);
    dg = __fadd_rn(dg, o); mi = __fadd_rn(mi, o);
  }
}

struct SlotRegs {
  float mm, dg, mi, gd, im;
  uint32_t tag, p0, p1;
};
__device__ __forceinline__ void ld_slot(const BndSlot* p, SlotRegs& s) {
  ld_slot(p, s.mm, s.dg, s.mi, s.gd, s.im, s.tag, s.p0, s.p1);
}

template <int R, bool LOCAL, bool SS, bool CELLOFF>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_viterbi2(const VitParams P) {
  static_assert(R % 4 == 0, "R must be a multiple of 4");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float4* qs = reinterpret_cast<float4*>(smem_raw) + (size_t)warp * R * 7;
  constexpr size_t kBarOff = (size_t)kWarpsPerCta * R * 112;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kBarOff);
  float* s33 = reinterpret_cast<float*>(smem_raw + kBarOff + 64);
  uint64_t* bar = bars + warp;

  if (lane == 0) mbar_init(bar, 1);
  if (SS) {
    for (int k = threadIdx.x; k < 44 * 44; k += blockDim.x) s33[k] = P.S33[k];
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int total_items = P.njobs * P.nstrips;
  uint32_t parity = 0;

  for (;;) {
    int item = 0;
    if (lane == 0) item = (int)atomicAdd(P.counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total_items) break;
    const int gsz = P.group_jobs * P.nstrips;
    const int g = item / gsz;
    const int rem = item - g * gsz;
    const int gjobs = min(P.group_jobs, P.njobs - g * P.group_jobs);
    const int s = rem / gjobs;
    const int job = g * P.group_jobs + (rem - s * gjobs);
    const int i0 = s * R;

    if (lane == 0) {
      mbar_expect_tx(bar, R * 112);
      tma_bulk_g2s(qs, P.qrec + (size_t)i0 * 7, R * 112, bar);
    }

    const int t = P.job_target[job * 32 + lane];
    const int Lt = P.Lt[t];
    const int Lmax = P.job_Lmax[job];   // even
    const float4* tc = P.cols + (size_t)P.col_off[t] * 7;
    uint32_t* btj = P.bt + P.job_bt_off[job] + lane;
    const size_t bt_row_stride = (size_t)(Lmax + 1) * 32;
    BndSlot* bnd = P.bnd + P.job_bnd_off[job] + lane;
    const uint32_t tag_in = P.tag_base + (uint32_t)s;
    const uint32_t tag_out = P.tag_base + (uint32_t)s + 1u;
    const bool last_strip = (s == P.nstrips - 1);
    const uint32_t* co = nullptr;
    if (CELLOFF) co = P.celloff + P.job_co_off[job] + (size_t)s * (Lmax + 1) * 32 + lane;

    float MM[R], GD[R], IM[R], DG[R], MI[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      MM[r] = __fmul_rn((float)(-(i0 + 1 + r)), P.egq);
      GD[r] = IM[r] = DG[r] = MI[r] = HHG_NEG;
    }
    float dtMM = __fmul_rn((float)(-i0), P.egq), dtDG = HHG_NEG, dtMI = HHG_NEG, dtGD = HHG_NEG,
          dtIM = HHG_NEG;
    float best = HHG_NEG;
    int bi = 0, bj = 0;

    SlotRegs nA{}, nB{};
    if (s > 0) { ld_slot(bnd + 32, nA); ld_slot(bnd + 64, nB); }

    // small strips leave room for a register prefetch of the next column pair (one iteration ahead)
    constexpr bool kRegPrefetch = (R <= 8);
    float4 nxA[7], nxB[7];
    if (kRegPrefetch) {
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        nxA[k] = __ldg(tc + k);
        nxB[k] = __ldg(tc + (size_t)(min(2, Lt) - 1) * 7 + k);
      }
    }

    mbar_wait(bar, parity);
    parity ^= 1u;

    for (int j = 1; j <= Lmax; j += 2) {
      ColOps tA, tB;
      if (kRegPrefetch) {
        unpack_col(tA, nxA);
        unpack_col(tB, nxB);
        const float4* sa = tc + (size_t)(min(j + 2, Lt) - 1) * 7;
        const float4* sb = tc + (size_t)(min(j + 3, Lt) - 1) * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) { nxA[k] = __ldg(sa + k); nxB[k] = __ldg(sb + k); }
      } else {
        load_col(tA, tc + (size_t)(min(j, Lt) - 1) * 7);
        load_col(tB, tc + (size_t)(min(j + 1, Lt) - 1) * 7);
        const char* nxt = reinterpret_cast<const char*>(tc + (size_t)(min(j + 2, Lt) - 1) * 7);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128));
      }
      // ---- boundary row i0 at columns j and j+1
      float aMM, aDG, aMI, aGD, aIM, bMM, bDG, bMI, bGD, bIM;
      if (s == 0) {
        aMM = __fmul_rn((float)(-j), P.egt);
        bMM = __fmul_rn((float)(-(j + 1)), P.egt);
        aDG = aMI = aGD = aIM = bDG = bMI = bGD = bIM = HHG_NEG;
      } else {
        while (((nA.tag ^ tag_in) | nA.p0 | nA.p1) != 0u) { __nanosleep(20); ld_slot(bnd + (size_t)j * 32, nA); }
        while (((nB.tag ^ tag_in) | nB.p0 | nB.p1) != 0u) { __nanosleep(20); ld_slot(bnd + (size_t)(j + 1) * 32, nB); }
        aMM = nA.mm; aDG = nA.dg; aMI = nA.mi; aGD = nA.gd; aIM = nA.im;
        bMM = nB.mm; bDG = nB.dg; bMI = nB.mi; bGD = nB.gd; bIM = nB.im;
        if (j + 2 <= Lmax) { ld_slot(bnd + (size_t)(j + 2) * 32, nA); ld_slot(bnd + (size_t)(j + 3) * 32, nB); }
      }
      uint32_t cowA = 0, cowB = 0;
      if (CELLOFF) { cowA = __ldg(co + (size_t)j * 32); cowB = __ldg(co + (size_t)(j + 1) * 32); }

      // column A = j: diag (i0, j-1), up (i0, j); column B = j+1: diag (i0, j), up (i0, j+1)
      float dAMM = dtMM, dAGD = dtGD, dAIM = dtIM, dADG = dtDG, dAMI = dtMI;
      float uAMM = aMM, uADG = aDG, uAMI = aMI;
      float dBMM = aMM, dBGD = aGD, dBIM = aIM, dBDG = aDG, dBMI = aMI;
      float uBMM = bMM, uBDG = bDG, uBMI = bMI;
      float bcA = (j <= Lt) ? best : INFINITY;

      uint32_t wordA = SS ? 0u : (tA.ss & P.zero), wordB = SS ? 0u : (tB.ss & P.zero);
      float lastA_mm = 0.f, lastA_dg = 0.f, lastA_mi = 0.f, lastA_gd = 0.f, lastA_im = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        QRow q;
#pragma unroll
        for (int k = 0; k < 5; ++k) q.q[k] = qs[r * 7 + k];
        {
          const float4 qa = qs[r * 7 + 5], qb = qs[r * 7 + 6];
          q.m2m = qa.x; q.m2d = qa.y; q.d2m = qa.z; q.d2d = qa.w;
          q.i2m = qb.x; q.i2i = qb.y; q.m2i = qb.z; q.ss = __float_as_uint(qb.w);
        }
        const float oMM = MM[r], oGD = GD[r], oIM = IM[r], oDG = DG[r], oMI = MI[r];   // (i, j-1)
        float mmA, gdA, imA, dgA, miA, mmB, gdB, imB, dgB, miB;
        uint32_t bA, bB;
        dp_cell<LOCAL, SS, CELLOFF>(tA, q, s33, P.ssw, P.shift, P.one2, dAMM, dAGD, dAIM, dADG, dAMI, oMM, oGD,
                                    oIM, uAMM, uADG, uAMI, (cowA >> r) & 1u, mmA, gdA, imA, dgA, miA, bA);
        dp_cell<LOCAL, SS, CELLOFF>(tB, q, s33, P.ssw, P.shift, P.one2, dBMM, dBGD, dBIM, dBDG, dBMI, mmA, gdA,
                                    imA, uBMM, uBDG, uBMI, (cowB >> r) & 1u, mmB, gdB, imB, dgB, miB, bB);
        // running maximum, column A (column B is examined after the row loop from MM[])
        if (mmA >= bcA) {
          const int i = i0 + 1 + r;
          const bool cand = (i <= P.Lq) && (LOCAL || i == P.Lq || j == Lt);
          if (cand && (mmA > best || i < bi)) { best = mmA; bi = i; bj = j; bcA = mmA; }
        }
        wordA |= bA << (8 * (r & 3));
        wordB |= bB << (8 * (r & 3));
        if ((r & 3) == 3) {
          uint32_t* dst = btj + (size_t)((i0 >> 2) + (r >> 2)) * bt_row_stride + (size_t)j * 32;
          __stcs(dst, wordA);
          __stcs(dst + 32, wordB);
          wordA = 0; wordB = 0;
        }
        dAMM = oMM; dAGD = oGD; dAIM = oIM; dADG = oDG; dAMI = oMI;
        uAMM = mmA; uADG = dgA; uAMI = miA;
        dBMM = mmA; dBGD = gdA; dBIM = imA; dBDG = dgA; dBMI = miA;
        uBMM = mmB; uBDG = dgB; uBMI = miB;
        MM[r] = mmB; GD[r] = gdB; IM[r] = imB; DG[r] = dgB; MI[r] = miB;
        if (r == R - 1) { lastA_mm = mmA; lastA_dg = dgA; lastA_mi = miA; lastA_gd = gdA; lastA_im = imA; }
      }
      dtMM = bMM; dtDG = bDG; dtMI = bMI; dtGD = bGD; dtIM = bIM;

      // running maximum, column B: one test on the max of the R values, rows examined only if needed
      {
        const float bcB0 = (j + 1 <= Lt) ? best : INFINITY;
        float cm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) cm[r] = MM[r];
#pragma unroll
        for (int n = R; n > 1; n = (n + 1) / 2) {
#pragma unroll
          for (int r = 0; r < n / 2; ++r) cm[r] = fmaxf(cm[r], cm[r + (n + 1) / 2]);
        }
        if (cm[0] >= bcB0) {
          float bcB = bcB0;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float mm = MM[r];
            if (mm >= bcB) {
              const int i = i0 + 1 + r;
              const bool cand = (i <= P.Lq) && (LOCAL || i == P.Lq || (j + 1) == Lt);
              if (cand && (mm > best || i < bi)) { best = mm; bi = i; bj = j + 1; bcB = mm; }
            }
          }
        }
      }

      if (!last_strip) {
        st_slot(bnd + (size_t)j * 32, lastA_mm, lastA_dg, lastA_mi, lastA_gd, lastA_im, tag_out);
        st_slot(bnd + (size_t)(j + 1) * 32, MM[R - 1], DG[R - 1], MI[R - 1], GD[R - 1], IM[R - 1], tag_out);
      }
    }
    const size_t o = ((size_t)job * P.nstrips + s) * 32 + lane;
    P.strip_score[o] = best;
    P.strip_ij[o] = (bi << 16) | bj;
    __syncwarp();
  }
}

}  // namespace hhg
