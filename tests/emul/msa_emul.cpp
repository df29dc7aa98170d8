// tests/emul/msa_emul.cpp -- TEST INFRASTRUCTURE: runs the product's alignment -> HMM kernels
// (hh-suite_b200/csrc/hhg_msa.cuh, unmodified source) on the CPU through tests/emul/cuda_emul_mw.h.
// Built by tests/test_msa_emul_cpu.py:
//   g++ -O1 -std=c++20 -ffp-contract=off -fPIC -shared -pthread -DHHG_EMUL -o tests/emul/libmsaemul.so tests/emul/msa_emul.cpp
#include "cuda_emul_mw.h"

#include <cfloat>
#include <xmmintrin.h>
#include <cstdio>
#include <cstdlib>

#include "../../hh-suite_b200/csrc/hhg_msa.cuh"
#include "../../hh-suite_b200/csrc/hhg_crf.cuh"

using namespace hhg;

// One A3M record -> raw HMM, the steps of msa_chunk_run (hhg_api.cu) with host memory.  threads: block size of the
// emulated launches (a multiple of 32).  lg2 / dif: the fast_log2 tables (1025 floats each).
extern "C" int emul_msa_to_hmm(const char* rec, long long len, const int* ip /*maxseq,maxcol,maxres,max_seqid,coverage,qid,Ndiff,wg,M,Mgaps*/,
                               float qsc, const float* S, const float* pb, const float* lg2, const float* dif, int threads,
                               int* dims, signed char* keep_out, float* wg_out, float* f, float* tr, float* neff,
                               float* neff_hmm) {
  MsaHost H;
  const std::string msg = MsaScanner::parse(rec, len, ip[0], ip[1], ip[2], &H, ip[8], ip[9]);
  if (!msg.empty()) return -1;
  const int N = H.N_in, L = H.L;
  MsaDesc d{N, L, H.stride, H.kfirst, 0, 0, 0, 0};
  std::vector<int> in_(N), inkk(N), sp(N), acc(N), Ncnt(L + 2), Nmax(L + 2), idw(L + 2), ni((size_t)(L + 2) * 21);
  std::vector<float> wg(N), F((size_t)(L + 2) * 20), TR((size_t)(L + 2) * 7, 0.f), nm(L + 2), nif(L + 2), nd(L + 2), nseg(L + 2, 0.f);
  int nfil = 0, status = 0, counter = 0;
  float nh = 0.f;
  std::vector<int8_t> keep(H.keep);
  MsaArrays A{};
  A.desc = &d; A.X = H.X.data(); A.keep = keep.data(); A.display = H.display.data();
  A.first = H.first.data(); A.last = H.last.data(); A.nres = H.nres.data(); A.ksort = H.ksort.data();
  A.in_ = in_.data(); A.inkk = inkk.data(); A.seqid_prev = sp.data(); A.acc = acc.data();
  A.Ncnt = Ncnt.data(); A.Nmax = Nmax.data(); A.idmaxwin = idw.data(); A.wg = wg.data();
  std::vector<int> ins_k(H.ins_k.empty() ? 1 : H.ins_k.size());
  std::vector<uint16_t> ins_cnt(ins_k.size());
  std::copy(H.ins_k.begin(), H.ins_k.end(), ins_k.begin());
  std::copy(H.ins_cnt.begin(), H.ins_cnt.end(), ins_cnt.begin());
  A.ins_off = H.ins_off.data(); A.ins_k = ins_k.data(); A.ins_cnt = ins_cnt.data();
  A.n_filtered = &nfil; A.status = &status;
  A.f = F.data(); A.tr = TR.data(); A.neff_m = nm.data(); A.neff_i = nif.data(); A.neff_d = nd.data(); A.neff_seg = nseg.data();
  A.neff_hmm = &nh;
  MsaFilterParams P{};
  P.max_seqid = ip[3]; P.coverage = ip[4]; P.qid = ip[5]; P.Ndiff = ip[6]; P.qsc = qsc;
  if (S) memcpy(P.S, S, sizeof(P.S));
  if (getenv("EMUL_TRACE")) fprintf(stderr, "filter\n");
  emul_launch(1, threads, k_msa_filter, A, P);
  if (getenv("EMUL_TRACE")) fprintf(stderr, "weights\n");
  emul_launch(1, threads, k_msa_weights, A, ni.data());
  std::vector<float> rcp(MSA_RCP_N);
  for (int m = 0; m < MSA_RCP_N; m += 4) {
    const __m128 v = _mm_set_ps((float)(m + 3), (float)(m + 2), (float)(m + 1), (float)m);
    _mm_storeu_ps(&rcp[m], _mm_rcp_ps(v));
  }
  if (getenv("EMUL_TRACE")) fprintf(stderr, "mstate\n");
  const int nblk = 2;
  std::vector<int> cnt((size_t)nblk * (L + 2) * 24);
  std::vector<float> wc((size_t)nblk * (L + 2) * 24 + 4), wi((size_t)nblk * N);
  std::vector<uint8_t> member((size_t)nblk * N);
  long long item_off = 0;
  float* wc_al = wc.data();
  while ((uintptr_t)wc_al & 15) ++wc_al;
  emul_launch(nblk, threads, k_msa_mstate, A, 1, (const long long*)&item_off, (long long)L, &counter, cnt.data(), wc_al,
              wi.data(), member.data(), L, N, (const float*)rcp.data(), pb, ip[7], lg2, dif);
  if (getenv("EMUL_TRACE")) fprintf(stderr, "finish\n");
  emul_launch(1, threads, k_msa_finish, A, pb, ip[7], lg2, dif);
  dims[0] = L; dims[1] = N; dims[2] = nfil; dims[3] = status;
  memcpy(keep_out, keep.data(), N);
  memcpy(wg_out, wg.data(), (size_t)N * 4);
  memcpy(f, F.data(), (size_t)(L + 2) * 80);
  memcpy(tr, TR.data(), (size_t)(L + 1) * 28);
  memcpy(neff, nm.data(), (size_t)(L + 1) * 4);
  memcpy(neff + (L + 1), nif.data(), (size_t)(L + 1) * 4);
  memcpy(neff + 2 * (L + 1), nd.data(), (size_t)(L + 1) * 4);
  *neff_hmm = nh;
  return 0;
}


// k_crf_scores (hh-suite_b200/csrc/hhg_crf.cuh) on host memory: w[W*20*K] ([window][aa][state]), bias[K], counts[L*20]
extern "C" void emul_crf_scores(int L, int K, int W, const double* w, const double* bias, const double* counts, double* score) {
  emul_launch2((unsigned)((K + 255) / 256), (unsigned)L, 256, k_crf_scores, L, K, W, w, bias, counts, score);
}
