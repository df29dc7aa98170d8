// tests/emul/mac_emul.cpp -- TEST INFRASTRUCTURE: runs the product's MAC kernels (hh-suite_b200/csrc/hhg_mac.cuh,
// unmodified source) on the CPU through tests/emul/cuda_emul.h.  Built by tests/test_mac_emul_cpu.py:
//   g++ -O1 -std=c++20 -ffp-contract=off -fPIC -shared -pthread -o tests/emul/libmacemul.so tests/emul/mac_emul.cpp
#include "cuda_emul.h"

#include <cfloat>

namespace hhg {
struct alignas(16) ColRec {   // as in hh-suite_b200/csrc/hhg_kernels.cuh
  float p[20];
  float m2m, m2d, d2m, d2d, i2m, i2i, m2i;
  uint32_t ss;
};
static_assert(sizeof(ColRec) == 112, "ColRec");
}  // namespace hhg

namespace hhg {
alignas(16) unsigned char mac_smem[256 * 1024];   // the block's dynamic shared memory (blocks run one at a time)
}

#include "../../hh-suite_b200/csrc/hhg_mac.cuh"

using namespace hhg;

// -excl / -template_excl ranges for the following emul_mac_realign calls: q_lo[nq], q_hi[nq], t_lo[nt], t_hi[nt]
static std::vector<int> g_reg;
static int g_reg_nq = 0, g_reg_nt = 0;
extern "C" void emul_mac_set_regions(int nq, int nt, const int* reg) {
  g_reg.assign(reg, reg + 2 * (nq + nt)); g_reg_nq = nq; g_reg_nt = nt;
}

// One request, inputs in the reference layout (like the oracle's hho_mac_realign): prepared template p/(log2) tr,
// query p and LINEAR tr.  t_tr_lin: linear template transitions (boundary rows are set here like the host API does).
extern "C" int emul_mac_realign(int Lq, const float* q_p, const float* q_tr_lin, int Lt, const float* t_p,
                                const float* t_tr_log2, const float* t_tr_lin, int local, double Cshift, float mact,
                                const int* vit5, const int* vit_i, const int* vit_j, int n_excl, const int* excl_i,
                                const int* excl_j, int smem_bytes, int band_scan, int* res6, float* sum_of_probs,
                                double* pforward, int* out_i, int* out_j, uint8_t* out_states, float* out_post,
                                float* post_matrix) {
  // shard records of this one template from p / log2 tr (what k_pack_cols would build)
  std::vector<ColRec> cols((size_t)Lt);
  for (int j = 1; j <= Lt; ++j) {
    ColRec& r = cols[j - 1];
    for (int a = 0; a < 20; ++a) r.p[a] = t_p[(size_t)j * 20 + a];
    const float* t1 = t_tr_log2 + (size_t)(j - 1) * 7;
    const float* t0 = t_tr_log2 + (size_t)j * 7;
    r.m2m = t1[0]; r.m2d = t1[2]; r.d2m = t1[5]; r.d2d = t1[6]; r.i2m = t1[3]; r.i2i = t0[4]; r.m2i = t0[1]; r.ss = 0;
  }
  // query transitions with the boundary rows of initializeQueryHMMTransitions, template ones of initializeForAlignment
  std::vector<float> qtr(q_tr_lin, q_tr_lin + (size_t)(Lq + 1) * 7), ttr((size_t)(Lt + 1) * 7);
  qtr[1] = qtr[2] = qtr[3] = qtr[4] = qtr[5] = qtr[6] = 0.f;
  { float* e = qtr.data() + (size_t)Lq * 7; e[0] = 1.f; e[1] = e[2] = e[3] = e[4] = 0.f; e[5] = 1.f; e[6] = 0.f; }
  // gather kernel (log2 rows), then the host step of hhg_mac_realign with the caller's powf results
  std::vector<long long> rec0{0}, tr_off{0}, cell_off{0}, row_off{0}, path_off{0}, vit_off{0, vit5[4]}, excl_off{0, n_excl};
  std::vector<int> Ltv{Lt};
  emul_launch(emul_dim3(8, 1), 128, k_mac_gather_tr, 1, cols.data(), rec0.data(), Ltv.data(), tr_off.data(), ttr.data());
  for (int i = 1; i < Lt; ++i)
    for (int k = 0; k < 7; ++k) {
      if (ttr[(size_t)i * 7 + k] != t_tr_log2[(size_t)i * 7 + k]) return -10;          // gather must reproduce the rows
      ttr[(size_t)i * 7 + k] = t_tr_lin[(size_t)i * 7 + k];
    }
  { float* b = ttr.data(); b[0] = 1.f; b[1] = b[2] = b[3] = b[4] = b[5] = b[6] = 0.f;
    float* e = ttr.data() + (size_t)Lt * 7; e[0] = 1.f; e[1] = e[2] = e[3] = e[4] = 0.f; e[5] = 1.f; e[6] = 0.f; }
  const size_t ncell = (size_t)(Lq + 1) * (Lt + 1);
  std::vector<float> post(ncell, 0.f);
  std::vector<uint8_t> off(ncell, 0), bt(ncell, 0);
  std::vector<double> rows((size_t)11 * (Lt + 3) + (Lt + 3 + 7) / 8 + 1, 0.0), scale((size_t)Lq + 3, 0.0);
  const size_t cap = (size_t)Lq + Lt + 2;
  std::vector<int> oi(cap, 0), oj(cap, 0);
  std::vector<uint8_t> os(cap, 0);
  std::vector<float> op(cap, 0.f);
  MacHitOut out{};
  MacArgs A{};
  A.n = 1; A.Lq = Lq; A.local = local; A.mact = mact; A.Cshift = Cshift;
  A.q_p = q_p; A.q_tr = qtr.data(); A.cols = cols.data(); A.rec0 = rec0.data(); A.Lt = Ltv.data();
  A.t_tr = ttr.data(); A.tr_off = tr_off.data(); A.vit = vit5; A.vit_off = vit_off.data();
  A.vit_i = vit_i; A.vit_j = vit_j;
  A.excl_off = n_excl ? excl_off.data() : nullptr; A.excl_i = excl_i; A.excl_j = excl_j;
  A.cell_off = cell_off.data(); A.post = post.data(); A.off = off.data(); A.bt = bt.data();
  A.row_off = row_off.data(); A.rows = rows.data(); A.scale = scale.data(); A.out = &out;
  A.path_off = path_off.data(); A.out_i = oi.data(); A.out_j = oj.data(); A.out_states = os.data(); A.out_post = op.data();
  A.smem_rows = smem_bytes;
  A.band_scan = band_scan;
  if (g_reg_nq + g_reg_nt) { A.reg = g_reg.data(); A.reg_nq = g_reg_nq; A.reg_nt = g_reg_nt; }
  emul_launch(emul_dim3(1), 256, k_mac_band, A);
  emul_launch(emul_dim3(1), 32, k_mac_realign, A);
  res6[0] = out.i1; res6[1] = out.i2; res6[2] = out.j1; res6[3] = out.j2; res6[4] = out.nsteps; res6[5] = out.matched_cols;
  *sum_of_probs = out.sum_of_probs; *pforward = out.pforward;
  for (int s = 0; s <= out.nsteps; ++s) { out_i[s] = oi[s]; out_j[s] = oj[s]; out_states[s] = os[s]; out_post[s] = op[s]; }
  std::memcpy(post_matrix, post.data(), ncell * sizeof(float));
  return out.nsteps;
}
