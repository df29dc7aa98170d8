// oracle/ref_gpu_adapter.cpp -- TEST INFRASTRUCTURE: the drop-in check.
//
// Links the UNMODIFIED reference (oracle/_ref/libhhref.a) and the product library (libhhg.so) into one
// binary that runs the reference's own HHalign front half (HMM::Read, PrepareQueryHMM, HMMSimd::MapOneHMM,
// HHFileEntry::getTemplateHMM, PrepareTemplateHMM steps; src/hhalign.cpp:590-645) and then aligns the same
// template list twice:
//   (1) with the reference's ViterbiRunner::alignment (src/hhviterbirunner.cpp:75, AVX2, OpenMP), and
//   (2) with GpuViterbiRunner::alignment below = the adapter of INTEGRATION.md section 2 on the C-ABI,
// and compares every Hit the two produce: Hit.score / score_ss (bits), i1,i2,j1,j2, nsteps, matched_cols,
// irep, lastrep and the whole path (i[], j[], states[]) for every alternative alignment.
// It contains no reference code; it includes the reference headers at build time and is compiled by
// oracle/ref_build.mk into oracle/_ref/hh_dropin_check (shipped prebuilt to the GPU box).
//
// usage: hh_dropin_check [--hhm-loader] [--mac] [--gpus N] <query.hhm> <template.hhm> [more templates ...]   exit 0 = identical
//   --gpus N      shard the templates over N GPUs (template k -> GPU k mod N), one host thread + hhg_ctx + hhg_comm
//                 per GPU; every Hit must still equal the reference's, and the hit list merged over NCCL by
//                 hhg_plan_topk / hhg_plan_topk_paths (K = all templates, key = Hit.score) must be the reference's
//                 first-alignment hits in (score descending, template index ascending) order, paths included
//   --hhm-loader  templates are loaded by hhg_db_create_hhm from the HHM text (or, when the template files are A3M
//                 alignments, by hhg_db_create_a3m) instead of the reference's preparation
//   --mac         additionally realign every hit: PosteriorDecoderRunner::executeComputation vs hhg_mac_realign
//                 (written in round 1, first exercised on a GPU in round 2; the MAC parity of round 1 is established
//                 by tests/test_mac_gpu.py against the same reference code through oracle/ref_shim.cpp)

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <set>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "hhdecl.h"
#include "hhhmm.h"
#include "hhhmmsimd.h"
#include "hhviterbi.h"
#include "hhviterbimatrix.h"
#include "hhviterbirunner.h"
#include "hhmatrices.h"
#include "hhfunc.h"
#include "hhdatabase.h"
#include "hhhit.h"
#include "hhposteriordecoder.h"
#include "hhposteriordecoderrunner.h"
#include "hhposteriormatrix.h"
#include "hhprefilter.h"
#include "ffindexdatabase.h"
#include "hash.h"
#include "cs219.lib.h"
#undef private
#undef protected

#include "../include/hhg.h"

namespace {

struct GpuHit {
  int target, irep, lastrep;
  float score, score_ss;
  int i1, i2, j1, j2, nsteps, matched_cols;
  std::vector<int> i, j;
  std::vector<char> states;
};

#define HHG_CHECK(call)                                                        \
  do {                                                                         \
    if ((call) != HHG_OK) {                                                    \
      fprintf(stderr, "hhg error: %s (%s)\n", hhg_last_error(), #call);        \
      exit(4);                                                                 \
    }                                                                          \
  } while (0)

// The adapter of INTEGRATION.md: same arguments as ViterbiRunner::alignment; templates are addressed by
// their index in `dbfiles`.
class GpuViterbiRunner {
 public:
  explicit GpuViterbiRunner(int device = 0) { HHG_CHECK(hhg_ctx_create(device, nullptr, &ctx_)); }
  ~GpuViterbiRunner() { if (comm_) hhg_comm_destroy(comm_); if (db_) hhg_db_destroy(db_); hhg_ctx_destroy(ctx_); }

  // multi-GPU: this runner owns the templates global_ids[0..] of the whole list and is rank `rank` of `world`
  void JoinComm(int rank, int world, const void* uid, const std::vector<int32_t>& global_ids, int n_total) {
    HHG_CHECK(hhg_comm_create(ctx_, rank, world, uid, &comm_));
    gids_ = global_ids;
    n_total_ = n_total;
  }
  std::vector<hhg_topk_rec> merged;      // the merged first-alignment hit list (every rank holds the same)
  std::vector<uint8_t> merged_paths;     // [merged.size() x merged_width]
  int merged_width = 0;

  // one-time shard upload: what PrepareTemplateHMM leaves BEFORE the query-dependent null model
  void Upload(Parameters& par, std::vector<HHEntry*>& entries, float* pb, const float S[20][20],
              const float Sim[20][20], const float R[20][20]) {
    std::vector<int32_t> L;
    std::vector<int64_t> p_off, tr_off, ss_off;
    std::vector<float> p, tr, pav;
    std::vector<uint8_t> ss;
    all_have_ss_ = true;
    HMM t(MAXSEQDIS, par.maxres);
    for (HHEntry* e : entries) {
      seqlen_.push_back(e->sequence_length);
      int format = 0;
      e->getTemplateHMM(par, 1, par.qsc_db, format, pb, S, Sim, &t);
      t.AddTransitionPseudocounts(par.gapd, par.gape, par.gapf, par.gapg, par.gaph, par.gapi, par.gapb, par.gapb);
      t.PreparePseudocounts(R);
      t.AddAminoAcidPseudocounts(par.pc_hhm_nocontext_mode, par.pc_hhm_nocontext_a, par.pc_hhm_nocontext_b,
                                 par.pc_hhm_nocontext_c);
      t.CalculateAminoAcidBackground(pb);
      L.push_back(t.L);
      p_off.push_back((int64_t)p.size()); tr_off.push_back((int64_t)tr.size()); ss_off.push_back((int64_t)ss.size());
      if (t.nss_pred < 0) all_have_ss_ = false;
      neff_.push_back(t.Neff_HMM);
      has_pred_.push_back(t.nss_pred >= 0 ? 1 : 0);
      has_dssp_.push_back(t.nss_dssp >= 0 ? 1 : 0);
      for (int i = 0; i <= t.L + 1; ++i) {
        p.insert(p.end(), t.p[i], t.p[i] + 20);
        ss.push_back(t.nss_pred >= 0 ? (uint8_t)(t.ss_pred[i] * MAXCF + t.ss_conf[i]) : 0);
      }
      for (int i = 0; i <= t.L; ++i) tr.insert(tr.end(), t.tr[i], t.tr[i] + 7);
      pav.insert(pav.end(), t.pav, t.pav + 20);
    }
    HHG_CHECK(hhg_db_create_raw(ctx_, (int)L.size(), L.data(), p_off.data(), tr_off.data(), ss_off.data(), p.data(),
                                tr.data(), ss.data(), pav.data(), &db_));
    L_ = L;
  }

  // alternative upload: hand the HHM TEXT of the files to the library (hhg_db_create_hhm); nothing of the
  // reference's HMM::Read / PrepareTemplateHMM runs for the templates on this side.
  void UploadText(Parameters& par, const std::vector<std::string>& files, const float R[20][20]) {
    std::string data;
    std::vector<int64_t> off, len;
    all_have_ss_ = true;
    for (const std::string& f : files) {
      FILE* fp = fopen(f.c_str(), "rb");
      if (!fp) { perror(f.c_str()); exit(2); }
      std::string rec;
      char buf[65536];
      size_t n;
      while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) rec.append(buf, n);
      fclose(fp);
      rec.push_back('\0');                                        // like an ffindex entry
      off.push_back((int64_t)data.size()); len.push_back((int64_t)rec.size());
      int32_t L = 0, has_ss = 0;
      const bool is_msa = rec[0] == '>' || rec[0] == '#';         // the format test of HHEntry::getTemplateHMM (:441)
      if (files.front() == f) msa_ = is_msa;
      if (is_msa != msa_) { fprintf(stderr, "templates must be all HHM or all A3M\n"); exit(2); }
      if (is_msa) {
        // alignment branch: Read / Compress / Filter / FrequenciesAndTransitions run inside hhg_db_create_a3m; the
        // Viterbi stage reads templates with GLOBAL sequence weights (char wg = 1, src/hhviterbirunner.cpp:143)
        hhg_msa_params_default(&mp_);
        mp_.maxseq = par.maxseq; mp_.maxcol = par.maxcol; mp_.maxres = par.maxres;
        mp_.max_seqid = par.max_seqid_db; mp_.coverage = par.coverage_db; mp_.qid = par.qid_db; mp_.qsc = par.qsc_db;
        mp_.Ndiff = par.Ndiff_db; mp_.wg = 1;
        int32_t N = 0;
        HHG_CHECK(hhg_a3m_scan(rec.data(), (int64_t)rec.size(), &mp_, &L, &N, &has_ss));
        if (!has_ss) all_have_ss_ = false;
        L_.push_back(L);
        // Neff_HMM of the record for Hit.Neff_HMM: one extra pass through the alignment kernels
        std::vector<float> f((size_t)(L + 2) * 20), tr((size_t)(L + 1) * 7), ne((size_t)3 * (L + 1));
        int32_t dims[6]; float nh = 0;
        HHG_CHECK(hhg_msa_to_hmm(ctx_, rec.data(), (int64_t)rec.size(), &mp_, S_, pb_, L, N, dims, nullptr, nullptr, f.data(),
                                 tr.data(), ne.data(), &nh, nullptr));
        neff_.push_back(nh);
      } else {
      HHG_CHECK(hhg_hhm_scan(rec.data(), (int64_t)rec.size(), &L, &has_ss));
      if (!has_ss) all_have_ss_ = false;
      L_.push_back(L);
      }
      if (!is_msa) {
        // Neff_HMM of the record (NEFF line) through the library's tokeniser; ss_dssp lines do not occur in the test sets
        std::vector<int32_t> f((size_t)L * 20), trn((size_t)(L + 1) * 10), nul(20);
        std::vector<uint8_t> ssb(L);
        float neff = 0; int32_t has_pc = 0;
        HHG_CHECK(hhg_hhm_parse(rec.data(), (int64_t)rec.size(), L, f.data(), trn.data(), ssb.data(), nul.data(), &neff, &has_pc));
        neff_.push_back(neff);
      }
      has_pred_.push_back(has_ss ? 1 : 0);
      has_dssp_.push_back(0);
      seqlen_.push_back(seqlen_override_.empty() ? par.maxres : seqlen_override_[has_pred_.size() - 1]);
      data += rec;
    }
    hhg_prep_params pp = {par.gapb, par.gapd, par.gape, par.gapf, par.gapg, par.gaph, par.gapi,
                          par.pc_hhm_nocontext_mode, par.pc_hhm_nocontext_a, par.pc_hhm_nocontext_b,
                          par.pc_hhm_nocontext_c};
    if (msa_)
      HHG_CHECK(hhg_db_create_a3m(ctx_, (int)files.size(), data.data(), off.data(), len.data(), &mp_, S_, pb_, &pp, &R[0][0], &db_));
    else
      HHG_CHECK(hhg_db_create_hhm(ctx_, (int)files.size(), data.data(), off.data(), len.data(), &pp, &R[0][0], &db_));
  }
  bool msa_ = false;
  hhg_msa_params mp_;
  const float* S_ = nullptr;      // substitution matrix / background the reference holds when it reads the templates
  const float* pb_ = nullptr;

  std::vector<GpuHit> alignment(Parameters& par, HMMSimd* q_simd, int n_targets, float* pb,
                                const float S33[NSSPRED][MAXCF][NSSPRED][MAXCF]) {
    HMM* q = q_simd->GetHMM(0);
    std::vector<float> qp((size_t)(q->L + 2) * 20), qtr((size_t)(q->L + 1) * 7);
    std::vector<uint8_t> qss(q->L + 2, 0);
    for (int i = 0; i <= q->L + 1; ++i) {
      memcpy(&qp[(size_t)i * 20], q->p[i], 80);
      if (q->nss_pred >= 0) qss[i] = (uint8_t)(q->ss_pred[i] * MAXCF + q->ss_conf[i]);
    }
    for (int i = 0; i <= q->L; ++i) memcpy(&qtr[(size_t)i * 7], q->tr[i], 28);
    // the query is sent once with the ss term armed if it could ever be used; per batch the runner decides
    const bool q_pred = q->nss_pred >= 0, q_dssp = q->nss_dssp >= 0;
    const int can_ss = (par.ssm == 2 && q_pred) ? 1 : 0;
    hhg_params hp = {par.loc, par.egq, par.egt, par.shift, par.ssw, can_ss, par.corr, par.ssm};
    HHG_CHECK(hhg_query_set(ctx_, q->L, qp.data(), qtr.data(), qss.data(), &S33[0][0][0][0], &hp));
    HHG_CHECK(hhg_db_apply_null_model(ctx_, db_, q->pav, pb, par.columnscore));
    {
      // par.exclstr / par.template_exclstr: "a-b,c-d" -> ranges, read like the reference does (pairs of integers)
      auto parse = [](const char* str, std::vector<int32_t>& lo, std::vector<int32_t>& hi) {
        std::vector<int> v;
        for (const char* c = str; c && *c;) {
          if (*c >= '0' && *c <= '9') { int x = 0; while (*c >= '0' && *c <= '9') x = x * 10 + (*c++ - '0'); v.push_back(x); }
          else ++c;
        }
        for (size_t k = 0; k + 1 < v.size(); k += 2) { lo.push_back(v[k]); hi.push_back(v[k + 1]); }
      };
      std::vector<int32_t> ql, qh, tl, th;
      parse(par.exclstr, ql, qh);
      parse(par.template_exclstr, tl, th);
      HHG_CHECK(hhg_set_excluded_regions(ctx_, (int)ql.size(), ql.data(), qh.data(), (int)tl.size(), tl.data(), th.data()));
    }

    std::vector<int32_t> ids(n_targets);
    for (int k = 0; k < n_targets; ++k) ids[k] = k;
    std::map<int, std::pair<std::vector<int32_t>, std::vector<int32_t>>> excl;   // accumulated paths per target
    std::vector<GpuHit> ret;
    for (int alignment = 0; alignment < par.altali && !ids.empty(); alignment++) {
      const unsigned all = (unsigned)ids.size();
      unsigned block = all;
      if (alignment == 0 && par.early_stopping_filter) block = 2000;          // src/hhviterbirunner.cpp:109-111
      std::vector<int32_t> next;
      for (unsigned start = 0; start < all; start += block) {
        const unsigned cnt = std::min(all - start, block);
        // the chunk is sorted by HHEntry::sequence_length, descending, with std::sort exactly like the reference
        // (:117-119; same comparison sequence => same permutation, also among equal keys), then cut into batches of
        // VECSIZE_FLOAT = 8 lanes; the ss mode of a batch is the consensus over its lanes (:14-22)
        std::sort(ids.begin() + start, ids.begin() + start + cnt,
                  [&](int32_t a, int32_t b) { return seqlen_[a] > seqlen_[b]; });
        std::vector<int32_t> grp[2];                                             // [0]: plain kernels, [1]: *AndSS
        for (unsigned b0 = start; b0 < start + cnt; b0 += 8) {
          int consensus = 0xFF;
          const unsigned b1 = std::min(b0 + 8, start + cnt);
          for (unsigned k = b0; k < b1; ++k) {
            const int t = ids[k];
            int mode = 0;                                                        // HMM::computeScoreSSMode
            if (q_pred && has_dssp_[t]) mode |= HMM::PRED_DSSP;
            if (q_dssp && has_pred_[t]) mode |= HMM::DSSP_PRED;
            if (q_pred && has_pred_[t]) mode |= HMM::PRED_PRED;
            consensus &= mode;
          }
          int ss_mode = consensus & HMM::PRED_DSSP;
          ss_mode = (ss_mode == 0) ? consensus & HMM::DSSP_PRED : 0;
          ss_mode = (ss_mode == 0) ? consensus & HMM::PRED_PRED : 0;
          const int use_ss = (par.ssm == 2 && ss_mode != HMM::NO_SS_INFORMATION) ? 1 : 0;   // Viterbi::Align, src/hhviterbi.cpp:177
          for (unsigned k = b0; k < b1; ++k) grp[use_ss].push_back(ids[k]);
        }
        std::vector<float> chunk_score; std::vector<int32_t> chunk_L; std::vector<float> chunk_neff;
        for (int g = 0; g < 2; ++g) {
          const std::vector<int32_t>& gi = grp[g];
          if (gi.empty()) continue;
          if (can_ss) HHG_CHECK(hhg_set_use_ss(ctx_, g));
          std::vector<hhg_hit> hits(gi.size());
          size_t cap = 0;
          for (int id : gi) cap += (size_t)q->L + L_[id] + 2;
          std::vector<uint8_t> paths(cap);
          std::vector<int64_t> eoff(gi.size() + 1, 0);
          std::vector<int32_t> ei, ej;
          if (alignment > 0) {
            for (size_t k = 0; k < gi.size(); ++k) {
              auto& e = excl[gi[k]];
              ei.insert(ei.end(), e.first.begin(), e.first.end());
              ej.insert(ej.end(), e.second.begin(), e.second.end());
              eoff[k + 1] = (int64_t)ei.size();
            }
          }
          HHG_CHECK(hhg_viterbi_search(ctx_, db_, (int)gi.size(), gi.data(), hits.data(), paths.data(), paths.size(),
                                       alignment ? eoff.data() : nullptr, ei.data(), ej.data()));
          if (comm_ && alignment == 0 && g == (grp[1].empty() ? 0 : 1) && grp[g == 0 ? 1 : 0].empty()) {
            // merged hit list of the first alignment round over all GPUs (collective: every rank calls it); only
            // formed when the round is one search (homogeneous ss mode, one chunk)
            merged.resize(n_total_);
            int m = 0;
            std::vector<int32_t> gg(gi.size());
            for (size_t k = 0; k < gi.size(); ++k) gg[k] = gids_[gi[k]];
            HHG_CHECK(hhg_plan_topk(ctx_, hhg_ctx_last_plan(ctx_), comm_, n_total_, 1, 0, gg.data(), merged.data(), &m));
            merged.resize(m);
            merged_width = 1;
            for (const hhg_topk_rec& r : merged) merged_width = std::max(merged_width, (int)r.hit.nsteps);
            merged_paths.assign((size_t)m * merged_width, 0);
            HHG_CHECK(hhg_plan_topk_paths(ctx_, hhg_ctx_last_plan(ctx_), comm_, m, merged.data(), merged_width,
                                          merged_paths.data()));
          }
          for (size_t k = 0; k < gi.size(); ++k) {
            const hhg_hit& h = hits[k];
            GpuHit gh;
            gh.target = gids_.empty() ? gi[k] : gids_[gi[k]]; gh.irep = alignment + 1;   // :257 (global template index)
            gh.lastrep = (h.hit_score <= par.smin) ? 1 : 0;                  // :36
            gh.score = h.hit_score; gh.score_ss = h.score_ss;
            gh.i1 = h.i1; gh.i2 = h.i2; gh.j1 = h.j1; gh.j2 = h.j2; gh.nsteps = h.nsteps; gh.matched_cols = h.matched_cols;
            gh.i.assign(h.nsteps + 1, 0); gh.j.assign(h.nsteps + 1, 0); gh.states.assign(h.nsteps + 1, 0);
            int i = h.i2, j = h.j2;                                           // replay of Viterbi::Backtrace
            for (int st = 1; st <= h.nsteps; ++st) {
              const char c = (char)paths[h.path_off + st - 1];
              gh.i[st] = i; gh.j[st] = j; gh.states[st] = c;
              if (st < h.nsteps) {
                if (c == 2) { --i; --j; } else if (c == 3 || c == 4) { --j; } else { --i; }
              }
            }
            ret.push_back(gh);
            chunk_score.push_back(h.hit_score); chunk_L.push_back(L_[gi[k]]); chunk_neff.push_back(neff_[gi[k]]);
            if (h.hit_score > par.smin) {                                     // :260-268
              next.push_back(gi[k]);
              auto& e = excl[gi[k]];
              for (int st = 1; st < h.nsteps; ++st) { e.first.push_back(gh.i[st]); e.second.push_back(gh.j[st]); }
            }
          }
        }
        if (alignment == 0 && par.early_stopping_filter) {                     // :178-188
          const float sum = hhg_early_stop_sum((int)chunk_score.size(), chunk_score.data(), chunk_L.data(), chunk_neff.data(),
                                               q->L, q->Neff_HMM, par.prefilter, par.dbsize, par.alphaa, par.alphab,
                                               par.alphac, par.prefilter_evalue_thresh);
          if (sum < cnt * par.filter_thresh) { early_stopped_at = (int)(start + cnt); break; }
        }
      }
      ids.swap(next);
    }
    return ret;
  }
  int early_stopped_at = -1;   // number of database entries after which the first round stopped (-1: it did not)

  // The adapter of INTEGRATION.md section 2b: PosteriorDecoderRunner::executeComputation on the C-ABI.  `vit` are the
  // Viterbi hits to realign; q is the query AFTER the reference put it into linear transition space.
  struct GpuMac { int target, irep, i1, i2, j1, j2, nsteps, matched_cols; float sum_of_probs; double pforward;
                  std::vector<int> i, j; std::vector<char> states; std::vector<float> post; };
  std::vector<GpuMac> realign(Parameters& par, HMM* q, const std::vector<GpuHit>& vit) {
    std::vector<float> qp((size_t)(q->L + 2) * 20), qtr((size_t)(q->L + 1) * 7);
    for (int i = 0; i <= q->L + 1; ++i) memcpy(&qp[(size_t)i * 20], q->p[i], 80);
    for (int i = 0; i <= q->L; ++i) memcpy(&qtr[(size_t)i * 7], q->tr[i], 28);
    HHG_CHECK(hhg_mac_query_set(ctx_, q->L, qp.data(), qtr.data()));
    std::map<int, std::vector<const GpuHit*>> by_target;                 // alignments_map + sort by irep (:54-66)
    for (const GpuHit& h : vit) if (h.nsteps > 0) by_target[h.target].push_back(&h);
    for (auto& kv : by_target)
      std::sort(kv.second.begin(), kv.second.end(), [](const GpuHit* a, const GpuHit* b) { return a->irep < b->irep; });
    std::map<int, std::pair<std::vector<int32_t>, std::vector<int32_t>>> alt;   // Hit.alt_i / alt_j per template
    std::vector<GpuMac> out;
    for (size_t round = 0;; ++round) {
      std::vector<const GpuHit*> batch;
      for (auto& kv : by_target) if (kv.second.size() > round) batch.push_back(kv.second[round]);
      if (batch.empty()) break;
      std::vector<int32_t> target, vitv, vi, vj, ei, ej;
      std::vector<int64_t> voff(1, 0), eoff(1, 0);
      size_t cap = 0;
      for (const GpuHit* h : batch) {
        target.push_back(h->target);
        const int32_t v5[5] = {h->i1, h->i2, h->j1, h->j2, h->nsteps};
        vitv.insert(vitv.end(), v5, v5 + 5);
        vi.insert(vi.end(), h->i.begin() + 1, h->i.end()); vj.insert(vj.end(), h->j.begin() + 1, h->j.end());
        voff.push_back((int64_t)vi.size());
        auto& a = alt[h->target];
        ei.insert(ei.end(), a.first.begin(), a.first.end()); ej.insert(ej.end(), a.second.begin(), a.second.end());
        eoff.push_back((int64_t)ei.size());
        cap += (size_t)q->L + L_[h->target] + 2;
      }
      if (ei.empty()) { ei.push_back(0); ej.push_back(0); }
      std::vector<hhg_mac_hit> mh(batch.size());
      std::vector<int32_t> oi(cap), oj(cap);
      std::vector<uint8_t> os(cap);
      std::vector<float> op(cap);
      hhg_mac_params mp = {par.loc, par.shift, par.mact};
      HHG_CHECK(hhg_mac_realign(ctx_, db_, (int)batch.size(), target.data(), vitv.data(), voff.data(), vi.data(), vj.data(),
                                round ? eoff.data() : nullptr, ei.data(), ej.data(), &mp, mh.data(), oi.data(), oj.data(),
                                os.data(), op.data(), cap));
      for (size_t k = 0; k < batch.size(); ++k) {
        const hhg_mac_hit& m = mh[k];
        GpuMac g;
        g.target = batch[k]->target; g.irep = batch[k]->irep;
        g.i1 = m.i1; g.i2 = m.i2; g.j1 = m.j1; g.j2 = m.j2; g.nsteps = m.nsteps; g.matched_cols = m.matched_cols;
        g.sum_of_probs = m.sum_of_probs; g.pforward = m.pforward;
        g.i.assign(oi.begin() + m.path_off, oi.begin() + m.path_off + m.nsteps + 1);
        g.j.assign(oj.begin() + m.path_off, oj.begin() + m.path_off + m.nsteps + 1);
        g.states.assign(os.begin() + m.path_off, os.begin() + m.path_off + m.nsteps + 1);
        g.post.assign(op.begin() + m.path_off, op.begin() + m.path_off + m.nsteps + 1);
        auto& a = alt[g.target];                                         // backtraceMAC pushes every visited (i,j)
        if (m.nsteps) { a.first.insert(a.first.end(), g.i.begin() + 1, g.i.end()); a.second.insert(a.second.end(), g.j.begin() + 1, g.j.end()); }
        else { a.first.push_back(g.i[0]); a.second.push_back(g.j[0]); }
        out.push_back(g);
      }
    }
    return out;
  }

 private:
  hhg_ctx* ctx_ = nullptr;
  hhg_db* db_ = nullptr;
  hhg_comm* comm_ = nullptr;
  std::vector<int32_t> gids_;
  int n_total_ = 0;
  std::vector<int32_t> L_;
  std::vector<int> seqlen_, has_pred_, has_dssp_;   // HHEntry::sequence_length, nss_pred >= 0, nss_dssp >= 0 per template
  std::vector<float> neff_;                         // Neff_HMM per template
  bool all_have_ss_ = false;
 public:
  std::vector<int> seqlen_override_;                // UploadText: the sequence_length the reference's entries carry
};


// The adapter of INTEGRATION.md section 3: Prefilter::prefilter_db (src/hhprefilter.h:80-85, .cpp:430-606) with the same
// signature, both scoring stages on the C-ABI, the selection logic in between restated: stage-1 list chosen on the
// device (hhg_prefilter_select = length correction :477, sort, keep rule :489-506), stage-2 E-values (:529), coarse cut
// (:530), ascending sort by (E-value, index) (:545), keep rule (:547-558), de-duplication by name, split into
// new / old hits by previous_hits (name without extension + "__1", :561-588), maxnumdb cap (:590).
class GpuPrefilter {
 public:
  GpuPrefilter(hhg_ctx* ctx, FFindexDatabase* cs219_database) : ctx_(ctx) {
    // init_prefilter (:314-335): one entry per ffindex record, length = entry->length - 1 (the record's NUL)
    ffindex_index_t* idx = cs219_database->db_index;
    n_ = (int)idx->n_entries;
    std::vector<int32_t> L(n_);
    std::vector<int64_t> off(n_);
    std::vector<uint8_t> seq;
    for (int k = 0; k < n_; ++k) {
      ffindex_entry_t* e = ffindex_get_entry_by_index(idx, k);
      const unsigned char* d = (const unsigned char*)ffindex_get_data_by_entry(cs219_database->db_data, e);
      L[k] = (int32_t)e->length - 1;
      off[k] = (int64_t)seq.size();
      seq.insert(seq.end(), d, d + L[k]);
      names_.push_back(e->name);
    }
    L_ = L;
    HHG_CHECK(hhg_csdb_create(ctx_, n_, L.data(), off.data(), seq.data(), &db_));
    // the 219 column states in linear space (Prefilter ctor :28-47: cs219.lib + TransformToLin)
    FILE* fin = fmemopen((void*)cs219_lib, cs219_lib_len, "r");
    cs::ContextLibrary<cs::AA> lib(fin);
    fclose(fin);
    cs::TransformToLin(lib);
    lib219_.resize(219 * 20);
    for (int k = 0; k < 219; ++k)
      for (int a = 0; a < 20; ++a) lib219_[k * 20 + a] = lib[k].probs[0][a];
  }
  ~GpuPrefilter() { hhg_csdb_destroy(db_); }

  void prefilter_db(HMM* q_tmp, Hash<Hit>* previous_hits, const int threads, const int prefilter_gap_open,
                    const int prefilter_gap_extend, const int prefilter_score_offset, const int prefilter_bit_factor,
                    const double prefilter_evalue_thresh, const double prefilter_evalue_coarse_thresh,
                    const int preprefilter_smax_thresh, const int min_prefilter_hits, const int maxnumdb,
                    const float R[20][20], std::vector<std::pair<int, std::string> >& new_prefilter_hits,
                    std::vector<std::pair<int, std::string> >& old_prefilter_hits) {
    (void)threads; (void)R;
    const int LQ = q_tmp->L;
    std::vector<float> qp((size_t)(LQ + 2) * 20);
    for (int i = 0; i <= LQ + 1; ++i) memcpy(&qp[(size_t)i * 20], q_tmp->p[i], 80);
    std::vector<uint8_t> prof((size_t)220 * LQ);
    HHG_CHECK(hhg_prefilter_build_profile(LQ, qp.data(), q_tmp->pav, lib219_.data(), prefilter_score_offset,
                                          prefilter_bit_factor, prof.data()));
    // stage 1 on the whole shard, list chosen on the device
    HHG_CHECK(hhg_prefilter_ungapped_run(ctx_, db_, LQ, prof.data(), prefilter_score_offset, 1));
    std::vector<int32_t> first(n_), first_score(n_);
    int nfirst = 0;
    HHG_CHECK(hhg_prefilter_select(ctx_, db_, LQ, prefilter_bit_factor, preprefilter_smax_thresh, min_prefilter_hits,
                                   first.data(), first_score.data(), n_, &nfirst));
    first.resize(nfirst);
    // stage 2: gapped byte SW on the stage-1 list, E-values with the database size
    std::vector<int32_t> sw(nfirst), len(nfirst);
    if (nfirst)
      HHG_CHECK(hhg_prefilter_sw(ctx_, db_, nfirst, first.data(), LQ, prof.data(), prefilter_gap_open + prefilter_gap_extend,
                                 prefilter_gap_extend, prefilter_score_offset, sw.data()));
    for (int k = 0; k < nfirst; ++k) len[k] = L_[first[k]];
    std::vector<double> ev(nfirst);
    if (nfirst) HHG_CHECK(hhg_prefilter_evalues(nfirst, sw.data(), len.data(), n_, LQ, prefilter_bit_factor, ev.data()));
    std::vector<std::pair<double, int> > hits;
    for (int k = 0; k < nfirst; ++k)
      if (ev[k] < prefilter_evalue_coarse_thresh) hits.push_back(std::make_pair(ev[k], (int)first[k]));
    // :545 sorts with comparePair, whose arguments are std::pair<int,int>: the E-value of a std::pair<double,int>
    // is TRUNCATED TO int by the implicit conversion, so the order is ascending ((int)evalue, index)
    std::sort(hits.begin(), hits.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) {
      const int ea = (int)a.first, eb = (int)b.first;
      if (ea != eb) return ea < eb;
      return a.second < b.second;
    });
    size_t keep = 0;
    for (; keep < hits.size(); ++keep)                                         // :547-558
      if (!((int)keep < min_prefilter_hits || hits[keep].first <= prefilter_evalue_thresh)) break;
    hits.resize(keep);
    std::set<std::string> doubled;
    int count_dbs = 0;
    for (const auto& h : hits) {
      ++count_dbs;
      const std::string& db_name = names_[h.second];
      if (!doubled.count(db_name)) {
        doubled.insert(db_name);
        std::string name = db_name;                                            // RemoveExtension
        const size_t dot = name.rfind('.');
        if (dot != std::string::npos) name.resize(dot);
        const std::string key = name + "__1";
        std::pair<int, std::string> result(L_[h.second], db_name);
        if (previous_hits->Contains(key.c_str())) old_prefilter_hits.push_back(result);
        else new_prefilter_hits.push_back(result);
      }
      if (count_dbs >= maxnumdb) break;
    }
  }

 private:
  hhg_ctx* ctx_;
  hhg_csdb* db_ = nullptr;
  int n_ = 0;
  std::vector<int32_t> L_;
  std::vector<std::string> names_;
  std::vector<float> lib219_;
};

uint32_t bits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }

}  // namespace

int main(int argc, char** argv) {
  bool text_loader = false, with_mac = false, real_lengths = false;
  int gpus = 1, hhblits_dbsize = 0, maxnumdb = 20000;
  const char* prefilter_db_base = nullptr;
  char* exclstr = nullptr;
  char* template_exclstr = nullptr;
  std::vector<std::string> previous;
  while (argc > 1 && argv[1][0] == '-' && argv[1][1] == '-') {
    if (!strcmp(argv[1], "--hhm-loader")) text_loader = true;
    else if (!strcmp(argv[1], "--mac")) with_mac = true;
    else if (!strcmp(argv[1], "--gpus") && argc > 2) { gpus = atoi(argv[2]); --argc; ++argv; }
    else if (!strcmp(argv[1], "--real-lengths")) real_lengths = true;
    else if (!strcmp(argv[1], "--excl") && argc > 2) { exclstr = argv[2]; --argc; ++argv; }
    else if (!strcmp(argv[1], "--template-excl") && argc > 2) { template_exclstr = argv[2]; --argc; ++argv; }
    else if (!strcmp(argv[1], "--prefilter") && argc > 2) { prefilter_db_base = argv[2]; --argc; ++argv; }
    else if (!strcmp(argv[1], "--maxnumdb") && argc > 2) { maxnumdb = atoi(argv[2]); --argc; ++argv; }
    else if (!strcmp(argv[1], "--previous") && argc > 2) {
      std::stringstream ss(argv[2]); std::string item;
      while (std::getline(ss, item, ',')) previous.push_back(item);
      --argc; ++argv;
    }
    else if (!strcmp(argv[1], "--hhblits") && argc > 2) { hhblits_dbsize = atoi(argv[2]); --argc; ++argv; }
    else break;
    --argc; ++argv;
  }
  if (gpus > 1 && with_mac) { fprintf(stderr, "--mac is checked on one GPU\n"); return 2; }
  if (prefilter_db_base && argc < 2) { fprintf(stderr, "usage: %s --prefilter <db>_cs219 query.hhm\n", argv[0]); return 2; }
  if (!prefilter_db_base && argc < 3) { fprintf(stderr, "usage: %s [--hhm-loader] [--mac] query.hhm template.hhm [...]\n", argv[0]); return 2; }
  Log::reporting_level() = WARNING;
  const char* pargv[] = {"hhalign"};
  Parameters par(1, pargv);
  par.nocontxt = 1;
  par.maxres = 4096;
  par.threads = 2;
  par.exclstr = exclstr;                  // -excl / -template_excl (src/hhdecl.h; ViterbiRunner::exclude_regions)
  par.template_exclstr = template_exclstr;
  if (hhblits_dbsize > 0) {               // what hhblits sets on top of hhsearch (src/hhblits.cpp:87-88): early stopping
    par.early_stopping_filter = true;
    par.filter_thresh = 0.01;
    par.prefilter = 1;
    par.dbsize = hhblits_dbsize;
  }
  float pb[21] __attribute__((aligned(32)));
  float P[20][20] __attribute__((aligned(32))), R[20][20] __attribute__((aligned(32)));
  float S[20][20] __attribute__((aligned(32))), Sim[20][20] __attribute__((aligned(32)));
  static float S73[NDSSP][NSSPRED][MAXCF], S37[NSSPRED][MAXCF][NDSSP], S33[NSSPRED][MAXCF][NSSPRED][MAXCF];
  SetSubstitutionMatrix(par.matrix, pb, P, R, S, Sim);
  SetSecStrucSubstitutionMatrix(par.ssa, S73, S37, S33);

  // ---- query: src/hhalign.cpp:610-626
  HMM* q = new HMM(MAXSEQDIS, par.maxres);
  HMMSimd q_vec(par.maxres);
  {
    FILE* f = fopen(argv[1], "r");
    if (!f) { perror(argv[1]); return 2; }
    char path[NAMELEN];
    Pathname(path, argv[1]);
    q->Read(f, par.maxcol, par.nseqdis, pb, path);
    fclose(f);
  }
  char input_format = 0;
  PrepareQueryHMM(par, input_format, q, nullptr, nullptr, pb, R);
  q_vec.MapOneHMM(q);

  if (prefilter_db_base) {
    // ---- seam 2: Prefilter::prefilter_db of the reference vs GpuPrefilter::prefilter_db on the same cs219 ffindex
    const std::string data = std::string(prefilter_db_base) + ".ffdata", index = std::string(prefilter_db_base) + ".ffindex";
    FFindexDatabase csdb(data.c_str(), index.c_str(), false);
    Hash<Hit> previous_hits;
    previous_hits.New(1631, Hit());
    for (const std::string& nm : previous) { std::string key = nm + "__1"; Hit h; previous_hits.Add((char*)key.c_str(), h); }
    std::vector<std::pair<int, std::string> > rn, ro, gn, go;
    Prefilter ref_pf("", &csdb);
    ref_pf.prefilter_db(q, &previous_hits, par.threads, par.prefilter_gap_open, par.prefilter_gap_extend,
                        par.prefilter_score_offset, par.prefilter_bit_factor, par.prefilter_evalue_thresh,
                        par.prefilter_evalue_coarse_thresh, par.preprefilter_smax_thresh, par.min_prefilter_hits, maxnumdb, R,
                        rn, ro);
    hhg_ctx* pctx = nullptr;
    HHG_CHECK(hhg_ctx_create(0, nullptr, &pctx));
    {
      GpuPrefilter gpu_pf(pctx, &csdb);
      gpu_pf.prefilter_db(q, &previous_hits, par.threads, par.prefilter_gap_open, par.prefilter_gap_extend,
                          par.prefilter_score_offset, par.prefilter_bit_factor, par.prefilter_evalue_thresh,
                          par.prefilter_evalue_coarse_thresh, par.preprefilter_smax_thresh, par.min_prefilter_hits, maxnumdb,
                          R, gn, go);
    }
    hhg_ctx_destroy(pctx);
    const bool same = rn == gn && ro == go;
    if (!same) {
      printf("PREFILTER MISMATCH: reference %zu new / %zu old, gpu %zu new / %zu old\n", rn.size(), ro.size(), gn.size(), go.size());
      for (size_t k = 0; k < std::min(rn.size(), gn.size()); ++k)
        if (rn[k] != gn[k]) { printf("  first difference at new[%zu]: ref (%d,%s) gpu (%d,%s)\n", k, rn[k].first, rn[k].second.c_str(), gn[k].first, gn[k].second.c_str()); break; }
    }
    printf("hh_dropin_check --prefilter: %zu sequences, query L=%d: %zu new + %zu old prefilter hits: %s\n",
           (size_t)csdb.db_index->n_entries, q->L, rn.size(), ro.size(), same ? "identical lists" : "MISMATCH");
    return same ? 0 : 1;
  }

  // entries carry a sequence_length like the ffindex entries of a real database (the runner sorts each chunk by it);
  // --real-lengths: the template's own length, otherwise the same value for all
  std::vector<HHEntry*> entries;
  std::vector<int> seqlens;
  for (int a = 2; a < argc; ++a) {
    int len = par.maxres;
    if (real_lengths) {
      FILE* fp = fopen(argv[a], "rb");
      if (!fp) { perror(argv[a]); return 2; }
      std::string rec; char buf[65536]; size_t n;
      while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) rec.append(buf, n);
      fclose(fp);
      int32_t L = 0, has_ss = 0;
      if (rec[0] == '>' || rec[0] == '#') {
        hhg_msa_params mp; hhg_msa_params_default(&mp);
        int32_t N = 0;
        HHG_CHECK(hhg_a3m_scan(rec.c_str(), (int64_t)rec.size() + 1, &mp, &L, &N, &has_ss));
      } else {
        HHG_CHECK(hhg_hhm_scan(rec.c_str(), (int64_t)rec.size() + 1, &L, &has_ss));
      }
      len = L;
    }
    seqlens.push_back(len);
    entries.push_back(new HHFileEntry(argv[a], len));
  }

  // ---- (1) the reference runner
  std::vector<ViterbiMatrix*> mats(par.threads);
  for (auto& m : mats) { m = new ViterbiMatrix(); m->AllocateBacktraceMatrix(q->L, par.maxres); }
  std::vector<HHblitsDatabase*> nodb;
  ViterbiRunner ref_runner(mats.data(), nodb, par.threads);
  std::vector<Hit> ref = ref_runner.alignment(par, &q_vec, entries, par.qsc_db, pb, S, Sim, R, par.ssm, S73, S33, S37);

  // ---- (2) the GPU adapter
  GpuViterbiRunner gpu_runner;
  std::vector<GpuHit> gpu;
  std::vector<hhg_topk_rec> merged;
  std::vector<uint8_t> merged_paths;
  int merged_width = 0;
  if (gpus <= 1) {
    if (text_loader) {
      std::vector<std::string> files(argv + 2, argv + argc);
      gpu_runner.seqlen_override_ = seqlens;
      gpu_runner.S_ = &S[0][0]; gpu_runner.pb_ = pb;
      gpu_runner.UploadText(par, files, R);
    } else {
      gpu_runner.Upload(par, entries, pb, S, Sim, R);
    }
    gpu = gpu_runner.alignment(par, &q_vec, (int)entries.size(), pb, S33);
  } else {
    // the shard of GPU r = templates r, r+gpus, ...; uploads run on this thread (they call into the reference),
    // the searches run concurrently, one host thread per GPU, and meet in hhg_plan_topk
    char uid[128];
    HHG_CHECK(hhg_comm_unique_id(uid));
    std::vector<std::unique_ptr<GpuViterbiRunner>> runners;
    std::vector<std::vector<int32_t>> gid(gpus);
    for (int r = 0; r < gpus; ++r) {
      runners.emplace_back(new GpuViterbiRunner(r));
      std::vector<HHEntry*> mine;
      std::vector<std::string> files;
      for (size_t k = r; k < entries.size(); k += gpus) { gid[r].push_back((int32_t)k); mine.push_back(entries[k]); files.push_back(argv[2 + k]); }
      if (mine.empty()) { fprintf(stderr, "--gpus %d needs at least %d templates\n", gpus, gpus); return 2; }
      for (size_t k = r; k < entries.size(); k += gpus) runners[r]->seqlen_override_.push_back(seqlens[k]);
      if (text_loader) runners[r]->UploadText(par, files, R); else runners[r]->Upload(par, mine, pb, S, Sim, R);
    }
    std::vector<std::vector<GpuHit>> part(gpus);
    std::vector<std::thread> th;
    for (int r = 0; r < gpus; ++r)
      th.emplace_back([&, r] {
        runners[r]->JoinComm(r, gpus, uid, gid[r], (int)entries.size());
        part[r] = runners[r]->alignment(par, &q_vec, (int)gid[r].size(), pb, S33);
      });
    for (auto& t : th) t.join();
    for (int r = 0; r < gpus; ++r) gpu.insert(gpu.end(), part[r].begin(), part[r].end());
    merged = runners[0]->merged; merged_paths = runners[0]->merged_paths; merged_width = runners[0]->merged_width;
    for (int r = 1; r < gpus; ++r)
      if (runners[r]->merged.size() != merged.size() ||
          memcmp(runners[r]->merged.data(), merged.data(), merged.size() * sizeof(hhg_topk_rec)) ||
          runners[r]->merged_paths != merged_paths) { printf("MISMATCH: rank %d holds a different merged list\n", r); return 1; }
  }

  // ---- compare
  std::map<HHEntry*, int> index;
  for (size_t k = 0; k < entries.size(); ++k) index[entries[k]] = (int)k;
  std::map<std::pair<int, int>, const GpuHit*> gmap;
  for (const GpuHit& g : gpu) gmap[{g.target, g.irep}] = &g;
  int bad = 0;
  if (ref.size() != gpu.size()) { printf("MISMATCH: %zu reference hits vs %zu gpu hits\n", ref.size(), gpu.size()); ++bad; }
  int maxrep = 0;
  for (Hit& h : ref) {
    const int t = index[h.entry];
    maxrep = std::max(maxrep, (int)h.irep);
    auto it = gmap.find({t, h.irep});
    if (it == gmap.end()) { printf("MISMATCH: no gpu hit for template %d irep %d\n", t, h.irep); ++bad; continue; }
    const GpuHit& g = *it->second;
    bool ok = bits(h.score) == bits(g.score) && bits(h.score_ss) == bits(g.score_ss) && h.i1 == g.i1 && h.i2 == g.i2 &&
              h.j1 == g.j1 && h.j2 == g.j2 && h.nsteps == g.nsteps && h.matched_cols == g.matched_cols &&
              h.lastrep == g.lastrep;
    for (int s = 1; ok && s <= h.nsteps; ++s) ok = h.i[s] == g.i[s] && h.j[s] == g.j[s] && h.states[s] == g.states[s];
    if (!ok) {
      printf("MISMATCH: template %d (%s) irep %d: ref score %.6f (%d-%d,%d-%d, %d steps) gpu %.6f (%d-%d,%d-%d, %d steps)\n",
             t, entries[t]->getName(), h.irep, h.score, h.i1, h.i2, h.j1, h.j2, h.nsteps, g.score, g.i1, g.i2, g.j1,
             g.j2, g.nsteps);
      ++bad;
    }
  }
  if (gpus > 1) {
    // the NCCL-merged list must be the reference's first-alignment hits, best Hit.score first, template index on ties
    std::vector<Hit*> first;
    for (Hit& h : ref) if (h.irep == 1) first.push_back(&h);
    std::sort(first.begin(), first.end(), [&](Hit* a, Hit* b) {
      if (a->score != b->score) return a->score > b->score;
      return index[a->entry] < index[b->entry];
    });
    int mbad = first.size() != merged.size();
    for (size_t r = 0; !mbad && r < first.size(); ++r) {
      const Hit& h = *first[r];
      const hhg_topk_rec& m = merged[r];
      bool ok = m.target == index[h.entry] && m.owner == m.target % gpus && bits(m.hit.hit_score) == bits(h.score) &&
                m.hit.i1 == h.i1 && m.hit.i2 == h.i2 && m.hit.j1 == h.j1 && m.hit.j2 == h.j2 && m.hit.nsteps == h.nsteps;
      for (int s = 1; ok && s <= h.nsteps; ++s) ok = merged_paths[r * merged_width + s - 1] == (uint8_t)h.states[s];
      if (!ok) { printf("MERGE MISMATCH at rank-list position %zu: template %d vs reference template %d\n", r, m.target, index[h.entry]); ++mbad; }
    }
    printf("hh_dropin_check --gpus %d: merged list of %zu hits over NCCL: %s\n", gpus, merged.size(),
           mbad ? "MISMATCH" : "identical to the reference's sorted first-round hits, paths included");
    bad += mbad;
  }
  if (text_loader) printf("(templates loaded by hhg_db_create_hhm from the HHM text)\n");
  // ---- (3) optional: MAC realignment of all hits, reference runner vs C-ABI
  if (with_mac && !bad) {
    std::vector<PosteriorMatrix*> pms(par.threads);
    int Lmax = 0;
    for (Hit& h : ref) Lmax = std::max(Lmax, (int)h.L);
    for (auto& pm : pms) { pm = new PosteriorMatrix(); pm->allocateMatrix(q->L, Lmax + 2); }
    std::vector<Hit*> hit_ptrs;
    for (Hit& h : ref) if (h.nsteps > 0) hit_ptrs.push_back(&h);
    par.ssw = par.ssw_realign;                                           // src/hhalign.cpp:651
    PosteriorDecoderRunner prunner(pms.data(), mats.data(), par.threads, par.ssw, S73, S33, S37);
    prunner.executeComputation(*q, hit_ptrs, par, par.qsc_db, pb, S, Sim, R);   // puts q into linear space, realigns in place
    std::vector<GpuViterbiRunner::GpuMac> gm = gpu_runner.realign(par, q, gpu);
    std::map<std::pair<int, int>, const GpuViterbiRunner::GpuMac*> mmap;
    for (const auto& g : gm) mmap[{g.target, g.irep}] = &g;
    int mbad = 0;
    for (Hit* hp : hit_ptrs) {
      Hit& h = *hp;
      auto it = mmap.find({index[h.entry], h.irep});
      if (it == mmap.end()) { printf("MAC MISMATCH: no gpu result for template %d irep %d\n", index[h.entry], h.irep); ++mbad; continue; }
      const auto& g = *it->second;
      bool ok = h.i1 == g.i1 && h.i2 == g.i2 && h.j1 == g.j1 && h.j2 == g.j2 && h.nsteps == g.nsteps &&
                h.matched_cols == g.matched_cols && bits(h.sum_of_probs) == bits(g.sum_of_probs) && h.Pforward == g.pforward;
      for (int s = 1; ok && s <= h.nsteps; ++s)
        ok = h.i[s] == g.i[s] && h.j[s] == g.j[s] && h.states[s] == g.states[s] && bits(h.P_posterior[s]) == bits(g.post[s]);
      if (!ok) {
        printf("MAC MISMATCH: template %d irep %d: ref %d-%d,%d-%d %d steps sum %.6f Pf %.17g | gpu %d-%d,%d-%d %d steps sum %.6f Pf %.17g\n",
               index[h.entry], h.irep, h.i1, h.i2, h.j1, h.j2, h.nsteps, h.sum_of_probs, h.Pforward, g.i1, g.i2, g.j1, g.j2,
               g.nsteps, g.sum_of_probs, g.pforward);
        ++mbad;
      }
    }
    printf("hh_dropin_check --mac: %zu hits realigned: %s\n", hit_ptrs.size(), mbad ? "MISMATCH" : "all MAC alignments identical");
    bad += mbad;
  }
  if (hhblits_dbsize > 0)
    printf("hh_dropin_check --hhblits: early stop after %d of %zu database entries (reference aligned %zu first-round hits)\n",
           gpu_runner.early_stopped_at, entries.size(), (size_t)std::count_if(ref.begin(), ref.end(), [](const Hit& h) { return h.irep == 1; }));
  printf("hh_dropin_check: query L=%d, %zu templates, %zu hits (up to irep %d): %s\n", q->L, entries.size(), ref.size(),
         maxrep, bad ? "MISMATCH" : "all hits identical");
  return bad ? 1 : 0;
}
