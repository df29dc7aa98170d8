#!/usr/bin/env python
"""bench.py -- Viterbi GCUPS (query_L x sum(target_L) / s) of the B200 hot path.

Workload (BASELINE.json configs[1], the one the metric is quoted on at one GPU): a synthetic query
profile L=400 against 100,000 synthetic profile HMMs per GPU (lengths lognormal, median 200, clipped
[30,2000]), Viterbi only (forward pass + backtrace of every target, as ViterbiRunner::alignment does).
With N GPUs every rank holds its own 100k-target shard (the DB is sharded by target, no data-path
collective) and the ranks exchange their top-K hit records with one NCCL all_gather per step.

    python bench.py --gpus N --steps K --warmup W            (driver; torchrun for N > 1)
    python bench.py --impl reference ...                      (the reference's AVX2 Viterbi on host cores)

value  : whole-job GCUPS, database resident in HBM, device-timed (CUDA events, max over ranks)
e2e    : the same through the host-buffer C-ABI call (hhg_query_set + hhg_viterbi_search): per step the
         query profile and the target-id list go H2D from pinned memory, hits and paths come back D2H.
roofline : algorithmic bytes (112 B per target column + 1 B per DP cell + 32 B per hit) / forward-kernel
         time, against the measured HBM peak in MEASURED_PEAKS.json.
cpu_baseline : the reference's own Viterbi::Align + Backtrace (oracle/_ref, AVX2, all host threads) on a
         bounded sample of the same shard.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOPK = 500          # realign_max of the reference (src/hhdecl.cpp): records exchanged per rank


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--targets", type=int, default=100000, help="targets per GPU")
    ap.add_argument("--lq", type=int, default=400)
    ap.add_argument("--cpu-sample", type=int, default=4000, help="targets in the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=16000, help="targets per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prefilter", action="store_true")
    return ap.parse_args()


def workload(args, rank):
    from hhsuite_b200 import synth
    qp, qtr, qss, qpav, qcols = synth.query_profile(args.lq, seed=1)
    db = synth.prepared_db(args.targets, seed=1000 + rank, query_cols=qcols, planted=64, fast=True)
    return (qp, qtr, qss, qpav), db


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        # nvidia-smi needs ~0.2 s to start, so it is launched before the warm-up; only samples that arrived inside
        # the timed region count
        t0 = getattr(self, "t_begin", 0.0)
        t1 = getattr(self, "t_end", float("inf"))
        inside = [(ts, ln) for ts, ln in self.lines if t0 <= ts <= t1 + 0.02]
        for ts, ln in (inside or self.lines[-3:]):      # a very short region may fall between two samples
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(args, qprof, db, sample, threads=None):
    """Reference AVX2 Viterbi::Align + Backtrace on `sample` targets of the shard, all host threads."""
    threads = threads or (os.cpu_count() or 1)
    qp, qtr, qss, qpav = qprof
    n = min(sample, len(db["L"]))
    from hhsuite_b200 import synth  # noqa: F401
    sub = dict(L=db["L"][:n], p=db["p"], tr=db["tr"], p_off=db["p_off"][:n], tr_off=db["tr_off"][:n])
    try:
        from oracle.binding import RefShim
        R = RefShim(nocontxt=True, maxres=max(4096, int(db["L"].max()) + 8))
        R.set_query(qp, qtr, qpav, None)
        R.viterbi_bench(dict(sub, L=sub["L"][:64], p_off=sub["p_off"][:64], tr_off=sub["tr_off"][:64]), threads)  # warm
        # best of 3: the GPU boxes' hosts are shared and the OpenMP timing is noisy (3.6 .. 13 GCUPS observed)
        sec, cells = min((R.viterbi_bench(sub, threads, with_backtrace=True, repeats=1)[:2] for _ in range(3)),
                         key=lambda x: x[0])
        return dict(value=cells / sec / 1e9, unit="GCUPS", cores=threads, kind="reference",
                    sample=f"first {n} targets of the rank-0 shard (sum L={int(sub['L'].sum())}), Viterbi::Align+Backtrace "
                           f"only, AVX2 no-FMA build of the unmodified reference, OpenMP dynamic over 8-target batches",
                    seconds=sec)
    except (FileNotFoundError, OSError):
        from oracle.binding import Oracle
        O = Oracle()
        n = min(n, 200)
        t0 = time.time()
        cells = 0
        for k in range(n):
            L = int(db["L"][k])
            tp = db["p"][db["p_off"][k]:db["p_off"][k] + L + 2]
            ttr = db["tr"][db["tr_off"][k]:db["tr_off"][k] + L + 1]
            O.viterbi(qp, qtr, tp, ttr)
            cells += args.lq * L
        sec = time.time() - t0
        return dict(value=cells / sec / 1e9, unit="GCUPS", cores=1, kind="port",
                    sample=f"first {n} targets, scalar C restatement (oracle/hh_oracle.c), 1 thread", seconds=sec)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    qprof, db = workload(argparse.Namespace(**{**vars(args), "targets": max(args.ref_sample, 64)}), 0)
    threads = os.cpu_count() or 1
    times = []
    cb = None
    for s in range(args.warmup + args.steps):
        cb = cpu_baseline(args, qprof, db, args.ref_sample, threads)
        if s >= args.warmup:
            times.append(cb["seconds"])
    cells = float(args.lq) * float(db["L"][:args.ref_sample].sum())
    t = float(np.mean(times))
    val = cells / t / 1e9
    cb = dict(cb, value=val)
    cb.pop("seconds", None)
    print(json.dumps({
        "impl": "reference", "metric": "Viterbi GCUPS (query_L x sum target_L / s)", "value": val, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"query L={args.lq} vs synthetic profile HMMs (median L=200), Viterbi only; each step = "
                               f"{args.ref_sample} targets of the 100k shard on the host CPU (bounded sample)",
                   "threads": threads},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def prefilter_figure(hh, ctx, lq, n=200000):
    """Secondary figure (BASELINE configs[2] stage 1): cs219 ungapped prefilter on a synthetic 200k shard."""
    import torch
    from hhsuite_b200 import synth
    cs = synth.cs219_db(n, seed=3)
    rng = np.random.default_rng(1)
    prof = rng.integers(30, 66, (220, lq), dtype=np.uint8)
    db = hh.CsDB(ctx, cs["L"], cs["off"], cs["seq"])
    db.run(prof, 50, upload=True)
    ctx.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        db.run(prof, 50, upload=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    cells = float(lq) * float(cs["L"].sum())
    db.close()
    return {"kernel": "k_prefilter_ungapped (DPX s16x2)", "sequences": n, "query_L": lq, "ms": ms,
            "tcells_per_s": cells / ms / 1e9, "unit": "1e12 byte-cells/s"}


def cuda_array(ptr, nbytes):
    """A torch uint8 view of device memory owned by the library (no copy)."""
    import torch

    class _W:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_W(), device="cuda")


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import hhsuite_b200 as hh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    qprof, db_h = workload(args, rank)
    qp, qtr, qss, qpav = qprof
    # a dedicated (non-default) torch stream: the library launches on it and torch.cuda.Event brackets it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = hh.Context(device=local_rank, stream=stream.cuda_stream)
    assert stream.cuda_stream != 0
    ctx.set_query(qp, qtr)
    db = hh.TargetDB(ctx, db_h["L"], db_h["p"], db_h["tr"], db_h["p_off"], db_h["tr_off"])
    plan = hh.Plan(ctx, db)
    cells_rank = plan.cells
    n = plan.n
    base_id = rank * n

    # device views for the top-K exchange: HitRec = 10 x 4 bytes, score first
    def topk_exchange(hits_dev_i32):
        scores = hits_dev_i32[:, 0].view(torch.float32)
        k = min(TOPK, n)
        top = torch.topk(scores, k)
        rec = torch.cat([(top.indices + base_id).to(torch.int32).unsqueeze(1), hits_dev_i32[top.indices]], dim=1)
        if world > 1:
            out = torch.empty((world * k, rec.shape[1]), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(out, rec.contiguous())
            sc = out[:, 1].view(torch.float32)
            best = torch.topk(sc, k)
            return out[best.indices]
        return rec

    hits_ptr = ctx.L.hhg_plan_hits_devptr(plan.h)
    hits_dev = cuda_array(hits_ptr, n * 40).view(torch.int32).view(n, 10)

    def step():
        plan.run()
        return topk_exchange(hits_dev)

    sampler = ClockSampler(local_rank)
    sampler.start()                      # before the warm-up: nvidia-smi takes a moment to deliver its first sample
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    l0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        merged = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_end()
    clocks = sampler.stop()
    launches = ctx.launches - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_step = ms_total / args.steps
    total_cells = cells_rank * world
    gcups = total_cells / (ms_step * 1e-3) / 1e9

    # ---- per-kernel roofline figure (forward kernel timed alone with CUDA events on the same stream)
    kt = [plan.run_timed() for _ in range(3)]
    ms_vit = float(np.mean([a for a, b in kt])); ms_bt = float(np.mean([b for a, b in kt]))
    peaks, peak_src = measured_peaks()
    ach = plan.alg_bytes / (ms_vit * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_viterbi launch of exactly this workload
    # (ncu --set full, profiles/r1_ncu_viterbi_bench_workload.txt); only valid for the default workload
    traffic = 32.75e9 if (args.targets == 100000 and args.lq == 400) else None
    roofline = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                "kernel": "k_viterbi<16,local>", "kernel_ms": ms_vit, "backtrace_ms": ms_bt,
                "algorithmic_bytes_per_launch": plan.alg_bytes,
                "note": "exact-fp32 recurrence is FP32-issue-bound, not HBM-bound (DESIGN.md, SURVEY 8d): "
                        "kernel GCUPS=%.1f" % (cells_rank / (ms_vit * 1e-3) / 1e9)}

    # ---- end to end through the host-buffer C-ABI call, pinned host buffers
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    qp_pin, qtr_pin = pin(qp), pin(qtr)
    ids_pin = pin(np.arange(n, dtype=np.int32))
    hits_pin = torch.empty(n * 40, dtype=torch.uint8).pin_memory().numpy().view(hh.capi.HIT_DTYPE)
    paths_pin = torch.empty(plan.path_cap, dtype=torch.uint8).pin_memory().numpy()
    h2d = qp_pin.nbytes + qtr_pin.nbytes + ids_pin.nbytes

    def e2e_step():
        ctx.set_query(qp_pin, qtr_pin)
        hits, paths = hh.viterbi_search(ctx, db, ids=ids_pin, hits=hits_pin, paths=paths_pin)
        k = min(TOPK, n)
        top = np.argpartition(-hits["score"], k - 1)[:k]
        if world > 1:
            rec = torch.from_numpy(np.concatenate([(top + base_id).astype(np.int32)[:, None],
                                                   hits[top].view(np.int32).reshape(k, 10)], axis=1)).to(dev)
            out = torch.empty((world * k, 11), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(out, rec)
            return out.cpu()
        return hits[top]

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    # D2H per step: the hit records + the tightly packed path strings (compacted on the device)
    d2h = hits_pin.nbytes + int(hits_pin["nsteps"].sum()) + (0 if world == 1 else 0)
    if world > 1:
        dist.barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_t = torch.tensor([(t1 - t0) / e2e_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_gcups = total_cells / float(e2e_t.item()) / 1e9

    out = {
        "metric": "Viterbi GCUPS (query_L x sum target_L / s)", "value": gcups, "unit": "GCUPS", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"query L={args.lq} vs {n} synthetic profile HMMs per GPU (lognormal lengths, median "
                               f"200, clip [30,2000]), Viterbi only: forward pass + backtrace of every target, "
                               f"top-{TOPK} hit records" + (" all-gathered over NCCL" if world > 1 else ""),
                   "targets_per_gpu": n, "query_L": args.lq, "sum_target_L_per_gpu": int(db_h["L"].sum()),
                   "parallelism": f"db-shard x{world}",
                   "l2": "inputs larger than L2 (1.8 GB of column records + 9 GB of backtrace bytes per step)",
                   "strip_rows": 16, "db_resident": True},
        "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": float(e2e_t.item()) * 1e3,
                "note": "hhg_query_set + hhg_viterbi_search with pinned host buffers (plan reused across queries, paths "
                        "compacted on the device before D2H); the target DB stays resident "
                        "on the GPU (loaded once, like the reference's mmap'd ffindex DB)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_prefilter:
        out["prefilter"] = prefilter_figure(hh, ctx, args.lq)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args, qprof, db_h, args.cpu_sample)
        cb.pop("seconds", None)
        out["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(out))
    plan.close(); db.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
