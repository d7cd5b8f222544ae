#!/usr/bin/env python
"""bench.py — frame pairs/s (match + RANSAC) at 5 000 keypoints, 256-bit descriptors, 1 024 hypotheses.

Workload (BASELINE.json configs[1]): every consecutive pair of a synthetic monocular sequence with
5 000 keypoints per frame goes through match_features = Hamming kNN-2 + ratio test + RansacFilter
(1 024 hypotheses, threshold 10) + inlier copy-out. One STEP = one pass over this rank's sequence
(--frames frames, frames-1 pairs) in a fixed number of batched kernel launches. Per-rank inputs are
~205 MB (larger than the 126 MB L2), so consecutive steps do not find their inputs in L2.

  value  pairs/s, inputs resident in HBM, steps issued as a streaming caller does (vb_pairs_submit_d / vb_pairs_wait, two
         in flight: the counting of one step runs beside the matcher of the next), CUDA events on the launching stream;
         value_one_stream = the same steps one after the other through vb_pairs_run_d (round 1's measurement)
  e2e    pairs/s through the host-pointer C-ABI (vb_pairs_submit / vb_pairs_wait, three submissions in flight): pinned
         host buffers, H2D of every frame and D2H of results + compact matches inside the timed region, every step;
         e2e.frac_of_copy_ceiling compares it with plain cudaMemcpyAsync traffic of the same byte counts on this box.
         (e2e_blocking_call = the blocking vb_pairs_run, e2e_pageable = the same calls on pageable memory.)
  roofline      the dominant kernel of the step, k_knn2_tc4 (tcgen05 kind::mxf4), against the fp4 tensor rate measured
                on this GPU by vb_probe_tensor_peak (roofline.peak_measured_fp4); `counting` carries the RANSAC inlier
                counting stage (the north star's named kernel) with SURVEY 8d's 16 logical bytes per evaluation
  config3       BASELINE configs[2]: one pair, 20 000 keypoints, 128-d float descriptors, 4 096 hypotheses
  config4       BASELINE configs[3]: ONE 10 001-frame sequence cut into contiguous pair ranges over the ranks (strong scaling)
  score_sweep   BASELINE configs[4]: matches x hypotheses corners of the scoring entry point
  cpu_baseline  the C oracle (port of the reference path) on the box's host cores, bounded sample
  --impl reference   times that CPU path with all host threads instead of the GPU path

N > 1 (torchrun): every rank owns its own sequence shard on its own GPU (weak scaling), no data-path
collective; gloo is used only for the barrier and the max-over-ranks of the device time. e2e_multi is the same work
driven by ONE process (rank 0) through vb_multi — one host thread + context per GPU, results landing in one host array.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "frame pairs/sec (match+RANSAC) at 5k kpts"
UNIT = "pairs/s"
E2E_DEPTH = 3   # host-pointer submissions in flight: one uploading, two computing (the library allows three per context)
DTYPE = "u8 descriptors as e2m1 +-1 on tcgen05 kind::mxf4 (exact integer Hamming) + f32/f64 residual"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1025, help="frames per rank (pairs = frames-1)")
    ap.add_argument("--kpts", type=int, default=5000)
    ap.add_argument("--hyps", type=int, default=1024)
    ap.add_argument("--threshold", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="(kept for compatibility: the config-5 corners are always run)")
    ap.add_argument("--frames-total", type=int, default=10001, help="config 4: frames of the ONE sequence sharded over the ranks")
    ap.add_argument("--quick", action="store_true", help="skip config3 / config4 / score_sweep / kd-tree extras")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s, source). Driver-written MEASURED_PEAKS.json, else the profiling guide's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1614.4)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1614.4, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, pairs, kpts, hyps):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture of
    this same workload (profiles/ncu_traffic.json, written by tools/ncu_traffic.py), or None when the shapes differ. The
    captured launch may cover a different number of pairs than the launch bench.py times (every pair moves the same bytes:
    its two frames' operands in, its partial results out), so the figure is scaled to `pairs`."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        if (d["kpts"], d["hyps"]) == (kpts, hyps) and d["pairs"] > 0:
            return float(d["dram_bytes_per_launch"]) * pairs / d["pairs"]
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------
def cpu_oracle(native=True):
    """The checker, used here only as the timed CPU baseline. Rebuilt with -march=native on this box."""
    from oracle_lib import Oracle, build_oracle
    path = None
    if native:
        try:
            path = build_oracle(out="_native", march="-march=native")
        except Exception:
            path = None
    return Oracle(path)


def cpu_pairs_per_s(orc, pts, desc, npairs, hyps, thr, seed0, threads=0):
    used = C.c_int()
    sub_p = np.ascontiguousarray(pts[:npairs + 1])
    sub_d = np.ascontiguousarray(desc[:npairs + 1])
    t0 = time.perf_counter()
    tot = orc.lib.vbo_pairs_run(sub_p, sub_d, npairs + 1, pts.shape[1], desc.shape[2], 0.7, hyps, thr, seed0, threads,
                                C.byref(used))
    dt = time.perf_counter() - t0
    return npairs / dt, used.value, dt, tot


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port: the matcher is an OpenCV call in the
    reference, so there is no reference source to compile for it) on all host threads, rank 0 only."""
    if rank != 0:
        return
    from vslam_b200 import synth
    orc = cpu_oracle()
    cores = os.cpu_count() or 1
    sample = int(min(args.frames - 1, max(8 * cores, 64)))   # ~1-2 s of host work per step
    pts, desc = synth.sequence(sample + 1, args.kpts, 1000)
    for _ in range(args.warmup):
        cpu_pairs_per_s(orc, pts, desc, min(sample, cores), args.hyps, args.threshold, 1, threads=cores)
    t0 = time.perf_counter()
    used = 1
    for _ in range(args.steps):
        _, used, _, _ = cpu_pairs_per_s(orc, pts, desc, sample, args.hyps, args.threshold, 1, threads=cores)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 XOR+popcount Hamming + f32/f64 residual (host cores)", "data": "synthetic",
            "config": workload_config(args, sample, device=False),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                             "sample": f"{sample} pairs per step x {args.steps} steps, OpenMP over pairs"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, pairs_per_rank, device=True):
    mb = (pairs_per_rank + 1) * args.kpts * 40 / 1e6
    return {"workload": "BASELINE configs[1]: frame pairs of 5000 keypoints, 256-bit binary descriptors, "
                        "1024 RANSAC hypotheses (synthetic forward-motion sequence, SURVEY 8d C2/C4)",
            "kpts": args.kpts, "descriptor_bits": 256, "hypotheses": args.hyps, "threshold": args.threshold, "ratio": 0.7,
            "pairs_per_step_per_gpu": pairs_per_rank,
            "l2": ("per-rank inputs %.0f MB %s 126 MB L2; no explicit flush" % (mb, ">" if mb > 126 else "<")) if device
                  else "host run: bounded sample of the same sequence, %.0f MB of inputs per step" % mb}


def copy_ceiling(torch, dev, h2d_bytes, d2h_bytes, steps, barrier, max_over_ranks):
    """Seconds per step (max over ranks) and GB/s per GPU of plain cudaMemcpyAsync traffic: h2d_bytes up and d2h_bytes down
    per step from / to pinned memory on two streams at once, every rank simultaneously, no kernels."""
    src_h = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    dst_d = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    src_d = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
    dst_h = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(n):
        for _ in range(n):
            with torch.cuda.stream(s_up):
                dst_d.copy_(src_h, non_blocking=True)
            with torch.cuda.stream(s_dn):
                dst_h.copy_(src_d, non_blocking=True)
        s_up.synchronize()
        s_dn.synchronize()
    run(2)
    barrier()
    t0 = time.perf_counter()
    run(steps)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0) / steps
    return dt, (h2d_bytes + d2h_bytes) / dt / 1e9


def reflected_frames(first, count, period):
    """Frame indices into a base sequence of period + 1 frames for frames [first, first + count) of a long sequence that
    walks the base forwards, then backwards, then forwards ... (every consecutive pair is a genuine consecutive pair of
    the base sequence, traversed in one direction or the other)."""
    j = np.arange(first, first + count) % (2 * period)
    return np.where(j <= period, j, 2 * period - j)


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from vslam_b200 import synth
    from vslam_b200.lib import PAIR_RESULT_DTYPE, Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if world > 1:
        dist.init_process_group("gloo", init_method="env://")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = Context(local_rank)
    # a real (non-default) stream: the ABI treats a NULL stream as "use the context's own stream", and
    # torch.cuda.Event only sees the stream it is recorded on
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    nframes, k, nbytes = args.frames, args.kpts, 32
    P = nframes - 1
    pts, desc = synth.sequence(nframes, k, 1000 + rank)           # every rank owns its own shard (weak scaling)
    prm = ctx.params(0.7, 8, args.hyps, args.threshold, 1)

    # device-resident inputs / outputs (torch = device memory plumbing only)
    pts_d = torch.from_numpy(pts).to(dev)
    desc_d = torch.from_numpy(desc).to(dev)
    res_d = torch.zeros(P * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    out_d = torch.zeros((P, k, 2), dtype=torch.int32, device=dev)
    # pinned host buffers for the e2e leg
    pts_h = torch.from_numpy(pts).pin_memory()
    desc_h = torch.from_numpy(desc).pin_memory()
    res_h = torch.zeros(P * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    out_h = torch.zeros((P, k, 2), dtype=torch.int32).pin_memory()

    def step_device():
        ctx._chk(ctx.L.vb_pairs_run_d(ctx.h, pts_d.data_ptr(), desc_d.data_ptr(), nframes, k, nbytes, C.byref(prm),
                                      res_d.data_ptr(), out_d.data_ptr()))

    # device-resident steps as a streaming caller issues them: two submissions in flight (vb_pairs_submit_d), so that the
    # counting and the small kernels of one step run beside the matcher of the next; second set of output buffers
    res_d2 = torch.zeros_like(res_d)
    out_d2 = torch.zeros_like(out_d)
    dev_outs = [(res_d, out_d), (res_d2, out_d2)]

    def steps_device_stream(n):
        def sub(i):
            t = C.c_int(-1)
            r_, o_ = dev_outs[i & 1]
            ctx._chk(ctx.L.vb_pairs_submit_d(ctx.h, pts_d.data_ptr(), desc_d.data_ptr(), nframes, k, nbytes, C.byref(prm),
                                             r_.data_ptr(), o_.data_ptr(), C.byref(t)))
            return t.value
        tk = sub(0)
        for i in range(1, n):
            tn = sub(i)
            ctx._chk(ctx.L.vb_pairs_wait(ctx.h, tk, None))
            tk = tn
        ctx._chk(ctx.L.vb_pairs_wait(ctx.h, tk, None))

    def step_e2e():
        ctx._chk(ctx.L.vb_pairs_run(ctx.h, pts_h.data_ptr(), desc_h.data_ptr(), nframes, k, nbytes, C.byref(prm),
                                    res_h.data_ptr(), out_h.data_ptr()))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------------------
    # (a) one step after the other on one stream (vb_pairs_run_d) — round 1's measurement, kept as value_one_stream
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_one_stream = max_over_ranks(e0.elapsed_time(e1))
    # (b) the same steps with two submissions in flight (vb_pairs_submit_d / vb_pairs_wait): `value`. vb_pairs_wait returns
    # when the ticket's kernels have finished, so the closing event is recorded after all of the timed work.
    steps_device_stream(max(args.warmup, 4))
    barrier()
    ctx.ransac_prune_stats(reset=True)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    steps_device_stream(args.steps)
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - launches0
    value = world * P * args.steps / (ms_total * 1e-3)
    evals_done, evals_full = ctx.ransac_prune_stats()

    # sanity: the timed path produced real results
    res = res_d.cpu().numpy().view(PAIR_RESULT_DTYPE)
    ok_pairs = int((res["status"] == 0).sum())
    mean_matches = float(res["n_matches"].mean())
    sum_tent = int(res["n_tentative"].sum())

    # ---- per-kernel device times (one extra profiled step, CUDA events inside the library) -------
    ctx.profile(True)
    step_device()
    torch.cuda.synchronize(dev)
    kt = {name: ctx.profile_ms(name) for name in ("expand", "hamming", "knnfix", "finish", "sample", "solve", "score", "select")}
    ctx.profile(False)
    # the two heavy stages over several profiled passes of the timed workload (events inside the library, launching stream)
    score_ms, ham_ms = [], []
    for _ in range(3):
        ctx.profile(True)
        step_device()
        torch.cuda.synchronize(dev)
        score_ms.append(ctx.profile_ms("score"))
        ham_ms.append(ctx.profile_ms("hamming"))
        ctx.profile(False)
    score_ms_avg, ham_ms_avg = float(np.mean(score_ms)), float(np.mean(ham_ms))
    peak, bf16_peak, peak_src = measured_peaks()
    # with the profiling brackets on, vb_pairs_run_d works in batches of 1 024 pairs and the brackets keep the last batch
    P_prof = P - 1024 * ((P - 1) // 1024)
    tent_prof = int(res["n_tentative"][P - P_prof:].sum())
    # dominant kernel of the step: the tensor-core matcher
    probe = tensor_probe(ctx)
    roofline = hamming_roofline(P_prof, k, ham_ms_avg, bf16_peak, peak_src, probe)
    if roofline is not None:
        roofline["traffic"] = ncu_traffic("k_knn2_tc4", P_prof, k, args.hyps)
        roofline["pairs_per_launch"] = P_prof
    # second stage: RANSAC inlier counting. SURVEY 8d's unit is 16 B per (hypothesis, match) evaluation; the tiles live in
    # shared memory / L2 and the stage is bound by the FP32 pipe, and the bounded counting performs only part of the
    # evaluations a full pass would (the rest provably cannot change the winner), so both figures are given.
    evals = float(args.hyps) * float(tent_prof)
    frac_done = (evals_done / evals_full) if evals_full else None
    counting = {"kernels": "k_bq_init + k_count_queue (persistent, work queue; packed fp32 residual with exact fallback)",
                "ms_per_launch": score_ms_avg, "pairs_per_launch": P_prof, "evaluations_full": evals,
                "fraction_evaluated": frac_done,
                "evaluations_per_s": (evals * frac_done / (score_ms_avg * 1e-3)) if frac_done else None,
                "hypotheses_decided_per_s": args.hyps * P_prof / (score_ms_avg * 1e-3),
                "logical_GBps_full_pass_equivalent": (16.0 * evals) / (score_ms_avg * 1e-3) / 1e9,
                "hbm_peak_GBps": peak,
                "dram_traffic_k_count_queue": ncu_traffic("k_count", P_prof, k, args.hyps),
                "note": "a hypothesis is abandoned only when its count so far plus every match it has not seen is below a "
                        "count another hypothesis is known to reach: winner, count, score and mask are bit-identical to "
                        "counting everything (tests/test_gpu_bounded_count.py); option ransac_prune = 0 runs the full count "
                        "(k_count2, 4.05 ms on this workload)"}
    step_ms = ms_total / args.steps
    # kernel_ms covers the last batch of P_prof pairs; its share of the step is scaled to the step's P pairs
    shares = {n: (v * (P / P_prof) / step_ms if v and v > 0 else None) for n, v in kt.items()}

    # ---- end to end through the host-pointer C ABI ---------------------------------------------------------------
    # Streaming submission: three tickets in flight — one uploading while two compute on the context's two streams — so the
    # upload of step i+2 and the download of step i-1 overlap the kernels of steps i and i+1. Every step uploads all of its frames from pinned host memory and downloads every pair's result + matches.
    e_steps = max(20, args.steps)
    outs = [(torch.zeros(P * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
             torch.zeros(P, dtype=torch.int32).pin_memory(),
             torch.zeros((P * k, 2), dtype=torch.int16).pin_memory()) for _ in range(E2E_DEPTH)]

    def submit(i, pp, dd, oo):
        t = C.c_int(-1)
        r_, o_, m_ = oo[i % len(oo)]
        ctx._chk(ctx.L.vb_pairs_submit(ctx.h, pp, dd, nframes, k, nbytes, C.byref(prm), r_.data_ptr(), o_.data_ptr(),
                                       m_.data_ptr(), P * k, C.byref(t)))
        return t.value

    def stream_steps(n, pp, dd, oo):
        """n pipelined steps; returns the match count of the last one."""
        tot = C.c_uint64(0)
        inflight = []
        for i in range(n):
            inflight.append(submit(i, pp, dd, oo))
            if len(inflight) == len(oo):
                ctx._chk(ctx.L.vb_pairs_wait(ctx.h, inflight.pop(0), C.byref(tot)))
        while inflight:
            ctx._chk(ctx.L.vb_pairs_wait(ctx.h, inflight.pop(0), C.byref(tot)))
        return int(tot.value)

    stream_steps(max(3, args.warmup), pts_h.data_ptr(), desc_h.data_ptr(), outs)
    barrier()
    t0 = time.perf_counter()
    total_matches = stream_steps(e_steps, pts_h.data_ptr(), desc_h.data_ptr(), outs)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()   # sampled every 20 ms across both timed regions (device-resident and end-to-end)
    last = outs[(e_steps - 1) % len(outs)]
    res_e = last[0].numpy().view(PAIR_RESULT_DTYPE)
    off_e = last[1].numpy().view(np.uint32)
    m16_e = last[2].numpy().view(np.uint16)
    assert np.array_equal(res_e["n_matches"], res["n_matches"]), "e2e and device-resident paths disagree"
    h2d_bytes = int(pts_h.numel() * 4 + desc_h.numel())
    d2h_bytes = int(P * PAIR_RESULT_DTYPE.itemsize + P * 4 + total_matches * 4)
    e2e = {"value": world * P * e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": d2h_bytes, "steps": e_steps,
           "api": "vb_pairs_submit / vb_pairs_wait, %d submissions in flight, pinned host buffers, compact uint16 matches" % E2E_DEPTH,
           "frac_of_device_resident": (world * P * e_steps / e2e_s) / value}
    # the ceiling of this box for the same traffic: plain cudaMemcpyAsync H2D + D2H of the same byte counts, both
    # directions at once, every rank at the same time, nothing computing
    ceil_s, ceil_gbs = copy_ceiling(torch, dev, h2d_bytes, d2h_bytes, 10, barrier, max_over_ranks)
    e2e["copy_ceiling_pairs_per_s"] = world * P / ceil_s
    e2e["copy_ceiling_GBps_per_gpu"] = ceil_gbs
    e2e["frac_of_copy_ceiling"] = e2e["value"] / e2e["copy_ceiling_pairs_per_s"]
    e2e["bound"] = "compute (device-resident step time)" if e2e["copy_ceiling_pairs_per_s"] > value else "host copies"

    # the blocking call (vb_pairs_run: int32 [pairs][k][2] match slab) and the streaming calls on PAGEABLE memory
    def step_blocking():
        ctx._chk(ctx.L.vb_pairs_run(ctx.h, pts_h.data_ptr(), desc_h.data_ptr(), nframes, k, nbytes, C.byref(prm),
                                    res_h.data_ptr(), out_h.data_ptr()))
    step_blocking()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        step_blocking()
    barrier()
    blocking_s = max_over_ranks(time.perf_counter() - t0)
    out_e = out_h.numpy()
    pouts = [(np.zeros(P, PAIR_RESULT_DTYPE), np.zeros(P, np.uint32), np.zeros((P * k, 2), np.uint16)) for _ in range(2)]

    class _NP:   # pageable numpy buffers behind the same .data_ptr() interface
        def __init__(self, a): self.a = a
        def data_ptr(self): return self.a.ctypes.data
    pouts_w = [tuple(_NP(a) for a in o) for o in pouts]
    stream_steps(2, pts.ctypes.data, desc.ctypes.data, pouts_w)
    barrier()
    t0 = time.perf_counter()
    stream_steps(5, pts.ctypes.data, desc.ctypes.data, pouts_w)
    barrier()
    pageable_s = max_over_ranks(time.perf_counter() - t0)
    assert np.array_equal(pouts[0][0]["n_matches"], res["n_matches"])
    e2e_other = {"e2e_blocking_call": {"value": world * P * 5 / blocking_s, "unit": UNIT, "api": "vb_pairs_run (blocking, int32 slab)",
                                       "d2h_bytes_per_step": int(res_h.numel() + out_h.numel() * 4), "steps": 5},
                 "e2e_pageable": {"value": world * P * 5 / pageable_s, "unit": UNIT, "steps": 5,
                                  "api": "vb_pairs_submit / vb_pairs_wait on pageable (malloc) host memory"}}

    # ---- single-pair latency through the reference-shaped call (match_features) ------------------
    fp0 = (pts[0], desc[0], pts[1], desc[1])
    for _ in range(5):
        ctx.match_features(*fp0, prm)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.match_features(*fp0, prm)
    single_ms = (time.perf_counter() - t0) / 20 * 1e3

    # ---- CPU baseline on this box's host cores (rank 0, bounded sample) ---------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        orc = cpu_oracle()
        cores = os.cpu_count() or 1
        v1, _, dt1, _ = cpu_pairs_per_s(orc, pts, desc, 2, args.hyps, args.threshold, 1, threads=1)
        vq, _, _, _ = cpu_pairs_per_s(orc, pts, desc, int(min(P, max(cores, 8))), args.hyps, args.threshold, 1, threads=cores)
        sample = int(min(P, max(cores, 8, vq * 12.0)))   # about 10 s of host work, never more than the step itself
        vN, used, dtN, _ = cpu_pairs_per_s(orc, pts, desc, sample, args.hyps, args.threshold, 1, threads=cores)
        # parity spot check of the timed GPU output (device-resident results, and the matches of the end-to-end call) against
        # the same oracle: first pair and five more spread over the step — counts, winner, every match, F bit for bit
        out_e = out_h.numpy()
        sample_pairs = sorted(set([0, 1, P // 3, P // 2, (2 * P) // 3, P - 1]))
        n_same = 0
        for i in sample_pairs:
            o = orc.match_features(pts[i], desc[i], pts[i + 1], desc[i + 1], 0.7, 8, args.hyps, args.threshold, 1 + i)
            same = (o["n"] == int(res["n_matches"][i]) == int(res_e["n_matches"][i]) and o["best"] == int(res["best_hyp"][i])
                    and o["n_tentative"] == int(res["n_tentative"][i])
                    and np.array_equal(out_e[i, :max(o["n"], 0)], o["matches"])
                    and np.array_equal(m16_e[int(off_e[i]):int(off_e[i]) + max(o["n"], 0)].astype(np.int32), o["matches"])
                    and np.array_equal(np.ascontiguousarray(res["F"][i], np.float32).view(np.uint32).reshape(-1),
                                       np.ascontiguousarray(o["F"], np.float32).view(np.uint32).reshape(-1)))
            n_same += int(bool(same))
            if i == 0:
                parity = bool(same)
        cpu = {"value": vN, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{sample} pairs over {used} OpenMP threads in {dtN:.1f} s (oracle C port, -O3 -march=native)",
               "single_thread_pairs_per_s": v1, "host_cpus": cores, "gpu_matches_oracle_on_pair0": parity,
               "gpu_matches_oracle_on_sampled_pairs": "%d of %d" % (n_same, len(sample_pairs))}

    # ---- BASELINE configs[3]: ONE long sequence cut over the ranks (strong scaling), every rank takes part --------------
    config4 = None if args.quick else config4_strong(args, ctx, torch, dev, stream, pts, desc, prm, rank, world, barrier,
                                                     max_over_ranks)
    # ---- the same weak-scaling work driven by ONE process through vb_multi (rank 0; the other ranks idle at the barrier)
    e2e_multi = None
    if world > 1 and not args.quick:
        barrier()
        if rank == 0:
            e2e_multi = multi_stage(args, torch, pts, desc, prm, world, P, k, nbytes)
        barrier()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": workload_config(args, P), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu,
                "kernel_ms": kt, "kernel_share_of_step": shares,
                "counting": counting,
                "bounded_counting": {"evaluations_done": evals_done, "evaluations_full": evals_full,
                                     "fraction": (evals_done / evals_full) if evals_full else None},
                "single_pair_match_features_ms": single_ms,
                "value_one_stream": {"value": world * P * args.steps / (ms_one_stream * 1e-3), "unit": UNIT,
                                     "ms_per_step": ms_one_stream / args.steps,
                                     "api": "vb_pairs_run_d back to back on one stream (round 1's `value`); `value` itself = "
                                            "vb_pairs_submit_d / vb_pairs_wait, two submissions in flight on two compute streams"},
                "check": {"pairs_ok": ok_pairs, "pairs": P, "mean_final_matches": mean_matches,
                          "mean_tentative": sum_tent / P}}
        line.update(e2e_other)
        line["tensor_probe"] = probe
        if config4 is not None:
            line["config4"] = config4
        if e2e_multi is not None:
            line["e2e_multi"] = e2e_multi
        if not args.quick:
            line["config3"] = config3_stage(ctx, torch, dev, None if args.no_cpu_baseline else cpu_oracle())
            line["score_sweep"] = score_sweep(ctx, torch, dev, peak)
            line["kdtree"] = kdtree_stage(ctx, torch, dev, pts, k)
            line["search_by_projection"] = projection_stage(ctx, k, None if args.no_cpu_baseline else cpu_oracle())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def hamming_roofline(P, k, ms, bf16_peak, peak_src, probe):
    """Dominant kernel of the step: k_knn2_tc4, tensor-pipe work. Algorithmic work = 2 * 256 flop per descriptor pair
    (256-term +-1 dot product). The peak it is divided by is MEASURED on this GPU in this run: back-to-back
    tcgen05.mma kind::mxf4 (the kernel's own instruction shape, M128 N240 K64) from one thread per SM with resident
    operands and no epilogue (vb_probe_tensor_peak). MEASURED_PEAKS.json has no fp4 entry; its bf16 figure and the
    probe's own bf16 reading are printed next to it so the method can be judged."""
    if not ms or ms <= 0:
        return None
    fp4 = "hamming_fp4=0" not in os.environ.get("VB_OPTIONS", "")
    tf = 2.0 * 256.0 * float(P) * k * k / (ms * 1e-3) / 1e12
    peak = probe["mxf4_n240"] if fp4 else probe["f8f6f4_n256"]
    return {"kernel": "k_knn2_tc4<10> (tcgen05 kind::mxf4, e2m1 +-1 on denormal accumulators: ue8m0 2^-71 x 2^-72 scales, packed 16-bit drain)" if fp4 else
                      "k_knn2_tc (tcgen05 kind::f8f6f4, e4m3 +-128)",
            "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
            "peak_source": "measured in this run: vb_probe_tensor_peak, UMMA-only loop of the kernel's instruction shape "
                           "(no MEASURED_PEAKS.json entry exists for fp4)",
            "peak_measured_fp4": probe["mxf4_n256"], "peak_measured_fp4_n240": probe["mxf4_n240"],
            "peak_measured_fp8": probe["f8f6f4_n256"], "peak_measured_bf16_probe": probe["bf16_n256"],
            "peak_bf16_measured_peaks_json": bf16_peak, "peak_bf16_source": peak_src,
            "frac_of_4x_bf16_measured_peaks": tf / (4.0 * bf16_peak), "frac_of_nominal_9000": tf / 9000.0,
            "ms_per_launch": ms, "pair_distances_per_s": float(P) * k * k / (ms * 1e-3),
            "padding": "tiles are 256 queries x 240 train rows: 5120 x 5040 evaluated for 5000 x 5000 (3.2 % of the MMA work)"}


def tensor_probe(ctx):
    """UMMA-only rates (TFLOP/s) on this GPU: the matcher's fp4 shape, the square fp4 / fp8 shapes, and bf16 as a cross-check
    of the method against the cuBLAS figure in MEASURED_PEAKS.json."""
    return {"mxf4_n240": ctx.probe_tensor_peak(0, 240), "mxf4_n256": ctx.probe_tensor_peak(0, 256),
            "f8f6f4_n256": ctx.probe_tensor_peak(1, 256, iters=2048), "bf16_n256": ctx.probe_tensor_peak(2, 256, iters=1024),
            "how": "k_probe_umma: 1 CTA per SM, one thread issues 4 K-steps x iters tcgen05.mma (M128) on shared-memory "
                   "resident operands into one TMEM accumulator, best of 5 launches, CUDA events"}


def config4_strong(args, ctx, torch, dev, stream, pts, desc, prm, rank, world, barrier, max_over_ranks):
    """BASELINE configs[3] as written: ONE synthetic sequence of --frames-total frames (10 001: 10 000 pairs), its pairs cut
    into contiguous ranges over the ranks, each range with a one-frame halo (strong scaling: total work fixed). The long
    sequence walks this run's 1 025-frame base sequence forwards and backwards (every pair is a genuine consecutive pair);
    pair i samples with seed0 + i whatever rank it lands on. Device-resident time by CUDA events, end to end through
    vb_pairs_submit / vb_pairs_wait in 1 024-pair chunks from pinned memory; max over ranks."""
    from vslam_b200.lib import PAIR_RESULT_DTYPE
    from vslam_b200.sequence import shard_pairs
    total_pairs = args.frames_total - 1
    k, nbytes = pts.shape[1], desc.shape[2]
    from vslam_b200 import synth
    base_p, base_d = (pts, desc) if rank == 0 else synth.sequence(pts.shape[0], k, 1000)   # every rank: rank 0's base
    b, e = shard_pairs(total_pairs, world)[rank]
    np_r = e - b
    fidx = reflected_frames(b, np_r + 1, pts.shape[0] - 1)
    sp = torch.from_numpy(np.ascontiguousarray(base_p[fidx])).pin_memory()
    sd = torch.from_numpy(np.ascontiguousarray(base_d[fidx])).pin_memory()
    sp_d, sd_d = sp.to(dev), sd.to(dev)
    res_d = torch.zeros(max(np_r, 1) * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    out_d = torch.zeros((max(np_r, 1), k, 2), dtype=torch.int32, device=dev)
    prm4 = ctx.params(0.7, 8, args.hyps, args.threshold, 1 + b)

    CH = 1024
    chunks = [(c, min(c + CH, np_r)) for c in range(0, np_r, CH)]
    RS = PAIR_RESULT_DTYPE.itemsize

    def dev_pass():   # 1 024-pair submissions, two in flight (vb_pairs_submit_d), results written in place
        tickets = []
        for c0, c1 in chunks:
            t = C.c_int(-1)
            pc = ctx.params(0.7, 8, args.hyps, args.threshold, 1 + b + c0)
            ctx._chk(ctx.L.vb_pairs_submit_d(ctx.h, sp_d.data_ptr() + c0 * k * 8, sd_d.data_ptr() + c0 * k * nbytes, c1 - c0 + 1, k,
                                             nbytes, C.byref(pc), res_d.data_ptr() + c0 * RS, out_d.data_ptr() + c0 * k * 8, C.byref(t)))
            tickets.append(t.value)
            if len(tickets) == 2:
                ctx._chk(ctx.L.vb_pairs_wait(ctx.h, tickets.pop(0), None))
        for t in tickets:
            ctx._chk(ctx.L.vb_pairs_wait(ctx.h, t, None))
    dev_pass()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(stream)
    for _ in range(reps):
        dev_pass()
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    r4 = res_d.cpu().numpy().view(PAIR_RESULT_DTYPE)[:np_r]
    # end to end: the same chunks from pinned host memory
    res_h = torch.zeros(max(np_r, 1) * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    off_h = torch.zeros(max(np_r, 1), dtype=torch.int32).pin_memory()
    m16_h = torch.zeros((max(np_r, 1) * k, 2), dtype=torch.int16).pin_memory()

    def e2e_pass():
        tickets, tot = [], C.c_uint64(0)
        for ci, (c0, c1) in enumerate(chunks):
            t = C.c_int(-1)
            pc = ctx.params(0.7, 8, args.hyps, args.threshold, 1 + b + c0)
            ctx._chk(ctx.L.vb_pairs_submit(ctx.h, sp.data_ptr() + c0 * k * 8, sd.data_ptr() + c0 * k * nbytes, c1 - c0 + 1, k, nbytes,
                                           C.byref(pc), res_h.data_ptr() + c0 * PAIR_RESULT_DTYPE.itemsize,
                                           off_h.data_ptr() + c0 * 4, m16_h.data_ptr() + c0 * k * 4, (c1 - c0) * k, C.byref(t)))
            tickets.append(t.value)
            if len(tickets) == E2E_DEPTH:
                ctx._chk(ctx.L.vb_pairs_wait(ctx.h, tickets.pop(0), C.byref(tot)))
        for t in tickets:
            ctx._chk(ctx.L.vb_pairs_wait(ctx.h, t, C.byref(tot)))
    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        e2e_pass()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / reps
    rh = res_h.numpy().view(PAIR_RESULT_DTYPE)[:np_r]
    same = bool(np.array_equal(rh["n_matches"], r4["n_matches"]) and np.array_equal(rh["best_hyp"], r4["best_hyp"]))
    ok = int((r4["status"] == 0).sum())
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ok, int(same)], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ok, same = int(t[0]), bool(int(t[1]) == world)
    return {"workload": "BASELINE configs[3]: one %d-frame sequence, %d kpts/frame, pairs sharded over %d GPU(s) in contiguous "
                        "ranges with a one-frame halo" % (args.frames_total, k, world),
            "scaling": "strong", "pairs_total": total_pairs, "pairs_this_rank": np_r, "n_gpus": world,
            "device_resident": {"value": total_pairs / (dev_ms * 1e-3), "unit": UNIT, "ms_per_pass": dev_ms},
            "e2e": {"value": total_pairs / e2e_s, "unit": UNIT, "ms_per_pass": e2e_s * 1e3,
                    "h2d_bytes_per_pass_per_rank": int(sp.numel() * 4 + sd.numel()), "chunk_pairs": CH},
            "pairs_with_model": ok, "e2e_equals_device_resident": same, "passes": reps}


def multi_stage(args, torch, pts, desc, prm, world, P, k, nbytes):
    """The weak-scaling workload (world x P pairs) driven by ONE process through vb_multi: one host thread + context + streams
    per GPU, every GPU's results landing in one host array (the host gather). Three submissions in flight."""
    from vslam_b200.lib import PAIR_RESULT_DTYPE, Multi
    tp = world * P
    fidx = reflected_frames(0, tp + 1, pts.shape[0] - 1)
    ph = torch.from_numpy(np.ascontiguousarray(pts[fidx])).pin_memory()
    dh = torch.from_numpy(np.ascontiguousarray(desc[fidx])).pin_memory()
    outs = [(torch.zeros(tp * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
             torch.zeros(tp, dtype=torch.int32).pin_memory(),
             torch.zeros((tp * k, 2), dtype=torch.int16).pin_memory()) for _ in range(E2E_DEPTH)]
    m = Multi(list(range(world)))
    try:
        def submit(i):
            t = C.c_int(-1)
            r_, o_, m_ = outs[i % len(outs)]
            m._chk(m.L.vb_multi_pairs_submit(m.h, ph.data_ptr(), dh.data_ptr(), tp + 1, k, nbytes, C.byref(prm), r_.data_ptr(),
                                             o_.data_ptr(), m_.data_ptr(), tp * k, C.byref(t)))
            return t.value

        def steps(n):
            tot, inflight = C.c_uint64(0), []
            for i in range(n):
                inflight.append(submit(i))
                if len(inflight) == len(outs):
                    m._chk(m.L.vb_multi_pairs_wait(m.h, inflight.pop(0), C.byref(tot)))
            while inflight:
                m._chk(m.L.vb_multi_pairs_wait(m.h, inflight.pop(0), C.byref(tot)))
            return int(tot.value)
        steps(3)
        n = max(10, min(args.steps, 20))
        t0 = time.perf_counter()
        tot = steps(n)
        dt = time.perf_counter() - t0
        r = outs[(n - 1) % len(outs)][0].numpy().view(PAIR_RESULT_DTYPE)
        return {"value": tp * n / dt, "unit": UNIT, "steps": n, "pairs_per_step": tp, "n_gpus": world,
                "api": "vb_multi_pairs_submit / vb_multi_pairs_wait: 1 process, %d host threads, results gathered in one host array" % world,
                "h2d_bytes_per_step": int(ph.numel() * 4 + dh.numel()), "d2h_bytes_per_step": int(tp * 64 + tot * 4),
                "pairs_with_model": int((r["status"] == 0).sum())}
    finally:
        m.close()


def config3_stage(ctx, torch, dev, orc):
    """BASELINE configs[2]: one pair, 20 000 keypoints, 128-d float descriptors, 4 096 hypotheses — match_features in one call
    (tcgen05 bf16 candidate GEMM + exact fp32 re-evaluation, ratio test, RansacFilter, inlier copy-out)."""
    from vslam_b200 import synth
    from vslam_b200.lib import PAIR_RESULT_DTYPE
    n, dim, H = 20000, 128, 4096
    fp = synth.frame_pair_float(n, 5, dim=dim)
    prm = ctx.params(0.7, 8, H, 10.0, 77)
    t = {x: torch.from_numpy(fp[x]).to(dev) for x in ("p1", "d1", "p2", "d2")}
    res_d = torch.zeros(PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    out_d = torch.zeros((n, 2), dtype=torch.int32, device=dev)
    call = lambda: ctx._chk(ctx.L.vb_match_features_l2f_d(ctx.h, t["p1"].data_ptr(), t["d1"].data_ptr(), n, t["p2"].data_ptr(),
                                                          t["d2"].data_ptr(), n, dim, C.byref(prm), out_d.data_ptr(), res_d.data_ptr()))
    for _ in range(3):
        call()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream(dev)
    reps = 10
    e0.record(st)
    for _ in range(reps):
        call()
    e1.record(st)
    torch.cuda.synchronize(dev)
    dev_ms = e0.elapsed_time(e1) / reps
    ctx.profile(True)
    call()
    torch.cuda.synchronize(dev)
    kms = {nm: ctx.profile_ms(nm) for nm in ("l2f", "l2f_gemm", "l2f_rerank", "finish", "sample", "solve", "score", "select")}
    ctx.profile(False)
    g = ctx.match_features_l2f(fp["p1"], fp["d1"], fp["p2"], fp["d2"], prm)
    t0 = time.perf_counter()
    for _ in range(5):
        g = ctx.match_features_l2f(fp["p1"], fp["d1"], fp["p2"], fp["d2"], prm)
    host_ms = (time.perf_counter() - t0) / 5 * 1e3
    r = res_d.cpu().numpy().view(PAIR_RESULT_DTYPE)[0]
    out = {"workload": "BASELINE configs[2]: 1 pair, 20000 kpts, 128-d float descriptors, 4096 hypotheses, threshold 10",
           "device_resident_ms": dev_ms, "pairs_per_s_device_resident": 1e3 / dev_ms, "host_call_ms": host_ms,
           "pairs_per_s_host_call": 1e3 / host_ms, "kernel_ms": {a: b for a, b in kms.items() if b and b > 0},
           "gemm_TFLOPs": (2.0 * n * n * dim / (kms["l2f_gemm"] * 1e-3) / 1e12) if kms.get("l2f_gemm", -1) > 0 else None,
           "n_tentative": int(r["n_tentative"]), "n_matches": int(r["n_matches"]), "best_hyp": int(r["best_hyp"]),
           "host_call_equals_device": bool(g["n"] == int(r["n_matches"]) and g["best"] == int(r["best_hyp"]))}
    if orc is not None:
        # parity on a bounded sample: the oracle's kNN-2 + ratio on 256 queries against the full train set, and the oracle's
        # find_fundamental on the GPU's tentative list
        qs = np.arange(0, n, n // 256)[:256]
        oi, od = orc.knn2_l2f(np.ascontiguousarray(fp["d1"][qs]), fp["d2"])
        gi, gd = ctx.knn2_l2f(fp["d1"], fp["d2"])
        knn_same = bool(np.array_equal(gi[qs], oi) and np.array_equal(gd[qs].view(np.uint32), od.view(np.uint32)))
        tent = ctx.match_l2f(fp["d1"], fp["d2"], 0.7)
        t0 = time.perf_counter()
        o = orc.find_fundamental(fp["p1"], fp["p2"], tent, 8, H, 10.0, 77)
        cpu_ransac_ms = (time.perf_counter() - t0) * 1e3
        out["oracle_check"] = {"knn2_on_256_queries": knn_same,
                               "find_fundamental": bool(o["best"] == g["best"] and o["n_inliers"] == g["n_inliers"]
                                                        and np.array_equal(tent[o["mask"].astype(bool)], g["matches"])
                                                        and np.array_equal(o["F"].view(np.uint32), g["F"].view(np.uint32))),
                               "cpu_oracle_find_fundamental_ms_1core": cpu_ransac_ms}
    return out


def kdtree_stage(ctx, torch, dev, pts, k, nq_big=1 << 20):
    """KD-tree rows of the path (not part of the pairs/s metric).
    (1) One frame: tree build, k nearest queries and k radius-2 queries (the search-by-projection radius, src/vslam.cpp:149)
        on the GPU next to the reference's own src/KDTree.cpp on one host core (oracle/_ref, when it travelled with the repo);
        at k queries the kernels are less than one wave — these are latency figures.
    (2) Throughput: 2^20-query batches of nearest / radius-2 against trees of k and 20 000 points, with SURVEY 8d's
        algorithmic unit (24 * ceil(log2 N) bytes per query) against the measured HBM figure — the tree itself (12 B per
        node) lives in L1/L2, so this is a statement of how far an irregular gather is from a streaming roofline.
    (3) A/B of the lane mapping the north star asked about: lanes per query 1 (product), 8, 32 (option kd_lanes_per_query)."""
    p = np.ascontiguousarray(pts[1])
    rng = np.random.default_rng(0)
    q = np.ascontiguousarray(p + rng.uniform(-2, 2, p.shape), np.float32)
    p_d, q_d = torch.from_numpy(p).to(dev), torch.from_numpy(q).to(dev)
    out_pt = torch.zeros((k, 2), dtype=torch.float32, device=dev)
    out_idx = torch.zeros(k, dtype=torch.int32, device=dev)
    out_d2 = torch.zeros(k, dtype=torch.float32, device=dev)
    offs = torch.zeros(k + 1, dtype=torch.int32, device=dev)
    hits = torch.zeros(16 * k, dtype=torch.int32, device=dev)
    L = ctx.L
    res = {}
    tree = C.c_void_p()
    tot = C.c_uint64()
    ms = {"kd_build": [], "kd_nearest": [], "kd_radius": []}
    for it in range(8):
        ctx.profile(it >= 3)
        if tree:
            L.vb_kdtree_free(tree)
        ctx._chk(L.vb_kdtree_build_d(ctx.h, p_d.data_ptr(), k, C.byref(tree)))
        ctx._chk(L.vb_kdtree_nearest_d(tree, q_d.data_ptr(), k, float("inf"), out_pt.data_ptr(), out_idx.data_ptr(), out_d2.data_ptr()))
        ctx._chk(L.vb_kdtree_radius_d(tree, q_d.data_ptr(), k, 2.0, offs.data_ptr(), hits.data_ptr(), 16 * k, C.byref(tot)))
        torch.cuda.synchronize(dev)
        if it >= 3:
            for name in ms:
                ms[name].append(ctx.profile_ms(name))
    ctx.profile(False)
    L.vb_kdtree_free(tree)
    # batched build: one tree per frame of a 256-frame block in one launch (one CTA per tree)
    nb = min(256, len(pts))
    blk_d = torch.from_numpy(np.ascontiguousarray(pts[:nb])).to(dev)
    handles = (C.c_void_p * nb)()
    bms = []
    for it in range(5):
        ctx.profile(it >= 2)
        ctx._chk(L.vb_kdtree_build_batch_d(ctx.h, blk_d.data_ptr(), nb, k, handles))
        torch.cuda.synchronize(dev)
        if it >= 2:
            bms.append(ctx.profile_ms("kd_build"))
        L.vb_kdtree_free_batch(handles, nb)
    ctx.profile(False)
    hbm = measured_peaks()[0]
    res["gpu_batched_build"] = {"trees": nb, "ms": float(np.mean(bms)), "trees_per_s": nb / (float(np.mean(bms)) * 1e-3),
                                "us_per_tree": float(np.mean(bms)) * 1e3 / nb,
                                "algorithmic_GBps": nb * 8.0 * k * np.ceil(np.log2(k)) / (float(np.mean(bms)) * 1e-3) / 1e9,
                                "frac_of_hbm_peak": nb * 8.0 * k * np.ceil(np.log2(k)) / (float(np.mean(bms)) * 1e-3) / 1e9 / hbm}
    res["gpu_ms"] = {n: float(np.mean(v)) for n, v in ms.items()}
    res["gpu_queries_per_s"] = {"nearest": k / (res["gpu_ms"]["kd_nearest"] * 1e-3), "radius2": k / (res["gpu_ms"]["kd_radius"] * 1e-3)}
    res["radius2_hits"] = int(tot.value)

    # ---- throughput batches and the lane-mapping A/B ------------------------------------------------------------------
    def timed(name, call, reps=5):
        call(); call()
        torch.cuda.synchronize(dev)
        v = []
        for _ in range(reps):
            ctx.profile(True)
            call()
            torch.cuda.synchronize(dev)
            v.append(ctx.profile_ms(name))
            ctx.profile(False)
        return float(np.mean(v))

    big = []
    for n_tree in sorted({k, 20000}):
        tp = np.stack([rng.uniform(0, 1280, n_tree), rng.uniform(0, 720, n_tree)], 1).astype(np.float32)
        qq = np.ascontiguousarray(tp[rng.integers(0, n_tree, nq_big)] + rng.uniform(-3, 3, (nq_big, 2)), np.float32)
        tp_d, qq_d = torch.from_numpy(tp).to(dev), torch.from_numpy(qq).to(dev)
        o_pt = torch.zeros((nq_big, 2), dtype=torch.float32, device=dev)
        o_idx = torch.zeros(nq_big, dtype=torch.int32, device=dev)
        o_d2 = torch.zeros(nq_big, dtype=torch.float32, device=dev)
        o_off = torch.zeros(nq_big + 1, dtype=torch.int32, device=dev)
        o_hit = torch.zeros(8 * nq_big, dtype=torch.int32, device=dev)
        tr = C.c_void_p()
        build_ms = timed("kd_build", lambda: (L.vb_kdtree_free(tr) if tr else None,
                                              ctx._chk(L.vb_kdtree_build_d(ctx.h, tp_d.data_ptr(), n_tree, C.byref(tr)))), reps=3)
        near = lambda: ctx._chk(L.vb_kdtree_nearest_d(tr, qq_d.data_ptr(), nq_big, float("inf"), o_pt.data_ptr(), o_idx.data_ptr(), o_d2.data_ptr()))
        rad = lambda: ctx._chk(L.vb_kdtree_radius_d(tr, qq_d.data_ptr(), nq_big, 2.0, o_off.data_ptr(), o_hit.data_ptr(), 8 * nq_big, C.byref(tot)))
        unit = 24.0 * np.ceil(np.log2(n_tree))
        row = {"tree_points": n_tree, "queries": nq_big, "build_ms_single_tree": build_ms,
               "build_algorithmic_GBps": 8.0 * n_tree * np.ceil(np.log2(n_tree)) / (build_ms * 1e-3) / 1e9,
               "algorithmic_bytes_per_query": unit}
        ab = {}
        for lpq in (1, 8, 32):
            ctx.set_option("kd_lanes_per_query", lpq)
            t_ms = timed("kd_nearest", near, reps=3 if lpq > 1 else 5)
            ab[str(lpq)] = {"ms": t_ms, "queries_per_s": nq_big / (t_ms * 1e-3)}
        ctx.reset_options()
        t_near, t_rad = ab["1"]["ms"], timed("kd_radius", rad)
        row["nearest"] = {"ms": t_near, "queries_per_s": nq_big / (t_near * 1e-3), "algorithmic_GBps": unit * nq_big / (t_near * 1e-3) / 1e9,
                          "frac_of_hbm_peak": unit * nq_big / (t_near * 1e-3) / 1e9 / hbm}
        row["radius2"] = {"ms": t_rad, "queries_per_s": nq_big / (t_rad * 1e-3), "hits": int(tot.value),
                          "algorithmic_GBps": unit * nq_big / (t_rad * 1e-3) / 1e9, "frac_of_hbm_peak": unit * nq_big / (t_rad * 1e-3) / 1e9 / hbm,
                          "note": "count pass + scan + fill pass (CSR output)"}
        row["lanes_per_query_ab"] = ab
        big.append(row)
        L.vb_kdtree_free(tr)
    res["throughput"] = big
    res["roofline_note"] = ("bound = latency / L1-L2 gather of a dependent chain (one node per step), not HBM: the tree (12 B per node) "
                            "stays on chip; algorithmic_GBps uses SURVEY 8d's 24 * ceil(log2 N) B per query against hbm_peak %.0f GB/s" % hbm)
    try:
        from oracle_lib import Ref
        if Ref.available():
            R = Ref()
            t = lambda which, reps: float(R.lib.vbref_kdtree_time_ms(which, p, k, q, k, 2.0, reps))
            res["reference_cpu_ms_1core"] = {"build_value_tree": t(0, 20), "nearest_x_k": t(1, 10), "radius2_x_k": t(2, 10),
                                             "build_frame_kdtree_as_written": t(3, 1), "radius2_frame_kdtree_x_k": t(4, 10)}
            res["reference_note"] = ("reference src/KDTree.cpp compiled unmodified (oracle/_ref); frame_kdtree build is slow "
                                     "as written because its comparator copies the point vector (src/KDTree.cpp:128)")
    except Exception as e:  # the checker is optional here
        res["reference_cpu_ms_1core"] = f"unavailable: {e}"
    return res


def projection_stage(ctx, k, orc):
    """SURVEY 8f rank 1 (not part of the pairs/s metric): search by projection of a 20 000-point map into one frame of k
    keypoints (reference src/vslam.cpp:129-161) through the host-pointer C-ABI call, next to the oracle on one host core."""
    from vslam_b200 import synth
    n_map = 20000
    s = synth.projection_scene(n_map, k, 77)
    tree = ctx.kdtree_build(s["pts"])
    run = lambda: ctx.search_by_projection(tree, s["X"], s["c2"], 1280, 720, s["desc"], s["ids"], s["obs_off"], s["obs_desc"])
    for _ in range(3):
        run()
    ctx.profile(True)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        g = run()
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    dev_ms = ctx.profile_ms("sbp")
    ctx.profile(False)
    tree.free()
    res = {"map_points": n_map, "keypoints": k, "claimed": int(g[4]), "in_view": int(g[3].sum()),
           "gpu_ms_host_call": e2e_ms, "gpu_ms_device": dev_ms, "map_points_per_s_host_call": n_map / (e2e_ms * 1e-3)}
    if orc is not None:
        t0 = time.perf_counter()
        o = orc.search_by_projection(s["X"], s["c2"], 1280, 720, s["pts"], s["desc"], s["ids"], s["obs_off"], s["obs_desc"])
        res["cpu_oracle_ms_1core"] = (time.perf_counter() - t0) * 1e3
        res["matches_oracle"] = bool(np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]))
    return res


def score_sweep(ctx, torch, dev, peak):
    """BASELINE configs[4] corners: matches x hypotheses through vb_ransac_score_d (every hypothesis' inlier count AND residual
    sum, the reference's compute_fundamental_residual :105-140), device-resident, CUDA events inside the library: the scoring
    kernel alone and kernel + fold. logical_GBps = SURVEY 8d's 16 B per (hypothesis, match) evaluation / time."""
    from oracle_lib import Oracle
    from vslam_b200 import synth
    orc = Oracle()
    out = []
    rng = np.random.default_rng(0)
    base = synth.correspondences(4000, 3)
    bank = np.zeros((512, 9), np.float32)
    for i in range(len(bank)):
        sel = rng.choice(len(base), 8, replace=False)
        bank[i] = orc.compute_fundamental(base[sel, :2], base[sel, 2:]).reshape(-1)
    corr_all = synth.correspondences(1000000, 11)
    for m, h in ((1000, 256), (16384, 1024), (1000000, 256), (1000000, 1024), (262144, 4096), (1000000, 16384)):
        corr = np.ascontiguousarray(corr_all[:m])
        Fs = bank[rng.integers(0, len(bank), h)] if h > len(bank) else bank[:h]
        Fs = np.ascontiguousarray(Fs)
        cd, fd = torch.from_numpy(corr).to(dev), torch.from_numpy(Fs).to(dev)
        cnt = torch.zeros(h, dtype=torch.int32, device=dev)
        sc = torch.zeros(h, dtype=torch.float32, device=dev)
        call = lambda: ctx._chk(ctx.L.vb_ransac_score_d(ctx.h, cd.data_ptr(), m, fd.data_ptr(), h, 10.0, cnt.data_ptr(), sc.data_ptr()))
        for _ in range(3):
            call()
        torch.cuda.synchronize(dev)
        ms, ms_sel = [], []
        for _ in range(5):
            ctx.profile(True)
            call()
            torch.cuda.synchronize(dev)
            ms.append(ctx.profile_ms("score"))
            ms_sel.append(ctx.profile_ms("select"))
            ctx.profile(False)
        t = float(np.mean(ms)) * 1e-3
        tf = t + float(np.mean(ms_sel)) * 1e-3
        gbs = 16.0 * m * h / t / 1e9
        row = {"matches": m, "hypotheses": h, "score_kernel_ms": t * 1e3, "score_plus_fold_ms": tf * 1e3, "hyp_per_s": h / tf,
               "evals_per_s": m * h / tf, "logical_GBps": gbs, "logical_GBps_with_fold": 16.0 * m * h / tf / 1e9,
               "frac_of_hbm_peak": gbs / peak}
        # oracle check of three hypotheses (bounded: 3 x m residuals on one core)
        p1, p2 = np.ascontiguousarray(corr[:, :2]), np.ascontiguousarray(corr[:, 2:])
        mm = np.stack([np.arange(m), np.arange(m)], 1).astype(np.int32)
        gc, gs_ = cnt.cpu().numpy(), sc.cpu().numpy()
        okk = True
        for hh in (0, h // 2, h - 1):
            _, _, n_o, s_o = orc.residual(p1, p2, mm, Fs[hh].reshape(3, 3), 10.0)
            okk = okk and int(gc[hh]) == n_o and np.float32(s_o).view(np.uint32) == gs_[hh:hh + 1].view(np.uint32)[0]
        row["matches_oracle_on_3_hypotheses"] = bool(okk)
        out.append(row)
    return out


if __name__ == "__main__":
    main()
