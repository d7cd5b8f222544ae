#!/usr/bin/env python
"""Run the BASELINE.json configs that are not the bench line (C1, C3, C5 of SURVEY §8d) on one GPU and
write one JSON document (profiles/rNN_configs.json): parity against the oracle where the oracle finishes
in seconds, device times from CUDA events, and the CPU oracle timed beside it.

    python tools/run_configs.py --out gpurun_out/configs.json [--quick]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from oracle_lib import Oracle, build_oracle  # noqa: E402
from vslam_b200 import synth  # noqa: E402
from vslam_b200.lib import Context  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps * 1e3, r


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def config1(ctx, orc):
    """C1: two frames x 2000 kpts with known F; the reference's CPU path (oracle) vs the GPU path, H=100 (as
    src/vslam.cpp:19) and H=1024, seeds 0-9."""
    rows = []
    for H in (100, 1024):
        ok, cpu_ms, gpu_ms, inl = 0, [], [], []
        for seed in range(10):
            fp = synth.frame_pair(2000, seed)
            prm = ctx.params(0.7, 8, H, 10.0, 100 + seed)
            t0 = time.perf_counter()
            o = orc.match_features(fp["p1"], fp["d1"], fp["p2"], fp["d2"], 0.7, 8, H, 10.0, 100 + seed)
            cpu_ms.append((time.perf_counter() - t0) * 1e3)
            ms, g = timed(lambda: ctx.match_features(fp["p1"], fp["d1"], fp["p2"], fp["d2"], prm), reps=3, warm=1)
            gpu_ms.append(ms)
            same = (g["n"] == o["n"] and np.array_equal(g["matches"], o["matches"]) and np.array_equal(bits(g["F"]), bits(o["F"])))
            ok += bool(same)
            inl.append(o["n"])
        rows.append({"hypotheses": H, "seeds": 10, "bit_exact_pairs": ok, "cpu_oracle_ms_per_pair_1thread": float(np.mean(cpu_ms)),
                     "gpu_e2e_ms_per_pair_single_call": float(np.mean(gpu_ms)), "mean_final_matches": float(np.mean(inl))})
    return rows


def config3(ctx, orc, quick):
    """C3: one pair, 20000 kpts, 128-d float descriptors, 4096 hypotheses. The oracle's brute-force float
    matcher takes minutes at this size, so parity is checked on a 1500-query slice of the same data."""
    k = 4000 if quick else 20000
    fp = synth.frame_pair_float(k, 17)
    ctx.profile(True)
    ms_match, tent = timed(lambda: ctx.match_l2f(fp["d1"], fp["d2"], 0.7), reps=2, warm=1)
    l2_kernel_ms = ctx.profile_ms("l2f")
    ms_ransac, g = timed(lambda: ctx.ransac_fundamental(fp["p1"], fp["p2"], tent, 8, 4096, 10.0, 5), reps=3, warm=1)
    kms = {n: ctx.profile_ms(n) for n in ("sample", "solve", "score", "select")}
    ctx.profile(False)
    nq = 1500
    oidx, odist = orc.knn2_l2f(np.ascontiguousarray(fp["d1"][:nq]), fp["d2"])
    gidx, gdist = ctx.knn2_l2f(np.ascontiguousarray(fp["d1"][:nq]), fp["d2"])
    o = orc.find_fundamental(fp["p1"], fp["p2"], tent, 8, 4096, 10.0, 5)
    return {"kpts": k, "dim": 128, "hypotheses": 4096, "tentative": int(len(tent)),
            "match_l2f_ms_e2e": ms_match, "l2f_kernel_ms": l2_kernel_ms, "ransac_ms_e2e": ms_ransac, "ransac_kernel_ms": kms,
            "pair_distances_per_s": k * k / (l2_kernel_ms * 1e-3) if l2_kernel_ms > 0 else None,
            "knn_slice_bit_exact": bool(np.array_equal(gidx, oidx) and np.array_equal(bits(gdist), bits(odist))),
            "ransac_bit_exact": bool(g["best"] == o["best"] and np.array_equal(g["mask"], o["mask"]) and np.array_equal(bits(g["F"]), bits(o["F"]))),
            "n_inliers": int(g["n_inliers"])}


def config5(ctx, orc, quick):
    """C5: scoring sweep, matches x hypotheses, device-resident, CUDA-event kernel times, vs measured HBM peak."""
    import torch
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rng = np.random.default_rng(0)
    base = synth.correspondences(4000, 3)
    bank = np.stack([orc.compute_fundamental(base[s, :2], base[s, 2:]).reshape(-1)
                     for s in (rng.choice(4000, 8, replace=False) for _ in range(512))]).astype(np.float32)
    ms_list = (1000, 4096, 16384, 65536) if quick else (1000, 4096, 16384, 65536, 262144, 1000000)
    rows = []
    for m in ms_list:
        corr = synth.correspondences(m, 11)
        cd = torch.from_numpy(corr).to(dev)
        for h in (256, 1024, 4096, 16384):
            Fs = bank[rng.integers(0, 512, h)]
            fd = torch.from_numpy(np.ascontiguousarray(Fs)).to(dev)
            cnt = torch.zeros(h, dtype=torch.int32, device=dev)
            sc = torch.zeros(h, dtype=torch.float32, device=dev)
            call = lambda: ctx._chk(ctx.L.vb_ransac_score_d(ctx.h, cd.data_ptr(), m, fd.data_ptr(), h, 10.0, cnt.data_ptr(), sc.data_ptr()))
            for _ in range(2):
                call()
            torch.cuda.synchronize(dev)
            tk, tt = [], []
            for _ in range(4):
                ctx.profile(True)
                call()
                ctx.synchronize()
                tk.append(ctx.profile_ms("score")); tt.append(ctx.profile_ms("score") + ctx.profile_ms("select"))
                ctx.profile(False)
            t = float(np.mean(tk)) * 1e-3
            gbs = 16.0 * m * h / t / 1e9
            row = {"matches": m, "hypotheses": h, "score_kernel_ms": t * 1e3, "score_plus_fold_ms": float(np.mean(tt)),
                   "hypotheses_per_s": h / t, "evals_per_s": m * h / t, "logical_GBps": gbs, "frac_of_measured_hbm": gbs / peak}
            if m * h <= 70e6:      # parity spot check where the oracle is quick
                pick = rng.choice(h, 3, replace=False)
                p1, p2 = np.ascontiguousarray(corr[:, :2]), np.ascontiguousarray(corr[:, 2:])
                mm = np.stack([np.arange(m), np.arange(m)], 1).astype(np.int32)
                gc, gs = cnt.cpu().numpy(), sc.cpu().numpy()
                row["bit_exact_vs_oracle"] = all(
                    (lambda r: r[2] == gc[i] and bits(np.array([r[3]]))[0] == bits(gs[i:i + 1])[0])(orc.residual(p1, p2, mm, Fs[i], 10.0))
                    for i in pick)
            rows.append(row)
    return {"peak_hbm_GBps": peak, "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    try:
        orc = Oracle(build_oracle(out="_native", march="-march=native"))
    except Exception:
        orc = Oracle()
    ctx = Context(0)
    doc = {"C1_two_frames_2000": config1(ctx, orc), "C3_float_20000": config3(ctx, orc, args.quick),
           "C5_scoring_sweep": config5(ctx, orc, args.quick)}
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(doc, open(args.out, "w"), indent=1)
    print(json.dumps(doc)[:3000])


if __name__ == "__main__":
    main()
