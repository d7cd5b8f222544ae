"""Oracle vs an all-cv2 replay of RansacFilter::find_fundamental (reference src/RansacFilter.cpp:36-67) on 54 seeded
problems of BASELINE configs 1 and 2 (tests/golden/gen_golden_e2e.py -> tests/golden/find_fundamental_cv2_4_13.npz).

The only step of the path that is oracle-DEFINED rather than pinned is the 8-point solve (cv::SVDecomp is OpenBLAS sgesdd
in this cv2 build). This test measures what that costs at the OUTPUT of find_fundamental — winner
index, inlier mask, inlier count, F — and gates it, so the divergence is a number with a bar instead of a footnote.
Everything else (matcher, sample sets, residual given F, update rule) is bit-pinned elsewhere; the GPU path is bit-equal
to the oracle (tests/test_gpu_parity.py), so these figures are the GPU path's too.

Measured (2026-10, cv2 4.13.0 + OpenBLAS):  winner index equal on 48 of 54 problems; inlier-mask Hamming distance mean
1.4 % of the matches (<= 1.8 % whenever the winner is the same hypothesis, <= 11.4 % when a near-tied other hypothesis wins);
winner inlier COUNT within 0.6 %; F of the same winner within 3e-3 relative Frobenius (median 1.3e-4); per-hypothesis inlier
counts differ by 4.4 of ~3 000 on average (max 71): the un-normalised fp32 8x9 system (//TODO: normalize, :40) is that
ill-conditioned.
"""
import hashlib
import os

import numpy as np
import pytest

from vslam_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "find_fundamental_cv2_4_13.npz")


def rel_frob(a, b):
    a, b = a.reshape(-1).astype(np.float64), b.reshape(-1).astype(np.float64)
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b))


@pytest.fixture(scope="module")
def e2e_rows(oracle):
    g = np.load(GOLD)
    rows = []
    for i, (k, H, s) in enumerate(g["problems"]):
        k, H, s = int(k), int(H), int(s)
        fp = synth.frame_pair(k, s)
        tent = oracle.match_hamming(fp["d1"], fp["d2"], float(g["ratio"]))
        h = hashlib.sha1()
        for a in (fp["p1"], fp["p2"], fp["d1"], fp["d2"], tent):
            h.update(np.ascontiguousarray(a).tobytes())
        assert h.hexdigest() == str(g["digests"][i]), "synthetic inputs drifted: regenerate tests/golden/gen_golden_e2e.py"
        assert len(tent) == int(g[f"m_{i}"])
        o = oracle.find_fundamental(fp["p1"], fp["p2"], tent, 8, H, float(g["thr"]), 7000 + 13 * s + k, want_all=True)
        cvmask = np.unpackbits(g[f"mask_{i}"])[:len(tent)]
        rows.append(dict(k=k, H=H, m=len(tent), same=int(o["best"]) == int(g[f"best_{i}"]),
                         ham=int((cvmask != o["mask"]).sum()) / len(tent),
                         dn=abs(int(o["n_inliers"]) - int(g[f"n_{i}"])) / max(int(g[f"n_{i}"]), 1),
                         fd=rel_frob(o["F"], g[f"F_{i}"]),
                         dcnt=np.abs(o["cnt_all"] - g[f"cnts_{i}"]),
                         # the oracle's count for the hypothesis cv2 chose, against the oracle's own maximum
                         regret=(int(o["cnt_all"].max()) - int(o["cnt_all"][int(g[f"best_{i}"])])) / len(tent)))
    return rows


def test_e2e_problem_count(e2e_rows):
    assert len(e2e_rows) >= 50
    assert {r["k"] for r in e2e_rows} == {2000, 5000}


def test_e2e_winner_agreement(e2e_rows):
    agree = sum(r["same"] for r in e2e_rows)
    print(f"winner index equal on {agree} of {len(e2e_rows)} problems")
    assert agree >= 0.8 * len(e2e_rows)
    # where the winner differs, cv2's choice is a near-tie in the oracle's own counts (a different but equally good model)
    assert max(r["regret"] for r in e2e_rows) <= 0.01


def test_e2e_inlier_mask_distance(e2e_rows):
    ham = np.array([r["ham"] for r in e2e_rows])
    same = np.array([r["same"] for r in e2e_rows])
    print(f"mask Hamming distance / matches: mean {ham.mean():.4f}, max same-winner {ham[same].max():.4f}, max {ham.max():.4f}")
    assert ham.mean() <= 0.025
    assert ham[same].max() <= 0.03
    assert ham.max() <= 0.15
    assert max(r["dn"] for r in e2e_rows) <= 0.01           # the winners explain the same number of matches


def test_e2e_fundamental_distance(e2e_rows):
    fd = np.array([r["fd"] for r in e2e_rows if r["same"]])
    print(f"F (same winner) relative Frobenius: median {np.median(fd):.2e}, max {fd.max():.2e}")
    assert np.median(fd) <= 5e-4 and fd.max() <= 5e-3


def test_e2e_per_hypothesis_counts(e2e_rows):
    mean = np.mean([r["dcnt"].mean() / r["m"] for r in e2e_rows])
    worst = max(r["dcnt"].max() / r["m"] for r in e2e_rows)
    print(f"per-hypothesis inlier count difference / matches: mean {mean:.5f}, max {worst:.4f}")
    assert mean <= 0.003 and worst <= 0.04
