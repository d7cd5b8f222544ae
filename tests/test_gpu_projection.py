"""Search by projection (reference src/vslam.cpp:129-161 + src/PointMap.cpp:36-46) through the C ABI vs the oracle.
Bars: projected coordinates bit-exact, in-view flags, claims (assign), updated map_point_ids and claim count identical."""
import numpy as np
import pytest

from vslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from vslam_b200.lib import Context
    c = Context(0)
    yield c
    c.close()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _run_both(ctx, oracle, s, W=1280, H=720, radius=2.0, thr=64):
    t = ctx.kdtree_build(s["pts"])
    try:
        g = ctx.search_by_projection(t, s["X"], s["c2"], W, H, s["desc"], s["ids"], s["obs_off"], s["obs_desc"], radius, thr)
    finally:
        t.free()
    o = oracle.search_by_projection(s["X"], s["c2"], W, H, s["pts"], s["desc"], s["ids"], s["obs_off"], s["obs_desc"], radius, thr)
    return g, o


def _check(g, o):
    ga, gids, gxy, ginv, gcnt = g
    oa, oids, oxy, oinv = o
    assert np.array_equal(ginv, oinv)
    # projected coordinates: bit-exact wherever they are numbers (x/0 and 0/0 rows compare as NaN/inf classes)
    fin = np.isfinite(oxy).all(1)
    assert np.array_equal(_bits(gxy[fin]), _bits(oxy[fin]))
    assert np.array_equal(np.isnan(gxy), np.isnan(oxy))
    assert np.array_equal(ga, oa)
    assert np.array_equal(gids, oids)
    assert gcnt == int((oa >= 0).sum())


@pytest.mark.parametrize("n_map,k,seed", [(3000, 5000, 1), (50, 500, 2), (99, 800, 3), (100, 800, 4), (20000, 5000, 5),
                                          (1, 10, 6), (700, 3, 7)])
def test_search_by_projection_matches_sequential_reference(ctx, oracle, n_map, k, seed):
    s = synth.projection_scene(n_map, k, seed)
    g, o = _run_both(ctx, oracle, s)
    _check(g, o)
    if n_map >= 3000:
        assert (o[0] >= 0).sum() > min(n_map, k) // 5   # the scene really exercises claims


def test_search_by_projection_long_displacement_chains(ctx, oracle):
    """2000 map points fighting over 40 keypoints in tight clusters: many rounds of deferred acceptance."""
    s = synth.projection_scene(2000, 40, 11, contested=0.9, claimed=0.0)
    s["pts"][:] = s["pts"][0] + np.random.default_rng(0).uniform(-1.2, 1.2, s["pts"].shape).astype(np.float32)
    rng = np.random.default_rng(1)
    # every map point projects into the cluster and matches every keypoint (same base descriptor, few flips)
    K = np.array([[525, 0, 640], [0, 525, 360], [0, 0, 1.0]])
    s["c2"] = np.concatenate([K, np.zeros((3, 1))], 1).astype(np.float32)
    uv = s["pts"][0].astype(np.float64) + rng.uniform(-0.8, 0.8, (2000, 2))
    z = rng.uniform(3, 9, 2000)
    s["X"] = np.stack([(uv[:, 0] - 640) / 525 * z, (uv[:, 1] - 360) / 525 * z, z, np.ones(2000)], 1).astype(np.float32)
    base = synth.random_descriptors(rng, 1, 32)
    s["desc"] = synth.flip_bits(rng, np.repeat(base, 40, 0), 10)
    s["obs_off"] = np.arange(2001, dtype=np.int32)
    s["obs_desc"] = synth.flip_bits(rng, np.repeat(base, 2000, 0), 10)
    g, o = _run_both(ctx, oracle, s)
    _check(g, o)
    assert (o[0] >= 0).sum() == 40   # every keypoint ends up claimed, by the 40 lowest-indexed contenders that reach it


def test_search_by_projection_descriptor_widths_and_threshold(ctx, oracle):
    for nbytes, thr in ((16, 30), (64, 120)):
        s = synth.projection_scene(1500, 2000, 21, nbytes=nbytes)
        g, o = _run_both(ctx, oracle, s, thr=thr)
        _check(g, o)


def test_search_by_projection_empty_inputs(ctx):
    s = synth.projection_scene(10, 100, 9)
    t = ctx.kdtree_build(s["pts"])
    a, ids, xy, inv, cnt = ctx.search_by_projection(t, s["X"][:0], s["c2"], 1280, 720, s["desc"], s["ids"], np.zeros(1, np.int32),
                                                    s["obs_desc"][:0])
    t.free()
    assert len(a) == 0 and cnt == 0 and np.array_equal(ids, s["ids"])
