"""vb_pairs_submit / vb_pairs_wait (three submissions in flight, compact download) and vb_multi (one host thread + context per GPU in one
process): same results, bit for bit, as the blocking vb_pairs_run and as the oracle's match_features
(reference src/Frame.cpp:82-105), whatever the number of tickets in flight or the device list."""
import ctypes as C

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


def _same(res_a, matches_a, res_b, matches_b):
    for f in ("status", "n_tentative", "n_matches", "best_hyp", "n_inliers"):
        assert np.array_equal(res_a[f], res_b[f]), f
    assert np.array_equal(_bits(res_a["score"]), _bits(res_b["score"])) and np.array_equal(_bits(res_a["F"]), _bits(res_b["F"]))
    for i, n in enumerate(res_a["n_matches"]):
        assert np.array_equal(matches_a[i], matches_b[i][:n]), i


def test_compact_equals_blocking_and_oracle(ctx, oracle):
    from vslam_b200.lib import unpack_compact
    pts, desc = synth.sequence(7, 1200, 5)
    desc[3] = synth.random_descriptors(np.random.default_rng(2), 1200)        # two pairs without a model (< 8 tentative)
    prm = ctx.params(0.7, 8, 128, 10.0, 40)
    res_b, out_b = ctx.pairs_run(pts, desc, prm)
    res_c, off, m16 = ctx.pairs_run_compact(pts, desc, prm)
    mc = unpack_compact(res_c, off, m16)
    _same(res_c, mc, res_b, out_b)
    assert off[0] == 0 and np.array_equal(np.diff(off.astype(np.int64)), res_c["n_matches"][:-1])   # packed back to back
    assert (res_c["status"] == 3).sum() == 2
    for i in range(6):
        o = oracle.match_features(pts[i], desc[i], pts[i + 1], desc[i + 1], 0.7, 8, 128, 10.0, 40 + i)
        assert res_c["n_matches"][i] == max(o["n"], 0)
        if o["n"] > 0:
            assert np.array_equal(mc[i], o["matches"]) and np.array_equal(_bits(res_c["F"][i]), _bits(o["F"].reshape(-1)))


def test_lone_submission_uploads_in_two_pieces(ctx):
    """A submission with nothing else in flight and >= 512 pairs uploads in two pieces and starts on the first half early;
    with another one in flight it stays whole. Same bits either way (pair i samples with seed0 + i wherever its batch ends)."""
    from vslam_b200.lib import unpack_compact
    pts, desc = synth.sequence(1300, 260, 31)          # 1 299 pairs: pieces of 649 and 650, the second one in two batches
    prm = ctx.params(0.7, 8, 32, 10.0, 5)
    res_b, out_b = ctx.pairs_run(pts, desc, prm)
    res_c, off, m16 = ctx.pairs_run_compact(pts, desc, prm)                     # alone: split
    _same(res_c, unpack_compact(res_c, off, m16), res_b, out_b)
    small = synth.sequence(4, 260, 32)
    from vslam_b200.lib import PAIR_RESULT_DTYPE
    r0, o0, m0 = np.zeros(3, PAIR_RESULT_DTYPE), np.zeros(3, np.uint32), np.zeros((3 * 260, 2), np.uint16)
    r1, o1, m1 = np.zeros(1299, PAIR_RESULT_DTYPE), np.zeros(1299, np.uint32), np.zeros((1299 * 260, 2), np.uint16)
    t0 = ctx.pairs_submit(small[0], small[1], prm, r0, o0, m0)
    t1 = ctx.pairs_submit(pts, desc, prm, r1, o1, m1)                           # not alone: whole
    ctx.pairs_wait(t0); ctx.pairs_wait(t1)
    _same(r1, unpack_compact(r1, o1, m1), res_b, out_b)


def test_three_tickets_in_flight(ctx):
    from vslam_b200.lib import PAIR_RESULT_DTYPE, VbError, pinned_empty, unpack_compact
    seqs = [synth.sequence(5, 900, s) for s in (1, 2, 3, 4, 5)]
    prm = ctx.params(0.7, 8, 64, 10.0, 7)
    want = [ctx.pairs_run(p, d, prm) for p, d in seqs]
    bufs, keep = [], []
    for p, d in seqs:
        # pinned copies of the inputs and pinned outputs: the asynchronous path proper
        pp, h1 = pinned_empty(ctx.L, p.shape, np.float32); pp[:] = p
        dd, h2 = pinned_empty(ctx.L, d.shape, np.uint8); dd[:] = d
        res, h3 = pinned_empty(ctx.L, (4,), PAIR_RESULT_DTYPE)
        off, h4 = pinned_empty(ctx.L, (4,), np.uint32)
        m16, h5 = pinned_empty(ctx.L, (4 * 900, 2), np.uint16)
        bufs.append((pp, dd, res, off, m16)); keep += [h1, h2, h3, h4, h5]
    sub = lambda i: ctx.pairs_submit(bufs[i][0], bufs[i][1], prm, *bufs[i][2:])
    t0, t1, t2 = sub(0), sub(1), sub(2)                                     # one uploading, two computing (context + twin)
    with pytest.raises(VbError) as e:                                       # a fourth one needs a free slot
        sub(3)
    assert e.value.code == 4
    tots = [ctx.pairs_wait(t0)]
    t3 = sub(3)
    tots.append(ctx.pairs_wait(t1))
    t4 = sub(4)
    tots += [ctx.pairs_wait(t) for t in (t2, t3, t4)]
    with pytest.raises(VbError):
        ctx.pairs_wait(t1)                                                  # already completed
    for (pp, dd, res, off, m16), (rb, ob), tot in zip(bufs, want, tots):
        _same(res, unpack_compact(res, off, m16), rb, ob)
        assert tot == int(res["n_matches"].sum())
    for h in keep:
        ctx.L.vb_host_free(h)


def test_compact_capacity_contract(ctx):
    from vslam_b200.lib import PAIR_RESULT_DTYPE, VbError
    pts, desc = synth.sequence(3, 800, 9)
    prm = ctx.params(0.7, 8, 64, 10.0, 1)
    res, off, small = np.zeros(2, PAIR_RESULT_DTYPE), np.zeros(2, np.uint32), np.zeros((10, 2), np.uint16)
    t = ctx.pairs_submit(pts, desc, prm, res, off, small)
    tot = C.c_uint64(0)
    rc = ctx.L.vb_pairs_wait(ctx.h, t, C.byref(tot))
    assert rc == 4 and tot.value == int(res["n_matches"].sum()) > 10        # results are valid, required size reported
    # no match download at all
    t = ctx.pairs_submit(pts, desc, prm, res, None, None)
    assert ctx.pairs_wait(t) == 0 and (res["n_matches"] > 100).all()
    big = np.zeros((3, 70000, 2), np.float32)
    with pytest.raises(VbError):                                            # uint16 indices: k <= 65535
        ctx.pairs_submit(big, np.zeros((3, 70000, 32), np.uint8), prm, res, off, small)


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0], "all"])
def test_multi_equals_single_context(ctx, devices):
    """The device list (here: also several workers on one GPU, and every GPU of the box) never changes a bit of the output."""
    import torch
    from vslam_b200.lib import Multi
    if devices == "all":
        devices = list(range(torch.cuda.device_count()))
    pts, desc = synth.sequence(12, 700, 21)
    prm = ctx.params(0.7, 8, 96, 10.0, 300)
    res_b, out_b = ctx.pairs_run(pts, desc, prm)
    m = Multi(devices)
    try:
        for _ in range(2):                                                  # second run reuses the workers' buffers
            res, off, m16 = m.pairs_run(pts, desc, prm)
            mm = [m16[int(o):int(o) + int(n)].astype(np.int32) for o, n in zip(off, res["n_matches"])]
            _same(res, mm, res_b, out_b)
        # ranges are packed from first_pair * k
        nw, P, k = len(devices), 11, 700
        firsts = sorted({P * w // nw for w in range(nw)})
        for f in firsts:
            assert off[f] == f * k
        # fewer pairs than workers
        res2, off2, m162 = m.pairs_run(pts[:2], desc[:2], prm)
        assert res2["n_matches"][0] == res_b["n_matches"][0]
        assert np.array_equal(m162[int(off2[0]):int(off2[0]) + int(res2["n_matches"][0])].astype(np.int32), out_b[0][:res_b["n_matches"][0]])
    finally:
        m.close()


def test_device_resident_submissions_overlap_and_agree(ctx):
    """vb_pairs_submit_d: three device-resident submissions in flight on two compute streams (the context's and its twin's) ==
    vb_pairs_run_d on one stream, bit for bit; option pairs_overlap = 0 puts both on one stream with the same results."""
    import torch
    from vslam_b200.lib import PAIR_RESULT_DTYPE, VbError
    nframes, k = 40, 1500
    seqs = [synth.sequence(nframes, k, s) for s in (11, 12, 13)]
    prm = ctx.params(0.7, 8, 128, 10.0, 21)
    P = nframes - 1
    dev = [(torch.from_numpy(p).cuda(), torch.from_numpy(d).cuda()) for p, d in seqs]

    def outs():
        return torch.zeros(P * PAIR_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda"), torch.zeros((P, k, 2), dtype=torch.int32, device="cuda")

    want = []
    for pd, dd in dev:
        r, o = outs()
        ctx._chk(ctx.L.vb_pairs_run_d(ctx.h, C.c_void_p(pd.data_ptr()), C.c_void_p(dd.data_ptr()), nframes, k, 32, C.byref(prm),
                                      C.c_void_p(r.data_ptr()), C.c_void_p(o.data_ptr())))
        ctx.synchronize()
        want.append((r.cpu().numpy().view(PAIR_RESULT_DTYPE), o.cpu().numpy()))
    for overlap in (1, 0):
        ctx.set_option("pairs_overlap", overlap)
        got = [outs() for _ in dev]
        tickets = []
        for (pd, dd), (r, o) in zip(dev, got):
            t = C.c_int(-1)
            ctx._chk(ctx.L.vb_pairs_submit_d(ctx.h, C.c_void_p(pd.data_ptr()), C.c_void_p(dd.data_ptr()), nframes, k, 32, C.byref(prm),
                                             C.c_void_p(r.data_ptr()), C.c_void_p(o.data_ptr()), C.byref(t)))
            tickets.append(t.value)
        with pytest.raises(VbError):                                            # three are in flight: no free slot
            t = C.c_int(-1)
            ctx._chk(ctx.L.vb_pairs_submit_d(ctx.h, C.c_void_p(dev[0][0].data_ptr()), C.c_void_p(dev[0][1].data_ptr()), nframes, k, 32,
                                             C.byref(prm), C.c_void_p(got[0][0].data_ptr()), None, C.byref(t)))
        for t in tickets:
            ctx.pairs_wait(t)
        for (r, o), (wr, wo) in zip(got, want):
            rr = r.cpu().numpy().view(PAIR_RESULT_DTYPE)
            _same(rr, [o.cpu().numpy()[i, :n] for i, n in enumerate(rr["n_matches"])], wr, wo)
