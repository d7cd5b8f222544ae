"""numpy model of the tensor matcher's packed drain (vslam_b200/csrc/hamming_tc.cu, tc_drain = 6 .. 10) and of the stride-2 fix
pass: the ARITHMETIC of the scheme, checked on the CPU against the definition — BFMatcher(NORM_HAMMING) knnMatch k = 2
(reference src/Frame.cpp:83-85: nearest and second nearest by (distance, train index)) and the ratio test of :91.

What is modelled, step by step as the kernels do it:
  * an accumulator word = fp32(MAGIC) + 64 * dot, dot = 256 - 2 * distance, accumulated in fp32 (numpy float32 adds): the word's
    upper half stays 0x4B40, its low half is 128 * (257 - distance);
  * tcgen05.ld.pack::16b: register i of a 32-column load = low half of column 2i | low half of column 2i + 1 << 16;
  * a group = the 8 even or the 8 odd columns of a 16-column span: packed unsigned 16-bit maxima over the span's 8 registers;
  * key = that maximum + (127 - position), position = 4 * tile + span within the warp's 64-column part; running (best, best
    other) per half-word by min / max / max; columns past n2 contribute a zero half-word; keys below 128 = none;
  * per (row, part) the two smallest of the four group keys (distance << 22 | first column) go to the fix pass, which merges the
    parts, evaluates the best group exactly (8 columns, stride 2), takes the other group's distance as second-distance
    candidate, and evaluates BOTH groups when the odd group of the best group's span ties with it.
The GPU tests (tests/test_gpu_parity.py) check the kernels themselves against the oracle; this file checks that the scheme is
right for every input the model is fed, including the ties it was designed around, without a GPU."""
import numpy as np
import pytest

MAGIC = np.uint32(0x4B404080)
NCOLS, PARTS = 240, ((0, 64), (64, 64), (128, 64), (192, 48))
IDX_BITS = 22
NONE = 0xFFFFFFFF


def accumulate(dist):
    """fp32 accumulator words for a [rows, cols] table of distances: MAGIC + four K-steps of +-64 products."""
    acc = np.full(dist.shape, MAGIC, np.uint32).view(np.float32).copy()
    dot = (256 - 2 * dist.astype(np.int64)).astype(np.int64)
    # four K-steps: split the dot product into four partial sums of the right parity (any split is exact)
    parts = np.stack([dot // 4, dot // 4, dot // 4, dot - 3 * (dot // 4)])
    for k in range(4):
        acc = (acc + (64.0 * parts[k]).astype(np.float32)).astype(np.float32)
    return acc.view(np.uint32)


def group_key(key16, part_c0, parity):
    if key16 < 128:
        return NONE
    pos = 127 - (key16 & 127)
    col = (pos >> 2) * NCOLS + part_c0 + (pos & 3) * 16 + parity
    return ((257 - (key16 >> 7)) << IDX_BITS) | col


def drain_row(dist_row, n2):
    """One query row against n2 train columns: the (k1, k2) group keys of every column part, as k_knn2_tc4<6> writes them."""
    ntiles = (n2 + NCOLS - 1) // NCOLS
    out = []
    for c0, cw in PARTS:
        r0 = [0, 0]      # [even-column half, odd-column half]
        r1 = [0, 0]
        for j in range(ntiles):
            tile0 = j * NCOLS + c0
            if tile0 >= n2:
                continue
            nvalid = n2 - tile0
            cols = np.arange(cw)
            d = np.where(cols < nvalid, dist_row[np.minimum(tile0 + cols, n2 - 1)], 0)
            words = accumulate(d[None, :])[0]
            assert np.all((words >> 16) == 0x4B40)                      # the upper half-word never moves
            low = np.where(cols < nvalid, words & 0xFFFF, 0)            # masked columns: zero half-word
            for span in range(cw // 16):
                posc = 127 - (4 * j + span)
                for parity in (0, 1):
                    key = int(low[16 * span + parity:16 * span + 16:2].max()) + posc
                    t = min(r0[parity], key)
                    r0[parity] = max(r0[parity], key)
                    r1[parity] = max(r1[parity], t)
        ka, kb = group_key(r0[0], c0, 0), group_key(r0[1], c0, 1)
        kc, kd = group_key(r1[0], c0, 0), group_key(r1[1], c0, 1)
        out.append((min(ka, kb), min(max(ka, kb), min(kc, kd))))
    return out


def fix8(dist_row, n2, parts):
    """k_knn2_tc_fix<8, 2>: (best key, second key) with key = distance << 22 | column; the second key's column is only a group's
    first column when it comes from the other group."""
    k1 = k2 = NONE
    for a, b in parts:
        lo, hi = min(k1, a), max(k1, a)
        k2 = min(k2, b, hi)
        k1 = lo
    keys, others = [], []
    mask = (1 << IDX_BITS) - 1
    tie = k2 != NONE and (k2 >> IDX_BITS) == (k1 >> IDX_BITS) and (k2 & mask) == (k1 & mask) + 1
    for sub in range(8):
        key, other = NONE, k2
        col = (k1 & mask) + 2 * sub
        if k1 != NONE and col < n2:
            key = (int(dist_row[col]) << IDX_BITS) | col
        if tie:
            col2 = (k2 & mask) + 2 * sub
            if col2 < n2:
                key2 = (int(dist_row[col2]) << IDX_BITS) | col2
                other = max(key, key2)
                key = min(key, key2)
        keys.append(key)
        others.append(other)
    best = min(keys)
    second = min(others[i] if keys[i] == best else keys[i] for i in range(8))
    return best, second


def brute(dist_row):
    order = np.lexsort((np.arange(len(dist_row)), dist_row))
    return int(order[0]), int(dist_row[order[0]]), int(dist_row[order[1]])


def check_row(dist_row, n2):
    best, second = fix8(dist_row, n2, drain_row(dist_row, n2))
    bi, bd, sd = brute(dist_row[:n2])
    assert (best & ((1 << IDX_BITS) - 1), best >> IDX_BITS, second >> IDX_BITS) == (bi, bd, sd)


@pytest.mark.parametrize("n2", [2, 7, 9, 16, 17, 239, 240, 241, 479, 500, 1000])
def test_packed_drain_random_rows(n2):
    rng = np.random.default_rng(n2)
    for _ in range(12):
        d = rng.integers(90, 170, n2)
        d[rng.integers(0, n2, max(1, n2 // 40))] = rng.integers(0, 60)        # a few close candidates
        check_row(d, n2)


def test_packed_drain_extreme_distances():
    for n2 in (16, 33, 250):
        for val in (0, 1, 255, 256):
            check_row(np.full(n2, val), n2)                                   # all equal: lowest index wins, second = same
        d = np.full(n2, 256)
        d[n2 - 1] = 255
        check_row(d, n2)


@pytest.mark.parametrize("a,b", [(1, 4), (4, 1), (2, 3), (0, 15), (15, 16), (14, 17), (63, 64), (61, 66), (239, 240), (236, 243),
                                 (191, 192), (33, 36), (250, 265)])
def test_packed_drain_ties_between_interleaved_groups(a, b):
    """Equal distances in the even and the odd group of one span (and across spans, parts, tiles): for the best, for the second,
    and three-way."""
    n2 = 600
    rng = np.random.default_rng(a * 1000 + b)
    for case in range(6):
        d = rng.integers(100, 160, n2)
        if case % 3 == 0:        # the pair ties for best
            d[a] = d[b] = 20
        elif case % 3 == 1:      # a unique best elsewhere, the pair ties for second
            d[(a + 97) % n2] = 10
            d[a] = d[b] = 20
        else:                    # three-way tie including a later span
            d[a] = d[b] = d[(b + 40) % n2] = 20
        check_row(d, n2)


def test_packed_drain_position_range():
    """32 tiles x 4 spans = 128 positions: the last span of the last tile still gets a key above the 'none' range, and a key never
    reaches into the other half-word (largest key = 128 * 257 + 127 < 2^16)."""
    n2 = 32 * NCOLS
    d = np.full(n2, 256)
    d[n2 - 1] = 256          # worst distance at the very last position
    check_row(d, n2)
    d[n2 - 1] = 0
    check_row(d, n2)
    assert 128 * 257 + 127 < 65536


def test_denormal_accumulators_carry_the_same_low_half_word():
    """Variants 9 and 10 start an accumulator from the DENORMAL 16512 * 2^-149 (bits 0x00004080) and scale every product to
    +-64 * 2^-149 (ue8m0 factors 2^-71 and 2^-72), so that the constant can be written back with tcgen05.st.unpack::16b, which
    zero-fills the upper half-words. In IEEE fp32 with gradual underflow that sum is exact in any order: the bits are
    16512 + 64 * dot, upper half-word zero, the low half-word the one variant 6 reads. (That the tensor pipe does not flush
    denormals is checked on the GPU: tools/probe/tmem_probe.cu and every parity test of the default matcher.)"""
    rng = np.random.default_rng(4)
    unit = np.float32(2.0 ** -149)
    assert unit > 0 and np.array([unit]).view(np.uint32)[0] == 1      # numpy keeps denormals
    prod = np.float32(2.0 ** -71) * np.float32(2.0 ** -72)           # one +1 * +1 product after both scale factors
    assert np.array([prod]).view(np.uint32)[0] == 64
    for dist in list(range(0, 257, 16)) + [1, 255, 37, 200]:
        bits = np.concatenate([np.ones(256 - dist, np.float32), -np.ones(dist, np.float32)])
        rng.shuffle(bits)
        acc = np.array([0x00004080], np.uint32).view(np.float32)[0]
        for k in range(4):                                             # four K-steps of 64 products, each summed exactly
            step = np.float32(bits[64 * k:64 * k + 64].sum()) * prod
            acc = np.float32(acc + step)
        word = int(np.array([acc]).view(np.uint32)[0])
        ref = int(accumulate(np.array([[dist]]))[0, 0])
        assert word >> 16 == 0 and word == 16512 + 64 * (256 - 2 * dist)
        assert word & 0xFFFF == ref & 0xFFFF                           # the half-word the packed loads deliver
