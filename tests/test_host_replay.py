"""The host half of the product path — what sift_gpu_run does between its two device stages: the reference's first
cleanup sort with its u16 size, the orientation-stage bounds test, the second cleanup sort and the descriptor-stage
bounds test (sift.cpp:37-55, :65-70, :173-178) — against the pinned oracle, on the CPU (no GPU needed).

`sift_gpu_debug_host_replay` (include/sift_gpu.h) runs exactly the function the pipeline runs per image
(`replay_image`, sift_b200/csrc/sift_gpu.cu) on a candidate list supplied by the caller.  Here that list is the
oracle's own `_eliminateEdgeResponses` output, so the test isolates the ordering logic: the unstable std::sort replay
(order_replay.h), the truncation and the bounds tests.  The oracle is only the checker."""
import numpy as np
import pytest

import oracle_lib as ol
import ref_cases as rc
from sift_b200 import capi

# "negative" leaves keypoints with several orientation peaks: the reference then appends copies behind the vector
# (sift.cpp:194-200), which the pipeline handles in a second replay once the device has delivered the peaks
# (redo_with_extra_orientations, GPU test test_extra_orientation_peaks_*); the single-pass entry point cannot see them.
CASES = [n for n in rc.CASES if n != "negative"]


def oracle_case(name):
    make, p, throws, _ = rc.CASES[name]
    img = make()
    o = ol.Oracle(p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"], strict=False)
    kp = o.calculate(img)
    return img, o, kp, p, throws


def replay(img, o, p, flags=0):
    c = o.candidates()
    keep = np.flatnonzero(c["filtered"] == 0).astype(np.uint32)
    h, w = img.shape
    return capi.host_replay(w, h, c["x"].size, keep, c["x"][keep], c["y"][keep], c["octave"][keep].astype(np.uint8),
                            c["index"][keep].astype(np.uint8), dogs_per_epoch=p["dpe"], octaves=p["octaves"],
                            sigma=p["sigma"], k=p["k"], subpixel=p["subpixel"], flags=flags)


@pytest.mark.parametrize("name", CASES)
def test_host_replay_reproduces_the_oracles_vector_order(name):
    img, o, kp, p, throws = oracle_case(name)
    got, n_surv = replay(img, o, p)
    assert n_surv == o.survivors()["x"].size                      # (uint16_t) size after the first cleanup (sift.cpp:41)
    assert got.size == kp["x"].size
    for f in ("x", "y", "octave", "index", "filtered"):
        assert np.array_equal(got[f], kp[f]), f"keypoint field {f}: the vector order differs from the reference's"
    assert np.array_equal(got["scale"], kp["scale"])              # DoG scale label, bit for bit
    assert np.array_equal(got["desc_len"].astype(np.int32), kp["desc_len"])
    if throws:                                                    # the dead blur of sift.cpp:184 throws in the reference
        with pytest.raises(capi.SiftGpuPrecondition):
            replay(img, o, p, flags=capi.FLAG_STRICT)
    else:
        strict, _ = replay(img, o, p, flags=capi.FLAG_STRICT)
        assert np.array_equal(strict, got)


def test_host_replay_u16_wrap_drops_what_the_reference_drops():
    """More than 65535 unfiltered candidates: the reference keeps (uint16_t)count points of the sorted vector."""
    img, o, kp, p, _ = oracle_case("u16_wrap")
    n_unf = int((o.candidates()["filtered"] == 0).sum())
    assert n_unf > 65535
    got, n_surv = replay(img, o, p)
    assert n_surv == n_unf % 65536 == o.survivors()["x"].size
    assert got.size == kp["x"].size


def test_host_replay_canonical_mode_keeps_the_set():
    img, o, kp, p, _ = oracle_case("ragged")
    got, _ = replay(img, o, p, flags=capi.FLAG_ORDER_CANONICAL)
    ref = sorted(zip(kp["octave"].tolist(), kp["index"].tolist(), kp["y"].tolist(), kp["x"].tolist()))
    mine = sorted(zip(got["octave"].tolist(), got["index"].tolist(), got["y"].tolist(), got["x"].tolist()))
    assert ref == mine


def test_host_replay_rejects_bad_lists():
    with pytest.raises(capi.SiftGpuError):   # canon must ascend
        capi.host_replay(64, 64, 10, [3, 2], [20, 21], [20, 21], [0, 0], [1, 1])
    with pytest.raises(capi.SiftGpuError):   # index outside the scanned layers
        capi.host_replay(64, 64, 10, [1], [20], [20], [0], [0])
    got, ns = capi.host_replay(64, 64, 10, [], [], [], [], [])
    assert got.size == 0 and ns == 0


# ---- the lossless f32 -> u8 packing of the upload path (sift_b200/csrc/pack_host.cpp) -----------------------------
def test_frame_packing_accepts_exactly_the_8_bit_valued_frames():
    from sift_b200.synth import synth_frame

    img = synth_frame(333, 77, 4)
    out = capi.pack_u8(img)
    assert out is not None and np.array_equal(out, img.astype(np.uint8))
    assert np.array_equal(out.astype(np.float32).view(np.uint32), img.view(np.uint32))      # the device widening restores every bit
    for bad in (0.5, -1.0, 256.0, 255.00002, 1e9, float("nan"), float("inf"), -float("inf"), -0.0, 1e-30):
        for pos in ((0, 0), (76, 332), (40, 161), (7, 17), (76, 320)):   # vector body, row tails, first and last pixel
            t = img.copy()
            t[pos] = bad
            assert capi.pack_u8(t) is None, (bad, pos)
    for v in (0.0, 255.0):
        t = img.copy()
        t[3, 5] = v
        assert capi.pack_u8(t) is not None


@pytest.mark.parametrize("w", [1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 211])
def test_frame_packing_ragged_widths_strides_and_pitches(w):
    a = np.random.default_rng(w).integers(0, 256, (13, w)).astype(np.float32)
    o = capi.pack_u8(a, dst_pitch=((w + 31) // 32) * 32)
    assert o is not None and np.array_equal(o, a.astype(np.uint8))
    big = np.full((13, w + 5), 0.25, np.float32)     # a strided view: the padding columns hold non-packable values
    big[:, :w] = a
    o = capi.pack_u8(big[:, :w])
    assert o is not None and np.array_equal(o, a.astype(np.uint8))
    big[6, w - 1] = 7.5
    assert capi.pack_u8(big[:, :w]) is None


# ---- randomized: replay_image against a numpy model of sift.cpp:37-55 built on the real std::sort ------------------
def _model(w, h, octaves, dpe, n_cand, canon, xs, ys, oc, ix, sigma=1.6, k=capi.SQRT2_F32):
    """sift.cpp:37-55 spelled out: std::sort(cmpByFilter) over the whole candidate vector (the library's debug entry runs
    the real std::sort), u16 size, orientation-stage bounds test, second sort + u16 size, descriptor-stage bounds test."""
    flags = np.ones(n_cand, np.uint8)
    flags[canon] = 0
    order = capi.sort_order(flags)                                  # real std::sort permutation of the n_cand elements
    pos_of = np.full(n_cand, -1, np.int64)
    pos_of[canon] = np.arange(canon.size)
    n1 = canon.size % 65536                                         # (uint16_t) count
    l1 = pos_of[order[:n1]]                                         # survivor slots in vector order
    assert (l1 >= 0).all()
    # nearest Gaussian level of every DoG class, as Sift::_findNearestGaussian (sift.cpp:205-218) over the schedule of :381-417
    g = np.zeros((octaves, dpe + 1), np.float32)
    g[0, 0] = sigma
    e = 0
    for o in range(octaves):
        for j in range(1, dpe + 1):
            g[o, j] = np.float32(np.float64(k) ** e * np.float64(np.float32(sigma)))
            e += 1
        if o < octaves - 1:
            g[o + 1, 0] = g[o, dpe - 1]
            e -= 2
    d = g[:, 1:] - g[:, :-1]
    dims = [(w, h)]
    for _ in range(1, octaves):
        dims.append(((dims[-1][0] + 1) // 2, (dims[-1][1] + 1) // 2))

    def target_dims(o, i):
        best, bo = np.float32(100), 0
        for oo in range(octaves):
            for ii in range(dpe + 1):
                cur = np.abs(np.float32(g[oo, ii] - d[o, i]))
                if cur < best:
                    best, bo = cur, oo
        return dims[bo]

    tw = np.array([target_dims(o, i)[0] for o, i in zip(oc[l1], ix[l1])], np.int64)
    th = np.array([target_dims(o, i)[1] for o, i in zip(oc[l1], ix[l1])], np.int64)
    x, y = xs[l1].astype(np.int64), ys[l1].astype(np.int64)
    outside = (x < 8) | (x >= tw - 8) | (y < 8) | (y >= th - 8)      # sift.cpp:173-178
    order2 = capi.sort_order(outside.astype(np.uint8))
    n2 = int((~outside).sum()) % 65536
    kept = order2[:n2]
    rej = (x[kept] < 8) | (x[kept] > tw[kept] - 8) | (y[kept] < 8) | (y[kept] > th[kept] - 8)   # sift.cpp:65-70
    return n1, l1[kept], rej, d


@pytest.mark.parametrize("seed,w,h,octaves,dpe,n_cand,density", [
    (0, 320, 240, 3, 3, 5000, 0.02), (1, 320, 240, 3, 3, 5000, 0.5), (2, 211, 157, 2, 4, 20000, 0.1),
    (3, 1920, 1080, 5, 3, 140000, 0.014), (4, 1120, 1120, 1, 3, 200000, 0.45), (5, 64, 64, 2, 3, 40, 0.3),
    (6, 640, 480, 4, 3, 70000, 0.97), (7, 97, 131, 3, 3, 17, 1.0), (8, 400, 300, 3, 3, 3000, 0.0)])
def test_host_replay_random_lists_against_the_std_sort_model(seed, w, h, octaves, dpe, n_cand, density):
    rng = np.random.default_rng(seed)
    canon = np.flatnonzero(rng.uniform(size=n_cand) < density).astype(np.uint32)
    m = canon.size
    oc = rng.integers(0, octaves, m).astype(np.uint8)
    ix = rng.integers(1, dpe - 1, m).astype(np.uint8) if dpe > 3 else np.ones(m, np.uint8)
    ow = np.array([max(1, (w + (1 << o) - 1) >> o) for o in range(octaves)])
    oh = np.array([max(1, (h + (1 << o) - 1) >> o) for o in range(octaves)])
    xs = (rng.uniform(size=m) * ow[oc]).astype(np.uint16)           # anywhere in the point's octave: the bounds tests must bite
    ys = (rng.uniform(size=m) * oh[oc]).astype(np.uint16)
    got, n_surv = capi.host_replay(w, h, n_cand, canon, xs, ys, oc, ix, dogs_per_epoch=dpe, octaves=octaves)
    n1, slots, rej, d = _model(w, h, octaves, dpe, n_cand, canon, xs, ys, oc, ix)
    assert n_surv == n1 and got.size == slots.size
    assert np.array_equal(got["x"], xs[slots]) and np.array_equal(got["y"], ys[slots])
    assert np.array_equal(got["octave"], oc[slots]) and np.array_equal(got["index"], ix[slots])
    assert np.array_equal(got["filtered"], rej.astype(np.uint8))
    assert np.array_equal(got["desc_len"], np.where(rej, 0, 128).astype(np.uint8))
    assert np.array_equal(got["scale"], d[oc[slots], ix[slots]])
