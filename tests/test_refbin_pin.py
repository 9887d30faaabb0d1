"""Pins the oracle to the reference's SHIPPED EXECUTABLE (CPU only, no GPU).

/root/reference/bin/arch_x64/sift is the binary the reference's author built (GCC 7, -O3) against the real Vigra 1.11: the
Vigra templates on the path — Kernel1D::initGaussian, the reflect line convolution, resizeImageNoInterpolation,
linalg::inverse / linearSolve / qrDecomposition — and libstdc++ 7's std::sort are compiled into it.  It cannot be started
here (Vigra-impex, OpenCV and Boost are missing), but oracle/refbin_run.cpp maps it and calls its own sift::alg:: and
sift::Sift:: functions in place (tests/refbin.py).  That closes what oracle/_ref leaves open: there the reference's sources run
over stand-in headers, so the Vigra routines themselves were restated; here they are the real thing.

Two layers, as in test_ref_pin.py:
  1. the oracle equals the committed digests of the executable's outputs (tests/golden/refbin_digests.json, generator
     tests/golden/make_refbin_golden.py) — runs everywhere, also where /root/reference does not exist;
  2. the oracle equals the executable run live, stage by stage, on the cases it finishes in seconds, and its alg:: functions
     on random images — runs where the executable and the helper exist.
The executable is the literal reference (quadratic in the image size: config 2 takes it 51 minutes): the six-octave cases stay
with oracle/_ref."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
import ref_cases as rc
import refbin
from sift_b200.synth import synth_frame

DIGESTS = json.load(open(os.path.join(rc.GOLDEN, "refbin_digests.json")))
REF_DIGESTS = json.load(open(os.path.join(rc.GOLDEN, "ref_digests.json")))
CASES = [n for n in DIGESTS if not n.startswith("_")]
NOTES = ("seconds", "only")   # bookkeeping entries of a case, not digests
LIVE = ["tiny", "flat", "ragged", "negative", "sub_small", "sigma_k"]
live = pytest.mark.skipif(not refbin.available(), reason="reference executable or oracle/_ref/refbin_run absent")


def oracle_case(name):
    make, p, throws, _ = rc.CASES[name]
    o = ol.Oracle(p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"], strict=True)
    try:
        kp = o.calculate(make())
    except ol.OraclePrecondition:
        return None, None, p
    return o, kp, p


# ---- 1. committed digests of the executable's outputs ----------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_shipped_executables_digests(name):
    want = {k: v for k, v in DIGESTS[name].items() if k not in NOTES}
    o, kp, p = oracle_case(name)
    if want.get("throws"):
        assert o is None, "the executable leaves calculate() with a vigra exception here, the strict oracle must as well"
        return
    assert o is not None
    got = rc.pyramid_and_point_digests(o, kp, p)
    bad = [k for k in want if got.get(k) != want[k]]
    assert not bad, f"stages differing from the shipped executable: {bad[:8]}"


def test_executable_digests_agree_with_the_source_build():
    """Where both exist, the reference's sources over the stand-in headers (oracle/_ref, ref_digests.json) and its shipped
    executable produce the same digests — two independent routes to the same numbers."""
    n = 0
    for name in CASES:
        a, b = DIGESTS[name], REF_DIGESTS[name]
        if a.get("throws"):
            assert "throws" in b
            continue
        for k, v in a.items():
            if k not in NOTES:
                assert b[k] == v, (name, k)
                n += 1
    assert n > 200
    assert len(DIGESTS["_executable"]["sha256"]) == 64


# ---- 2. the executable run live ----------------------------------------------------------------------------------------
@live
@pytest.mark.parametrize("name", LIVE)
def test_oracle_equals_the_live_executable(name):
    make, p, throws, _ = rc.CASES[name]
    s = refbin.run_stages(make(), p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"])
    o, kp, _ = oracle_case(name)
    assert s is not None and o is not None
    for oc in range(p["octaves"]):
        for i in range(p["dpe"] + 1):
            (a, sa), (b, sb) = s.gauss(oc, i), o.gauss(oc, i)
            assert sa == sb and np.array_equal(a, b), f"gauss({oc},{i})"
        for i in range(p["dpe"]):
            (a, sa), (b, sb) = s.dog(oc, i), o.dog(oc, i)
            assert sa == sb and np.array_equal(a, b), f"dog({oc},{i})"
    a, b = s.candidates(), o.candidates()
    for f in a:
        assert np.array_equal(a[f], b[f]), f"candidates field {f}"     # `filtered`: Vigra's own QR inverse + linearSolve
    k = s.keypoints()
    for f in k:
        assert np.array_equal(k[f], kp[f]), f"keypoint field {f}"      # order: libstdc++ 7's std::sort, compiled in


@live
@pytest.mark.parametrize("w,h,sigma,seed", [(97, 61, 1.6, 3), (64, 64, 3.2, 1), (211, 157, 0.8, 9), (33, 200, 2.2627417, 4), (200, 33, 6.4, 5)])
def test_executables_blur_reduce_increase(w, h, sigma, seed):
    """alg::convolveWithGauss / reduceToNextLevel / increaseToNextLevel of the executable (real Kernel1D::initGaussian, reflect
    convolution, resizeImageNoInterpolation) against the oracle's unit functions, also on float images with negative values."""
    img = synth_frame(w, h, seed) - (100.0 if seed % 2 else 0.0) + np.random.default_rng(seed).uniform(-0.5, 0.5, (h, w)).astype(np.float32)
    blur, red, inc = refbin.run_unit(img, sigma)
    assert np.array_equal(blur, ol.convolve(img, sigma))
    assert np.array_equal(red, ol.reduce(img, sigma))
    assert np.array_equal(inc, ol.increase(img, sigma))


@live
@pytest.mark.parametrize("seed", range(6))
def test_executables_elimination_with_vigras_own_linear_algebra(seed):
    """Sift::_eliminateEdgeResponses of the executable — vigra::linalg::inverse and linearSolve (Householder QR with Vigra's
    rank rule) as Vigra compiled them — on EVERY interior pixel of random DoG stacks: smooth ones, quantised ones full of ties
    and singular Hessians, steep ones.  One flag per pixel, equal to the oracle's (and so to the CUDA kernel's, which the GPU
    tests hold to the oracle on the same construction)."""
    rng = np.random.default_rng(seed)
    w, h = 61 + seed, 47
    if seed == 3:    # quantised values: ties, zero Hessians (inverse() fails), rank-deficient systems
        d = [rng.integers(126, 131, (h, w)).astype(np.float32) for _ in range(3)]
    elif seed == 4:  # pure noise around 128: every branch of the reject chain
        d = [(128 + rng.normal(0, 6, (h, w))).astype(np.float32) for _ in range(3)]
    elif seed == 5:  # two identical layers: the scale derivatives vanish
        a = (128 + rng.normal(0, 3, (h, w))).astype(np.float32)
        d = [a, a.copy(), (128 + rng.normal(0, 3, (h, w))).astype(np.float32)]
    else:
        base = ol.convolve(rng.uniform(0, 255, (h, w)).astype(np.float32), 1.6)
        d = [np.float32(128) + (ol.convolve(base, s) - base) * np.float32(g) for s, g in ((1.6, 1), (2.26, 3), (3.2, 5))]
    gx, gy = np.meshgrid(np.arange(1, w - 1), np.arange(1, h - 1), indexing="ij")
    gx, gy = gx.ravel().astype(np.uint16), gy.ravel().astype(np.uint16)
    fb, fo = refbin.run_eliminate(*d, gx, gy), ol.eliminate(*d, gx, gy)
    assert fb.size == fo.size == gx.size
    assert np.array_equal(fb, fo), f"{int((fb != fo).sum())} of {fb.size} flags differ"
    assert 0 < int(fb.sum())


@live
def test_executables_vertex_parabola_and_peaks():
    rng = np.random.default_rng(8)
    rows = [(355, 0.0, 5, 1234.5, 15, 0.0), (355, 0.0, 5, 10.0, 15, 0.0), (5, 0.0, 15, 0.0, 25, 0.0)]
    for _ in range(300):
        lx, px, rx = (int(v) for v in rng.choice(np.arange(5, 360, 10), 3, replace=False))
        ly, py, ry = (float(np.float32(v)) for v in rng.uniform(0, 5000, 3))
        rows.append((lx, ly, px, py, rx, ry))
    got = refbin.run_vertex(rows)
    want = np.array([ol.vertex_parabola(int(r[0]), r[1], int(r[2]), r[3], int(r[4]), r[5]) for r in rows], np.float32)
    assert np.array_equal(got, want, equal_nan=True)
    assert abs(got[1] - 177.4913) < 1e-3        # what every orientation comes out as
    hs = []
    for trial in range(300):
        hh = rng.uniform(0, 100, 36).astype(np.float32)
        if trial % 4 == 0:
            hh[rng.integers(0, 36, 30)] = 0
        if trial % 7 == 0:
            hh[:] = 0                           # all-zero histogram: 0/0 vertex, NaN in a std::set
            hh[rng.integers(0, 36)] = trial
        if trial % 5 == 0:
            hh[rng.integers(0, 36, 3)] = hh.max()   # equal maxima
        hs.append(hh)
    for trial, (a, hh) in enumerate(zip(refbin.run_peaks(hs), hs)):
        b = ol.find_peaks(hh)
        assert a.size == b.size and np.array_equal(a, b, equal_nan=True), trial


@live
@pytest.mark.parametrize("name", ["ragged", "sub_small", "sigma_k"])
def test_product_host_replay_reproduces_the_executables_vector_order(name):
    """The PRODUCT's host half (sift_gpu_debug_host_replay: both cleanup sorts replayed by order_replay.h, the u16 size, the
    bounds tests) fed with the executable's own candidate list comes out in the executable's own keypoint order — the order
    libstdc++ 7's std::sort, compiled into that binary, left the reference's vector in."""
    from sift_b200 import capi

    make, p, _, _ = rc.CASES[name]
    img = make()
    s = refbin.run_stages(img, p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"])
    c, k = s.candidates(), s.keypoints()
    keep = np.flatnonzero(c["filtered"] == 0).astype(np.uint32)
    h, w = img.shape
    got, _ = capi.host_replay(w, h, c["x"].size, keep, c["x"][keep], c["y"][keep], c["octave"][keep].astype(np.uint8),
                              c["index"][keep].astype(np.uint8), dogs_per_epoch=p["dpe"], octaves=p["octaves"], sigma=p["sigma"],
                              k=p["k"], subpixel=p["subpixel"])
    assert got.size == k["x"].size > 0
    for f in ("x", "y", "octave", "index", "scale", "filtered"):
        assert np.array_equal(got[f], k[f]), f


@live
@pytest.mark.parametrize("w,h,sigma", [(5, 20, 1.6), (6, 20, 1.6), (20, 5, 1.6), (20, 6, 1.6), (10, 10, 3.2), (11, 11, 3.2), (2, 2, 0.3),
                                       (1, 1, 0.3), (2, 1, 0.3), (3, 3, 0.5), (4, 4, 1.0), (3, 2, 0.3), (19, 40, 6.4), (20, 40, 6.4)])
def test_executables_preconditions_sit_where_the_oracle_puts_them(w, h, sigma):
    """Vigra's own precondition checks (kernel longer than line; resize source / destination too small), reached through the
    executable's alg:: functions: the oracle — and the product's plan check, which uses the same rule — must throw on exactly
    the same shapes, with Vigra's message."""
    img = np.random.default_rng(w * 100 + h).integers(0, 256, (h, w)).astype(np.float32)
    theirs = refbin.unit_exception(img, sigma)
    mine = None
    for fn in (ol.convolve, ol.reduce, ol.increase):     # the order the helper calls them in
        try:
            fn(img, sigma)
        except ol.OraclePrecondition as e:
            mine = str(e)
            break
    assert (theirs is None) == (mine is None), (theirs, mine)
    if mine is not None:   # the oracle's unit helpers only name the function that threw; Vigra's text comes from the executable
        want = {"kernel longer than line": "kernel longer than line", "reduce": "resizeImageNoInterpolation()",
                "increase": "resizeImageNoInterpolation()"}[mine]
        assert want in theirs, (theirs, mine)
