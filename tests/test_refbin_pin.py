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
The executable is the literal reference (quadratic in the image size): 1080p, the u16 wrap and six octaves stay with oracle/_ref."""
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
    want = dict(DIGESTS[name])
    want.pop("seconds", None)
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
            if k != "seconds":
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
