"""Pins the oracle to the reference's own code (CPU only, no GPU).

oracle/_ref/libref[_fast].so is /root/reference/{sift.cpp,algorithms.cpp} compiled UNMODIFIED against Vigra
stand-in headers (oracle/ref_shim/, oracle/Makefile target `ref`).  Three layers:
  1. the oracle restatement equals the committed digests of the reference's outputs (always runs, also where
     /root/reference and oracle/_ref do not exist);
  2. the restatement equals the live reference library stage by stage (runs wherever oracle/_ref was built);
  3. the reference's own alg:: functions and private stages, called one at a time on random inputs, equal the
     oracle's unit functions — and the copy-on-write build equals the deep-copy build.
What stays restated on both sides here: the Vigra routines behind the shim — Gaussian taps, reflect line convolution,
nearest-neighbour resize walk, Householder-QR inverse/linearSolve; test_refbin_pin.py pins those against the reference's shipped
executable (real Vigra 1.11 compiled in) — see DESIGN.md §5."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
import ref_cases as rc
from sift_b200.synth import synth_frame

DIGESTS = json.load(open(os.path.join(rc.GOLDEN, "ref_digests.json")))
have_ref = pytest.mark.skipif(not (ol.ref_available(True) or os.path.exists("/root/reference/sift.cpp")),
                              reason="oracle/_ref not built and /root/reference absent")


def run_case(name, L=None, strict=True):
    make, p, throws, _ = rc.CASES[name]
    o = ol.Oracle(p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"], strict=strict, L=L)
    try:
        kp = o.calculate(make())
    except ol.OraclePrecondition:
        return None, None, p
    return o, kp, p


FAST_CASES = [n for n, c in rc.CASES.items() if not c[3]]
SLOW_CASES = [n for n, c in rc.CASES.items() if c[3]]


# ---- 1. restatement vs committed digests of the reference's outputs ----------------------------------------
@pytest.mark.parametrize("name", FAST_CASES + SLOW_CASES)
def test_oracle_matches_reference_digests(name):
    o, kp, p = run_case(name)
    want = DIGESTS[name]
    if "throws" in want:
        assert o is None, "the reference throws here (sift.cpp:184), the strict oracle must as well"
        return
    assert o is not None
    got = rc.stage_digests(o, kp, p)
    bad = [k for k in want if got.get(k) != want[k]]
    assert not bad, f"stages differing from the reference build: {bad[:8]}"


def test_digest_file_covers_every_case():
    assert set(rc.CASES) <= set(DIGESTS)
    assert "reference sources" in DIGESTS["_library"]


# ---- 2. restatement vs the live reference library -------------------------------------------------------------
@have_ref
@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_equals_live_reference(name):
    L = ol.ref_lib(fast=True)
    r, rk, p = run_case(name, L=L)
    o, okp, _ = run_case(name)
    if rc.CASES[name][2]:
        assert r is None and o is None
        return
    for f in rk:
        assert np.array_equal(rk[f], okp[f]), f"keypoint field {f}"
    for oc in range(p["octaves"]):
        for i in range(p["dpe"] + 1):
            (a, sa), (b, sb) = r.gauss(oc, i), o.gauss(oc, i)
            assert sa == sb and np.array_equal(a, b), f"gauss({oc},{i})"
        for i in range(p["dpe"]):
            (a, sa), (b, sb) = r.dog(oc, i), o.dog(oc, i)
            assert sa == sb and np.array_equal(a, b), f"dog({oc},{i})"
    for getter in ("candidates", "survivors"):
        a, b = getattr(r, getter)(), getattr(o, getter)()
        for f in a:
            assert np.array_equal(a[f], b[f]), f"{getter} field {f}"
    assert r.text() == o.text()
    for scale in (0.3, 0.6627, 0.9373, 1.3255, 1.8745, 2.651, 3.749, 7.0):
        assert r.nearest_gaussian(scale) == o.nearest_gaussian(scale)


@have_ref
@pytest.mark.parametrize("name", ["negative", "ragged", "sub_small", "flat", "dpe4_oct3_throws"])
def test_copy_on_write_build_equals_deep_copy_build(name):
    """libref_fast.so (shim-level copy-on-write + memoised blur) against libref.so (every copy deep, as in Vigra)."""
    a, ak, p = run_case(name, L=ol.ref_lib(fast=True))
    b, bk, _ = run_case(name, L=ol.ref_lib(fast=False))
    if a is None or b is None:
        assert a is None and b is None
        return
    assert rc.stage_digests(a, ak, p) == rc.stage_digests(b, bk, p)


@have_ref
def test_literal_reference_cost_is_quadratic_the_fast_build_is_not():
    """Sanity of the two builds' purpose: same results, very different time on a frame with many candidates."""
    img = synth_frame(256, 256, 21)
    slow, fast = ol.Oracle(3, 3, L=ol.ref_lib(False)), ol.Oracle(3, 3, L=ol.ref_lib(True))
    t_slow, n1 = slow.time_calculate(img)
    t_fast, n2 = fast.time_calculate(img)
    assert n1 == n2 and n1 > 0
    assert t_slow > 2 * t_fast


# ---- 3. the reference's functions one at a time ------------------------------------------------------------------
@have_ref
@pytest.mark.parametrize("w,h,sigma", [(37, 23, 1.6), (64, 48, 3.2), (11, 40, 2.2627417), (6, 6, 1.6), (90, 70, 6.4), (33, 20, 1.0)])
def test_ref_convolve_reduce_increase(w, h, sigma):
    L = ol.ref_lib(True)
    img = np.random.default_rng(w * 131 + h).uniform(-20, 255, (h, w)).astype(np.float32)
    assert np.array_equal(ol.convolve(img, sigma, L=L), ol.convolve(img, sigma))
    assert np.array_equal(ol.reduce(img, sigma, L=L), ol.reduce(img, sigma))
    assert np.array_equal(ol.increase(img, sigma, L=L), ol.increase(img, sigma))


@have_ref
def test_ref_preconditions():
    L = ol.ref_lib(True)
    for fn in (ol.convolve, ol.reduce, ol.increase):
        with pytest.raises(ol.OraclePrecondition):
            fn(np.zeros((5, 40), np.float32), 1.6, L=L)  # h = 5 <= r = 5
        with pytest.raises(ol.OraclePrecondition):
            fn(np.zeros((5, 40), np.float32), 1.6)
    ol.convolve(np.zeros((6, 6), np.float32), 1.6, L=L)


@have_ref
def test_ref_dog_and_gradient():
    L = ol.ref_lib(True)
    rng = np.random.default_rng(3)
    a, b = (rng.uniform(-50, 300, (40, 50)).astype(np.float32) for _ in range(2))
    assert np.array_equal(ol.dog(a, b, L=L), ol.dog(a, b))
    m1, o1 = ol.gradient(a, L=L)
    m2, o2 = ol.gradient(a)
    assert np.array_equal(m1, m2) and np.array_equal(o1, o2)
    flat = np.full((12, 9), 7, np.float32)  # atan2f(0, 0)
    assert np.array_equal(ol.gradient(flat, L=L)[1], ol.gradient(flat)[1])


@have_ref
@pytest.mark.parametrize("seed", range(4))
def test_ref_extrema_and_elimination(seed):
    L = ol.ref_lib(True)
    rng = np.random.default_rng(seed)
    w, h = 61 + seed, 47
    if seed == 3:  # quantised values: many ties (the predicate counts them)
        d = [rng.integers(126, 131, (h, w)).astype(np.float32) for _ in range(3)]
    else:
        base = ol.convolve(rng.uniform(0, 255, (h, w)).astype(np.float32), 1.6)
        d = [np.float32(128) + (ol.convolve(base, s) - base) * np.float32(g) for s, g in ((1.6, 1), (2.26, 3), (3.2, 5))]
    xs, ys = ol.extrema(*d, L=L)
    xo, yo = ol.extrema(*d)
    assert np.array_equal(xs, xo) and np.array_equal(ys, yo) and xs.size > 0
    # every interior pixel as a candidate: exercises all reject branches incl. singular and det < 0
    gx, gy = np.meshgrid(np.arange(1, w - 1), np.arange(1, h - 1), indexing="ij")
    gx, gy = gx.ravel().astype(np.uint16), gy.ravel().astype(np.uint16)
    fr, fo = ol.eliminate(*d, gx, gy, L=L), ol.eliminate(*d, gx, gy)
    assert np.array_equal(fr, fo)
    assert 0 < int(fr.sum()) <= fr.size


@have_ref
def test_ref_vertex_parabola_and_peaks():
    L = ol.ref_lib(True)
    assert ol.vertex_parabola(355, 0.0, 5, 1234.5, 15, 0.0, L=L) == ol.vertex_parabola(355, 0.0, 5, 1234.5, 15, 0.0)
    assert abs(ol.vertex_parabola(355, 0.0, 5, 10.0, 15, 0.0, L=L) - 177.4913) < 1e-3
    rng = np.random.default_rng(8)
    for trial in range(200):
        lx, px, rx = (int(v) for v in rng.choice(np.arange(5, 360, 10), 3, replace=False))
        ly, py, ry = (float(np.float32(v)) for v in rng.uniform(0, 5000, 3))
        a, b = ol.vertex_parabola(lx, ly, px, py, rx, ry, L=L), ol.vertex_parabola(lx, ly, px, py, rx, ry)
        assert a == b or (np.isnan(a) and np.isnan(b))
    for trial in range(200):
        h = rng.uniform(0, 100, 36).astype(np.float32)
        if trial % 4 == 0:
            h[rng.integers(0, 36, 30)] = 0
        if trial % 7 == 0:
            h[:] = 0  # all-zero histogram: 0/0 vertex, NaN in a std::set
            h[rng.integers(0, 36)] = trial
        if trial % 5 == 0:
            h[rng.integers(0, 36, 3)] = h.max()  # equal maxima
        a, b = ol.find_peaks(h, L=L), ol.find_peaks(h)
        assert a.size == b.size and np.array_equal(a, b, equal_nan=True), trial


@have_ref
def test_ref_std_sort_over_real_interest_points():
    """std::sort(cmpByFilter) over the reference's own 56-byte sift::InterestPoint objects gives the permutation the
    oracle derives from a light (flag, index) array — the order descriptors depend on."""
    L = ol.ref_lib(True)
    rng = np.random.default_rng(17)
    for n in (0, 1, 2, 15, 16, 17, 33, 100, 1000, 4097, 70000):
        for density in (0.0, 0.02, 0.3, 0.5, 0.9, 1.0):
            flags = (rng.uniform(0, 1, n) >= density).astype(np.uint8)
            assert np.array_equal(ol.sort_order(flags, L=L), ol.sort_order(flags)), (n, density)


@have_ref
def test_ref_normalize_vector():
    L = ol.ref_lib(True)
    rng = np.random.default_rng(4)
    for v in (rng.uniform(0, 9, 8), np.zeros(8), np.array([1, -1, 0, 0, 0, 0, 0, 0.0]), rng.uniform(-3, 3, 8)):
        v = v.astype(np.float32)
        got = v.copy()
        L.oracle_normalize(got, got.size)
        s = np.float32(0)
        for e in v:
            s = np.float32(s + e)
        want = v if s == 0 else (v / s).astype(np.float32)
        assert np.array_equal(got, want, equal_nan=True)
