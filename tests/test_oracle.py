"""CPU tests of the oracle (oracle/): the known-answer facts SURVEY.md Appendix A/B pins from the
reference's shipped binary, independent cross-checks (scipy mirror convolution, numpy lstsq) and the
committed golden vectors.  The reference has no tests of its own, so these are what keep
the restatement honest.  No GPU needed."""
import glob
import os

import numpy as np
import pytest
import scipy.ndimage as ndi

import oracle_lib as ol
from sift_b200.synth import synth_frame

K = float(np.float32(np.sqrt(2.0)))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- Appendix A.1 / B: Gaussian taps and radii -------------------------------------------------
@pytest.mark.parametrize("sigma,radius", [(1.0, 3), (1.6, 5), (2.2627417, 7), (3.2, 10), (4.5254834, 14), (6.4, 19),
                                          (9.0509668, 27), (12.8, 38), (18.1019336, 54), (0.1, 1)])
def test_gaussian_radius_and_normalisation(sigma, radius):
    taps, r = ol.gaussian_taps(sigma)
    assert r == radius and taps.size == 2 * radius + 1
    assert abs(float(taps.sum(dtype=np.float64)) - 1.0) < 1e-6
    assert np.array_equal(taps, taps[::-1])  # bitwise symmetric
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    ref = np.exp(-0.5 * x * x / (np.float32(sigma).astype(np.float64) ** 2))
    ref /= ref.sum()
    assert np.allclose(taps, ref, rtol=2e-6, atol=1e-9)


def test_sigma_zero_is_identity_tap():
    taps, r = ol.gaussian_taps(0.0)
    assert r == 0 and taps.tolist() == [1.0]


# ---- Appendix A.2: reflect-101 separable convolution ---------------------------------------------
@pytest.mark.parametrize("w,h,sigma", [(37, 23, 1.6), (64, 64, 3.2), (11, 40, 2.2627417), (6, 6, 1.6)])
def test_blur_matches_scipy_mirror(w, h, sigma):
    rng = np.random.default_rng(w * 100 + h)
    img = rng.uniform(0, 255, (h, w)).astype(np.float32)
    taps, r = ol.gaussian_taps(sigma)
    ref = ndi.correlate1d(ndi.correlate1d(img.astype(np.float64), taps.astype(np.float64), axis=1, mode="mirror"),
                          taps.astype(np.float64), axis=0, mode="mirror")
    out = ol.convolve(img, sigma)
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-4)


def test_blur_is_sequential_fp32_mul_add():
    """Bit-level restatement: sum += tap*src in ascending source order, fp32 temp between passes."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (9, 13)).astype(np.float32)
    taps, r = ol.gaussian_taps(1.0)

    def line(v):
        n = v.size
        out = np.zeros(n, np.float32)
        for x in range(n):
            s = np.float32(0)
            for j in range(-r, r + 1):
                i = x + j
                i = -i if i < 0 else i
                i = 2 * (n - 1) - i if i >= n else i
                s = np.float32(s + np.float32(taps[r - j] * v[i]))
            out[x] = s
        return out

    tmp = np.stack([line(row) for row in img])
    ref = np.stack([line(col) for col in tmp.T]).T
    assert np.array_equal(ol.convolve(img, 1.0), ref)


def test_blur_precondition_kernel_longer_than_line():
    img = np.zeros((5, 40), np.float32)  # h = 5 <= r = 5
    with pytest.raises(ol.OraclePrecondition):
        ol.convolve(img, 1.6)
    ol.convolve(np.zeros((6, 6), np.float32), 1.6)  # w = h = r + 1 is allowed


# ---- Appendix A.3: nearest-neighbour resize index walk ------------------------------------------
@pytest.mark.parametrize("n_old,seam", [(1920, 479), (1080, 269), (600, 149), (488, 121), (270, 66)])
def test_even_halving_has_a_seam(n_old, seam):
    m = ol.resize_map(n_old, n_old // 2)
    i = np.arange(n_old // 2)
    expect = np.where(i <= seam, 2 * i, 2 * i + 1)
    assert np.array_equal(m, expect)


@pytest.mark.parametrize("n_old", [135, 375, 75, 61])
def test_odd_halving_is_exact_stride_two(n_old):
    m = ol.resize_map(n_old, (n_old + 1) // 2)
    assert np.array_equal(m, 2 * np.arange((n_old + 1) // 2))


@pytest.mark.parametrize("n", [488, 600, 2160, 3840, 7])
def test_doubling_is_pixel_duplication(n):
    assert np.array_equal(ol.resize_map(n, 2 * n), np.arange(2 * n) // 2)


def test_increase_and_reduce_shapes_and_values():
    img = synth_frame(30, 22, 1)
    up = ol.increase(img, 1.0)
    assert up.shape == (44, 60)
    b = ol.convolve(img, 1.0)
    assert np.array_equal(up, np.repeat(np.repeat(b, 2, 0), 2, 1))
    dn = ol.reduce(img, 1.6)
    assert dn.shape == (11, 15)
    bb = ol.convolve(img, 1.6)
    assert np.array_equal(dn, bb[np.ix_(ol.resize_map(22, 11), ol.resize_map(30, 15))])


# ---- Appendix A.5 / F3: linear algebra ------------------------------------------------------------
def test_inverse3_matches_numpy_and_detects_singular():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.normal(size=(3, 3)).astype(np.float32) + 3 * np.eye(3, dtype=np.float32)
        ok, inv = ol.inverse3(a)
        assert ok
        assert np.allclose(inv, np.linalg.inv(a.astype(np.float64)), rtol=1e-3, atol=1e-4)
    ok, _ = ol.inverse3(np.array([[1, 2, 3], [2, 4, 6], [1, 0, 1]], np.float32))
    assert not ok
    ok, _ = ol.inverse3(np.zeros((3, 3), np.float32))
    assert not ok


def test_linear_solve3_full_rank_and_min_norm():
    rng = np.random.default_rng(1)
    for _ in range(50):
        a = rng.normal(size=(3, 3)).astype(np.float32) + 2 * np.eye(3, dtype=np.float32)
        b = rng.normal(size=3).astype(np.float32)
        ok, x = ol.linear_solve3(a, b)
        assert ok
        assert np.allclose(x, np.linalg.solve(a.astype(np.float64), b.astype(np.float64)), rtol=2e-3, atol=2e-4)
    # rank 2 (third column zero): minimum-norm least squares, return value False
    a = np.array([[126025, 355, 0], [25, 5, 0], [225, 15, 0]], np.float32)
    b = np.array([0, 1, 0], np.float32)
    ok, x = ol.linear_solve3(a, b)
    assert not ok
    ref = np.linalg.lstsq(a.astype(np.float64), b.astype(np.float64), rcond=None)[0]
    assert np.allclose(x, ref, rtol=1e-3, atol=1e-7) and x[2] == 0


def test_vertex_parabola_constant_orientation():
    """the bin-0 parabola through (355,0),(5,S),(15,0) has its LS vertex at 177.4913 for any S > 0."""
    for s in (1e-3, 1.0, 77.5, 12345.678, 3.0e7):
        assert abs(ol.vertex_parabola(355, 0.0, 5, s, 15, 0.0) - 177.4913) < 2e-3


def test_find_peaks_typical_and_degenerate():
    h = np.zeros(36, np.float32)
    h[0] = 42.0
    p = ol.find_peaks(h)
    assert p.size == 1 and abs(p[0] - 177.4913) < 2e-3
    p = ol.find_peaks(np.zeros(36, np.float32))  # all-zero histogram: every vertex is NaN, the set keeps one
    assert p.size == 1 and np.isnan(p[0])
    h = np.zeros(36, np.float32)
    h[10], h[20] = 5.0, 4.5  # two real peaks -> two entries, ascending
    p = ol.find_peaks(h)
    assert p.size == 2 and p[0] < p[1]


# ---- schedule (Appendix B) ---------------------------------------------------------------------
def test_scale_schedule_and_nearest_gaussian():
    img = synth_frame(64, 48, 0)
    o = ol.Oracle(3, 2, 1.6, K, False)
    o.calculate(img)
    sig = np.float32(1.6)
    assert o.gauss(0, 0)[1] == sig and o.gauss(0, 1)[1] == sig
    assert o.dog(0, 0)[1] == 0.0
    assert abs(o.gauss(0, 2)[1] - 2.2627417) < 1e-6 and abs(o.gauss(0, 3)[1] - 3.2) < 1e-6
    assert o.gauss(1, 0)[1] == o.gauss(0, 2)[1] == o.gauss(1, 1)[1]
    assert abs(o.dog(0, 1)[1] - 0.6627417) < 1e-6 and abs(o.dog(1, 1)[1] - 0.9372583) < 1e-6
    assert o.level_dims(1) == (32, 24)
    # keypoint scales of octaves 0..3 all pick gaussians(0,0); 2.651 -> (0,2); 3.749 -> (0,3)   
    o6 = ol.Oracle(3, 6, 1.6, K, False)
    try:
        o6.calculate(synth_frame(64, 48, 0))
    except ol.OraclePrecondition:
        pass
    o4 = ol.Oracle(3, 2, 1.6, K, False)
    o4.calculate(img)
    for s in (0.6627, 0.9373, 1.3255, 1.8745):
        assert o4.nearest_gaussian(s) == (0, 0)


def test_too_small_image_raises_precondition():
    with pytest.raises(ol.OraclePrecondition):
        ol.Oracle(3, 4, 1.6, K, False).calculate(synth_frame(40, 40, 0))  # octave 3 is 5x5, radius 14


# ---- extrema / elimination / ordering ------------------------------------------------------------
def test_extrema_ties_count_on_flat_image():
    """on a flat DoG every interior pixel is a candidate (none greater OR none less)."""
    d = np.full((9, 12), 128.0, np.float32)
    xs, ys = ol.extrema(d, d, d)
    assert xs.size == (12 - 2) * (9 - 2)
    # x-outer / y-inner emission order
    assert np.array_equal(xs, np.repeat(np.arange(1, 11), 7)) and np.array_equal(ys, np.tile(np.arange(1, 8), 10))


def test_extrema_uses_the_2x2_half_open_neighbourhood():
    """With the layers below/above far smaller, (x, y) is a candidate iff it is the maximum of its own
    {x-1,x} x {y-1,y} block in the current layer: values at x+1 / y+1 must not matter (half-open subarray)."""
    rng = np.random.default_rng(9)
    cur = rng.permutation(20 * 17).reshape(17, 20).astype(np.float32) + 1000
    low = np.zeros_like(cur)
    xs, ys = ol.extrema(low, cur, low)
    blockmax = np.maximum(np.maximum(cur[1:, 1:], cur[:-1, 1:]), np.maximum(cur[1:, :-1], cur[:-1, :-1]))  # at (x>=1, y>=1)
    expect = (cur[1:, 1:] == blockmax)[: 17 - 2, : 20 - 2]  # interior x in [1, w-2], y in [1, h-2]
    got = np.zeros_like(expect)
    got[ys.astype(int) - 1, xs.astype(int) - 1] = True
    assert np.array_equal(got, expect) and 0 < xs.size < expect.size
    # a full 3x3 maximum test would give fewer candidates
    full = ndi.maximum_filter(cur, size=3) == cur
    assert expect.sum() > full[1:-1, 1:-1].sum()


def test_sort_order_is_a_partition_and_not_stable():
    rng = np.random.default_rng(3)
    flags = (rng.uniform(size=5000) > 0.07).astype(np.uint8)
    order = ol.sort_order(flags)
    assert np.array_equal(np.sort(order), np.arange(flags.size))
    sorted_flags = flags[order]
    n_unf = int((flags == 0).sum())
    assert not sorted_flags[:n_unf].any() and sorted_flags[n_unf:].all()
    assert not np.array_equal(order[:n_unf], np.flatnonzero(flags == 0))  # introsort permutes equal keys


def test_hoisted_flavour_equals_literal_flavour():
    img = synth_frame(72, 56, 11)
    a = ol.Oracle(3, 2, 1.6, K, False, literal=False).calculate(img)
    b = ol.Oracle(3, 2, 1.6, K, False, literal=True).calculate(img)
    assert a["x"].size == b["x"].size > 0
    for f in ("x", "y", "octave", "index", "scale", "orientation", "desc"):
        assert np.array_equal(a[f], b[f]), f


def test_descriptor_structure():
    """16 cells x 8 bins, bin 7 never written (o/45 % 7), every non-empty cell sums to 1 (L1)."""
    o = ol.Oracle(3, 2, 1.6, K, False)
    kp = o.calculate(synth_frame(96, 80, 3))
    d = kp["desc"].reshape(-1, 16, 8)
    assert d.shape[0] > 0 and (kp["desc_len"] == 128).all()
    assert (d[:, :, 7] == 0).all()
    sums = d.sum(-1)
    assert np.all((np.abs(sums - 1) < 1e-5) | (sums == 0))
    assert np.all(np.abs(kp["orientation"] - 177.4913) < 2e-3)


def test_result_text_format():
    o = ol.Oracle(3, 2, 1.6, K, False)
    kp = o.calculate(synth_frame(96, 80, 3))
    lines = o.text().split("\n")
    assert lines[0] == "Location\tscale\torientation\tdescriptors"
    first = lines[1].split("\t")
    assert first[0] == f"[{kp['x'][0]}, {kp['y'][0]}]" and first[1] in ("0.662742", "0.937258") and first[2] == "177.491"
    assert first[3].startswith("[") and first[3].endswith(", ]") and first[3].count(",") == 128
    assert len(lines) == kp["x"].size + 2


# ---- committed golden vectors ---------------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    o = ol.Oracle(3, int(g["octaves"]), 1.6, K, bool(g["subpixel"]))
    kp = o.calculate(g["img"].astype(np.float32))
    for oc in range(int(g["octaves"])):
        assert np.array_equal(o.dog(oc, 1)[0], g[f"dog_{oc}_1"])
    assert np.array_equal(o.gauss(int(g["octaves"]) - 1, 3)[0], g["g_last"])
    c = o.candidates()
    for f in ("x", "y", "octave", "index", "filtered"):
        assert np.array_equal(c[f], g["cand_" + f])
    for f in ("x", "y", "octave", "index", "scale", "orientation"):
        assert np.array_equal(kp[f], g["kp_" + f])
    assert np.array_equal(kp["desc"], g["desc"])


def test_parrot_counts(parrot):
    """Config 1 (example/parrot.jpg band 0, CLI defaults).  The survey's numpy model saw 24,870 / 1,556 / 1,506."""
    o = ol.Oracle(3, 4, 1.6, K, False)
    kp = o.calculate(parrot)
    assert o.candidates()["x"].size == 24869
    assert o.survivors()["x"].size == 1557
    assert kp["x"].size == 1507 and not kp["filtered"].any()
    assert o.level_dims(3) == (61, 75)
