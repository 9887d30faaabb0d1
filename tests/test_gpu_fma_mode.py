"""SIFT_GPU_FLAG_FMA_BLUR on the BASELINE configs (-m gpu).  The opt-in throughput mode fuses multiply and add in the
Gaussian blur (one rounding per tap instead of the reference's two), so it is checked against north_star's tolerances
instead of bit-identity:
  * every DoG level within 1e-4 relative of the reference's (measured: < 1e-6);
  * >= 99 % of the reference's final keypoints present at the identical position (0 px <= 0.5 px), same octave and DoG level,
    identical scale label, orientation within one 10-degree bin; at most 1 % extra keypoints (measured: 100 %, 100 %, 99.9 %);
  * descriptors.  The reference accumulates descriptors IN PLACE in vector order after an unstable std::sort
    (F5): a single tie-candidate that appears or disappears (FMA mode: +2 of 24 869 on the parrot) shifts every later
    element of the candidate vector, the introsort permutes the survivors differently, and every keypoint whose window
    overlaps an earlier one sees different predecessors.  Against the reference's vector order the descriptor distance is
    therefore NOT bounded (measured median L2 0.8 - 3.7 of a maximum of 4); what FMA mode preserves is the descriptor as a
    function of (pyramid, order): with the order pinned (SIFT_GPU_FLAG_ORDER_CANONICAL on both sides) at least 99 % of the
    shared keypoints have L2 <= 1e-3 against the exact mode.
Because of the third point the exact mode stays the default and the headline; FMA mode is what it says: same keypoints,
faster pyramid, descriptors comparable only under a pinned order.
The exact mode is bit-identical to the reference build (tests/test_gpu_parity.py, tests/test_gpu_ref_pin.py)."""
import numpy as np
import pytest

import oracle_lib as ol  # checker only
import ref_cases as rc
from sift_b200 import capi

pytestmark = pytest.mark.gpu


def key(d):
    return (d["octave"].astype(np.int64) << 48) | (d["index"].astype(np.int64) << 32) | (d["x"].astype(np.int64) << 16) | d["y"].astype(np.int64)


@pytest.mark.parametrize("name", ["parrot", "600up", "1080p"])
def test_fma_blur_within_north_star_tolerances(built, name):
    make, p, _, _ = rc.CASES[name]
    img = make()
    h, w = img.shape
    mk = lambda flags: capi.SiftGpu(p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"], max_width=w, max_height=h, flags=flags)
    g = mk(capi.FLAG_FMA_BLUR)
    r = g.run([img])[0]
    o = ol.Oracle(p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"])
    okp = o.calculate(img)
    worst = 0.0
    for oc in range(p["octaves"]):
        for i in range(p["dpe"]):
            a, b = g.level(0, oc, i, capi.KIND_DOG)[0], o.dog(oc, i)[0]
            worst = max(worst, float(np.max(np.abs(a - b) / np.abs(b))))
    assert worst <= 1e-4, f"DoG relative error {worst}"

    k = r["kps"]
    kr, kg = key(okp), key(k)
    common, ir, ig = np.intersect1d(kr, kg, return_indices=True)
    n_ref = kr.size
    assert common.size >= 0.99 * n_ref, f"{common.size} of {n_ref} reference keypoints found"
    assert kg.size <= 1.01 * n_ref
    assert np.all(np.abs(k["orientation"][ig] - okp["orientation"][ir]) < 10.0)
    assert np.array_equal(k["scale"][ig], okp["scale"][ir])
    both = (k["desc_len"][ig] == 128) & (okp["desc_len"][ir] == 128)
    l2 = np.sqrt(((r["desc"][ig][both] - okp["desc"][ir][both]) ** 2).sum(1))
    same_order = kr.size == kg.size and bool(np.array_equal(kr, kg))
    print(f"{name}: DoG rel {worst:.2e}; {common.size}/{n_ref} keypoints shared, {kg.size} returned; candidates {r['n_candidates']} vs "
          f"{o.candidates()['x'].size}; vector order identical: {same_order}; descriptor L2 vs the reference's order: median {np.median(l2):.2e}, "
          f"{100 * float((l2 <= 1e-3).mean()):.1f} % <= 1e-3")
    g.close()

    # descriptors with the order pinned: canonical (octave, index, x, y) order on both sides, exact vs FMA
    ge, gf = mk(capi.FLAG_ORDER_CANONICAL), mk(capi.FLAG_ORDER_CANONICAL | capi.FLAG_FMA_BLUR)
    re_, rf = ge.run([img])[0], gf.run([img])[0]
    ke, kf = key(re_["kps"]), key(rf["kps"])
    common, ie, i_f = np.intersect1d(ke, kf, return_indices=True)
    assert common.size >= 0.99 * ke.size
    both = (re_["kps"]["desc_len"][ie] == 128) & (rf["kps"]["desc_len"][i_f] == 128)
    l2c = np.sqrt(((re_["desc"][ie][both] - rf["desc"][i_f][both]) ** 2).sum(1))
    frac = float((l2c <= 1e-3).mean())
    print(f"{name}: pinned order: {common.size}/{ke.size} shared, descriptor L2 median {np.median(l2c):.2e}, {100 * frac:.2f} % <= 1e-3, max {l2c.max():.3f}")
    assert frac >= 0.99
    ge.close(); gf.close()
