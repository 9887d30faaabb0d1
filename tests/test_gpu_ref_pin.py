"""GPU parity against outputs of the REFERENCE'S OWN code (-m gpu).

tests/golden/ref_digests.json holds sha256 digests of every stage that oracle/_ref/libref_fast.so — the reference's
sift.cpp + algorithms.cpp compiled unmodified against Vigra stand-in headers — produced for the cases of
tests/ref_cases.py (generator: tests/golden/make_ref_golden.py).  The CUDA path, called through the C ABI in its
default exact mode, must reproduce them bit for bit: pyramid levels and scale labels, candidate list with flags,
survivor count, keypoints in the reference's vector order, descriptors.  Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol  # checker only
import ref_cases as rc
from sift_b200 import capi

pytestmark = pytest.mark.gpu
DIGESTS = json.load(open(os.path.join(rc.GOLDEN, "ref_digests.json")))


def gpu_stage_digests(g, r, p):
    d = {"n_keypoints": int(r["kps"].size), "n_candidates": int(r["n_candidates"]), "n_survivors": int(r["n_survivors"])}
    for oc in range(p["octaves"]):
        for i in range(p["dpe"] + 1):
            a, s = g.level(0, oc, i, capi.KIND_GAUSS)
            d[f"gauss_{oc}_{i}"], d[f"gauss_scale_{oc}_{i}"] = rc.digest(a), float(s)
        for i in range(p["dpe"]):
            a, s = g.level(0, oc, i, capi.KIND_DOG)
            d[f"dog_{oc}_{i}"], d[f"dog_scale_{oc}_{i}"] = rc.digest(a), float(s)
    c = g.candidates(0)
    d["n_unfiltered"] = int((c["filtered"] == 0).sum())
    for f in ("x", "y", "octave", "index", "filtered"):
        d[f"cand_{f}"] = rc.digest(c[f])
    k = r["kps"]
    for f, dt in (("x", np.uint16), ("y", np.uint16), ("octave", np.uint16), ("index", np.uint16), ("scale", np.float32),
                  ("orientation", np.float32), ("filtered", np.uint8), ("desc_len", np.int32)):
        d[f"kp_{f}"] = rc.digest(k[f].astype(dt))
    d["kp_desc"] = rc.digest(r["desc"])
    return d


def run_gpu(name, flags=0):
    make, p, throws, _ = rc.CASES[name]
    img = make()
    h, w = img.shape
    g = capi.SiftGpu(p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"], max_width=w, max_height=h, flags=flags)
    r = g.run([img], raise_on_error=False)[0]
    return g, r, p, img


@pytest.mark.parametrize("name", [n for n, c in rc.CASES.items() if not c[2]])
def test_gpu_reproduces_the_reference_builds_output(built, name):
    g, r, p, _ = run_gpu(name)
    assert r["status"] == 0
    got, want = gpu_stage_digests(g, r, p), DIGESTS[name]
    bad = [k for k in got if k in want and got[k] != want[k]]
    assert not bad, f"stages differing from the reference build: {bad[:8]}"
    assert len([k for k in got if k in want]) >= 20
    g.close()


def test_u16_wrap_case_really_wraps(built):
    """sift.cpp:41: `u16_t size = distance(...)` — more than 65535 unfiltered candidates leave count mod 65536 keypoints."""
    w = DIGESTS["u16_wrap"]
    assert w["n_unfiltered"] > 65535 and w["n_survivors"] == w["n_unfiltered"] - 65536


@pytest.mark.parametrize("name", [n for n, c in rc.CASES.items() if c[2]])
def test_strict_mode_throws_where_the_reference_throws(built, name):
    """The reference leaves calculate() with a vigra::PreconditionViolation from the dead blur of sift.cpp:184;
    SIFT_GPU_FLAG_STRICT reports it as SIFT_GPU_E_PRECONDITION.  Without the flag the library skips the dead blur
    (its result is never used) and must then equal the non-strict oracle restatement."""
    assert DIGESTS[name] == {"throws": "PreconditionViolation"}
    g, r, p, img = run_gpu(name, flags=capi.FLAG_STRICT)
    assert r["status"] == capi.E_PRECONDITION and r["kps"].size == 0
    g.close()
    g, r, p, img = run_gpu(name)
    assert r["status"] == 0
    o = ol.Oracle(p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"], strict=False)
    okp = o.calculate(img)
    assert r["kps"].size == okp["x"].size > 0
    for f in ("x", "y", "octave", "index", "scale", "orientation", "filtered"):
        assert np.array_equal(r["kps"][f], okp[f]), f
    assert np.array_equal(r["desc"], okp["desc"])
    if p["octaves"] == 6:  # octave-4/5 keypoints: nearest Gaussian (0,2) / (0,3), sift.cpp:205-218
        assert (okp["octave"] == 5).any() and (okp["octave"] == 4).any()
        assert o.nearest_gaussian(float(okp["scale"][okp["octave"] == 5][0])) == (0, 3)
        assert o.nearest_gaussian(float(okp["scale"][okp["octave"] == 4][0])) == (0, 2)
    g.close()
