"""Parity cases shared by the reference-pinning tests (CPU: oracle vs oracle/_ref; GPU: CUDA path vs the committed
digests of oracle/_ref's outputs, tests/golden/ref_digests.json).  TEST INFRASTRUCTURE ONLY.

Every case is (image, ctor arguments of sift::Sift in sift.hpp:66-71 order).  `throws` marks cases where the
reference itself leaves calculate() with a vigra::PreconditionViolation (the dead 16x16 blur of sift.cpp:184 once
1.5*scale needs a radius above 15)."""
import hashlib
import os

import numpy as np

from sift_b200.synth import synth_frame

K = float(np.float32(np.sqrt(2.0)))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def lattice(w, h, pitch=10, blob_sigma=2.5, amp=180.0, seed=0):
    """Dense grid of small blobs (+-2 noise): about 5.8 % of the pixels survive elimination, so ~1.2 Mpixel are enough
    to push the survivor count past 65535 and exercise the u16 truncation of sift.cpp:41."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    dx, dy = (x % pitch) - pitch / 2, (y % pitch) - pitch / 2
    img = 40 + amp * np.exp(-(dx * dx + dy * dy) / (2 * blob_sigma * blob_sigma))
    img += np.random.default_rng(seed).uniform(-2, 2, (h, w))
    return np.clip(np.round(img), 0, 255).astype(np.float32)


def large_blobs(w, h, seed=0):
    """A dimmed synthetic frame plus ten blobs of sigma 25-70 px: structure coarse enough to leave survivors in
    octave 5 of a six-octave pyramid (the plain generator's blobs stop at sigma 14)."""
    img = synth_frame(w, h, seed).astype(np.float64) * 0.5
    rng = np.random.default_rng(1000 + seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    for _ in range(10):
        cx, cy = rng.uniform(0.15 * w, 0.85 * w), rng.uniform(0.15 * h, 0.85 * h)
        s, a = rng.uniform(25, 70), rng.uniform(60, 120)
        img += a * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * s * s))
    return np.clip(np.rint(img), 0, 255).astype(np.float32)


def _parrot():
    return np.load(os.path.join(GOLDEN, "parrot_r.npy")).astype(np.float32)


# name -> (image factory, dict(dpe, octaves, sigma, k, subpixel), throws, slow)
CASES = {
    # BASELINE.json configs 1-3
    "parrot": (_parrot, dict(dpe=3, octaves=4, sigma=1.6, k=K, subpixel=False), False, False),
    "600up": (lambda: synth_frame(600, 600, 0), dict(dpe=3, octaves=4, sigma=1.6, k=K, subpixel=True), False, False),
    "1080p": (lambda: synth_frame(1920, 1080, 0), dict(dpe=3, octaves=5, sigma=1.6, k=K, subpixel=False), False, False),
    # parameter / shape edge cases
    "dpe4": (lambda: synth_frame(640, 480, 3), dict(dpe=4, octaves=2, sigma=1.6, k=K, subpixel=False), False, False),
    "dpe4_oct3_throws": (lambda: synth_frame(640, 480, 3), dict(dpe=4, octaves=3, sigma=1.6, k=K, subpixel=False), True, False),
    "negative": (lambda: synth_frame(320, 240, 5) - 100.0, dict(dpe=3, octaves=3, sigma=1.6, k=K, subpixel=False), False, False),
    "flat": (lambda: np.full((96, 128), 50, np.float32), dict(dpe=3, octaves=2, sigma=1.6, k=K, subpixel=False), False, False),
    "sigma_k": (lambda: synth_frame(400, 300, 2), dict(dpe=3, octaves=3, sigma=1.2, k=1.3, subpixel=False), False, False),
    "ragged": (lambda: synth_frame(211, 157, 9), dict(dpe=3, octaves=3, sigma=1.6, k=K, subpixel=False), False, False),
    "sub_small": (lambda: synth_frame(160, 120, 4), dict(dpe=3, octaves=3, sigma=1.6, k=K, subpixel=True), False, False),
    "tiny": (lambda: synth_frame(24, 24, 1), dict(dpe=3, octaves=1, sigma=1.6, k=K, subpixel=False), False, False),
    # more than 65535 survivors: the u16 count wraps (sift.cpp:41) after the unstable sort
    "u16_wrap": (lambda: lattice(1120, 1120), dict(dpe=3, octaves=1, sigma=1.6, k=K, subpixel=False), False, True),
    # six octaves.  Octave-4 keypoints take gaussians(0,2) as their nearest Gaussian (sift.cpp:205-218); the plain
    # generator leaves no survivors in octave 5, so the reference finishes ...
    "oct6": (lambda: synth_frame(2048, 1800, 11), dict(dpe=3, octaves=6, sigma=1.6, k=K, subpixel=False), False, True),
    # ... while with coarse structure an octave-5 keypoint reaches the dead blur with radius 17 on a 16x16 window
    "oct6_throws": (lambda: large_blobs(2048, 1800, 0), dict(dpe=3, octaves=6, sigma=1.6, k=K, subpixel=False), True, True),
}


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()[:32]


def stage_digests(o, kp, p):
    """Digest of every stage of one finished calculate() on an oracle_lib.Oracle-like object `o`."""
    d = {"n_keypoints": int(kp["x"].size)}
    for oc in range(p["octaves"]):
        for i in range(p["dpe"] + 1):
            g, s = o.gauss(oc, i)
            d[f"gauss_{oc}_{i}"] = digest(g)
            d[f"gauss_scale_{oc}_{i}"] = float(s)
        for i in range(p["dpe"]):
            g, s = o.dog(oc, i)
            d[f"dog_{oc}_{i}"] = digest(g)
            d[f"dog_scale_{oc}_{i}"] = float(s)
    c = o.candidates()
    d["n_candidates"] = int(c["x"].size)
    d["n_unfiltered"] = int((c["filtered"] == 0).sum())
    for f in ("x", "y", "octave", "index", "filtered", "scale"):
        d[f"cand_{f}"] = digest(c[f])
    s = o.survivors()
    d["n_survivors"] = int(s["x"].size)
    for f in ("x", "y", "octave", "index"):
        d[f"surv_{f}"] = digest(s[f])
    for f in ("x", "y", "octave", "index", "scale", "orientation", "filtered", "desc_len"):
        d[f"kp_{f}"] = digest(kp[f])
    d["kp_desc"] = digest(kp["desc"])
    d["text"] = hashlib.sha256(o.text().encode()).hexdigest()[:32]
    return d


def pyramid_and_point_digests(o, kp, p):
    """The stages the shipped executable can be made to show (tests/refbin.py): every Gaussian and DoG level with its scale
    label, the candidate list with the elimination flags, the final keypoints with orientation and descriptors.  Same keys
    and digests as stage_digests(), minus the intermediate survivor list and the result text."""
    d = {"n_keypoints": int(kp["x"].size)}
    for oc in range(p["octaves"]):
        for i in range(p["dpe"] + 1):
            g, s = o.gauss(oc, i)
            d[f"gauss_{oc}_{i}"] = digest(g)
            d[f"gauss_scale_{oc}_{i}"] = float(s)
        for i in range(p["dpe"]):
            g, s = o.dog(oc, i)
            d[f"dog_{oc}_{i}"] = digest(g)
            d[f"dog_scale_{oc}_{i}"] = float(s)
    c = o.candidates()
    d["n_candidates"] = int(c["x"].size)
    d["n_unfiltered"] = int((c["filtered"] == 0).sum())
    for f in ("x", "y", "octave", "index", "filtered", "scale"):
        d[f"cand_{f}"] = digest(c[f])
    for f in ("x", "y", "octave", "index", "scale", "orientation", "filtered", "desc_len"):
        d[f"kp_{f}"] = digest(kp[f])
    d["kp_desc"] = digest(kp["desc"])
    return d
