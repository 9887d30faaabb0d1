"""Seeded synthetic frame generator: 8-bit-valued float32 grayscale made of a
background, isotropic Gaussian blobs, axis-aligned rectangles ("corners") and +-2 grey levels of
iid noise so that no neighbourhood is exactly flat (the reference's tie-inclusive extrema test,
sift.cpp:366-371, would otherwise turn every flat pixel into a candidate)."""
import numpy as np


def synth_frame(w, h, seed=0):
    """Return an (h, w) float32 array with integer values in [0, 255]; seed = frame index."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 40.0, np.float64)
    n = max(8, (w * h) // 20000)
    # blobs
    cx = rng.uniform(0, w, n)
    cy = rng.uniform(0, h, n)
    sg = rng.uniform(2.0, 14.0, n)
    am = rng.uniform(0.0, 180.0, n)
    for i in range(n):
        r = int(np.ceil(4.0 * sg[i]))
        x0, x1 = max(0, int(cx[i]) - r), min(w, int(cx[i]) + r + 1)
        y0, y1 = max(0, int(cy[i]) - r), min(h, int(cy[i]) + r + 1)
        if x0 >= x1 or y0 >= y1:
            continue
        xs = np.arange(x0, x1)[None, :] - cx[i]
        ys = np.arange(y0, y1)[:, None] - cy[i]
        img[y0:y1, x0:x1] += am[i] * np.exp(-(xs * xs + ys * ys) / (2.0 * sg[i] * sg[i]))
    # corners (rectangles)
    rx = rng.uniform(0, w, n)
    ry = rng.uniform(0, h, n)
    sw = rng.uniform(8.0, 68.0, n)
    sh = rng.uniform(8.0, 68.0, n)
    st = rng.uniform(0.0, 120.0, n)
    for i in range(n):
        x0, x1 = max(0, int(rx[i])), min(w, int(rx[i] + sw[i]))
        y0, y1 = max(0, int(ry[i])), min(h, int(ry[i] + sh[i]))
        img[y0:y1, x0:x1] += st[i]
    img += rng.uniform(-2.0, 2.0, (h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.float32)


def synth_batch(w, h, n, first_seed=0):
    return np.stack([synth_frame(w, h, first_seed + i) for i in range(n)])
