"""One device pass of 1080p frames through the C ABI, for `ncu` (development aid).
usage: profile_pass.py [--fma] [--frames 16] [--runs 2] [--w 1920 --h 1080 --octaves 5] [--subpixel]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sift_b200 import capi  # noqa: E402
from sift_b200.synth import synth_frame  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fma", action="store_true")
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--octaves", type=int, default=5)
ap.add_argument("--subpixel", action="store_true")
a = ap.parse_args()
os.environ.setdefault("SIFT_GPU_GRAPHS", "0")
flags = capi.FLAG_SERIAL | (capi.FLAG_FMA_BLUR if a.fma else 0)
g = capi.SiftGpu(3, a.octaves, 1.6, capi.SQRT2_F32, a.subpixel, max_width=a.w, max_height=a.h, max_batch=a.frames, flags=flags)
frames = [synth_frame(a.w, a.h, s % 4).astype(np.uint8) for s in range(a.frames)]
for _ in range(a.runs):
    r = g.run(frames, raise_on_error=False)
t = g.timings()
print({k: round(v / a.frames, 5) if k.endswith("_ms") else v for k, v in t.items()})
g.close()
