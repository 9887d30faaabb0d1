"""Writes the JPEG fixtures of tests/test_gpu_zz_jpeg.py: a colour JPEG whose three channels are different synthetic
frames (so "band 0" can only mean R) and a grey one, plus the planes PIL/libjpeg decodes from them (for a loose sanity
comparison: decoders differ by a few grey levels, parity is defined on the plane the device decoded).

    python tests/golden/make_jpeg_fixtures.py
"""
import io
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from sift_b200.synth import synth_frame  # noqa: E402

W, H = 320, 240
rgb = np.stack([synth_frame(W, H, s) for s in (5, 6, 7)], axis=2).astype(np.uint8)
buf = io.BytesIO()
Image.fromarray(rgb, "RGB").save(buf, "JPEG", quality=92, subsampling=0)  # 4:4:4: chroma upsampling differences stay out of it
open(os.path.join(HERE, "synth_rgb_320x240.jpg"), "wb").write(buf.getvalue())
np.save(os.path.join(HERE, "synth_rgb_320x240_band0_pil.npy"), np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))[:, :, 0])

grey = synth_frame(W, H, 9).astype(np.uint8)
buf = io.BytesIO()
Image.fromarray(grey, "L").save(buf, "JPEG", quality=95)
open(os.path.join(HERE, "synth_grey_320x240.jpg"), "wb").write(buf.getvalue())
np.save(os.path.join(HERE, "synth_grey_320x240_band0_pil.npy"), np.asarray(Image.open(io.BytesIO(buf.getvalue()))))
print("written", [f for f in os.listdir(HERE) if "320x240" in f])
