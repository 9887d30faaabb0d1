"""Regenerates the small golden vectors under tests/golden/ from the CPU oracle (oracle/).

The reference ships no golden vectors, so these record the ORACLE's outputs: they make its behaviour
auditable and let the GPU tests run against committed numbers.  (The oracle itself is held to the reference's source build
and to its shipped executable by tests/test_ref_pin.py and tests/test_refbin_pin.py.)
Inputs are the seeded synthetic generator and a crop of the reference's example/parrot.jpg (band 0).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from sift_b200.synth import synth_frame  # noqa: E402

K = float(np.float32(np.sqrt(2.0)))


def make(name, img, octaves, subpixel):
    o = ol.Oracle(3, octaves, 1.6, K, subpixel)
    kp = o.calculate(img)
    c = o.candidates()
    d = {"img": img.astype(np.uint8), "octaves": octaves, "subpixel": int(subpixel)}
    for oc in range(octaves):
        d[f"dog_{oc}_1"] = o.dog(oc, 1)[0]
    d["g_last"] = o.gauss(octaves - 1, 3)[0]
    for f in ("x", "y", "octave", "index", "filtered"):
        d["cand_" + f] = c[f]
    for f in ("x", "y", "octave", "index", "scale", "orientation"):
        d["kp_" + f] = kp[f]
    d["desc"] = kp["desc"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, img.shape, "cands", c["x"].size, "kps", kp["x"].size)


if __name__ == "__main__":
    make("synth96_oct2", synth_frame(96, 80, 3), 2, False)
    make("synth48_sub_oct2", synth_frame(48, 40, 5), 2, True)
    parrot = np.load(os.path.join(HERE, "parrot_r.npy")).astype(np.float32)
    make("parrot_crop128_oct3", np.ascontiguousarray(parrot[200:328, 180:308]), 3, False)
