"""Regenerates tests/golden/parrot_r.npy: band 0 (red) of the reference's example/parrot.jpg as uint8,
i.e. what vigra::importImage leaves in the scalar image before the float widening (main.cpp:52-54,
SURVEY A.8).  Run in the build container (needs /root/reference and PIL); the GPU box only sees the .npy."""
import numpy as np
from PIL import Image

img = np.asarray(Image.open("/root/reference/example/parrot.jpg"))
assert img.shape == (600, 488, 3), img.shape
np.save(__file__.replace("make_parrot_fixture.py", "parrot_r.npy"), np.ascontiguousarray(img[..., 0]))
print("ok", img[..., 0].mean())
