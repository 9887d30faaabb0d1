import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from sift_b200 import capi
from sift_b200.synth import synth_frame
img = synth_frame(3840, 2160, 0)
flags = (int(sys.argv[1]) if len(sys.argv) > 1 else 0) | capi.FLAG_SERIAL
g = capi.SiftGpu(3, 6, 1.6, capi.SQRT2_F32, True, max_width=3840, max_height=2160, max_batch=1, flags=flags)
for i in range(3):
    r = g.run([img], raise_on_error=False)[0]
print({k: round(v, 4) for k, v in g.timings().items()})
