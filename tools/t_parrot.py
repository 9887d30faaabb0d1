import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from sift_b200 import capi
img = np.load('/root/repo/tests/golden/parrot_r.npy').astype(np.float32)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
g = capi.SiftGpu(3, 4, 1.6, capi.SQRT2_F32, False, max_width=488, max_height=600, flags=flags)
r = g.run([img], raise_on_error=False)[0]
print('status', r['status'], r['kps'].size, g.L.sift_gpu_last_error(g.h))
