"""The C++ surface sift::Sift (include/sift/sift.hpp -> libsift_host.so) driven the way a caller of the reference
drives it (main.cpp:56-57, :90-92), through sift_b200/host_selftest (-m gpu): calculate() results and the image it
overwrites when subpixel (sift.cpp:21), calculateBatch(), and the exception path."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol  # checker only
from sift_b200 import capi
from sift_b200.synth import synth_frame

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = capi.SQRT2_F32


def drive(tmp_path, frames, dpe, octaves, subpixel):
    args = [os.path.join(ROOT, "sift_b200", "host_selftest"), str(tmp_path), str(dpe), str(octaves), str(int(subpixel))]
    for i, f in enumerate(frames):
        p = tmp_path / f"in_{i}.f32"
        f.astype(np.float32).tofile(p)
        args += [str(f.shape[1]), str(f.shape[0]), str(p)]
    return subprocess.run(args, capture_output=True, text=True)


@pytest.mark.parametrize("subpixel", [False, True])
def test_sift_class_calculate_batch_and_exceptions(built, tmp_path, subpixel):
    frames = [synth_frame(200, 150, 1), synth_frame(160, 120, 4), synth_frame(200, 150, 2)]
    p = drive(tmp_path, frames, 3, 3, subpixel)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "batch_images 3" in p.stdout
    assert f"first_width_after {400 if subpixel else 200}" in p.stdout  # calculateBatch overwrites its images too
    assert "caught PreconditionViolation" in p.stdout and "kernel longer than line" in p.stdout
    for i, f in enumerate(frames):
        o = ol.Oracle(3, 3, 1.6, K, subpixel)
        o.calculate(f)
        want = o.text()
        assert open(tmp_path / f"calc_{i}.txt").read() == want, f"calculate() text of image {i}"
        assert open(tmp_path / f"batch_{i}.txt").read() == want, f"calculateBatch() text of image {i}"
        w, h = (int(v) for v in open(tmp_path / f"img_{i}.dims").read().split())
        after = np.fromfile(tmp_path / f"img_{i}.f32", np.float32).reshape(h, w)
        if subpixel:  # sift.cpp:21: img = increaseToNextLevel(img, 1.0)
            assert (w, h) == (2 * f.shape[1], 2 * f.shape[0])
            assert np.array_equal(after, ol.increase(f, 1.0))
        else:
            assert np.array_equal(after, f)


def test_keep_upsampled_flag_through_the_c_abi(built):
    """SIFT_GPU_FLAG_KEEP_UPSAMPLED + sift_gpu_image.upsampled_out: the 2x image the reference leaves in `img`."""
    import ctypes as C

    img = synth_frame(96, 80, 3)
    g = capi.SiftGpu(3, 2, 1.6, K, True, max_width=96, max_height=80, flags=capi.FLAG_KEEP_UPSAMPLED)
    up = np.zeros((160, 192), np.float32)
    descs = (capi.Image * 1)()
    descs[0] = capi.Image(img.ctypes.data, 96, 80, 0, capi.DTYPE_F32, capi.MEM_HOST, up.ctypes.data)
    rc, res = g.run_raw(descs, 1)
    assert rc == 0 and res[0].out_width == 192 and res[0].out_height == 160
    assert np.array_equal(up, ol.increase(img, 1.0))
    g.close()
