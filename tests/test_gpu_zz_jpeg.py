"""GPU tests of the nvJPEG front end (include/sift_gpu_jpeg.h): JPEG bytes are decoded on the device and band 0 feeds the
SIFT path.  Parity is defined on the decoded plane: the oracle runs on exactly the pixels the device decoded and must
agree bit for bit; against PIL's decode of the same file only a loose closeness is asserted (decoders differ)."""
import os

import numpy as np
import pytest

import oracle_lib as ol  # checker only
from sift_b200 import capi

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
K = capi.SQRT2_F32


@pytest.fixture(scope="module")
def front(built):
    try:
        capi.load_jpeg()
    except OSError as e:  # libnvjpeg missing on this box: the core path does not depend on it
        pytest.skip(f"libsift_gpu_jpeg.so not loadable: {e}")
    g = capi.SiftGpu(3, 3, 1.6, K, False, max_width=320, max_height=240, max_batch=2)
    j = capi.SiftGpuJpeg(g, max_images=4)
    yield g, j
    j.close()
    g.close()


def _blob(name):
    return open(os.path.join(GOLDEN, name), "rb").read()


def test_jpeg_frames_match_the_oracle_on_the_decoded_plane(front):
    g, j = front
    blobs = [_blob("synth_rgb_320x240.jpg"), _blob("synth_grey_320x240.jpg"), _blob("synth_rgb_320x240.jpg")]
    res = j.run(blobs)
    assert len(res) == 3
    for i, r in enumerate(res):
        plane = j.decoded(i)
        assert plane.shape == (240, 320)
        okp = ol.Oracle(3, 3, 1.6, K, False).calculate(plane.astype(np.float32))
        assert r["status"] == 0 and r["kps"].size == okp["x"].size and r["kps"].size > 10
        for f in ("x", "y", "octave", "index", "scale", "orientation", "filtered"):
            assert np.array_equal(r["kps"][f], okp[f]), (i, f)
        assert np.array_equal(r["desc"], okp["desc"]), i
    # the same file twice gives the same result, and the front end gives what the plain u8 entry gives on the same pixels
    assert np.array_equal(res[0]["desc"], res[2]["desc"])
    direct = g.run([j.decoded(0)])[0]
    assert np.array_equal(direct["desc"], res[0]["desc"]) and np.array_equal(direct["kps"], res[0]["kps"])


def test_band0_is_the_red_channel_and_close_to_libjpeg(front):
    _, j = front
    j.run([_blob("synth_rgb_320x240.jpg"), _blob("synth_grey_320x240.jpg")])
    for i, name in enumerate(("synth_rgb_320x240_band0_pil.npy", "synth_grey_320x240_band0_pil.npy")):
        ref = np.load(os.path.join(GOLDEN, name)).astype(np.int32)
        got = j.decoded(i).astype(np.int32)
        d = np.abs(got - ref)
        assert d.mean() < 1.0 and d.max() <= 8, (name, float(d.mean()), int(d.max()))


def test_jpeg_errors(front):
    _, j = front
    with pytest.raises(capi.SiftGpuError):
        j.run([b"not a jpeg at all, just bytes" * 4])
    with pytest.raises(capi.SiftGpuError):
        j.run([_blob("synth_grey_320x240.jpg")] * 5)  # more than max_images
    assert len(j.run([])) == 0
