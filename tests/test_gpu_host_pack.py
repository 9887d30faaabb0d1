"""GPU tests (-m gpu) of the packed upload: host f32 frames whose pixels are exact 8-bit values (what importImage delivers,
main.cpp:52-54) are packed to bytes on host threads ahead of their pass, uploaded as bytes and widened on the device.
The results must be bit-identical to the plain f32 upload and to the oracle; frames that are not 8-bit valued must make
their pass travel as f32; the byte counters must say what actually travelled."""
import numpy as np
import pytest

import oracle_lib as ol  # checker only
from sift_b200 import capi
from sift_b200.synth import synth_frame

pytestmark = pytest.mark.gpu
K = capi.SQRT2_F32
# SIFT_GPU_HOST_PACK: "0" off, "1" whole passes (stage A waits for the packers: deterministic counts), "2" split (it does not:
# whatever is packed when the pass is due goes up as bytes, the rest as f32)


def same(a, b):
    assert len(a) == len(b)
    for ra, rb in zip(a, b):
        assert ra["status"] == rb["status"] == 0
        assert np.array_equal(ra["kps"], rb["kps"]) and np.array_equal(ra["desc"], rb["desc"])
        assert ra["n_candidates"] == rb["n_candidates"] and ra["n_survivors"] == rb["n_survivors"]


def vs_oracle(r, img, octaves):
    okp = ol.Oracle(3, octaves, 1.6, K, False).calculate(img)
    assert r["kps"].size == okp["x"].size
    for f in ("x", "y", "octave", "index", "scale", "orientation", "filtered"):
        assert np.array_equal(r["kps"][f], okp[f]), f
    assert np.array_equal(r["desc"], okp["desc"])


def test_packed_upload_is_lossless_and_counted(built, monkeypatch):
    # ragged width: the staging pitch (224) differs from the width, one transfer per image
    frames = [synth_frame(200, 150, s) for s in range(12)]
    g = capi.SiftGpu(3, 3, max_width=200, max_height=150, max_batch=12)
    monkeypatch.setenv("SIFT_GPU_HOST_PACK", "1")
    a, ta = g.run(frames), g.timings()
    assert ta["packed_images"] == 12 and ta["h2d_bytes"] == 12 * 224 * 150
    monkeypatch.setenv("SIFT_GPU_HOST_PACK", "0")
    b, tb = g.run(frames), g.timings()
    assert tb["packed_images"] == 0 and tb["h2d_bytes"] == 12 * 200 * 150 * 4
    same(a, b)
    vs_oracle(a[0], frames[0], 3)
    vs_oracle(a[7], frames[7], 3)
    g.close()


def test_a_frame_that_is_not_8_bit_valued_sends_its_pass_as_f32(built, monkeypatch):
    # dense layout (width a multiple of 32): the packed pass goes up as one transfer
    frames = [synth_frame(256, 96, 20 + s) for s in range(9)]
    g = capi.SiftGpu(3, 2, max_width=256, max_height=96, max_batch=9)
    monkeypatch.setenv("SIFT_GPU_HOST_PACK", "1")
    a, ta = g.run(frames), g.timings()
    assert ta["packed_images"] == 9 and ta["h2d_bytes"] == 9 * 256 * 96
    vs_oracle(a[3], frames[3], 2)
    for bad in (0.5, -3.0, 300.0, -0.0):
        odd = [f.copy() for f in frames]
        odd[4][50, 100] = bad if bad != 0.5 else odd[4][50, 100] + 0.5
        c, tc = g.run(odd), g.timings()
        assert tc["packed_images"] == 0 and tc["h2d_bytes"] == 9 * 256 * 96 * 4, bad
        monkeypatch.setenv("SIFT_GPU_HOST_PACK", "0")
        d = g.run(odd)
        monkeypatch.setenv("SIFT_GPU_HOST_PACK", "1")
        same(c, d)
    vs_oracle(c[4], odd[4], 2)   # the frame holding -0.0f: still what the reference computes on it
    g.close()


def test_packing_ahead_across_many_passes_and_small_or_non_f32_passes_are_left_alone(built, monkeypatch):
    frames = [synth_frame(160, 128, 40 + s) for s in range(5)]
    seq = [frames[i % 5] for i in range(28)]          # passes of 8, 8, 8, 4: the last one is below the packing minimum
    monkeypatch.setenv("SIFT_GPU_HOST_PACK", "1")
    for flags in (0, capi.FLAG_SERIAL):
        g = capi.SiftGpu(3, 3, max_width=160, max_height=128, max_batch=8, flags=flags)
        for _ in range(2):                                # second round: captured graphs, reused staging
            a, ta = g.run(seq), g.timings()
            assert ta["packed_images"] == 24, flags
        monkeypatch.setenv("SIFT_GPU_HOST_PACK", "0")
        b = g.run(seq)
        monkeypatch.setenv("SIFT_GPU_HOST_PACK", "1")
        same(a, b)
        same(a[:5], a[5:10])
        # u8 frames and a mixed sequence (u8 passes between f32 passes)
        u8 = [f.astype(np.uint8) for f in seq[:8]]
        c, tcn = g.run(seq[:8] + u8 + seq[:8]), g.timings()
        assert tcn["packed_images"] == 16 and tcn["h2d_bytes"] == 24 * 160 * 128
        same(c[:8], a[:8]); same(c[8:16], a[:8]); same(c[16:], a[:8])
        g.close()
    vs_oracle(a[2], frames[2], 3)


def test_split_upload_mixes_bytes_and_floats_inside_a_pass(built, monkeypatch):
    """Mode 2 never waits for the packers: a pass goes up as n packed frames followed by nb - n raw ones, n anywhere in
    [0, nb].  Whatever n turns out to be, the results are those of the plain upload and the counters add up."""
    frames = [synth_frame(192, 144, 60 + s) for s in range(6)]
    seq = [frames[i % 6] for i in range(44)]          # passes of 10, 10, 10, 10, 4
    g = capi.SiftGpu(3, 3, max_width=192, max_height=144, max_batch=10)
    monkeypatch.setenv("SIFT_GPU_HOST_PACK", "0")
    ref = g.run(seq)
    monkeypatch.setenv("SIFT_GPU_HOST_PACK", "2")
    seen = set()
    for _ in range(4):
        a, t = g.run(seq), g.timings()
        n = int(t["packed_images"])
        assert 0 <= n <= 40
        assert t["h2d_bytes"] == n * 192 * 144 + (44 - n) * 192 * 144 * 4
        same(a, ref)
        seen.add(n)
    monkeypatch.delenv("SIFT_GPU_HOST_PACK")          # the library's own policy
    same(g.run(seq), ref)
    vs_oracle(a[3], frames[3], 3)
    vs_oracle(a[43], frames[43 % 6], 3)
    g.close()
