"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs, against the committed golden vectors, and — at BASELINE.json's full sizes — through
size-independent properties.  Bar: bit-exact for everything in the default (exact mul+add) mode; with
SIFT_GPU_FLAG_FMA_BLUR the DoG must agree within 1e-4 relative and >= 99 % of keypoints must match
(north_star tolerances)."""
import glob
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol  # checker only
from sift_b200 import capi
from sift_b200.synth import synth_frame

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
K = capi.SQRT2_F32


def full_compare(img, octaves, subpixel=False, dpe=3, batch=1):
    h, w = img.shape
    g = capi.SiftGpu(dpe, octaves, 1.6, K, subpixel, max_width=w, max_height=h, max_batch=batch)
    res = g.run([img] * batch)
    o = ol.Oracle(dpe, octaves, 1.6, K, subpixel)
    okp = o.calculate(img)
    b = batch - 1
    for oc in range(octaves):
        for i in range(dpe + 1):
            a, sa = g.level(b, oc, i, capi.KIND_GAUSS)
            r, sr = o.gauss(oc, i)
            assert sa == sr and np.array_equal(a, r), f"gaussian({oc},{i})"
        for i in range(dpe):
            a, sa = g.level(b, oc, i, capi.KIND_DOG)
            r, sr = o.dog(oc, i)
            assert sa == sr and np.array_equal(a, r), f"dog({oc},{i})"
    gc, occ = g.candidates(b), o.candidates()
    for f in ("x", "y", "octave", "index", "filtered"):
        assert np.array_equal(gc[f], occ[f]), f"candidate {f}"
    for r in res:
        assert r["status"] == 0
        assert r["n_candidates"] == occ["x"].size and r["n_survivors"] == o.survivors()["x"].size
        k = r["kps"]
        assert k.size == okp["x"].size
        for f in ("x", "y", "octave", "index", "scale", "orientation", "filtered"):
            assert np.array_equal(k[f], okp[f]), f"keypoint {f}"
        assert np.array_equal(k["desc_len"], okp["desc_len"])
        assert np.array_equal(r["desc"], okp["desc"]), "descriptors"
    g.close()
    return okp


@pytest.mark.parametrize("w,h,octaves,seed", [(64, 64, 2, 0), (200, 150, 3, 1), (97, 131, 3, 2), (333, 77, 2, 4)])
def test_small_synthetic_bit_exact(built, w, h, octaves, seed):
    full_compare(synth_frame(w, h, seed), octaves, batch=2)


@pytest.mark.parametrize("offset,seed,w,h", [(300.0, 1, 200, 150), (120.0, 1, 200, 150), (140.0, 3, 320, 240)])
def test_extra_orientation_peaks(built, offset, seed, w, h):
    """sift.cpp:194-200: a keypoint with several orientation peaks is appended once per peak (the first one too).
    Unreachable for non-negative images (every sample falls into bin 0 and the bin sum is positive), but a float
    image with negative grey values makes the bin sum negative and the histogram grows two peaks."""
    img = synth_frame(w, h, seed) - np.float32(offset)
    kp = full_compare(img, 3, batch=2)
    plain = ol.Oracle(3, 3, 1.6, K, False).calculate(synth_frame(w, h, seed))
    assert kp["x"].size > plain["x"].size and np.unique(kp["orientation"]).size > 2


def test_extra_orientation_peaks_mixed_batch(built):
    """Only one image of the pass takes the redo path; its neighbours must come out unchanged."""
    a, b = synth_frame(200, 150, 1), synth_frame(200, 150, 1) - np.float32(300)
    g = capi.SiftGpu(3, 3, 1.6, K, False, max_width=200, max_height=150, max_batch=4)
    res = g.run([a, b, a, b])
    ref = [ol.Oracle(3, 3, 1.6, K, False).calculate(x) for x in (a, b)]
    for i, r in enumerate(res):
        okp = ref[i % 2]
        assert r["status"] == 0 and r["kps"].size == okp["x"].size
        for f in ("x", "y", "octave", "index", "scale", "orientation", "filtered"):
            assert np.array_equal(r["kps"][f], okp[f], equal_nan=f == "orientation"), f
        assert np.array_equal(r["desc"], okp["desc"])
    g.close()


def test_repeated_runs_replay_the_captured_stage_graph(built):
    """The first pass of a shape on a slot runs eagerly, the second is captured into a CUDA graph, later ones replay it:
    every run must give the oracle's result, with different frames flowing through the same graph."""
    frames = [synth_frame(320, 240, s) for s in range(4)]
    ref = [ol.Oracle(3, 3, 1.6, K, False).calculate(f) for f in frames]
    g = capi.SiftGpu(3, 3, 1.6, K, False, max_width=320, max_height=240, max_batch=2)
    for rnd in range(4):
        order = [(rnd + i) % 4 for i in range(4)]
        res = g.run([frames[i] for i in order])
        for i, r in zip(order, res):
            okp = ref[i]
            assert r["status"] == 0 and r["kps"].size == okp["x"].size, (rnd, i)
            for f in ("x", "y", "octave", "index", "orientation", "filtered"):
                assert np.array_equal(r["kps"][f], okp[f]), (rnd, i, f)
            assert np.array_equal(r["desc"], okp["desc"]), (rnd, i)
    g.close()


def test_serial_context_and_long_batches_agree_with_the_pipelined_one(built):
    """SIFT_GPU_FLAG_SERIAL (one pass at a time, what the bench's roofline is timed on) and a batch of many passes (every
    slot reused several times, passes of 2, the last one ragged) must give the same results as the oracle."""
    frames = [synth_frame(256, 192, s) for s in range(3)]
    ref = [ol.Oracle(3, 3, 1.6, K, False).calculate(f) for f in frames]
    order = [i % 3 for i in range(15)]
    for flags in (0, capi.FLAG_SERIAL):
        g = capi.SiftGpu(3, 3, 1.6, K, False, max_width=256, max_height=192, max_batch=2, flags=flags)
        res = g.run([frames[i] for i in order])
        assert len(res) == 15
        for i, r in zip(order, res):
            okp = ref[i]
            assert r["status"] == 0 and r["kps"].size == okp["x"].size, (flags, i)
            for f in ("x", "y", "octave", "index", "orientation", "filtered"):
                assert np.array_equal(r["kps"][f], okp[f]), (flags, i, f)
            assert np.array_equal(r["desc"], okp["desc"]), (flags, i)
        g.close()


def test_config1_parrot_defaults(built, parrot):
    """BASELINE config 1: example/parrot.jpg band 0, sigma 1.6, k sqrt2, 4 octaves, 3 DoGs, subpixel 0."""
    kp = full_compare(parrot, 4)
    assert kp["x"].size == 1507


def test_config2_600_subpixel(built):
    """BASELINE config 2: 600x600 synthetic, subpixel=1 (2x upsampled base), 4 octaves."""
    full_compare(synth_frame(600, 600, 0), 4, subpixel=True)


def test_config3_1080p_5_octaves(built):
    """BASELINE config 3: 1920x1080 synthetic, 5 octaves."""
    kp = full_compare(synth_frame(1920, 1080, 0), 5, batch=2)
    assert kp["x"].size < 65536  # parity is only defined below the u16 wrap


def test_four_dogs_per_octave(built):
    full_compare(synth_frame(160, 120, 6), 2, dpe=4)


def test_non_default_sigma_k(built):
    img = synth_frame(180, 140, 8)
    g = capi.SiftGpu(3, 2, 1.2, 1.5, False, max_width=180, max_height=140)
    r = g.run([img])[0]
    o = ol.Oracle(3, 2, 1.2, 1.5, False)
    okp = o.calculate(img)
    assert np.array_equal(r["kps"]["x"], okp["x"]) and np.array_equal(r["kps"]["orientation"], okp["orientation"])
    assert np.array_equal(r["desc"], okp["desc"])
    g.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_against_committed_golden(built, path):
    gd = np.load(path)
    img = gd["img"].astype(np.float32)
    octaves, sub = int(gd["octaves"]), bool(gd["subpixel"])
    g = capi.SiftGpu(3, octaves, 1.6, K, sub, max_width=img.shape[1], max_height=img.shape[0])
    r = g.run([img])[0]
    for oc in range(octaves):
        assert np.array_equal(g.level(0, oc, 1, capi.KIND_DOG)[0], gd[f"dog_{oc}_1"])
    assert np.array_equal(g.level(0, octaves - 1, 3, capi.KIND_GAUSS)[0], gd["g_last"])
    c = g.candidates(0)
    for f in ("x", "y", "octave", "index", "filtered"):
        assert np.array_equal(c[f], gd["cand_" + f])
    for f in ("x", "y", "octave", "index", "scale", "orientation"):
        assert np.array_equal(r["kps"][f], gd["kp_" + f])
    assert np.array_equal(r["desc"], gd["desc"])
    g.close()


# ---- stage-level entry points ----------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx(built):
    g = capi.SiftGpu(3, 3, 1.6, K, False, max_width=512, max_height=512, max_batch=4)
    yield g
    g.close()


@pytest.mark.parametrize("w,h,sigma", [(6, 6, 1.6), (37, 23, 1.6), (130, 67, 3.2), (65, 33, 4.5254834), (300, 200, 9.0509668),
                                       (129, 257, 12.8), (64, 64, 1.0), (70, 80, 18.1019336), (33, 2, 0.1)])
def test_blur_bit_exact_ragged_sizes(ctx, w, h, sigma):
    rng = np.random.default_rng(w + 1000 * h)
    img = rng.integers(0, 256, (h, w)).astype(np.float32)
    assert np.array_equal(ctx.blur(img, sigma), ol.convolve(img, sigma))


def test_blur_precondition_maps_to_exception(ctx):
    with pytest.raises(capi.SiftGpuPrecondition):
        ctx.blur(np.zeros((5, 40), np.float32), 1.6)


@pytest.mark.parametrize("w,h", [(64, 48), (61, 75), (135, 270), (4, 3), (3, 5), (488, 512), (333, 251), (511, 203), (64, 201), (130, 509), (244, 300)])
def test_reduce_and_increase(ctx, w, h):
    img = synth_frame(max(w, 8), max(h, 8), w * h)[:h, :w].copy()
    sigma = 0.3 if min(w, h) < 6 else 1.6
    assert np.array_equal(ctx.reduce(img, sigma), ol.reduce(img, sigma))
    assert np.array_equal(ctx.increase(img, 0.3 if min(w, h) < 4 else 1.0), ol.increase(img, 0.3 if min(w, h) < 4 else 1.0))


def test_extrema_bit_exact_on_the_oracles_dog(ctx, parrot):
    """north_star: 'the extrema candidate set must be bit-exact when fed the reference's DoG'."""
    o = ol.Oracle(3, 2, 1.6, K, False)
    o.calculate(parrot)
    for oc in range(2):
        d0, d1, d2 = (o.dog(oc, i)[0] for i in range(3))
        gx, gy = ctx.extrema(d0, d1, d2)
        ox, oy = ol.extrema(d0, d1, d2)
        assert np.array_equal(gx, ox) and np.array_equal(gy, oy)
        assert np.array_equal(ctx.eliminate(d0, d1, d2, gx, gy), ol.eliminate(d0, d1, d2, ox, oy))


def test_extrema_flat_image_every_interior_pixel(ctx):
    d = np.full((70, 45), 128.0, np.float32)
    xs, ys = ctx.extrema(d, d, d)
    assert xs.size == 43 * 68
    assert np.array_equal(xs, np.repeat(np.arange(1, 44), 68)) and np.array_equal(ys, np.tile(np.arange(1, 69), 43))
    f = ctx.eliminate(d, d, d, xs, ys)
    assert f.all()  # singular Hessian -> inverse() fails -> rejected (sift.cpp:306)


def test_extrema_nan_and_ties(ctx):
    rng = np.random.default_rng(1)
    d = rng.integers(120, 136, (3, 40, 50)).astype(np.float32)  # many exact ties
    gx, gy = ctx.extrema(d[0], d[1], d[2])
    ox, oy = ol.extrema(d[0], d[1], d[2])
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and gx.size > 0
    assert np.array_equal(ctx.eliminate(d[0], d[1], d[2], gx, gy), ol.eliminate(d[0], d[1], d[2], ox, oy))


def test_eliminate_random_candidates(ctx):
    rng = np.random.default_rng(7)
    d = (128 + rng.normal(0, 6, (3, 60, 80))).astype(np.float32)
    xs = rng.integers(1, 79, 3000).astype(np.uint16)
    ys = rng.integers(1, 59, 3000).astype(np.uint16)
    a, b = ctx.eliminate(d[0], d[1], d[2], xs, ys), ol.eliminate(d[0], d[1], d[2], xs, ys)
    assert np.array_equal(a, b) and 0 < int((a == 0).sum()) < 3000


# ---- boundary behaviour -------------------------------------------------------------------------------
def test_u8_input_equals_f32_input(built):
    img = synth_frame(150, 110, 3)
    g = capi.SiftGpu(3, 2, max_width=150, max_height=110, max_batch=2)
    a = g.run([img])[0]
    b = g.run([img.astype(np.uint8)])[0]
    assert np.array_equal(a["kps"], b["kps"]) and np.array_equal(a["desc"], b["desc"])
    g.close()


def test_mixed_sizes_and_batch_position_independence(built):
    imgs = [synth_frame(120, 90, 1), synth_frame(120, 90, 2), synth_frame(90, 120, 3), synth_frame(120, 90, 1)]
    g = capi.SiftGpu(3, 2, max_width=120, max_height=120, max_batch=3)
    res = g.run(imgs)
    assert all(r["status"] == 0 for r in res)
    assert np.array_equal(res[0]["kps"], res[3]["kps"]) and np.array_equal(res[0]["desc"], res[3]["desc"])
    for im, r in zip(imgs, res):
        okp = ol.Oracle(3, 2, 1.6, K, False).calculate(im)
        assert np.array_equal(r["kps"]["x"], okp["x"]) and np.array_equal(r["desc"], okp["desc"])
    g.close()


def test_too_small_image_is_a_precondition_error(built):
    g = capi.SiftGpu(3, 4, max_width=64, max_height=64)
    with pytest.raises(capi.SiftGpuPrecondition):
        g.run([synth_frame(40, 40, 0)])
    res = g.run([synth_frame(40, 40, 0)], raise_on_error=False)
    assert res[0]["status"] == capi.E_PRECONDITION and res[0]["kps"].size == 0
    with pytest.raises(capi.SiftGpuError) as e:
        g.run([synth_frame(80, 40, 0)])
    assert e.value.code == capi.E_CAPACITY
    g.close()


def test_empty_batch_and_flat_image(built):
    g = capi.SiftGpu(3, 2, max_width=80, max_height=64)
    assert g.run([]) == []
    flat = np.full((64, 80), 77.0, np.float32)
    r = g.run([flat])[0]
    o = ol.Oracle(3, 2, 1.6, K, False)
    okp = o.calculate(flat)
    assert r["n_candidates"] == o.candidates()["x"].size == (78 * 62 + 38 * 30)
    assert r["kps"].size == okp["x"].size == 0
    g.close()


def test_canonical_order_gives_the_same_keypoint_set(built):
    img = synth_frame(400, 300, 5)
    a = capi.SiftGpu(3, 3, max_width=400, max_height=300)
    b = capi.SiftGpu(3, 3, max_width=400, max_height=300, flags=capi.FLAG_ORDER_CANONICAL)
    ra, rb = a.run([img])[0], b.run([img])[0]
    key = lambda k: sorted(zip(k["octave"].tolist(), k["index"].tolist(), k["x"].tolist(), k["y"].tolist()))
    assert key(ra["kps"]) == key(rb["kps"]) and ra["kps"].size > 50
    kb = rb["kps"]
    order = np.lexsort((kb["y"], kb["x"], kb["index"], kb["octave"]))
    assert np.array_equal(order, np.arange(kb.size))  # canonical = (octave, index, x, y)
    a.close(); b.close()


def test_fma_blur_within_north_star_tolerances(built):
    img = synth_frame(640, 480, 9)
    g = capi.SiftGpu(3, 4, max_width=640, max_height=480, flags=capi.FLAG_FMA_BLUR)
    r = g.run([img])[0]
    o = ol.Oracle(3, 4, 1.6, K, False)
    okp = o.calculate(img)
    for oc in range(4):
        for i in range(3):
            a, b = g.level(0, oc, i, capi.KIND_DOG)[0], o.dog(oc, i)[0]
            assert np.max(np.abs(a - b) / np.abs(b)) <= 1e-4  # DoG within 1e-4 relative
    ref = set(zip(okp["octave"].tolist(), okp["x"].tolist(), okp["y"].tolist()))
    got = set(zip(r["kps"]["octave"].tolist(), r["kps"]["x"].tolist(), r["kps"]["y"].tolist()))
    assert len(ref & got) >= 0.99 * len(ref) and len(got) <= 1.01 * len(ref)
    assert np.all(np.abs(r["kps"]["orientation"] - 177.4913) < 10.0)  # within one orientation bin
    g.close()


# ---- full-size properties (no oracle needed) --------------------------------------------------------
def test_1080p_properties_and_idempotence(built):
    imgs = [synth_frame(1920, 1080, s) for s in (1, 2)]
    g = capi.SiftGpu(3, 5, max_width=1920, max_height=1080, max_batch=2)
    r1 = g.run(imgs)
    # DoG = 128 + (g_next - g_prev): recompute from the returned Gaussians (linearity of the epilogue)
    for oc in (0, 4):
        ga, gb, d = g.level(1, oc, 1)[0], g.level(1, oc, 2)[0], g.level(1, oc, 1, capi.KIND_DOG)[0]
        assert np.array_equal(d, np.float32(128) + (gb - ga))
    # decimation: octave o+1 level 0 is a sub-sampling of blur(g(o,2)) -> its values all occur in the right rows
    assert g.level(1, 1, 0)[0].shape == (540, 960) and g.level(1, 4, 0)[0].shape == (68, 120)
    c = g.candidates(1)
    keyc = (c["octave"].astype(np.int64) << 40) | (c["index"].astype(np.int64) << 32) | (c["x"].astype(np.int64) << 16) | c["y"]
    assert np.all(np.diff(keyc) > 0)  # canonical emission order, no duplicates
    assert np.all((c["x"] >= 1) & (c["y"] >= 1))
    r2 = g.run(imgs)
    for a, b in zip(r1, r2):
        assert np.array_equal(a["kps"], b["kps"]) and np.array_equal(a["desc"], b["desc"])
        d = a["desc"].reshape(-1, 16, 8)
        s = d.sum(-1)
        assert np.all((np.abs(s - 1) < 1e-5) | (s == 0)) and (d[:, :, 7] == 0).all()
        assert 0 < a["kps"].size <= a["n_survivors"] < 65536
    g.close()


def test_config4_4k_subpixel_pyramid_only(built):
    """BASELINE config 4 (3840x2160, subpixel, 6 octaves) is graded on the pyramid: spot-check levels against
    oracle blurs of the device's own previous level (keeps the CPU work bounded)."""
    img = synth_frame(3840, 2160, 0)
    g = capi.SiftGpu(3, 6, 1.6, K, True, max_width=3840, max_height=2160, max_batch=1)
    r = g.run([img], raise_on_error=False)[0]
    assert r["status"] == 0 and r["out_width"] == 7680 and r["out_height"] == 4320
    for oc in (3, 4, 5):
        prev, _ = g.level(0, oc, 1)
        cur, s = g.level(0, oc, 2)
        assert np.array_equal(cur, ol.convolve(prev, s))
        assert np.array_equal(g.level(0, oc, 1, capi.KIND_DOG)[0], ol.dog(prev, cur))
    g2, s2 = g.level(0, 4, 2)
    assert np.array_equal(g.level(0, 5, 0)[0], ol.reduce(g2, s2))
    assert g.level(0, 5, 3)[0].shape == (135, 240)
    g.close()


def test_strict_mode_reproduces_the_dead_blur_exception(built):
    """with 6 octaves an octave-5 keypoint reaches sift.cpp:184 with radius 17 > 16."""
    img = synth_frame(2048, 1792, 3)
    g = capi.SiftGpu(3, 6, 1.6, K, False, max_width=2048, max_height=1792, flags=capi.FLAG_STRICT)
    r = g.run([img], raise_on_error=False)[0]
    lax = capi.SiftGpu(3, 6, 1.6, K, False, max_width=2048, max_height=1792)
    rl = lax.run([img])[0]
    has_oct5 = bool((rl["kps"]["octave"] == 5).any())
    assert (r["status"] == capi.E_PRECONDITION) == has_oct5
    g.close(); lax.close()


# ---- the C++ surface (sift::Sift + text writer) through the command-line shim --------------------------
def test_cli_text_output_equals_oracle_text(built, parrot, tmp_path):
    pgm = tmp_path / "parrot.pgm"
    with open(pgm, "wb") as f:
        f.write(b"P5\n488 600\n255\n" + parrot.astype(np.uint8).tobytes())
    out = tmp_path / "sift.txt"
    p = subprocess.run([os.path.join(ROOT, "sift_b200", "sift"), str(pgm), "-r", "1", "--out", str(out)], capture_output=True, text=True)
    assert p.returncode == 0 and "1507 interest points" in p.stdout, p.stderr
    o = ol.Oracle(3, 4, 1.6, K, False)
    o.calculate(parrot)
    assert open(out).read() == o.text()
    # main.cpp:59-76: the overlay image is written next to the input, as <image>_orientation.png
    from PIL import Image

    px = np.asarray(Image.open(str(pgm) + "_orientation.png").convert("RGB"))
    assert px.shape == (600, 488, 3)
    assert int(np.all(px == np.array([0, 0, 255], np.uint8), axis=2).sum()) > 1507  # every keypoint left an outline


def test_cli_reads_jpeg_and_png_band0(built, parrot, tmp_path):
    """`./sift example/parrot.jpg -r 1` (reference README): JPEG and PNG files go through the shim's own readers, band 0 feeds
    the pipeline (main.cpp:52-54).  JPEG decoders differ by a grey level here and there, so the check is against the oracle run
    on the plane the shim decoded (the same rule as the nvJPEG front end's test)."""
    import ctypes

    from PIL import Image

    rgb = np.stack([parrot.astype(np.uint8), np.roll(parrot, 7, 1).astype(np.uint8), np.flipud(parrot).astype(np.uint8)], axis=2)
    lib = ctypes.CDLL(os.path.join(ROOT, "sift_b200", "libsift_host.so"))
    lib.sift_host_read_image.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_int]
    for ext, kw in (("png", {}), ("jpg", {"quality": 95, "subsampling": 0})):
        f = tmp_path / f"parrot.{ext}"
        Image.fromarray(rgb, "RGB").save(f, **kw)
        w, h = ctypes.c_int(0), ctypes.c_int(0)
        err = ctypes.create_string_buffer(256)
        dec = np.zeros((600, 488, 3), np.uint8)
        assert lib.sift_host_read_image(str(f).encode(), dec.ctypes.data, ctypes.byref(w), ctypes.byref(h), err, 256) == 0, err.value
        assert (w.value, h.value) == (488, 600)
        band0 = dec[:, :, 0].astype(np.float32)
        if ext == "png":
            assert np.array_equal(band0, parrot)                       # lossless: band 0 is the R channel
        else:
            assert np.abs(band0 - parrot).mean() < 4.0                   # lossy, but the same picture
        out = tmp_path / f"sift_{ext}.txt"
        p = subprocess.run([os.path.join(ROOT, "sift_b200", "sift"), str(f), "-r", "1", "--out", str(out)], capture_output=True, text=True)
        assert p.returncode == 0 and "interest points" in p.stdout, p.stderr
        o = ol.Oracle(3, 4, 1.6, K, False)
        o.calculate(band0)
        assert open(out).read() == o.text(), ext
        assert os.path.exists(str(f) + "_orientation.png")
