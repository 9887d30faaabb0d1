"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/sift_gpu.h declares, create() fails loudly (no CPU fallback), the host replay of the
reference's std::sort equals the oracle's, the C++ host layer is built, and the data-parallel plumbing
works across two gloo ranks."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sift_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sift_gpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    from sift_b200 import capi

    lib = C.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sift_gpu.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == names
    lib.sift_gpu_version.restype = C.c_char_p
    assert b"sm_100a" in lib.sift_gpu_version()


def test_no_cpu_fallback_create_fails_without_gpu(built):
    import torch

    from sift_b200 import capi

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SiftGpuError) as e:
        capi.SiftGpu(3, 3, max_width=64, max_height=64)
    assert e.value.code == capi.E_CUDA


def test_reference_asserts_map_to_error_codes(built):
    from sift_b200 import capi

    for dpe, octv in ((2, 3), (3, 0)):
        with pytest.raises(capi.SiftGpuError) as e:
            capi.SiftGpu(dpe, octv, max_width=64, max_height=64)
        assert e.value.code == capi.E_ASSERT  # sift.cpp:382-383


def test_host_sort_replay_equals_oracle(built):
    import oracle_lib as ol
    from sift_b200 import capi

    rng = np.random.default_rng(0)
    for n, keep in ((0, 0.1), (1, 0.5), (15, 0.3), (16, 0.3), (17, 0.5), (1000, 0.06), (70000, 0.05), (200000, 0.06), (5000, 1.0), (5000, 0.0)):
        flags = (rng.uniform(size=n) >= keep).astype(np.uint8)
        assert np.array_equal(capi.sort_order(flags), ol.sort_order(flags)), (n, keep)


def test_sparse_sort_simulation_equals_std_sort(built):
    """The product replays libstdc++'s introsort on the sparse set of unfiltered positions (order_replay.h); the order
    it returns must be exactly where the real std::sort(cmpByFilter) puts the unfiltered elements."""
    from sift_b200 import capi

    rng = np.random.default_rng(1)
    cases = 0
    for n in list(range(0, 40)) + [63, 64, 65, 127, 129, 1000, 1025, 4096, 30000, 140000]:
        for keep in (0.0, 0.015, 0.06, 0.3, 0.5, 0.9, 1.0):
            for _ in range(2 if n > 2000 else 4):
                flags = (rng.uniform(size=n) >= keep).astype(np.uint8)
                full, fast = capi.sort_order(flags), capi.sort_order_fast(flags)
                assert np.array_equal(full[: int((flags == 0).sum())], fast), (n, keep)
                cases += 1
    for n in (17, 33, 100, 1000, 5000):  # structured patterns: long runs push the partition into its corner cases
        for pat in range(5):
            f = np.ones(n, np.uint8)
            if pat == 0: f[::2] = 0
            if pat == 1: f[(np.arange(n) // 7) % 2 == 0] = 0
            if pat == 2: f[n // 2:] = 0
            if pat == 3: f[: n // 2] = 0
            if pat == 4: f[n // 3: 2 * n // 3] = 0
            assert np.array_equal(capi.sort_order(f)[: int((f == 0).sum())], capi.sort_order_fast(f)), (n, pat)
    assert cases > 400


def test_cpp_host_layer_is_built(built):
    for f in ("libsift_host.so", "sift"):
        assert os.path.exists(os.path.join(ROOT, "sift_b200", f))
    out = subprocess.run([os.path.join(ROOT, "sift_b200", "sift"), "--help"], capture_output=True, text=True)
    assert out.returncode == 1 and "--dogsPerEpoch" in out.stdout  # main.cpp:47-50 returns 1 after printing the options


def test_shard_ranges_cover_everything():
    from sift_b200.shard import shard_range

    for n in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= hi - lo <= n // world + 1
                got += list(range(lo, hi))
            assert got == list(range(n))


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch
from sift_b200 import shard
rank, local, world = shard.world()
dist = shard.init_process_group("gloo")
lo, hi = shard.shard_range(4096, rank, world)
mx, sm = shard.reduce_max_sum(dist, torch.device("cpu"), [1.0 + rank, 5.0 - rank], [hi - lo, 1])
dist.barrier()
if rank == 0:
    print("RESULT", mx, sm)
"""


def test_two_rank_gloo_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
    assert "[2.0, 5.0]" in line and "[4096.0, 2.0]" in line


def test_bench_pyramid_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench

    assert abs(bench.pyramid_bytes(1920, 1080, 5, 3, False) / 1e6 - 118.75) < 0.01
    assert abs(bench.pyramid_bytes(488, 600, 4, 3, False) / 1e6 - 16.71) < 0.01      # config 1
    assert abs(bench.pyramid_bytes(600, 600, 4, 3, True) / 1e6 - 80.73) < 0.01       # config 2
    assert abs(bench.pyramid_bytes(3840, 2160, 6, 3, True) / 1e6 - 1868.44) < 0.05   # config 4


def test_keypoint_overlay_matches_the_reference_geometry(built):
    """main.cpp:59-75: rotated square of side int(scale*10) at (x*2^octave/div, y*2^octave/div), blue outline."""
    import ctypes

    lib = ctypes.CDLL(os.path.join(ROOT, "sift_b200", "libsift_host.so"))
    lib.sift_host_draw_points.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.sift_host_draw_points.restype = None

    def draw(pts, w=200, h=160, subpixel=0, fill=7):
        img = np.full((h, w, 3), fill, np.uint8)
        p = np.asarray(pts, np.float32).reshape(-1, 5)
        lib.sift_host_draw_points(img.ctypes.data, w, h, p.ctypes.data, p.shape[0], subpixel)
        return img

    blue = np.array([0, 0, 255], np.uint8)
    # orientation 0: axis-aligned square of side int(1.6*10) = 16 centred at (50, 40) * 2^1 = (100, 80)
    img = draw([[50, 40, 1, 1.6, 0.0]])
    on = np.all(img == blue, axis=2)
    ys, xs = np.nonzero(on)
    assert xs.min() == 92 and xs.max() == 108 and ys.min() == 72 and ys.max() == 88
    assert on[72, 92:109].all() and on[88, 92:109].all() and on[72:89, 92].all() and on[72:89, 108].all()
    assert on.sum() == 4 * 16 and not on[80, 100]          # outline only
    assert np.all(img[~on] == 7)                            # nothing else touched
    # 90 degrees gives the same square; 45 degrees a diamond whose corners sit on the axes
    assert np.array_equal(np.all(draw([[50, 40, 1, 1.6, 90.0]]) == blue, axis=2), on)
    d = np.all(draw([[50, 40, 1, 1.6, 45.0]]) == blue, axis=2)
    ys, xs = np.nonzero(d)
    assert abs((xs.max() - xs.min()) - 16 * np.sqrt(2)) <= 1.5 and d[80, xs.min()] and d[ys.min(), 100]
    # subpixel halves the coordinates (main.cpp:60); scale is not rescaled
    s = np.all(draw([[50, 40, 1, 1.6, 0.0]], subpixel=1) == blue, axis=2)
    ys, xs = np.nonzero(s)
    assert (xs.min(), xs.max(), ys.min(), ys.max()) == (42, 58, 32, 48)
    # clipping: a square hanging over the border, a huge one, NaN orientation (flat window) — no crash, pixels stay inside
    img = draw([[2, 2, 0, 3.0, 30.0], [100, 80, 0, 500.0, 12.0], [60, 60, 0, 2.0, float("nan")]])
    assert img.shape == (160, 200, 3)


def test_bench_reference_arm_prints_the_contract_line(built):
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU: one JSON line with the contract keys."""
    import json
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    import oracle_lib as ol

    # the reference's own sources when oracle/_ref is built (here, or prebuilt on the GPU box), else the restatement
    assert line["cpu_baseline"]["kind"] == ("reference" if ol.ref_available(True) else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_sparse_sort_simulation_on_clustered_and_threshold_patterns(built):
    """Survivors come in bursts (candidates are listed column by column), and the simulation switches to libstdc++'s own
    loop once half of a segment is unfiltered: exercise bursts, densities around that switch and the depth-limit fallback."""
    from sift_b200 import capi

    rng = np.random.default_rng(7)
    cases = 0
    for n in (200, 1905, 5000, 30000, 139000):
        for burst in (1, 3, 17, 200):
            for density in (0.002, 0.014, 0.1, 0.24, 0.26, 0.45, 0.5, 0.55, 0.8, 0.99):
                f = np.ones(n, np.uint8)
                n_bursts = max(1, int(n * density / burst))
                starts = rng.integers(0, max(1, n - burst), size=n_bursts)
                for s0 in starts:
                    f[s0:s0 + burst] = 0
                full, fast = capi.sort_order(f), capi.sort_order_fast(f)
                assert np.array_equal(full[: int((f == 0).sum())], fast), (n, burst, density)
                cases += 1
    # adversarial for median-of-three: sawtooth / organ-pipe layouts drive introsort towards its depth limit
    for n in (1000, 4097, 20000):
        i = np.arange(n)
        for f in ((i % 3 == 0), (i < n // 2) ^ (i % 2 == 0), (np.minimum(i, n - 1 - i) % 5 < 2), (i * 2654435761 % 97 < 40)):
            f = np.ascontiguousarray(f.astype(np.uint8))
            assert np.array_equal(capi.sort_order(f)[: int((f == 0).sum())], capi.sort_order_fast(f)), n
            cases += 1
    assert cases >= 200


def test_jpeg_front_end_library_exports_its_header(built):
    """include/sift_gpu_jpeg.h <-> libsift_gpu_jpeg.so: every declared entry point is exported (no compute without a GPU)."""
    import ctypes
    import re

    from sift_b200 import capi

    path = os.path.join(ROOT, "sift_b200", "libsift_gpu_jpeg.so")
    if not os.path.exists(path):
        pytest.skip("nvJPEG front end not built on this box")
    declared = set(re.findall(r"\b(sift_gpu_jpeg_\w+)\s*\(", open(os.path.join(ROOT, "include", "sift_gpu_jpeg.h")).read()))
    assert declared == set(capi.JPEG_EXPORTED_SYMBOLS)
    capi.load()
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), name
    # without a device the constructor fails loudly
    sift = None
    with pytest.raises(capi.SiftGpuError):
        sift = capi.SiftGpu(3, 3, max_width=64, max_height=64)
    assert sift is None



def _host_lib():
    import ctypes

    lib = ctypes.CDLL(os.path.join(ROOT, "sift_b200", "libsift_host.so"))
    lib.sift_host_read_image.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_int]
    lib.sift_host_read_image.restype = ctypes.c_int
    lib.sift_host_write_png.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.sift_host_write_png.restype = ctypes.c_int
    return lib


def _read_image(lib, path):
    import ctypes

    w, h = ctypes.c_int(0), ctypes.c_int(0)
    err = ctypes.create_string_buffer(256)
    if lib.sift_host_read_image(str(path).encode(), None, ctypes.byref(w), ctypes.byref(h), err, 256) != 0:
        raise ValueError(err.value.decode())
    px = np.zeros((h.value, w.value, 3), np.uint8)
    assert lib.sift_host_read_image(str(path).encode(), px.ctypes.data, ctypes.byref(w), ctypes.byref(h), err, 256) == 0
    return px


def test_cli_image_files_png_and_pnm(built, tmp_path):
    """main.cpp:52-54 / :59: the shim's own PNG reader (all five scanline filters, grey / RGB / RGBA / palette) and the PNM
    reader give the pixels PIL gives; the overlay writer (main.cpp:76) produces a PNG PIL reads back unchanged."""
    from PIL import Image

    lib = _host_lib()
    rng = np.random.default_rng(3)
    h, w = 37, 53
    smooth = (np.add.outer(np.arange(h) * 3, np.arange(w) * 2) % 256).astype(np.uint8)      # makes PIL pick Sub/Up/Average/Paeth
    rgb = np.stack([smooth, rng.integers(0, 256, (h, w)).astype(np.uint8), smooth[::-1]], axis=2)
    cases = {"rgb": Image.fromarray(rgb, "RGB"), "grey": Image.fromarray(smooth, "L"), "rgba": Image.fromarray(np.dstack([rgb, smooth]), "RGBA"),
             "palette": Image.fromarray(rgb, "RGB").quantize(17), "grey_alpha": Image.fromarray(np.dstack([smooth, smooth[:, ::-1]]), "LA")}
    for name, im in cases.items():
        p = tmp_path / f"{name}.png"
        im.save(p, optimize=(name == "rgb"))
        want = np.asarray(im.convert("RGB"))
        assert np.array_equal(_read_image(lib, p), want), name
    # binary PGM / PPM
    (tmp_path / "g.pgm").write_bytes(b"P5\n# comment\n53 37\n255\n" + smooth.tobytes())
    (tmp_path / "c.ppm").write_bytes(b"P6 53 37 255\n" + rgb.tobytes())
    assert np.array_equal(_read_image(lib, tmp_path / "g.pgm"), np.repeat(smooth[:, :, None], 3, 2))
    assert np.array_equal(_read_image(lib, tmp_path / "c.ppm"), rgb)
    # writer
    out = tmp_path / "out.png"
    assert lib.sift_host_write_png(str(out).encode(), rgb.ctypes.data, w, h) == 0
    assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), rgb)
    assert np.array_equal(_read_image(lib, out), rgb)
    # 16-bit grey: band 0 keeps the file's sample values (vigra::importImage does not scale), the colour view is the high byte
    import ctypes

    g16 = (rng.integers(0, 65536, (h, w))).astype(np.uint16)
    Image.fromarray(g16, "I;16").save(tmp_path / "g16.png")
    lib.sift_host_read_band0.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_int]
    lib.sift_host_read_band0.restype = ctypes.c_int
    ww, hh = ctypes.c_int(0), ctypes.c_int(0)
    b0 = np.zeros((h, w), np.float32)
    assert lib.sift_host_read_band0(str(tmp_path / "g16.png").encode(), b0.ctypes.data, ctypes.byref(ww), ctypes.byref(hh), None, 0) == 0
    assert (ww.value, hh.value) == (w, h) and np.array_equal(b0, g16.astype(np.float32))
    assert np.array_equal(_read_image(lib, tmp_path / "g16.png"), np.repeat((g16 >> 8).astype(np.uint8)[:, :, None], 3, 2))
    assert lib.sift_host_read_band0(str(tmp_path / "rgb.png").encode(), b0.ctypes.data, ctypes.byref(ww), ctypes.byref(hh), None, 0) == 0
    assert np.array_equal(b0, rgb[:, :, 0].astype(np.float32))
    # not an image / unsupported: an error message, no crash
    (tmp_path / "x.txt").write_bytes(b"hello world, not an image")
    with pytest.raises(ValueError):
        _read_image(lib, tmp_path / "x.txt")
    Image.fromarray(rgb, "RGB").save(tmp_path / "i.png", interlace=True) if False else None
    with pytest.raises(ValueError):
        _read_image(lib, tmp_path / "missing.png")


def test_order_replay_portable_dense_sort(tmp_path):
    """order_replay.h hands dense segments to libstdc++'s own introsort loop; where that internal is not available it uses its
    own restatement of the algorithm.  Both builds must reproduce the real std::sort permutation (tests/native/order_replay_check.cpp)."""
    src = os.path.join(ROOT, "tests", "native", "order_replay_check.cpp")
    for name, defs in (("own", []), ("lib", ["-DSIFT_ORDER_REPLAY_LIBSTDCXX_DENSE"])):
        exe = str(tmp_path / f"orc_{name}")
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", *defs, "-o", exe, src])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and "cases agree" in out.stdout, out.stdout


def test_worker_pool_back_to_back_jobs(tmp_path):
    """sift_b200/csrc/pool.h under sift_gpu_run's usage pattern (two pools, the next job begun right after the previous one
    was joined): every item exactly once, no hang — also under ThreadSanitizer where the toolchain has it
    (tests/native/pool_check.cpp).  The first version of the pool let a worker that was still leaving job k draw from job k+1."""
    src = os.path.join(ROOT, "tests", "native", "pool_check.cpp")
    exe = str(tmp_path / "pool_check")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, src, "-lpthread"])
    out = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "pool ok" in out.stdout, out.stdout + out.stderr
    tsan = str(tmp_path / "pool_check_tsan")
    if subprocess.run(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-o", tsan, src, "-lpthread"],
                      capture_output=True).returncode == 0:
        out = subprocess.run([tsan, "2000"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "pool ok" in out.stdout and "WARNING: ThreadSanitizer" not in out.stderr, out.stdout + out.stderr


def test_frame_packer_stays_inside_its_rows(tmp_path):
    """tests/native/pack_check.cpp: pack_host.cpp on ragged shapes with exact-size buffers, under ASan/UBSan where available."""
    srcs = [os.path.join(ROOT, "tests", "native", "pack_check.cpp"), os.path.join(ROOT, "sift_b200", "csrc", "pack_host.cpp")]
    exe = str(tmp_path / "pack_check")
    san = ["-fsanitize=address,undefined"]
    if subprocess.run(["/usr/bin/g++", "-O1", "-g", "-std=c++17", *san, "-o", exe, *srcs], capture_output=True).returncode != 0:
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, *srcs])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "pack edge cases ok" in out.stdout, out.stdout + out.stderr
