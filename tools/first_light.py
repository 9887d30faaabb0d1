"""Stage-by-stage comparison of the CUDA path (through the C ABI) with the CPU oracle on a few inputs.
Diagnostic script for GPU bring-up; the assertions proper live in tests/test_gpu_parity.py."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402  (checker only)
from sift_b200 import capi  # noqa: E402
from sift_b200.synth import synth_frame  # noqa: E402


def compare(name, img, octaves, subpixel=False, dpe=3, flags=0):
    h, w = img.shape
    g = capi.SiftGpu(dpe, octaves, 1.6, capi.SQRT2_F32, subpixel, max_width=w, max_height=h, max_batch=2, flags=flags)
    t0 = time.time()
    res = g.run([img, img])
    t_gpu = time.time() - t0
    t0 = time.time()
    res = g.run([img, img])
    t_gpu2 = time.time() - t0
    o = ol.Oracle(dpe, octaves, 1.6, capi.SQRT2_F32, subpixel)
    t0 = time.time()
    okp = o.calculate(img)
    t_cpu = time.time() - t0
    print(f"== {name}: {w}x{h} oct={octaves} sub={subpixel}  gpu(2 imgs) first {t_gpu*1e3:.1f} ms, second {t_gpu2*1e3:.1f} ms; oracle {t_cpu*1e3:.1f} ms")
    print("   timings", {k: round(v, 3) for k, v in g.timings().items()})
    bad = 0
    for oc in range(octaves):
        for i in range(dpe + 1):
            a, sa = g.level(0, oc, i, capi.KIND_GAUSS)
            b, sb = o.gauss(oc, i)
            eq = np.array_equal(a, b)
            if not eq or sa != sb:
                bad += 1
                print(f"   gauss({oc},{i}) equal={eq} maxabs={np.abs(a-b).max():.3e} scale {sa} vs {sb}")
        for i in range(dpe):
            a, sa = g.level(1, oc, i, capi.KIND_DOG)
            b, sb = o.dog(oc, i)
            eq = np.array_equal(a, b)
            if not eq or sa != sb:
                bad += 1
                print(f"   dog({oc},{i}) equal={eq} maxabs={np.abs(a-b).max():.3e} scale {sa} vs {sb}")
    print("   pyramid mismatching levels:", bad)
    gc, oc_ = g.candidates(0), o.candidates()
    same_c = all(np.array_equal(gc[k], oc_[k]) for k in ("x", "y", "octave", "index", "filtered"))
    print(f"   candidates gpu {gc['x'].size} oracle {oc_['x'].size} identical(list+flags)={same_c}")
    if not same_c and gc["x"].size == oc_["x"].size:
        for k in ("x", "y", "octave", "index", "filtered"):
            print("     field", k, "mismatches", int((gc[k] != oc_[k]).sum()))
    for r in res:
        k = r["kps"]
        print(f"   result: status {r['status']} n={k.size} cands={r['n_candidates']} surv={r['n_survivors']} | oracle n={okp['x'].size} surv={o.survivors()['x'].size}")
    k = res[0]["kps"]
    if k.size == okp["x"].size:
        same_k = all(np.array_equal(k[f], okp[f]) for f in ("x", "y", "octave", "index"))
        print("   keypoint order identical:", same_k, " scale equal:", np.array_equal(k["scale"], okp["scale"]))
        do = np.abs(k["orientation"] - okp["orientation"])
        print("   orientation bit-equal:", np.array_equal(k["orientation"], okp["orientation"]), "max diff", do.max() if do.size else 0)
        dd = np.abs(res[0]["desc"] - okp["desc"])
        l2 = np.sqrt(((res[0]["desc"] - okp["desc"]) ** 2).sum(1))
        print("   descriptors bit-equal rows:", int((dd.max(1) == 0).sum()), "/", k.size, " max L2", l2.max() if l2.size else 0)
    g.close()


def main():
    compare("synth64", synth_frame(64, 64, 0), 2)
    compare("synth200x150", synth_frame(200, 150, 1), 3)
    compare("parrot", np.load(os.path.join(ROOT, "tests", "golden", "parrot_r.npy")).astype(np.float32), 4)
    compare("synth600 sub", synth_frame(600, 600, 0), 4, subpixel=True)
    compare("1080p", synth_frame(1920, 1080, 0), 5)


if __name__ == "__main__":
    main()
