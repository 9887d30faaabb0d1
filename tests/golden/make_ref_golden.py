"""Regenerates tests/golden/ref_digests.json from oracle/_ref/libref_fast.so — the REFERENCE'S OWN sift.cpp and
algorithms.cpp compiled unmodified against the Vigra stand-in headers (oracle/Makefile, target `ref`).  Run in the
build container (needs /root/reference):  python tests/golden/make_ref_golden.py

Each entry holds sha256 digests of every stage the reference produced for one case of tests/ref_cases.py (pyramid
levels, candidate list with flags, post-sort survivors, keypoints, descriptors, result text) or "throws".  The
CPU suite checks the oracle restatement against them; the GPU suite checks the CUDA path against them, so both
are pinned to outputs of the reference's own code even where oracle/_ref cannot be rebuilt (the GPU box)."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import ref_cases as rc  # noqa: E402


def main():
    L = ol.ref_lib(fast=True)
    out = {"_generator": "tests/golden/make_ref_golden.py", "_library": L.oracle_flavour().decode()}
    only = sys.argv[1:]
    path = os.path.join(HERE, "ref_digests.json")
    if only and os.path.exists(path):
        out = json.load(open(path))
    for name, (make, p, throws, _slow) in rc.CASES.items():
        if only and name not in only:
            continue
        t = time.time()
        r = ol.Oracle(p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"], L=L)
        try:
            kp = r.calculate(make())
            assert not throws, f"{name}: expected the reference to throw"
            out[name] = rc.stage_digests(r, kp, p)
        except ol.OraclePrecondition as e:
            assert throws, f"{name}: the reference threw: {e}"
            out[name] = {"throws": "PreconditionViolation"}
        print(f"{name}: {time.time() - t:.1f} s, {out[name].get('n_keypoints', 'throws')} keypoints", flush=True)
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
