"""Generates tests/golden/refbin_digests.json: SHA-256 digests of what the reference's SHIPPED EXECUTABLE
(/root/reference/bin/arch_x64/sift: GCC 7, real Vigra 1.11 compiled in) computes on the parity cases of tests/ref_cases.py,
obtained by calling its own functions in place (oracle/refbin_run.cpp, tests/refbin.py).  Run here, where /root/reference
exists; the JSON travels, the executable does not.  TEST INFRASTRUCTURE.

The executable is the literal reference (per-candidate image copies, a full-image blur per keypoint): the cases it finishes in
minutes to an hour are listed in CASES below (600up: 51 min, u16_wrap: 24 min); the six-octave cases would take many hours
and stay pinned through oracle/_ref (the reference's sources over stand-in headers) only.  The 1080p entry holds the result
vector of Sift::calculate alone (refbin.run_calculate; the stage-by-stage run does the work twice).

    python tests/golden/make_refbin_golden.py [case ...]
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_cases as rc  # noqa: E402
import refbin  # noqa: E402

CASES = ["tiny", "flat", "ragged", "negative", "sub_small", "sigma_k", "dpe4", "dpe4_oct3_throws", "parrot", "600up", "u16_wrap"]
CALCULATE_ONLY = ["1080p"]
OUT = os.path.join(HERE, "refbin_digests.json")


def main():
    assert refbin.available(), "needs /root/reference/bin/arch_x64/sift and oracle/_ref/refbin_run (make -C oracle)"
    names = sys.argv[1:] or CASES + CALCULATE_ONLY
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    out["_executable"] = {"path": "bin/arch_x64/sift of the reference repository",
                          "sha256": hashlib.sha256(open(refbin.EXE, "rb").read()).hexdigest(),
                          "how": "its own sift::Sift::_createDOGs / _findScaleSpaceExtrema / _eliminateEdgeResponses / calculate called in place (oracle/refbin_run.cpp)"}
    for name in names:
        make, p, throws, _ = rc.CASES[name]
        t = time.time()
        if name in CALCULATE_ONLY:
            k = refbin.run_calculate(make(), p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"])
            out[name] = {"n_keypoints": int(k["x"].size), "kp_desc": rc.digest(k["desc"]),
                         "only": "Sift::calculate (keypoints, orientations, descriptors); the stage-by-stage run would take twice as long"}
            for f in ("x", "y", "octave", "index", "scale", "orientation", "filtered", "desc_len"):
                out[name][f"kp_{f}"] = rc.digest(k[f])
            out[name]["seconds"] = round(time.time() - t, 1)
            print(name, out[name]["n_keypoints"], out[name]["seconds"], flush=True)
            json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
            continue
        s = refbin.run_stages(make(), p["dpe"], p["octaves"], p["sigma"], p["k"], p["subpixel"], timeout=4 * 3600)
        out[name] = {"throws": True} if s is None else rc.pyramid_and_point_digests(s, s.keypoints(), p)
        out[name]["seconds"] = round(time.time() - t, 1)
        assert (s is None) == throws, name
        print(name, "throws" if s is None else out[name]["n_keypoints"], out[name]["seconds"], flush=True)
        json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
