"""Drives oracle/_ref/refbin_run (oracle/refbin_run.cpp): code of the reference's SHIPPED EXECUTABLE
(/root/reference/bin/arch_x64/sift — GCC 7, real Vigra 1.11 compiled in) called in place.  TEST INFRASTRUCTURE ONLY.

`run_stages(img, dpe, octaves, sigma, k, subpixel)` returns an object with the accessors of oracle_lib.Oracle
(gauss, dog, candidates, keypoints) filled from what the executable's own Sift::_createDOGs, _findScaleSpaceExtrema,
_eliminateEdgeResponses and Sift::calculate produced, or None when the executable left with a vigra exception."""
import os
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.environ.get("SIFT_REF_EXE", "/root/reference/bin/arch_x64/sift")
RUN = os.path.join(ROOT, "oracle", "_ref", "refbin_run")


_probe = None


def available():
    """The executable and the helper exist AND the helper can map and run it here (its segments want fixed low addresses: an
    environment that refuses those makes the live tests skip, not fail)."""
    global _probe
    if _probe is None:
        _probe = False
        if os.path.exists(EXE) and os.path.exists(RUN):
            try:
                a = np.arange(64, dtype=np.float32).reshape(8, 8)
                _probe = len(run_unit(a, 0.5, timeout=60)) == 3
            except Exception:
                _probe = False
    return _probe


class Stages:
    def __init__(self, octaves, dpe, g, d, cands, kps):
        self.octaves, self.dpe, self._g, self._d, self._c, self._k = octaves, dpe, g, d, cands, kps

    def gauss(self, o, i):
        return self._g[o, i]

    def dog(self, o, i):
        return self._d[o, i]

    def candidates(self):
        return self._c

    def keypoints(self):
        return self._k


def _points(buf, off, with_desc):
    (n,) = struct.unpack_from("<i", buf, off)
    off += 4
    d = dict(x=np.zeros(n, np.uint16), y=np.zeros(n, np.uint16), octave=np.zeros(n, np.uint16), index=np.zeros(n, np.uint16),
             scale=np.zeros(n, np.float32), filtered=np.zeros(n, np.uint8))
    if with_desc:
        d["orientation"] = np.zeros(n, np.float32)
        d["desc"] = np.zeros((n, 128), np.float32)
        d["desc_len"] = np.zeros(n, np.int32)
    for j in range(n):
        x, y, oc, ix, sc, ori, fl = struct.unpack_from("<HHHHffB", buf, off)
        off += 17
        d["x"][j], d["y"][j], d["octave"][j], d["index"][j], d["scale"][j], d["filtered"][j] = x, y, oc, ix, sc, fl
        if with_desc:
            (nd,) = struct.unpack_from("<i", buf, off)
            off += 4
            d["orientation"][j], d["desc_len"][j] = ori, nd
            d["desc"][j, :nd] = np.frombuffer(buf, np.float32, nd, off)
            off += 4 * nd
    return d, off


def _call(cmd, payload, timeout):
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        with open(fin, "wb") as f:
            f.write(payload)
        r = subprocess.run([RUN, EXE, cmd, fin, fout], capture_output=True, text=True, timeout=timeout)
        if r.returncode not in (0, 3):
            raise RuntimeError(f"refbin_run failed ({r.returncode}): {r.stderr[-400:]}")
        with open(fout, "rb") as f:
            return r.returncode, f.read()


def run_stages(img, dpe, octaves, sigma, k, subpixel, timeout=1200):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    rc, buf = _call("stages", struct.pack("<iiiiffi", w, h, dpe, octaves, sigma, k, int(subpixel)) + img.tobytes(), timeout)
    if rc == 3:
        return None   # vigra::PreconditionViolation (or another std::exception) left the executable's code
    off = 0
    O, D = struct.unpack_from("<ii", buf, off)
    off += 8

    def image():
        nonlocal off
        ww, hh, sc = struct.unpack_from("<iif", buf, off)
        off += 12
        a = np.frombuffer(buf, np.float32, ww * hh, off).reshape(hh, ww).copy()
        off += 4 * ww * hh
        return a, sc

    g, d = {}, {}
    for o in range(O):
        for i in range(D + 1):
            g[o, i] = image()
        for i in range(D):
            d[o, i] = image()
    cands, off = _points(buf, off, False)
    kps, off = _points(buf, off, True)
    assert off == len(buf)
    return Stages(O, D, g, d, cands, kps)


def run_unit(img, sigma, timeout=300):
    """alg::convolveWithGauss, reduceToNextLevel, increaseToNextLevel of the executable on one image."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    _, buf = _call("unit", struct.pack("<iif", w, h, sigma) + img.tobytes(), timeout)
    out, off = [], 0
    for _ in range(3):
        ww, hh, _sc = struct.unpack_from("<iif", buf, off)
        off += 12
        out.append(np.frombuffer(buf, np.float32, ww * hh, off).reshape(hh, ww).copy())
        off += 4 * ww * hh
    return out


def run_eliminate(d0, d1, d2, xs, ys, timeout=600):
    """Sift::_eliminateEdgeResponses of the executable (Vigra's own inverse + linearSolve) on three DoG layers: one flag per point."""
    d = [np.ascontiguousarray(a, np.float32) for a in (d0, d1, d2)]
    h, w = d[1].shape
    xy = np.stack([np.asarray(xs, np.uint16), np.asarray(ys, np.uint16)], axis=1).astype(np.uint16)
    _, buf = _call("eliminate", struct.pack("<iii", w, h, xy.shape[0]) + b"".join(a.tobytes() for a in d) + xy.tobytes(), timeout)
    return np.frombuffer(buf, np.uint8).copy()


def run_vertex(triples, timeout=120):
    """alg::vertexParabola of the executable on rows (lx, ly, px, py, rx, ry)."""
    t = np.ascontiguousarray(triples, np.float32).reshape(-1, 6)
    _, buf = _call("vertex", struct.pack("<i", t.shape[0]) + t.tobytes(), timeout)
    return np.frombuffer(buf, np.float32).copy()


def run_peaks(histos, timeout=120):
    """Sift::_findPeaks of the executable on rows of 36 bins: list of the returned std::set contents (ascending, NaN as stored)."""
    hs = np.ascontiguousarray(histos, np.float32).reshape(-1, 36)
    _, buf = _call("peaks", struct.pack("<i", hs.shape[0]) + hs.tobytes(), timeout)
    out, off = [], 0
    for _ in range(hs.shape[0]):
        (n,) = struct.unpack_from("<i", buf, off)
        off += 4
        out.append(np.frombuffer(buf, np.float32, n, off).copy())
        off += 4 * n
    return out


def time_calculate(img, dpe, octaves, sigma, k, subpixel, timeout=4 * 3600):
    """Seconds the executable's own Sift::calculate takes on this host (one thread), and the number of keypoints."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    _, buf = _call("time", struct.pack("<iiiiffi", w, h, dpe, octaves, sigma, k, int(subpixel)) + img.tobytes(), timeout)
    return struct.unpack("<di", buf)


def run_calculate(img, dpe, octaves, sigma, k, subpixel, timeout=8 * 3600):
    """Only the executable's Sift::calculate: the keypoint dict (orientation and descriptors included), or None on a vigra exception."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    rc, buf = _call("calculate", struct.pack("<iiiiffi", w, h, dpe, octaves, sigma, k, int(subpixel)) + img.tobytes(), timeout)
    if rc == 3:
        return None
    kps, off = _points(buf, 0, True)
    assert off == len(buf)
    return kps


def unit_exception(img, sigma, timeout=120):
    """None when the executable's convolveWithGauss, reduceToNextLevel and increaseToNextLevel all return on `img`, else the
    what() of the vigra exception the first failing one left with."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    rc, buf = _call("unit", struct.pack("<iif", w, h, sigma) + img.tobytes(), timeout)
    return buf[4:].decode(errors="replace") if rc == 3 else None
