"""ctypes binding of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's hot path (oracle/sift_oracle.cpp).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    """Compile the oracle if the .so is missing or stale (gcc only; no GPU needed)."""
    src = [os.path.join(_ROOT, "oracle", f) for f in ("sift_oracle.cpp", "capi.cpp", "sift_oracle.hpp", "vigra_linalg.hpp")]
    if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_ref_libs = {}
REF_DIR = os.path.join(_ROOT, "oracle", "_ref")


def ref_available(fast=True):
    """True when the reference's own sources were compiled here (oracle/_ref/, see oracle/Makefile)."""
    return os.path.exists(os.path.join(REF_DIR, "libref_fast.so" if fast else "libref.so"))


def ref_lib(fast=True):
    """oracle/_ref/libref[_fast].so: /root/reference/{sift,algorithms}.cpp compiled unmodified against the Vigra
    stand-in headers; same entry points as liboracle.so (oracle/ref_capi.cpp)."""
    if fast not in _ref_libs:
        path = os.path.join(REF_DIR, "libref_fast.so" if fast else "libref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/sift.cpp"):
            env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL, env=env)
        _ref_libs[fast] = _bind(C.CDLL(path), ref=True)
    return _ref_libs[fast]


def lib():
    global _lib
    if _lib is None:
        _lib = _bind(C.CDLL(build()), ref=False)
    return _lib


def _bind(L, ref):
    L.oracle_create.restype = C.c_void_p
    L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
    L.oracle_destroy.argtypes = [C.c_void_p]
    L.oracle_last_error.restype = C.c_char_p
    L.oracle_last_error.argtypes = [C.c_void_p]
    L.oracle_calculate.restype = C.c_int
    L.oracle_calculate.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.oracle_level_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.oracle_get_gauss.restype = C.c_float
    L.oracle_get_gauss.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.oracle_get_dog.restype = C.c_float
    L.oracle_get_dog.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.oracle_nearest_gaussian.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.oracle_n_candidates.restype = C.c_int
    L.oracle_n_candidates.argtypes = [C.c_void_p]
    L.oracle_get_candidates.argtypes = [C.c_void_p, _u16p, _u16p, _u16p, _u16p, _f32p, _u8p]
    L.oracle_n_survivors.restype = C.c_int
    L.oracle_n_survivors.argtypes = [C.c_void_p]
    L.oracle_get_survivors.argtypes = [C.c_void_p, _u16p, _u16p, _u16p, _u16p, _f32p, _u8p]
    L.oracle_n_keypoints.restype = C.c_int
    L.oracle_n_keypoints.argtypes = [C.c_void_p]
    L.oracle_get_keypoints.argtypes = [C.c_void_p, _u16p, _u16p, _u16p, _u16p, _f32p, _f32p, _u8p, _f32p, _i32p]
    L.oracle_format_results.restype = C.c_long
    L.oracle_format_results.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
    if not ref:
        L.oracle_gaussian_taps.restype = C.c_int
        L.oracle_gaussian_taps.argtypes = [C.c_float, _f32p, C.c_int]
        L.oracle_resize_map.argtypes = [C.c_int, C.c_int, _i32p]
        L.oracle_resize.restype = C.c_int
        L.oracle_resize.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int]
        L.oracle_inverse3.restype = C.c_int
        L.oracle_inverse3.argtypes = [_f32p, _f32p]
        L.oracle_linear_solve3.restype = C.c_int
        L.oracle_linear_solve3.argtypes = [_f32p, _f32p, _f32p]
    else:
        L.oracle_flavour.restype = C.c_char_p
        L.oracle_normalize.argtypes = [_f32p, C.c_int]
    L.oracle_convolve.restype = C.c_int
    L.oracle_convolve.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
    L.oracle_reduce.restype = C.c_int
    L.oracle_reduce.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
    L.oracle_increase.restype = C.c_int
    L.oracle_increase.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
    L.oracle_dog.argtypes = [_f32p, _f32p, C.c_long, _f32p]
    L.oracle_extrema.restype = C.c_long
    L.oracle_extrema.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, _u16p, _u16p, C.c_long]
    L.oracle_eliminate.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, _u16p, _u16p, C.c_long, _u8p]
    L.oracle_vertex_parabola.restype = C.c_float
    L.oracle_vertex_parabola.argtypes = [C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float]
    L.oracle_find_peaks.restype = C.c_int
    L.oracle_find_peaks.argtypes = [_f32p, _f32p]
    L.oracle_sort_order.argtypes = [_u8p, C.c_long, _u32p]
    L.oracle_gradient.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p]
    L.oracle_time_calculate.restype = C.c_double
    L.oracle_time_calculate.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    return L


class OraclePrecondition(Exception):
    pass


class Oracle:
    """One reference-equivalent Sift object (ctor order of sift.hpp:66-71)."""

    def __init__(self, dogs_per_epoch=3, octaves=3, sigma=1.6, k=float(np.float32(np.sqrt(2.0))), subpixel=False,
                 literal=False, strict=False, L=None):
        self.L = L if L is not None else lib()
        self.dpe, self.octaves = dogs_per_epoch, octaves
        self.h = self.L.oracle_create(dogs_per_epoch, octaves, sigma, k, int(subpixel), int(literal), int(strict))

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    def calculate(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        h, w = img.shape
        ow, oh = C.c_int(), C.c_int()
        n = self.L.oracle_calculate(self.h, img, w, h, C.byref(ow), C.byref(oh))
        if n == -1:
            raise OraclePrecondition(self.L.oracle_last_error(self.h).decode())
        if n < 0:
            raise RuntimeError(self.L.oracle_last_error(self.h).decode())
        self.out_w, self.out_h = ow.value, oh.value
        return self.keypoints()

    def time_calculate(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        h, w = img.shape
        n = C.c_int()
        t = self.L.oracle_time_calculate(self.h, img, w, h, C.byref(n))
        return t, n.value

    def level_dims(self, o):
        w, h = C.c_int(), C.c_int()
        self.L.oracle_level_dims(self.h, o, C.byref(w), C.byref(h))
        return w.value, h.value

    def gauss(self, o, i):
        w, h = self.level_dims(o)
        out = np.empty((h, w), np.float32)
        s = self.L.oracle_get_gauss(self.h, o, i, out.ctypes.data)
        return out, s

    def dog(self, o, i):
        w, h = self.level_dims(o)
        out = np.empty((h, w), np.float32)
        s = self.L.oracle_get_dog(self.h, o, i, out.ctypes.data)
        return out, s

    def nearest_gaussian(self, scale):
        o, i = C.c_int(), C.c_int()
        self.L.oracle_nearest_gaussian(self.h, scale, C.byref(o), C.byref(i))
        return o.value, i.value

    def _points(self, n, getter, with_desc):
        d = dict(x=np.zeros(n, np.uint16), y=np.zeros(n, np.uint16), octave=np.zeros(n, np.uint16),
                 index=np.zeros(n, np.uint16), scale=np.zeros(n, np.float32), filtered=np.zeros(n, np.uint8))
        if with_desc:
            d["orientation"] = np.zeros(n, np.float32)
            d["desc"] = np.zeros((n, 128), np.float32)
            d["desc_len"] = np.zeros(n, np.int32)
            getter(self.h, d["x"], d["y"], d["octave"], d["index"], d["scale"], d["orientation"], d["filtered"],
                   d["desc"], d["desc_len"])
        else:
            getter(self.h, d["x"], d["y"], d["octave"], d["index"], d["scale"], d["filtered"])
        return d

    def candidates(self):
        return self._points(self.L.oracle_n_candidates(self.h), self.L.oracle_get_candidates, False)

    def survivors(self):
        return self._points(self.L.oracle_n_survivors(self.h), self.L.oracle_get_survivors, False)

    def keypoints(self):
        return self._points(self.L.oracle_n_keypoints(self.h), self.L.oracle_get_keypoints, True)

    def text(self):
        n = self.L.oracle_format_results(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.L.oracle_format_results(self.h, buf, n + 1)
        return buf.value.decode()


# ---- unit helpers --------------------------------------------------------------------------
def gaussian_taps(sigma):
    buf = np.zeros(1024, np.float32)
    r = lib().oracle_gaussian_taps(sigma, buf, buf.size)
    return buf[: 2 * r + 1].copy(), r


def convolve(img, sigma, L=None):
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty_like(img)
    if (L or lib()).oracle_convolve(img, img.shape[1], img.shape[0], sigma, out) != 0:
        raise OraclePrecondition("kernel longer than line")
    return out


def resize_map(n_old, n_new):
    m = np.zeros(n_new, np.int32)
    lib().oracle_resize_map(n_old, n_new, m)
    return m


def reduce(img, sigma, L=None):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    if (L or lib()).oracle_reduce(img, w, h, sigma, out) != 0:
        raise OraclePrecondition("reduce")
    return out


def increase(img, sigma, L=None):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty((2 * h, 2 * w), np.float32)
    if (L or lib()).oracle_increase(img, w, h, sigma, out) != 0:
        raise OraclePrecondition("increase")
    return out


def dog(lower, higher, L=None):
    lower = np.ascontiguousarray(lower, np.float32)
    higher = np.ascontiguousarray(higher, np.float32)
    out = np.empty_like(lower)
    (L or lib()).oracle_dog(lower, higher, lower.size, out)
    return out


def extrema(d0, d1, d2, L=None):
    d0, d1, d2 = (np.ascontiguousarray(a, np.float32) for a in (d0, d1, d2))
    h, w = d1.shape
    cap = w * h
    xs, ys = np.zeros(cap, np.uint16), np.zeros(cap, np.uint16)
    n = (L or lib()).oracle_extrema(d0, d1, d2, w, h, xs, ys, cap)
    return xs[:n].copy(), ys[:n].copy()


def eliminate(d0, d1, d2, xs, ys, L=None):
    d0, d1, d2 = (np.ascontiguousarray(a, np.float32) for a in (d0, d1, d2))
    h, w = d1.shape
    xs, ys = np.ascontiguousarray(xs, np.uint16), np.ascontiguousarray(ys, np.uint16)
    f = np.zeros(xs.size, np.uint8)
    (L or lib()).oracle_eliminate(d0, d1, d2, w, h, xs, ys, xs.size, f)
    return f


def inverse3(a):
    a = np.ascontiguousarray(a, np.float32).reshape(9)
    out = np.zeros(9, np.float32)
    ok = lib().oracle_inverse3(a, out)
    return bool(ok), out.reshape(3, 3)


def linear_solve3(a, b):
    a = np.ascontiguousarray(a, np.float32).reshape(9)
    b = np.ascontiguousarray(b, np.float32).reshape(3)
    out = np.zeros(3, np.float32)
    ok = lib().oracle_linear_solve3(a, b, out)
    return bool(ok), out


def vertex_parabola(lx, ly, px, py, rx, ry, L=None):
    return (L or lib()).oracle_vertex_parabola(lx, ly, px, py, rx, ry)


def find_peaks(histo, L=None):
    histo = np.ascontiguousarray(histo, np.float32)
    out = np.zeros(36, np.float32)
    n = (L or lib()).oracle_find_peaks(histo, out)
    return out[:n].copy()


def sort_order(flags, L=None):
    flags = np.ascontiguousarray(flags, np.uint8)
    order = np.zeros(flags.size, np.uint32)
    (L or lib()).oracle_sort_order(flags, flags.size, order)
    return order


def gradient(img, L=None):
    img = np.ascontiguousarray(img, np.float32)
    mag, ori = np.empty_like(img), np.empty_like(img)
    (L or lib()).oracle_gradient(img, img.shape[1], img.shape[0], mag, ori)
    return mag, ori
