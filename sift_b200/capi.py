"""ctypes binding of libsift_gpu.so (include/sift_gpu.h) — the same calls the C++ host class
sift::Sift (include/sift/sift.hpp) makes.  Fails loudly when the CUDA library is missing: there is
no CPU fallback anywhere in the product path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIFT_GPU_LIB") or os.path.join(_HERE, "libsift_gpu.so")   # SIFT_GPU_LIB: development builds of the same ABI

OK, E_INVALID, E_CUDA, E_PRECONDITION, E_CAPACITY, E_ASSERT, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
FLAG_ORDER_CANONICAL, FLAG_STRICT, FLAG_FMA_BLUR, FLAG_KEEP_UPSAMPLED, FLAG_SERIAL = 1, 2, 4, 8, 16
DTYPE_F32, DTYPE_U8 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
KIND_GAUSS, KIND_DOG = 0, 1
SQRT2_F32 = float(np.float32(np.sqrt(2.0)))


class Params(C.Structure):
    _fields_ = [("sigma", C.c_float), ("k", C.c_float), ("octaves", C.c_uint16), ("dogs_per_epoch", C.c_uint16),
                ("subpixel", C.c_uint8), ("device", C.c_int32), ("max_width", C.c_int32), ("max_height", C.c_int32),
                ("max_batch", C.c_int32), ("flags", C.c_uint32)]


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("row_stride_bytes", C.c_int64),
                ("dtype", C.c_int32), ("memory", C.c_int32), ("upsampled_out", C.c_void_p)]


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("octave", C.c_uint16), ("index", C.c_uint16),
                ("scale", C.c_float), ("orientation", C.c_float), ("filtered", C.c_uint8), ("desc_len", C.c_uint8),
                ("reserved", C.c_uint16)]


KP_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("octave", "<u2"), ("index", "<u2"), ("scale", "<f4"),
                     ("orientation", "<f4"), ("filtered", "u1"), ("desc_len", "u1"), ("reserved", "<u2")])
assert KP_DTYPE.itemsize == C.sizeof(Keypoint) == 20


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("n", C.c_uint32), ("kps", C.POINTER(Keypoint)), ("desc", C.POINTER(C.c_float)),
                ("n_candidates", C.c_uint32), ("n_survivors", C.c_uint32), ("out_width", C.c_int32),
                ("out_height", C.c_int32)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "pyramid_ms", "extrema_ms", "eliminate_ms", "d2h_survivors_ms",
                                         "host_order_ms", "h2d_keypoints_ms", "orientation_ms", "descriptor_ms",
                                         "d2h_results_ms", "device_total_ms", "wall_ms", "span_ms")] + [
                    ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("packed_images", C.c_uint32), ("reserved", C.c_uint32)]


class SiftGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sift_gpu error {code}: {msg}")
        self.code = code


class SiftGpuPrecondition(SiftGpuError):
    """What the reference surfaces as vigra::PreconditionViolation (a std::exception)."""


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    f32p = C.POINTER(C.c_float)
    L.sift_gpu_create.restype = C.c_int
    L.sift_gpu_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    L.sift_gpu_run.restype = C.c_int
    L.sift_gpu_run.argtypes = [C.c_void_p, C.POINTER(Image), C.c_int, C.POINTER(Result)]
    L.sift_gpu_destroy.argtypes = [C.c_void_p]
    L.sift_gpu_last_error.restype = C.c_char_p
    L.sift_gpu_last_error.argtypes = [C.c_void_p]
    L.sift_gpu_get_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
    L.sift_gpu_version.restype = C.c_char_p
    L.sift_gpu_debug_get_level.restype = C.c_int
    L.sift_gpu_debug_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int),
                                           C.POINTER(C.c_int), f32p]
    for name in ("sift_gpu_debug_blur", "sift_gpu_debug_reduce", "sift_gpu_debug_increase"):
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    L.sift_gpu_debug_extrema.restype = C.c_int
    L.sift_gpu_debug_extrema.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.sift_gpu_debug_eliminate.restype = C.c_int
    L.sift_gpu_debug_eliminate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_uint32, C.c_void_p]
    L.sift_gpu_debug_get_candidates.restype = C.c_int
    L.sift_gpu_debug_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.sift_gpu_debug_sort_order.restype = C.c_int
    L.sift_gpu_debug_sort_order.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.sift_gpu_debug_sort_order_fast.restype = C.c_int
    L.sift_gpu_debug_sort_order_fast.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    L.sift_gpu_debug_host_replay.restype = C.c_int
    L.sift_gpu_debug_host_replay.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                             C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.sift_gpu_debug_pack_rows_u8.restype = C.c_int
    L.sift_gpu_debug_pack_rows_u8.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "sift_gpu_create", "sift_gpu_run", "sift_gpu_destroy", "sift_gpu_last_error", "sift_gpu_get_timings",
    "sift_gpu_version", "sift_gpu_debug_get_level", "sift_gpu_debug_blur", "sift_gpu_debug_reduce",
    "sift_gpu_debug_increase", "sift_gpu_debug_extrema", "sift_gpu_debug_eliminate", "sift_gpu_debug_get_candidates",
    "sift_gpu_debug_sort_order", "sift_gpu_debug_sort_order_fast", "sift_gpu_debug_host_replay", "sift_gpu_debug_pack_rows_u8",
]


def sort_order(flags):
    """The std::sort(cmpByFilter) permutation the host replays (reference sift.cpp:37)."""
    flags = np.ascontiguousarray(flags, np.uint8)
    order = np.zeros(flags.size, np.uint32)
    rc = load().sift_gpu_debug_sort_order(flags.ctypes.data, flags.size, order.ctypes.data)
    if rc != 0:
        raise SiftGpuError(rc, "sort_order")
    return order


def sort_order_fast(flags):
    """Post-sort order of the unfiltered elements from the sparse introsort simulation (order_replay.h)."""
    flags = np.ascontiguousarray(flags, np.uint8)
    order = np.zeros(max(1, flags.size), np.uint32)
    n = C.c_uint32()
    rc = load().sift_gpu_debug_sort_order_fast(flags.ctypes.data, flags.size, order.ctypes.data, C.byref(n))
    if rc != 0:
        raise SiftGpuError(rc, "sort_order_fast")
    return order[: n.value].copy()


def host_replay(width, height, n_candidates, canon, xs, ys, octave, index, *, dogs_per_epoch=3, octaves=3, sigma=1.6,
                k=SQRT2_F32, subpixel=False, flags=0):
    """The host half of Sift::calculate between the two device stages (sift.cpp:37-55), run without a device: unfiltered
    candidates of one image in canonical order in, keypoint records in the reference's final vector order out.
    Returns (kps as a KP_DTYPE array, n_survivors after the first cleanup)."""
    canon = np.ascontiguousarray(canon, np.uint32)
    xs, ys = np.ascontiguousarray(xs, np.uint16), np.ascontiguousarray(ys, np.uint16)
    octave, index = np.ascontiguousarray(octave, np.uint8), np.ascontiguousarray(index, np.uint8)
    prm = Params(sigma, k, octaves, dogs_per_epoch, 1 if subpixel else 0, 0, width, height, 1, flags)
    kps = np.zeros(max(1, canon.size), KP_DTYPE)
    n, ns = C.c_uint32(), C.c_uint32()
    rc = load().sift_gpu_debug_host_replay(C.byref(prm), width, height, n_candidates, canon.ctypes.data, xs.ctypes.data,
                                           ys.ctypes.data, octave.ctypes.data, index.ctypes.data, canon.size,
                                           kps.ctypes.data, kps.size, C.byref(n), C.byref(ns))
    if rc == E_PRECONDITION:
        raise SiftGpuPrecondition(rc, "host_replay")
    if rc != 0:
        raise SiftGpuError(rc, "host_replay")
    return kps[: n.value].copy(), ns.value


def pack_u8(img, dst_pitch=None):
    """The lossless f32 -> u8 frame packing of the upload path (pack_host.cpp).  Returns the bytes, or None when some pixel
    is not an exact 8-bit value."""
    img = np.asarray(img, np.float32)
    assert img.ndim == 2 and img.strides[1] == 4
    h, w = img.shape
    pitch = dst_pitch or w
    out = np.zeros((h, pitch), np.uint8)
    ok = load().sift_gpu_debug_pack_rows_u8(img.ctypes.data, img.strides[0], w, h, out.ctypes.data, pitch)
    return out[:, :w] if ok else None


def results_to_dicts(res):
    """Copies a ctypes Result array into plain numpy (the library's buffers are only valid until the next run)."""
    out = []
    for r in res:
        if r.n:
            kps = np.ctypeslib.as_array(C.cast(r.kps, C.POINTER(C.c_uint8)), (r.n * KP_DTYPE.itemsize,)).view(KP_DTYPE).copy()
            desc = np.ctypeslib.as_array(r.desc, (r.n, 128)).copy()
        else:
            kps, desc = np.zeros(0, KP_DTYPE), np.zeros((0, 128), np.float32)
        out.append(dict(status=r.status, kps=kps, desc=desc, n_candidates=r.n_candidates, n_survivors=r.n_survivors,
                        out_width=r.out_width, out_height=r.out_height))
    return out


class SiftGpu:
    """One device context; mirrors the constructor of the reference's sift::Sift (sift.hpp:66-71)."""

    def __init__(self, dogs_per_epoch=3, octaves=3, sigma=1.6, k=SQRT2_F32, subpixel=False, *, max_width, max_height,
                 max_batch=1, device=0, flags=0):
        self.L = load()
        self.prm = Params(sigma, k, octaves, dogs_per_epoch, int(subpixel), device, max_width, max_height, max_batch, flags)
        h = C.c_void_p()
        rc = self.L.sift_gpu_create(C.byref(self.prm), C.byref(h))
        if rc != 0:
            raise SiftGpuError(rc, self.L.sift_gpu_last_error(None).decode())
        self.h = h
        self.octaves, self.dpe, self.subpixel = octaves, dogs_per_epoch, subpixel

    def close(self):
        if getattr(self, "h", None):
            self.L.sift_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        msg = self.L.sift_gpu_last_error(self.h).decode()
        if rc == E_PRECONDITION:
            return SiftGpuPrecondition(rc, msg)
        return SiftGpuError(rc, msg)

    def run_raw(self, descs, n):
        """descs: ctypes array of Image.  Returns the ctypes Result array (buffers valid until the next run)."""
        res = (Result * n)()
        rc = self.L.sift_gpu_run(self.h, descs, n, res)
        return rc, res

    def run(self, images, raise_on_error=True):
        """images: list of 2-D float32/uint8 numpy arrays (host) — or (ptr, w, h, dtype) tuples for device memory.
        Returns a list of dicts: kps (structured array copy), desc (n x 128 copy), n_candidates, n_survivors, status."""
        n = len(images)
        descs = (Image * n)()
        keep = []
        for i, im in enumerate(images):
            if isinstance(im, tuple):
                ptr, w, h, dt = im
                descs[i] = Image(ptr, w, h, 0, dt, MEM_DEVICE, None)
                continue
            if im.dtype == np.uint8:
                a, dt = np.ascontiguousarray(im), DTYPE_U8
            else:
                a, dt = np.ascontiguousarray(im, np.float32), DTYPE_F32
            keep.append(a)
            descs[i] = Image(a.ctypes.data, a.shape[1], a.shape[0], 0, dt, MEM_HOST, None)
        rc, res = self.run_raw(descs, n)
        if rc != 0 and raise_on_error:
            raise self._err(rc)
        return results_to_dicts(res)

    def timings(self):
        t = Timings()
        self.L.sift_gpu_get_timings(self.h, C.byref(t))
        return {n: getattr(t, n) for n, _ in Timings._fields_}

    # ---- stage-level entry points ----
    def level(self, image_idx, octave, elem, kind=KIND_GAUSS):
        w, h, s = C.c_int(), C.c_int(), C.c_float()
        rc = self.L.sift_gpu_debug_get_level(self.h, image_idx, octave, elem, kind, None, C.byref(w), C.byref(h), C.byref(s))
        if rc != 0:
            raise self._err(rc)
        out = np.empty((h.value, w.value), np.float32)
        rc = self.L.sift_gpu_debug_get_level(self.h, image_idx, octave, elem, kind, out.ctypes.data, None, None, None)
        if rc != 0:
            raise self._err(rc)
        return out, s.value

    def _blur_like(self, fn, img, sigma, shape):
        img = np.ascontiguousarray(img, np.float32)
        out = np.empty(shape, np.float32)
        rc = fn(self.h, img.ctypes.data, img.shape[1], img.shape[0], sigma, out.ctypes.data)
        if rc != 0:
            raise self._err(rc)
        return out

    def blur(self, img, sigma):
        return self._blur_like(self.L.sift_gpu_debug_blur, img, sigma, img.shape)

    def reduce(self, img, sigma):
        return self._blur_like(self.L.sift_gpu_debug_reduce, img, sigma, ((img.shape[0] + 1) // 2, (img.shape[1] + 1) // 2))

    def increase(self, img, sigma):
        return self._blur_like(self.L.sift_gpu_debug_increase, img, sigma, (img.shape[0] * 2, img.shape[1] * 2))

    def extrema(self, d0, d1, d2):
        d0, d1, d2 = (np.ascontiguousarray(a, np.float32) for a in (d0, d1, d2))
        h, w = d1.shape
        cap = w * h
        xs, ys, n = np.zeros(cap, np.uint16), np.zeros(cap, np.uint16), C.c_uint32()
        rc = self.L.sift_gpu_debug_extrema(self.h, d0.ctypes.data, d1.ctypes.data, d2.ctypes.data, w, h, xs.ctypes.data,
                                           ys.ctypes.data, cap, C.byref(n))
        if rc != 0:
            raise self._err(rc)
        return xs[: n.value].copy(), ys[: n.value].copy()

    def eliminate(self, d0, d1, d2, xs, ys):
        d0, d1, d2 = (np.ascontiguousarray(a, np.float32) for a in (d0, d1, d2))
        xs, ys = np.ascontiguousarray(xs, np.uint16), np.ascontiguousarray(ys, np.uint16)
        h, w = d1.shape
        f = np.zeros(xs.size, np.uint8)
        rc = self.L.sift_gpu_debug_eliminate(self.h, d0.ctypes.data, d1.ctypes.data, d2.ctypes.data, w, h, xs.ctypes.data,
                                             ys.ctypes.data, xs.size, f.ctypes.data)
        if rc != 0:
            raise self._err(rc)
        return f

    def candidates(self, image_idx):
        n = C.c_uint32()
        rc = self.L.sift_gpu_debug_get_candidates(self.h, image_idx, None, None, None, None, None, 0, C.byref(n))
        if rc != 0:
            raise self._err(rc)
        m = n.value
        d = dict(x=np.zeros(m, np.uint16), y=np.zeros(m, np.uint16), octave=np.zeros(m, np.uint16),
                 index=np.zeros(m, np.uint16), filtered=np.zeros(m, np.uint8))
        rc = self.L.sift_gpu_debug_get_candidates(self.h, image_idx, d["x"].ctypes.data, d["y"].ctypes.data,
                                                  d["octave"].ctypes.data, d["index"].ctypes.data, d["filtered"].ctypes.data,
                                                  m, C.byref(n))
        if rc != 0:
            raise self._err(rc)
        return d


# ---- nvJPEG front end (include/sift_gpu_jpeg.h, libsift_gpu_jpeg.so) --------------------------------------------
class Jpeg(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t)]


_jpeg_lib = None


def load_jpeg():
    global _jpeg_lib
    if _jpeg_lib is None:
        load()  # libsift_gpu.so first: the front end links against it
        L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsift_gpu_jpeg.so"))
        L.sift_gpu_jpeg_create.restype = C.c_int
        L.sift_gpu_jpeg_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.sift_gpu_jpeg_run.restype = C.c_int
        L.sift_gpu_jpeg_run.argtypes = [C.c_void_p, C.POINTER(Jpeg), C.c_int, C.POINTER(Result)]
        L.sift_gpu_jpeg_decoded.restype = C.c_int
        L.sift_gpu_jpeg_decoded.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.sift_gpu_jpeg_last_error.restype = C.c_char_p
        L.sift_gpu_jpeg_last_error.argtypes = [C.c_void_p]
        L.sift_gpu_jpeg_destroy.restype = None
        L.sift_gpu_jpeg_destroy.argtypes = [C.c_void_p]
        _jpeg_lib = L
    return _jpeg_lib


JPEG_EXPORTED_SYMBOLS = ["sift_gpu_jpeg_create", "sift_gpu_jpeg_run", "sift_gpu_jpeg_decoded", "sift_gpu_jpeg_last_error",
                         "sift_gpu_jpeg_destroy"]


class SiftGpuJpeg:
    """JPEG byte streams in, keypoints out: nvJPEG decode on the device, band 0 into the context `sift`."""

    def __init__(self, sift, max_images):
        self.L = load_jpeg()
        self.sift = sift
        h = C.c_void_p()
        rc = self.L.sift_gpu_jpeg_create(sift.h, sift.prm.device, sift.prm.max_width, sift.prm.max_height, max_images, C.byref(h))
        if rc != 0:
            raise SiftGpuError(rc, self.L.sift_gpu_jpeg_last_error(None).decode())
        self.h = h

    def run(self, blobs):
        n = len(blobs)
        keep = [np.frombuffer(b, np.uint8) for b in blobs]
        arr = (Jpeg * n)()
        for i, a in enumerate(keep):
            arr[i] = Jpeg(a.ctypes.data, a.size)
        res = (Result * n)()
        rc = self.L.sift_gpu_jpeg_run(self.h, arr, n, res)
        if rc != 0:
            raise SiftGpuError(rc, self.L.sift_gpu_jpeg_last_error(self.h).decode())
        return results_to_dicts(res)

    def decoded(self, i):
        """Band 0 of image i of the last run, exactly the pixels the pipeline saw."""
        w, h = C.c_int(), C.c_int()
        rc = self.L.sift_gpu_jpeg_decoded(self.h, i, None, C.byref(w), C.byref(h))
        if rc != 0:
            raise SiftGpuError(rc, self.L.sift_gpu_jpeg_last_error(self.h).decode())
        out = np.empty((h.value, w.value), np.uint8)
        rc = self.L.sift_gpu_jpeg_decoded(self.h, i, out.ctypes.data, None, None)
        if rc != 0:
            raise SiftGpuError(rc, self.L.sift_gpu_jpeg_last_error(self.h).decode())
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.sift_gpu_jpeg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

