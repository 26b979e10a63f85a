"""TEST INFRASTRUCTURE — ctypes loader for oracle/_ref/libjxlref.so (the reference's own code + its prebuilt libjxl
0.12.0 / weaver, built by oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this module; the product (jxl_coder_b200) must never do so.

Every function here calls the *unmodified reference*: see oracle/ref_api.cpp for the file:line of each entry.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_lib = None


class RefUnavailable(RuntimeError):
    pass


class RefImage(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("width", C.c_uint32), ("height", C.c_uint32), ("stride", C.c_uint32),
                ("config", C.c_char * 24), ("color_space", C.c_char * 24), ("error_class", C.c_char * 80),
                ("error_msg", C.c_char * 256)]


class RefRaw(C.Structure):
    _fields_ = [("pixels", C.POINTER(C.c_uint8)), ("xsize", C.c_size_t), ("ysize", C.c_size_t), ("bit_depth", C.c_uint32),
                ("use_floats", C.c_int), ("alpha_premultiplied", C.c_int), ("orientation", C.c_int),
                ("prefer_encoding", C.c_int), ("has_alpha", C.c_int), ("intensity_target", C.c_float),
                ("color_space", C.c_int), ("white_point", C.c_int), ("primaries", C.c_int),
                ("transfer_function", C.c_int), ("gamma", C.c_double), ("icc_size", C.c_size_t)]


class RefError(Exception):
    """Mirrors the Java exception the reference raised (class name + message)."""

    def __init__(self, cls, msg):
        super().__init__(f"{cls}: {msg}")
        self.java_class = cls
        self.message = msg


def available():
    return os.path.exists(os.path.join(_REF, "libjxlref.so"))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RefUnavailable("oracle/_ref/libjxlref.so missing: run `make -C oracle ref` where /root/reference exists")
    g = C.RTLD_GLOBAL
    # bionic sonames first (by absolute path: the loader then satisfies NEEDED libc.so/libm.so/... by soname)
    for n in ("libc.so", "libm.so", "libdl.so", "liblog.so", "libbrotlicommon.so", "libbrotlidec.so",
              "libbrotlienc.so", "libjxl_cms.so", "libjxl_threads.so", "libjxl.so"):
        C.CDLL(os.path.join(_REF, "lib", n), mode=g)
    L = C.CDLL(os.path.join(_REF, "libjxlref.so"))
    u8p = C.POINTER(C.c_uint8)
    L.ref_decode_sampled.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(RefImage)]
    L.ref_get_size.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.ref_decode_oneshot.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(RefRaw)]
    L.ref_decode_oneshot_discard.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    L.ref_free.argtypes = [C.c_void_p]
    L.ref_anim_open.restype = C.c_int64
    L.ref_anim_open.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    for n in ("ref_anim_close", "ref_anim_num_frames", "ref_anim_loops", "ref_anim_width", "ref_anim_height"):
        getattr(L, n).argtypes = [C.c_int64]
    L.ref_anim_frame_duration.argtypes = [C.c_int64, C.c_int]
    L.ref_anim_get_frame.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(RefImage)]
    L.ref_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(u8p), C.POINTER(C.c_size_t)]
    L.ref_encode_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_float,
                                C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(u8p), C.POINTER(C.c_size_t)]
    L.ref_anim_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.POINTER(u8p), C.POINTER(C.c_size_t)]
    L.ref_weave_u8.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.POINTER(u8p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.ref_weave_u16.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.POINTER(C.POINTER(C.c_uint16)), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                C.POINTER(C.c_uint32)]
    L.ref_libjxl_version.restype = C.c_uint32
    _lib = L
    return L


def _buf(data):
    b = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    return (C.c_char * len(b)).from_buffer_copy(b), len(b)


_BPP = {"ARGB_8888": 4, "RGBA_F16": 8, "RGB_565": 2, "RGBA_1010102": 4}


def _take_image(img):
    """RefImage → dict(pixels=np.uint8 [h, stride], width, height, stride, config, color_space)."""
    n = img.stride * img.height
    arr = np.ctypeslib.as_array(img.data, shape=(n,)).copy().reshape(img.height, img.stride)
    lib().ref_free(img.data)
    return dict(pixels=arr, width=img.width, height=img.height, stride=img.stride, config=img.config.decode(),
                color_space=img.color_space.decode())


def set_api_level(level):
    lib().ref_set_api_level(int(level))


def decode_sampled(data, w=-1, h=-1, cfg=1, scale_mode=1, filt=4, buffer_kind=0, api_level=34):
    """JxlCoder.decodeSampled through the reference's JNI entry (JniDecoding.cpp:333-392)."""
    L = lib()
    L.ref_set_api_level(api_level)
    b, n = _buf(data)
    img = RefImage()
    rc = L.ref_decode_sampled(b, n, w, h, cfg, scale_mode, filt, buffer_kind, C.byref(img))
    if rc:
        raise RefError(img.error_class.decode(), img.error_msg.decode())
    return _take_image(img)


def get_size(data):
    b, n = _buf(data)
    w, h = C.c_uint32(), C.c_uint32()
    return (w.value, h.value) if lib().ref_get_size(b, n, C.byref(w), C.byref(h)) else None


def decode_oneshot(data, allowed_floats=True):
    """DecodeJpegXlOneShot (interop/JxlDecoding.cpp:36-176): raw libjxl RGBA u8/u16 + metadata."""
    b, n = _buf(data)
    r = RefRaw()
    rc = lib().ref_decode_oneshot(b, n, int(allowed_floats), C.byref(r))
    if rc:
        raise RefError("decode", f"DecodeJpegXlOneShot rc={rc}")
    bps = 2 if r.use_floats else 1
    # libjxl applies the orientation: transposing orientations swap the buffer dims (JniDecoding.cpp:95-100)
    xs, ys = r.xsize, r.ysize
    if r.orientation >= 5:
        xs, ys = ys, xs
    nbytes = xs * ys * 4 * bps
    raw = np.ctypeslib.as_array(r.pixels, shape=(nbytes,)).copy()
    lib().ref_free(r.pixels)
    px = raw.view(np.uint16 if bps == 2 else np.uint8).reshape(ys, xs, 4)
    meta = {k: getattr(r, k) for k, _ in RefRaw._fields_ if k != "pixels"}
    return px, meta


def decode_discard(data, allowed_floats=True):
    b, n = _buf(data)
    return lib().ref_decode_oneshot_discard(b, n, int(allowed_floats))


def _take_bytes(out, n):
    b = C.string_at(out, n.value)
    lib().ref_free(out)
    return b


def encode(pixels, w, h, colorspace=1, compression=2, data_format=1, effort=7, quality=90, decoding_speed=0,
           primaries=1, transfer=13, icc=None):
    """EncodeJxlOneshot (interop/JxlEncoding.cpp:48-193); colorspace 1 rgb / 2 rgba / 3 mono; compression 1 lossless / 2 lossy."""
    px = np.ascontiguousarray(pixels)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    iccb = bytes(icc) if icc else None
    rc = lib().ref_encode(px.ctypes.data, px.nbytes, w, h, colorspace, compression, data_format, effort, quality,
                          decoding_speed, primaries, transfer, iccb, len(iccb) if iccb else 0, C.byref(out), C.byref(n))
    if rc:
        raise RefError("encode", f"EncodeJxlOneshot rc={rc}")
    return _take_bytes(out, n)


# JxlEncoderFrameSettingId values (jxl/encode.h:132-248)
OPT = dict(EFFORT=0, DECODING_SPEED=1, RESAMPLING=2, NOISE=6, DOTS=7, PATCHES=8, EPF=9, GABORISH=10, MODULAR=11,
           KEEP_INVISIBLE=12, GROUP_ORDER=13, RESPONSIVE=16, PROGRESSIVE_AC=17, QPROGRESSIVE_AC=18, PROGRESSIVE_DC=19,
           MODULAR_GROUP_SIZE=23, MODULAR_PREDICTOR=24, MODULAR_NB_PREV_CHANNELS=26)


def encode_ex(pixels, w, h, channels, bits=8, lossless=False, distance=1.0, alpha_distance=-1.0, options=None,
              primaries=0, transfer=0, orientation=0):
    px = np.ascontiguousarray(pixels)
    options = options or {}
    ids = (C.c_int * max(1, len(options)))(*[OPT[k] if isinstance(k, str) else k for k in options])
    vals = (C.c_int * max(1, len(options)))(*list(options.values()))
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    rc = lib().ref_encode_ex(px.ctypes.data, px.nbytes, w, h, channels, bits, int(lossless), distance, alpha_distance,
                             ids, vals, len(options), primaries, transfer, orientation, C.byref(out), C.byref(n))
    if rc:
        raise RefError("encode_ex", f"rc={rc}")
    return _take_bytes(out, n)


def encode_layers(layers, w, h, channels=4, lossless=True, distance=1.0, animation=False, effort=5):
    """layers: list of dict(pixels=[h_i, w_i, channels] u8, x0, y0, mode, source, save, duration).  Blend modes as in
    jxl/codestream_header.h: 0 replace, 1 add, 2 blend, 3 muladd, 4 mul.  TEST INPUT GENERATION ONLY."""
    L = lib()
    L.ref_encode_layers.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
    px = np.concatenate([np.ascontiguousarray(l["pixels"], dtype=np.uint8).reshape(-1) for l in layers])
    dims = np.array([[l.get("x0", 0), l.get("y0", 0), l["pixels"].shape[1], l["pixels"].shape[0]] for l in layers], np.int32)
    blend = np.array([[l.get("mode", 0), l.get("source", 0), l.get("save", 0), l.get("duration", 0)] for l in layers], np.int32)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    rc = L.ref_encode_layers(px.ctypes.data, w, h, channels, int(lossless), distance, len(layers), dims.ctypes.data, blend.ctypes.data,
                             int(animation), effort, C.byref(out), C.byref(n))
    if rc:
        raise RefError("encode_layers", f"rc={rc}")
    return _take_bytes(out, n)


def anim_encode(frames, w, h, colorspace=2, compression=2, duration=40, num_loops=0, quality=90, effort=7,
                decoding_speed=0):
    fr = np.ascontiguousarray(frames)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    rc = lib().ref_anim_encode(fr.ctypes.data, w, h, colorspace, compression, fr.shape[0], duration, num_loops, quality,
                               effort, decoding_speed, C.byref(out), C.byref(n))
    if rc:
        raise RefError("anim_encode", f"rc={rc}")
    return _take_bytes(out, n)


class Anim:
    """JxlAnimatedImage through the reference's coordinator JNI (JxlAnimatedDecoderCoordinator.cpp:45-425)."""

    def __init__(self, data, cfg=1, scale_mode=1, filt=1, use_bytebuffer=False, api_level=34):
        L = lib()
        L.ref_set_api_level(api_level)
        self._keep, n = _buf(data)
        err = C.create_string_buffer(400)
        self.h = L.ref_anim_open(self._keep, n, cfg, scale_mode, filt, int(use_bytebuffer), err, 400)
        if not self.h:
            cls, _, msg = err.value.decode().partition(": ")
            raise RefError(cls, msg)

    def __len__(self):
        return lib().ref_anim_num_frames(self.h)

    def duration(self, i):
        return lib().ref_anim_frame_duration(self.h, i)

    @property
    def loops(self):
        return lib().ref_anim_loops(self.h)

    @property
    def size(self):
        return lib().ref_anim_width(self.h), lib().ref_anim_height(self.h)

    def frame(self, i, w=0, h=0):
        img = RefImage()
        if lib().ref_anim_get_frame(self.h, i, w, h, C.byref(img)):
            raise RefError(img.error_class.decode(), img.error_msg.decode())
        return _take_image(img)

    def close(self):
        if self.h:
            lib().ref_anim_close(self.h)
            self.h = 0


def weave_u8(img, nw, nh, fn=4, premul=False, mode=0):
    """weave_scale_u8 (weaver/src/scale.rs:294-326); img [h,w,4] u8; mode 0 resize / 1 fill / 2 fit."""
    a = np.ascontiguousarray(img)
    h, w, _ = a.shape
    out, ow, oh, os_ = C.POINTER(C.c_uint8)(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    if lib().ref_weave_u8(a.ctypes.data, w * 4, w, h, nw, nh, fn, int(premul), mode, C.byref(out), C.byref(ow), C.byref(oh), C.byref(os_)):
        raise RefError("weave", "null result")
    r = np.ctypeslib.as_array(out, shape=(oh.value, os_.value)).copy()[:, :ow.value * 4].reshape(oh.value, ow.value, 4)
    lib().ref_free(out)
    return r


def weave_u16(img, nw, nh, depth=16, fn=4, premul=False, mode=0):
    a = np.ascontiguousarray(img)
    h, w, _ = a.shape
    out, ow, oh, os_ = C.POINTER(C.c_uint16)(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    if lib().ref_weave_u16(a.ctypes.data, w * 8, w, h, nw, nh, depth, fn, int(premul), mode, C.byref(out), C.byref(ow), C.byref(oh), C.byref(os_)):
        raise RefError("weave", "null result")
    r = np.ctypeslib.as_array(out, shape=(oh.value, os_.value)).copy()[:, :ow.value * 4].reshape(oh.value, ow.value, 4)
    lib().ref_free(out)
    return r


def fnv1a64(buf):
    """64-bit FNV-1a — vectorised via Python int loop on bytes is slow; use numpy trick per byte."""
    h = 1469598103934665603
    for b in bytes(buf):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def dither_table():
    return np.fromfile(os.path.join(_REF, "dither_table.bin"), dtype="<f4").reshape(32, 32)
