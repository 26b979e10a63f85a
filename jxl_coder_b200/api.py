"""ctypes binding of libjxlb200.so and the Python mirror of the reference's Kotlin API."""
import ctypes as C
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path():
    return os.path.join(_HERE, "libjxlb200.so")


class PreferredColorConfig(enum.IntEnum):   # PreferredColorConfig.kt
    DEFAULT = 1
    RGBA_8888 = 2
    RGBA_F16 = 3
    RGB_565 = 4
    RGBA_1010102 = 5
    HARDWARE = 6


class ScaleMode(enum.IntEnum):              # ScaleMode.kt
    FIT = 1
    FILL = 2
    RESIZE = 3


class JxlResizeFilter(enum.IntEnum):        # JxlResizeFilter.kt
    BILINEAR = 1
    NEAREST = 2
    CUBIC = 3
    MITCHELL_NETRAVALI = 4
    LANCZOS = 5
    CATMULL_ROM = 6
    HERMITE = 7
    BSPLINE = 8
    HANN = 9
    BICUBIC = 10


class JxlCoderError(Exception):
    """java.lang.Exception of the reference; .status is the jxlb_status code."""

    def __init__(self, status, message=""):
        super().__init__(f"[{status}] {message}")
        self.status = status
        self.message = message


class InvalidJXLException(JxlCoderError):
    pass


class InvalidImageSizeException(JxlCoderError):
    pass


class UnsupportedJXLException(JxlCoderError):
    pass


STATUS = dict(OK=0, INVALID_JXL=1, INVALID_SIZE=2, OOM=3, BAD_ARG=4, ERROR=5, UNSUPPORTED=6, NO_DEVICE=7, NOT_JXL=8)
_FORMAT_NAMES = {0: "ARGB_8888", 1: "RGBA_F16", 2: "RGB_565", 3: "RGBA_1010102"}
_FORMAT_BPP = {0: 4, 1: 8, 2: 2, 3: 4}
_CS_NAMES = {0: "", 1: "SRGB", 2: "BT2020_PQ", 3: "BT2020_HLG", 4: "DISPLAY_P3", 5: "LINEAR_SRGB", 6: "DCI_P3", 7: "BT709"}


class _Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("stride_bytes", C.c_uint32),
                ("format", C.c_int32), ("color_space", C.c_int32), ("premultiplied", C.c_int32), ("device", C.c_int32),
                ("message", C.c_char * 128)]


class _Request(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("width", C.c_int32), ("height", C.c_int32),
                ("color_config", C.c_int32), ("scale_mode", C.c_int32), ("filter", C.c_int32)]


class _BatchOpts(C.Structure):
    _fields_ = [("api_level", C.c_int32), ("output_device", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32)]


def load_library():
    """Loads libjxlb200.so; raises if it has not been built (python -m jxl_coder_b200.build / __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: build it with `python jxl_coder_b200/build.py` (needs nvcc); there is no CPU fallback")
    L = C.CDLL(p)
    L.jxlb_decode_sampled.argtypes = [C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_Image)]
    L.jxlb_decode_batch.argtypes = [C.POINTER(_Request), C.c_size_t, C.POINTER(_Image), C.POINTER(C.c_int32), C.POINTER(_BatchOpts)]
    L.jxlb_decode_batch_submit.restype = C.c_void_p
    L.jxlb_decode_batch_submit.argtypes = [C.POINTER(_Request), C.c_size_t, C.POINTER(_BatchOpts)]
    L.jxlb_decode_batch_collect.argtypes = [C.c_void_p, C.POINTER(_Image), C.POINTER(C.c_int32)]
    L.jxlb_get_size.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.jxlb_image_free.argtypes = [C.POINTER(_Image)]
    L.jxlb_anim_open.restype = C.c_void_p
    L.jxlb_anim_open.argtypes = [C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    for n in ("jxlb_anim_num_frames", "jxlb_anim_loops", "jxlb_anim_width", "jxlb_anim_height", "jxlb_anim_close"):
        getattr(L, n).argtypes = [C.c_void_p]
    L.jxlb_anim_close.restype = None
    L.jxlb_anim_frame_duration_ms.argtypes = [C.c_void_p, C.c_int32]
    L.jxlb_anim_get_frame.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_Image)]
    L.jxlb_batch_prepare.restype = C.c_void_p
    L.jxlb_batch_prepare.argtypes = [C.POINTER(_Request), C.c_size_t, C.POINTER(_BatchOpts), C.POINTER(C.c_int32)]
    L.jxlb_batch_run.argtypes = [C.c_void_p]
    L.jxlb_batch_run_async.argtypes = [C.c_void_p]
    L.jxlb_batch_wait.argtypes = [C.c_void_p]
    L.jxlb_batch_reset_stats.argtypes = [C.c_void_p]
    L.jxlb_batch_span_ms.restype = C.c_float
    L.jxlb_batch_span_ms.argtypes = [C.c_void_p, C.c_void_p]
    L.jxlb_batch_stage_ms_mean.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    L.jxlb_batch_fetch.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_Image)]
    L.jxlb_batch_device_pixels.restype = C.c_void_p
    L.jxlb_batch_device_pixels.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.jxlb_batch_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.jxlb_batch_free.argtypes = [C.c_void_p]
    L.jxlb_kernel_launches.restype = C.c_uint64
    L.jxlb_last_batch_timings.argtypes = [C.POINTER(C.c_float)]
    L.jxlb_version.restype = C.c_char_p
    _LIB = L
    return L


def _raise(status, message):
    cls = {1: InvalidJXLException, 2: InvalidImageSizeException, 6: UnsupportedJXLException}.get(status, JxlCoderError)
    raise cls(status, message)


class Bitmap:
    """What android.graphics.Bitmap carries: pixels + config + colour space (host copy as a numpy array)."""

    def __init__(self, img, keep_native=False):
        self.width, self.height, self.stride = img.width, img.height, img.stride_bytes
        self.config = _FORMAT_NAMES[img.format]
        self.color_space = _CS_NAMES.get(img.color_space, "")
        self.premultiplied = bool(img.premultiplied)
        self.device = img.device
        self._img = img if keep_native else None
        self.device_ptr = img.data if img.device >= 0 else None
        if img.device < 0:
            n = img.stride_bytes * img.height
            raw = np.ctypeslib.as_array(C.cast(img.data, C.POINTER(C.c_uint8)), shape=(n,))
            if not keep_native:  # own copy; the native buffer goes back to the pinned pool
                raw = raw.copy()
                load_library().jxlb_image_free(C.byref(img))
            # keep_native: zero-copy view of the pinned result, valid until free()
            self.pixels = raw.reshape(img.height, img.stride_bytes)  # uint8 [h, stride]; see as_array()
        else:
            self.pixels = None

    def as_array(self):
        """Typed view: ARGB_8888 -> uint8 [h,w,4]; RGBA_F16 -> float16 [h,w,4]; RGB_565 -> uint16 [h,w]; 1010102 -> uint32 [h,w]."""
        if self.config == "ARGB_8888":
            return self.pixels.reshape(self.height, self.width, 4)
        if self.config == "RGBA_F16":
            return self.pixels.view(np.float16).reshape(self.height, self.width, 4)
        if self.config == "RGB_565":
            return self.pixels.view(np.uint16).reshape(self.height, self.width)
        return self.pixels.view(np.uint32).reshape(self.height, self.width)

    def free(self):
        if self._img is not None:
            load_library().jxlb_image_free(C.byref(self._img))
            self._img = None


def _as_buffer(data):
    if isinstance(data, bytes):  # immutable: hand the library the object's own storage (it copies what it keeps)
        return C.c_char_p(data), len(data)
    if isinstance(data, bytearray):
        return (C.c_char * len(data)).from_buffer_copy(data), len(data)
    b = bytes(data)
    return (C.c_char * len(b)).from_buffer_copy(b), len(b)


def decode_batch(datas, width=-1, height=-1, config=PreferredColorConfig.DEFAULT, scale_mode=ScaleMode.FIT,
                 filt=JxlResizeFilter.MITCHELL_NETRAVALI, api_level=34, device=-1, output_device=-1, raise_on_error=True,
                 keep_native=False):
    """n independent decodeSampled calls decoded together on one GPU (jxlb_decode_batch).  Returns a list of Bitmap
    (or JxlCoderError instances when raise_on_error is False)."""
    L = load_library()
    n = len(datas)
    bufs = [_as_buffer(d) for d in datas]
    reqs = (_Request * n)()
    for i, (b, ln) in enumerate(bufs):
        reqs[i] = _Request(C.cast(b, C.c_void_p), ln, width, height, int(config), int(scale_mode), int(filt))
    outs = (_Image * n)()
    st = (C.c_int32 * n)()
    opts = _BatchOpts(api_level, output_device, device, 0)
    L.jxlb_decode_batch(reqs, n, outs, st, C.byref(opts))
    res = []
    for i in range(n):
        if st[i] != 0:
            if raise_on_error:
                for j in range(n):
                    if st[j] == 0:
                        L.jxlb_image_free(C.byref(outs[j]))
                _raise(st[i], outs[i].message.decode(errors="replace"))
            cls = {1: InvalidJXLException, 2: InvalidImageSizeException, 6: UnsupportedJXLException}.get(st[i], JxlCoderError)
            res.append(cls(st[i], outs[i].message.decode(errors="replace")))
        else:
            res.append(Bitmap(outs[i], keep_native=keep_native))
    return res


class PendingBatch:
    """jxlb_decode_batch_submit / _collect: submit() returns at once (the inputs have been copied), result() waits and
    returns the list of Bitmap.  Keep two or three in flight to overlap the batches on the GPU."""

    def __init__(self, datas, width=-1, height=-1, config=PreferredColorConfig.DEFAULT, scale_mode=ScaleMode.FIT,
                 filt=JxlResizeFilter.MITCHELL_NETRAVALI, api_level=34, device=-1, output_device=-1, keep_native=False):
        L = load_library()
        self.n = len(datas)
        self.keep_native = keep_native
        bufs = [_as_buffer(d) for d in datas]
        reqs = (_Request * self.n)()
        for i, (b, ln) in enumerate(bufs):
            reqs[i] = _Request(C.cast(b, C.c_void_p), ln, width, height, int(config), int(scale_mode), int(filt))
        opts = _BatchOpts(api_level, output_device, device, 0)
        self.h = L.jxlb_decode_batch_submit(reqs, self.n, C.byref(opts))
        if not self.h:
            raise JxlCoderError(5, "jxlb_decode_batch_submit failed")

    def result(self, raise_on_error=True):
        L = load_library()
        outs = (_Image * self.n)()
        st = (C.c_int32 * self.n)()
        L.jxlb_decode_batch_collect(self.h, outs, st)
        self.h = None
        res = []
        for i in range(self.n):
            if st[i] != 0:
                if raise_on_error:
                    for j in range(self.n):
                        if st[j] == 0:
                            L.jxlb_image_free(C.byref(outs[j]))
                    _raise(st[i], outs[i].message.decode(errors="replace"))
                res.append(JxlCoderError(st[i], outs[i].message.decode(errors="replace")))
            else:
                res.append(Bitmap(outs[i], keep_native=self.keep_native))
        return res


class JxlCoder:
    """Mirror of com.awxkee.jxlcoder.JxlCoder (decode side)."""

    api_level = 34

    @staticmethod
    def decode(byte_array, preferred_color_config=PreferredColorConfig.DEFAULT, scale_mode=ScaleMode.FIT):
        # JxlCoder.kt:50-63: w = h = -1, CATMULL_ROM
        return JxlCoder.decode_sampled(byte_array, -1, -1, preferred_color_config, scale_mode, JxlResizeFilter.CATMULL_ROM)

    @staticmethod
    def decode_sampled(byte_array, width, height, preferred_color_config=PreferredColorConfig.DEFAULT, scale_mode=ScaleMode.FIT,
                       jxl_resize_filter=JxlResizeFilter.MITCHELL_NETRAVALI):
        L = load_library()
        b, n = _as_buffer(byte_array)
        img = _Image()
        st = L.jxlb_decode_sampled(b, n, width, height, int(preferred_color_config), int(scale_mode), int(jxl_resize_filter),
                                   JxlCoder.api_level, C.byref(img))
        if st != 0:
            _raise(st, img.message.decode(errors="replace"))
        return Bitmap(img)

    @staticmethod
    def get_size(byte_array):
        """(width, height) or None (JxlCoder.kt:191-193)."""
        L = load_library()
        b, n = _as_buffer(byte_array)
        w, h = C.c_uint32(), C.c_uint32()
        return (w.value, h.value) if L.jxlb_get_size(b, n, C.byref(w), C.byref(h)) == 0 else None

    @staticmethod
    def is_jxl(byte_array):
        """JxlCoder.kt:244-258: signature test."""
        d = bytes(byte_array[:12])
        return d[:2] == b"\xff\x0a" or d == bytes([0, 0, 0, 0xC, 0x4A, 0x58, 0x4C, 0x20, 0xD, 0xA, 0x87, 0xA])


class JxlAnimatedImage:
    """Mirror of com.awxkee.jxlcoder.JxlAnimatedImage (JxlAnimatedImage.kt:43-193)."""

    def __init__(self, byte_array, preferred_color_config=PreferredColorConfig.DEFAULT, scale_mode=ScaleMode.FIT,
                 jxl_resize_filter=JxlResizeFilter.BILINEAR, api_level=34):
        L = load_library()
        self._buf, n = _as_buffer(byte_array)
        st = C.c_int32()
        self._h = L.jxlb_anim_open(self._buf, n, int(preferred_color_config), int(scale_mode), int(jxl_resize_filter), api_level, C.byref(st))
        if not self._h:
            _raise(st.value, "cannot open animated image")

    @property
    def number_of_frames(self):
        return load_library().jxlb_anim_num_frames(self._h)

    def get_frame_duration(self, frame):
        return load_library().jxlb_anim_frame_duration_ms(self._h, frame)

    @property
    def loops_count(self):
        return load_library().jxlb_anim_loops(self._h)

    def get_width(self):
        return load_library().jxlb_anim_width(self._h)

    def get_height(self):
        return load_library().jxlb_anim_height(self._h)

    def get_frame(self, frame, scale_width=0, scale_height=0):
        img = _Image()
        st = load_library().jxlb_anim_get_frame(self._h, frame, scale_width, scale_height, C.byref(img))
        if st != 0:
            _raise(st, img.message.decode(errors="replace"))
        return Bitmap(img)

    def close(self):
        if self._h:
            load_library().jxlb_anim_close(self._h)
            self._h = None


def kernel_launches():
    return int(load_library().jxlb_kernel_launches())


def last_batch_timings():
    """Device time (ms) of the last batch: upload, entropy, reconstruction, filters+colour+pack, download, total."""
    buf = (C.c_float * 6)()
    load_library().jxlb_last_batch_timings(buf)
    return dict(zip(("upload", "entropy", "recon", "filters_color_pack", "download", "total"), [float(v) for v in buf]))


class PreparedBatch:
    """Throughput interface: parse + upload once (inputs resident in HBM), run the kernels repeatedly, results stay in HBM."""

    STAGES = ("upload", "lf_sections", "group_sections", "reconstruction_phase", "inverse_transforms", "filters_color_pack", "download", "all_kernels")

    def __init__(self, datas, width=-1, height=-1, config=PreferredColorConfig.RGBA_8888, scale_mode=ScaleMode.FIT,
                 filt=JxlResizeFilter.MITCHELL_NETRAVALI, api_level=34, device=-1):
        L = load_library()
        n = len(datas)
        self._bufs = [_as_buffer(d) for d in datas]
        reqs = (_Request * n)()
        for i, (b, ln) in enumerate(self._bufs):
            reqs[i] = _Request(C.cast(b, C.c_void_p), ln, width, height, int(config), int(scale_mode), int(filt))
        st = (C.c_int32 * n)()
        opts = _BatchOpts(api_level, -1, device, 0)
        self._h = L.jxlb_batch_prepare(reqs, n, C.byref(opts), st)
        self.status = list(st)
        self.n = n
        if not self._h:
            raise JxlCoderError(5, "jxlb_batch_prepare failed")

    def run(self):
        return load_library().jxlb_batch_run(self._h)

    def fetch(self, i):
        img = _Image()
        st = load_library().jxlb_batch_fetch(self._h, i, C.byref(img))
        if st != 0:
            _raise(st, img.message.decode(errors="replace"))
        return Bitmap(img)

    def device_pixels(self, i):
        n = C.c_size_t()
        p = load_library().jxlb_batch_device_pixels(self._h, i, C.byref(n))
        return p, n.value

    def run_async(self):
        """Enqueues one run on this batch's own stream; runs of different PreparedBatch objects overlap on the GPU."""
        return load_library().jxlb_batch_run_async(self._h)

    def wait(self):
        return load_library().jxlb_batch_wait(self._h)

    def stage_ms(self):
        buf = (C.c_float * 8)()
        load_library().jxlb_batch_stage_ms(self._h, buf)
        return dict(zip(self.STAGES, [float(v) for v in buf]))

    def reset_stats(self):
        load_library().jxlb_batch_reset_stats(self._h)

    def span_ms(self, last=None):
        """Device time from this batch's first run since reset_stats() to the end of `last`'s (default: this batch's) latest run."""
        return float(load_library().jxlb_batch_span_ms(self._h, (last or self)._h))

    def stage_ms_mean(self):
        """(mean stage times over every run waited for so far, number of runs)"""
        buf = (C.c_float * 8)()
        n = C.c_int32(0)
        load_library().jxlb_batch_stage_ms_mean(self._h, buf, C.byref(n))
        return dict(zip(self.STAGES, [float(v) for v in buf])), int(n.value)

    def free(self):
        if self._h:
            load_library().jxlb_batch_free(self._h)
            self._h = None
