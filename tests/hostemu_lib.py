"""TEST TOOL: builds and loads tests/hostemu (the shared host/device section decoders run serially on the CPU)."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = ["tests/hostemu/hostemu.cc", "jxl_coder_b200/csrc/frame_parser.cc", "jxl_coder_b200/csrc/plan.cc",
       "jxl_coder_b200/csrc/natural_orders.cc", "jxl_coder_b200/csrc/numeric_tables.cc", "jxl_coder_b200/csrc/color_params.cc",
       "jxl_coder_b200/csrc/resize.cc", "jxl_coder_b200/csrc/color_matrix.cc", "jxl_coder_b200/csrc/resize_host.cc",
       "jxl_coder_b200/csrc/color_matrix_host.cc"]
OUT = os.path.join(HERE, "hostemu", "_build", "libhostemu.so")
_lib = None


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [os.path.join(ROOT, s) for s in SRC] + [os.path.join(ROOT, "jxl_coder_b200/csrc", h) for h in
                                                  os.listdir(os.path.join(ROOT, "jxl_coder_b200/csrc")) if h.endswith(".h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", OUT] + SRC, cwd=ROOT)
    return OUT


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emu_decode.restype = C.c_void_p
        L.emu_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]
        for n in ("emu_free", "emu_status", "emu_failed_stream", "emu_late_status"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.emu_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        for n, t in (("emu_lf_quant", C.c_int32), ("emu_xfromy", C.c_int32), ("emu_bfromy", C.c_int32), ("emu_cell_strategy", C.c_uint8),
                     ("emu_cell_hfmul", C.c_uint16), ("emu_cell_sharp", C.c_uint8), ("emu_coef", C.c_int16), ("emu_mod", C.c_int32)):
            getattr(L, n).restype = C.POINTER(t)
            getattr(L, n).argtypes = [C.c_void_p]
        L.emu_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.emu_plane_stride.argtypes = [C.c_void_p]
        L.emu_plane_h.argtypes = [C.c_void_p]
        L.emu_logcount.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.emu_natural_order.argtypes = [C.c_uint32, C.c_void_p]
        L.emu_resize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t,
                                 C.POINTER(C.c_uint32)]
        L.emu_color_matrix.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32]
        L.emu_color_matrix16.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32]
        L.emu_expand_lehmer.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.emu_rcp_check.argtypes = [C.c_long]
        L.emu_rcp_check.restype = C.c_long
        _lib = L
    return _lib


INFO = ["width", "height", "w8", "h8", "lf_stride", "coef_stride", "coef_h", "num_groups", "num_lf_groups", "encoding",
        "num_mod_channels", "mod_stride", "w64", "h64", "single_section", "global_nb_transforms", "xsize", "ysize", "bits",
        "num_extra", "is_last", "orientation", "upsampling", "up_width", "up_height"]


class Decoded:
    def __init__(self, data, frame=0):
        L = lib()
        err = C.create_string_buffer(512)
        self.h = L.emu_decode(bytes(data), len(data), frame, err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())
        buf = (C.c_uint32 * len(INFO))()
        L.emu_info(self.h, buf)
        self.info = dict(zip(INFO, list(buf)))
        self.status = L.emu_status(self.h)
        self.failed_stream = L.emu_failed_stream(self.h)

    def _arr(self, fn, shape):
        p = getattr(lib(), fn)(self.h)
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(shape).copy()

    def lf_quant(self):
        i = self.info
        return self._arr("emu_lf_quant", (3, i["h8"], i["lf_stride"]))[:, :, :i["w8"]]

    def cells(self):
        i = self.info
        s = self._arr("emu_cell_strategy", (i["h8"], i["w8"]))
        q = self._arr("emu_cell_hfmul", (i["h8"], i["w8"]))
        sh = self._arr("emu_cell_sharp", (i["h8"], i["w8"]))
        return s, q, sh

    def cfl(self):
        i = self.info
        return self._arr("emu_xfromy", (i["h64"], i["w64"])), self._arr("emu_bfromy", (i["h64"], i["w64"]))

    def coef(self):
        i = self.info
        return self._arr("emu_coef", (3, i["coef_h"], i["coef_stride"]))[:, :, :i["w8"] * 8]

    def mod(self):
        i = self.info
        return self._arr("emu_mod", (i["num_mod_channels"], i["height"], i["mod_stride"]))[:, :, :i["width"]]

    def render(self, bits16=False, want_planes=False):
        """RGBA [h, w, 4] through the numeric host/device functions (+ XYB planes after IDCT / after filters)."""
        i = self.info
        L = lib()
        dt = np.uint16 if bits16 else np.uint8
        ow, oh = (i["up_width"], i["up_height"]) if i["upsampling"] == 2 else (i["width"], i["height"])
        out = np.zeros((oh, ow, 4), dt)
        ps, ph = L.emu_plane_stride(self.h), L.emu_plane_h(self.h)
        a = np.zeros((3, ph, ps), np.float32) if want_planes else None
        b = np.zeros((3, ph, ps), np.float32) if want_planes else None
        rc = L.emu_render(self.h, out.ctypes.data, out.strides[0], int(bits16), a.ctypes.data if want_planes else None,
                          b.ctypes.data if want_planes else None)
        if rc:
            raise RuntimeError("emu_render rc=%d" % rc)
        if want_planes:
            return out, a[:, :i["height"], :i["width"]], b[:, :i["height"], :i["width"]]
        return out

    def late_status(self):
        """Status raised by a pixel stage during render() (what the product reads back after its kernels)."""
        return lib().emu_late_status(self.h)

    def close(self):
        if self.h:
            lib().emu_free(self.h)
            self.h = None


def resize_rgba8(img, req_w, req_h, scale_mode, filt, has_alpha=False):
    """CPU restatement of the rescale plan + passes (csrc/resize.cc).  Returns the output array or the kResize* status."""
    h, w, _ = img.shape
    src = np.ascontiguousarray(img, dtype=np.uint8)
    cap = max(w * h * 4 * 64, 1 << 24)
    out = np.zeros(cap, np.uint8)
    dims = (C.c_uint32 * 2)()
    st = lib().emu_resize(src.ctypes.data, w, h, req_w, req_h, scale_mode, filt, int(has_alpha), out.ctypes.data, cap, dims)
    if st:
        return st
    return out[: dims[0] * dims[1] * 4].reshape(dims[1], dims[0], 4).copy()


def color_matrix(jxl, rgba):
    """api_level < 34 colour pass (csrc/color_matrix.cc) on a copy of rgba [h,w,4] u8.  Returns (status, array)."""
    a = np.ascontiguousarray(rgba, dtype=np.uint8).copy()
    st = lib().emu_color_matrix(bytes(jxl), len(jxl), a.ctypes.data, a.shape[1], a.shape[0])
    return st, a


def color_matrix16(jxl, rgba16):
    """The same on RGBA16 samples [h,w,4] u16 (applyColorMatrix16Bit).  Returns (status, array)."""
    a = np.ascontiguousarray(rgba16, dtype=np.uint16).copy()
    st = lib().emu_color_matrix16(bytes(jxl), len(jxl), a.ctypes.data, a.shape[1], a.shape[0])
    return st, a
