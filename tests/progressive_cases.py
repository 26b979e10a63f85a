"""Progressive VarDCT files (JXL_ENC_FRAME_SETTING_PROGRESSIVE_AC / QPROGRESSIVE_AC): the AC coefficients of every group
arrive in several pass sections that add up (frame.h: PassDev, vardct_sections.h: DecodeAcGroup(pass))."""
import cases
from oracle import synth

GRID = [(w, h, opt, dist, effort) for (w, h) in [(600, 400), (256, 256), (1100, 700), (97, 33)]
        for opt in ("PROGRESSIVE_AC", "QPROGRESSIVE_AC") for (dist, effort) in [(1.0, 7), (3.0, 3)]]


def name(w, h, opt, dist, effort):
    return "prog_%s_%dx%d_d%g_e%d" % (opt.lower(), w, h, dist, effort)


def make(ref, w, h, opt, dist, effort):
    img = synth.synth_image(w, h, 5)
    return cases._cached(name(w, h, opt, dist, effort), lambda: ref.encode_ex(img, w, h, 3, distance=dist, options={opt: 1, "EFFORT": effort}))
