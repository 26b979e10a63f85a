"""Progressive VarDCT files (JXL_ENC_FRAME_SETTING_PROGRESSIVE_AC / QPROGRESSIVE_AC): the AC coefficients of every group
arrive in several pass sections that add up (frame.h: PassDev, vardct_sections.h: DecodeAcGroup(pass)), and the extra
channels are split over the passes by their shifts -- a lossless alpha comes with the last pass, a squeezed (lossy) alpha
pyramid level by level."""
import cases
from oracle import synth

# (w, h, option, distance, effort, alpha_distance): alpha_distance None = no alpha channel
GRID = [(w, h, opt, dist, effort, None) for (w, h) in [(600, 400), (256, 256), (1100, 700), (97, 33)]
        for opt in ("PROGRESSIVE_AC", "QPROGRESSIVE_AC") for (dist, effort) in [(1.0, 7), (3.0, 3)]]
GRID += [(w, h, opt, 1.0, 7, ad) for (w, h) in [(600, 400), (2200, 300)] for opt in ("PROGRESSIVE_AC", "QPROGRESSIVE_AC")
         for ad in (0.0, 1.0)]
# no larger than one group: still several sections (one per pass), and the extra channels then live in the global stream
GRID += [(w, h, "QPROGRESSIVE_AC", 1.5, 5, ad) for (w, h) in [(185, 226), (9, 22), (256, 256)] for ad in (0.0, 1.0)]


def name(w, h, opt, dist, effort, ad):
    return "prog_%s_%dx%d_d%g_e%d%s" % (opt.lower(), w, h, dist, effort, "" if ad is None else "_a%g" % ad)


def make(ref, w, h, opt, dist, effort, ad):
    img = synth.synth_image(w, h, 5, alpha=ad is not None)
    if ad is None:
        return cases._cached(name(w, h, opt, dist, effort, ad), lambda: ref.encode_ex(img, w, h, 3, distance=dist, options={opt: 1, "EFFORT": effort}))
    return cases._cached(name(w, h, opt, dist, effort, ad),
                         lambda: ref.encode_ex(img, w, h, 4, distance=dist, alpha_distance=ad, options={opt: 1, "EFFORT": effort}))
