import glob
import os
import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["jxl"] = d["jxl"].tobytes()
    return d


def lossy_close(mine, ref, what="", min_exact=0.985):
    """Tolerance of the lossy (XYB / float) path, stated by BASELINE.json: within 1 LSB per channel -- NO sample further
    than 1 from the reference's, and (a regression guard, far tighter than the bound) at least 98.5 % exactly equal (lowest seen: 98.9 % at distance 4 / effort 5, where three filter iterations run; 99.98 % at distance 1).
    Measured with libjxl's own approximations restated (FastLinearToSRGB, RCPPS in the edge-preserving filter and the
    quant bias, its sRGB primaries): 99.98 % exact on 8-bit sRGB / P3 / Rec.709 output."""
    d = np.abs(mine.astype(np.int32) - ref.astype(np.int32))
    exact = float((d == 0).mean())
    assert d.max() <= 1, (what, int(d.max()), float((d > 1).mean()))
    assert exact >= min_exact, (what, exact)
    return exact, int(d.max())


def pq_close(mine, ref, what=""):
    """Rec.2100 PQ output.  The PQ curve rises by ~10^6 code values per unit of linear light at black, and a channel of a
    saturated colour is the difference of two terms of magnitude ~2 there, so the last-bit differences between two correct
    f32 pipelines (libjxl's own SIMD targets included) reach the 8-bit code on isolated near-black samples.  Bound: at
    most 3 samples in 10^5 further than 1 LSB (measured: 1 in 10^6 .. 1 in 10^5), those only where the reference's value is
    below 64 (a quarter of the range), >= 99.5 % exact."""
    d = np.abs(mine.astype(np.int32) - ref.astype(np.int32))
    far = d > 1
    assert float(far.mean()) <= 3e-5, (what, float(far.mean()), int(d.max()))
    if far.any():
        assert int(ref[far].max()) < 64, (what, int(ref[far].max()))
    exact = float((d == 0).mean())
    assert exact >= 0.995, (what, exact)
    return exact, int(d.max())


def restatement_close(mine, ref, what=""):
    """Bound for oracle/pyjxl, the float64 closed-form restatement used as the STAGE oracle (token streams, coefficient and
    metadata planes are compared exactly; its pixels only pin that the stages compose).  It does not emulate libjxl's
    approximations (FastLinearToSRGB, RCPPS, FastPowf), which the CUDA path does, so it sits a little further from the
    reference than the product: >= 99.9 % within 1 LSB, none beyond 3, >= 95 % exact."""
    d = np.abs(mine.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 3, (what, int(d.max()))
    assert float((d <= 1).mean()) >= 0.999, (what, float((d <= 1).mean()))
    assert float((d == 0).mean()) >= 0.95, (what, float((d == 0).mean()))
