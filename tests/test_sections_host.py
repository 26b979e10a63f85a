"""CPU tests of the shared host/device section decoders (run through tests/hostemu) against the Python oracle:
quantised LF, HF metadata, block placement, AC coefficient planes and modular samples must be bit-exact."""
import numpy as np
import pytest

import cases
import hostemu_lib as H
from oracle.pyjxl import decode as D
from oracle.pyjxl import entropy as ent
from oracle.pyjxl import vardct as vd


def test_tables_match_oracle():
    L = H.lib()
    assert [L.emu_freq_ctx(k) for k in range(1, 64)] == vd.FREQ_CTX[1:]
    assert [L.emu_nnz_ctx(k) for k in range(1, 64)] == vd.NNZ_CTX[1:]
    import ctypes as C
    for i in range(128):
        nb, sym = C.c_uint32(), C.c_uint32()
        L.emu_logcount(i, C.byref(nb), C.byref(sym))
        assert (nb.value, sym.value) == ent.LOGCOUNT_LUT[i], i
    for o, s in vd.ORDER_REP.items():
        buf = np.zeros(65536, np.uint16)
        n = L.emu_natural_order(o, buf.ctypes.data)
        assert buf[:n].tolist() == vd.natural_order(vd.CBX[s], vd.CBY[s])


@pytest.mark.parametrize("name", cases.SMALL)
def test_sections_bit_exact(name, ref):
    data = cases.get(name)
    e = H.Decoded(data)
    assert e.status == 0, (e.status, e.failed_stream)
    r = D.decode(data)
    st = r["stages"]
    i = e.info
    if i["encoding"] == 0:
        lf = e.lf_quant()
        for c in range(3):
            assert (lf[c] == st["lf_quant"][c]).all()
        s, q, sh = e.cells()
        assert ((s & 0x7F) == st["strategy"]).all()
        assert (((s & 0x80) != 0) == st["first"]).all()
        assert (q == st["hf_mul"]).all()
        assert (sh == st["sharpness"]).all()
        xf, bf = e.cfl()
        assert (xf == st["xfromy"]).all() and (bf == st["bfromy"]).all()
        cp = D.coefficient_planes(r["fh"], st["hf_global"], st["strategy"], st["first"], st["coef_list"], i["width"], i["height"])
        assert (e.coef() == cp).all()
    if i["num_mod_channels"]:
        m = e.mod()
        want = np.stack(st["modular_image"])
        assert (m == want).all()
    e.close()


@pytest.mark.parametrize("kind", ["rgb_lossy", "rgba_lossless"])
def test_animation_frames_host(kind, ref):
    """Displayed frame i of an animation decodes as an independent picture (same host/device code as the kernels, run on
    the CPU) and matches the reference's coalesced frame i."""
    import golden_lib
    data = cases.anim_case(kind)
    ra = ref.Anim(data, cfg=2)
    for i in range(cases.ANIM_N):
        e = H.Decoded(data, frame=i)
        assert e.status == 0
        out = e.render()
        e.close()
        want = ra.frame(i)["pixels"][:, : cases.ANIM_W * 4].reshape(cases.ANIM_H, cases.ANIM_W, 4)
        if kind == "rgba_lossless":
            a = out[..., 3:4].astype(np.uint16)
            out[..., :3] = (out[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)
            assert (out == want).all()
        else:
            golden_lib.lossy_close(out, want, "anim %d" % i)
    ra.close()
