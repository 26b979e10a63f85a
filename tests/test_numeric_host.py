"""CPU tests of the numeric host/device functions (dequant, LLF, inverse VarDCT incl. the special 8x8 transforms,
Gaborish, EPF, colour, dither, pack) run through tests/hostemu, against the golden vectors / the reference."""
import numpy as np
import pytest

import cases
import golden_lib
import hostemu_lib as H


@pytest.mark.parametrize("name", golden_lib.names())
def test_pipeline_matches_golden(name):
    g = golden_lib.load(name)
    e = H.Decoded(g["jxl"])
    assert e.status == 0
    out = e.render()
    if "lossless" in name:
        assert (out == g["raw"]).all()
    else:
        golden_lib.lossy_close(out, g["raw"], name)
    e.close()


@pytest.mark.parametrize("name", [n for n in cases.SMALL if n not in golden_lib.names()])
def test_pipeline_matches_reference(name, ref):
    data = cases.get(name)
    raw, _ = ref.decode_oneshot(data)
    e = H.Decoded(data)
    assert e.status == 0
    out = e.render()
    if "lossless" in name:
        assert (out == raw).all()
    else:
        golden_lib.lossy_close(out, raw, name)
    e.close()
