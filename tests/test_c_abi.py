"""The C-ABI library loads and exports every symbol include/jxlb200.h declares; the header-only entry points and the
argument validation work without a GPU; decode entry points fail loudly (no CPU fallback) when no device exists."""
import os
import re

import pytest

import golden_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def J():
    import jxl_coder_b200 as J
    if not os.path.exists(J.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    return J


def test_exports_every_declared_symbol(J):
    hdr = open(os.path.join(ROOT, "include", "jxlb200.h")).read()
    declared = set(re.findall(r"JXLB_API[^;]*?\b(jxlb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    L = J.load_library()
    for name in declared:
        assert hasattr(L, name), name


def test_get_size_and_signature(J):
    g = golden_lib.load("rgba_lossy_300x203")
    assert J.JxlCoder.get_size(g["jxl"]) == (300, 203)
    assert J.JxlCoder.is_jxl(g["jxl"])
    assert J.JxlCoder.get_size(b"definitely not a jxl file") is None
    assert J.JxlCoder.get_size(g["jxl"][:3]) is None


def test_argument_validation_matches_reference_messages(J):
    g = golden_lib.load("rgb_lossy_64")
    with pytest.raises(J.JxlCoderError) as e:
        J.JxlCoder.decode(g["jxl"], 0)
    assert e.value.status == 4 and "Invalid Color Config: 0 was passed" in e.value.message
    with pytest.raises(J.JxlCoderError) as e:
        J.JxlCoder.decode_sampled(g["jxl"], -1, -1, 2, 0, 4)
    assert "Invalid Scale Mode was passed" in e.value.message
    with pytest.raises(J.JxlCoderError) as e:
        J.JxlCoder.decode_sampled(g["jxl"], -1, -1, 2, 1, 0)
    assert "Invalid Sampler: 0 was passed" in e.value.message


def test_no_cpu_fallback(J):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = golden_lib.load("rgb_lossy_64")
    with pytest.raises(J.JxlCoderError) as e:
        J.JxlCoder.decode(g["jxl"])
    assert e.value.status == 7  # JXLB_ERROR_NO_DEVICE


def test_animation_header_queries(J):
    # frame table parsing is CPU-only
    g = golden_lib.load("rgb_lossy_64")
    a = J.JxlAnimatedImage(g["jxl"])
    assert a.number_of_frames == 1 and a.get_width() == 64 and a.get_height() == 64
    a.close()
