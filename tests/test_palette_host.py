"""Palette transforms through the CPU emulation of the kernel code (tests/hostemu: the same modular.h / vardct_sections.h
the CUDA kernels compile), against the source image (lossless: integer work, bit-exact) and the reference's alpha."""
import numpy as np
import pytest

import hostemu_lib as H
import palette_cases as P


@pytest.mark.parametrize("name", sorted(P.CASES))
def test_palette_files_decode_exactly(name, ref):
    data, img, bits, kw = P.make(ref, name)
    e = H.Decoded(data)
    assert e.status == 0, (name, e.status, e.failed_stream)
    out = e.render(bits16=bits == 16)
    e.close()
    ch = img.shape[2]
    if kw.get("lossless") and bits == 8:
        # parity is with the REFERENCE's decode; the source image is a second check wherever the reference reproduces it (the
        # delta-palette files do not: libjxl's encoder, asked for lossless with these options, comes back up to 75 off the
        # source through its own decoder)
        want = ref.decode_sampled(data, cfg=2)["pixels"][:, : img.shape[1] * 4].reshape(img.shape[0], img.shape[1], 4)
        if np.array_equal(want[..., :ch][want[..., 3] == 255], img[want[..., 3] == 255]):
            assert np.array_equal(out[..., :ch], img), name
        else:
            assert name.startswith("pal_delta"), name
        a = out[..., 3:4].astype(np.uint16)
        out[..., :3] = (out[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)  # ReformatColorConfig premultiplies
        assert np.array_equal(out, want), name
    elif kw.get("lossless"):
        assert np.array_equal(out[..., :ch], img), name
    else:
        assert np.array_equal(out[..., 3], img[..., 3]), name
        r = ref.decode_sampled(data, cfg=2)
        want = r["pixels"][:, : img.shape[1] * 4].reshape(img.shape[0], img.shape[1], 4)
        assert np.array_equal(out[..., 3], want[..., 3])


def test_channel_plan_replays_transforms():
    """PlanChannels on a hand-made header: palette over channels 0..2 of RGBA, then an RCT can no longer find 3 colour
    channels; a palette on alpha after a colour palette addresses the list that already holds one meta channel."""
    L = H.lib()
    import ctypes as C
    L.emu_plan_channels.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    out = (C.c_uint32 * 32)()
    # transforms as (id, begin_c, num_c, nb_colours)
    tr = (C.c_uint32 * 8)(1, 0, 3, 16, 1, 2, 1, 4)  # palette(RGB) -> list [M0, idx, A]; palette(begin 2 = A, 1 channel)
    assert L.emu_plan_channels(4, 2, tr, out) == 0
    nb_meta, ncoded = out[0], out[1]
    assert (nb_meta, ncoded) == (2, 2)
    assert list(out[2:4]) == [0, 3]          # coded channels live in planes 0 (index of the RGB palette) and 3 (alpha index)
    assert list(out[10:12]) == [1, 0]        # meta channels: the later palette's colours come first in the stream
    tr = (C.c_uint32 * 8)(1, 0, 3, 16, 0, 1, 0, 0)  # RCT at begin 1 needs 3 channels after the palette: only idx, A remain
    assert L.emu_plan_channels(4, 2, tr, out) != 0
