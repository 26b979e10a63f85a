"""Inputs that make the reference's encoder emit palette transforms (csrc/modular.h: PlanChannels / InversePalettePixel):
frame-level colour palettes, per-channel palettes (16-bit data with few distinct values, two-level alpha), palettes inside
single-section frames and inside the groups of large frames, and a lossless alpha beside lossy colour -- plus effort-1
lossless files, whose modular streams use LZ77."""
import numpy as np

import cases
from oracle import synth


def few_colours(w, h, seed, n, alpha=False):
    """RGB(A) picture with n distinct colours laid out in soft blobs (so that the encoder's palette heuristic accepts it)."""
    rng = np.random.default_rng(seed)
    base = synth.synth_image(w, h, seed, alpha=False).reshape(h, w, 3).astype(np.int32)
    idx = ((base[..., 0] * 3 + base[..., 1] * 5 + base[..., 2] * 7) // 97) % n
    pal = rng.integers(0, 256, (n, 4 if alpha else 3)).astype(np.uint8)
    if alpha:
        pal[:, 3] = np.where(rng.random(n) < 0.5, 255, pal[:, 3])
    return pal[idx]


def gradient(w, h):
    yy, xx = np.mgrid[0:h, 0:w]
    return np.stack([xx * 255 // max(1, w - 1), yy * 255 // max(1, h - 1), (xx + yy) % 256], axis=2).astype(np.uint8)


def two_level_alpha(w, h, seed):
    img = synth.synth_image(w, h, seed, alpha=True).reshape(h, w, 4).copy()
    img[..., 3] = np.where(img[..., 3] > 128, 255, 0)
    return img


# name -> (builder of (img, w, h, ch, bits), encode kwargs)
CASES = {
    "pal_rgb_12colours_700x300": (lambda: (few_colours(700, 300, 1, 12), 700, 300, 3, 8), dict(lossless=True, options={"EFFORT": 7})),
    "pal_rgba_40colours_300x200": (lambda: (few_colours(300, 200, 2, 40, alpha=True), 300, 200, 4, 8), dict(lossless=True, options={"EFFORT": 7})),
    "pal_rgb_200colours_1100x600": (lambda: (few_colours(1100, 600, 3, 200), 1100, 600, 3, 8), dict(lossless=True, options={"EFFORT": 5})),
    "pal_16bit_x257_455x482": (lambda: (synth.synth_image(455, 482, 16).astype(np.uint16) * 257, 455, 482, 3, 16), dict(lossless=True, options={"EFFORT": 3})),
    "pal_16bit_x257_rgba_498x320": (lambda: (synth.synth_image(498, 320, 4, alpha=True).astype(np.uint16) * 257, 498, 320, 4, 16), dict(lossless=True, options={"EFFORT": 8})),
    "pal_local_rgba_822x769": (lambda: (synth.synth_image(822, 769, 81, alpha=True), 822, 769, 4, 8), dict(lossless=True, options={"EFFORT": 5, "EPF": 0})),
    "pal_local_rgba_102x307": (lambda: (synth.synth_image(102, 307, 1, alpha=True), 102, 307, 4, 8), dict(lossless=True, options={"EFFORT": 6})),
    "pal_alpha2_rgba_822x769": (lambda: (two_level_alpha(822, 769, 6), 822, 769, 4, 8), dict(lossless=True, options={"EFFORT": 5})),
    "pal_lossy_colour_lossless_alpha2_560x561": (lambda: (two_level_alpha(560, 561, 7), 560, 561, 4, 8), dict(distance=1.0, alpha_distance=0.0, options={"EFFORT": 7})),
    # libjxl's effort-1 ("fast lossless") encoder: LZ77 run-length copies inside prefix-coded modular streams
    "e1_rgb_1145x612": (lambda: (synth.synth_image(1145, 612, 91), 1145, 612, 3, 8), dict(lossless=True, options={"EFFORT": 1})),
    "e1_rgba_208x244": (lambda: (synth.synth_image(208, 244, 21, alpha=True), 208, 244, 4, 8), dict(lossless=True, options={"EFFORT": 1})),
    "e1_rgb16_485x332": (lambda: (synth.synth_image(485, 332, 13).astype(np.uint16) * 257, 485, 332, 3, 16), dict(lossless=True, options={"EFFORT": 1})),
    # delta palettes (found by the randomised CPU sweep): implicit delta colours through negative indices under predictor 13,
    # and explicit delta entries (nb_deltas = 3) -- frame-level, so the serial inverse kernel runs
    "pal_delta_implicit_215x409": (lambda: (synth.synth_image(215, 409, 18 + 1000 * 102), 215, 409, 3, 8),
                                   dict(lossless=True, options={"EFFORT": 5, "MODULAR_GROUP_SIZE": 1})),
    "pal_delta_entries_2137x567": (lambda: (gradient(2137, 567), 2137, 567, 3, 8), dict(lossless=True, options={"EFFORT": 5, "MODULAR_GROUP_SIZE": 1})),
    "pal_delta_noise_rgba_650x362": (lambda: (np.random.default_rng(5).integers(0, 256, (362, 650, 4)).astype(np.uint8), 650, 362, 4, 8),
                                     dict(lossless=True, options={"EFFORT": 4, "DECODING_SPEED": 4, "MODULAR_GROUP_SIZE": 1})),
    "pal_lossy_colour_lossless_alpha_82x569": (lambda: (synth.synth_image(82, 569, 72, alpha=True), 82, 569, 4, 8), dict(distance=2.0, alpha_distance=0.0, options={"EFFORT": 4})),
}


def make(ref, name):
    build, kw = CASES[name]
    img, w, h, ch, bits = build()
    img = np.ascontiguousarray(img)
    data = cases._cached(name, lambda: ref.encode_ex(img, w, h, ch, bits=bits, **kw))
    return data, img.reshape(h, w, ch), bits, kw
