"""TEST INFRASTRUCTURE (oracle) — VarDCT section decode (LfGlobal, LfGroup, HfGlobal, PassGroup) and numeric
reconstruction (dequant, LLF, inverse VarDCT, CfL, adaptive LF smoothing, Gaborish, EPF, XYB→RGB, dither) as done by
libjxl 0.12.0 behind the reference's DecodeJpegXlOneShot
(/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:46-175).  Restates SURVEY.md App. B.5, B.7, App. C;
float64 math (the libjxl binary is float32 + approximations: parity is ±1 LSB).
"""
import math
import numpy as np
from .entropy import BitReader, Code, ceil_log2, unpack_signed, read_context_map, read_permutation
from . import modular as mod

# strategy tables (App. B.7): cells covered x / y, coefficient-order id, quant-table id
CBX = [1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32]
CBY = [1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16]
ORDER_ID = [0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12]
QUANT_ID = [0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16]
ORDER_REP = {0: 0, 1: 1, 2: 4, 3: 5, 4: 6, 5: 8, 6: 10, 7: 18, 8: 19, 9: 21, 10: 22, 11: 24, 12: 25}
FREQ_CTX = [None, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22,
            23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30]
NNZ_CTX = [None, 0, 31, 62, 62, 93, 93, 93, 93, 123, 123, 123, 123] + [152] * 8 + [180] * 12 + [206] * 31
DEFAULT_BCTX = [0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14]

_T = [-1.025, -0.78, -0.65012, -0.19041574, -0.208193958, -0.421064, -0.327338457]
_U = [-0.304195821, -0.363303632, -0.356603801, -0.344307452, -0.336995929, -0.301808655, -0.273216844]
_W = [-1.2, -1.2, -0.8, -0.7, -0.7, -0.4, -0.5]
# default DCT quant-weight band parameters, X / Y / B (App. C, read out of the shipped libjxl.so)
QUANT_PARAMS = {
    0: [[3150.0, 0.0, -0.4, -0.4, -0.4, -2.0], [560.0, 0.0, -0.3, -0.3, -0.3, -0.3], [512.0, -2.0, -1.0, 0.0, -1.0, -2.0]],
    4: [[8996.8725711814115328, -1.3000777393353804, -0.49424529824571225, -0.439093774457103443, -0.6350101832695744, -0.90177264050827612, -1.6162099239887414],
        [3191.48366296844234752, -0.67424582104194355, -0.80745813428471001, -0.44925837484843441, -0.35865440981033403, -0.31322389111877305, -0.37615025315725483],
        [1157.50408145487200256, -2.0531423165804414, -1.4, -0.50687130033378396, -0.42708730624733904, -1.4856834539296244, -4.9209142884401604]],
    5: [[15718.40830982518931456, -1.025, -0.98, -0.9012, -0.4, -0.48819395464, -0.421064, -0.27],
        [7305.7636810695983104, -0.8041958212306401, -0.7633036457487539, -0.55660379990111464, -0.49785304658857626, -0.43699592683512467, -0.40180866526242109, -0.27321683125358037],
        [3803.53173721215041536, -3.060733579805728, -2.0413270132490346, -2.0235650159727417, -0.5495389509954993, -0.4, -0.4, -0.3]],
    6: [[7240.7734, -0.7, -0.7, -0.2, -0.2, -0.2, -0.5], [1448.15466, -0.5, -0.5, -0.5, -0.2, -0.2, -0.2], [506.854126, -1.4, -0.2, -0.5, -0.5, -1.5, -3.6]],
    7: [[16283.249, -1.78128457, -1.63090587, -1.0382179, -0.85, -0.7, -0.9, -1.23606384], [5089.15771, -0.320049405, -0.353628486, -0.3034, -0.61, -0.5, -0.5, -0.6],
        [3397.77612, -0.321327358, -0.345076203, -0.7034, -0.9, -1.0, -1.0, -1.17546058]],
    8: [[13844.9707, -0.971138, -0.658, -0.42026, -0.22712, -0.2206, -0.226, -0.6], [4798.96387, -0.611253083, -0.837707877, -0.790148616, -0.269272745, -0.382727683, -0.229242221, -0.20719099],
        [1807.23694, -1.2, -1.2, -0.7, -0.7, -0.7, -0.4, -0.5]],
    9: [[2198.05054, -0.962696254, -0.761942506, -0.655114055], [764.36554, -0.926302016, -0.967522979, -0.278452903], [527.107544, -1.45943856, -1.45008206, -1.58437228]],
    11: [[23966.166] + _T, [8380.19141] + _U, [4493.02393] + _W],
    12: [[15358.8984] + _T, [5597.36035] + _U, [2919.96167] + _W],
    13: [[47932.332] + _T, [16760.3828] + _U, [8986.04785] + _W],
    14: [[30717.7969] + _T, [11194.7207] + _U, [5839.92334] + _W],
    15: [[95864.6641] + _T, [33520.7656] + _U, [17972.0957] + _W],
    16: [[61435.5938] + _T, [22389.4414] + _U, [11679.8467] + _W],
}


def lf_global(cs, fh, md):
    """LfGlobal section (App. B.5).  Returns (g, global_tree, global_code, bitreader positioned after)."""
    br = BitReader(cs, fh['sec_offs'][0] * 8)
    g = {}
    if fh['flags'] & ~0x80:
        raise NotImplementedError('patches/splines/noise/lf-frame flags 0x%x' % fh['flags'])
    if not br.Bool():
        g['lf_dequant'] = [br.F16() for _ in range(3)]
    else:
        g['lf_dequant'] = [1 / 4096, 1 / 512, 1 / 256]
    if fh['encoding'] == 0:
        g['global_scale'] = br.U32((1, 11), (2049, 11), (4097, 12), (8193, 16))
        g['quant_lf'] = br.U32(16, (1, 5), (1, 8), (1, 16))
        if br.Bool():
            g['bctx'] = dict(lf_thr=[[], [], []], qf_thr=[], map=DEFAULT_BCTX)
        else:
            lft = []
            n = 1
            for _ in range(3):
                k = br.u(4)
                lft.append([unpack_signed(br.U32((0, 4), (16, 8), (272, 16), (65808, 32))) for _ in range(k)])
                n *= k + 1
            k = br.u(4)
            qft = [br.U32((0, 2), (4, 3), (12, 5), (44, 8)) + 1 for _ in range(k)]
            n *= k + 1
            cm = read_context_map(br, 3 * 13 * n)
            g['bctx'] = dict(lf_thr=lft, qf_thr=qft, map=cm)
        g['nb_block_ctx'] = max(g['bctx']['map']) + 1
        if br.Bool():
            g['cfl'] = dict(colour_factor=84, base_x=0.0, base_b=1.0, x_lf=128, b_lf=128)
        else:
            g['cfl'] = dict(colour_factor=br.U32(84, 256, (2, 8), (258, 16)), base_x=br.F16(), base_b=br.F16(), x_lf=br.u(8), b_lf=br.u(8))
    has_tree = g['has_tree'] = br.Bool()
    tree = code = None
    if has_tree:
        tree = mod.decode_tree(br)
        code = Code(br, (len(tree) + 1) // 2)
    return g, tree, code, br


def lf_group_rect(fh, lfg):
    W, H = fh['cw'], fh['ch']
    d = fh['group_dim'] * 8
    gx, gy = lfg % fh['nlfx'], lfg // fh['nlfx']
    x0, y0 = gx * d, gy * d
    return x0, y0, min(d, W - x0), min(d, H - y0)


def lf_group(cs, fh, g, tree, code, lfg, br=None):
    """LfGroup section (App. B.7): LF coefficients (Y,X,B order) + HF metadata."""
    x0, y0, w, h = lf_group_rect(fh, lfg)
    w8, h8 = -(-w // 8), -(-h // 8)
    nlf = fh['num_lf_groups']
    if br is None:
        br = BitReader(cs, fh['sec_offs'][1 + lfg] * 8)
    o = {}
    o['extra_precision'] = br.u(2)
    ch, info = mod.decode_channels(br, [(w8, h8)] * 3, 1 + lfg, tree, code)
    o['lf_info'] = info
    o['lf'] = ch
    nb = br.u(ceil_log2(w8 * h8)) + 1
    o['nb_blocks'] = nb
    w64, h64 = -(-w // 64), -(-h // 64)
    ch2, info2 = mod.decode_channels(br, [(w64, h64), (w64, h64), (nb, 2), (w8, h8)], 1 + 2 * nlf + lfg, tree, code)
    o['hf_info'] = info2
    o['xfromy'], o['bfromy'], o['blockinfo'], o['sharpness'] = ch2
    o['end_bit'] = br.p
    return o


def natural_order(cx, cy):
    """Natural coefficient order for a cx×cy-cell block (App. B.7); positions index the 8·min × 8·max array."""
    if cy > cx:
        cx, cy = cy, cx
    xs = cx // cy
    xsm = xs - 1
    xss = (xs - 1).bit_length()
    cur = cx * cy
    out = [0] * (64 * cx * cy)
    N = cx * 8
    for i in range(N):
        for j in range(i + 1):
            x, y = j, i - j
            if i % 2:
                x, y = y, x
            if y & xsm:
                continue
            y >>= xss
            if x < cx and y < cy:
                val = y * cx + x
            else:
                val = cur
                cur += 1
            out[val] = y * N + x
    for ip in range(N - 1, 0, -1):
        i = ip - 1
        for j in range(i + 1):
            x = N - 1 - (i - j)
            y = N - 1 - j
            if i % 2:
                x, y = y, x
            if y & xsm:
                continue
            y >>= xss
            out[cur] = y * N + x
            cur += 1
    assert cur == 64 * cx * cy
    return out


def hf_global(cs, fh, g, br=None):
    """HfGlobal section (App. B.7): coefficient orders + AC code.  Returns (o, accode)."""
    nlf = fh['num_lf_groups']
    ng = fh['num_groups']
    if br is None:
        br = BitReader(cs, fh['sec_offs'][1 + nlf] * 8)
    o = {}
    if not br.u(1):
        raise NotImplementedError('custom quant tables')
    o['num_presets'] = br.u(ceil_log2(ng)) + 1
    assert fh['num_passes'] == 1
    uo = br.U32(0x5F, 0x13, 0, (0, 13))
    o['used_orders'] = uo
    perms = {}
    if uo:
        code = Code(br, 8)
        code.begin(br)
        for ord_ in range(13):
            if not (uo >> ord_) & 1:
                continue
            s = ORDER_REP[ord_]
            llf = CBX[s] * CBY[s]
            size = 64 * llf
            for c in range(3):
                perms[(ord_, c)] = read_permutation(br, code, size, llf)
        assert code.final_ok(), 'order final state'
    orders = {}
    for ord_, s in ORDER_REP.items():
        nat = natural_order(CBX[s], CBY[s])
        for c in range(3):
            p = perms.get((ord_, c))
            orders[(ord_, c)] = [nat[k] for k in p] if p else nat
    o['orders'] = orders  # c index: 0=X 1=Y 2=B (bitstream order of the permutations)
    nctx = o['num_presets'] * g['nb_block_ctx'] * 495
    accode = Code(br, nctx)
    o['end_bit'] = br.p
    return o, accode


def block_context(g, ord_, hf_mul, c, lf_idx=0):
    b = g['bctx']
    qthr = b['qf_thr']
    qi = sum(1 for t in qthr if hf_mul > t)
    nlfctx = 1
    for t in b['lf_thr']:
        nlfctx *= len(t) + 1
    idx = (c ^ 1) if c < 2 else 2  # Y→0, X→1, B→2
    idx = idx * 13 + ord_
    idx = idx * (len(qthr) + 1) + qi
    idx = idx * nlfctx + lf_idx
    return b['map'][idx]


class BlockMap:
    """Places the LF group's BlockInfo entries on the 8×8 cell grid (App. B.7)."""

    def __init__(self, lfo, w8, h8):
        types = lfo['blockinfo'][0]
        muls = lfo['blockinfo'][1]
        cov = np.full((h8, w8), -1, np.int32)
        self.first = {}
        k = 0
        for y in range(h8):
            for x in range(w8):
                if cov[y, x] >= 0:
                    continue
                t = int(types[k])
                q = int(muls[k]) + 1
                k += 1
                assert y + CBY[t] <= h8 and x + CBX[t] <= w8, 'block crosses LF group'
                assert (cov[y:y + CBY[t], x:x + CBX[t]] < 0).all()
                cov[y:y + CBY[t], x:x + CBX[t]] = t
                self.first[(y, x)] = (t, q)
        assert k == lfo['nb_blocks'], (k, lfo['nb_blocks'])
        self.cov = cov


def pass_group(cs, fh, g, bm, lf_rect, lfq, hfo, accode, gidx, br=None):
    """PassGroup section AC part (App. B.7).  Returns (list of (by, bx, c, k, value) with by/bx LF-group-relative cell
    coordinates, c: 0=X 1=Y 2=B, k = coefficient index in coded order), bitreader)."""
    nlf = fh['num_lf_groups']
    if br is None:
        br = BitReader(cs, fh['sec_offs'][1 + nlf + 1 + gidx] * 8)
    gx, gy = gidx % fh['ngx'], gidx // fh['ngx']
    lx0, ly0, lw, lh = lf_rect
    w8, h8 = -(-lw // 8), -(-lh // 8)
    # group origin in cells relative to the LF group
    bx0 = (gx * 256 - lx0) // 8
    by0 = (gy * 256 - ly0) // 8
    bw = min(32, w8 - bx0)
    bh = min(32, h8 - by0)
    hfp = br.u(ceil_log2(hfo['num_presets']))
    nbc = g['nb_block_ctx']
    ctx_off = hfp * 495 * nbc
    accode.begin(br)
    nzmap = [[[0] * bw for _ in range(bh)] for _ in range(3)]
    out = []
    lfthr = g['bctx']['lf_thr']
    for by in range(bh):
        for bx in range(bw):
            key = (by0 + by, bx0 + bx)
            if key not in bm.first:
                continue
            t, q = bm.first[key]
            cx, cy = CBX[t], CBY[t]
            covered = cx * cy
            l2 = covered.bit_length() - 1
            size = 64 * covered
            ord_ = ORDER_ID[t]
            lf_idx = 0
            if any(lfthr):
                # bucket = (bx * (nB+1) + bb) * (nY+1) + by from the quantised LF values; thresholds are stored X,Y,B,
                # lfq planes are in Y,X,B order  [M: libjxl DequantDC; no fixture exercises LF thresholds]
                bkt = [sum(1 for th in lfthr[ci] if int(lfq[pl_][key[0], key[1]]) > th) for ci, pl_ in ((0, 1), (1, 0), (2, 2))]
                lf_idx = (bkt[0] * (len(lfthr[2]) + 1) + bkt[2]) * (len(lfthr[1]) + 1) + bkt[1]
            for c in (1, 0, 2):
                nzm = nzmap[c]
                if bx == 0 and by == 0:
                    pred = 32
                elif bx == 0:
                    pred = nzm[by - 1][bx]
                elif by == 0:
                    pred = nzm[by][bx - 1]
                else:
                    pred = (nzm[by - 1][bx] + nzm[by][bx - 1] + 1) // 2
                bc = block_context(g, ord_, q, c, lf_idx)
                nzc = pred if pred < 8 else (36 if pred >= 64 else 4 + pred // 2)
                nz = accode.read(br, ctx_off + nzc * nbc + bc)
                assert nz <= size - covered, ('nz too big', nz, size, covered)
                v = (nz + covered - 1) >> l2
                for yy in range(cy):
                    for xx in range(cx):
                        nzm[by + yy][bx + xx] = v
                ho = ctx_off + nbc * 37 + 458 * bc
                prev = 0 if nz > size // 16 else 1
                kk = covered
                while kk < size and nz != 0:
                    nl = (nz + covered - 1) >> l2
                    ctx = ho + (NNZ_CTX[nl] + FREQ_CTX[kk >> l2]) * 2 + prev
                    u = accode.read(br, ctx)
                    prev = 1 if u else 0
                    nz -= prev
                    if u:
                        out.append((key[0], key[1], c, kk, unpack_signed(u)))
                    kk += 1
                assert nz == 0, 'nonzeros left'
    assert accode.final_ok(), 'AC final state'
    return out, br


# ------------------------------------------------------------------------------------------------ numeric pipeline
ID_WEIGHTS = [[280.0, 3160.0, 3160.0], [60.0, 864.0, 864.0], [18.0, 200.0, 200.0]]
DCT2_WEIGHTS = [[3840.0, 2560.0, 1280.0, 640.0, 480.0, 300.0], [960.0, 640.0, 320.0, 180.0, 140.0, 120.0], [640.0, 320.0, 128.0, 64.0, 32.0, 16.0]]
DCT4_PARAMS = [[2200.0, 0.0, 0.0, 0.0], [392.0, 0.0, 0.0, 0.0], [112.0, -0.25, -0.25, -0.5]]
AFV_WEIGHTS = [[3072.0, 3072.0, 256.0, 256.0, 256.0, 414.0, 0.0, 0.0, 0.0], [1024.0, 1024.0, 50.0, 50.0, 50.0, 58.0, 0.0, 0.0, 0.0],
               [384.0, 384.0, 12.0, 12.0, 12.0, 22.0, -0.25, -0.25, -0.25]]
AFV_FREQS = [0, 0, 0.8517778890324296, 5.37778436506804, 0, 0, 4.734747904497923, 5.449245381693219, 1.6598270267479331, 4,
             7.275749096817861, 10.423227632456525, 2.662932286148962, 7.630657783650829, 8.962388608184032, 12.97166202570235]


def afv_weights(c):
    """AFV quant weights (table 10): corner weights + frequency-interpolated bands + 4x8 and 4x4 tables interleaved."""
    a = AFV_WEIGHTS[c]
    w48 = quant_weights(4, 8, QUANT_PARAMS[9][c])
    w44 = quant_weights(4, 4, DCT4_PARAMS[c])
    lo = 0.8517778890324296
    hi = 12.97166202570235 - lo + 1e-6
    bands = [a[5]]
    for i in range(1, 4):
        v = a[i + 5]
        bands.append(bands[-1] * ((1 + v) if v > 0 else 1 / (1 - v)))
    w = np.zeros((8, 8))
    w[0, 0] = 1.0
    w[1, 0] = a[0]   # set_weight(x=0, y=1)
    w[0, 1] = a[1]   # set_weight(x=1, y=0)
    w[2, 0] = a[2]   # (0, 2)
    w[0, 2] = a[3]   # (2, 0)
    w[2, 2] = a[4]   # (2, 2)
    for y in range(4):
        for x in range(4):
            if x < 2 and y < 2:
                continue
            pos = (AFV_FREQS[y * 4 + x] - lo) * 3 / hi
            i = int(pos)
            w[2 * y, 2 * x] = bands[i] * (bands[i + 1] / bands[i]) ** (pos - i)
    for y in range(4):
        for x in range(8):
            if x == 0 and y == 0:
                continue
            w[2 * y + 1, x] = w48[y, x]
    for y in range(4):
        for x in range(4):
            if x == 0 and y == 0:
                continue
            w[2 * y, 2 * x + 1] = w44[y, x]
    return w


def quant_weights(rows, cols, params):
    nb = len(params)
    bands = [params[0]]
    for v in params[1:]:
        bands.append(bands[-1] * ((1 + v) if v > 0 else 1 / (1 - v)))
    scale = (nb - 1) / (math.sqrt(2) + 1e-6)
    rc = scale / (cols - 1)
    rr = scale / (rows - 1)
    w = np.zeros((rows, cols))
    for y in range(rows):
        for x in range(cols):
            d = math.hypot(x * rc, y * rr)
            i = int(d)
            fr = d - i
            a = bands[i]
            bb = bands[min(i + 1, nb - 1)]
            w[y, x] = a * (bb / a) ** fr
    return w


_wcache = {}


def dequant_matrix(t, c):
    """1/weight for strategy t, channel c (0=X,1=Y,2=B), shaped 8·min × 8·max."""
    cx, cy = CBX[t], CBY[t]
    kr, kc = 8 * min(cx, cy), 8 * max(cx, cy)
    key = (QUANT_ID[t], c)
    if key not in _wcache:
        q = QUANT_ID[t]
        if q == 9:
            w48 = quant_weights(4, 8, QUANT_PARAMS[9][c])
            full = np.zeros((8, 8))
            for y in range(8):
                full[y] = w48[y // 2]
            _wcache[key] = 1.0 / full
        elif q in QUANT_PARAMS:
            _wcache[key] = 1.0 / quant_weights(kr, kc, QUANT_PARAMS[q][c])
        elif q == 1:  # IDENTITY: 3 weights
            w = np.full((8, 8), ID_WEIGHTS[c][0])
            w[0, 1] = w[1, 0] = ID_WEIGHTS[c][1]
            w[1, 1] = ID_WEIGHTS[c][2]
            _wcache[key] = 1.0 / w
        elif q == 2:  # DCT2X2: 6 weights by frequency ring
            d = DCT2_WEIGHTS[c]
            w = np.zeros((8, 8))
            w[0, 0] = 1.0
            w[0, 1] = w[1, 0] = d[0]
            w[1, 1] = d[1]
            w[0:2, 2:4] = d[2]
            w[2:4, 0:2] = d[2]
            w[2:4, 2:4] = d[3]
            w[0:4, 4:8] = d[4]
            w[4:8, 0:4] = d[4]
            w[4:8, 4:8] = d[5]
            _wcache[key] = 1.0 / w
        elif q == 3:  # DCT4X4: 4x4 weights, each used for a 2x2 cell group
            w44 = quant_weights(4, 4, DCT4_PARAMS[c])
            w = np.repeat(np.repeat(w44, 2, 0), 2, 1)
            _wcache[key] = 1.0 / w
        elif q == 10:  # AFV
            _wcache[key] = 1.0 / afv_weights(c)
        else:
            raise NotImplementedError('quant table %d (strategy %d)' % (q, t))
    return _wcache[key]


def dct_matrix(N):
    M = np.zeros((N, N))
    for n in range(N):
        for k in range(N):
            M[n, k] = (1.0 if k == 0 else math.sqrt(2)) * math.cos((2 * n + 1) * k * math.pi / (2 * N))
    return M


_C = {}


def Cm(N):
    if N not in _C:
        _C[N] = dct_matrix(N)
    return _C[N]


def llf_scale(N, k):
    return 1.0 if k == 0 else 8 * math.sin(k * math.pi / (16 * N)) / math.sin(k * math.pi / (2 * N))


QUANT_BIAS = [0.945349932, 0.929945469, 0.950064898, 0.145]


def adjust_bias(q, c):
    a = np.abs(q)
    with np.errstate(divide='ignore', invalid='ignore'):
        big = q - QUANT_BIAS[3] / q
    return np.where(a == 0, 0.0, np.where(a == 1, np.sign(q) * QUANT_BIAS[c], big))


def lf_dequant(g, lfo):
    """Returns [X, Y, B] float LF planes after dequant + CfL-DC (App. B.7), and per-channel multipliers."""
    inv_gs = 65536.0 / g['global_scale']
    mul = [g['lf_dequant'][c] * inv_gs / g['quant_lf'] for c in range(3)]
    ep = 1 << lfo['extra_precision']
    lfq = lfo['lf']
    Y = lfq[0].astype(np.float64) * mul[1] / ep
    X = lfq[1].astype(np.float64) * mul[0] / ep
    B = lfq[2].astype(np.float64) * mul[2] / ep
    cf = g['cfl']
    X = X + (cf['base_x'] + (cf['x_lf'] - 128) / cf['colour_factor']) * Y
    B = B + (cf['base_b'] + (cf['b_lf'] - 128) / cf['colour_factor']) * Y
    return [X, Y, B], mul


def adaptive_lf_smooth(dc, mul):
    kW1 = 0.20345139757231578
    kW2 = 0.0334829185968739
    kW0 = 1 - 4 * (kW1 + kW2)
    if dc[0].shape[0] < 3 or dc[0].shape[1] < 3:
        return [a.copy() for a in dc]
    sm = []
    for c in range(3):
        a = dc[c]
        s = a.copy()
        s[1:-1, 1:-1] = kW0 * a[1:-1, 1:-1] + kW1 * (a[:-2, 1:-1] + a[2:, 1:-1] + a[1:-1, :-2] + a[1:-1, 2:]) + \
            kW2 * (a[:-2, :-2] + a[:-2, 2:] + a[2:, :-2] + a[2:, 2:])
        sm.append(s)
    gap = np.full(dc[0].shape, 0.5)
    for c in range(3):
        gap = np.maximum(gap, np.abs((dc[c] - sm[c]) / mul[c]))
    fac = np.maximum(0, 3 - 4 * gap)
    out = []
    for c in range(3):
        o = (sm[c] - dc[c]) * fac + dc[c]
        o2 = dc[c].copy()
        o2[1:-1, 1:-1] = o[1:-1, 1:-1]
        out.append(o2)
    return out


SUPPORTED_STRATEGIES = {0, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 18, 19, 20, 21, 22, 23, 24, 25, 26}


def idct_block(t, K, dcs, by, bx):
    """K: dequantised coefficient array (8·min × 8·max) for one channel; dcs: LF plane; returns pixels (8cy × 8cx)."""
    cx, cy = CBX[t], CBY[t]
    R, Cw = 8 * cy, 8 * cx
    if t in (12, 13):
        Kc = K.copy()
        Kc[0, 0] = dcs[by, bx]
        out = np.zeros((8, 8))
        b0, b1 = Kc[0, 0], Kc[1, 0]
        dd = [b0 + b1, b0 - b1]
        for sub in range(2):
            blk = np.zeros((4, 8))
            for iy in range(4):
                blk[iy, :] = Kc[sub + iy * 2, :]
            blk[0, 0] = dd[sub]
            px48 = Cm(4) @ blk @ Cm(8).T
            if t == 12:
                out[sub * 4:sub * 4 + 4, :] = px48
            else:
                out[:, sub * 4:sub * 4 + 4] = px48.T
        return out
    if t in (1, 2, 3) or 14 <= t <= 17:
        Kc = K.copy()
        Kc[0, 0] = dcs[by, bx]
        return special_8x8(t, Kc)
    if t not in SUPPORTED_STRATEGIES:
        raise NotImplementedError('strategy %d in oracle' % t)
    tall = R >= Cw
    d = dcs[by:by + cy, bx:bx + cx]
    D = (Cm(cy).T / cy) @ d @ (Cm(cx).T / cx).T
    for ky in range(cy):
        for kx in range(cx):
            D[ky, kx] *= llf_scale(cy, ky) * llf_scale(cx, kx)
    Kc = K.copy()
    if tall:
        Kc[:cx, :cy] = D.T
        return Cm(R) @ Kc.T @ Cm(Cw).T
    Kc[:cy, :cx] = D
    return Cm(R) @ Kc @ Cm(Cw).T


_AFV = None


def afv_basis():
    """16x16 AFV basis (JPEG XL spec constant; rows 1..15 read out of the reference binary by
    oracle/gen_tables.py, row 0 = 0.25)."""
    global _AFV
    if _AFV is None:
        import os
        here = os.path.dirname(os.path.abspath(__file__))
        txt = open(os.path.join(here, '..', '..', 'jxl_coder_b200', 'csrc', 'tables', 'afv_basis.inc')).read()
        vals = [float(v.rstrip('f')) for v in txt.replace('\n', ' ').split(',') if v.strip() and not v.strip().startswith('//')]
        _AFV = np.array(vals).reshape(16, 16)
    return _AFV


def _idct2_top(blk, S):
    """libjxl IDCT2TopBlock<S>: one level of 2x2 Hadamard synthesis on the top-left SxS of an 8x8 array."""
    n = S // 2
    out = blk.copy()
    for y in range(n):
        for x in range(n):
            c00, c01, c10, c11 = blk[y, x], blk[y, n + x], blk[y + n, x], blk[y + n, n + x]
            out[2 * y, 2 * x] = c00 + c01 + c10 + c11
            out[2 * y, 2 * x + 1] = c00 + c01 - c10 - c11
            out[2 * y + 1, 2 * x] = c00 - c01 + c10 - c11
            out[2 * y + 1, 2 * x + 1] = c00 - c01 - c10 + c11
    return out


def special_8x8(t, K):
    """IDENTITY (1), DCT2X2 (2), DCT4X4 (3), AFV0-3 (14-17) on the 8x8 coefficient array K (K[0,0] = LF value).
    Restated from the JPEG XL inverse transforms and checked against the reference binary's own
    TransformToPixels (oracle/ref_api.cpp: ref_transform_to_pixels) in tests/test_oracle_pinned.py."""
    px = np.zeros((8, 8))
    if t == 2:
        b = _idct2_top(K, 2)
        b = _idct2_top(b, 4)
        return _idct2_top(b, 8)
    if t in (1, 3):
        b00, b01, b10, b11 = K[0, 0], K[0, 1], K[1, 0], K[1, 1]
        dcs = [b00 + b01 + b10 + b11, b00 + b01 - b10 - b11, b00 - b01 + b10 - b11, b00 - b01 - b10 + b11]
        for y in range(2):
            for x in range(2):
                sub = np.zeros((4, 4))
                for iy in range(4):
                    for ix in range(4):
                        sub[iy, ix] = K[y + iy * 2, x + ix * 2]
                if t == 3:
                    sub[0, 0] = dcs[y * 2 + x]
                    px[4 * y:4 * y + 4, 4 * x:4 * x + 4] = Cm(4) @ sub.T @ Cm(4).T
                else:
                    resid = sub.sum() - sub[0, 0]
                    base = dcs[y * 2 + x] - resid / 16.0
                    q = sub + base
                    q[1, 1] = base
                    q[0, 0] = sub[1, 1] + base
                    px[4 * y:4 * y + 4, 4 * x:4 * x + 4] = q
        return px
    kind = t - 14
    ax, ay = kind & 1, kind >> 1
    b00, b01, b10 = K[0, 0], K[0, 1], K[1, 0]
    dcs = [(b00 + b10 + b01) * 4.0, b00 + b10 - b01, b00 - b10]
    co = np.zeros(16)
    for iy in range(4):
        for ix in range(4):
            co[iy * 4 + ix] = K[iy * 2, ix * 2]
    co[0] = dcs[0]
    blk = (afv_basis().T @ co).reshape(4, 4)
    for iy in range(4):
        for ix in range(4):
            px[iy + ay * 4, ax * 4 + ix] = blk[3 - iy if ay else iy, 3 - ix if ax else ix]
    sub = np.zeros((4, 4))
    for iy in range(4):
        for ix in range(4):
            sub[iy, ix] = K[iy * 2, ix * 2 + 1]
    sub[0, 0] = dcs[1]
    x0 = 0 if ax else 4
    px[ay * 4:ay * 4 + 4, x0:x0 + 4] = Cm(4) @ sub.T @ Cm(4).T
    sub = np.zeros((4, 8))
    for iy in range(4):
        sub[iy, :] = K[1 + iy * 2, :]
    sub[0, 0] = dcs[2]
    y0 = 0 if ay else 4
    px[y0:y0 + 4, :] = Cm(4) @ sub @ Cm(8).T
    return px


def mpad(a, p):
    return np.pad(a, p, mode='symmetric')


def gaborish(pl, w=None):
    w1 = [0.115169525] * 3
    w2 = [0.061248592] * 3
    if w is not None:
        w1 = [w[0], w[2], w[4]]
        w2 = [w[1], w[3], w[5]]
    out = []
    for c, a in enumerate(pl):
        mul = 1 / (1 + 4 * (w1[c] + w2[c]))
        p = mpad(a, 1)
        out.append(mul * (p[1:-1, 1:-1] + w1[c] * (p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]) +
                          w2[c] * (p[:-2, :-2] + p[:-2, 2:] + p[2:, :-2] + p[2:, 2:])))
    return out


def epf_inv_sigma(g, fh, hfq, sharp):
    """Per-cell 1/sigma (negative) as libjxl computes it (App. B.7)."""
    lut = np.array(fh['epf_sharp_lut']) if fh['epf_sharp_lut'] else np.arange(8) / 7.0
    qs = g['global_scale'] / 65536.0
    with np.errstate(divide='ignore'):
        sig = fh['epf_quant_mul'] / (qs * hfq * (-1.1715728752538099)) * lut[sharp]
    sig = np.minimum(-1e-4, sig)
    return 1.0 / sig


def epf(pl, inv_sigma_cell, fh, W, H):
    """EPF stages per epf_iters (App. B.7): iters==3: stage0; iters>=1: stage1; iters>=2: stage2."""
    inv_px = np.repeat(np.repeat(inv_sigma_cell, 8, 0), 8, 1)[:H, :W]
    scale = fh['epf_ch_scale'] or [40.0, 5.0, 3.5]
    border = fh['epf_border']
    skip = inv_px < -3.90524291751269967
    xb = (np.arange(W) % 8 == 0) | (np.arange(W) % 8 == 7)
    yb = (np.arange(H) % 8 == 0) | (np.arange(H) % 8 == 7)
    isborder = yb[:, None] | xb[None, :]
    plus = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]

    def stage(pl, neigh, sad_plus, sm):
        P = [mpad(a, 3) for a in pl]

        def sh(p, dy, dx):
            return p[3 + dy:3 + dy + H, 3 + dx:3 + dx + W]
        mul = np.where(isborder, sm * border, sm)
        isg = inv_px * mul
        wacc = np.ones((H, W))
        acc = [a.copy() for a in pl]
        for dy, dx in neigh:
            sad = 0
            for c in range(3):
                if sad_plus:
                    for py, px_ in plus:
                        sad = sad + np.abs(sh(P[c], dy + py, dx + px_) - sh(P[c], py, px_)) * scale[c]
                else:
                    sad = sad + np.abs(sh(P[c], dy, dx) - sh(P[c], 0, 0)) * scale[c]
            wgt = np.maximum(0, 1 + sad * isg)
            wacc += wgt
            for c in range(3):
                acc[c] += wgt * sh(P[c], dy, dx)
        return [np.where(skip, pl[c], acc[c] / wacc) for c in range(3)]
    n4 = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    n12 = [(dy, dx) for dy in range(-2, 3) for dx in range(-2, 3) if 0 < abs(dy) + abs(dx) <= 2]
    it = fh['epf_iters']
    if it == 3:
        pl = stage(pl, n12, True, 1.65 * fh['epf_pass0'])
    if it >= 1:
        pl = stage(pl, n4, True, 1.65)
    if it >= 2:
        pl = stage(pl, n4, False, 1.65 * fh['epf_pass2'])
    return pl


OPSIN_INV = np.array([[11.031566901960783, -9.866943921568629, -0.16462299647058826],
                      [-3.254147380392157, 4.418770392156863, -0.16462299647058826],
                      [-3.6588512862745097, 2.7129230470588235, 1.9459282392156863]])
OPSIN_BIAS = 0.0037930732552754493


def xyb_to_linear(pl, intensity_target=255.0):
    X, Y, B = pl
    cb = OPSIN_BIAS ** (1 / 3)
    mr = (Y + X + cb) ** 3 - OPSIN_BIAS
    mg = (Y - X + cb) ** 3 - OPSIN_BIAS
    mb = (B + cb) ** 3 - OPSIN_BIAS
    s = 255.0 / intensity_target
    return np.stack([(OPSIN_INV[i][0] * mr + OPSIN_INV[i][1] * mg + OPSIN_INV[i][2] * mb) * s for i in range(3)], -1)


def srgb_oetf(v):
    v = np.clip(v, 0, 1)
    return np.where(v <= 0.0031308, 12.92 * v, 1.055 * np.power(np.maximum(v, 1e-12), 1 / 2.4) - 0.055)


def rgb_to_xyz_matrix(prim, wp):
    (xr, yr), (xg, yg), (xb, yb) = prim
    xw, yw = wp
    P = np.array([[xr / yr, xg / yg, xb / yb], [1, 1, 1], [(1 - xr - yr) / yr, (1 - xg - yg) / yg, (1 - xb - yb) / yb]])
    Wv = np.array([xw / yw, 1, (1 - xw - yw) / yw])
    S = np.linalg.solve(P, Wv)
    return P * S


PRIMARIES = {1: [(0.64, 0.33), (0.30, 0.60), (0.15, 0.06)], 9: [(0.708, 0.292), (0.170, 0.797), (0.131, 0.046)],
             11: [(0.680, 0.320), (0.265, 0.690), (0.150, 0.060)]}
D65 = (0.3127, 0.3290)


def to_u8(rgbf, dither):
    H, W, _ = rgbf.shape
    yy, xx = np.mgrid[0:H, 0:W]
    return np.clip(np.rint(rgbf * 255 + dither[yy % 32, xx % 32][:, :, None]), 0, 255).astype(np.uint8)
