"""TEST INFRASTRUCTURE (oracle) — modular sub-bitstream decode (MA tree, predictors incl. the weighted predictor,
inverse RCT) as done by libjxl 0.12.0 for LF coefficients, HF metadata, alpha and lossless frames behind the
reference's DecodeJpegXlOneShot (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:46-175).
Restates SURVEY.md App. B.6.  Pure-Python loops: small cases only.
"""
import numpy as np
from .entropy import Code, unpack_signed


def decode_tree(br, limit=1 << 22):
    """MA tree: list of nodes (prop, splitval, left|ctx, right, predictor, offset, multiplier); prop == -1 ⇒ leaf."""
    code = Code(br, 6)
    code.begin(br)
    tree = []
    to_decode = 1
    leaf = 0
    while to_decode > 0:
        to_decode -= 1
        p1 = code.read(br, 1)
        assert p1 <= 256
        prop = p1 - 1
        if prop == -1:
            pred = code.read(br, 2)
            assert pred < 14, pred
            off = unpack_signed(code.read(br, 3))
            ml = code.read(br, 4)
            assert ml < 31
            mb = code.read(br, 5)
            tree.append((-1, 0, leaf, 0, pred, off, (mb + 1) << ml))
            leaf += 1
            continue
        sv = unpack_signed(code.read(br, 0))
        tree.append((prop, sv, len(tree) + to_decode + 1, len(tree) + to_decode + 2, 0, 0, 1))
        to_decode += 2
        assert len(tree) < limit
    assert code.final_ok(), 'tree final state'
    return tree


class WPHeader:
    def __init__(self, br=None):
        self.p1C = 16
        self.p2C = 10
        self.p3 = [7, 7, 7, 0, 0]
        self.w = [0xd, 0xc, 0xc, 0xc]
        if br is not None and not br.Bool():
            self.p1C = br.u(5)
            self.p2C = br.u(5)
            self.p3 = [br.u(5) for _ in range(5)]
            self.w = [br.u(4) for _ in range(4)]


def read_transform(br):
    t = {'id': br.u(2)}
    bc = ((0, 3), (8, 6), (72, 10), (1096, 13))
    if t['id'] == 0:
        t['begin_c'] = br.U32(*bc)
        t['rct_type'] = br.U32(6, (0, 2), (2, 4), (10, 6))
    elif t['id'] == 1:
        t['begin_c'] = br.U32(*bc)
        t['num_c'] = br.U32(1, 3, 4, (1, 13))
        t['nb_colours'] = br.U32((0, 8), (256, 10), (1280, 12), (5376, 16))
        t['nb_deltas'] = br.U32(0, (1, 8), (257, 10), (1281, 16))
        t['d_pred'] = br.u(4)
    elif t['id'] == 2:
        n = br.U32(0, (1, 4), (9, 6), (41, 8))
        t['sq'] = []
        for _ in range(n):
            t['sq'].append({'horizontal': br.Bool(), 'in_place': br.Bool(), 'begin_c': br.U32(*bc), 'num_c': br.U32(1, 2, 3, (4, 4))})
    return t


class WPState:
    def __init__(self, h, xs):
        self.h = h
        self.xs = xs
        self.pe = [[0] * ((xs + 2) * 2) for _ in range(4)]
        self.err = [0] * ((xs + 2) * 2)
        self.div = [(1 << 24) // (i + 1) for i in range(64)]
        self.pred = 0
        self.prediction = [0] * 4
        self.prop = 0

    def _ew(self, x, mw):
        sh = max(0, (x + 1).bit_length() - 1 - 5)
        return 4 + ((mw * self.div[x >> sh]) >> sh)

    def predict(self, x, y, N, W, NE, NW, NN):
        xs = self.xs
        cur = 0 if y & 1 else xs + 2
        prv = xs + 2 if y & 1 else 0
        pN = prv + x
        pNE = pN + 1 if x < xs - 1 else pN
        pNW = pN - 1 if x > 0 else pN
        w = [self._ew(self.pe[i][pN] + self.pe[i][pNE] + self.pe[i][pNW], self.h.w[i]) for i in range(4)]
        N <<= 3
        W <<= 3
        NE <<= 3
        NW <<= 3
        NN <<= 3
        teW = 0 if x == 0 else self.err[cur + x - 1]
        teN = self.err[pN]
        teNW = self.err[pNW]
        sumWN = teN + teW
        teNE = self.err[pNE]
        p = teW
        if abs(teN) > abs(p):
            p = teN
        if abs(teNW) > abs(p):
            p = teNW
        if abs(teNE) > abs(p):
            p = teNE
        self.prop = p
        P = self.prediction
        P[0] = W + NE - N
        P[1] = N - (((sumWN + teNE) * self.h.p1C) >> 5)
        P[2] = W - (((sumWN + teNW) * self.h.p2C) >> 5)
        P[3] = N - ((teNW * self.h.p3[0] + teN * self.h.p3[1] + teNE * self.h.p3[2] + (NN - N) * self.h.p3[3] + (NW - W) * self.h.p3[4]) >> 5)
        ws = sum(w)
        lw = ws.bit_length() - 1
        w = [wi >> (lw - 4) for wi in w]
        ws = sum(w)
        sm = (ws >> 1) - 1
        for i in range(4):
            sm += P[i] * w[i]
        self.pred = (sm * self.div[ws - 1]) >> 24
        if ((teN ^ teW) | (teN ^ teNW)) > 0:
            return (self.pred + 3) >> 3
        mx = max(W, NE, N)
        mn = min(W, NE, N)
        self.pred = max(mn, min(mx, self.pred))
        return (self.pred + 3) >> 3

    def update(self, val, x, y):
        xs = self.xs
        cur = 0 if y & 1 else xs + 2
        prv = xs + 2 if y & 1 else 0
        val <<= 3
        self.err[cur + x] = self.pred - val
        for i in range(4):
            e = (abs(self.prediction[i] - val) + 3) >> 3
            self.pe[i][cur + x] = e
            self.pe[i][prv + x + 1] += e


def _tdiv(a, b):
    return a // b if a >= 0 else -((-a) // b)


def decode_channels(br, chans, stream_id, gtree, gcode, max_chan_size=1 << 30, nb_meta=0):
    """chans: list of (w, h).  Reads GroupHeader + channel data.  Returns (list of np.int64 [h,w], info); transforms
    are NOT undone here (see undo_transforms)."""
    if not chans:
        return [], {}
    use_global = br.Bool()
    wp = WPHeader(br)
    nt = br.U32(0, 1, (2, 4), (18, 8))
    tr = [read_transform(br) for _ in range(nt)]
    info = {'use_global': use_global, 'transforms': tr, 'wp': wp}
    if any(t['id'] != 0 for t in tr):
        raise NotImplementedError('palette/squeeze in oracle: %r' % tr)
    if use_global:
        tree, code = gtree, gcode
    else:
        tree = decode_tree(br)
        code = Code(br, (len(tree) + 1) // 2)
    info['tree_size'] = len(tree)
    dm = 0
    for i, (w, h) in enumerate(chans):
        if not w or not h:
            continue
        if i >= nb_meta and (w > max_chan_size or h > max_chan_size):
            break
        dm = max(dm, w)
    code.begin(br)
    uses_wp = any(n[0] == 15 or (n[0] == -1 and n[4] == 6) for n in tree)
    out = []
    for ci, (w, h) in enumerate(chans):
        if not w or not h:
            out.append(np.zeros((h, w), np.int64))
            continue
        if ci >= nb_meta and (w > max_chan_size or h > max_chan_size):
            break  # this and all later channels are coded per group
        img = [[0] * w for _ in range(h)]
        wps = WPState(wp, w) if uses_wp else None
        for y in range(h):
            row = img[y]
            rN = img[y - 1] if y > 0 else None
            rNN = img[y - 2] if y > 1 else None
            prev9 = 0
            for x in range(w):
                W = row[x - 1] if x > 0 else (rN[x] if y > 0 else 0)
                N = rN[x] if y > 0 else W
                NW = rN[x - 1] if (x > 0 and y > 0) else W
                NE = rN[x + 1] if (x + 1 < w and y > 0) else N
                NN = rNN[x] if y > 1 else N
                WW = row[x - 2] if x > 1 else W
                NEE = rN[x + 2] if (x + 2 < w and y > 0) else NE
                wpred = 0
                wprop = 0
                if wps:
                    wpred = wps.predict(x, y, N, W, NE, NW, NN)
                    wprop = wps.prop
                p9 = W + N - NW
                props = (ci, stream_id, y, x, abs(N), abs(W), N, W, W - prev9, p9, W - NW, NW - N, N - NE, N - NN, W - WW, wprop)
                prev9 = p9
                n = tree[0]
                while n[0] >= 0:
                    assert n[0] < 16, 'prev-channel property %d unsupported in oracle' % n[0]
                    n = tree[n[2]] if props[n[0]] > n[1] else tree[n[3]]
                pr = n[4]
                if pr == 0:
                    g = 0
                elif pr == 1:
                    g = W
                elif pr == 2:
                    g = N
                elif pr == 3:
                    g = _tdiv(W + N, 2)
                elif pr == 4:
                    p = W + N - NW
                    g = W if abs(p - W) < abs(p - N) else N
                elif pr == 5:
                    g = max(min(W, N), min(max(W, N), W + N - NW))
                elif pr == 6:
                    g = wpred
                elif pr == 7:
                    g = NE
                elif pr == 8:
                    g = NW
                elif pr == 9:
                    g = WW
                elif pr == 10:
                    g = _tdiv(W + NW, 2)
                elif pr == 11:
                    g = _tdiv(N + NW, 2)
                elif pr == 12:
                    g = _tdiv(N + NE, 2)
                else:
                    g = _tdiv(6 * N - 2 * NN + 7 * W + WW + NEE + 3 * NE + 8, 16)
                v = code.read(br, n[2], dm)
                val = unpack_signed(v) * n[6] + n[5] + g
                row[x] = val
                if wps:
                    wps.update(val, x, y)
        out.append(np.array(img, dtype=np.int64).reshape(h, w))
    assert code.final_ok(), 'modular final state'
    info['decoded'] = len(out)
    while len(out) < len(chans):
        out.append(None)
    return out, info


def undo_rct(chn, t):
    """Inverse RCT (App. B.6) in place on the list of channel arrays."""
    bc = t['begin_c']
    ty = t['rct_type']
    perm = ty // 7
    k = ty % 7
    F, S, T = chn[bc], chn[bc + 1], chn[bc + 2]
    if k == 6:
        tmp = F - (T >> 1)
        G = T + tmp
        Bq = tmp - (S >> 1)
        Rq = Bq + S
        F, S, T = Rq, G, Bq
    else:
        if k & 1:
            T = T + F
        if (k >> 1) == 1:
            S = S + F
        elif (k >> 1) == 2:
            S = S + ((F + T) >> 1)
    v0 = bc + perm % 3
    v1 = bc + (perm + 1 + perm // 3) % 3
    v2 = bc + (perm + 2 - perm // 3) % 3
    chn[v0], chn[v1], chn[v2] = F, S, T


def undo_transforms(chn, transforms):
    for t in reversed(transforms):
        assert t['id'] == 0
        undo_rct(chn, t)
