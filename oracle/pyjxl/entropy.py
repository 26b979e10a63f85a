"""TEST INFRASTRUCTURE (oracle) — CPU restatement of the JPEG XL entropy decoder used by libjxl 0.12.0, the third-party
codec behind the reference's DecodeJpegXlOneShot (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:46-175).
libjxl's source is NOT in /root/reference (prebuilt .so only, build_jxl.sh:8 clones HEAD); this follows the published
format (ISO/IEC 18181-1 §C: ANS / prefix codes, hybrid-uint, LZ77, context maps) as digested in SURVEY.md App. B.4 and
was validated against the shipped binary (exact section-length consumption, ANS final state 0x130000, final pixels).
Pure-Python loops: small cases only.
"""


def ceil_log2(x):
    return 0 if x <= 1 else (x - 1).bit_length()


def floor_log2(x):
    return x.bit_length() - 1


def unpack_signed(v):
    return (v >> 1) if not (v & 1) else -((v + 1) >> 1)


class BitReader:
    def __init__(self, b, p=0):
        self.b = b
        self.p = p  # bit position
        self.n = len(b) * 8

    def u(self, n):
        v = 0
        p = self.p
        b = self.b
        for i in range(n):
            byte = b[p >> 3] if (p >> 3) < len(b) else 0
            v |= ((byte >> (p & 7)) & 1) << i
            p += 1
        self.p = p
        return v

    def peek(self, n):
        p = self.p
        v = self.u(n)
        self.p = p
        return v

    def U32(self, *d):
        x = d[self.u(2)]
        return x if isinstance(x, int) else x[0] + self.u(x[1])

    def Bool(self):
        return self.u(1)

    def U64(self):
        sel = self.u(2)
        if sel == 0:
            return 0
        if sel == 1:
            return 1 + self.u(4)
        if sel == 2:
            return 17 + self.u(8)
        v = self.u(12)
        sh = 12
        while self.u(1):
            if sh == 60:
                v |= self.u(4) << sh
                break
            v |= self.u(8) << sh
            sh += 8
        return v

    def F16(self):
        import struct
        return struct.unpack('<e', struct.pack('<H', self.u(16)))[0]

    def Enum(self):
        return self.U32(0, 1, (2, 4), (18, 6))

    def align(self):
        self.p = (self.p + 7) & ~7


# Fixed prefix code for ANS log-counts on a 7-bit peek: (nbits, symbol)   [App. B.4 step 5, B.8 pitfall 1]
_H = [(3, 10), (7, 12), (3, 7), (4, 3), (3, 6), (3, 8), (3, 9), (4, 5), (3, 10), (4, 4), (3, 7), (4, 1), (3, 6), (3, 8), (3, 9), (4, 2),
      (3, 10), (5, 0), (3, 7), (4, 3), (3, 6), (3, 8), (3, 9), (4, 5), (3, 10), (4, 4), (3, 7), (4, 1), (3, 6), (3, 8), (3, 9), (4, 2),
      (3, 10), (6, 11), (3, 7), (4, 3), (3, 6), (3, 8), (3, 9), (4, 5), (3, 10), (4, 4), (3, 7), (4, 1), (3, 6), (3, 8), (3, 9), (4, 2),
      (3, 10), (5, 0), (3, 7), (4, 3), (3, 6), (3, 8), (3, 9), (4, 5), (3, 10), (4, 4), (3, 7), (4, 1), (3, 6), (3, 8), (3, 9), (4, 2)]
LOGCOUNT_LUT = _H + [_H[0], (7, 13)] + _H[2:]


def varlen8(br):
    if br.u(1):
        n = br.u(3)
        return 1 if n == 0 else br.u(n) + (1 << n)
    return 0


def varlen16(br):
    if br.u(1):
        n = br.u(4)
        return 1 if n == 0 else br.u(n) + (1 << n)
    return 0


def read_ans_histogram(br):
    """Returns the symbol counts (sum 4096)."""
    if br.u(1):  # simple
        ns = br.u(1) + 1
        syms = [varlen8(br) for _ in range(ns)]
        c = [0] * (max(syms) + 1)
        if ns == 1:
            c[syms[0]] = 4096
        else:
            assert syms[0] != syms[1]
            c[syms[0]] = br.u(12)
            c[syms[1]] = 4096 - c[syms[0]]
        return c
    if br.u(1):  # flat
        n = varlen8(br) + 1
        c = [4096 // n] * n
        for i in range(4096 % n):
            c[i] += 1
        return c
    log = 0
    while log < 3 and br.u(1):
        log += 1
    shift = (br.u(log) | (1 << log)) - 1
    assert shift <= 13
    length = varlen8(br) + 3
    logc = [0] * length
    same = [0] * length
    omit_log = -1
    omit_pos = -1
    i = 0
    while i < length:
        nb, val = LOGCOUNT_LUT[br.peek(7)]
        br.p += nb
        logc[i] = val
        if val == 13:
            rl = varlen8(br)
            same[i] = rl + 5
            i += rl + 3 + 1
            continue
        if val > omit_log:
            omit_log = val
            omit_pos = i
        i += 1
    assert omit_pos >= 0
    c = [0] * length
    total = 0
    prev = 0
    numsame = 0
    for i in range(length):
        if same[i]:
            numsame = same[i] - 1
            prev = c[i - 1] if i > 0 else 0
        if numsame > 0:
            c[i] = prev
            numsame -= 1
        else:
            code = logc[i]
            if i == omit_pos or code == 0:
                continue
            elif code == 1:
                c[i] = 1
            else:
                lc = code - 1
                bc = min(lc, max(0, shift - ((12 - lc) >> 1)))
                c[i] = (1 << lc) + (br.u(bc) << (lc - bc))
        total += c[i]
    c[omit_pos] = 4096 - total
    assert c[omit_pos] > 0, 'bad histogram'
    return c


class AliasTable:
    """Vose-style alias table as libjxl builds it (App. B.4 step 5)."""

    def __init__(self, dist, log_alpha):
        d = list(dist)
        while d and d[-1] == 0:
            d.pop()
        if not d:
            d = [4096]
        T = 1 << log_alpha
        es = 4096 >> log_alpha
        self.log_es = 12 - log_alpha
        self.mask = es - 1
        self.right = [0] * T
        self.cut = [0] * T
        self.off1 = [0] * T
        self.f0 = [0] * T
        self.f1 = [0] * T
        for sym, v in enumerate(d):
            if v == 4096:
                for i in range(T):
                    self.right[i] = sym
                    self.cut[i] = 0
                    self.off1[i] = es * i
                    self.f0[i] = 0
                    self.f1[i] = 4096
                return
        assert len(d) <= T, (len(d), T)
        cut = [0] * T
        under = []
        over = []
        for i in range(len(d)):
            cut[i] = d[i]
            if cut[i] > es:
                over.append(i)
            elif cut[i] < es:
                under.append(i)
        for i in range(len(d), T):
            under.append(i)
        while over:
            o = over.pop()
            u_ = under.pop()
            by = es - cut[u_]
            cut[o] -= by
            self.right[u_] = o
            self.off1[u_] = cut[o]
            if cut[o] < es:
                under.append(o)
            elif cut[o] > es:
                over.append(o)
        for i in range(T):
            if cut[i] == es:
                self.right[i] = i
                self.off1[i] = 0
                self.cut[i] = 0
            else:
                self.off1[i] -= cut[i]
                self.cut[i] = cut[i]
            self.f0[i] = d[i] if i < len(d) else 0
            r = self.right[i]
            self.f1[i] = d[r] if r < len(d) else 0

    def look(self, v):
        i = v >> self.log_es
        pos = v & self.mask
        if pos >= self.cut[i]:
            return self.right[i], self.off1[i] + pos, self.f1[i]
        return i, pos, self.f0[i]


# ---- Brotli-style prefix codes (RFC 7932 §3.4/3.5)
_CL_ORDER = [1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15]
_CL_TAB = {0b0000: (2, 0), 0b0100: (2, 0), 0b1000: (2, 0), 0b1100: (2, 0), 0b0001: (2, 4), 0b0101: (2, 4), 0b1001: (2, 4),
           0b1101: (2, 4), 0b0010: (2, 3), 0b0110: (2, 3), 0b1010: (2, 3), 0b1110: (2, 3), 0b0011: (3, 2), 0b1011: (3, 2),
           0b0111: (4, 1), 0b1111: (4, 5)}


def _canonical(lengths):
    maxl = max(lengths) if lengths else 0
    code = 0
    table = {}
    for l in range(1, maxl + 1):
        for s, ls in enumerate(lengths):
            if ls == l:
                table[(l, code)] = s
                code += 1
        code <<= 1
    return table, maxl


class PrefixCode:
    def __init__(self, br, alphabet):
        self.single = None
        self.lengths = None
        if alphabet == 1:
            self.single = 0
            return
        hskip = br.u(2)
        if hskip == 1:
            mb = (alphabet - 1).bit_length()
            ns = br.u(2) + 1
            syms = [br.u(mb) for _ in range(ns)]
            lengths = [0] * alphabet
            if ns == 1:
                self.single = syms[0]
                return
            if ns == 2:
                lengths[syms[0]] = 1
                lengths[syms[1]] = 1
            elif ns == 3:
                lengths[syms[0]] = 1
                lengths[syms[1]] = 2
                lengths[syms[2]] = 2
            else:
                if br.u(1):
                    lengths[syms[0]] = 1
                    lengths[syms[1]] = 2
                    lengths[syms[2]] = 3
                    lengths[syms[3]] = 3
                else:
                    for x in syms:
                        lengths[x] = 2
            self.lengths = lengths
            self.table, self.maxl = _canonical(lengths)
            return
        cl = [0] * 18
        space = 32
        num = 0
        for i in range(hskip, 18):
            nb, v = _CL_TAB[br.peek(4)]
            br.p += nb
            cl[_CL_ORDER[i]] = v
            if v:
                space -= 32 >> v
                num += 1
            if space <= 0:
                break
        assert num == 1 or space == 0, ('clcl', space, num)
        cltab, clmax = _canonical(cl)

        def rd(tab, maxl):
            code = 0
            for l in range(1, maxl + 1):
                code = (code << 1) | br.u(1)
                if (l, code) in tab:
                    return tab[(l, code)]
            raise ValueError('bad prefix')

        lengths = [0] * alphabet
        i = 0
        prev = 8
        rep = 0
        rep_len = 0
        space = 32768
        single_cl = [k for k, v in enumerate(cl) if v] if num == 1 else None
        while i < alphabet and space > 0:
            c = single_cl[0] if single_cl else rd(cltab, clmax)
            if c < 16:
                rep = 0
                lengths[i] = c
                i += 1
                if c:
                    prev = c
                    space -= 32768 >> c
            else:
                extra = c - 14
                new_len = prev if c == 16 else 0
                if rep_len != new_len:
                    rep = 0
                    rep_len = new_len
                old = rep
                if rep > 0:
                    rep = (rep - 2) << extra
                rep += br.u(extra) + 3
                delta = rep - old
                assert i + delta <= alphabet
                for _ in range(delta):
                    lengths[i] = rep_len
                    i += 1
                if rep_len:
                    space -= delta << (15 - rep_len)
        assert space == 0, ('space', space)
        self.lengths = lengths
        self.table, self.maxl = _canonical(lengths)

    def read(self, br):
        if self.single is not None:
            return self.single
        code = 0
        for l in range(1, self.maxl + 1):
            code = (code << 1) | br.u(1)
            k = (l, code)
            if k in self.table:
                return self.table[k]
        raise ValueError('bad prefix symbol')


class HybridUint:
    def __init__(self, br, log_alpha):
        self.split_exp = br.u(ceil_log2(log_alpha + 1))
        self.msb = self.lsb = 0
        if self.split_exp != log_alpha:
            self.msb = br.u(ceil_log2(self.split_exp + 1))
            assert self.msb <= self.split_exp
            self.lsb = br.u(ceil_log2(self.split_exp - self.msb + 1))
            assert self.msb + self.lsb <= self.split_exp
        self.split = 1 << self.split_exp

    def val(self, tok, br):
        if tok < self.split:
            return tok
        nb = self.split_exp - (self.msb + self.lsb) + ((tok - self.split) >> (self.msb + self.lsb))
        low = tok & ((1 << self.lsb) - 1)
        tok >>= self.lsb
        bits = br.u(nb)
        return (((((1 << self.msb) | (tok & ((1 << self.msb) - 1))) << nb) | bits) << self.lsb) | low


# LZ77 special distances for streams with a 2-D distance multiplier (modular channel data) [spec Table; M in survey]
_SPECIAL = [(0, 1), (1, 0), (1, 1), (-1, 1), (0, 2), (2, 0), (1, 2), (-1, 2), (2, 1), (-2, 1), (2, 2), (-2, 2), (0, 3), (3, 0),
            (1, 3), (-1, 3), (3, 1), (-3, 1), (2, 3), (-2, 3), (3, 2), (-3, 2), (0, 4), (4, 0), (1, 4), (-1, 4), (4, 1), (-4, 1),
            (3, 3), (-3, 3), (2, 4), (-2, 4), (4, 2), (-4, 2), (0, 5), (3, 4), (-3, 4), (4, 3), (-4, 3), (5, 0), (1, 5), (-1, 5),
            (5, 1), (-5, 1), (2, 5), (-2, 5), (5, 2), (-5, 2), (4, 4), (-4, 4), (3, 5), (-3, 5), (5, 3), (-5, 3), (0, 6), (6, 0),
            (1, 6), (-1, 6), (6, 1), (-6, 1), (2, 6), (-2, 6), (6, 2), (-6, 2), (4, 5), (-4, 5), (5, 4), (-5, 4), (3, 6), (-3, 6),
            (6, 3), (-6, 3), (0, 7), (7, 0), (1, 7), (-1, 7), (5, 5), (-5, 5), (7, 1), (-7, 1), (4, 6), (-4, 6), (6, 4), (-6, 4),
            (2, 7), (-2, 7), (7, 2), (-7, 2), (3, 7), (-3, 7), (7, 3), (-7, 3), (5, 6), (-5, 6), (6, 5), (-6, 5), (8, 0), (4, 7),
            (-4, 7), (7, 4), (-7, 4), (8, 1), (8, 2), (6, 6), (-6, 6), (8, 3), (5, 7), (-5, 7), (7, 5), (-7, 5), (8, 4), (6, 7),
            (-6, 7), (7, 6), (-7, 6), (8, 5), (7, 7), (-7, 7), (8, 6), (8, 7)]


class Code:
    """An entropy code over `nctx` contexts (App. B.4 steps 1-8)."""

    def __init__(self, br, nctx, allow_lz77=True):
        self.lz = br.u(1)
        if self.lz:
            assert allow_lz77
            self.min_symbol = br.U32(224, 512, 4096, (8, 15))
            self.min_length = br.U32(3, 4, (5, 2), (9, 8))
            self.lz_cfg = HybridUint(br, 8)
            nctx += 1
        self.nctx = nctx
        self.ctx_map = [0] * nctx
        if nctx > 1:
            self.ctx_map = read_context_map(br, nctx)
        nh = max(self.ctx_map) + 1
        self.nh = nh
        self.prefix = br.u(1)
        self.log_alpha = 15 if self.prefix else 5 + br.u(2)
        self.cfg = [HybridUint(br, self.log_alpha) for _ in range(nh)]
        if self.prefix:
            sizes = [varlen16(br) + 1 for _ in range(nh)]
            self.codes = [PrefixCode(br, a) for a in sizes]
        else:
            self.hists = [read_ans_histogram(br) for _ in range(nh)]
            self.alias = [AliasTable(h, self.log_alpha) for h in self.hists]
        self.state = None
        self.win = None
        self.ncopy = 0

    def begin(self, br):
        if not self.prefix:
            self.state = br.u(32)
        if self.lz:
            self.win = [0] * (1 << 20)
            self.nd = 0
            self.ncopy = 0
            self.cpos = 0

    def sym(self, br, h):
        if self.prefix:
            return self.codes[h].read(br)
        sy, off, fr = self.alias[h].look(self.state & 0xfff)
        self.state = fr * (self.state >> 12) + off
        if self.state < (1 << 16):
            self.state = (self.state << 16) | br.u(16)
        return sy

    def read(self, br, ctx, dist_mult=0):
        if self.lz and self.ncopy > 0:
            v = self.win[self.cpos & 0xfffff]
            self.cpos += 1
            self.ncopy -= 1
            self.win[self.nd & 0xfffff] = v
            self.nd += 1
            return v
        h = self.ctx_map[ctx]
        tok = self.sym(br, h)
        if self.lz and tok >= self.min_symbol:
            n = self.lz_cfg.val(tok - self.min_symbol, br) + self.min_length
            dh = self.ctx_map[self.nctx - 1]
            dt = self.sym(br, dh)
            d = self.cfg[dh].val(dt, br)
            if dist_mult == 0:
                d += 1
            elif d >= 120:
                d -= 119
            else:
                dx, dy = _SPECIAL[d]
                d = max(1, dx + dist_mult * dy)
            d = min(d, self.nd, 1 << 20)
            self.cpos = self.nd - d
            self.ncopy = n
            return self.read(br, ctx, dist_mult)
        v = self.cfg[h].val(tok, br)
        if self.lz:
            self.win[self.nd & 0xfffff] = v
            self.nd += 1
        return v

    def final_ok(self):
        return self.prefix or self.state == 0x130000


def read_context_map(br, n):
    if br.u(1):
        b = br.u(2)
        return [br.u(b) if b else 0 for _ in range(n)]
    mtf = br.u(1)
    c = Code(br, 1, allow_lz77=(n > 2))
    c.begin(br)
    m = [c.read(br, 0) for _ in range(n)]
    assert c.final_ok(), 'context map final state'
    if mtf:
        l = list(range(256))
        for i, v in enumerate(m):
            x = l[v]
            m[i] = x
            if v:
                del l[v]
                l.insert(0, x)
    assert max(m) < 256
    return m


def read_permutation(br, code, size, skip):
    """Lehmer-coded permutation (App. B.3 / B.7): returns perm list of `size`."""
    def cctx(v):
        return 0 if v == 0 else min(7, v.bit_length())
    end = code.read(br, cctx(size)) + skip
    assert end <= size
    leh = [0] * size
    last = 0
    for i in range(skip, end):
        leh[i] = code.read(br, cctx(last))
        last = leh[i]
        assert leh[i] < size - i
    tmp = list(range(size))
    perm = []
    for i in range(size):
        perm.append(tmp.pop(leh[i]))
    return perm
