"""TEST INFRASTRUCTURE (oracle) — whole-frame JPEG XL decode to RGBA8 with every intermediate stage exposed, i.e. the
CPU restatement of what libjxl 0.12.0 computes inside the reference's DecodeJpegXlOneShot
(/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:36-176: {4, JXL_TYPE_UINT8} interleaved output,
alpha not premultiplied).  The per-stage outputs are what tests/ compare the CUDA kernels against; the restatement
itself is pinned against the reference's own binary (oracle/_ref) in tests/test_oracle_pinned.py.

Scope: single-pass VarDCT frames (flags 0 / 0x80) with the strategies in vardct.SUPPORTED_STRATEGIES, 8-bit alpha;
modular (lossless) frames with RCT only.  Pure-Python entropy loops: small images only.
"""
import os
import numpy as np
from . import headers as hd
from . import modular as mod
from . import vardct as vd
from .entropy import BitReader

_DITHER = None


def dither_table():
    global _DITHER
    if _DITHER is None:
        here = os.path.dirname(os.path.abspath(__file__))
        for p in (os.path.join(here, '..', '_ref', 'dither_table.bin'), os.path.join(here, '..', '..', 'jxl_coder_b200', 'csrc', 'dither_table.bin')):
            if os.path.exists(p):
                _DITHER = np.fromfile(p, dtype='<f4').reshape(32, 32).astype(np.float64)
                break
        else:
            raise FileNotFoundError('dither_table.bin')
    return _DITHER


def _modular_channels(md, fh):
    """(w, h) of every channel of the frame's modular image (colour channels only for modular encoding)."""
    W, H = fh['cw'], fh['ch']
    ch = []
    if fh['encoding'] == 1:
        ncol = 1 if md['cs'] == 1 else 3
        ch += [(W, H)] * ncol
    for ec in md['extra']:
        assert ec['dim_shift'] == 0
        ch.append((W, H))
    return ch


def decode_frame(cs, br, md, stages=None):
    """Decodes one frame starting at br.  Returns dict(rgba=uint8 [H,W,4], fh=..., + stages)."""
    st = {} if stages is None else stages
    fh = hd.parse_frame_header(br, md)
    st['fh'] = fh
    W, H = fh['cw'], fh['ch']
    assert fh['upsampling'] == 1 and fh['num_passes'] == 1 and fh['frame_type'] == 0
    nlf, ng = fh['num_lf_groups'], fh['num_groups']
    single = fh['toc_entries'] == 1
    g, gtree, gcode, gbr = vd.lf_global(cs, fh, md)
    st['lf_global'] = g
    # ---- global modular stream
    mchans = _modular_channels(md, fh)
    gdim = fh['group_dim']
    mimg = [np.zeros((h, w), np.int64) for (w, h) in mchans]
    ginfo = {'transforms': []}
    global_done = [False] * len(mchans)
    if mchans:
        part, ginfo = mod.decode_channels(gbr, mchans, 0, gtree, gcode, max_chan_size=gdim)
        for i in range(ginfo['decoded']):
            mimg[i] = part[i]
            global_done[i] = True
    st['global_modular'] = ginfo
    vardct = fh['encoding'] == 0
    if vardct:
        w8t, h8t = -(-W // 8), -(-H // 8)
        lfq_full = [np.zeros((h8t, w8t), np.int64) for _ in range(3)]
        dc_full = [np.zeros((h8t, w8t)) for _ in range(3)]
        strategy = np.full((h8t, w8t), -1, np.int32)      # strategy id per covered cell
        first = np.zeros((h8t, w8t), bool)                 # top-left cell of a block
        hfmul = np.zeros((h8t, w8t), np.int32)
        sharp = np.zeros((h8t, w8t), np.int32)
        xfromy = np.zeros((-(-H // 64), -(-W // 64)), np.int32)
        bfromy = np.zeros((-(-H // 64), -(-W // 64)), np.int32)
        lfos, bms, rects = [], [], []
        sbr = gbr if single else None
        for l in range(nlf):
            rect = vd.lf_group_rect(fh, l)
            lfo = vd.lf_group(cs, fh, g, gtree, gcode, l, br=sbr)
            x0, y0, w, h = rect
            w8, h8 = -(-w // 8), -(-h // 8)
            bm = vd.BlockMap(lfo, w8, h8)
            cx0, cy0 = x0 // 8, y0 // 8
            for c in range(3):
                lfq_full[c][cy0:cy0 + h8, cx0:cx0 + w8] = lfo['lf'][c]
            dc, mul = vd.lf_dequant(g, lfo)
            for c in range(3):
                dc_full[c][cy0:cy0 + h8, cx0:cx0 + w8] = dc[c]
            strategy[cy0:cy0 + h8, cx0:cx0 + w8] = bm.cov
            for (by, bx), (t, q) in bm.first.items():
                first[cy0 + by, cx0 + bx] = True
                hfmul[cy0 + by:cy0 + by + vd.CBY[t], cx0 + bx:cx0 + bx + vd.CBX[t]] = q
            sharp[cy0:cy0 + h8, cx0:cx0 + w8] = lfo['sharpness']
            xfromy[y0 // 64:y0 // 64 + lfo['xfromy'].shape[0], x0 // 64:x0 // 64 + lfo['xfromy'].shape[1]] = lfo['xfromy']
            bfromy[y0 // 64:y0 // 64 + lfo['bfromy'].shape[0], x0 // 64:x0 // 64 + lfo['bfromy'].shape[1]] = lfo['bfromy']
            lfos.append(lfo)
            bms.append(bm)
            rects.append(rect)
        st.update(lf_quant=lfq_full, strategy=strategy, first=first, hf_mul=hfmul, sharpness=sharp, xfromy=xfromy, bfromy=bfromy,
                  lf_mul=mul, lf_groups=lfos)
        hfo, accode = vd.hf_global(cs, fh, g, br=sbr)
        st['hf_global'] = hfo
        # adaptive LF smoothing is done on the whole frame's LF image
        if not (fh['flags'] & 0x80):
            dcs = vd.adaptive_lf_smooth(dc_full, mul)
        else:
            dcs = dc_full
        st['lf_dequant'] = dc_full
        st['lf_smooth'] = dcs
        # ---- AC
        coef_list = []  # (cell y, cell x (frame coords), c, k, v)
        for gi in range(ng):
            gx, gy = gi % fh['ngx'], gi // fh['ngx']
            l = (gy // 8) * fh['nlfx'] + (gx // 8)
            toks, pbr = vd.pass_group(cs, fh, g, bms[l], rects[l], lfos[l]['lf'], hfo, accode, gi, br=sbr)
            cx0, cy0 = rects[l][0] // 8, rects[l][1] // 8
            coef_list += [(cy0 + by, cx0 + bx, c, k, v) for (by, bx, c, k, v) in toks]
            _decode_group_modular(pbr, fh, md, mchans, global_done, mimg, gtree, gcode, gi, st)
        st['coef_list'] = coef_list
        planes = _reconstruct(fh, g, hfo, strategy, first, hfmul, xfromy, bfromy, dcs, coef_list, W, H, st)
        PW, PH = planes[0].shape[1], planes[0].shape[0]
        st['xyb_idct'] = [p[:H, :W].copy() for p in planes]
        pl = [p[:H, :W] for p in planes]
        if fh['gab']:
            pl = vd.gaborish(pl, fh['gab_w'])
        if fh['epf_iters']:
            inv_sigma = vd.epf_inv_sigma(g, fh, hfmul, sharp)
            pl = vd.epf(pl, inv_sigma, fh, W, H)
        st['xyb_final'] = pl
        lin = vd.xyb_to_linear(pl, md['intensity_target'])
        if md['prim'] != 1 and md['prim'] in vd.PRIMARIES:
            Ms = vd.rgb_to_xyz_matrix(vd.PRIMARIES[1], vd.D65)
            Mt = vd.rgb_to_xyz_matrix(vd.PRIMARIES[md['prim']], vd.D65)
            lin = lin @ (np.linalg.inv(Mt) @ Ms).T
        st['linear_rgb'] = lin
        if md['tf'] == 13:
            enc = vd.srgb_oetf(lin)
        elif md['tf'] == 8:
            enc = np.clip(lin, 0, 1)
        else:
            raise NotImplementedError('transfer function %r in oracle' % md['tf'])
        rgb = vd.to_u8(enc, dither_table())
    else:
        for gi in range(ng):
            gbr2 = gbr if single else BitReader(cs, fh['sec_offs'][1 + nlf + 1 + gi] * 8)
            if not single:
                _decode_group_modular(gbr2, fh, md, mchans, global_done, mimg, gtree, gcode, gi, st)
        mod.undo_transforms(mimg, ginfo['transforms'])
        ncol = 1 if md['cs'] == 1 else 3
        bits = md['bit_depth']['bits']
        col = [np.clip(mimg[i], 0, (1 << bits) - 1) for i in range(ncol)]
        if ncol == 1:
            col = col * 3
        assert bits == 8
        rgb = np.stack(col, -1).astype(np.uint8)
    rgba = np.full((H, W, 4), 255, np.uint8)
    rgba[:, :, :3] = rgb
    nmod_col = len(mchans) - len(md['extra'])
    for i, ec in enumerate(md['extra']):
        if ec['type'] == 0:
            assert ec['bit_depth']['bits'] == 8
            rgba[:, :, 3] = np.clip(mimg[nmod_col + i], 0, 255).astype(np.uint8)
            break
    st['modular_image'] = mimg
    return dict(rgba=rgba, fh=fh, stages=st)


def _decode_group_modular(pbr, fh, md, mchans, global_done, mimg, gtree, gcode, gi, st):
    """Per-group modular data at the end of a PassGroup section (stream id 1+3nlf+17+g)."""
    todo = [i for i in range(len(mchans)) if not global_done[i]]
    if not todo:
        return
    gdim = fh['group_dim']
    gx, gy = gi % fh['ngx'], gi // fh['ngx']
    x0, y0 = gx * gdim, gy * gdim
    chans = []
    for i in todo:
        w, h = mchans[i]
        chans.append((max(0, min(gdim, w - x0)), max(0, min(gdim, h - y0))))
    nlf = fh['num_lf_groups']
    part, info = mod.decode_channels(pbr, chans, 1 + 3 * nlf + 17 + gi, gtree, gcode)
    mod.undo_transforms(part, info['transforms'])
    st.setdefault('group_modular', {})[gi] = info
    for j, i in enumerate(todo):
        w, h = chans[j]
        mimg[i][y0:y0 + h, x0:x0 + w] = part[j]


def coefficient_planes(fh, hfo, strategy, first, coef_list, W, H):
    """int planes [3][H8*8, W8*8] (c: 0=X 1=Y 2=B) in the layout DESIGN.md defines for the GPU path: every varblock's
    pixel rectangle holds its quantised coefficient array; square and tall blocks store it transposed, so the plane
    holds F[v][u] (vertical, horizontal frequency) at (row v, col u) for every block."""
    h8, w8 = strategy.shape
    planes = np.zeros((3, h8 * 8, w8 * 8), np.int32)
    orders = hfo['orders']
    for (by, bx, c, k, v) in coef_list:
        t = int(strategy[by, bx])
        cx, cy = vd.CBX[t], vd.CBY[t]
        kc = 8 * max(cx, cy)
        pos = orders[(vd.ORDER_ID[t], c)][k]
        r, col = pos // kc, pos % kc
        if cy >= cx:
            r, col = col, r
        planes[c, by * 8 + r, bx * 8 + col] += v
    return planes


def _reconstruct(fh, g, hfo, strategy, first, hfmul, xfromy, bfromy, dcs, coef_list, W, H, st):
    h8, w8 = strategy.shape
    orders = hfo['orders']
    blocks = {}
    for (by, bx, c, k, v) in coef_list:
        t = int(strategy[by, bx])
        key = (by, bx)
        if key not in blocks:
            blocks[key] = [np.zeros(64 * vd.CBX[t] * vd.CBY[t]) for _ in range(3)]
        blocks[key][c][orders[(vd.ORDER_ID[t], c)][k]] += v
    inv_gs = 65536.0 / g['global_scale']
    xm = 0.8 ** (fh['x_qm_scale'] - 2)
    bm_ = 0.8 ** (fh['b_qm_scale'] - 2)
    cf = g['cfl']
    planes = [np.zeros((h8 * 8, w8 * 8)) for _ in range(3)]
    ys, xs = np.nonzero(first)
    for by, bx in zip(ys.tolist(), xs.tolist()):
        t = int(strategy[by, bx])
        q = int(hfmul[by, bx])
        cx, cy = vd.CBX[t], vd.CBY[t]
        kr, kc = 8 * min(cx, cy), 8 * max(cx, cy)
        co = blocks.get((by, bx)) or [np.zeros(kr * kc) for _ in range(3)]
        sd = inv_gs / q
        K = [vd.adjust_bias(co[c], c).reshape(kr, kc) * sd * vd.dequant_matrix(t, c) for c in range(3)]
        K[0] = K[0] * xm
        K[2] = K[2] * bm_
        kx = cf['base_x'] + xfromy[by // 8, bx // 8] / cf['colour_factor']
        kb = cf['base_b'] + bfromy[by // 8, bx // 8] / cf['colour_factor']
        K[0] = K[0] + kx * K[1]
        K[2] = K[2] + kb * K[1]
        for c in range(3):
            planes[c][by * 8:by * 8 + 8 * cy, bx * 8:bx * 8 + 8 * cx] = vd.idct_block(t, K[c], dcs[c], by, bx)
    return planes


def decode(data, want_stages=True):
    """Decodes the (last) frame of a still image.  Returns dict(rgba, md, fh, stages)."""
    cs = hd.extract_codestream(data)
    br, md = hd.parse_image_header(cs)
    res = None
    while True:
        st = {}
        res = decode_frame(cs, br, md, st)
        if res['fh']['is_last']:
            break
    res['md'] = md
    res['codestream'] = cs
    return res


def basic_info(data):
    cs = hd.extract_codestream(data)
    _, md = hd.parse_image_header(cs)
    return md
