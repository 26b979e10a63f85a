"""TEST INFRASTRUCTURE (oracle) — container, image/frame headers and TOC of a JPEG XL codestream, as consumed by
libjxl 0.12.0 behind the reference's DecodeJpegXlOneShot / DecodeBasicInfo
(/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:36-226).  Restates SURVEY.md App. B.1-B.3.
"""
import math
from .entropy import BitReader, Code, unpack_signed, read_permutation

SIG_CONTAINER = bytes([0, 0, 0, 0xC, 0x4A, 0x58, 0x4C, 0x20, 0xD, 0xA, 0x87, 0xA])


def extract_codestream(data):
    """Bare codestream (FF 0A) or ISOBMFF container with jxlc / jxlp boxes (App. B.1; cf. JxlCoder.kt:244-258)."""
    if data[:2] == b'\xff\x0a':
        return bytes(data)
    if data[:12] != SIG_CONTAINER:
        raise ValueError('not a JXL file')
    pos = 0
    parts = []
    while pos + 8 <= len(data):
        size = int.from_bytes(data[pos:pos + 4], 'big')
        typ = data[pos + 4:pos + 8]
        hdr = 8
        if size == 1:
            size = int.from_bytes(data[pos + 8:pos + 16], 'big')
            hdr = 16
        elif size == 0:
            size = len(data) - pos
        body = data[pos + hdr:pos + size]
        if typ == b'jxlc':
            parts.append(body)
        elif typ == b'jxlp':
            parts.append(body[4:])
        pos += size
    if not parts:
        raise ValueError('no codestream box')
    return b''.join(parts)


def _size_header(br):
    small = br.u(1)

    def dim():
        return (br.u(5) + 1) * 8 if small else br.U32((1, 9), (1, 13), (1, 18), (1, 30))
    h = dim()
    ratio = br.u(3)
    w = dim() if ratio == 0 else {1: h, 2: h * 12 // 10, 3: h * 4 // 3, 4: h * 3 // 2, 5: h * 16 // 9, 6: h * 5 // 4, 7: h * 2}[ratio]
    return w, h


def _bit_depth(br):
    if not br.Bool():
        return dict(float=False, bits=br.U32(8, 10, 12, (1, 6)), exp_bits=0)
    return dict(float=True, bits=br.U32(32, 16, 24, (1, 6)), exp_bits=1 + br.u(4))


def parse_image_header(cs):
    assert cs[:2] == b'\xff\x0a'
    br = BitReader(cs, 16)
    w, h = _size_header(br)
    md = dict(w=w, h=h, orientation=1, bit_depth=dict(float=False, bits=8, exp_bits=0), mod16=True, extra=[], xyb=True,
              have_animation=False, cs=0, wp=1, prim=1, tf=13, gamma=None, want_icc=False, intent=1,
              intensity_target=255.0, min_nits=0.0, linear_below=0.0, tps_num=0, tps_den=0, loops=0, timecodes=False)
    all_default = br.Bool()
    extra_fields = False
    if not all_default:
        extra_fields = br.Bool()
        if extra_fields:
            md['orientation'] = 1 + br.u(3)
            if br.Bool():
                md['intrinsic'] = _size_header(br)
            if br.Bool():  # preview header
                div8 = br.Bool()
                if div8:
                    ph = br.U32(16, 32, (1, 5), (33, 9)) * 8
                else:
                    ph = br.U32((1, 6), (65, 8), (321, 10), (1345, 12))
                pr = br.u(3)
                if pr == 0:
                    if div8:
                        br.U32(16, 32, (1, 5), (33, 9))
                    else:
                        br.U32((1, 6), (65, 8), (321, 10), (1345, 12))
                md['have_preview'] = True
            md['have_animation'] = bool(br.Bool())
            if md['have_animation']:
                md['tps_num'] = br.U32(100, 1000, (1, 10), (1, 30))
                md['tps_den'] = br.U32(1, 1001, (1, 8), (1, 10))
                md['loops'] = br.U32(0, (0, 3), (0, 16), (0, 32))
                md['timecodes'] = bool(br.Bool())
        md['bit_depth'] = _bit_depth(br)
        md['mod16'] = bool(br.Bool())
        nextra = br.U32(0, 1, (2, 4), (1, 12))
        for _ in range(nextra):
            ec = dict(type=0, bit_depth=dict(float=False, bits=8, exp_bits=0), dim_shift=0, premul=False)
            if not br.Bool():
                ec['type'] = br.Enum()
                ec['bit_depth'] = _bit_depth(br)
                ec['dim_shift'] = br.U32(0, 3, 4, (1, 3))
                nl = br.U32(0, (0, 4), (16, 5), (48, 10))
                br.p += 8 * nl
                if ec['type'] == 0:
                    ec['premul'] = bool(br.Bool())
                elif ec['type'] == 2:  # spot colour
                    for _ in range(4):
                        br.F16()
                elif ec['type'] == 5:  # CFA
                    br.U32(1, (0, 2), (3, 4), (19, 8))
            md['extra'].append(ec)
        md['xyb'] = bool(br.Bool())
        if not br.Bool():  # colour encoding not all_default
            md['want_icc'] = bool(br.Bool())
            md['cs'] = br.Enum()
            if not md['want_icc']:
                if md['cs'] != 2:
                    md['wp'] = br.Enum()
                    if md['wp'] == 2:
                        md['wp_xy'] = [_custom_xy(br)]
                if md['cs'] not in (1, 2):
                    md['prim'] = br.Enum()
                    if md['prim'] == 2:
                        md['prim_xy'] = [_custom_xy(br) for _ in range(3)]
                if br.Bool():
                    md['gamma'] = br.u(24)
                    md['tf'] = None
                else:
                    md['tf'] = br.Enum()
                md['intent'] = br.Enum()
        if extra_fields:
            if not br.Bool():
                md['intensity_target'] = br.F16()
                md['min_nits'] = br.F16()
                md['rel_to_max'] = bool(br.Bool())
                md['linear_below'] = br.F16()
        ext = br.U64()
        assert ext == 0, 'metadata extensions'
    md['default_m'] = bool(br.Bool())
    if not md['default_m']:
        if md['xyb']:
            if not br.Bool():
                md['opsin_inverse'] = [br.F16() for _ in range(9)]
                md['opsin_bias'] = [br.F16() for _ in range(3)]
                md['quant_bias'] = [br.F16() for _ in range(3)]
                md['quant_bias_num'] = br.F16()
        cw_mask = br.u(3)
        if cw_mask & 1:
            md['up2'] = [br.F16() for _ in range(15)]
        if cw_mask & 2:
            md['up4'] = [br.F16() for _ in range(55)]
        if cw_mask & 4:
            md['up8'] = [br.F16() for _ in range(210)]
    if md['want_icc']:
        raise NotImplementedError('ICC stream in codestream')
    br.align()
    return br, md


def _custom_xy(br):
    def c():
        return unpack_signed(br.U32((0, 19), (524288, 19), (1048576, 20), (2097152, 21)))
    return (c(), c())


_U32_DIM = ((0, 8), (256, 11), (2304, 14), (18688, 30))


def parse_frame_header(br, md):
    """Returns fh dict with section offsets (bytes, into the codestream) in logical order (App. B.2-B.3)."""
    nextra = len(md['extra'])
    xyb = md['xyb']
    fh = dict(pos_bytes=br.p // 8, frame_type=0, encoding=0, flags=0, ycbcr=False, upsampling=1, ec_upsampling=[1] * nextra,
              group_size_shift=1, x_qm_scale=3, b_qm_scale=2, num_passes=1, have_crop=False, x0=0, y0=0, width=md['w'],
              height=md['h'], blend=dict(mode=0, source=0), ec_blend=[dict(mode=0, source=0) for _ in range(nextra)], duration=0,
              is_last=True, save_as_ref=0, save_before_ct=False, gab=True, gab_w=None, epf_iters=2, epf_sharp_lut=None,
              epf_ch_scale=None, epf_quant_mul=0.46, epf_pass0=0.9, epf_pass2=6.5, epf_border=2.0 / 3.0, lf_level=0,
              jpeg_upsampling=[0, 0, 0], pass_shift=[], name_len=0)
    all_default = br.Bool()
    if not all_default:
        ft = fh['frame_type'] = br.u(2)
        enc = fh['encoding'] = br.u(1)
        flags = fh['flags'] = br.U64()
        if not xyb:
            fh['ycbcr'] = bool(br.Bool())
        if fh['ycbcr'] and not (flags & 0x20):
            fh['jpeg_upsampling'] = [br.u(2) for _ in range(3)]
        if not (flags & 0x20):
            fh['upsampling'] = br.U32(1, 2, 4, 8)
            fh['ec_upsampling'] = [br.U32(1, 2, 4, 8) for _ in range(nextra)]
        if enc == 1:
            fh['group_size_shift'] = br.u(2)
        if xyb and enc == 0:
            fh['x_qm_scale'] = br.u(3)
            fh['b_qm_scale'] = br.u(3)
        if ft != 2:
            passes = fh['num_passes'] = br.U32(1, 2, 3, (4, 3))
            if passes != 1:
                nds = br.U32(0, 1, 2, (3, 1))
                fh['pass_shift'] = [br.u(2) for _ in range(passes - 1)]
                fh['pass_ds'] = [br.U32(1, 2, 4, 8) for _ in range(nds)]
                fh['pass_last'] = [br.U32(0, 1, 2, (0, 3)) for _ in range(nds)]
        if ft == 1:
            fh['lf_level'] = 1 + br.u(2)
        else:
            fh['have_crop'] = bool(br.Bool())
            if fh['have_crop']:
                if ft != 2:
                    fh['x0'] = unpack_signed(br.U32(*_U32_DIM))
                    fh['y0'] = unpack_signed(br.U32(*_U32_DIM))
                fh['width'] = br.U32(*_U32_DIM)
                fh['height'] = br.U32(*_U32_DIM)
        normal = ft in (0, 3)
        full = (not fh['have_crop']) or (fh['x0'] <= 0 and fh['y0'] <= 0 and fh['width'] + fh['x0'] >= md['w'] and fh['height'] + fh['y0'] >= md['h'])
        if normal:
            def blend():
                bi = dict(mode=br.U32(0, 1, 2, (3, 2)), source=0, alpha_ch=0, clamp=False)
                if nextra > 0 and bi['mode'] in (2, 3):
                    bi['alpha_ch'] = br.U32(0, 1, 2, (3, 3))
                if nextra > 0 and bi['mode'] in (2, 3, 4):
                    bi['clamp'] = bool(br.Bool())
                if bi['mode'] != 0 or not full:
                    bi['source'] = br.u(2)
                return bi
            fh['blend'] = blend()
            fh['ec_blend'] = [blend() for _ in range(nextra)]
            if md['have_animation']:
                fh['duration'] = br.U32(0, 1, (0, 8), (0, 32))
                if md['timecodes']:
                    fh['timecode'] = br.u(32)
            fh['is_last'] = bool(br.Bool())
        else:
            fh['is_last'] = False
        if ft != 1 and not fh['is_last']:
            fh['save_as_ref'] = br.u(2)
        if ft != 1:
            resets = full and normal and fh['blend']['mode'] == 0
            can_ref = (not fh['is_last']) and (fh['duration'] == 0 or fh['save_as_ref'] != 0)
            if ft == 2 or (resets and can_ref):
                fh['save_before_ct'] = bool(br.Bool())
        nl = fh['name_len'] = br.U32(0, (0, 4), (16, 5), (48, 10))
        br.p += 8 * nl
        if not br.Bool():  # restoration filter
            fh['gab'] = bool(br.Bool())
            if fh['gab'] and br.Bool():
                fh['gab_w'] = [br.F16() for _ in range(6)]
            fh['epf_iters'] = br.u(2)
            if fh['epf_iters']:
                if enc == 0 and br.Bool():
                    fh['epf_sharp_lut'] = [br.F16() for _ in range(8)]
                if br.Bool():
                    fh['epf_ch_scale'] = [br.F16() for _ in range(3)]
                    br.u(32)
                if br.Bool():
                    if enc == 0:
                        fh['epf_quant_mul'] = br.F16()
                    fh['epf_pass0'] = br.F16()
                    fh['epf_pass2'] = br.F16()
                    fh['epf_border'] = br.F16()
                if enc == 1:
                    fh['epf_sigma_modular'] = br.F16()
            assert br.U64() == 0
        assert br.U64() == 0
    W, H = fh['width'], fh['height']
    ups = fh['upsampling']
    W, H = -(-W // ups), -(-H // ups)
    if fh['lf_level']:
        d = 1 << (3 * fh['lf_level'])
        W, H = -(-W // d), -(-H // d)
    fh['cw'], fh['ch'] = W, H  # coded size
    gdim = 256 if fh['encoding'] == 0 else (128 << fh['group_size_shift'])
    fh['group_dim'] = gdim
    ngx = -(-W // gdim)
    ngy = -(-H // gdim)
    nlfx = -(-W // (8 * gdim))
    nlfy = -(-H // (8 * gdim))
    ng, nlf = ngx * ngy, nlfx * nlfy
    passes = fh['num_passes']
    ntoc = 1 if (ng == 1 and passes == 1) else 1 + nlf + 1 + ng * passes
    fh.update(num_groups=ng, num_lf_groups=nlf, toc_entries=ntoc, ngx=ngx, ngy=ngy, nlfx=nlfx, nlfy=nlfy)
    perm = None
    fh['toc_permuted'] = bool(br.Bool())
    if fh['toc_permuted']:
        code = Code(br, 8)
        code.begin(br)
        perm = read_permutation(br, code, ntoc, 0)
        assert code.final_ok(), 'TOC permutation final state'
    br.align()
    sizes = [br.U32((0, 10), (1024, 14), (17408, 22), (4211712, 30)) for _ in range(ntoc)]
    br.align()
    base = br.p // 8
    offs = []
    acc = base
    for z in sizes:
        offs.append(acc)
        acc += z
    if perm:
        # logical section j lives at bitstream slot perm[j]
        offs = [offs[j] for j in perm]
        sizes = [sizes[j] for j in perm]
    fh['sec_sizes'] = sizes
    fh['sec_offs'] = offs
    fh['payload_start'] = base
    fh['end_bytes'] = acc
    br.p = acc * 8
    return fh
