"""Closed-form numpy model of the reference's pixel-reconstruction path.

TEST INFRASTRUCTURE.  This is a second, independently derived statement of what
`worker::post_process` (reference src/worker.rs:32-251) computes for a whole image: where oracle/zj_oracle.c
transliterates the reference line by line (iterators, intrinsics and all), this file writes down the *closed
forms* those lines reduce to, the same closed forms the CUDA kernels implement.
tests/test_oracle_vs_model.py requires the two to agree bit-for-bit on random data; a disagreement means one
of the two readings of the reference is wrong.

All arithmetic is wrapping at the reference's operand width (numpy int32/int16 arrays wrap silently).
"""
from __future__ import annotations

import numpy as np

X86, SCALAR = 0, 1
CS_RGB, CS_GRAY, CS_YCBCR, CS_CMYK, CS_YCCK, CS_RGBA, CS_RGBX = range(7)
NCOMP = {CS_RGB: 3, CS_YCBCR: 3, CS_GRAY: 1, CS_CMYK: 4, CS_YCCK: 4, CS_RGBA: 4, CS_RGBX: 4}

SCALE_BITS = np.int32(512 + 65536 + (128 << 17))


class RefPanic(Exception):
    """The reference would panic on this input."""


# --------------------------------------------------------------------------- IDCT
def _kernel_1d(s, bias, sh):
    """1-D 8-point kernel, reference src/idct/scalar.rs:79-166 == src/idct/avx2.rs:251-331.
    s: int32 array [..., 8] (transform along the last axis)."""
    i32 = np.int32
    s0, s1, s2, s3, s4, s5, s6, s7 = (s[..., k] for k in range(8))
    p1 = (s2 + s6) * i32(2217)
    t2 = p1 + s6 * i32(-7567)
    t3 = p1 + s2 * i32(3135)
    t0 = (s0 + s4) << i32(12)
    t1 = (s0 - s4) << i32(12)
    x0 = t0 + t3 + bias
    x3 = t0 - t3 + bias
    x1 = t1 + t2 + bias
    x2 = t1 - t2 + bias
    a, b, c, d = s7, s5, s3, s1
    p3 = a + c
    p4 = b + d
    q1 = a + d
    q2 = b + c
    p5 = (p3 + p4) * i32(4816)
    a = a * i32(1223)
    b = b * i32(8410)
    c = c * i32(12586)
    d = d * i32(6149)
    q1 = p5 + q1 * i32(-3685)
    q2 = p5 + q2 * i32(-10497)
    p3 = p3 * i32(-8034)
    p4 = p4 * i32(-1597)
    d = d + q1 + p4
    c = c + q2 + p3
    b = b + q2 + p4
    a = a + q1 + p3
    out = np.stack([x0 + d, x1 + c, x2 + b, x3 + a, x3 - a, x2 - b, x1 - c, x0 - d], axis=-1)
    return out >> i32(sh)


def idct_blocks(coef: np.ndarray, qt: np.ndarray, variant: int) -> np.ndarray:
    """coef int16 [N,64] natural order, qt int32 [64] -> int16 [N,8,8] samples."""
    with np.errstate(over="ignore"):
        n = coef.shape[0]
        c = coef.reshape(n, 8, 8)
        d = c.astype(np.int32) * qt.reshape(8, 8).astype(np.int32)
        if variant == X86:  # rows first, then columns (Q2; avx2.rs:333-347)
            a = _kernel_1d(d, np.int32(512), 10)                      # along rows
            b = _kernel_1d(a.transpose(0, 2, 1), SCALE_BITS, 17)      # down columns -> [N, col, row]
            full = b.transpose(0, 2, 1)
        else:  # columns first, then rows (scalar.rs:79-274)
            a = _kernel_1d(d.transpose(0, 2, 1), np.int32(512), 10)   # [N, col, k] -> [N, col, row_out]
            a = a.transpose(0, 2, 1)                                  # [N, row, col]
            full = _kernel_1d(a, SCALE_BITS, 17)
        full = np.clip(full, 0, 255).astype(np.int16)
        # DC-only shortcut (Q3): i16 wrapping, clamped only on X86 (avx2.rs:159-167 / scalar.rs:45-48)
        ac_zero = ~np.any(coef[:, 1:] != 0, axis=1)
        dc = (coef[:, 0].astype(np.int16) * qt[0].astype(np.int16)).astype(np.int16)  # wrapping i16 mul
        dcv = ((dc >> np.int16(3)) + np.int16(128)).astype(np.int16)
        if variant == X86:
            dcv = np.clip(dcv, 0, 255).astype(np.int16)
        full[ac_zero] = dcv[ac_zero][:, None, None]
        return full


def plane_from_blocks(samples: np.ndarray, blocks_per_row: int) -> np.ndarray:
    """[N,8,8] raster blocks -> int16 plane [N/bpr*8, bpr*8]."""
    n = samples.shape[0]
    rows = n // blocks_per_row
    return samples[: rows * blocks_per_row].reshape(rows, blocks_per_row, 8, 8).transpose(0, 2, 1, 3).reshape(rows * 8, blocks_per_row * 8)


# --------------------------------------------------------------------------- up-samplers (flat strip arrays)
def _T(a, b):
    """(3a + b + 2) >> 2 in wrapping i16"""
    with np.errstate(over="ignore"):
        return ((np.int16(3) * a.astype(np.int16) + b.astype(np.int16) + np.int16(2)).astype(np.int16)) >> np.int16(2)


def up_h_scalar(x: np.ndarray) -> np.ndarray:
    """upsampler/scalar.rs:5-60 closed form (SURVEY A.4 'H, SCALAR')"""
    n = x.size
    if n <= 2:
        raise RefPanic
    out = np.zeros(2 * n, np.int16)
    i = np.arange(1, n - 1)
    out[2 * i] = _T(x[i], x[i - 1])
    out[2 * i + 1] = _T(x[i], x[i + 1])
    out[0] = x[0]
    out[1] = _T(x[0:1], x[1:2])[0]
    out[2 * n - 2] = _T(x[n - 2 : n - 1], x[n - 1 : n])[0]  # sic
    out[2 * n - 1] = x[n - 1]
    return out


def up_h_sse(x: np.ndarray) -> np.ndarray:
    """upsampler/sse.rs:24-134 closed form (Q4b)"""
    n = x.size
    if n <= 5:
        raise RefPanic
    out = np.zeros(2 * n, np.int16)
    i = np.arange(1, n - 1)
    out[2 * i] = _T(x[i], x[i - 1])
    out[2 * i + 1] = _T(x[i], x[i + 1])
    out[0] = x[0]
    out[1] = _T(x[0:1], x[1:2])[0]
    il = n - 4
    t = lambda a, b: _T(x[a : a + 1], x[b : b + 1])[0]
    tail = [t(il, il - 1), t(il, il + 1), t(il + 1, il), x[il + 1], x[il + 2], t(il + 2, il + 1), t(il + 2, il + 3), x[il + 3]]
    out[2 * n - 8 :] = np.array(tail, np.int16)
    return out


def up_v(x: np.ndarray) -> np.ndarray:
    """upsampler/scalar.rs:64-147 closed form: 8 rows in, 16 rows out (Q4c)"""
    n = x.size
    w = n >> 3
    if w == 0:
        raise RefPanic
    r = x[: 8 * w].reshape(8, w)
    out = np.zeros((16, w), np.int16)
    out[0] = r[0]
    out[1] = r[0]
    for k in range(1, 7):
        out[2 * k] = _T(r[k], r[k + 1])
        out[2 * k + 1] = _T(r[k + 1], r[k])
    out[14] = r[7]
    out[15] = r[7]
    return out.reshape(-1)


def up_hv_scalar(x: np.ndarray) -> np.ndarray:
    """upsampler/scalar.rs:148-166: H(V(x)), V seeing the strip as 8 double-rows (Q4d)"""
    return up_h_scalar(up_v(x))


def up_hv_avx2(x: np.ndarray) -> np.ndarray:
    """upsampler/avx2.rs:29-342 closed form (SURVEY A.4 'HV, X86/AVX2', Q4e/f/g); n >= 500, n % 128 == 0."""
    n = x.size
    if n < 500:
        return up_hv_scalar(x)
    assert n % 128 == 0
    S = n // 8
    L = 2 * S
    xi = x.astype(np.int32)

    def get(idx):  # .get(i).unwrap_or(&0)
        return int(xi[idx]) if 0 <= idx < n else 0

    def w16(v):
        return ((int(v) + 32768) & 0xFFFF) - 32768

    out = np.zeros(4 * n, np.int16)
    strides = [0, S, S, S, S, S, S, 0]
    nvec = S // 16 - 1
    for j in range(8):
        sj = strides[j]
        near_in = x[j * S : (j + 1) * S]
        far_in = x[j * S + sj : (j + 1) * S + sj]
        rn = _T(near_in, far_in)
        rf = _T(far_in, near_in)
        for R, base in ((rn, 2 * j * L), (rf, (2 * j + 1) * L)):
            o = np.zeros(L, np.int16)
            i = np.arange(0, S - 16)
            m = np.zeros(S - 16, np.int16)
            q = np.zeros(S - 16, np.int16)
            m[1:] = R[0 : S - 17]
            q[:] = R[1 : S - 15]
            for t in range(nvec):
                if t >= 1:
                    p, s = j * S + 16 * t, sj
                    prev = w16(3 * w16(w16(get(p) + get(p + s)) + 2)) >> 2
                    pfar = w16(3 * w16(w16(get(p + 16) + get(p + 16 + s)) + 2)) >> 2
                elif j == 0:
                    prev, pfar = int(x[0]), int(x[16])
                else:
                    p, s = j * S - 16, strides[j - 1]
                    prev = w16(3 * w16(w16(get(p) + get(p + s)) + 2)) >> 2
                    pfar = w16(3 * w16(w16(get(p + 16) + get(p + 16 + s)) + 2)) >> 2
                m[16 * t] = prev
                q[16 * t + 15] = pfar
            o[2 * i] = _T(R[: S - 16], m)
            o[2 * i + 1] = _T(R[: S - 16], q)
            out[base : base + L] = o
        # tail: last 32 outputs of each out double-row from raw shifted input
        P = (j + 1) * S - 16
        for sh, base in ((0, 2 * j * L), (sj, (2 * j + 1) * L)):
            O = base + L - 32
            for k in range(15):
                c = P - 17 + k + sh
                out[O + 2 * k] = _T(x[c : c + 1], x[c - 1 : c])[0]
                out[O + 2 * k + 1] = _T(x[c : c + 1], x[c + 1 : c + 2])[0]
            out[O + 30] = out[O + 28]
            out[O + 31] = out[O + 29]
        fb = (2 * j + 1) * L
        out[fb] = out[fb + 1]
    return out


# --------------------------------------------------------------------------- colour writer
def conv_rgb(y, cb, cr):
    """color_convert/scalar.rs:66-85 == avx.rs:123-192, wrapping i16 -> three uint8 arrays"""
    with np.errstate(over="ignore"):
        i16 = np.int16
        y = y.astype(i16)
        cb = (cb.astype(i16) - i16(128)).astype(i16)
        cr = (cr.astype(i16) - i16(128)).astype(i16)
        r = (y + ((i16(45) * cr).astype(i16) >> i16(5))).astype(i16)
        g = (y - (((i16(11) * cb).astype(i16) + (i16(23) * cr).astype(i16)).astype(i16) >> i16(5))).astype(i16)
        b = (y + ((i16(113) * cb).astype(i16) >> i16(6))).astype(i16)
        return (np.clip(r, 0, 255).astype(np.uint8), np.clip(g, 0, 255).astype(np.uint8), np.clip(b, 0, 255).astype(np.uint8))


def write_rgb_rows(Y, Cb, Cr, width, nc, out_rows):
    """color_convert_ycbcr, worker.rs:143-251 closed form.
    Y/Cb/Cr: int16 [rows, Wp]; out_rows: uint8 [rows, width*nc] (zero-initialised)."""
    rows, Wp = Y.shape
    stride = width * nc
    r, g, b = conv_rgb(Y, Cb, Cr)
    rgb = np.stack([r, g, b], axis=-1).reshape(rows, Wp * 3)
    if width < 16:
        if Wp > 16:
            raise RefPanic
        pad = np.zeros((rows, 48), np.uint8)
        # samples beyond Wp are 0 -> conv of (0,0,0)
        z = conv_rgb(np.zeros(1, np.int16), np.zeros(1, np.int16), np.zeros(1, np.int16))
        pad[:, :] = np.tile(np.array([z[0][0], z[1][0], z[2][0]], np.uint8), 16)
        pad[:, : Wp * 3] = rgb
        temp = np.zeros((rows, 16 * nc), np.uint8)
        temp[:, :48] = pad
        out_rows[:, : width * nc] = temp[:, : width * nc]
        return
    E = max(Wp // 16 - 1, 0)
    P = 48 * E
    if P > stride:
        raise RefPanic
    out_rows[:, :P] = rgb[:, :P]
    d = max(0, 64 - (stride - P))
    T = max(0, P - d)
    if T + 48 > stride:
        raise RefPanic
    out_rows[:, T : T + 48] = rgb[:, (Wp - 16) * 3 : Wp * 3]


def write_gray_rows(Yflat, width, out_flat):
    """ycbcr_to_grayscale, color_convert/scalar.rs:91-114 (Q7)"""
    n = Yflat.size
    width_mcu = n // width
    if width_mcu == 0:
        raise RefPanic
    width_chunk = n // width_mcu
    nchunks = n // width_chunk
    if width > width_chunk or nchunks * width > out_flat.size:
        raise RefPanic
    v = Yflat[: nchunks * width_chunk].reshape(nchunks, width_chunk)[:, :width].astype(np.uint16).astype(np.uint8)
    out_flat[: nchunks * width] = v.reshape(-1)


# --------------------------------------------------------------------------- whole image
def geometry(width, height, h_max, v_max, nc, interleaved_flag_progressive=False):
    mcu_x = -(-width // (8 * h_max))
    mcu_y = -(-height // (8 * v_max))
    if (h_max, v_max) == (1, 1):
        bw, strips, ybr, cbr = -(-width // 8), -(-height // 8), 1, 1
    elif (h_max, v_max) == (2, 1):
        bw, strips, ybr, cbr = mcu_x, mcu_y // 2, 2, 2
    elif (h_max, v_max) == (1, 2):
        bw, strips, ybr, cbr = mcu_x, mcu_y, 2, 1
    elif (h_max, v_max) == (2, 2):
        bw, strips, ybr, cbr = mcu_x, mcu_y // 2, 4, 2
    else:
        raise ValueError
    out_chunk = width * nc * 8 * h_max * v_max
    interleaved = (h_max, v_max) != (1, 1)
    alloc = ((width + 8) & 0xFFFF) * ((height + 8) & 0xFFFF) * nc + (128 * height * nc if interleaved else 0)
    avail = alloc // out_chunk
    if strips > avail:
        if interleaved_flag_progressive:
            strips = avail
        else:
            raise RefPanic
    return dict(mcu_x=mcu_x, mcu_y=mcu_y, strips=strips, y_block_rows=ybr, c_block_rows=cbr, out_chunk=out_chunk)


def reconstruct(width, height, planes, qts, h_max, v_max, out_cs, variant, progressive=False):
    """planes: list of int16 arrays (flat coefficient planes, 1 or 3); qts: list of int32[64].
    Returns uint8 [width*height*nc]."""
    n_comp = len(planes)
    nc = NCOMP[out_cs]
    g = geometry(width, height, h_max, v_max, nc, progressive)
    x = min(n_comp, nc)
    out = np.zeros(max(g["strips"] * g["out_chunk"], width * height * nc), np.uint8)
    rows_out = 8 * h_max * v_max
    ybpr = h_max * g["mcu_x"]          # Y blocks per block-row
    cbpr = g["mcu_x"]
    for s in range(g["strips"]):
        P = []
        for z in range(x):
            bpr = ybpr if z == 0 else cbpr
            brs = g["y_block_rows"] if z == 0 else g["c_block_rows"]
            nblk = bpr * brs
            blk = planes[z][s * nblk * 64 : (s + 1) * nblk * 64].reshape(nblk, 64)
            P.append(plane_from_blocks(idct_blocks(blk, qts[z], variant), bpr))
        Y = P[0]
        Wp = Y.shape[1]
        if (h_max, v_max) != (1, 1):
            up = {(2, 1): up_h_sse if variant == X86 else up_h_scalar,
                  (1, 2): up_v,
                  (2, 2): up_hv_avx2 if variant == X86 else up_hv_scalar}[(h_max, v_max)]
            for z in range(1, x):
                P[z] = up(P[z].reshape(-1)).reshape(rows_out, Wp)
        chunk = out[s * g["out_chunk"] : (s + 1) * g["out_chunk"]]
        if out_cs == CS_GRAY:
            write_gray_rows(Y.reshape(-1), width, chunk)
        elif n_comp == 3 and out_cs == CS_YCBCR:
            v = np.stack([P[0][:, :width], P[1][:, :width], P[2][:, :width]], axis=-1).astype(np.uint16).astype(np.uint8)
            chunk[:] = v.reshape(-1)
        elif n_comp == 3 and out_cs in (CS_RGB, CS_RGBA, CS_RGBX):
            write_rgb_rows(P[0], P[1], P[2], width, nc, chunk.reshape(rows_out, width * nc))
        # anything else: zeros (worker.rs:131-132)
    return out[: width * height * nc].copy()
