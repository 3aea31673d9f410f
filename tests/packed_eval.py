"""Numpy evaluation of the PACKED weight form (what the CUDA kernels consume), used to validate the
packer's layouts and algebraic folds against the oracle without a GPU.  Test-only."""
import struct

import numpy as np

import rtdm_b200.pack as P


def parse_blob(blob):
    magic, version, arch, prec, n, *_ = struct.unpack_from("<8I", blob, 0)
    assert magic == P.MAGIC and version == P.VERSION
    out = {}
    for i in range(n):
        tid, dt, off, nb = struct.unpack_from("<IIQQ", blob, 32 + 24 * i)
        assert off % 256 == 0 and off + nb <= len(blob)
        out[tid] = (dt, bytes(blob[off:off + nb]))
    return {"arch": arch, "precision": prec, "tensors": out}


def f32(parsed, tid, shape):
    dt, raw = parsed["tensors"][tid]
    assert dt == P.DT_F32
    return np.frombuffer(raw, dtype=np.float32).reshape(shape).astype(np.float64)


def forward_packed(blob, x, arch):
    """x: (B,3,140,140) -> logits (B,5), computed the way the simt kernels do (NHWC, packed weights)."""
    pz = parse_blob(blob)
    red = arch == "squeeze-redconv"
    w = P.widths(arch)
    cs = w[0][0]
    xh = np.transpose(x.astype(np.float64), (0, 2, 3, 1))                      # NHWC
    sw = f32(pz, P.T_STEM_W, (3, 3, 3, cs))
    sb = f32(pz, P.T_STEM_B, (cs,))
    B = x.shape[0]
    a = np.zeros((B, 69, 69, cs)) + sb
    for ky in range(3):
        for kx in range(3):
            a += np.einsum("bhwc,co->bhwo", xh[:, ky:ky + 137:2, kx:kx + 137:2, :], sw[ky, kx])
    taps = {"stem": a}

    def block(k, a, out_hw, pool, c, co):
        base = P.T_BLOCK_BASE + 8 * k
        dw = f32(pz, base + P.T_DW_W, (3, 9, c))
        db = f32(pz, base + P.T_DW_B, (3, c))
        H = a.shape[1]
        ap = np.zeros((B, H + 6, H + 6, c))
        ap[:, 2:2 + H, 2:2 + H] = a
        cat = np.zeros((B, out_hw, out_hw, 3 * c))
        for d in range(3):
            dil = d + 1
            acc = np.zeros((B, out_hw, out_hw, c)) + db[d]
            for ky in range(3):
                for kx in range(3):
                    oy = 2 + ky * dil - (dil - 1)
                    ox = 2 + kx * dil - (dil - 1)
                    acc += ap[:, oy:oy + out_hw, ox:ox + out_hw] * dw[d, ky * 3 + kx]
            cat[..., d * c:(d + 1) * c] = acc
        pw = f32(pz, base + P.T_PW_W, (3 * c, co))
        z = cat @ pw + f32(pz, base + P.T_PW_B, (co,))
        z = np.where(z > 0, z, 0.01 * z)
        z = z * f32(pz, base + P.T_BN_S, (co,)) + f32(pz, base + P.T_BN_T, (co,))
        return z

    def pool(z):
        return np.maximum(np.maximum(z[:, 0::2, 0::2], z[:, 0::2, 1::2]), np.maximum(z[:, 1::2, 0::2], z[:, 1::2, 1::2]))

    a = pool(block(0, a, 66, True, *w[0]))
    taps["pool1"] = a
    z = block(1, a, 30, True, *w[1])
    if red:
        z = z @ f32(pz, P.T_RED2_W, (96, 48)) + f32(pz, P.T_RED2_B, (48,))
    a = pool(z)
    taps["pool2"] = a
    a = pool(block(2, a, 12, True, *w[2]))
    if red:
        a = a @ f32(pz, P.T_RED3_W, (128, 64)) + f32(pz, P.T_RED3_B, (64,))
    taps["pool3"] = a
    a = block(3, a, 4, False, *w[3])
    taps["acff4"] = a
    s = a.sum(axis=(1, 2))                                                     # (B,256)
    logits = s @ f32(pz, P.T_HEAD_W, (5, 256)).T + f32(pz, P.T_HEAD_B, (5,))
    return logits, taps
