#!/usr/bin/env python3
"""Numpy model of the kernel's FFT blocking (development aid, CPU only).

Checks, against the oracle's cr4 restatement, the index algebra the CUDA kernel uses:
  * stage 0 collapses to y0[4*idx+m] = (x[bitrev8(idx)]>>2, 0) for a frame of <=256 real samples
  * block A (G,q1): 4 stage-1 butterflies (groups 4G+m2) then 4 stage-2 butterflies (q2=q1+4*m1)
  * block B (q3): 4 stage-3 butterflies (groups m4) then 4 stage-4 butterflies (q4=q3+64*m3)
Run: python tools/fft_block_model.py
"""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_tables  # noqa: E402

M32 = np.uint32(0xFFFFFFFF)


def asr(x, n):
    return (x.astype(np.int32) >> n).astype(np.uint32)


def sx16(x):
    return (x & np.uint32(0xFFFF)).astype(np.uint16).astype(np.int16).astype(np.int32).astype(np.uint32)


def cxmul(yr, yi, P, S):
    P = np.uint32(P & 0xFFFFFFFF) if np.isscalar(P) else P
    zr = yr * P + yi * S
    zi = yi * P - yr * S
    return zr, zi


def tree(A, B, Cc, D, s):
    Ar, Ai = A; Br, Bi = B; Cr, Ci = Cc; Dr, Di = D
    Cr, Ci, Dr, Di = Cr + Dr, Ci + Di, Cr - Dr, Ci - Di
    Ar, Ai = asr(Ar, 2), asr(Ai, 2)
    Ar, Ai = Ar + asr(Br, 2 + s), Ai + asr(Bi, 2 + s)
    Br, Bi = Ar - asr(Br, 1 + s), Ai - asr(Bi, 1 + s)
    Ar, Ai = Ar + asr(Cr, 2 + s), Ai + asr(Ci, 2 + s)
    Cr, Ci = Ar - asr(Cr, 1 + s), Ai - asr(Ci, 1 + s)
    Br = Br + asr(Di, 2 + s)
    Bi = Bi - asr(Dr, 2 + s)
    Yr = Br - asr(Di, 1 + s)      # leg-3 real  (asm keeps it in the register named Di)
    Yi = Bi + asr(Dr, 1 + s)      # leg-3 imag
    return (Ar, Ai), (Br, Bi), (Cr, Ci), (Yr, Yi)


def twiddles():
    t = np.array(gen_tables.twiddle_table(), dtype=np.int64).reshape(-1, 3, 2)   # [triple][leg3,leg2,leg1][Ka,Kb]
    P = (t[:, :, 0] + t[:, :, 1]).astype(np.int32).astype(np.uint32)
    S = t[:, :, 1].astype(np.int32).astype(np.uint32)
    return P, S   # index [triple][leg], leg order (K3, K2, K1)


def bitrev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def model_fft(w):
    """w: int16[160] windowed frame -> packed u32[1024] like the asm output (all bins)."""
    P, S = twiddles()
    OFF = {4: 0, 16: 4, 64: 20, 256: 84}
    wz = np.zeros(256, np.int32); wz[:len(w)] = w
    y2 = np.zeros((1024, 2), np.uint32)
    z = np.uint32(0)
    for G in range(16):
        r0 = bitrev(G, 4)
        for q1 in range(4):
            v = {}
            for m2 in range(4):
                r = r0 + 16 * bitrev(m2, 2)
                a = np.uint32(np.int32(wz[r]) >> 2)
                b = np.uint32(np.int32(wz[r + 128]) >> 2) if r + 128 < 256 else z
                c = np.uint32(np.int32(wz[r + 64]) >> 2)
                tr = OFF[4] + q1
                Cz = cxmul(c, z, P[tr, 1], S[tr, 1])
                Bz = cxmul(b, z, P[tr, 2], S[tr, 2])
                outs = tree((a, z), Bz, Cz, (z, z), 14)
                for m1 in range(4):
                    v[(m2, m1)] = (sx16(outs[m1][0]), sx16(outs[m1][1]))
            for m1 in range(4):
                q2 = q1 + 4 * m1
                tr = OFF[16] + q2
                D = cxmul(*v[(3, m1)], P[tr, 0], S[tr, 0])
                Cc = cxmul(*v[(2, m1)], P[tr, 1], S[tr, 1])
                B = cxmul(*v[(1, m1)], P[tr, 2], S[tr, 2])
                outs = tree(v[(0, m1)], B, Cc, D, 14)
                for m2o in range(4):
                    e = 64 * G + q2 + 16 * m2o
                    y2[e, 0] = sx16(outs[m2o][0]); y2[e, 1] = sx16(outs[m2o][1])
    out = np.zeros((1024, 2), np.uint32)
    for q3 in range(64):
        v = {}
        for m4 in range(4):
            tr = OFF[64] + q3
            legs = [tuple(y2[256 * m4 + q3 + 64 * m3]) for m3 in range(4)]
            D = cxmul(*legs[3], P[tr, 0], S[tr, 0])
            Cc = cxmul(*legs[2], P[tr, 1], S[tr, 1])
            B = cxmul(*legs[1], P[tr, 2], S[tr, 2])
            outs = tree(legs[0], B, Cc, D, 14)
            for m3 in range(4):
                v[(m4, m3)] = (sx16(outs[m3][0]), sx16(outs[m3][1]))
        for m3 in range(4):
            q4 = q3 + 64 * m3
            tr = OFF[256] + q4
            D = cxmul(*v[(3, m3)], P[tr, 0], S[tr, 0])
            Cc = cxmul(*v[(2, m3)], P[tr, 1], S[tr, 1])
            B = cxmul(*v[(1, m3)], P[tr, 2], S[tr, 2])
            outs = tree(v[(0, m3)], B, Cc, D, 14)
            for m4o in range(4):
                k = q4 + 256 * m4o
                out[k, 0] = sx16(outs[m4o][0]); out[k, 1] = sx16(outs[m4o][1])
    return (out[:, 0] & np.uint32(0xFFFF)) | (out[:, 1] << np.uint32(16))


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
    rng = np.random.default_rng(7)
    np.seterr(over="ignore")
    for it in range(6):
        amp = [50, 500, 3000, 20000, 32767, 32767][it]
        w = rng.integers(-amp, amp + 1, 160).astype(np.int16)
        inp = np.zeros(1024, np.uint32); inp[:160] = w.astype(np.uint16)
        ref = np.zeros(1024, np.uint32)
        lib.sro_fft_raw(inp.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p))
        got = model_fft(w)
        ok = (got == ref).all()
        print("amp", amp, "match" if ok else "MISMATCH %d" % int((got != ref).sum()))
        assert ok


if __name__ == "__main__":
    main()
