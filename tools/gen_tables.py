#!/usr/bin/env python3
"""Generate every constant table of the VAD->MFCC->DTW path from its closed form.

Nothing here is copied from the reference: each table is recomputed from the formula the
reference's Matlab tooling used, and tests/test_tables.py checks (when /root/reference is
mounted) that the result is identical, element for element, to what the reference ships:

  hamm[160]      Matlab/matlab仿真/speech_recog.m:217-225  -> Src/Speech_Recog/MFCC_Arg.h:6-9
  tri_cen[24]    speech_recog.m:240-271                     -> MFCC_Arg.h:12-15
  tri_odd/even   speech_recog.m:274-313                     -> MFCC_Arg.h:18-27
                 (Matlab is 1-based: Matlab "odd" is the C array tri_even and vice versa)
  dct_arg[12*24] Matlab/matlab仿真/teat.m:19-27             -> MFCC_Arg.h:30-44
  twiddles       Src/BSP/cr4_fft_1024_stm32.s:285-629 (TableFFT_V7): 340 triples of
                 (Ka,Kb) = (round(2^14(cos t - sin t)), round(2^14 sin t)), t = 3p, p, 2p
  log table      thr[L] = min{ v : floor(100*ln v) >= L } -- replaces the per-filter
                 (u32)(log((double)v)*100) of MFCC.C:168 by an exact integer lookup.

Usage: python tools/gen_tables.py   (rewrites stm32-speech-recognition_b200/csrc/sr_tables.h)
"""
import math
import os
from decimal import Decimal, getcontext

FS = 8000
FRAME_LEN = 160
FFT_POINT = 1024
FRQ_MAX = FFT_POINT // 2
TRI_NUM = 24
MFCC_NUM = 12


def mround(x):
    """Matlab int32(): round half away from zero."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


# GEOM_B (BASELINE configs[0]: 25 ms frames, 10 ms hop, 256-point FFT): an EXTENSION -- the reference only ships the
# 160/80/1024 geometry; the same Matlab formulas are evaluated with frame_len = 200 and fft_point = 256
FRAME_LEN_B = 200
FFT_POINT_B = 256
FRQ_MAX_B = FFT_POINT_B // 2


def hamm_table(frame_len=FRAME_LEN):
    return [mround(10000 * (0.54 - 0.46 * math.cos(2 * math.pi * i / (frame_len - 1))))
            for i in range(frame_len)]


def dct_table():
    return [mround(100 * math.cos((c + 1) * (2 * h + 1) * math.pi / (2 * TRI_NUM)))
            for c in range(MFCC_NUM) for h in range(TRI_NUM)]


def tri_tables(frq_max=FRQ_MAX):
    FRQ_MAX = frq_max
    f_max = FS / 2
    mel_max = 2595 * math.log10(1 + f_max / 700)
    mel_step = mel_max / (TRI_NUM + 1)
    mel_thl = 1000
    cen = []
    for i in range(1, TRI_NUM + 1):
        if i < mel_thl / mel_step:
            v = mel_step * i
        else:
            v = (math.exp(math.log(10) * (mel_step * i) / 2595) - 1) * 700
        cen.append(mround(v / (f_max / FRQ_MAX)))
    n, top = FRQ_MAX, 1000
    tc = [0] + cen                      # 1-based like the Matlab source
    m_odd = [0.0] * (n + 1)
    m_even = [0.0] * (n + 1)
    for j in range(1, tc[1] + 1):
        m_odd[j] = top * j / tc[1]
    for j in range(tc[1] + 1, tc[2] + 1):
        m_odd[j] = top * (tc[2] - j) / (tc[2] - tc[1])
    for h in range(3, TRI_NUM + 1, 2):
        for j in range(tc[h - 1], tc[h] + 1):
            m_odd[j] = top * (j - tc[h - 1]) / (tc[h] - tc[h - 1])
        for j in range(tc[h] + 1, tc[h + 1] + 1):
            m_odd[j] = top * (tc[h + 1] - j) / (tc[h + 1] - tc[h])
    for h in range(2, TRI_NUM - 1, 2):
        for j in range(tc[h - 1], tc[h] + 1):
            m_even[j] = top * (j - tc[h - 1]) / (tc[h] - tc[h - 1])
        for j in range(tc[h] + 1, tc[h + 1] + 1):
            m_even[j] = top * (tc[h + 1] - j) / (tc[h + 1] - tc[h])
    for j in range(tc[TRI_NUM - 1], tc[TRI_NUM] + 1):
        m_even[j] = top * (j - tc[TRI_NUM - 1]) / (tc[TRI_NUM] - tc[TRI_NUM - 1])
    for j in range(tc[TRI_NUM] + 1, n + 1):
        m_even[j] = top * (n - j) / (n - tc[TRI_NUM])
    # C arrays are 0-based: C tri_even[k] = Matlab odd(k+1), C tri_odd[k] = Matlab even(k+1)
    c_even = [mround(v) for v in m_odd[1:]]
    c_odd = [mround(v) for v in m_even[1:]]
    return cen, c_odd, c_even


def twiddle_table():
    """340 triples x (Ka,Kb) in TableFFT_V7 order: blocks for stride s=4,16,64,256; inside a
    block butterfly q=0..s-1; inside a triple the legs p3,p2,p1 with angles 3p,p,2p."""
    out = []
    for s in (4, 16, 64, 256):
        for q in range(s):
            phi = 2 * math.pi * q / (4 * s)
            for mult in (3, 1, 2):
                th = mult * phi
                out.append(mround(16384 * (math.cos(th) - math.sin(th))))
                out.append(mround(16384 * math.sin(th)))
    return out


def log_thresholds():
    """thr[L], L=0..2218: smallest u32 v>=1 with floor(100 ln v) >= L; thr[2219] = 2^32-1 pad."""
    getcontext().prec = 60
    lmax = int((Decimal(2 ** 32 - 1).ln() * 100).to_integral_value(rounding="ROUND_FLOOR"))
    thr = [1]
    for L in range(1, lmax + 1):
        e = (Decimal(L) / 100).exp()
        v = int(e.to_integral_value(rounding="ROUND_CEILING"))
        # guard: floor(100 ln v) >= L and floor(100 ln (v-1)) < L
        assert (Decimal(v).ln() * 100) >= L and (Decimal(v - 1).ln() * 100) < L
        thr.append(v)
    return thr, lmax


def c_array(ctype, name, vals, per_line=16):
    lines = ["static const %s %s[%d] = {" % (ctype, name, len(vals))]
    for i in range(0, len(vals), per_line):
        lines.append("  " + ",".join(str(v) for v in vals[i:i + per_line]) + ",")
    lines.append("};")
    return "\n".join(lines)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "stm32-speech-recognition_b200", "csrc", "sr_tables.h")
    cen, odd, even = tri_tables()
    cen_b, odd_b, even_b = tri_tables(FRQ_MAX_B)
    thr, lmax = log_thresholds()
    parts = [
        "// GENERATED by tools/gen_tables.py -- do not edit. Closed-form tables of the MFCC path;",
        "// tests/test_tables.py proves them identical to Src/Speech_Recog/MFCC_Arg.h:6-44 and",
        "// Src/BSP/cr4_fft_1024_stm32.s:285-629 of the reference.",
        "#ifndef SR_TABLES_H_",
        "#define SR_TABLES_H_",
        "#include <stdint.h>",
        "#define SR_LOG_LMAX %d" % lmax,
        c_array("uint16_t", "sr_tab_hamm", hamm_table()),
        c_array("uint16_t", "sr_tab_tri_cen", cen),
        c_array("uint16_t", "sr_tab_tri_odd", odd),
        c_array("uint16_t", "sr_tab_tri_even", even),
        c_array("int8_t", "sr_tab_dct", dct_table(), 24),
        "/* GEOM_B (200/80/256): extension, no counterpart in the reference's headers */",
        c_array("uint16_t", "sr_tab_b_hamm", hamm_table(FRAME_LEN_B)),
        c_array("uint16_t", "sr_tab_b_tri_cen", cen_b),
        c_array("uint16_t", "sr_tab_b_tri_odd", odd_b),
        c_array("uint16_t", "sr_tab_b_tri_even", even_b),
        c_array("int16_t", "sr_tab_twiddle", twiddle_table(), 12),
        c_array("uint32_t", "sr_tab_log_thr", thr + [2 ** 32 - 1], 8),
        "#endif",
        "",
    ]
    with open(dst, "w") as f:
        f.write("\n".join(parts))
    print("wrote", os.path.normpath(dst))
    # Q15 sine table of the synthetic-PCM workload generator (sr_synth.cu); not part of the reference path
    dst2 = os.path.join(here, "..", "stm32-speech-recognition_b200", "csrc", "sr_synth_tables.h")
    sine = [mround(32767 * math.sin(2 * math.pi * k / 1024)) for k in range(1024)]
    with open(dst2, "w") as f:
        f.write("// GENERATED by tools/gen_tables.py -- Q15 sine, round(32767*sin(2*pi*k/1024))\n")
        f.write("#ifndef SR_SYNTH_TABLES_H_\n#define SR_SYNTH_TABLES_H_\n#include <stdint.h>\n")
        f.write(c_array("int16_t", "sr_synth_sine", sine) + "\n#endif\n")
    print("wrote", os.path.normpath(dst2))


if __name__ == "__main__":
    main()
