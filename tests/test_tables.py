"""Closed-form tables (tools/gen_tables.py -> csrc/sr_tables.h) against the reference's own numbers, the
log threshold table against the host libm expression of MFCC.C:168, and the kernel's filter partition
against a brute-force restatement of MFCC.C:136-162. CPU only."""
import ctypes as C
import math
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_tables  # noqa: E402
import oracle_bind as ob  # noqa: E402

REF = "/root/reference"
need_ref = pytest.mark.skipif(not os.path.isdir(REF + "/Src/Speech_Recog"), reason="reference tree not mounted")


def _ref_arrays():
    src = open(REF + "/Src/Speech_Recog/MFCC_Arg.h", "rb").read().decode("gb18030").replace("\r", "")
    def arr(name):
        m = re.search(name + r"\[\]=\s*\{([^}]*)\}", src)
        return [int(x) for x in re.findall(r"-?\d+", m.group(1))]
    return {k: arr(k) for k in ("hamm", "tri_cen", "tri_odd", "tri_even", "dct_arg")}


@need_ref
def test_mfcc_tables_equal_reference_header():
    r = _ref_arrays()
    cen, odd, even = gen_tables.tri_tables()
    assert gen_tables.hamm_table() == r["hamm"]           # MFCC_Arg.h:6-9
    assert cen == r["tri_cen"]                            # MFCC_Arg.h:12-15
    assert odd == r["tri_odd"]                            # MFCC_Arg.h:18-21
    assert even == r["tri_even"]                          # MFCC_Arg.h:24-27
    assert gen_tables.dct_table() == r["dct_arg"]         # MFCC_Arg.h:30-44


@need_ref
def test_twiddles_equal_reference_asm_table():
    text = open(REF + "/Src/BSP/cr4_fft_1024_stm32.s", "rb").read().decode("gb18030").replace("\r", "")
    body = text[text.index("TableFFT_V7\n"):]
    vals = []
    for line in body.split("\n"):
        if "DCW" in line:
            vals += [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{4})", line)]
    vals = [v - 65536 if v >= 32768 else v for v in vals]
    assert len(vals) == 2040
    assert gen_tables.twiddle_table() == vals             # .s:285-629


def test_committed_header_is_current():
    """sr_tables.h in the tree is what the generator produces now"""
    path = os.path.join(ROOT, "stm32-speech-recognition_b200", "csrc", "sr_tables.h")
    text = open(path).read()
    hamm = [int(x) for x in re.search(r"sr_tab_hamm\[160\] = \{([^}]*)\}", text).group(1).replace("\n", "").split(",") if x.strip()]
    assert hamm == gen_tables.hamm_table()
    tw = [int(x) for x in re.search(r"sr_tab_twiddle\[2040\] = \{([^}]*)\}", text).group(1).replace("\n", "").split(",") if x.strip()]
    assert tw == gen_tables.twiddle_table()


def test_log_threshold_table_matches_libm_expression():
    """thr[L] is the first v with (u32)(log((double)v)*100) >= L for the host's libm (MFCC.C:168)"""
    thr, lmax = gen_tables.log_thresholds()
    assert lmax == 2218 and thr[0] == 1
    o = ob.port()
    for L in range(1, lmax + 1):
        v = thr[L]
        assert o.lib.sro_log100(v) >= L, (L, v)          # plateaus: thr[1..69] = 2 because log100(2) = 69
        assert o.lib.sro_log100(v - 1) < L, (L, v)
    assert o.lib.sro_log100(0) == 0 and o.lib.sro_log100(1) == 0 and o.lib.sro_log100(0xFFFFFFFF) == 2218
    # python's own double log agrees as well (same expression)
    rng = np.random.default_rng(3)
    for v in rng.integers(1, 2 ** 32, 2000, dtype=np.uint64).tolist():
        assert o.lib.sro_log100(v) == int(math.log(float(v)) * 100)


def test_filter_partition_reproduces_reference_ranges():
    """the lane-chunk / partial-sum partition used by mfcc_kernel covers exactly the bins of MFCC.C:136-162"""
    import sr_b200
    L = sr_b200.lib()
    se, so = np.zeros(32, np.uint8), np.zeros(32, np.uint8)
    lo, hi = np.zeros(24, np.uint8), np.zeros(24, np.uint8)
    L.sr_debug_filter_partition(se.ctypes.data_as(C.c_void_p), so.ctypes.data_as(C.c_void_p),
                                lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p))
    cen, _, _ = gen_tables.tri_tables()
    # reference ranges
    rng = {0: (0, cen[1]), 23: (cen[22], 512)}
    for h in range(2, 24, 2):
        rng[h] = (cen[h - 1], cen[h + 1])
    for h in range(1, 22, 2):
        rng[h] = (cen[h - 1], cen[h + 1])
    for h in range(24):
        split = so if h & 1 else se
        bins = []
        for e in range(int(lo[h]), int(hi[h]) + 1):
            lane, part = e >> 1, e & 1
            a, b = (0, int(split[lane])) if part == 0 else (int(split[lane]), 16)
            bins += [16 * lane + i for i in range(a, b)]
        assert bins == list(range(*rng[h])), h
