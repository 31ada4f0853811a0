"""Host side of the packed PCM transport (csrc/sr_pack_host.cpp): every SIMD variant and the worker pool against a
numpy restatement of the format (3 bytes per sample pair: a | b << 12, little endian). No GPU needed."""
import numpy as np
import pytest

import sr_b200


def ref_pack(x):
    a, b = x[0::2].astype(np.uint32), x[1::2].astype(np.uint32)
    t = (a & 0xFFF) | ((b & 0xFFF) << 12)
    out = np.empty((a.size, 3), np.uint8)
    out[:, 0], out[:, 1], out[:, 2] = t & 0xFF, (t >> 8) & 0xFF, (t >> 16) & 0xFF
    return out.reshape(-1)


@pytest.mark.parametrize("variant", [-1, 0, 1, 2, 3, 101, 103, 107, 114])
def test_pack12_matches_format(variant):
    rng = np.random.default_rng(variant + 200)
    for n in [2, 14, 16, 30, 32, 34, 62, 64, 66, 126, 128, 130, 254, 256, 258, 1000, 4098, 100002, 2096 * 8000]:
        x = rng.integers(0, 4096, n).astype(np.uint16)
        r = sr_b200.pack12_host(x, variant)
        if r is None:
            pytest.skip("SIMD variant not available on this CPU")
        packed, orbits = r
        assert (orbits & 0xF000) == 0
        assert np.array_equal(packed, ref_pack(x)), n
        x[int(rng.integers(0, n))] = 0x1000 + int(rng.integers(0, 0xF000))   # one sample outside the 12-bit range
        assert sr_b200.pack12_host(x, variant)[1] & 0xF000, n              # -> the caller must send this chunk plain


def test_pack12_nontemporal_variant_alignment():
    """variant 3 streams whole cache lines when the destination is 64-byte aligned and falls back otherwise; neither may
    touch a byte outside the packed range"""
    import ctypes as C
    rng = np.random.default_rng(5)
    for n in [128, 130, 256, 1000, 128 * 77, 100002]:
        x = rng.integers(0, 4096, n).astype(np.uint16)
        raw = np.full(n // 2 * 3 + 256, 0xCD, np.uint8)
        off = (-raw.ctypes.data) % 64
        for extra in (0, 16):
            raw[:] = 0xCD
            dst = raw[off + extra: off + extra + n // 2 * 3]
            o = sr_b200.lib().sr_debug_pack12_host(3, C.c_void_p(x.ctypes.data), n, C.c_void_p(dst.ctypes.data))
            if o == 0xFFFFFFFF:
                pytest.skip("AVX-512 VBMI not available on this CPU")
            assert (o & 0xF000) == 0 and np.array_equal(dst, ref_pack(x)), (n, extra)
            assert (raw[:off + extra] == 0xCD).all() and (raw[off + extra + n // 2 * 3:] == 0xCD).all(), (n, extra)
