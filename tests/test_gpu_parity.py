"""GPU parity tests proper: every call goes through the C-ABI of libspeech_b200.so and is compared
BIT-EXACTLY (integer path: tolerance 0) with the oracle -- the reference's own C when oracle/_ref/libref.so
travelled with the repo, else the restatement -- and with the committed golden vectors."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_bind as ob
import sr_b200

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CAPS = np.load(os.path.join(HERE, "golden", "captures.npz"))
GOLD = np.load(os.path.join(HERE, "golden", "golden.npz"))


@pytest.fixture(scope="module")
def ora():
    return ob.best_oracle()


def _cmp_recog(out, ref, keys=("seg_off", "score", "best_idx", "best_dis", "cmd", "status")):
    for k in keys:
        assert np.array_equal(out[k].reshape(-1), ref[k].reshape(-1)), k
    assert ob.ftr_equal(out["ftr"], ref["ftr"])


def test_fast_sqrt_exhaustive(handle):
    """the branch-free correctly-rounded sqrt used by mag (MFCC.C:58) and get_dis (DTW.C:59) equals the IEEE
    intrinsic for EVERY float in [1, 2^33) -- a superset of what (float)(s32 pw) and (float)(u32 d) can be"""
    import struct
    lo = struct.unpack("<I", struct.pack("<f", 1.0))[0]
    hi = struct.unpack("<I", struct.pack("<f", 2.0 ** 33))[0]
    bad = C.c_uint64(123)
    rc = sr_b200.lib().sr_debug_sqrt_mismatches(handle._h, lo, hi, C.byref(bad))
    assert rc == 0 and bad.value == 0


# ---- FFT: the asm restatement on device, arbitrary complex inputs -----------------------------------
def test_fft_raw_bit_exact(handle, ora):
    rng = np.random.default_rng(21)
    x = rng.integers(0, 2 ** 32, (96, 1024), dtype=np.uint32)
    x[:4] = 0
    x[4:8] = 0x7FFF7FFF
    x[8:12] = 0x80008000                               # -32768 everywhere: pw = 2^31 corner of MFCC.C:56
    x[12:44, :] = rng.integers(-3000, 3000, (32, 1024)).astype(np.int16).astype(np.uint16)
    assert np.array_equal(handle.fft_raw(x), ora.fft_raw(x))


@pytest.mark.parametrize("length", [0, 1, 160, 161, 256, 1000, 1024])
def test_fft_mag_bit_exact(handle, ora, length):
    rng = np.random.default_rng(length)
    fr = np.ascontiguousarray(rng.integers(-32768, 32768, (16, max(length, 1))).astype(np.int16)[:, :length])
    want = ora.fft_mag(fr) if length else np.zeros((16, 512), np.uint32)      # all-zero input -> all-zero spectrum
    assert np.array_equal(handle.fft_mag(fr), want)


# ---- get_mfcc ----------------------------------------------------------------------------------------
def test_mfcc_fixed_segment_config2_shape(handle, ora):
    """BASELINE config 2 geometry: segment [80, 8000) of every utterance, mid = 2048 -> 98 frames"""
    B, U = 96, 8000
    pcm = sr_b200.synth_pcm_host(B, U, 0x5EED0000)
    seg = np.tile(np.array([80, 8000], np.uint32), (B, 1))
    atap = np.zeros(B, sr_b200.ATAP_DTYPE)
    atap["mid_val"] = 2048
    got = handle.mfcc(pcm, seg, atap)
    assert (got["frm_num"] == 98).all()
    assert ob.ftr_equal(got, ora.mfcc_batch(pcm, seg, atap))


def test_mfcc_extreme_inputs_and_ragged_segments(handle, ora):
    rng = np.random.default_rng(9)
    B, U = 64, 12000
    pcm = rng.integers(0, 4096, (B, U)).astype(np.uint16)
    pcm[:8] = rng.integers(0, 65536, (8, U))           # > 12 bit: s16 window wrap, u32 energy wrap
    pcm[8] = 0
    pcm[9] = 65535
    pcm[10] = 2048                                     # all-zero frames: filter sums 0 -> log(0) pinned to 0
    starts = rng.integers(1, 4000, B) // 1 * 1
    lens = rng.integers(0, 130, B) * 80 + rng.integers(0, 80, B)      # ragged, some < 160, some > 119 frames
    ends = np.minimum(starts + lens, U)
    seg = np.stack([starts, ends], 1).astype(np.uint32)
    seg[11] = [ob.NULL, ob.NULL]
    seg[12] = [100, ob.NULL]
    seg[13] = [80, 80 + 160]                           # exactly one frame
    seg[14] = [80, 80 + 118 * 80 + 160]               # exactly vv_frm_max = 119 frames
    seg[16] = [80, 80 + 119 * 80 + 160]               # 120 frames: rejected, frm_num = 0 (MFCC.C:103-107)
    seg[15] = [1, 1 + 159]                             # one sample short of a frame
    atap = np.zeros(B, sr_b200.ATAP_DTYPE)
    atap["mid_val"] = rng.integers(0, 4096, B)
    got = handle.mfcc(pcm, seg, atap)
    valid = (seg[:, 0] != ob.NULL) & (seg[:, 1] != ob.NULL)
    want = ora.mfcc_batch(pcm[valid], seg[valid], atap[valid])
    assert ob.ftr_equal(got[valid], want)
    assert (got["frm_num"][~valid] == 0).all()
    assert (got["frm_num"] == 0).any() and (got["frm_num"] > 100).any() and (got["frm_num"] == 1).any()


@pytest.mark.parametrize("U", [8000, 8003, 5001, 16000])
def test_mfcc_unaligned_utterance_stride(handle, ora, U):
    """utterance strides that break the 16-byte phase of the bulk copies; last utterance ends at the buffer end"""
    B = 37
    pcm = sr_b200.synth_pcm_host(B, U, 0x77 + U)
    rng = np.random.default_rng(U)
    st = rng.integers(1, 900, B)
    en = np.minimum(st + rng.integers(160, 4000, B), U)
    en[-1] = U
    st[0] = 1
    seg = np.stack([st, en], 1).astype(np.uint32)
    atap = np.zeros(B, sr_b200.ATAP_DTYPE)
    atap["mid_val"] = 2000
    assert ob.ftr_equal(handle.mfcc(pcm, seg, atap), ora.mfcc_batch(pcm, seg, atap))


@pytest.mark.parametrize("B", [1, 2, 37, 147, 148, 149, 295, 296, 297, 445, 1000])
def test_mfcc_batch_sizes_around_the_cta_count(handle, ora, B):
    """mfcc_kernel hands utterances out with an atomic counter and ends a CTA's walk with an end marker in its staging
    ring; batch sizes around 1x / 2x / 3x the CTA count make every mix of (utterance, marker) land in a CTA's first
    ring slots, in either claim order (the first version lost an utterance that sat behind a marker). Run twice: the
    last CTA out re-arms the counter for the next launch."""
    U = 4003
    pcm = sr_b200.synth_pcm_host(B, U, 0x4A00 + B)
    rng = np.random.default_rng(B)
    st = rng.integers(1, 900, B)
    en = np.minimum(st + rng.integers(160, 3000, B), U)
    seg = np.stack([st, en], 1).astype(np.uint32)
    atap = np.zeros(B, sr_b200.ATAP_DTYPE)
    atap["mid_val"] = 2000
    want = ora.mfcc_batch(pcm, seg, atap)
    for _ in range(2):
        assert ob.ftr_equal(handle.mfcc(pcm, seg, atap), want)


def test_mfcc_segment_at_sample_zero_reads_previous_utterance(handle, ora):
    """start == 0 makes MFCC.C:119 read vc_dat[-1]; for b > 0 that is the last sample of utterance b-1 in a
    contiguous batch (same as the reference on the same memory); for b == 0 it is pinned to mid_val."""
    B, U = 5, 4000
    pcm = sr_b200.synth_pcm_host(B, U, 0x99)
    seg = np.tile(np.array([0, 1600], np.uint32), (B, 1))
    atap = np.zeros(B, sr_b200.ATAP_DTYPE)
    atap["mid_val"] = 2100
    got = handle.mfcc(pcm, seg, atap)
    # oracle on a buffer with one leading sample = mid_val: utterance 0 then sees x[-1] = mid
    flat = np.concatenate([[np.uint16(2100)], pcm.reshape(-1)])
    for b in range(B):
        view = flat[b * U: b * U + 1 + U].copy().reshape(1, -1)
        want = ora.mfcc_batch(view, np.array([[1, 1601]], np.uint32), atap[b:b + 1])
        assert ob.ftr_equal(got[b:b + 1], want), b


# ---- noise_atap + VAD ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["stm32_123", "stm32_456", "stm32_noise", "stm32_voice_123", "v1"])
def test_vad_on_board_captures_matches_golden(handle, name):
    pcm = CAPS[name].reshape(1, -1)
    atap = handle.noise_atap(pcm, 2400)
    assert atap.tobytes() == GOLD[name + "/atap"].tobytes()
    seg = handle.vad(pcm, atap)
    assert seg.reshape(-1).tolist() == GOLD[name + "/seg"].tolist()


def test_vad_extremes_bit_exact(handle, ora):
    rng = np.random.default_rng(4)
    B, U = 48, 8000
    pcm = sr_b200.synth_pcm_host(B, U, 0xABCD0000)
    pcm[0] = rng.integers(0, 4096, U)
    pcm[1] = rng.integers(0, 65536, U)
    pcm[2] = 2048
    pcm[3, 2400:] = np.where(np.arange(U - 2400) % 2 == 0, 0, 4095)
    pcm[4, 3000:7900] = rng.integers(0, 4096, 4900)
    for b in range(5, 16):                            # sparse out-of-band spikes: exercises the carried last_sig
        pcm[b] = 2048 + rng.integers(-3, 4, U)
        idx = rng.integers(2400, U, 60)
        pcm[b, idx] = np.where(rng.integers(0, 2, 60) == 1, 2048 + 500, 2048 - 500)
    atap = handle.noise_atap(pcm, 2400)
    seg = handle.vad(pcm, atap)
    for b in range(B):
        a = ora.noise_atap(pcm[b], 2400)
        assert a.tobytes() == atap[b:b + 1].tobytes(), b
        assert ora.vad(pcm[b], U, a).tolist() == seg[b].reshape(-1).tolist(), b
    # n_len not a multiple of 240 leaves atap untouched (VAD.C:33-36)
    keep = atap.copy()
    keep["mid_val"] = 7
    assert handle.noise_atap(pcm, 2401, keep.copy()).tobytes() == keep.tobytes()


@pytest.mark.parametrize("U,buf_len", [(8000, 8000), (8000, 7777), (16000, 16000), (40000, 40000), (8003, 8003), (400, 400), (160, 160)])
def test_vad_lengths(handle, ora, U, buf_len):
    B = 9
    pcm = sr_b200.synth_pcm_host(B, U, 0x1234 + U, 3 if U > 20000 else 1)
    n_len = 2400 if U >= 2400 else (240 if U >= 240 else 0)    # n_len = 0: noise_atap leaves atap untouched
    atap0 = np.zeros(B, sr_b200.ATAP_DTYPE)
    atap0["mid_val"], atap0["n_thl"], atap0["z_thl"], atap0["s_thl"] = 2000, 40, 2, 3000
    atap = handle.noise_atap(pcm, n_len, atap0.copy())
    seg = handle.vad(pcm, atap, buf_len)
    for b in range(B):
        a = ora.noise_atap(pcm[b], n_len, atap0[b:b + 1]) if n_len else atap0[b:b + 1].copy()
        assert a.tobytes() == atap[b:b + 1].tobytes()
        if buf_len > 160:
            assert ora.vad(pcm[b], buf_len, a).tolist() == seg[b].reshape(-1).tolist(), b
        else:
            assert (seg[b] == ob.NULL).all()


def test_noise_atap_windows_longer_than_one_chunk(handle, ora):
    """noise windows beyond the 2 560 samples the kernel stages per chunk take the direct global-memory path"""
    B, U = 40, 16000
    pcm = sr_b200.synth_pcm_host(B, U, 0x5150)
    for n_len in (2640, 4800, 7200, 15840):
        atap = handle.noise_atap(pcm, n_len)
        seg = handle.vad(pcm, atap)
        for b in range(B):
            a = ora.noise_atap(pcm[b], n_len)
            assert a.tobytes() == atap[b:b + 1].tobytes(), (n_len, b)
            assert ora.vad(pcm[b], U, a).tolist() == seg[b].reshape(-1).tolist(), (n_len, b)


def test_vad_and_mfcc_fuzz_with_arbitrary_atap(handle, ora):
    """random PCM shapes and ARBITRARY adaptive parameters handed straight to VAD / get_mfcc (n_thl > mid makes the lower
    band edge wrap, VAD.C:113; huge mid_val exercises the s32 casts of MFCC.C:110,119)"""
    rng = np.random.default_rng(77)
    B, U = 256, 4000
    pcm = np.zeros((B, U), np.uint16)
    for b in range(B):
        kind = b % 6
        base = int(rng.integers(0, 4096))
        if kind == 0:
            pcm[b] = rng.integers(0, 4096, U)
        elif kind == 1:
            pcm[b] = np.clip(base + rng.integers(-30, 31, U), 0, 65535)
            idx = rng.integers(0, U, 200)
            pcm[b, idx] = rng.integers(0, 4096, 200)
        elif kind == 2:
            t = np.arange(U)
            pcm[b] = np.clip(2048 + 1500 * np.sin(t * rng.uniform(0.01, 1.5)) * (rng.random(U) < 0.7), 0, 4095)
        elif kind == 3:
            pcm[b] = rng.integers(0, 65536, U)
        elif kind == 4:
            pcm[b] = np.where(rng.random(U) < 0.05, rng.integers(0, 4096, U), base)
        else:
            blocks = rng.integers(0, 2, U // 80 + 1).repeat(80)[:U]
            pcm[b] = np.where(blocks == 1, rng.integers(0, 4096, U), base)
    atap = np.zeros(B, sr_b200.ATAP_DTYPE)
    atap["mid_val"] = rng.integers(0, 4096, B)
    atap["n_thl"] = rng.integers(0, 3000, B)          # often > mid_val: b_thl wraps
    atap["z_thl"] = rng.integers(0, 12, B)
    atap["s_thl"] = rng.integers(0, 200000, B)
    atap["mid_val"][:8] = rng.integers(60000, 2 ** 32, 8, dtype=np.uint64).astype(np.uint32)
    seg = handle.vad(pcm, atap)
    for b in range(B):
        assert ora.vad(pcm[b], U, atap[b:b + 1]).tolist() == seg[b].reshape(-1).tolist(), b
    st = rng.integers(1, 2000, B)
    en = np.minimum(st + rng.integers(160, 1800, B), U)
    sg = np.stack([st, en], 1).astype(np.uint32)
    assert ob.ftr_equal(handle.mfcc(pcm, sg, atap), ora.mfcc_batch(pcm, sg, atap))


def test_vad_threshold_corner_cases(handle, ora):
    """band edges at the corners of the packed 16-bit compare path: a_thl == 0, a_thl > 0xFFFF, b_thl == 0, b_thl wrapped
    (VAD.C:112-113 are u32), mid_val beyond the sample range, thresholds equal to sample values; full-range u16 samples"""
    rng = np.random.default_rng(2024)
    combos = [(0, 0), (0, 1), (1, 1), (5, 5), (5, 6), (100, 100), (2048, 0), (2048, 2048), (2048, 2049), (65535, 0),
              (65535, 1), (65000, 535), (65000, 536), (65000, 2000), (65535, 65535), (30000, 30000), (30000, 40000),
              (65536, 1), (65536, 65535), (70000, 5000), (70000, 4464), (70000, 4465), (2 ** 32 - 5, 10), (2 ** 32 - 1, 0),
              (2 ** 32 - 1, 1), (2 ** 31, 65535), (1, 0), (1, 2), (32768, 32768), (32768, 32767)]
    B, U = len(combos) * 3, 4000
    pcm = np.zeros((B, U), np.uint16)
    atap = np.zeros(B, sr_b200.ATAP_DTYPE)
    for i, (mid, n) in enumerate(combos * 3):
        kind = i // len(combos)
        centre = min(mid, 65535)
        if kind == 0:
            pcm[i] = rng.integers(0, 65536, U)
        elif kind == 1:                                  # hover around the band edges so >=, < and == all occur
            pcm[i] = np.clip(centre + rng.integers(-3, 4, U) + rng.choice([-n, 0, n], U), 0, 65535)
        else:                                            # sparse excursions: long in-band runs between markers
            pcm[i] = centre
            idx = rng.integers(0, U, 120)
            pcm[i, idx] = rng.choice([0, 65535, max(centre - n, 0), min(centre + n, 65535), max(centre - n - 1, 0)], 120)
        atap[i] = (mid, n, int(rng.integers(0, 6)), int(rng.integers(0, 400000)))
    seg = handle.vad(pcm, atap)
    for b in range(B):
        assert ora.vad(pcm[b], U, atap[b:b + 1]).tolist() == seg[b].reshape(-1).tolist(), (b, combos[b % len(combos)])


def test_dtw_limit_batch_and_drop_in_symbol(handle, ora):
    """dtw_limit (DTW.C:76-109): every lattice point of a few (I, M) shapes, and the reference-named symbol after dtw()"""
    L = sr_b200.lib()
    pts = []
    for (I, M) in ((36, 34), (50, 100), (100, 50), (119, 60), (1, 1), (8, 15)):
        for x in range(0, I + 3):
            for y in range(0, M + 3):
                pts.append((x, y, I, M))
    a = np.array(pts, np.uint16)
    out = np.zeros(len(pts), np.uint8)
    cols = [np.ascontiguousarray(a[:, k]) for k in range(4)]
    rc = L.sr_dtw_limit_batch(handle._h, *[c.ctypes.data_as(C.c_void_p) for c in cols], len(pts), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    po = ob.port()
    want = np.array([po.lib.sro_dtw_limit(int(x), int(y), int(I), int(M)) for x, y, I, M in pts], np.uint8)
    assert np.array_equal(out, want) and 0 < want.sum() < len(pts)
    f = sr_b200.synth_ftr_host(2, 0xABC, 30, 40).view(sr_b200.FTR_DTYPE).reshape(-1)
    L.dtw(f[0:1].ctypes.data_as(C.c_void_p), f[1:2].ctypes.data_as(C.c_void_p))
    I, M = int(f["frm_num"][0]), int(f["frm_num"][1])
    for x, y in ((1, 1), (5, 20), (20, 5), (I, M), (2, 9)):
        assert L.dtw_limit(x, y) == po.lib.sro_dtw_limit(x, y, I, M)
    # ... and against the reference's OWN dtw_limit, which reads the file statics its dtw() left behind (DTW.C:65-68,
    # 130-131, 141-142): call dtw() in libref, then compare every lattice point of that shape with the drop-in symbol
    if ob.have_ref():
        r = ob.ref()
        for seed, lo, hi in ((0xABC, 30, 40), (0x51, 50, 100), (0x52, 8, 15), (0x53, 100, 119)):
            f = sr_b200.synth_ftr_host(2, seed, lo, hi).view(sr_b200.FTR_DTYPE).reshape(-1)
            if seed == 0x51:
                f["frm_num"][0], f["frm_num"][1] = 50, 100        # the 2:1 edge of the guard (DTW.C:133)
            pa, pb = f[0:1].ctypes.data_as(C.c_void_p), f[1:2].ctypes.data_as(C.c_void_p)
            assert L.dtw(pa, pb) == r.lib.dtw(pa, pb)
            I, M = int(f["frm_num"][0]), int(f["frm_num"][1])
            mine = [L.dtw_limit(x, y) for x in range(0, I + 3) for y in range(0, M + 3)]
            theirs = [r.lib.dtw_limit(C.c_uint16(x), C.c_uint16(y)) for x in range(0, I + 3) for y in range(0, M + 3)]
            assert mine == theirs and 0 < sum(theirs) < len(theirs), (seed, I, M)


def test_dtw_fuzz_many_pairs(handle, ora):
    """20 000 random pairs over all length combinations, realistic and adversarial value ranges, T not a multiple of 32"""
    rng = np.random.default_rng(78)
    B, T = 400, 50
    ftr = sr_b200.synth_ftr_host(B, 0xF00D, 1, 119).view(sr_b200.FTR_DTYPE).reshape(-1).copy()
    ftr["mfcc_dat"][::3] = rng.integers(-32768, 32768, ftr["mfcc_dat"][::3].shape)      # every third: full-range values
    ftr["mfcc_dat"][1::7] //= 64                                                          # small values: many equal distances (ties)
    bank = sr_b200.synth_ftr_host(T, 0xFEED, 1, 119, stride=2860)
    bank[::5, 4:] = rng.integers(0, 256, bank[::5, 4:].shape)
    handle.set_bank(bank, T, 2860)
    want, _ = ora.dtw_batch(ftr, bank, T, 2860)
    score, bi, bd = handle.dtw(ftr)
    assert np.array_equal(score, want)


# ---- dtw ---------------------------------------------------------------------------------------------
def test_dtw_all_lengths_bit_exact(handle, ora):
    B, T = 150, 70
    ftr = sr_b200.synth_ftr_host(B, 0xD7A00000, 1, 119).view(sr_b200.FTR_DTYPE).reshape(-1)
    bank = sr_b200.synth_ftr_host(T, 0xD7A10000, 1, 119, stride=4096)
    bank[5, 0] = 0                                     # save_sign != 12345
    bank[17, :2] = 0xFF
    handle.set_bank(bank, T, 4096)
    want, _ = ora.dtw_batch(ftr, bank, T, 4096, check_sign=0)
    score, bi, bd = handle.dtw(ftr, flags=0)
    assert np.array_equal(score, want)
    key = (want.astype(np.uint64) << np.uint64(32)) | np.arange(T, dtype=np.uint64)[None, :]
    k = key.min(axis=1)
    assert np.array_equal(bi, (k & np.uint64(0xFFFFFFFF)).astype(np.uint32)) and np.array_equal(bd, (k >> np.uint64(32)).astype(np.uint32))
    want_s, _ = ora.dtw_batch(ftr, bank, T, 4096, check_sign=1)
    score_s, _, _ = handle.dtw(ftr, flags=sr_b200.DTW_CHECK_SIGN)
    assert np.array_equal(score_s, want_s) and (score_s[:, 5] == ob.NULL).all()


def test_dtw_200_templates_mixed_save_sign(handle, ora):
    """configs[2] bank width: T = 200 = 6 full template tiles + a remainder launch, a third of the slots erased / unsigned
    (main.c:283), against the reference's own dtw; argmin = strict '<' first-wins scan over the 200 slots (main.c:285-289)"""
    B, T = 96, 200
    fin = sr_b200.synth_ftr_host(B, 0xD7A00000, 50, 100).view(sr_b200.FTR_DTYPE).reshape(-1)
    bank = sr_b200.synth_ftr_host(T, 0xD7A10000, 50, 100, stride=4096)
    rng = np.random.default_rng(200)
    bad = rng.random(T) < 0.33
    bank[bad, 0:2] = 0xFF                                  # erased flash: save_sign != 12345
    bank[~bad, 0], bank[~bad, 1] = 12345 & 0xFF, 12345 >> 8
    bank[7] = bank[3]                                      # an exact duplicate: first wins
    handle.set_bank(bank, T, 4096)
    score, bi, bd = handle.dtw(fin, flags=sr_b200.DTW_CHECK_SIGN)
    want, _ = ora.dtw_batch(fin, bank, T, 4096, check_sign=1)
    assert np.array_equal(score, want)
    assert (score[:, bad] == ob.NULL).all() and (score[:, ~bad] != ob.NULL).any()
    key = (want.astype(np.uint64) << np.uint64(32)) | np.arange(T, dtype=np.uint64)[None, :]
    k = key.min(axis=1)
    assert np.array_equal(bi, (k & np.uint64(0xFFFFFFFF)).astype(np.uint32)) and np.array_equal(bd, (k >> np.uint64(32)).astype(np.uint32))
    # same bank without the save_sign check (dtw() itself never looks at it)
    score2, _, _ = handle.dtw(fin)
    want2, _ = ora.dtw_batch(fin, bank, T, 4096, check_sign=0)
    assert np.array_equal(score2, want2)


def test_dtw_empty_feature_sets_are_deterministic(handle, ora):
    """frm_num == 0 on both sides passes the 2:1 guard (DTW.C:133) and the do-while still reads rows 0 and 1 of both
    structs (DTW.C:146-160): the result is defined by the struct bytes, not by whatever was staged before"""
    f = sr_b200.synth_ftr_host(6, 0xE0, 20, 30).view(sr_b200.FTR_DTYPE).reshape(-1).copy()
    f["frm_num"][:3] = 0                                   # rows stay: the reference reads them
    bank = np.zeros((4, 4096), np.uint8)
    bank[:, :2860] = f[[0, 1, 3, 4]].view(np.uint8).reshape(4, 2860)
    handle.set_bank(bank, 4, 4096)
    for _ in range(2):                                     # twice: the second run sees different leftovers in shared memory
        score, _, _ = handle.dtw(f)
        want, _ = ora.dtw_batch(f, bank, 4, 4096)
        assert np.array_equal(score, want)
        handle.dtw(sr_b200.synth_ftr_host(64, 0xE1, 90, 119).view(sr_b200.FTR_DTYPE).reshape(-1))


@pytest.mark.parametrize("T,B,fr", [(1, 70, (1, 119)), (5, 333, (20, 45)), (20, 1500, (23, 43)), (32, 257, (50, 100)),
                                    (33, 640, (1, 119)), (70, 200, (30, 119)), (200, 300, (50, 100))])
def test_dtw_dynamic_pair_scheduling_equals_static_and_reference(ora, T, B, fr):
    """sr_dtw_dyn.cu (pairs pulled dynamically from a ring of staged utterances) == the static kernel == the reference's
    dtw, scores and first-wins argmin, over bank widths around the 32-template tile, all frame counts, mixed save_sign,
    the 2:1 guard, garbage headers; run twice so the second launch sees a dirty ring"""
    h = sr_b200.Handle(0)
    fin = sr_b200.synth_ftr_host(B, 0xD100 + T, fr[0], fr[1]).view(sr_b200.FTR_DTYPE).reshape(-1).copy()
    bank = sr_b200.synth_ftr_host(T, 0xD200 + T, fr[0], fr[1], stride=4096)
    rng = np.random.default_rng(T)
    bad = rng.random(T) < 0.2
    bank[bad, 0:2] = 0xFF
    if T > 3:
        bank[2, 2:4] = (200, 0)                            # frm_num 200 > vv_frm_max: never walked
        fin["frm_num"][1] = 0
        fin["frm_num"][2] = 300
    h.set_bank(bank, T, 4096)
    res = {}
    for v in (0, 1, 1):
        h.set_dtw_variant(v)
        res[v] = h.dtw(fin, flags=sr_b200.DTW_CHECK_SIGN)
        assert all(np.array_equal(a, b) for a, b in zip(res[v], res[0])), v
    ok = fin["frm_num"] <= 119
    want, _ = ora.dtw_batch(fin[ok], bank, T, 4096, check_sign=1)
    valid = np.ones(T, bool)
    if T > 3:
        valid[2] = False                                   # the reference would read past the struct: not comparable
    assert np.array_equal(res[1][0][ok][:, valid], want[:, valid])
    # through the recognise path (status gate: failed utterances never reach dtw)
    U = 8000
    pcm = sr_b200.synth_pcm_host(128, U, 0x77)
    pcm[::7] = 2048
    tb, _ = h.enrol(sr_b200.synth_pcm_host(min(T, 24), U, 0x7E3A0000), 2400)
    h.set_bank(tb, tb.shape[0], 4096)
    h.set_dtw_variant(0)
    a = h.recognise(pcm, 2400)
    h.set_dtw_variant(1)
    b = h.recognise(pcm, 2400)
    for k in ("score", "best_idx", "best_dis", "cmd", "status"):
        if k == "score":
            okr = a["status"] == 0
            assert np.array_equal(a[k][okr], b[k][okr])
        else:
            assert np.array_equal(a[k], b[k]), k
    h.close()


def test_dtw_extreme_values_wrap(handle, ora):
    """|dif| up to 65535 per dimension: the u32 accumulation of get_dis wraps (DTW.C:56)"""
    rng = np.random.default_rng(2)
    B, T = 40, 33
    ftr = sr_b200.synth_ftr_host(B, 1, 30, 60).view(sr_b200.FTR_DTYPE).reshape(-1).copy()
    ftr["mfcc_dat"] = rng.integers(-32768, 32768, ftr["mfcc_dat"].shape)
    bank = sr_b200.synth_ftr_host(T, 2, 30, 60)
    bank[:, 4:] = rng.integers(0, 256, bank[:, 4:].shape)
    handle.set_bank(bank, T, 2860)
    want, _ = ora.dtw_batch(ftr, bank, T, 2860)
    score, _, _ = handle.dtw(ftr)
    assert np.array_equal(score, want)
    a = rng.integers(-32768, 32768, (500, 12)).astype(np.int16)
    b = rng.integers(-32768, 32768, (500, 12)).astype(np.int16)
    assert np.array_equal(handle.get_dis(a, b), ora.get_dis(a, b))


@pytest.mark.parametrize("r", [10, 3, 15])
def test_dtw_band_extension_vs_own_dp_oracle(handle, r):
    """Sakoe-Chiba DP (BASELINE configs[2]); NOT in the reference -> checked against our own CPU DP (parity unpinned)"""
    B, T = 70, 45
    ftr = sr_b200.synth_ftr_host(B, 0xD7A00000, 1, 119).view(sr_b200.FTR_DTYPE).reshape(-1)
    ftr2 = sr_b200.synth_ftr_host(B, 0xD7A00100, 50, 100).view(sr_b200.FTR_DTYPE).reshape(-1)
    ftr = np.concatenate([ftr, ftr2])
    bank = sr_b200.synth_ftr_host(T, 0xD7A10000, 40, 110, stride=4096)
    handle.set_bank(bank, T, 4096)
    want, cells = ob.port().dtw_batch(ftr, bank, T, 4096, band_r=r, nthreads=4)
    score, bi, bd = handle.dtw(ftr, flags=sr_b200.DTW_BAND, band_r=r)
    assert np.array_equal(score, want) and cells > 0
    assert (want != ob.NULL).any() and (want == ob.NULL).any()
    key = (want.astype(np.uint64) << np.uint64(32)) | np.arange(T, dtype=np.uint64)[None, :]
    assert np.array_equal(bd, (key.min(axis=1) >> np.uint64(32)).astype(np.uint32))


# ---- spch_recg ---------------------------------------------------------------------------------------
def test_recognise_matches_golden_synthetic(handle):
    pcm = sr_b200.synth_pcm_host(24, 8000, 0x5EED0000)
    bank = GOLD["synth/bank"]
    handle.set_bank(bank, 8, 4096)
    out = handle.recognise(pcm, 2400)
    _cmp_recog(out, {k: GOLD["synth/" + k] for k in ("seg_off", "score", "best_idx", "best_dis", "cmd", "status", "ftr")})
    pcm5 = sr_b200.synth_pcm_host(4, 40000, 0x5EED5000, 3)
    out5 = handle.recognise(pcm5, 2400)
    _cmp_recog(out5, {k: GOLD["synth5/" + k] for k in ("seg_off", "score", "best_idx", "best_dis", "cmd", "status", "ftr")})


def test_recognise_mixed_failures_vs_oracle(handle, ora):
    rng = np.random.default_rng(6)
    B, U, T = 200, 8000, 11
    pcm = sr_b200.synth_pcm_host(B, U, 0xC0FFEE00)
    pcm[3] = 2048                                     # VAD fail
    pcm[4, 2500:7990] = rng.integers(0, 4096, 5490)   # segment never closes: VAD fail
    pcm[7] = rng.integers(0, 4096, U)
    tpl = sr_b200.synth_pcm_host(T, U, 0x7E3A0000)
    handle.set_bank(np.zeros((1, 4096), np.uint8), 0, 4096)
    e = handle.recognise(tpl, 2400, want=("ftr", "status"))
    bank = sr_b200.make_bank(e["ftr"], valid=[1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1])
    handle.set_bank(bank, T, 4096)
    out = handle.recognise(pcm, 2400)
    ref = ora.recognise_batch(pcm, 2400, bank, T, 4096)
    _cmp_recog(out, ref)
    assert set(out["status"].tolist()) >= {0, 1}
    # empty bank: idx 0, dis_max (main.c:276-278)
    handle.set_bank(bank, 0, 4096)
    o2 = handle.recognise(pcm[:5], 2400, want=("best_idx", "best_dis", "cmd"))
    assert (o2["best_idx"] == 0).all() and (o2["best_dis"] == ob.NULL).all()


def test_recognise_2s_buffers_and_long_segments(handle, ora):
    """the reference's native 2 s buffer; words long enough to exceed vv_frm_max -> MFCC fail (status 2)"""
    B, U = 16, 16000
    pcm = sr_b200.synth_pcm_host(B, U, 0x2222, 2)
    rng = np.random.default_rng(8)
    pcm[0, 3000:13500] = 2048 + (1200 * np.sin(np.arange(10500) * 0.3)).astype(np.int64) + rng.integers(-50, 50, 10500)
    tpl = sr_b200.synth_pcm_host(4, 8000, 0x7E3A0000)
    handle.set_bank(np.zeros((1, 4096), np.uint8), 0, 4096)
    bank = sr_b200.make_bank(handle.recognise(tpl, 2400, want=("ftr",))["ftr"])
    handle.set_bank(bank, 4, 4096)
    out = handle.recognise(pcm, 2400)
    _cmp_recog(out, ora.recognise_batch(pcm, 2400, bank, 4, 4096))
    assert out["status"][0] == 2


# ---- the reference's own entry points (batch of 1) -----------------------------------------------------
def test_reference_named_entry_points(handle, ora):
    L = sr_b200.lib()
    pcm = CAPS["stm32_123"].copy()
    atap = np.zeros(1, sr_b200.ATAP_DTYPE)
    L.noise_atap(pcm.ctypes.data_as(C.c_void_p), 2400, atap.ctypes.data_as(C.c_void_p))
    assert atap.tobytes() == GOLD["stm32_123/atap"].tobytes()
    vv = (sr_b200.ValidTag * 3)()
    L.VAD(pcm.ctypes.data_as(C.c_void_p), 16000, vv, atap.ctypes.data_as(C.c_void_p))
    base = pcm.ctypes.data
    offs = [((v.start - base) // 2 if v.start else ob.NULL, (v.end - base) // 2 if v.end else ob.NULL) for v in vv]
    assert [x for p in offs for x in p] == GOLD["stm32_123/seg"].tolist()
    f = np.zeros(2, sr_b200.FTR_DTYPE)
    f["save_sign"] = 4321
    L.get_mfcc(C.byref(vv[0]), f[0:1].ctypes.data_as(C.c_void_p), atap.ctypes.data_as(C.c_void_p))
    L.get_mfcc(C.byref(vv[1]), f[1:2].ctypes.data_as(C.c_void_p), atap.ctypes.data_as(C.c_void_p))
    assert ob.ftr_equal(f[0:1], GOLD["stm32_123/ftr0"]) and ob.ftr_equal(f[1:2], GOLD["stm32_123/ftr1"])
    assert (f["save_sign"] == 4321).all()              # MFCC.C never writes save_sign
    d01 = L.dtw(f[0:1].ctypes.data_as(C.c_void_p), f[1:2].ctypes.data_as(C.c_void_p))
    d00 = L.dtw(f[0:1].ctypes.data_as(C.c_void_p), f[0:1].ctypes.data_as(C.c_void_p))
    assert [[d00, d01]] == GOLD["stm32_123/dtw"][:1].tolist()
    r0, r1 = f["mfcc_dat"][0][:12].copy(), f["mfcc_dat"][1][:12].copy()
    assert L.get_dis(r0.ctypes.data_as(C.c_void_p), r1.ctypes.data_as(C.c_void_p)) == ora.get_dis(r0.reshape(1, 12), r1.reshape(1, 12))[0]
    fr = (pcm[4000:4160].astype(np.int32) - 2213).astype(np.int16)
    p = L.fft(fr.ctypes.data_as(C.c_void_p), 160)
    mag = np.ctypeslib.as_array(p, shape=(1024,))[:512].copy()
    assert np.array_equal(mag, ora.fft_mag(fr.reshape(1, -1))[0])
    assert not L.fft(fr.ctypes.data_as(C.c_void_p), 1025)  # MFCC.C:32-35


# ---- full-size, size-independent properties ------------------------------------------------------------
def test_large_batch_shard_invariance_and_sampled_parity(handle, ora):
    """BASELINE-size run (16 384 x 1 s here to bound host RAM/time): results do not depend on how the batch is
    sharded (what the multi-GPU split relies on), enrolment then recognition of the same audio gives
    distance 0 against its own template, and a random sample agrees with the oracle bit-for-bit."""
    B, U, T = 16384, 8000, 20
    pcm = sr_b200.synth_pcm_host(B, U, 0x5EED0000)
    handle.set_bank(np.zeros((1, 4096), np.uint8), 0, 4096)
    enrol = handle.recognise(pcm[:T], 2400, want=("ftr", "status"))
    assert (enrol["status"] == 0).all()
    bank = sr_b200.make_bank(enrol["ftr"])
    handle.set_bank(bank, T, 4096)
    full = handle.recognise(pcm, 2400)
    assert (full["best_idx"][:T] == np.arange(T)).all() and (full["best_dis"][:T] == 0).all()
    cut = 5000
    a = handle.recognise(pcm[:cut], 2400)
    b = handle.recognise(pcm[cut:], 2400)
    for k in ("seg_off", "score", "best_idx", "best_dis", "cmd", "status"):
        assert np.array_equal(full[k], np.concatenate([a[k], b[k]])), k
    assert ob.ftr_equal(full["ftr"], np.concatenate([a["ftr"], b["ftr"]]))
    idx = np.random.default_rng(0).choice(B, 192, replace=False)
    ref = ora.recognise_batch(np.ascontiguousarray(pcm[idx]), 2400, bank, T, 4096)
    _cmp_recog({k: v[idx] for k, v in full.items()}, ref)
    assert (full["status"] == 0).mean() > 0.99


# ---- device-pointer variants on a torch stream ----------------------------------------------------------
def test_device_pointer_api_on_torch_stream(ora):
    import torch
    dev = torch.device("cuda:0")
    B, U, T = 300, 8000, 9
    h = sr_b200.Handle(0)
    st = torch.cuda.Stream(dev)
    h.set_stream(st.cuda_stream)
    pcm_h = sr_b200.synth_pcm_host(B, U, 0x5151)
    with torch.cuda.stream(st):
        pcm = torch.empty((B, U), dtype=torch.int16, device=dev)
        sr_b200.synth_pcm_dev(pcm.data_ptr(), B, U, 0x5151, 1, st.cuda_stream)     # device generator == host generator
        tpl = torch.empty((T, U), dtype=torch.int16, device=dev)
        sr_b200.synth_pcm_dev(tpl.data_ptr(), T, U, 0x7E3A0000, 1, st.cuda_stream)
        ftr_t = torch.zeros((T, 2860), dtype=torch.uint8, device=dev)
        h.set_bank_dev(0, 0, 4096)
        h.recognise_dev(tpl.data_ptr(), U, T, 2400, ftr=ftr_t.data_ptr())
        bank = torch.full((T, 4096), 255, dtype=torch.uint8, device=dev)
        bank[:, :2860] = ftr_t
        bank[:, 0] = 12345 & 0xFF
        bank[:, 1] = 12345 >> 8
        h.set_bank_dev(bank.data_ptr(), T, 4096)
        score = torch.zeros((B, T), dtype=torch.int32, device=dev)
        bidx = torch.zeros(B, dtype=torch.int32, device=dev)
        bdis = torch.zeros(B, dtype=torch.int32, device=dev)
        cmd = torch.zeros(B, dtype=torch.int32, device=dev)
        status = torch.zeros(B, dtype=torch.uint8, device=dev)
        seg = torch.zeros((B, 6), dtype=torch.int32, device=dev)
        h.recognise_dev(pcm.data_ptr(), U, B, 2400, seg_off=seg.data_ptr(), score=score.data_ptr(),
                        best_idx=bidx.data_ptr(), best_dis=bdis.data_ptr(), cmd=cmd.data_ptr(), status=status.data_ptr())
    st.synchronize()
    assert np.array_equal(pcm.cpu().numpy().view(np.uint16), pcm_h)
    bank_h = bank.cpu().numpy()
    ref = ora.recognise_batch(pcm_h, 2400, bank_h, T, 4096)
    assert np.array_equal(score.cpu().numpy().view(np.uint32), ref["score"])
    assert np.array_equal(seg.cpu().numpy().view(np.uint32).reshape(-1), ref["seg_off"].reshape(-1))
    assert np.array_equal(bidx.cpu().numpy().view(np.uint32), ref["best_idx"])
    assert np.array_equal(bdis.cpu().numpy().view(np.uint32), ref["best_dis"])
    assert np.array_equal(cmd.cpu().numpy().view(np.uint32), ref["cmd"])
    assert np.array_equal(status.cpu().numpy(), ref["status"])
    assert h.launch_count() >= 6
    h.close()


def test_enrol_and_get_mdl(handle, ora):
    """save_mdl (main.c:121-138 + Flash.C:17-67) as a batch, and the reference's template averaging get_mdl"""
    B, U = 40, 8000
    pcm = sr_b200.synth_pcm_host(B, U, 0x7E3A0000)
    pcm[5] = 2048                                     # VAD_fail: slot stays erased
    bank, status = handle.enrol(pcm, 2400)
    ref = ora.recognise_batch(pcm, 2400, None, 0, 4096)
    assert np.array_equal(status, ref["status"]) and status[5] == 1
    want = sr_b200.make_bank(ref["ftr"])
    want[ref["status"] != 0] = 0xFF                   # save_ftr_mdl is never reached: the slot stays erased
    assert np.array_equal(bank, want) and (bank[5] == 0xFF).all()
    f1 = sr_b200.synth_ftr_host(64, 0xAA00, 1, 59).view(sr_b200.FTR_DTYPE).reshape(-1)
    f2 = sr_b200.synth_ftr_host(64, 0xBB00, 1, 59).view(sr_b200.FTR_DTYPE).reshape(-1)
    want_m, want_d = ora.get_mdl(f1, f2)
    pre = np.zeros(64, sr_b200.FTR_DTYPE)
    pre["save_sign"] = 777
    got_m, got_d = handle.get_mdl(f1, f2, pre)
    assert np.array_equal(got_d, want_d) and ob.ftr_equal(got_m, want_m) and (got_m["save_sign"] == 777).all()
    # long paths: frm_num clamps at 119 instead of the reference's out-of-bounds writes (port == kernel)
    g1 = sr_b200.synth_ftr_host(16, 0xCC00, 100, 119).view(sr_b200.FTR_DTYPE).reshape(-1)
    g2 = sr_b200.synth_ftr_host(16, 0xDD00, 100, 119).view(sr_b200.FTR_DTYPE).reshape(-1)
    pm, pd = ob.port().get_mdl(g1, g2)
    gm, gd = handle.get_mdl(g1, g2)
    assert np.array_equal(gd, pd) and ob.ftr_equal(gm, pm) and (gm["frm_num"] == 119).any()


@pytest.mark.parametrize("chunk", [80, 800, 777])
def test_streaming_equals_batch(handle, ora, chunk):
    """lock-step chunked capture (config 5 shape: 5 s streams, 3 words): the union of the streaming events equals
    the batch VAD on the finished buffers, and every event equals get_mfcc + dtw of that segment"""
    S, L, T = 24, 40000, 8
    pcm = sr_b200.synth_pcm_host(S, L, 0x5EED5000, 3)
    pcm[3] = 2048                                     # a silent stream: no event
    bank = GOLD["synth/bank"]
    handle.set_bank(bank, T, 4096)
    pool = sr_b200.StreamPool(handle, S, L, 2400)
    events, latest = [], {}
    pin_ptr = None
    if chunk == 800:                                   # pinned capture array: chunks are read zero-copy, strided ([S][L] rows)
        arr, pin_ptr = sr_b200.host_alloc_dev(0, S * L * 2)
        arr.view(np.uint16).reshape(S, L)[:] = pcm
    for n0 in range(0, L, chunk):
        c = np.ascontiguousarray(pcm[:, n0:n0 + chunk])
        evs = pool.push(pin_ptr + 2 * n0, c.shape[1], L) if pin_ptr else pool.push(c)
        for e in evs:
            # the segment closes with the chunk that delivers sample end+879 (last sample of the closing frame)
            assert n0 <= e["end"] + 879 < n0 + c.shape[1]
        events += evs
    seg, atap = pool.segments()
    pool.close()
    if pin_ptr:
        sr_b200.host_free(pin_ptr)
    batch_atap = handle.noise_atap(pcm, 2400)
    assert atap.tobytes() == batch_atap.tobytes()
    batch_seg = handle.vad(pcm, batch_atap)
    assert np.array_equal(seg, batch_seg)
    closed = [(s, k) for s in range(S) for k in range(3) if batch_seg[s, k, 1] != ob.NULL]
    assert sorted((e["stream"], e["segment"]) for e in events) == closed and len(closed) >= 3 * (S - 1) - 2
    for e in events:
        s, k = e["stream"], e["segment"]
        assert (e["start"], e["end"]) == tuple(batch_seg[s, k])
        f = ora.mfcc_batch(pcm[s:s + 1], batch_seg[s, k].reshape(1, 2), batch_atap[s:s + 1])
        assert e["frm_num"] == int(f["frm_num"][0]) and e["status"] == 0
        sc, _ = ora.dtw_batch(f, bank, T, 4096, check_sign=1)
        key = (sc[0].astype(np.uint64) << np.uint64(32)) | np.arange(T, dtype=np.uint64)
        assert e["best_dis"] == int(key.min() >> np.uint64(32)) and e["best_idx"] == int(key.min() & np.uint64(0xFFFFFFFF))
        assert e["cmd"] == e["best_idx"] // 4


def _check_stream_events(handle, ora, pcm, bank, T, events, seg, atap):
    S = pcm.shape[0]
    batch_atap = handle.noise_atap(pcm, 2400)
    assert atap.tobytes() == batch_atap.tobytes()
    batch_seg = handle.vad(pcm, batch_atap)
    assert np.array_equal(seg, batch_seg)
    closed = [(s, k) for s in range(S) for k in range(3) if batch_seg[s, k, 1] != ob.NULL]
    assert sorted((e["stream"], e["segment"]) for e in events) == closed and len(closed) >= 2 * S
    want = handle.recognise(pcm, 2400, want=("best_idx", "best_dis", "cmd", "status"))     # segment 0 == the batch call
    for e in events:
        s, k = e["stream"], e["segment"]
        assert (e["start"], e["end"]) == tuple(batch_seg[s, k])
        if k == 0:
            assert (e["best_idx"], e["best_dis"], e["cmd"], e["status"]) == tuple(int(want[q][s]) for q in ("best_idx", "best_dis", "cmd", "status"))
    for e in events[:: max(1, len(events) // 12)]:           # a sample against the oracle, segment by segment
        s, k = e["stream"], e["segment"]
        f = ora.mfcc_batch(pcm[s:s + 1], batch_seg[s, k].reshape(1, 2), batch_atap[s:s + 1])
        assert e["frm_num"] == int(f["frm_num"][0]) and e["status"] == 0
        sc, _ = ora.dtw_batch(f, bank, T, 4096, check_sign=1)
        key = (sc[0].astype(np.uint64) << np.uint64(32)) | np.arange(T, dtype=np.uint64)
        assert e["best_dis"] == int(key.min() >> np.uint64(32)) and e["best_idx"] == int(key.min() & np.uint64(0xFFFFFFFF))


def test_streaming_ragged_arrival_equals_batch(handle, ora):
    """every stream advances at its own pace (random chunk lengths incl. 0 and odd ones, some streams far ahead of
    others, a stream that starts late): events and final segments still equal the batch results"""
    S, L, T = 40, 40000, 8
    pcm = sr_b200.synth_pcm_host(S, L, 0x5EED6000, 3)
    bank = GOLD["synth/bank"]
    handle.set_bank(bank, T, 4096)
    pool = sr_b200.StreamPool(handle, S, L, 2400)
    rng = np.random.default_rng(77)
    pos = np.zeros(S, np.int64)
    events, pushes = [], 0
    while (pos < L).any():
        lens = rng.choice([0, 1, 79, 80, 81, 160, 333, 800, 1601, 4000], S).astype(np.int64)
        lens[5] = 0 if pushes < 30 else lens[5]            # stream 5 starts late
        lens = np.minimum(lens, L - pos)
        w = int(lens.max())
        if w == 0:
            continue
        chunk = np.zeros((S, w), np.uint16)
        for s in range(S):
            chunk[s, :lens[s]] = pcm[s, pos[s]:pos[s] + lens[s]]
        evs = pool.push_ragged(chunk, lens)
        for e in evs:                                       # an event appears with the push that delivers sample end+879
            s = e["stream"]
            assert pos[s] <= e["end"] + 879 < pos[s] + lens[s], (e, pos[s], lens[s])
        events += evs
        pos += lens
        pushes += 1
    seg, atap = pool.segments()
    pool.close()
    _check_stream_events(handle, ora, pcm, bank, T, events, seg, atap)


def test_streaming_small_event_buffer_keeps_events(handle, ora):
    """max_events smaller than what a push closes: nothing is lost, the rest comes with later pushes / sr_streams_fetch"""
    S, L, T = 32, 16000, 8
    pcm = sr_b200.synth_pcm_host(S, L, 0x5EED7000, 1)
    handle.set_bank(GOLD["synth/bank"], T, 4096)
    pool = sr_b200.StreamPool(handle, S, L, 2400)
    full = []
    for n0 in range(0, L, 4000):
        full += pool.push(np.ascontiguousarray(pcm[:, n0:n0 + 4000]))
    pool.reset()
    got = []
    for n0 in range(0, L, 4000):
        evs = pool.push(np.ascontiguousarray(pcm[:, n0:n0 + 4000]), max_events=3)
        assert len(evs) <= 3
        got += evs
    assert pool.pending() == len(full) - len(got) > 0
    while pool.pending():
        got += pool.fetch(max_events=5)
    pool.close()
    key = lambda e: (e["stream"], e["segment"])             # the order inside one push is the order the warps finished
    assert sorted(got, key=key) == sorted(full, key=key) and len(full) >= S - 2


def test_stream_group_shards_streams_over_handles(handle, ora):
    """sr_stream_group: streams sharded over several handles (all visible GPUs, or two handles on one GPU), lock-step and
    ragged pushes; events carry global stream numbers and equal the batch results"""
    import torch
    ng = max(1, torch.cuda.device_count())
    devs = list(range(ng)) if ng > 1 else [0, 0]
    S, L, T = 37, 40000, 8
    pcm = sr_b200.synth_pcm_host(S, L, 0x5EED8000, 3)
    bank = GOLD["synth/bank"]
    hs = [sr_b200.Handle(d) for d in devs]
    for h in hs:
        h.set_bank(bank, T, 4096)
    handle.set_bank(bank, T, 4096)
    pool = sr_b200.StreamPool(hs, S, L, 2400)
    events = []
    for n0 in range(0, 20000, 800):
        events += pool.push(np.ascontiguousarray(pcm[:, n0:n0 + 800]))
    rng = np.random.default_rng(5)
    pos = np.full(S, 20000, np.int64)
    while (pos < L).any():
        lens = np.minimum(rng.integers(0, 1500, S), L - pos)
        w = max(int(lens.max()), 1)
        chunk = np.zeros((S, w), np.uint16)
        for s in range(S):
            chunk[s, :lens[s]] = pcm[s, pos[s]:pos[s] + lens[s]]
        events += pool.push_ragged(chunk, lens)
        pos += lens
    seg, atap = pool.segments()
    pool.close()
    _check_stream_events(handle, ora, pcm, bank, T, events, seg, atap)
    for h in hs:
        h.close()


def test_geom_b_extension_vs_own_oracle():
    """GEOM_B (200/80/256, BASELINE configs[0]'s framing): PARITY UNPINNED -- the reference has no 256-point path, the
    checker is this repo's own restatement (oracle/sr_oracle.c::sro_mfcc_geom_b, whose FFT generalisation is pinned at
    N = 1024). get_mfcc alone on synthetic / full-range / ragged segments, then the whole recognise path"""
    po = ob.port()
    h = sr_b200.Handle(0)
    h.set_geometry(1)
    B, U = 96, 8000
    pcm = sr_b200.synth_pcm_host(B, U, 0xB0B0)
    rng = np.random.default_rng(0xB)
    pcm[80:] = rng.integers(0, 65536, (16, U)).astype(np.uint16)          # full-range samples: s16 / u32 wraps
    atap = np.zeros(B, sr_b200.ATAP_DTYPE)
    atap["mid_val"] = 2048
    atap["mid_val"][80:] = rng.integers(0, 65536, 16)
    seg = np.zeros((B, 2), np.uint32)
    seg[:, 0] = 80 * rng.integers(1, 30, B)
    seg[:, 1] = np.minimum(seg[:, 0] + 80 * rng.integers(1, 100, B), U)
    seg[1] = (80, 80 + 199)                                                # shorter than one 200-sample frame
    seg[2] = (80, 280)                                                     # exactly one frame
    seg[3] = (80, 80 + 200 + 80 * 119)                                     # 120 frames: rejected (vv_frm_max)
    seg[4] = (0xFFFFFFFF, 0xFFFFFFFF)
    seg[5] = (160, 8000)
    got = h.mfcc(pcm, seg, atap)
    want = po.mfcc_geom_b_batch(pcm, seg, atap)
    assert ob.ftr_equal(got, want)
    assert int(got["frm_num"][2]) == 1 and int(got["frm_num"][1]) == 0 and int(got["frm_num"][3]) == 0 and int(got["frm_num"][5]) == 96
    assert (got["frm_num"] > 0).sum() > 60
    # the same frames in the reference geometry are different numbers (this is a different front end, not a re-labelling)
    h.set_geometry(0)
    ref_geom = h.mfcc(pcm, seg, atap)
    assert not ob.ftr_equal(ref_geom, got)
    # whole path: enrol + recognise in GEOM_B == VAD (reference framing) -> GEOM_B features -> dtw -> argmin on the CPU
    h.set_geometry(1)
    T = 6
    tpl = sr_b200.synth_pcm_host(T, U, 0x7E3A0000)
    bank, est = h.enrol(tpl, 2400)
    assert (est == 0).all()
    h.set_bank(bank, T, 4096)
    utt = sr_b200.synth_pcm_host(32, U, 0x5EED0000)
    out = h.recognise(utt, 2400)
    a = np.zeros(32, sr_b200.ATAP_DTYPE)
    for b in range(32):
        a[b] = po.noise_atap(utt[b], 2400)[0]
    sg = np.stack([po.vad(utt[b], U, a[b:b + 1]) for b in range(32)]).reshape(32, 3, 2)
    assert np.array_equal(out["seg_off"], sg)
    f = po.mfcc_geom_b_batch(utt, sg[:, 0, :], a)
    ok = sg[:, 0, 1] != ob.NULL
    assert ob.ftr_equal(out["ftr"][ok], f[ok])
    sc, _ = po.dtw_batch(f, bank, T, 4096, check_sign=1)
    assert np.array_equal(out["score"][ok], sc[ok]) and ok.sum() >= 30
    h.close()


def test_recognise_multi_handle_sharding(ora):
    """sr_recognise_batch_multi: contiguous shards over several handles (all visible GPUs, or two handles on one GPU)
    == the single-handle result, bit for bit"""
    import torch
    ng = max(1, torch.cuda.device_count())
    devs = list(range(ng)) if ng > 1 else [0, 0]
    B, U, T = 1003, 8000, 9
    pcm = sr_b200.synth_pcm_host(B, U, 0x3131)
    hs = [sr_b200.Handle(d) for d in devs]
    bank, _ = hs[0].enrol(sr_b200.synth_pcm_host(T, U, 0x7E3A0000), 2400)
    for h in hs:
        h.set_bank(bank, T, 4096)
    multi = sr_b200.recognise_multi(hs, pcm, 2400)
    single = hs[0].recognise(pcm, 2400)
    for k in multi:
        assert np.array_equal(multi[k], single[k]), k
    ref = ora.recognise_batch(pcm[:64], 2400, bank, T, 4096)
    assert np.array_equal(multi["score"][:64], ref["score"])
    for h in hs:
        h.close()


def test_thin_c_host_links_reference_named_symbols():
    """host/spch_host.c = save_mdl + spch_recg transcribed against the reference's headers (include/compat) and linked
    straight against libspeech_b200.so: single-call path and batched path must agree"""
    import subprocess
    exe = os.path.join(os.path.dirname(HERE), "stm32-speech-recognition_b200", "host", "spch_host")
    assert os.path.exists(exe), "host binary not built (see __graft_entry__.build)"
    r = subprocess.run([exe, "20"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout[-2000:]


def test_unpack12_device_expander(handle):
    rng = np.random.default_rng(12)
    for n in [2, 14, 16, 18, 4098, 1000002]:
        x = rng.integers(0, 4096, n).astype(np.uint16)
        packed, orbits = sr_b200.pack12_host(x)
        assert (orbits & 0xF000) == 0
        assert np.array_equal(handle.unpack12(packed, n), x), n


def test_packed_transport_equals_plain(handle, ora):
    """sr_recognise_batch with the 12-bit packed PCIe transport: same results as the plain transport, chunks holding a sample
    >= 4096 travel plain, and both agree with the oracle on a sample"""
    B, U, T = 9000, 8000, 7                             # 5 chunks of 2096 utterances (32 MB)
    pcm = sr_b200.synth_pcm_host(B, U, 0x7A000000)
    pcm[2500, 17] = 4096                                # chunks 1 and 4 cannot be packed
    pcm[8999, 7999] = 65535
    handle.set_bank(np.zeros((1, 4096), np.uint8), 0, 4096)
    e = handle.recognise(sr_b200.synth_pcm_host(T, U, 0x7E3A0000), 2400, want=("ftr", "status"))
    bank = sr_b200.make_bank(e["ftr"])
    handle.set_bank(bank, T, 4096)
    try:
        handle.set_transport(0)
        plain = handle.recognise(pcm, 2400)
        assert handle.transport_stats()[0] == 0
        handle.set_transport(1)
        packed = handle.recognise(pcm, 2400)
        n_packed, n_plain, nbytes = handle.transport_stats()
    finally:
        handle.set_transport(-1)
    assert n_packed + n_plain == 5 and n_plain >= 2, (n_packed, n_plain)
    cpus = len(os.sched_getaffinity(0))
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        cpus = cpus if q == "max" else min(cpus, -(-int(q) // int(per)))
    except (OSError, ValueError):
        pass
    if cpus >= 8:                                       # with fewer CPUs the library creates no packer pool: everything plain
        assert n_packed >= 1 and nbytes < pcm.nbytes, (n_packed, n_plain, nbytes)
    for k in plain:
        assert plain[k].tobytes() == packed[k].tobytes(), k
    sel = np.array([0, 2095, 2096, 2500, 4191, 4192, 8383, 8384, 8999])
    ref = ora.recognise_batch(pcm[sel], 2400, bank, T, 4096)
    _cmp_recog({k: v[sel] for k, v in packed.items()}, ref)


def test_packed_transport_under_torchrun_two_ranks():
    """the 12-bit transport with two ranks of one node sharing the CPU quota (LOCAL_WORLD_SIZE = 2: each rank sizes its
    packer pool from its share): packed == plain == device path on every rank"""
    import subprocess
    import sys
    import torch
    script = os.path.join(HERE, "_torchrun_pack.py")
    port = 29500 + (os.getpid() % 500)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), script], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("rank ok") == 2, r.stdout[-3000:]


def test_multi_gpu_c_host_nccl_allgather():
    """host/spch_host_mgpu.c: plain C + pthreads, one rank per GPU, NCCL all-gather through the C-ABI
    (sr_recognise_batch_dev_allgather); every rank's gathered scores and argmin keys == the single-GPU batch result"""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (NCCL refuses two ranks on one device)")
    exe = os.path.join(os.path.dirname(HERE), "stm32-speech-recognition_b200", "host", "spch_host_mgpu")
    assert os.path.exists(exe), "host binary not built (see __graft_entry__.build)"
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([exe, str(n), "384"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and ": 0 mismatches" in r.stdout, r.stdout[-2000:]


def test_nccl_allgather_through_python_binding_two_ranks():
    """sr_comm_* from two processes (torchrun): the id travels through torch.distributed, the gather runs inside
    libspeech_b200.so; gathered block r on every rank == rank r's own results"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = os.path.join(HERE, "_torchrun_comm.py")
    port = 29500 + ((os.getpid() + 7) % 500)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), script], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("rank ok") == 2, r.stdout[-3000:]


def test_numa_helpers_and_labels(handle):
    """placement helpers never fail on a single-node box and report consistently; commstr labels (main.c:25-31, 295)"""
    L = sr_b200.lib()
    node = L.sr_device_numa_node(0)
    arr, p = sr_b200.host_alloc_dev(0, 1 << 20)
    arr[:] = 7
    got = L.sr_host_numa_node(C.c_void_p(p))
    assert node < 0 or got < 0 or got == node
    sr_b200.host_free(p)
    assert handle.label(0) == b"0 " and handle.label(9) == b"9 " and handle.label(10) == bytes([0xC9, 0xCF]) and handle.label(18) is None
    h2 = sr_b200.Handle(0)
    h2.set_labels([b"on", b"off", b"up"], 4)
    assert h2.label(1) == b"off" and h2.label(3) is None
    h2.close()


def test_empty_batch_and_argument_errors(handle):
    z = np.zeros((0, 8000), np.uint16)
    assert handle.recognise(z, 2400)["cmd"].shape == (0,)
    with pytest.raises(sr_b200.SrError):
        handle.set_bank(np.zeros((2, 100), np.uint8), 2, 100)          # slot stride < sizeof(v_ftr_tag)
