"""The drop-in boundary without a GPU: libspeech_b200.so loads, exports every symbol that include/*.h
declares, keeps the reference's struct layouts, and FAILS LOUDLY (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sr_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.split("\n") if not l.strip().startswith("#"))
    text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    names = []
    for stmt in text.split(";"):
        stmt = stmt.replace('extern "C" {', "").strip()
        if "(" in stmt and not stmt.startswith("typedef"):
            m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*\(", stmt)
            if m:
                names.append(m.group(1))
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    L = sr_b200.lib()
    names = _declared("speech_recog.h") + _declared("sr_synth.h")
    assert {"noise_atap", "VAD", "get_mfcc", "dtw", "fft", "get_dis", "sr_recognise_batch", "sr_mfcc_batch_dev"} <= set(names)
    for n in names:
        assert hasattr(L, n), n


def test_struct_layouts_match_reference_headers():
    assert sr_b200.ATAP_DTYPE.itemsize == 12               # VAD.H:10-16
    assert sr_b200.FTR_DTYPE.itemsize == 2860              # MFCC.H:18-25 (#pragma pack(1))
    assert sr_b200.FTR_DTYPE.fields["mfcc_dat"][1] == 4
    assert C.sizeof(sr_b200.ValidTag) == 2 * C.sizeof(C.c_void_p)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = sr_b200.lib()
    assert L.sr_device_count() == 0
    with pytest.raises(sr_b200.SrError):
        sr_b200.Handle(0)
    assert b"no CUDA device" in L.sr_last_error(None) or len(L.sr_last_error(None)) > 0
    # reference-named entry points return their failure sentinels
    a = np.zeros(12, np.int16)
    assert L.get_dis(a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p)) == 0xFFFFFFFF
    f = np.zeros(2, sr_b200.FTR_DTYPE)
    assert L.dtw(f[0:1].ctypes.data_as(C.c_void_p), f[1:2].ctypes.data_as(C.c_void_p)) == 0xFFFFFFFF
    assert not L.fft(a.ctypes.data_as(C.c_void_p), 12)


def test_host_synth_is_deterministic_and_in_range():
    a = sr_b200.synth_pcm_host(5, 8000, 123)
    b = sr_b200.synth_pcm_host(5, 8000, 123)
    c = sr_b200.synth_pcm_host(5, 8000, 124)
    assert np.array_equal(a, b) and np.array_equal(a[1:], c[:4]) and a.max() <= 4095
    # calibration window is noise only: |x - mid| <= 60
    mid = np.median(a[:, :2400], axis=1)
    assert (np.abs(a[:, :2400].astype(np.int64) - mid[:, None]) <= 61).all()


def test_wav_ingestion_matches_the_documented_mapping(tmp_path):
    """8/16-bit PCM WAV -> 12-bit unsigned ADC codes (x/16 + 2048), checked against python's wave module; and on the
    reference's own recordings when the tree is mounted (the oracle then finds speech in them)"""
    import io
    import wave
    rng = np.random.default_rng(0)
    for width, nch in ((2, 1), (1, 1), (2, 2)):
        n = 3000
        x = rng.integers(-32768, 32768, (n, nch)).astype(np.int16) if width == 2 else rng.integers(0, 256, (n, nch)).astype(np.uint8)
        bio = io.BytesIO()
        with wave.open(bio, "wb") as w:
            w.setnchannels(nch); w.setsampwidth(width); w.setframerate(8000); w.writeframes(x.tobytes())
        got, rate = sr_b200.wav_to_adc12(bio.getvalue())
        c0 = x[:, 0].astype(np.int64)
        want = np.trunc(c0 / 16).astype(np.int64) + 2048 if width == 2 else (c0 - 128) * 16 + 2048
        assert rate == 8000 and np.array_equal(got.astype(np.int64), np.clip(want, 0, 4095))
    with pytest.raises(sr_b200.SrError):
        sr_b200.wav_to_adc12(b"RIFFxxxxWAVEjunk")
    ref_wav = "/root/reference/Matlab/\u8bed\u97f3\u6837\u672c/\u4e0a\u4e0b\u524d\u540e\u5de6\u53f3.wav"
    if os.path.exists(ref_wav):
        import oracle_bind as ob
        pcm, rate = sr_b200.wav_to_adc12(open(ref_wav, "rb").read())
        assert rate == 8000 and len(pcm) > 8000
        o = ob.best_oracle()
        atap = o.noise_atap(pcm[:16000].copy(), 2400)
        seg = o.vad(pcm[:16000].copy(), min(len(pcm), 16000), atap)
        assert seg[0] != ob.NULL                       # the reference's VAD opens a segment on its own recording


def test_command_labels_without_a_device():
    """commstr[] of main.c:25-31 (what spch_recg returns, main.c:295): the reference's own 18 labels are built in and
    need neither a handle nor a GPU; sr_labels_batch maps failed utterances to NULL like spch_recg's early returns"""
    import ctypes as C
    import numpy as np
    import sr_b200
    L = sr_b200.lib()
    L.sr_label.argtypes = [C.c_void_p, C.c_uint32]
    L.sr_label.restype = C.c_void_p
    got = [C.string_at(L.sr_label(None, k)) for k in range(18)]
    assert got[:10] == [b"%d " % k for k in range(10)]
    assert [g.decode("gbk") for g in got[10:]] == ["上", "下", "前", "后", "左", "右", "大", "小"]
    assert L.sr_label(None, 18) is None and L.sr_label(None, 2 ** 31) is None
    L.sr_labels_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    cmd = np.array([3, 17, 5, 40], np.uint32)
    st = np.array([0, 0, 1, 0], np.uint8)
    out = (C.c_void_p * 4)()
    assert L.sr_labels_batch(None, cmd.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), 4, out) == 0
    assert C.string_at(out[0]) == b"3 " and C.string_at(out[1]).decode("gbk") == "小" and out[2] is None and out[3] is None


def test_placement_helpers_degrade_without_a_device():
    """no CUDA device: the NUMA helpers report 'unknown' instead of failing, and never touch the calling thread"""
    import ctypes as C
    import os
    import numpy as np
    import sr_b200
    L = sr_b200.lib()
    if L.sr_device_count() != 0:
        import pytest
        pytest.skip("a CUDA device is present")
    before = os.sched_getaffinity(0)
    assert L.sr_device_numa_node(0) == -1
    assert L.sr_bind_thread_to_device(0) == -1
    assert os.sched_getaffinity(0) == before
    a = np.ones(4096, np.uint8)
    assert L.sr_host_numa_node(a.ctypes.data_as(C.c_void_p)) >= -1
