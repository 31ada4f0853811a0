#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz. Needs /root/reference (the board captures) and oracle/_ref/libref.so
(the reference's own C compiled for the host): every expected value below is produced by EXECUTING THE
REFERENCE, not by our restatement.

  captures.npz : the reference's real ADC captures converted text -> u16
                 (Matlab/语音样本/STM32 {123,456,noise}.txt, Matlab/matlab仿真/{STM32_Voice - 123.txt, v1.c})
  golden.npz   : for every capture: noise_atap, VAD offsets, get_mfcc of every closed segment, dtw between
                 segments; for a seeded synthetic batch (inputs regenerated from the seed by the tests):
                 the full spch_recg outputs against an 8-slot bank.
Run:  python tests/golden/make_golden.py
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "stm32-speech-recognition_b200", "python"))
import oracle_bind as ob  # noqa: E402
import sr_b200  # noqa: E402

REF = "/root/reference/Matlab"
FILES = {
    "stm32_123": REF + "/语音样本/STM32 123.txt",
    "stm32_456": REF + "/语音样本/STM32 456.txt",
    "stm32_noise": REF + "/语音样本/STM32 noise.txt",
    "stm32_voice_123": REF + "/matlab仿真/STM32_Voice - 123.txt",
    "v1": REF + "/matlab仿真/v1.c",
}


def load_txt(path):
    txt = open(path, "rb").read().decode("latin1")
    return np.array([int(x) for x in re.findall(r"-?\d+", txt)], dtype=np.int64)


def checksum(v):
    cs = 0
    for x in v.tolist():
        cs = (cs * 31 + x) & 0xFFFFFFFFFFFFFFFF
    return cs - (1 << 64) if cs >= (1 << 63) else cs


def main():
    r = ob.ref()
    caps, gold = {}, {}
    for name, path in FILES.items():
        v = load_txt(path)
        assert v.min() >= 0 and v.max() <= 4095, (name, v.min(), v.max())
        pcm = v.astype(np.uint16)
        caps[name] = pcm
        n = len(pcm)
        atap = r.noise_atap(pcm, 2400)
        seg = r.vad(pcm, n, atap)
        gold[name + "/atap"] = atap
        gold[name + "/seg"] = seg
        ftrs = []
        for k in range(3):
            if seg[2 * k + 1] != ob.NULL:
                f = r.mfcc_batch(pcm.reshape(1, -1), seg[2 * k:2 * k + 2].reshape(1, 2), atap)
                ftrs.append(f)
                gold["%s/ftr%d" % (name, k)] = f
        if len(ftrs) >= 2:
            bank = np.concatenate(ftrs).view(np.uint8).reshape(len(ftrs), -1)
            sc, _ = r.dtw_batch(np.concatenate(ftrs), bank, len(ftrs), 2860)
            gold[name + "/dtw"] = sc
        print(name, n, atap, seg.tolist(), [int(f["frm_num"][0]) for f in ftrs])
    # KATs of SURVEY.md section 4 -- a mismatch means the FFT restatement differs from the survey's
    a = gold["stm32_123/atap"][0]
    assert (a["mid_val"], a["n_thl"], a["z_thl"], a["s_thl"]) == (2213, 172, 2, 9524)
    assert gold["stm32_123/seg"].tolist()[:4] == [3920, 6880, 8640, 11440]
    f0 = gold["stm32_123/ftr0"]
    assert int(f0["frm_num"][0]) == 36 and checksum(f0["mfcc_dat"][0][:432]) == -6592886050377706402
    f1 = gold["stm32_123/ftr1"]
    assert int(f1["frm_num"][0]) == 34 and checksum(f1["mfcc_dat"][0][:408]) == -6112762061601859578
    assert gold["stm32_123/dtw"].tolist() == [[0, 3874], [3874, 0]]

    # seeded synthetic batch through the reference's spch_recg flow
    B, U, T = 24, 8000, 8
    pcm = sr_b200.synth_pcm_host(B, U, 0x5EED0000)
    tpl = sr_b200.synth_pcm_host(T, U, 0x7E3A0000)
    e = r.recognise_batch(tpl, 2400, None, 0, 4096)
    assert (e["status"] == 0).all()
    bank = sr_b200.make_bank(e["ftr"])
    out = r.recognise_batch(pcm, 2400, bank, T, 4096)
    gold["synth/pcm_sum"] = np.array([int(pcm.astype(np.uint64).sum()), int(tpl.astype(np.uint64).sum())], np.uint64)
    gold["synth/bank"] = bank
    for k, v in out.items():
        gold["synth/" + k] = v
    print("synth status", out["status"].tolist(), "frames", out["ftr"]["frm_num"].tolist(), "cmd", out["cmd"].tolist())
    # 2 s / 3-word stream shape (config 5 building block): 40000 samples
    pcm5 = sr_b200.synth_pcm_host(4, 40000, 0x5EED5000, 3)
    out5 = r.recognise_batch(pcm5, 2400, bank, T, 4096)
    for k, v in out5.items():
        gold["synth5/" + k] = v
    print("synth5 seg", out5["seg_off"].reshape(4, 6).tolist())
    np.savez_compressed(os.path.join(HERE, "captures.npz"), **caps)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **gold)
    print("wrote captures.npz, golden.npz")


if __name__ == "__main__":
    main()
