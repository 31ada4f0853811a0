"""helper of test_packed_transport_under_torchrun_two_ranks: run by torchrun with 2 ranks (no collective needed)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "stm32-speech-recognition_b200", "python"))
import sr_b200  # noqa: E402

rank = int(os.environ.get("LOCAL_RANK", "0"))
import torch  # noqa: E402

dev = rank % max(1, torch.cuda.device_count())
sr_b200.lib().sr_bind_thread_to_device(dev)
B, U, T = 2048 * 5, 8000, 6                       # 5 chunks of 32 MB: the packed path engages (>= 4 chunks)
arr, p = sr_b200.host_alloc_dev(dev, B * U * 2)
pcm = arr.view(np.uint16).reshape(B, U)
pcm[:] = sr_b200.synth_pcm_host(B, U, 0x5EED0000 + rank * B)
pcm[2048 * 2 + 5, 100] = 60000                    # one chunk holds a sample >= 4096: it must travel plain
h = sr_b200.Handle(dev)
bank, _ = h.enrol(sr_b200.synth_pcm_host(T, U, 0x7E3A0000), 2400)
h.set_bank(bank, T, 4096)
want = ("best_idx", "best_dis", "cmd", "status", "seg_off")
h.set_transport(0)
plain = h.recognise(pcm, 2400, want=want)
h.set_transport(1)
packed = h.recognise(pcm, 2400, want=want)
pk, pl, nbytes = h.transport_stats()
h.set_transport(-1)
auto = h.recognise(pcm, 2400, want=want)
for k in want:
    assert np.array_equal(plain[k], packed[k]) and np.array_equal(plain[k], auto[k]), k
assert pk + pl == 5 and pl >= 1, (pk, pl)
print("rank %d: packed %d plain %d chunks, %d bytes" % (rank, pk, pl, nbytes))
h.close()
sr_b200.host_free(p)
print("rank ok")
