"""helper of test_nccl_allgather_through_python_binding_two_ranks: run by torchrun with 2 ranks, one GPU each"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "stm32-speech-recognition_b200", "python"))
import sr_b200  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
h = sr_b200.Handle(local)
idt = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(sr_b200.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
h.comm_create(rank, world, bytes(idt.cpu().numpy().tobytes()))
B, U, T = 1000, 8000, 7
bank, _ = h.enrol(sr_b200.synth_pcm_host(T, U, 0x7E3A0000), 2400)
h.set_bank(bank, T, 4096)
pcm_all = sr_b200.synth_pcm_host(B * world, U, 0x5EED0000)
pcm_all[::11] = 2048                                       # some VAD failures
stream = torch.cuda.Stream(dev)
h.set_stream(stream.cuda_stream)
pcm = torch.from_numpy(pcm_all[rank * B:(rank + 1) * B].view(np.int16)).to(dev)
score = torch.zeros((B, T), dtype=torch.int32, device=dev)
gs = torch.zeros((world * B, T), dtype=torch.int32, device=dev)
gb = torch.zeros(world * B, dtype=torch.int64, device=dev)
for _ in range(3):                                         # repeated: each call orders itself after the previous gather
    h.recognise_dev_allgather(pcm.data_ptr(), U, B, 2400, gathered_score=gs.data_ptr(), gathered_best=gb.data_ptr(),
                              score=score.data_ptr())
h.sync()
single = sr_b200.Handle(local)
single.set_bank(bank, T, 4096)
ref = single.recognise(pcm_all, 2400, want=("score", "best_idx", "best_dis", "status"))
ok = ref["status"] == 0
got_s = gs.cpu().numpy().view(np.uint32)
got_b = gb.cpu().numpy().view(np.uint64)
assert np.array_equal(got_s[ok], ref["score"][ok])
want_b = (ref["best_dis"].astype(np.uint64) << np.uint64(32)) | ref["best_idx"].astype(np.uint64)
assert np.array_equal(got_b, want_b)
single.close()
h.close()
dist.destroy_process_group()
print("rank ok")
