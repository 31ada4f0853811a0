"""world_size-2 gloo test of the multi-GPU host logic (sharding + the one all-gather of scores) on CPU.
The compute engine inside each rank is the oracle (test infrastructure) standing in for the GPU kernels --
what is under test is that sharded + gathered results equal the unsharded ones, bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "stm32-speech-recognition_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as ob
    import sr_b200
    from sr_b200.dist import gather_blocks, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, U = 6, 8000
    o = ob.port()
    tpl = sr_b200.synth_pcm_host(T, U, 0x7E3A0000)
    bank = sr_b200.make_bank(o.recognise_batch(tpl, 2400, None, 0, 4096)["ftr"])
    lo, hi = shard_range(B, rank, world)
    pcm = sr_b200.synth_pcm_host(hi - lo, U, 0x5EED0000 + lo)          # rank-local shard, seeded by global utterance id
    loc = o.recognise_batch(pcm, 2400, bank, T, 4096)
    score = gather_blocks(torch.from_numpy(loc["score"].astype(np.int64)), B)
    best = gather_blocks(torch.from_numpy(loc["best_idx"].astype(np.int64)), B)
    if rank == 0:
        np.save(os.path.join(out_dir, "score.npy"), score.numpy())
        np.save(os.path.join(out_dir, "best.npy"), best.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [16, 17])
def test_sharded_allgather_equals_unsharded(tmp_path, B):
    import oracle_bind as ob
    import sr_b200
    port = 29500 + (os.getpid() % 2000) + B
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    o = ob.port()
    T, U = 6, 8000
    tpl = sr_b200.synth_pcm_host(T, U, 0x7E3A0000)
    bank = sr_b200.make_bank(o.recognise_batch(tpl, 2400, None, 0, 4096)["ftr"])
    full = o.recognise_batch(sr_b200.synth_pcm_host(B, U, 0x5EED0000), 2400, bank, T, 4096)
    assert np.array_equal(np.load(tmp_path / "score.npy"), full["score"].astype(np.int64))
    assert np.array_equal(np.load(tmp_path / "best.npy"), full["best_idx"].astype(np.int64))


def test_shard_ranges_partition():
    from sr_b200.dist import shard_range
    for n in (0, 1, 7, 65536, 1048576):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
