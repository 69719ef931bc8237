"""N > 1 host logic on CPU: contiguous sequence sharding and the gather of keypoint records over
torch.distributed (gloo, world_size 2). The record packing is the same code the NCCL path runs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from object_keypoints_b200 import sharding


def test_shard_range_is_a_contiguous_balanced_partition():
    for n in (0, 1, 7, 64, 4096, 4099):
        for world in (1, 2, 3, 8):
            ranges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_shard_sequences_keeps_sequences_whole():
    lengths = [64] * 64                      # config 4: 64 sequences x 64 frames
    for world in (1, 2, 4, 8):
        covered = []
        for r in range(world):
            s0, s1, f0, f1 = sharding.shard_sequences(lengths, r, world)
            assert f1 - f0 == sum(lengths[s0:s1])
            covered.append((f0, f1))
        assert covered[0][0] == 0 and covered[-1][1] == 4096
    s0, s1, f0, f1 = sharding.shard_sequences([900, 30, 450], 1, 2)
    assert (s0, s1, f0, f1) == (2, 3, 930, 1380)


def fake_tables(n, seed):
    g = torch.Generator().manual_seed(seed)
    return {'n_objects': torch.randint(0, 5, (n,), generator=g, dtype=torch.int32),
            'flags': torch.randint(0, 64, (n,), generator=g, dtype=torch.int32),
            'kp_count': torch.randint(0, 3, (n, 4, 3), generator=g, dtype=torch.int32),
            'kp_point': torch.randn((n, 4, 3, 3, 3), generator=g, dtype=torch.float64)}


def test_record_roundtrip():
    t = fake_tables(6, 0)
    back = sharding.unpack_records(sharding.record_tensor(t), t)
    for key in t:
        assert torch.equal(back[key], t[key]), key


def _worker(rank, world, port, frames_total, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        f0, f1 = sharding.shard_range(frames_total, rank, world)
        everything = fake_tables(frames_total, 123)                      # the same "global" result on each rank
        mine = {k: v[f0:f1].clone() for k, v in everything.items()}     # this rank decoded its shard
        gathered = sharding.gather_keypoint_records(mine, world)
        got = sharding.unpack_records(gathered, mine)
        ok = all(torch.equal(got[k], everything[k]) for k in everything)
        np.save(os.path.join(out_dir, f'ok_{rank}.npy'), np.array([ok, gathered.shape[0]]))
    finally:
        dist.destroy_process_group()


def test_gather_over_gloo_world_size_2(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    world, frames_total = 2, 16
    mp.spawn(_worker, args=(world, port, frames_total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok, rows = np.load(os.path.join(str(tmp_path), f'ok_{r}.npy'))
        assert ok == 1 and rows == frames_total
