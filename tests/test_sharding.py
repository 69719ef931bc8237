"""N > 1 host logic on CPU: contiguous sequence sharding and the gather of keypoint records over
torch.distributed (gloo, world_size 2). The record packing is the same code the NCCL path runs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pytest

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


def consistent_tables(n, seed, cfg=(1, 3), O=4):
    """Fake decode tables whose counts respect keypoint_config (what the kernel can actually produce)."""
    g = torch.Generator().manual_seed(seed)
    config = [1] + list(cfg)
    C, S = len(config), max(config)
    limit = torch.tensor(config, dtype=torch.int32)[None, None, :] + 1
    count = (torch.randint(0, 9, (n, O, C), generator=g, dtype=torch.int32) % limit).to(torch.int32)
    return {'n_objects': torch.randint(0, O + 1, (n,), generator=g, dtype=torch.int32),
            'flags': torch.randint(0, 256, (n,), generator=g, dtype=torch.int32),
            'kp_count': count, 'kp_point': torch.randn((n, O, C, S, 3), generator=g, dtype=torch.float64)}


def _masked(t):
    """What a compact record preserves: counts of existing objects, points of kept slots."""
    O, S = t['kp_count'].shape[1], t['kp_point'].shape[3]
    valid = torch.arange(O)[None, :] < t['n_objects'][:, None]
    count = torch.where(valid[:, :, None], t['kp_count'], torch.zeros((), dtype=torch.int32))
    slot = torch.arange(S)[None, None, None, :] < count[..., None]
    return {'n_objects': t['n_objects'], 'flags': t['flags'], 'kp_count': count,
            'kp_point': torch.where(slot[..., None], t['kp_point'], torch.zeros((), dtype=torch.float64))}


def test_compact_record_layout_and_roundtrip():
    lay = sharding.compact_layout(16, {'keypoint_config': [1, 3]})
    assert (lay['C'], lay['P'], lay['slot_of']) == (3, 5, [0, 1, 2])
    assert lay['points_offset'] == 8 + 192 and lay['record_bytes'] == 8 + 192 + 16 * 5 * 24      # 2120, not 3856
    assert sharding.compact_layout(3, [1, 1, 1])['points_offset'] == 8 + 48                     # 4 * 3 * 4 = 48 is a multiple of 8
    assert sharding.compact_layout(3, [2])['points_offset'] == 8 + 24
    assert sharding.compact_layout(1, [])['points_offset'] == 16                                 # 4 bytes of counts, padded to 8
    for cfg, O in (([1, 3], 4), ([1, 1, 1], 3), ([], 2), ([2, 4], 5)):
        t = consistent_tables(9, 5, cfg, O)
        records = sharding.pack_compact_records(t, cfg)
        assert records.dtype == torch.uint8 and records.shape == (9, sharding.compact_layout(O, cfg)['record_bytes'])
        back = sharding.unpack_compact_records(records, O, cfg)
        want = _masked(t)
        for key in want:
            assert torch.equal(back[key], want[key]), (cfg, key)
        # stale bytes where the kernel does not write must not leak through the reader
        lay = sharding.compact_layout(O, cfg)
        noisy = records.clone()
        for n in range(9):
            for o in range(O):
                for c in range(lay['C']):
                    kept = int(want['kp_count'][n, o, c])
                    if o >= int(t['n_objects'][n]):
                        noisy[n, 8 + 4 * (o * lay['C'] + c):8 + 4 * (o * lay['C'] + c) + 4] = 0x5A
                    for s_ in range(kept, lay['cfg'][c]):
                        at = lay['points_offset'] + 24 * (o * lay['P'] + lay['slot_of'][c] + s_)
                        noisy[n, at:at + 24] = 0x5A
        assert not torch.equal(noisy, records)
        back = sharding.unpack_compact_records(noisy, O, cfg)
        for key in want:
            assert torch.equal(back[key], want[key]), (cfg, key, 'stale bytes leaked')


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
        # the compact records of the fused kernel travel the same way (NCCL transport: all_gather of the byte rows)
        everything = consistent_tables(frames_total, 321)
        mine = {k: v[f0:f1].clone() for k, v in everything.items()}
        local = sharding.pack_compact_records(mine, [1, 3])
        rows = torch.empty((world * local.shape[0], local.shape[1]), dtype=torch.uint8)
        dist.all_gather_into_tensor(rows, local)
        back = sharding.unpack_compact_records(rows, 4, [1, 3])
        want = _masked(everything)
        ok = ok and all(torch.equal(back[k], want[k]) for k in want)
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


@pytest.mark.gpu
def test_record_exchange_on_the_gpus_of_this_box():
    """The multi-GPU exchange on hardware: every transport (NCCL all_gather / gather, peer stores to every
    rank / to one root) through seven pipelined steps against the torch packing + plain all_gather of the same tables,
    including a frame whose record comes from the overflow fix-up. Runs tools/check_exchange.py under torchrun on all GPUs
    of the box; skipped on a single-GPU box (the N = 2 and N = 8 logs of the round are in profiles/)."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the exchange needs at least two ranks")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    run = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(root, 'tools', 'check_exchange.py')],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    lines = [line for line in run.stdout.splitlines() if line.startswith('rank ')]
    assert len(lines) == n and not any('FAILED' in line for line in lines), run.stdout[-2000:]
