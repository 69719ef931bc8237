"""Multi-GPU check of sharding.RecordExchange (run under torchrun on N GPUs of one box): every transport must
deliver exactly the records the plain torch packing + all_gather delivers, through several pipelined steps.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_exchange.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from object_keypoints_b200 import KeypointDecoder, synthetic, sharding

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
device = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=device)
cfg, size, frames = [1, 3], (64, 64), 257
decoder = KeypointDecoder(cfg, size, camera=synthetic.default_camera(size), device=device)
report = {}
for transport, root, staged in (('nccl', None, True), ('peer', None, True), ('nccl', 0, True), ('peer', 0, True), ('peer', world - 1, True),
                                ('peer', None, False), ('peer', 0, False)):
    label = f"{transport}{'' if staged or transport != 'peer' else '-direct'}/{'allgather' if root is None else 'gather->' + str(root)}"
    try:
        exchange = sharding.RecordExchange(decoder, frames, world=world, rank=rank, transport=transport, root=root, staged=staged)
        tables = decoder.tables(frames)
        for step in range(7):                                     # more steps than ring slots
            batch = synthetic.make_batch(frames, cfg, size, seed=100 * step + rank, objects=(1, 3))
            heat = batch.heat.copy()
            heat[rank % frames, 1] = 0.5                          # an overflowing map: that frame's record comes from the fix-up launch
            sink = exchange.begin()
            decoder.decode_batch(heat, batch.depth, batch.centers, tables=tables, records=sink)
            got, done = exchange.end()
            # what every rank should have sent: the torch packing of its own tables, gathered with a plain all_gather
            mine = sharding.unpack_compact_records(sharding.pack_compact_records(tables, cfg), 16, cfg)
            want = {}
            for key, value in mine.items():
                parts = [torch.empty_like(value) for _ in range(world)]
                dist.all_gather(parts, value.contiguous())
                want[key] = torch.cat(parts)
            done.synchronize()
            if root is None or rank == root:
                back = sharding.unpack_compact_records(got, 16, cfg)
                for key in want:
                    assert torch.equal(back[key], want[key]), f"{label}: step {step} {key} differs on rank {rank}"
                assert int(back['n_objects'].sum()) > 0
            dist.barrier()
        report[label] = 'ok'
    except Exception as error:
        report[label] = f"FAILED {type(error).__name__}: {error}"
    dist.barrier()
print(f"rank {rank}/{world}: {report}", flush=True)
dist.destroy_process_group()
