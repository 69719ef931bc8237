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
for transport, root in (('nccl', None), ('peer', None), ('nccl', 0), ('peer', 0), ('peer', world - 1)):
    label = f"{transport}/{'allgather' if root is None else 'gather->' + str(root)}"
    try:
        exchange = None
        for step in range(5):
            batch = synthetic.make_batch(frames, cfg, size, seed=100 * step + rank, objects=(1, 3))
            tables = decoder.decode_batch(batch.heat, batch.depth, batch.centers)
            if exchange is None:
                exchange = sharding.RecordExchange(tables, world=world, rank=rank, transport=transport, root=root)
            got, done = exchange.exchange(tables)
            want = sharding.gather_keypoint_records(tables, world)
            done.synchronize()
            if root is None or rank == root:
                assert torch.equal(got, want), f"{label}: step {step} differs on rank {rank}"
        report[label] = 'ok'
    except Exception as error:
        report[label] = f"FAILED {type(error).__name__}: {error}"
    dist.barrier()
print(f"rank {rank}/{world}: {report}", flush=True)
dist.destroy_process_group()
