"""Multi-GPU layer: frames / sequences shard independently (one process per GPU), the only
cross-GPU step is the gather of the 3D keypoint records. The reference
has no multi-GPU inference code (it asserts batch 1, perception/pipeline.py:183).

Works with any torch.distributed backend: NCCL over NVLink on the B200 box, gloo in CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced partition: rank r owns [start, stop). Sequences stay whole and in order,
    so concatenating the ranks' results restores the global order."""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sequences(sequence_lengths, rank, world):
    """Sequences (lists of frames) are assigned contiguously; returns (first_sequence, last_sequence,
    first_frame, last_frame) of this rank."""
    s0, s1 = shard_range(len(sequence_lengths), rank, world)
    f0 = int(sum(sequence_lengths[:s0]))
    f1 = f0 + int(sum(sequence_lengths[s0:s1]))
    return s0, s1, f0, f1


RECORD_FIELDS = ('n_objects', 'flags', 'kp_count', 'kp_point')


def record_tensor(tables):
    """Pack what the consumer of the pipeline needs per frame -- object count, flags, kept-keypoint
    counts and the camera-frame 3D points -- into one float64 [N, R] tensor (equal size on every
    rank, so a single all_gather moves it)."""
    t = tables.tensors if hasattr(tables, 'tensors') else tables
    N = t['n_objects'].shape[0]
    parts = [t['n_objects'].reshape(N, 1).double(), t['flags'].reshape(N, 1).double(),
             t['kp_count'].reshape(N, -1).double(), t['kp_point'].reshape(N, -1)]
    return torch.cat(parts, dim=1).contiguous()


def unpack_records(records, like):
    """Inverse of record_tensor for a gathered [M, R] tensor; `like` supplies the per-frame shapes."""
    t = like.tensors if hasattr(like, 'tensors') else like
    M = records.shape[0]
    n_count = t['kp_count'][0].numel()
    out = {'n_objects': records[:, 0].round().to(torch.int32), 'flags': records[:, 1].round().to(torch.int32)}
    out['kp_count'] = records[:, 2:2 + n_count].round().to(torch.int32).reshape((M,) + tuple(t['kp_count'].shape[1:]))
    out['kp_point'] = records[:, 2 + n_count:].reshape((M,) + tuple(t['kp_point'].shape[1:]))
    return out


def gather_keypoint_records(tables, world=None, out=None):
    """all_gather of the per-rank records -> [world * N, R] on every rank (rank order = frame order
    under shard_range). With world == 1 it is the local record."""
    record = record_tensor(tables)
    world = dist.get_world_size() if world is None else world
    if world == 1 or not dist.is_initialized():
        return record
    if out is None:
        out = torch.empty((world * record.shape[0], record.shape[1]), dtype=record.dtype, device=record.device)
    dist.all_gather_into_tensor(out, record)
    return out
