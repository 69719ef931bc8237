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


# ------------------------------------------------------------------------------------------------
# compact records: what the fused decode kernel emits per frame (include/okp.h, okp_decode_emit_*)
# ------------------------------------------------------------------------------------------------
def compact_layout(max_objects, keypoint_config):
    """Byte layout of a compact record: int32 n_objects, uint32 flags, int32 kp_count[O][C], (pad to 8),
    float64 point[O][P][3] with P = 1 + sum(keypoint_config) in (map, slot) order.
    -> dict(O, C, P, cfg, points_offset, record_bytes, slot_of) where slot_of[c] is the first point slot of map c."""
    cfg = [1] + [int(v) for v in (keypoint_config['keypoint_config'] if isinstance(keypoint_config, dict) else keypoint_config)]
    O, C, P = int(max_objects), len(cfg), sum(cfg)
    points_offset = 8 + (4 * O * C + 7) // 8 * 8
    slot_of = [sum(cfg[:c]) for c in range(C)]
    return {'O': O, 'C': C, 'P': P, 'cfg': cfg, 'points_offset': points_offset, 'record_bytes': points_offset + 24 * O * P,
            'slot_of': slot_of}


def pack_compact_records(tables, keypoint_config, record_bytes=None):
    """Plain torch statement of the record the kernel emits (tests, verification of the exchange): uint8 [N, record_bytes]
    with every part the kernel does not write left zero."""
    t = tables.tensors if hasattr(tables, 'tensors') else tables
    N, O = t['kp_count'].shape[0], t['kp_count'].shape[1]
    lay = compact_layout(O, keypoint_config)
    stride = lay['record_bytes'] if record_bytes is None else int(record_bytes)
    device = t['kp_count'].device
    out = torch.zeros((N, stride), dtype=torch.uint8, device=device)
    n_objects = t['n_objects'].to(torch.int32)
    header = torch.stack([n_objects, t['flags'].to(torch.int32)], dim=1).contiguous()
    out[:, :8] = header.view(torch.uint8).reshape(N, 8)
    valid_obj = torch.arange(O, device=device)[None, :] < n_objects[:, None]
    counts = torch.where(valid_obj[:, :, None], t['kp_count'].to(torch.int32), torch.zeros((), dtype=torch.int32, device=device))
    out[:, 8:8 + 4 * O * lay['C']] = counts.contiguous().view(torch.uint8).reshape(N, -1)
    points = torch.zeros((N, O, lay['P'], 3), dtype=torch.float64, device=device)
    for c in range(lay['C']):
        for s_ in range(lay['cfg'][c]):
            keep = valid_obj & (counts[:, :, c] > s_)
            points[:, :, lay['slot_of'][c] + s_] = torch.where(keep[:, :, None], t['kp_point'][:, :, c, s_], points[:, :, lay['slot_of'][c] + s_])
    out[:, lay['points_offset']:lay['record_bytes']] = points.view(torch.uint8).reshape(N, -1)
    return out


def unpack_compact_records(records, max_objects, keypoint_config):
    """uint8 [M, record_bytes] gathered records -> dict(n_objects [M], flags [M], kp_count [M,O,C], kp_point [M,O,C,S,3])
    in the layout of the decode tables. Parts the kernel did not write (objects >= n_objects, slots >= kp_count) hold
    whatever the buffer held before: they are reset to zero here, readers never see them."""
    lay = compact_layout(max_objects, keypoint_config)
    O, C, P = lay['O'], lay['C'], lay['P']
    M = records.shape[0]
    S = max(lay['cfg'])
    device = records.device
    header = records[:, :8].contiguous().view(torch.int32).reshape(M, 2)
    n_objects, flags = header[:, 0].clone(), header[:, 1].clone()
    counts = records[:, 8:8 + 4 * O * C].contiguous().view(torch.int32).reshape(M, O, C)
    valid_obj = torch.arange(O, device=device)[None, :] < n_objects[:, None]
    counts = torch.where(valid_obj[:, :, None], counts, torch.zeros((), dtype=torch.int32, device=device))
    points = records[:, lay['points_offset']:lay['record_bytes']].contiguous().view(torch.float64).reshape(M, O, P, 3)
    kp_point = torch.zeros((M, O, C, S, 3), dtype=torch.float64, device=device)
    for c in range(C):
        for s_ in range(lay['cfg'][c]):
            keep = counts[:, :, c] > s_
            kp_point[:, :, c, s_] = torch.where(keep[:, :, None], points[:, :, lay['slot_of'][c] + s_], kp_point[:, :, c, s_])
    return {'n_objects': n_objects, 'flags': flags, 'kp_count': counts, 'kp_point': kp_point}


class RecordExchange:
    """The per-step gather of the 3D keypoint records on CUDA, pipelined against the decode.

    The records are COMPACT (``compact_layout``) and the decode kernel writes them itself (okp_decode_emit_*): a step is

        sink = exchange.begin()                                    # where this step's records go
        decoder.decode_batch(heat, depth, centers, tables, records=sink)
        gathered, done = exchange.end()                            # cross-rank completion, on the exchange stream

    ``end`` is asynchronous: the completion (a device-side barrier, or the NCCL collective) runs on the exchange's own
    stream after everything enqueued on the current stream so far, so the gather of step k overlaps the decode of step
    k + 1 (frames are independent). It returns the gathered ``[world * frames, record_bytes]`` uint8 tensor
    (``unpack_compact_records``) plus the event that marks it complete. ``begin`` makes the current stream wait for the
    completion of step k - 2 before step k may write: no rank runs more than two steps ahead of the slowest, so with
    ``depth`` = 4 buffers the result of step k stays valid until the decode of step k + 2 has been ISSUED on the reading
    rank (consume it on the compute stream before that).

    Transports:
      'peer'  the gather buffers are symmetric (peer-mapped) memory; completion is a device-side barrier, no NCCL on the
              data path. ``staged=False`` (default): the sink of a step IS the destination rank's buffer, the grouping
              kernel's record stores travel over NVLink / NVSwitch while it runs -- no pack kernel, no copy.
              ``staged=True``: the kernel writes the step's records into a LOCAL buffer and the exchange stream pushes them
              into the destination rank's buffer with one peer copy (copy engine, no SM, overlapped with the next step's
              decode). Measured on 8 GPUs gathering to one root: 0.545 ms per step direct, 0.551 ms staged (the same on 2
              GPUs) -- the record traffic is not what the last 7 % of weak-scaling efficiency go to.
      'nccl'  the sink is a local send buffer, NCCL all_gather_into_tensor / gather moves it.
      'auto'  'peer' when symmetric memory can be set up on this box, else 'nccl'.

    ``root``: None = all_gather (every rank ends up with every record: the kernel stores each record ``world`` times); an
    integer = gather to that rank only (north_star: "a final NVLink gather of the 3D keypoints"): one store per record,
    into the root's buffer -- only the root's returned tensor is meaningful.
    """
    LAG = 2

    def __init__(self, decoder, frames, world=None, rank=None, transport='auto', depth=4, group=None, root=None, staged=False):
        import ctypes
        from . import _abi
        self._ctypes, self._abi = ctypes, _abi
        self.group = group
        self.world = dist.get_world_size(group) if world is None else int(world)
        self.rank = (dist.get_rank(group) if dist.is_initialized() else 0) if rank is None else int(rank)
        self.device = decoder.device
        self.N = int(frames)
        self.O = int(decoder.params.max_objects)
        self.cfg = list(decoder.cfg)
        self.record_bytes = decoder.record_bytes()
        assert self.record_bytes == compact_layout(self.O, self.cfg)['record_bytes']
        self.depth = int(depth)
        if self.depth < self.LAG + 2:
            raise ValueError(f"depth must be at least {self.LAG + 2}")
        self.root = None if root is None else int(root)
        self.stream = torch.cuda.Stream(device=self.device)
        self.calls = 0
        self.handles = None
        shape = (self.world * self.N, self.record_bytes)
        if self.world == 1:
            transport = 'local'
        if transport in ('auto', 'peer'):
            try:
                self._setup_peer(shape)
                transport = 'peer'
            except Exception as error:                      # no symmetric memory on this box / build
                if transport == 'peer':
                    raise
                self.peer_error = f"{type(error).__name__}: {error}"
                transport = 'nccl'
        self.transport = transport
        self.staged = bool(staged) and transport == 'peer'
        if transport != 'peer':
            self.buffers = [torch.zeros(shape, dtype=torch.uint8, device=self.device) for _ in range(self.depth)]
        self.send = [torch.zeros((self.N, self.record_bytes), dtype=torch.uint8, device=self.device) for _ in range(self.depth)] \
            if transport == 'nccl' or self.staged else None
        if self.staged:                                      # this rank's rows inside every destination's gather buffer
            targets = range(self.world) if self.root is None else [self.root]
            self._remote_rows = [[handle.get_buffer(t, shape, torch.uint8)[self.rank * self.N:(self.rank + 1) * self.N]
                                  for t in targets] for handle in self.handles]
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self._sinks = [self._make_sink(slot) for slot in range(self.depth)]

    def _setup_peer(self, shape):
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if self.group is None else self.group
        self.buffers, self.handles = [], []
        for _ in range(self.depth):
            buffer = symm_mem.empty(shape, dtype=torch.uint8, device=self.device)
            handle = symm_mem.rendezvous(buffer, group)
            if len(handle.buffer_ptrs) != self.world:
                raise RuntimeError("symmetric memory rendezvous returned the wrong number of peers")
            buffer.zero_()
            self.buffers.append(buffer)
            self.handles.append(handle)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)                            # nobody stores into a buffer that is still being zeroed

    def _make_sink(self, slot):
        if self.transport == 'peer' and not self.staged:
            peers = [int(p) for p in self.handles[slot].buffer_ptrs]
            targets = peers if self.root is None else [peers[self.root]]
            first_row = self.rank * self.N
        elif self.transport == 'peer':
            targets, first_row = [self.send[slot].data_ptr()], 0
        elif self.transport == 'nccl':
            targets, first_row = [self.send[slot].data_ptr()], 0
        else:
            targets, first_row = [self.buffers[slot].data_ptr()], 0
        array = (self._ctypes.c_void_p * len(targets))(*targets)
        sink = self._abi.OkpRecordSink(buffers_dev=array, n_buffers=len(targets), record_bytes=self.record_bytes,
                                       first_row=first_row)
        sink._keep = array                                   # the struct only holds a pointer to it
        return sink

    def begin(self):
        """-> the OkpRecordSink of the coming step (pass it as ``records=`` to KeypointDecoder.decode_batch)."""
        k = self.calls
        if k >= self.LAG:
            torch.cuda.current_stream(self.device).wait_event(self.done[(k - self.LAG) % self.depth])
        return self._sinks[k % self.depth]

    def end(self):
        """After the decode of the step has been enqueued on the current stream: completion on the exchange stream.
        -> (gathered uint8 [world * frames, record_bytes], completion event)."""
        slot = self.calls % self.depth
        self.calls += 1
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        out = self.buffers[slot]
        with torch.cuda.stream(self.stream):
            if self.transport == 'peer':
                if self.staged:
                    for rows in self._remote_rows[slot]:
                        rows.copy_(self.send[slot], non_blocking=True)
                self.handles[slot].barrier(channel=0)       # every rank's records have landed
            elif self.transport == 'nccl':
                if self.root is None:
                    dist.all_gather_into_tensor(out, self.send[slot], group=self.group)
                else:
                    rows = [out[r * self.N:(r + 1) * self.N] for r in range(self.world)] if self.rank == self.root else None
                    dist.gather(self.send[slot], rows, dst=self.root, group=self.group)
            self.done[slot].record(self.stream)
        return out, self.done[slot]

    def finish(self):
        """Makes the current stream wait for every exchange issued so far."""
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
