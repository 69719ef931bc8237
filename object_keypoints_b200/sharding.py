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


class RecordExchange:
    """The per-step gather of the 3D keypoint records on CUDA, pipelined against the decode.

    ``exchange(tables)`` is asynchronous: it runs on its own stream after everything enqueued on the current
    stream so far, so the gather of step k overlaps the decode of step k + 1 (frames are independent, SURVEY.md
    8e). It returns the gathered ``[world * frames, R]`` float64 tensor plus the event that marks it complete;
    the tensor stays valid until ``exchange`` has been called ``depth - 1`` more times.

    Transports:
      'peer'  the buffers are symmetric (peer-mapped) memory; ONE kernel (okp_pack_records_f64) packs the rank's
              records and stores them straight into every rank's buffer over NVLink / NVSwitch, followed by a
              device-side barrier. No NCCL on the data path.
      'nccl'  the same kernel packs into a local send buffer, NCCL all_gather_into_tensor moves it.
      'auto'  'peer' when symmetric memory can be set up on this box, else 'nccl'.

    ``root``: None = all_gather (every rank ends up with every record); an integer = gather to that rank only
    (north_star: "a final NVLink gather of the 3D keypoints"): each rank stores its records once, into the root's
    buffer, instead of ``world`` times -- only the root's returned tensor is meaningful.
    """

    def __init__(self, tables, world=None, rank=None, transport='auto', depth=3, group=None, root=None):
        import ctypes
        from . import _lib
        self._ctypes, self._lib_module, self._lib = ctypes, _lib, _lib.lib()
        self.group = group
        self.world = dist.get_world_size(group) if world is None else int(world)
        self.rank = (dist.get_rank(group) if dist.is_initialized() else 0) if rank is None else int(rank)
        t = tables.tensors
        self.device = t['kp_point'].device
        self.N = int(t['n_objects'].shape[0])
        _, self.O, self.C, self.S = (int(v) for v in t['kp_point'].shape[:4])
        self.R = int(self._lib.okp_record_doubles(self.O, self.C, self.S))
        self.depth = int(depth)
        self.root = None if root is None else int(root)
        self.stream = torch.cuda.Stream(device=self.device)
        self.calls = 0
        self.handles = None
        shape = (self.world * self.N, self.R)
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
        if transport != 'peer':
            self.buffers = [torch.empty(shape, dtype=torch.float64, device=self.device) for _ in range(self.depth)]
            self.send = [torch.empty((self.N, self.R), dtype=torch.float64, device=self.device) for _ in range(self.depth)] \
                if transport == 'nccl' else None
        self.done = [torch.cuda.Event() for _ in range(self.depth)]

    def _setup_peer(self, shape):
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if self.group is None else self.group
        self.buffers, self.handles = [], []
        for _ in range(self.depth):
            buffer = symm_mem.empty(shape, dtype=torch.float64, device=self.device)
            handle = symm_mem.rendezvous(buffer, group)
            if len(handle.buffer_ptrs) != self.world:
                raise RuntimeError("symmetric memory rendezvous returned the wrong number of peers")
            self.buffers.append(buffer)
            self.handles.append(handle)

    def _pack(self, tables, first_row, destinations):
        array = (self._ctypes.c_void_p * len(destinations))(*destinations)
        rc = self._lib.okp_pack_records_f64(self._ctypes.byref(tables.struct), self.N, self.O, self.C, self.S, first_row,
                                            array, len(destinations), self._ctypes.c_void_p(self.stream.cuda_stream))
        self._lib_module.check(rc, 'okp_pack_records_f64')

    def exchange(self, tables):
        slot = self.calls % self.depth
        self.calls += 1
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        out = self.buffers[slot]
        with torch.cuda.stream(self.stream):
            if self.transport == 'peer':
                peers = [int(p) for p in self.handles[slot].buffer_ptrs]
                if self.root is not None:                   # gather: one copy, into the root's buffer (and our own rows)
                    peers = [peers[self.root]]
                self._pack(tables, self.rank * self.N, peers)
                self.handles[slot].barrier(channel=0)       # every rank's stores have landed everywhere
            elif self.transport == 'nccl':
                self._pack(tables, 0, [self.send[slot].data_ptr()])
                if self.root is None:
                    dist.all_gather_into_tensor(out, self.send[slot], group=self.group)
                else:
                    rows = [out[r * self.N:(r + 1) * self.N] for r in range(self.world)] if self.rank == self.root else None
                    dist.gather(self.send[slot], rows, dst=self.root, group=self.group)
            else:
                self._pack(tables, 0, [out.data_ptr()])
            self.done[slot].record(self.stream)
        return out, self.done[slot]

    def finish(self):
        """Makes the current stream wait for every exchange issued so far."""
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
