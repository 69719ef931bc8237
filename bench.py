#!/usr/bin/env python
"""bench.py -- frames/s of the heatmap -> 3D keypoint path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path (K1 peaks + merge, K3/K4 grouping + 3D; for N > 1 followed by the
NCCL all_gather of the 3D keypoint records) over one batch of synthetic network outputs that is
already resident in HBM. Frames shard independently: every rank decodes its own `frames` frames
(weak scaling), `value` is the whole-job frames/s = N * frames / max-over-ranks step time.

Workloads (BASELINE.json configs):
  config4_180x320  4096 valve frames [1,3] of 180x320 per GPU, 8 objects per frame   (default)
  config4_64x64    4096 valve frames of 64x64 per GPU, 2 objects per frame
  config2_cups     128 cups frames [1,1,1] of 64x64 (64 stereo pairs), 1-4 objects

One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    'config4_180x320': dict(cfg=[1, 3], size=(180, 320), frames=4096, grid=(4, 2), config='valve.json'),
    'config4_64x64': dict(cfg=[1, 3], size=(64, 64), frames=4096, grid=(2, 1), config='valve.json'),
    'config2_cups': dict(cfg=[1, 1, 1], size=(64, 64), frames=128, grid=(2, 2), config='cups.json'),
}
METRIC = "frames/sec heatmap->3D keypoints"
UNIT = "frames/s"


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'sm_max_mhz': 1965.0}, 'fallback'       # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, sm_max, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                sm_max.append(float(parts[1]))
            except ValueError:
                continue
            for name, value in zip(names, parts[3:7]):
                if value.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(sm_max) if sm_max else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_inputs(workload, device, seed):
    import torch
    from object_keypoints_b200 import synthetic
    w = WORKLOADS[workload]
    if workload == 'config2_cups':
        batch = synthetic.make_batch(w['frames'], w['cfg'], w['size'], seed=1002 + seed, objects=(1, 4))
        return (torch.from_numpy(batch.heat).to(device), torch.from_numpy(batch.depth).to(device),
                torch.from_numpy(batch.centers).to(device))
    heat, depth, centers, _ = synthetic.torch_grid_batch(w['frames'], w['cfg'], w['size'], seed=1004 + seed,
                                                         grid=w['grid'], device=device)
    return heat, depth, centers


def cpu_baseline(workload, heat, depth, centers, budget_s=12.0):
    """The C oracle (OpenMP, every host core) on a bounded sample of the very same frames."""
    from object_keypoints_b200 import synthetic
    from oracle import c_oracle
    w = WORKLOADS[workload]
    camera = synthetic.default_camera(w['size'])
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else c_oracle.max_threads()
    probe = min(64, heat.shape[0])
    h, d, c = heat[:probe].cpu().numpy(), depth[:probe].cpu().numpy(), centers[:probe].cpu().numpy()
    c_oracle.decode(h, d, c, w['cfg'], camera, threads=cores)                       # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    c_oracle.decode(h, d, c, w['cfg'], camera, threads=cores)
    rate = probe / (time.perf_counter() - t0)
    sample = int(max(probe, min(heat.shape[0], rate * budget_s)))
    h, d, c = heat[:sample].cpu().numpy(), depth[:sample].cpu().numpy(), centers[:sample].cpu().numpy()
    t0 = time.perf_counter()
    c_oracle.decode(h, d, c, w['cfg'], camera, threads=cores)
    elapsed = time.perf_counter() - t0
    return {'value': sample / elapsed, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f"first {sample} frames of {workload} (C/OpenMP restatement oracle/okp_oracle.c, {elapsed:.2f} s)"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the C/OpenMP port; the Python original cannot
    travel to the GPU box) on every host core, same workload/metric; rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from object_keypoints_b200 import synthetic
    from oracle import c_oracle
    w = WORKLOADS[args.workload]
    camera = synthetic.default_camera(w['size'])
    sample = min(w['frames'], 512 if w['size'] != (64, 64) else 2048)
    if w['size'] == (64, 64) and args.workload == 'config2_cups':
        batch = synthetic.make_batch(w['frames'], w['cfg'], w['size'], seed=1002, objects=(1, 4))
        heat, depth, centers = batch.heat, batch.depth, batch.centers
    else:
        device = 'cuda' if torch.cuda.is_available() else 'cpu'
        h, d, c, _ = synthetic.torch_grid_batch(sample, w['cfg'], w['size'], seed=1004, grid=w['grid'], device=device, chunk=64)
        heat, depth, centers = h.cpu().numpy(), d.cpu().numpy(), c.cpu().numpy()
    sample = heat.shape[0]
    # every host core, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for its workers)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    for _ in range(max(args.warmup, 1)):
        c_oracle.decode(heat, depth, centers, w['cfg'], camera, threads=cores)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        c_oracle.decode(heat, depth, centers, w['cfg'], camera, threads=cores)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'keypoint_config': w['config'], 'frames_per_step': sample,
                   'prediction_size': list(w['size'])},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f"{sample} frames of {args.workload} per step, C/OpenMP port of perception/pipeline.py"},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from object_keypoints_b200 import KeypointDecoder, synthetic
    from object_keypoints_b200.pipeline import DecodeTables
    from object_keypoints_b200.sharding import RecordExchange

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    w = WORKLOADS[args.workload]
    frames = args.frames or w['frames']
    w = dict(w, frames=frames)
    WORKLOADS[args.workload] = w
    H, W = w['size']
    C = 1 + len(w['cfg'])
    camera = synthetic.default_camera(w['size'])
    heat, depth, centers = make_inputs(args.workload, device, seed=rank)
    decoder = KeypointDecoder(w['cfg'], w['size'], camera=camera, device=device)
    # two table sets: the gather of step k (own stream) overlaps the decode of step k + 1
    table_sets = [decoder.tables(frames), DecodeTables(frames, decoder.C, decoder.cfg, decoder.params, device)]
    tables = table_sets[0]
    root = 0 if args.exchange == 'gather' else None
    exchange = RecordExchange(tables, world=world, rank=rank, transport=args.transport, root=root) if world > 1 else None
    gathered = None
    table_free = [None, None]
    steps_done = [0]

    def step():
        index = steps_done[0] % 2
        steps_done[0] += 1
        current = table_sets[index]
        if table_free[index] is not None:
            torch.cuda.current_stream().wait_event(table_free[index])     # its previous gather has read it
        decoder.extract_peaks(heat, current)
        decoder.group_objects(depth, centers, current)
        if world > 1:
            result, done = exchange.exchange(current)
            table_free[index] = done
            return result
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        gathered = step()
    barrier()

    # ---- timed region: exactly K steps, device time, K1 bracketed by its own events ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    k1_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for i in range(args.steps):
        index = steps_done[0] % 2
        steps_done[0] += 1
        current = table_sets[index]
        if table_free[index] is not None:
            torch.cuda.current_stream().wait_event(table_free[index])
        k1_events[i][0].record()
        decoder.extract_peaks(heat, current)
        k1_events[i][1].record()
        decoder.group_objects(depth, centers, current)
        if world > 1:
            gathered, table_free[index] = exchange.exchange(current)
    if world > 1:
        exchange.finish()                               # every gather of the K steps is inside the timed region
    stop.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = start.elapsed_time(stop)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events) / args.steps

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region ----
    e2e_frames = min(frames, args.e2e_frames)
    host = [t[:e2e_frames].cpu().pin_memory() for t in (heat, depth, centers)]
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    result_host = None
    for _ in range(2):
        result_host = decoder.decode_host_batch(*host)
    barrier()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(e2e_steps):
        result_host = decoder.decode_host_batch(*host)
    t_stop.record()
    barrier()
    e2e_ms = t_start.elapsed_time(t_stop) / e2e_steps
    d2h = sum(v.numel() * v.element_size() for v in result_host.values())
    h2d = decoder.host_bytes_copied                     # bytes the copy engine moved; depth / centre maps are gathered in place

    times = torch.tensor([elapsed_ms, k1_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, k1_ms, e2e_ms = [float(v) for v in times.cpu()]

    # sanity: the decode found the objects that were drawn (guards against timing an empty kernel)
    found = float((tables['n_objects'] > 0).float().mean())
    assert found > 0.99, f"only {found:.3f} of the frames produced objects"
    if world > 1 and (rank == 0 or args.exchange == 'allgather'):   # the gathered records hold every rank's frames
        objects = gathered[:, 0].reshape(world, frames)
        assert bool((objects > 0).float().mean(dim=1).gt(0.99).all()), "gathered records are incomplete"

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        ms_per_step = elapsed_ms / args.steps
        value = world * frames / (ms_per_step / 1e3)
        algorithmic_bytes = frames * C * H * W * 4                 # every heatmap byte exactly once
        achieved = algorithmic_bytes / (k1_ms / 1e3) / 1e9
        traffic = None
        ncu_summary = os.path.join(ROOT, 'profiles', 'r01_k1_traffic.json')
        if os.path.exists(ncu_summary):
            with open(ncu_summary) as f:
                traffic = json.load(f).get(args.workload)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload, 'keypoint_config': w['config'], 'frames_per_gpu': frames,
                       'prediction_size': [H, W], 'objects_per_frame': w['grid'][0] * w['grid'][1],
                       'l2': f"inputs {algorithmic_bytes / 1e6:.0f} MB heatmaps per step exceed the 126 MB L2; no flush needed",
                       'parallelism': f"frames sharded over {world} GPU(s); per step one {'gather to rank 0' if args.exchange == 'gather' else 'all_gather'} of the 3D keypoint records"
                                      + (f" ({exchange.transport}: "
                                         + ('pack kernel stores straight into every peer over NVLink' if exchange.transport == 'peer'
                                            else 'pack kernel + NCCL all_gather') + ", overlapped with the next step's decode)"
                                         if world > 1 else "")},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                         'frac': achieved / peaks['hbm_gbs'], 'traffic': traffic, 'peak_source': peak_kind,
                         'kernel': 'K1 peaks (box sum + NMS + centroid) + merge', 'kernel_ms': k1_ms,
                         'algorithmic_bytes': algorithmic_bytes},
            'e2e': {'value': world * e2e_frames / (e2e_ms / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'frames_per_step': e2e_frames, 'steps': e2e_steps,
                    'host_input_bytes_per_step': sum(t.numel() * t.element_size() for t in host),
                    'sparse_chunks': decoder.host_chunks_sparse,
                    'note': 'pinned host tensors in, pinned host tables out, chunks of 128 frames pipelined (host pass / PCIe / '
                            'decode). Heatmaps: a host pass (OpenMP + AVX2, inside the timed region) marks the 4x16-pixel tiles '
                            'within reach of a value above threshold / 25 and only those cross PCIe (bit-identical tables, '
                            'csrc/okp_sparse.cuh; dense copy when more than half of a chunk is marked). Depth and centre maps '
                            '(gather-only) are read in place from pinned host memory over PCIe'},
            'gpu_launches': args.steps * (3 + (1 if world > 1 else 0)),
            'clocks': clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(args.workload, heat, depth, centers)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='config4_180x320', choices=sorted(WORKLOADS))
    ap.add_argument('--frames', type=int, default=0, help="frames per GPU (default: the workload's)")
    ap.add_argument('--e2e-frames', type=int, default=2048)
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--exchange', default='gather', choices=['gather', 'allgather'],
                    help="N > 1: gather the records to rank 0 (north_star) or to every rank")
    ap.add_argument('--transport', default='auto', choices=['auto', 'peer', 'nccl'],
                    help="N > 1: how the keypoint records are gathered (sharding.RecordExchange)")
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
