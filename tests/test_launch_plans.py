"""Launch plans of the peak kernel (host code of csrc/okp_peaks_stream.cuh, compiled with nvcc; no GPU needed): the shapes the
bench runs must keep TWO CTAs per SM -- at most 320 threads at 96 registers and at most ~113 KB of shared memory per CTA. Round 2
found the 64x64 bfloat16 plan at 352 threads, i.e. one CTA per SM and 719 us instead of 448 us; this pins the fix."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def plans(tmp_path_factory):
    if shutil.which('nvcc') is None:
        pytest.skip("nvcc not on PATH")
    out = str(tmp_path_factory.mktemp('plans') / 'print_plan')
    run = subprocess.run(['nvcc', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', out,
                          os.path.join(ROOT, 'tools', 'print_plan.cu')], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    text = subprocess.run([out], capture_output=True, text=True, timeout=60).stdout
    rows = []
    for line in text.splitlines():
        m = re.match(r'\s*(\d+)x\s*(\d+) esize (\d) fused (\d): M\s*(\d+) strips\s*(\d+) compute threads\s*(\d+) EW (\d) total threads\s*(\d+) '
                     r'smem\s*(\d+) NS (\d) nb (\d+) groups (\d+) F (\d+)', line)
        if m:
            keys = ('H', 'W', 'esize', 'fused', 'M', 'strips', 'compute', 'EW', 'threads', 'smem', 'NS', 'nb', 'groups', 'F')
            rows.append(dict(zip(keys, (int(v) for v in m.groups()))))
    assert rows, text
    return rows


def test_bench_shapes_keep_two_ctas_per_sm(plans):
    seen = set()
    for p in plans:
        if (p['H'], p['W']) not in ((180, 320), (64, 64)):
            continue
        seen.add((p['H'], p['W'], p['esize'], p['fused']))
        assert p['threads'] <= 320, p                      # 2 x 320 threads x 96 registers = 61440 of the SM's 65536
        assert p['smem'] <= 113 * 1024, p                  # 2 x (smem + 1 KB reserved) within 227 KB
        assert p['compute'] == p['M'] * p['strips'] and p['threads'] == (p['compute'] + 31) // 32 * 32 + 32 + 32 * p['EW'], p
    assert len(seen) == 8


def test_small_maps_take_the_small_map_plan(plans):
    for p in plans:
        if (p['H'], p['W']) == (64, 64):
            assert p['NS'] == 3 and p['compute'] <= 224 and p['EW'] == 2, p
        if (p['H'], p['W']) == (180, 320):
            assert p['NS'] == 4 and p['M'] == 3 and p['EW'] == 1, p
        if p['fused']:
            assert p['F'] >= 1 and p['M'] % 3 == 0, p      # a fused group is a whole number of frames (C = 3 in print_plan.cu)
