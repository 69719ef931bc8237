"""The Hartley-Sturm arithmetic of csrc/okp_stereo.cuh is __host__ __device__: compiled for the host (nvcc, no GPU needed) it
must reproduce the cv2.correctMatches goldens and the NumPy oracle -- the same statements the GPU kernel runs, so the
convergence rule of the root finder (which ends at the rounding floor instead of asking every root for 4e-16) is pinned on
the CPU as well."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def host_binary(tmp_path_factory):
    if shutil.which('nvcc') is None:
        pytest.skip("nvcc not on PATH")
    out = str(tmp_path_factory.mktemp('stereo') / 'host_check_stereo')
    run = subprocess.run(['nvcc', '-O2', '-fmad=false', '-std=c++17', '-o', out, os.path.join(ROOT, 'tools', 'host_check_stereo.cu')],
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    return out


def correct_on_host(binary, F, left, right):
    lines = [str(len(left)), ' '.join(repr(float(v)) for v in np.asarray(F).ravel())]
    lines += [f"{a[0]!r} {a[1]!r} {b[0]!r} {b[1]!r}" for a, b in zip(np.asarray(left).tolist(), np.asarray(right).tolist())]
    run = subprocess.run([binary], input='\n'.join(lines) + '\n', capture_output=True, text=True, timeout=120)
    assert run.returncode == 0
    values = np.array([[float(v) for v in line.split()] for line in run.stdout.strip().split('\n')])
    return values[:, :2], values[:, 2:]


def test_host_build_of_the_kernel_arithmetic_is_cv2(host_binary):
    g = load_golden('geometry.npz')
    left, right = correct_on_host(host_binary, g['F'], g['pairs_undistorted_left'], g['pairs_undistorted_right'])
    assert np.abs(left - g['pairs_corrected_left']).max() < 1e-9 and np.abs(right - g['pairs_corrected_right']).max() < 1e-9
    for i in range(len(g['hs_F'])):                      # forward motion (epipole inside the image), general pose
        left, right = correct_on_host(host_binary, g['hs_F'][i], g['hs_left'][i], g['hs_right'][i])
        assert np.abs(left - g['hs_corrected_left'][i]).max() < 1e-9
        assert np.abs(right - g['hs_corrected_right'][i]).max() < 1e-9


def test_host_build_matches_the_oracle_on_random_pairs(host_binary):
    from oracle import np_oracle
    g = load_golden('geometry.npz')
    rng = np.random.default_rng(11)
    n = 2000
    left = rng.uniform(100, 1100, (n, 2))
    right = left + np.stack([rng.uniform(-60, -5, n), rng.normal(0, 0.4, n)], axis=1)
    want_l, want_r = np_oracle.correct_matches(g['F'], left, right)
    got_l, got_r = correct_on_host(host_binary, g['F'], left, right)
    assert np.abs(got_l - want_l).max() < 1e-9 and np.abs(got_r - want_r).max() < 1e-9
