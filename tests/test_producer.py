"""The input producer (object_keypoints_b200/producer.py, BASELINE config 5) against the UNMODIFIED
reference network: same parameter names / shapes, same outputs for the same weights and input
(fixture tests/golden/producer_valve.npz, made by oracle/make_goldens.py::producer_case)."""
import numpy as np
import pytest
import torch

from helpers import load_golden


def _golden_net():
    from object_keypoints_b200 import producer
    g = load_golden('producer_valve.npz')
    net = producer.KeypointNet(heatmaps_out=int(g['heatmaps_out'])).eval()
    return g, net, producer


def test_state_dict_has_the_reference_names_and_sizes():
    g, net, _ = _golden_net()
    mine = net.state_dict()
    assert sorted(mine) == [str(n) for n in g['parameter_names']]
    assert [mine[str(n)].numel() for n in g['parameter_names']] == [int(v) for v in g['parameter_sizes']]


def test_forward_matches_the_reference_network():
    g, net, producer = _golden_net()
    seed = int(g['seed'])
    net.load_state_dict(producer.deterministic_state_dict(net, seed=seed), strict=True)
    frames = torch.from_numpy(np.random.default_rng(seed).normal(0.0, 1.0, (1, 3, 511, 511)).astype(np.float32))
    with torch.no_grad():
        heat, depth, centers = net(frames)
    assert heat.shape == (1, 3, 64, 64) and depth.shape == (1, 3, 64, 64) and centers.shape == (1, 2, 2, 64, 64)
    for got, want, name in ((heat, g['heat'], 'heat'), (depth, g['depth'], 'depth'), (centers, g['centers'], 'centers')):
        scale = max(float(np.abs(want).max()), 1.0)
        assert float(np.abs(got.numpy() - want).max()) <= 2e-4 * scale, name     # float32, same op order up to conv algorithms
    assert float(heat.std()) > 1e-3, "degenerate output would make the comparison vacuous"


@pytest.mark.gpu
def test_config5_bf16_producer_feeds_the_decode_kernels_in_place():
    """BASELINE config 5 at a test-sized batch: random-init network, bf16, outputs stay on the device and
    go straight into okp_decode_bf16. Random-init heatmaps sit near 0.5 everywhere: ~100
    noise peaks per map, so this exercises the capacity / overflow flags; tables must be bitwise the C
    oracle's on the up-cast maps."""
    from oracle import c_oracle
    from object_keypoints_b200 import KeypointDecoder, producer, synthetic
    cfg = [1, 3]
    net = producer.build_producer(cfg, device='cuda', dtype=torch.bfloat16, seed=0)
    frames = torch.randn(4, 3, 511, 511, device='cuda', dtype=torch.bfloat16,
                         generator=torch.Generator('cuda').manual_seed(5)).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        heat, depth, centers = net(frames)
    assert heat.dtype == torch.bfloat16 and heat.is_contiguous() and heat.shape == (4, 3, 64, 64)
    camera = synthetic.default_camera((64, 64))
    decoder = KeypointDecoder(cfg, (64, 64), camera=camera, max_peaks=128, max_objects=128, max_votes=64)
    got = decoder.decode_batch(heat, depth, centers).numpy()
    want = c_oracle.decode(heat.float().cpu().numpy(), depth.float().cpu().numpy(), centers.float().cpu().numpy(), cfg,
                           camera, max_peaks=128, max_objects=128, max_votes=64)
    for key in ['peak_count', 'peak_yx', 'peak_object', 'n_objects', 'flags', 'kp_assigned', 'kp_count', 'kp_peak', 'n_votes']:
        np.testing.assert_array_equal(got[key], want[key], err_msg=key)
    for key in ['peak_score', 'peak_xy', 'peak_conf', 'kp_xy']:
        np.testing.assert_array_equal(got[key].view(np.uint32), want[key].view(np.uint32), err_msg=key)
    assert got['peak_count'].sum() > 0


@pytest.mark.gpu
def test_learned_pipeline_runs_a_traced_keypoint_net(tmp_path):
    """A9 end to end, the way scripts/package_model.py:21-42 + scripts/eval_model.py:278-293 use it: the network is traced
    to TorchScript, LearnedKeypointTrackingPipeline loads the file, runs it on the device and decodes its outputs in place.
    The objects must be those of the decoder called directly on the network's outputs, and (objects, heatmap) the return."""
    from object_keypoints_b200 import KeypointDecoder, LearnedKeypointTrackingPipeline, producer, synthetic
    from object_keypoints_b200.pipeline import tables_to_objects
    cfg = {'keypoint_config': [1, 3]}
    net = producer.build_producer(cfg, device='cuda', dtype=torch.float32, seed=0)
    frame = torch.randn(1, 3, 511, 511, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        traced = torch.jit.trace(net, frame.cuda().contiguous(memory_format=torch.channels_last))
    path = str(tmp_path / 'keypoint_net.pt')
    traced.save(path)
    options = dict(max_peaks=128, max_objects=128, max_votes=64)          # random-init maps: ~100 noise peaks per map
    pipeline = LearnedKeypointTrackingPipeline(path, True, [64, 64], None, cfg, **options)
    camera = synthetic.default_camera((64, 64))
    pipeline.reset(camera)
    objects, heatmap = pipeline(frame)
    assert heatmap.is_cuda and heatmap.shape == (1, 3, 64, 64)
    with torch.no_grad():
        heat, depth, centers = torch.jit.load(path).cuda()(frame.cuda())
    assert torch.equal(heat, heatmap)
    decoder = KeypointDecoder(cfg, (64, 64), camera=camera, **options)
    want = tables_to_objects(decoder.decode_batch(heat, depth, centers).numpy(), 0)
    assert len(objects) == len(want) > 0
    for a, b in zip(objects, want):
        assert set(a) == {'p_centers', 'keypoints', 'p_C'}
        for c in range(3):
            np.testing.assert_array_equal(a['keypoints'][c], b['keypoints'][c])
            if b['p_C'][c] is None:
                assert a['p_C'][c] is None
            else:
                np.testing.assert_array_equal(a['p_C'][c], b['p_C'][c])
