"""Input producer for the decode path: the keypoint network, restated in plain PyTorch.

NOT the product (BASELINE.json: "The CornerNet-Squeeze backbone stays in PyTorch only as the input
producer"). It exists so that BASELINE config 5 -- a random-init CornerNet-Squeeze bf16 forward whose
head outputs stay on the device and feed the decode kernels -- can run without the reference tree,
and so that a checkpoint trained with the reference loads unchanged: parameter names and shapes are
those of the reference's ``KeypointNet.state_dict()`` (checked against the unmodified reference in
tests/golden/producer_*.npz, made by oracle/make_goldens.py).

Architecture restated from the reference (paths relative to its root):
  stem + two stacked hourglasses   perception/corner_net_lite/core/models/CornerNet_Squeeze.py:66-89
                                   perception/corner_net_lite/core/models/py_utils/modules.py:25-93
  fire module (squeeze 1x1, expand 1x1 || depthwise-grouped 3x3)           CornerNet_Squeeze.py:10-31
  residual / conv-bn-relu blocks   perception/corner_net_lite/core/models/py_utils/utils.py:142-184
  heatmap / depth / centre heads   perception/models.py:13-53
  deployed forward (sigmoid on the heatmaps, last stack only)              scripts/package_model.py:22-28
"""
import torch
from torch import nn
from torch.nn import functional as F

HOURGLASS_WIDTHS = (256, 256, 384, 384, 512)      # channels per depth (CornerNet_Squeeze.py:72)
HOURGLASS_REPEATS = (2, 2, 2, 2, 4)               # fire modules per depth
STACKS = 2
HEATMAP_BIAS = 0.01 / 0.99                        # perception/models.py:25-26


class ConvBnRelu(nn.Module):
    """`convolution` of the reference: conv (bias only without BN) -> BN -> ReLU."""

    def __init__(self, kernel, c_in, c_out, stride=1, with_bn=True):
        super().__init__()
        self.conv = nn.Conv2d(c_in, c_out, kernel, stride=stride, padding=(kernel - 1) // 2, bias=not with_bn)
        self.bn = nn.BatchNorm2d(c_out) if with_bn else nn.Sequential()

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class Residual(nn.Module):
    """Two 3x3 convs with BN; the skip is a strided 1x1 conv + BN whenever the shape changes."""

    def __init__(self, c_in, c_out, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, c_out, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(c_out)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(c_out)
        changes = stride != 1 or c_in != c_out
        self.skip = nn.Sequential(nn.Conv2d(c_in, c_out, 1, stride=stride, bias=False), nn.BatchNorm2d(c_out)) \
            if changes else nn.Sequential()

    def forward(self, x):
        y = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)))))
        return F.relu(y + self.skip(x))


class Fire(nn.Module):
    """Squeeze to c_out/2 by 1x1, expand by a 1x1 half and a grouped 3x3 half, BN, optional identity skip."""

    def __init__(self, c_in, c_out, stride=1):
        super().__init__()
        mid = c_out // 2
        self.conv1 = nn.Conv2d(c_in, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv_1x1 = nn.Conv2d(mid, c_out // 2, 1, stride=stride, bias=False)
        self.conv_3x3 = nn.Conv2d(mid, c_out // 2, 3, stride=stride, padding=1, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(c_out)
        self.identity = stride == 1 and c_in == c_out

    def forward(self, x):
        s = self.bn1(self.conv1(x))                                   # no activation after the squeeze
        y = self.bn2(torch.cat((self.conv_1x1(s), self.conv_3x3(s)), 1))
        return F.relu(y + x) if self.identity else F.relu(y)


def _fires(c_in, c_out, count, stride=1, widen_last=False):
    """`count` fire modules; the channel change sits on the first (default) or the last module."""
    if widen_last:
        return nn.Sequential(*[Fire(c_in, c_in) for _ in range(count - 1)], Fire(c_in, c_out))
    return nn.Sequential(Fire(c_in, c_out, stride=stride), *[Fire(c_out, c_out) for _ in range(count - 1)])


class Hourglass(nn.Module):
    """One recursion level: a full-resolution branch plus a stride-2 branch that is processed one level
    deeper and brought back by a transposed conv (the Squeeze variant pools by striding, modules.py:25-68)."""

    def __init__(self, depth, widths, repeats):
        super().__init__()
        here, below = widths[0], widths[1]
        self.up1 = _fires(here, here, repeats[0])
        self.max1 = nn.Sequential()
        self.low1 = _fires(here, below, repeats[0], stride=2)
        self.low2 = Hourglass(depth - 1, widths[1:], repeats[1:]) if depth > 1 else _fires(below, below, repeats[1])
        self.low3 = _fires(below, here, repeats[0], widen_last=True)
        self.up2 = nn.ConvTranspose2d(here, here, 4, stride=2, padding=1)

    def forward(self, x):
        return self.up1(x) + self.up2(self.low3(self.low2(self.low1(x))))


class SqueezeHourglassBackbone(nn.Module):
    """frames [N,3,511,511] -> one 256-channel 64x64 feature map per stack."""

    def __init__(self):
        super().__init__()
        self.pre = nn.Sequential(ConvBnRelu(7, 3, 128, stride=2), Residual(128, 256, stride=2), Residual(256, 256, stride=2))
        self.hgs = nn.ModuleList([Hourglass(4, HOURGLASS_WIDTHS, HOURGLASS_REPEATS) for _ in range(STACKS)])
        self.cnvs = nn.ModuleList([ConvBnRelu(3, 256, 256) for _ in range(STACKS)])
        self.inters = nn.ModuleList([Residual(256, 256) for _ in range(STACKS - 1)])
        merge = lambda: nn.Sequential(nn.Conv2d(256, 256, 1, bias=False), nn.BatchNorm2d(256))
        self.inters_ = nn.ModuleList([merge() for _ in range(STACKS - 1)])
        self.cnvs_ = nn.ModuleList([merge() for _ in range(STACKS - 1)])

    def forward(self, x):
        inter = self.pre(x)
        features = []
        for i in range(STACKS):
            feature = self.cnvs[i](self.hgs[i](inter))
            features.append(feature)
            if i + 1 < STACKS:
                inter = self.inters[i](F.relu(self.inters_[i](inter) + self.cnvs_[i](feature)))
        return features


def _head(width, outputs):
    return nn.Sequential(ConvBnRelu(1, 256, width), ConvBnRelu(1, width, 32), nn.Conv2d(32, outputs, 1))


class _Heads(nn.Module):
    def __init__(self, width, outputs):
        super().__init__()
        self.output_head1 = _head(width, outputs)          # intermediate supervision (training only)
        self.output_head2 = _head(width, outputs)


class KeypointNet(nn.Module):
    """perception/models.py:60-91 with the deployed forward of scripts/package_model.py:22-28:
    ``forward(frames) -> (sigmoid(heatmap) [N,C,64,64], depth [N,C,64,64], centers [N,C-1,2,64,64])`` from the
    LAST stack; the first stack's heads exist only so that reference checkpoints load with strict=True."""

    def __init__(self, heatmaps_out=2, features=128):
        super().__init__()
        self.heatmaps_out = heatmaps_out
        self.backbone = SqueezeHourglassBackbone()
        self.heatmap_head = _Heads(features, heatmaps_out)
        self.depth_head = _Heads(features, heatmaps_out)
        self.center_head = _Heads(features, (heatmaps_out - 1) * 2)
        for head in (self.heatmap_head.output_head1, self.heatmap_head.output_head2):
            nn.init.constant_(head[-1].bias, HEATMAP_BIAS)

    def forward(self, frames):
        feature = self.backbone(frames)[-1]
        heat = torch.sigmoid(self.heatmap_head.output_head2(feature))
        depth = self.depth_head.output_head2(feature)
        centers = self.center_head.output_head2(feature)
        N, _, H, W = centers.shape
        # the decode kernels stream dense NCHW maps; the heads are tiny (C <= 16 channels), so this is free
        return (heat.contiguous(memory_format=torch.contiguous_format),
                depth.contiguous(memory_format=torch.contiguous_format),
                centers.contiguous(memory_format=torch.contiguous_format).reshape(N, self.heatmaps_out - 1, 2, H, W))


def build_producer(keypoint_config, device=None, dtype=torch.bfloat16, seed=0, state_dict=None):
    """Random-init (or checkpoint-loaded) network for ``keypoint_config`` in eval mode, channels_last, `dtype`."""
    from . import _abi
    cfg = _abi.check_keypoint_config(keypoint_config)
    generator_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = KeypointNet(heatmaps_out=1 + len(cfg))
    torch.random.set_rng_state(generator_state)
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    net = net.eval().to(dtype=dtype)
    if device is not None:
        net = net.to(device)
    return net.to(memory_format=torch.channels_last)


def deterministic_state_dict(module, seed=0):
    """A state dict that depends only on parameter NAMES and shapes (not on construction order or torch's
    RNG stream), so that the reference network and this restatement can be given identical weights without
    shipping 100 MB of them: He-scaled normal conv weights, BN statistics near identity."""
    import zlib
    import numpy as np
    out = {}
    for name, tensor in module.state_dict().items():
        rng = np.random.default_rng(zlib.crc32(name.encode()) + seed)
        shape = tuple(tensor.shape)
        if name.endswith('num_batches_tracked'):
            value = np.zeros(shape, np.int64)
        elif name.endswith('running_var'):
            value = rng.uniform(0.5, 1.5, shape)
        elif name.endswith('running_mean'):
            value = rng.normal(0.0, 0.1, shape)
        elif tensor.dim() == 1 and name.endswith('weight'):
            value = rng.uniform(0.8, 1.2, shape)
        elif name.endswith('bias'):
            value = rng.normal(0.0, 0.05, shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            value = rng.normal(0.0, (2.0 / fan_in) ** 0.5, shape)
        out[name] = torch.from_numpy(np.asarray(value)).to(tensor.dtype)
    return out
