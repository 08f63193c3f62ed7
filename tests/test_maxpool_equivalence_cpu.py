"""CPU: the max-pool forward kernel takes the window arg-max on sign(scale) * x instead of on relu(x * scale + shift)
(csrc/elementwise.cu::maxpool_fwd_kernel). This test states that rule in torch and checks, against torch's own
max_pool2d + autograd, that (1) the pooled values are identical and (2) the gradient that reaches the batch-norm output
after the ReLU backward is identical -- windows that are entirely <= 0 after the ReLU may pick another tap, but the ReLU
mask zeroes both choices. (Reference semantics: torchvision resnet50's relu + maxpool, argus/models.py:84.)"""
import torch
import torch.nn.functional as F


def raw_argmax_pool(x, scale, shift):
    """x (N, C, H, W) float64. Returns pooled values and the flat arg-max position per window (first maximum)."""
    N, C, H, W = x.shape
    sgn = torch.where(scale < 0, -1.0, 1.0).view(1, C, 1, 1).to(x.dtype)
    xs = F.pad(x * sgn, (1, 1, 1, 1), value=float("-inf"))
    pos = F.pad(torch.arange(H * W, dtype=torch.float64).view(1, 1, H, W).expand(N, C, H, W), (1, 1, 1, 1), value=-1.0)
    win = xs.unfold(2, 3, 2).unfold(3, 3, 2).reshape(N, C, H // 2, W // 2, 9)
    wpos = pos.unfold(2, 3, 2).unfold(3, 3, 2).reshape(N, C, H // 2, W // 2, 9)
    best = win.argmax(-1, keepdim=True)          # torch.argmax returns the first maximum
    idx = wpos.gather(-1, best).squeeze(-1).long()
    raw = x.reshape(N, C, H * W).gather(2, idx.reshape(N, C, -1)).reshape(N, C, H // 2, W // 2)
    pooled = torch.relu(raw * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1))
    return pooled, idx


def test_raw_argmax_equals_activated_argmax_after_relu_mask():
    g = torch.Generator().manual_seed(0)
    N, C, H, W = 2, 16, 12, 10
    x = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    scale = torch.randn(C, generator=g, dtype=torch.float64)          # both signs
    scale[scale.abs() < 0.05] = 0.3                                    # gamma == 0 is the one case the rule does not cover
    shift = torch.randn(C, generator=g, dtype=torch.float64) - 0.5     # many all-negative windows
    bn = (x * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1)).requires_grad_(True)
    act = torch.relu(bn)
    ref = F.max_pool2d(act, 3, 2, 1)
    dy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    ref.backward(dy)
    want = bn.grad                                                     # gradient at the batch-norm output

    pooled, idx = raw_argmax_pool(x, scale, shift)
    assert torch.equal(pooled, ref.detach())
    routed = torch.zeros(N, C, H * W, dtype=torch.float64).scatter_add_(2, idx.reshape(N, C, -1), dy.reshape(N, C, -1))
    got = routed.reshape(N, C, H, W) * (bn.detach() > 0)               # ReLU backward
    assert torch.equal(got, want)
    # and the share of windows where the two arg-max rules disagree is not negligible: the mask really is what saves it
    assert (ref.detach() == 0).double().mean() > 0.05
