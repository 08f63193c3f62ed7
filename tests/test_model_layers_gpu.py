"""Layer-by-layer GPU parity: every bottleneck output of our network is compared with the fp32 reference
restatement, and — as a yardstick for what bf16 arithmetic alone costs — with the same reference run under
torch.autocast(bfloat16). A wiring / kernel bug shows up as a jump at one block that the autocast run does not have."""
import pytest
import torch

from gpu_util import rel, structured_images

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W,training,structured", [
    (8, 128, 128, True, True), (2, 256, 256, True, True), (8, 128, 128, True, False), (2, 128, 128, False, True)])
def test_blockwise_activations(cuda_device, B, H, W, training, structured):
    from argus_b200.models import NCameraCNN
    from oracle.ref_model import make_reference_model

    ref = make_reference_model(42).to(cuda_device)
    ours = NCameraCNN().to(cuda_device)
    ours.load_state_dict(ref.state_dict())
    ref.train(training); ours.train(training)
    if structured:
        x = structured_images(B, 6, H, W, 11, cuda_device)
    else:
        x = torch.rand(B, 6, H, W, generator=torch.Generator().manual_seed(11)).to(cuda_device)

    taps = []
    mods = [ref.resnet.maxpool] + [blk for layer in (ref.resnet.layer1, ref.resnet.layer2, ref.resnet.layer3,
                                                      ref.resnet.layer4) for blk in layer] + [ref.resnet.avgpool, ref.resnet.fc]
    hooks = [m.register_forward_hook(lambda _m, _i, o: taps.append(o.detach().float())) for m in mods]
    with torch.no_grad():
        y_ref = ref(x)
        fp32 = list(taps)
        taps.clear()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y_ac = ref(x).float()
        ac = list(taps)
        y = ours(x)
    for h in hooks:
        h.remove()

    def to_rows(t):
        if t.dim() == 4:
            return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
        return t.reshape(t.shape[0], -1)

    print(f"\n[B={B} {H}x{W} training={training} structured={structured}] block: ours-vs-fp32  autocast-vs-fp32  ours-vs-autocast")
    for i in range(len(mods)):
        mine = ours.probe_activation(i - 1).float()
        r_ours = rel(mine, to_rows(fp32[i]))
        r_ac = rel(to_rows(ac[i]), to_rows(fp32[i]))
        r_x = rel(mine, to_rows(ac[i]))
        print(f"blk {i - 1:3d}: {r_ours:.4e}   {r_ac:.4e}   {r_x:.4e}")
        # our arithmetic (bf16 storage, fp32 accumulate) must not be meaningfully worse than torch's own bf16 path
        assert r_ours < max(1.25 * r_ac, 2e-2), (i - 1, r_ours, r_ac)
    print("out:", rel(y, y_ref), rel(y_ac, y_ref), rel(y, y_ac))
    assert rel(y, y_ref) < max(1.25 * rel(y_ac, y_ref), 2e-2)
    if not training:
        assert rel(y, y_ref) < 2e-2  # the north star's bf16 tolerance holds outright for inference
