"""GPU: the CUDA-graph PoseEstimator (get_pose pipeline) returns exactly what the eager C-ABI path returns, for
uint8 and float inputs, and `get_pose` keeps the reference contract (tests/test_utils.py:82-87: (2,6,256,256) -> (2,7))."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_get_pose_contract(cuda_device):
    from argus_b200.models import NCameraCNN
    from argus_b200.utils import get_pose

    model = NCameraCNN().to(cuda_device).eval()
    pose = get_pose(torch.rand(2, 6, 256, 256, device=cuda_device), model)
    assert pose.shape == (2, 7)
    assert torch.allclose(pose[:, 3:].norm(dim=-1), torch.ones(2, device=cuda_device), atol=1e-5)


@pytest.mark.parametrize("B,u8", [(1, True), (4, True), (2, False)])
def test_pose_estimator_matches_eager(cuda_device, B, u8):
    from argus_b200.models import NCameraCNN
    from argus_b200.utils import PoseEstimator, get_pose, xyzxyzw_to_xyzwxyz_SE3

    torch.manual_seed(1)
    model = NCameraCNN().to(cuda_device).eval()
    est = PoseEstimator(model, B, 128, 128, uint8_input=u8)
    est_w = PoseEstimator(model, B, 128, 128, uint8_input=u8, wxyz=True)
    for seed in range(3):
        g = torch.Generator().manual_seed(seed)
        if u8:
            x = torch.randint(0, 256, (B, 2, 128, 128, 3), dtype=torch.uint8, generator=g)
        else:
            x = torch.rand(B, 6, 128, 128, generator=g)
        want = get_pose(x.to(cuda_device), model)
        got = est(x.pin_memory()).clone()
        assert torch.equal(got, want)
        assert torch.equal(est_w(x.to(cuda_device)).clone(), xyzxyzw_to_xyzwxyz_SE3(want))
    # a parameter update is picked up after re-creating the session (weights are baked into packed copies)
    with torch.no_grad():
        model.output_mlp._modules["4"].bias.add_(0.5)
    est2 = PoseEstimator(model, B, 128, 128, uint8_input=u8)
    assert not torch.equal(est2(x.to(cuda_device)), want)
