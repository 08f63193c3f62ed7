"""CPU: dataset contract of the reference (tests/test_data.py:25-54 and the fixture of tests/conftest.py:14-57) and
the quaternion-reorder golden vectors (tests/test_utils.py:17-47)."""
import numpy as np
import pytest
import torch

from argus_b200.dataset import CameraCubePoseDataset, CameraCubePoseDatasetConfig, write_dataset
from argus_b200.utils import xyzwxyz_to_xyzxyzw_SE3, xyzxyzw_to_xyzwxyz_SE3


@pytest.fixture(scope="module")
def dummy_data_path(tmp_path_factory):
    """15 random-noise 256x256 PNG pairs, train 10 / test 5, poses (x,y,z,qw,qx,qy,qz) — the reference fixture."""
    rng = np.random.default_rng(0)
    root = tmp_path_factory.mktemp("data") / "dummy"

    def poses(n):
        q = rng.normal(size=(n, 4))
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
        return np.concatenate([rng.normal(size=(n, 3)), q], -1)

    imgs = rng.integers(0, 256, (15, 2, 256, 256, 3), dtype=np.uint8)
    write_dataset(str(root), imgs[:10], poses(10), imgs[10:], poses(5))
    return str(root), imgs


def test_len_and_get_item(dummy_data_path):
    path, imgs = dummy_data_path
    cfg = CameraCubePoseDatasetConfig(dataset_path=path)
    train, test = CameraCubePoseDataset(cfg, train=True), CameraCubePoseDataset(cfg, train=False)
    assert len(train) == 10 and len(test) == 5
    ex = train[3]
    assert set(ex.keys()) == {"images", "cube_pose"}
    assert ex["images"].shape == (6, 256, 256) and ex["images"].dtype == torch.float32
    assert ex["cube_pose"].shape == (7,) and ex["cube_pose"].dtype == torch.float32
    want = torch.from_numpy(imgs[3]).permute(0, 3, 1, 2).reshape(6, 256, 256).float() / 255.0
    assert torch.equal(ex["images"], want)
    u8 = CameraCubePoseDataset(cfg, train=False, as_uint8=True)[0]["images"]
    assert u8.dtype == torch.uint8 and torch.equal(u8, torch.from_numpy(imgs[10]))


def test_center_crop(dummy_data_path):
    path, imgs = dummy_data_path
    ds = CameraCubePoseDataset(CameraCubePoseDatasetConfig(dataset_path=path, center_crop=(128, 128)), train=True)
    ex = ds[0]
    assert ex["images"].shape == (6, 128, 128)
    want = torch.from_numpy(imgs[0][:, 64:192, 64:192]).permute(0, 3, 1, 2).reshape(6, 128, 128).float() / 255.0
    assert torch.equal(ex["images"], want)


def test_pose_is_reordered_to_scalar_last(dummy_data_path):
    path, _ = dummy_data_path
    ds = CameraCubePoseDataset(CameraCubePoseDatasetConfig(dataset_path=path), train=True)
    z = np.load(path + "/dummy.npz")
    stored = z["train/cube_poses"][0]
    got = ds[0]["cube_pose"].numpy()
    assert np.allclose(got[:3], stored[:3]) and np.allclose(got[3:6], stored[4:7]) and np.isclose(got[6], stored[3])


def test_missing_path_raises():
    with pytest.raises(FileNotFoundError):
        CameraCubePoseDatasetConfig(dataset_path="/nonexistent/argus/data")


def test_quaternion_reorder_golden_vectors():
    """reference tests/test_utils.py:17-47"""
    a = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0])
    assert torch.equal(xyzwxyz_to_xyzxyzw_SE3(a), torch.tensor([1.0, 2.0, 3.0, 5.0, 6.0, 7.0, 4.0]))
    assert torch.equal(xyzxyzw_to_xyzwxyz_SE3(a), torch.tensor([1.0, 2.0, 3.0, 7.0, 4.0, 5.0, 6.0]))
    b = torch.arange(14.0).reshape(2, 7)
    assert torch.equal(xyzxyzw_to_xyzwxyz_SE3(xyzwxyz_to_xyzxyzw_SE3(b)), b)
    assert torch.equal(xyzwxyz_to_xyzxyzw_SE3(b)[1], torch.tensor([7.0, 8.0, 9.0, 11.0, 12.0, 13.0, 10.0]))
