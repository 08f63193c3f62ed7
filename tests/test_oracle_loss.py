"""CPU: pins the float64 loss oracle (oracle/se3_loss.py) against (a) the reference's own known-answer test,
(b) an independent ground truth (scipy expm/logm on 4x4 matrices, central finite differences), (c) the committed
golden vectors; and checks that the kernel's host-compiled SE(3) arithmetic equals the oracle."""
import ctypes
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import se3_loss as o

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def rand_poses(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    return np.concatenate([rng.normal(size=(n, 3)), q], -1)


def test_reference_known_answer_and_shapes():
    """Reference tests/test_train.py:18-36: shapes (6,)->() and (32,6)->(32,); loss(xi, Exp(xi)) == 0 (atol 1e-8)."""
    rng = np.random.default_rng(0)
    assert o.geometric_loss(rng.normal(size=6), rand_poses(rng, 1)[0]).shape == ()
    xi = rng.normal(size=(32, 6))
    loss = o.geometric_loss(xi, o.se3_exp(xi))
    assert loss.shape == (32,)
    assert np.allclose(loss, 0.0, atol=1e-8)


@pytest.mark.parametrize("scale", [1e-6, 1e-2, 1.0, 3.0])
def test_loss_against_matrix_exponential(scale):
    rng = np.random.default_rng(1)
    pred = rng.normal(size=(24, 6)) * scale
    tgt = rand_poses(rng, 24)
    want = np.array([o.geometric_loss_matrix(pred[i], tgt[i]) for i in range(24)])
    assert np.allclose(o.geometric_loss(pred, tgt), want, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("scale", [1e-6, 1.0, 3.0])
def test_gradient_against_finite_differences(scale):
    rng = np.random.default_rng(2)
    pred = rng.normal(size=(16, 6)) * scale
    tgt = rand_poses(rng, 16)
    _, g = o.geometric_loss_and_grad(pred, tgt)
    eps = 1e-6
    fd = np.zeros_like(g)
    for k in range(6):
        d = np.zeros(6)
        d[k] = eps
        fd[:, k] = (o.geometric_loss(pred + d, tgt) - o.geometric_loss(pred - d, tgt)) / (2 * eps)
    assert np.allclose(g, fd, rtol=1e-6, atol=1e-6)


def test_quaternion_double_cover_and_inverse():
    rng = np.random.default_rng(3)
    pred, tgt = rng.normal(size=(8, 6)), rand_poses(rng, 8)
    flipped = tgt.copy()
    flipped[:, 3:] *= -1  # q and -q are the same rotation: Log uses atan(|v|/w)
    assert np.allclose(o.geometric_loss(pred, tgt), o.geometric_loss(pred, flipped), rtol=1e-12)
    ident = o.se3_mul(tgt, o.se3_inv(tgt))
    assert np.allclose(ident[:, :3], 0, atol=1e-12) and np.allclose(np.abs(ident[:, 6]), 1, atol=1e-12)


def test_golden_vectors():
    gold = json.loads((GOLDEN / "loss_vectors.json").read_text())
    loss, grad = o.geometric_loss_and_grad(np.array(gold["pred"]), np.array(gold["target"]))
    assert np.allclose(loss, gold["loss"], rtol=1e-12, atol=1e-14)
    assert np.allclose(grad, gold["grad"], rtol=1e-10, atol=1e-12)
    assert np.allclose(o.get_pose(np.array(gold["pred"])), gold["pose"], rtol=1e-12, atol=1e-14)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("se3") / "libse3_host.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", str(ROOT / "tests" / "csrc" / "se3_host.cpp"), "-o",
                    str(out)], check=True)
    return ctypes.CDLL(str(out))


def test_kernel_math_on_host_matches_oracle(host_lib):
    """argus_b200/csrc/se3_math.cuh compiled for the host == oracle, incl. the small-angle series branches."""
    gold = json.loads((GOLDEN / "loss_vectors.json").read_text())
    pred = np.ascontiguousarray(gold["pred"], dtype=np.float64)
    tgt = np.ascontiguousarray(gold["target"], dtype=np.float64)
    n = pred.shape[0]
    loss = np.zeros(n)
    grad = np.zeros((n, 6))
    pose = np.zeros((n, 7))
    dp = ctypes.POINTER(ctypes.c_double)
    host_lib.se3_loss_and_grad_host(pred.ctypes.data_as(dp), tgt.ctypes.data_as(dp), ctypes.c_int(n),
                                    loss.ctypes.data_as(dp), grad.ctypes.data_as(dp))
    host_lib.se3_exp_host(pred.ctypes.data_as(dp), ctypes.c_int(n), pose.ctypes.data_as(dp))
    assert np.allclose(loss, gold["loss"], rtol=1e-10, atol=1e-13)
    assert np.allclose(grad, gold["grad"], rtol=1e-8, atol=1e-10)
    assert np.allclose(pose, gold["pose"], rtol=1e-12, atol=1e-14)
