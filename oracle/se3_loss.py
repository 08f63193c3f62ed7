"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement, in numpy float64, of the reference's geometric pose loss

    geometric_loss_fn(pred, target) = sum((pp.se3(pred).Exp() @ target.Inv()).Log() ** 2, -1)
        -- /root/reference/argus/train.py:105-119

and of `get_pose` = `pp.se3(model(x)).Exp()` (/root/reference/argus/utils.py:179-189).

The arithmetic lives in the un-vendored third-party dependency **pypose** (PyPI `pypose>=0.6.7`,
/root/reference/pyproject.toml:24; no lockfile, not installed, no network). pypose's published conventions are
restated here:
  * se3 tangent vectors are ordered [tau(3), phi(3)] (translation part first);
  * SE3 elements are [t(3), qx, qy, qz, qw] (scalar-last unit quaternion);
  * Exp: q = [phi * sin(th/2)/th, cos(th/2)], t = Jl(phi) tau;
  * Log: phi = v * 2*atan(|v|/w)/|v| (so q and -q give the same answer), tau = Jl^{-1}(phi) t;
  * `X @ Y` is group composition, `Inv` the group inverse.

Pinning: the reference's own known-answer test (tests/test_train.py:32-36: loss(xi, Exp(xi)) == 0, atol 1e-8) and
its shape contract (tests/test_train.py:21-30) are checked in tests/test_oracle_loss.py. Nothing in the reference
pins non-zero loss values or gradients, so those are pinned here against an independent ground truth: the matrix
exponential / logarithm of 4x4 homogeneous transforms (scipy.linalg.expm / logm) and central finite differences.
Beyond that: **parity with pypose itself is unpinned** (library not available offline).
"""
from __future__ import annotations

import numpy as np

_SMALL = 1e-4  # series / closed-form switch for float64


def _hat(v):
    """(...,3) -> (...,3,3) skew-symmetric matrices."""
    z = np.zeros_like(v[..., 0])
    return np.stack(
        [
            np.stack([z, -v[..., 2], v[..., 1]], -1),
            np.stack([v[..., 2], z, -v[..., 0]], -1),
            np.stack([-v[..., 1], v[..., 0], z], -1),
        ],
        -2,
    )


def _coef_A(th):  # sin(th/2)/th
    small = th < _SMALL
    ths = np.where(small, 1.0, th)
    return np.where(small, 0.5 - th**2 / 48 + th**4 / 3840, np.sin(ths / 2) / ths)


def _coef_B(th):  # (1-cos th)/th^2
    small = th < _SMALL
    ths = np.where(small, 1.0, th)
    return np.where(small, 0.5 - th**2 / 24 + th**4 / 720, 2 * np.sin(ths / 2) ** 2 / ths**2)


def _coef_C(th):  # (th - sin th)/th^3
    small = th < 1e-2
    ths = np.where(small, 1.0, th)
    return np.where(small, 1 / 6 - th**2 / 120 + th**4 / 5040 - th**6 / 362880, (ths - np.sin(ths)) / ths**3)


def _coef_D(th):  # 1/th^2 - (1+cos th)/(2 th sin th) = 1/th^2 - cot(th/2)/(2 th)
    small = th < 1e-2
    ths = np.where(small, 1.0, th)
    return np.where(
        small,
        1 / 12 + th**2 / 720 + th**4 / 30240 + th**6 / 1209600,
        1 / ths**2 - np.cos(ths / 2) / (2 * ths * np.sin(ths / 2)),
    )


def _coef_E(th):  # (th^2 + 2 cos th - 2)/(2 th^4)
    small = th < 5e-2
    ths = np.where(small, 1.0, th)
    return np.where(
        small,
        1 / 24 - th**2 / 720 + th**4 / 40320 - th**6 / 3628800,
        (ths**2 + 2 * np.cos(ths) - 2) / (2 * ths**4),
    )


def _coef_F(th):  # (2 th - 3 sin th + th cos th)/(2 th^5)
    small = th < 1e-1
    ths = np.where(small, 1.0, th)
    return np.where(
        small,
        1 / 120 - th**2 / 2520 + th**4 / 120960 - th**6 / 9979200,
        (2 * ths - 3 * np.sin(ths) + ths * np.cos(ths)) / (2 * ths**5),
    )


def quat_mul(a, b):
    """Hamilton product of scalar-last quaternions."""
    av, aw = a[..., :3], a[..., 3:4]
    bv, bw = b[..., :3], b[..., 3:4]
    v = aw * bv + bw * av + np.cross(av, bv)
    w = aw * bw - np.sum(av * bv, -1, keepdims=True)
    return np.concatenate([v, w], -1)


def quat_rotate(q, x):
    """R(q) x for scalar-last unit quaternions."""
    v, w = q[..., :3], q[..., 3:4]
    t = 2 * np.cross(v, x)
    return x + w * t + np.cross(v, t)


def left_jacobian_so3(phi):
    th = np.linalg.norm(phi, axis=-1)[..., None, None]
    K = _hat(phi)
    I = np.broadcast_to(np.eye(3), K.shape)
    return I + _coef_B(th) * K + _coef_C(th) * (K @ K)


def left_jacobian_so3_inv(phi):
    th = np.linalg.norm(phi, axis=-1)[..., None, None]
    K = _hat(phi)
    I = np.broadcast_to(np.eye(3), K.shape)
    return I - 0.5 * K + _coef_D(th) * (K @ K)


def se3_exp(xi):
    """se3 [tau, phi] -> SE3 [t, qx, qy, qz, qw]   (pp.se3(xi).Exp())."""
    xi = np.asarray(xi, dtype=np.float64)
    tau, phi = xi[..., :3], xi[..., 3:]
    th = np.linalg.norm(phi, axis=-1, keepdims=True)
    q = np.concatenate([phi * _coef_A(th), np.cos(th / 2)], -1)
    t = (left_jacobian_so3(phi) @ tau[..., None])[..., 0]
    return np.concatenate([t, q], -1)


def se3_inv(T):
    T = np.asarray(T, dtype=np.float64)
    t, q = T[..., :3], T[..., 3:]
    qc = np.concatenate([-q[..., :3], q[..., 3:]], -1)
    return np.concatenate([-quat_rotate(qc, t), qc], -1)


def se3_mul(X, Y):
    tx, qx = X[..., :3], X[..., 3:]
    ty, qy = Y[..., :3], Y[..., 3:]
    return np.concatenate([tx + quat_rotate(qx, ty), quat_mul(qx, qy)], -1)


def se3_log(T):
    """SE3 [t, q] -> se3 [tau, phi]   (LieTensor.Log())."""
    T = np.asarray(T, dtype=np.float64)
    t, q = T[..., :3], T[..., 3:]
    v, w = q[..., :3], q[..., 3:4]
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    small = n < _SMALL
    ns = np.where(small, 1.0, n)
    with np.errstate(divide="ignore"):
        factor = np.where(small, 2 / w - (2 / 3) * n**2 / w**3, 2 * np.arctan(ns / w) / ns)
    phi = v * factor
    tau = (left_jacobian_so3_inv(phi) @ t[..., None])[..., 0]
    return np.concatenate([tau, phi], -1)


def geometric_loss(pred, target):
    """Per-sample loss, float64. pred (...,6) se3; target (...,7) SE3 [t, q_xyzw]."""
    pred = np.asarray(pred, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    xi = se3_log(se3_mul(se3_exp(pred), se3_inv(target)))
    return np.sum(xi**2, -1)


def _Q_matrix(rho, phi):
    """Barfoot's Q(rho, phi): the off-diagonal block of the SE(3) left Jacobian."""
    th = np.linalg.norm(phi, axis=-1)[..., None, None]
    P, R = _hat(phi), _hat(rho)
    c1, c2, c3 = _coef_C(th), _coef_E(th), _coef_F(th)
    return (
        0.5 * R
        + c1 * (P @ R + R @ P + P @ R @ P)
        + c2 * (P @ P @ R + R @ P @ P - 3 * P @ R @ P)
        + c3 * (P @ R @ P @ P + P @ P @ R @ P)
    )


def geometric_loss_and_grad(pred, target):
    """Loss and its analytic gradient with respect to pred:

        xi = Log(Exp(pred) T^-1),  L = |xi|^2,  dL/dpred = 2 Jl6(pred)^T Jl6(xi)^-T xi
        Jl6([tau, phi]) = [[Jl(phi), Q(tau, phi)], [0, Jl(phi)]]
    """
    pred = np.asarray(pred, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    xi = se3_log(se3_mul(se3_exp(pred), se3_inv(target)))
    loss = np.sum(xi**2, -1)
    tau_e, phi_e = xi[..., :3, None], xi[..., 3:, None]
    Jinv = left_jacobian_so3_inv(xi[..., 3:])
    JinvT = np.swapaxes(Jinv, -1, -2)
    Qe = _Q_matrix(xi[..., :3], xi[..., 3:])
    u_tau = JinvT @ tau_e
    u_phi = phi_e - JinvT @ (np.swapaxes(Qe, -1, -2) @ u_tau)
    Jp = left_jacobian_so3(pred[..., 3:])
    JpT = np.swapaxes(Jp, -1, -2)
    Qp = _Q_matrix(pred[..., :3], pred[..., 3:])
    g_tau = 2 * (JpT @ u_tau)
    g_phi = 2 * (np.swapaxes(Qp, -1, -2) @ u_tau + JpT @ u_phi)
    return loss, np.concatenate([g_tau[..., 0], g_phi[..., 0]], -1)


def get_pose(pred):
    """`pp.se3(pred).Exp()` -> (...,7) [t, qx, qy, qz, qw]."""
    return se3_exp(pred)


# ------------------------------------------------------------------------------------------------------------------
# Independent ground truth used to pin the restatement (4x4 homogeneous matrices + scipy expm/logm)
# ------------------------------------------------------------------------------------------------------------------
def _twist_matrix(xi):
    X = np.zeros((4, 4))
    X[:3, :3] = _hat(np.asarray(xi[3:], dtype=np.float64))
    X[:3, 3] = xi[:3]
    return X


def _pose_matrix(T):
    from scipy.spatial.transform import Rotation

    M = np.eye(4)
    M[:3, :3] = Rotation.from_quat(T[3:]).as_matrix()
    M[:3, 3] = T[:3]
    return M


def geometric_loss_matrix(pred, target):
    """Same loss through expm/logm of 4x4 matrices, for one sample. Valid away from rotation angle pi."""
    from scipy.linalg import expm, logm

    E = expm(_twist_matrix(pred)) @ np.linalg.inv(_pose_matrix(target))
    Lg = np.real(logm(E))
    xi = np.array([Lg[0, 3], Lg[1, 3], Lg[2, 3], Lg[2, 1], Lg[0, 2], Lg[1, 0]])
    return float(np.sum(xi**2))
