// SE(3) / se(3) arithmetic of the pose loss, evaluated in double precision (one thread per sample; the whole
// batch is a few KB, so this is latency- not throughput-bound). Host+device so the CPU test-suite can exercise the
// exact code the kernel runs.
//
// Conventions (pypose, as used by /root/reference/argus/train.py:119 and utils.py:189):
//   se3 = [tau(3), phi(3)], SE3 = [t(3), qx, qy, qz, qw]; Exp/Log/Inv/Mul as restated in oracle/se3_loss.py.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ARGUS_HD __host__ __device__ __forceinline__
#else
#define ARGUS_HD inline
#endif

namespace argus {
namespace se3 {

struct V3 {
  double x, y, z;
};
struct M3 {
  double m[3][3];
};

ARGUS_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
ARGUS_HD V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
ARGUS_HD V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
ARGUS_HD V3 scl(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
ARGUS_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
ARGUS_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
ARGUS_HD double norm(V3 a) { return sqrt(dot(a, a)); }

ARGUS_HD M3 hat(V3 v) {
  M3 r;
  r.m[0][0] = 0; r.m[0][1] = -v.z; r.m[0][2] = v.y;
  r.m[1][0] = v.z; r.m[1][1] = 0; r.m[1][2] = -v.x;
  r.m[2][0] = -v.y; r.m[2][1] = v.x; r.m[2][2] = 0;
  return r;
}
ARGUS_HD M3 eye() {
  M3 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = (i == j) ? 1.0 : 0.0;
  return r;
}
ARGUS_HD M3 mm(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
ARGUS_HD M3 madd(const M3& a, const M3& b, double sb) {  // a + sb * b
  M3 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + sb * b.m[i][j];
  return r;
}
ARGUS_HD M3 mscale(const M3& a, double s) {
  M3 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * s;
  return r;
}
ARGUS_HD V3 mv(const M3& a, V3 v) {
  return v3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
ARGUS_HD V3 mtv(const M3& a, V3 v) {  // a^T v
  return v3(a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z, a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z,
            a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z);
}

// coefficient functions with series near 0 (thresholds chosen for double precision)
ARGUS_HD double coefA(double th) {  // sin(th/2)/th
  double t2 = th * th;
  return th < 1e-4 ? 0.5 - t2 / 48 + t2 * t2 / 3840 : sin(0.5 * th) / th;
}
ARGUS_HD double coefB(double th) {  // (1-cos th)/th^2
  double t2 = th * th;
  if (th < 1e-4) return 0.5 - t2 / 24 + t2 * t2 / 720;
  double s = sin(0.5 * th);
  return 2 * s * s / t2;
}
ARGUS_HD double coefC(double th) {  // (th-sin th)/th^3
  double t2 = th * th;
  return th < 1e-2 ? 1.0 / 6 - t2 / 120 + t2 * t2 / 5040 - t2 * t2 * t2 / 362880 : (th - sin(th)) / (t2 * th);
}
ARGUS_HD double coefD(double th) {  // 1/th^2 - cot(th/2)/(2 th)
  double t2 = th * th;
  return th < 1e-2 ? 1.0 / 12 + t2 / 720 + t2 * t2 / 30240 + t2 * t2 * t2 / 1209600
                   : 1.0 / t2 - cos(0.5 * th) / (2 * th * sin(0.5 * th));
}
ARGUS_HD double coefE(double th) {  // (th^2 + 2 cos th - 2)/(2 th^4)
  double t2 = th * th;
  return th < 5e-2 ? 1.0 / 24 - t2 / 720 + t2 * t2 / 40320 - t2 * t2 * t2 / 3628800
                   : (t2 + 2 * cos(th) - 2) / (2 * t2 * t2);
}
ARGUS_HD double coefF(double th) {  // (2 th - 3 sin th + th cos th)/(2 th^5)
  double t2 = th * th;
  return th < 1e-1 ? 1.0 / 120 - t2 / 2520 + t2 * t2 / 120960 - t2 * t2 * t2 / 9979200
                   : (2 * th - 3 * sin(th) + th * cos(th)) / (2 * t2 * t2 * th);
}

ARGUS_HD M3 Jl(V3 phi) {
  double th = norm(phi);
  M3 K = hat(phi);
  return madd(madd(eye(), K, coefB(th)), mm(K, K), coefC(th));
}
ARGUS_HD M3 JlInv(V3 phi) {
  double th = norm(phi);
  M3 K = hat(phi);
  return madd(madd(eye(), K, -0.5), mm(K, K), coefD(th));
}
// Barfoot's Q(rho, phi): upper-right block of the SE(3) left Jacobian for [rho, phi]
ARGUS_HD M3 Qmat(V3 rho, V3 phi) {
  double th = norm(phi);
  M3 P = hat(phi), R = hat(rho);
  M3 PR = mm(P, R), RP = mm(R, P), PRP = mm(PR, P), PP = mm(P, P);
  M3 t1 = madd(madd(PR, RP, 1.0), PRP, 1.0);
  M3 t2 = madd(madd(mm(PP, R), mm(R, PP), 1.0), PRP, -3.0);
  M3 t3 = madd(mm(PRP, P), mm(P, PRP), 1.0);
  M3 q = mscale(R, 0.5);
  q = madd(q, t1, coefC(th));
  q = madd(q, t2, coefE(th));
  q = madd(q, t3, coefF(th));
  return q;
}

struct Quat {
  V3 v;
  double w;
};
ARGUS_HD Quat qmul(Quat a, Quat b) {
  Quat r;
  r.v = add(add(scl(b.v, a.w), scl(a.v, b.w)), cross(a.v, b.v));
  r.w = a.w * b.w - dot(a.v, b.v);
  return r;
}
ARGUS_HD V3 qrot(Quat q, V3 x) {
  V3 t = scl(cross(q.v, x), 2.0);
  return add(add(x, scl(t, q.w)), cross(q.v, t));
}

// Exp: se3 [tau, phi] -> (t, q)
ARGUS_HD void exp_se3(V3 tau, V3 phi, V3& t, Quat& q) {
  double th = norm(phi);
  q.v = scl(phi, coefA(th));
  q.w = cos(0.5 * th);
  t = mv(Jl(phi), tau);
}

// loss = |Log(Exp(pred) * T^-1)|^2 and d loss / d pred (see oracle/se3_loss.py::geometric_loss_and_grad)
ARGUS_HD double pose_loss_and_grad(const double pred[6], const double target[7], double grad[6]) {
  V3 tau = v3(pred[0], pred[1], pred[2]), phi = v3(pred[3], pred[4], pred[5]);
  V3 tp;
  Quat qp;
  exp_se3(tau, phi, tp, qp);
  // T^-1 = (-R(conj q) t, conj q)
  Quat qc;
  qc.v = v3(-target[3], -target[4], -target[5]);
  qc.w = target[6];
  V3 tinv = scl(qrot(qc, v3(target[0], target[1], target[2])), -1.0);
  // E = Exp(pred) * T^-1
  Quat qe = qmul(qp, qc);
  V3 te = add(tp, qrot(qp, tinv));
  // Log(E)
  double n = norm(qe.v);
  double factor = (n < 1e-4) ? 2.0 / qe.w - (2.0 / 3.0) * n * n / (qe.w * qe.w * qe.w) : 2.0 * atan(n / qe.w) / n;
  V3 phi_e = scl(qe.v, factor);
  M3 Ji = JlInv(phi_e);
  V3 tau_e = mv(Ji, te);
  double loss = dot(tau_e, tau_e) + dot(phi_e, phi_e);
  // gradient: 2 Jl6(pred)^T Jl6(xi)^-T xi
  V3 u_tau = mtv(Ji, tau_e);
  M3 Qe = Qmat(tau_e, phi_e);
  V3 u_phi = sub(phi_e, mtv(Ji, mtv(Qe, u_tau)));
  M3 Jp = Jl(phi);
  M3 Qp = Qmat(tau, phi);
  V3 g_tau = scl(mtv(Jp, u_tau), 2.0);
  V3 g_phi = scl(add(mtv(Qp, u_tau), mtv(Jp, u_phi)), 2.0);
  grad[0] = g_tau.x; grad[1] = g_tau.y; grad[2] = g_tau.z;
  grad[3] = g_phi.x; grad[4] = g_phi.y; grad[5] = g_phi.z;
  return loss;
}

}  // namespace se3
}  // namespace argus
