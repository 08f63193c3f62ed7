// Host build of the exact SE(3) arithmetic the pose-loss kernel runs (argus_b200/csrc/se3_math.cuh is
// __host__ __device__), so the CPU test-suite can check it against the oracle without a GPU.
#include "../../argus_b200/csrc/se3_math.cuh"

extern "C" {
void se3_loss_and_grad_host(const double* pred, const double* target, int n, double* loss, double* grad) {
  for (int i = 0; i < n; ++i) loss[i] = argus::se3::pose_loss_and_grad(pred + 6 * i, target + 7 * i, grad + 6 * i);
}
void se3_exp_host(const double* pred, int n, double* pose) {
  for (int i = 0; i < n; ++i) {
    argus::se3::V3 t;
    argus::se3::Quat q;
    argus::se3::exp_se3(argus::se3::v3(pred[6 * i], pred[6 * i + 1], pred[6 * i + 2]),
                        argus::se3::v3(pred[6 * i + 3], pred[6 * i + 4], pred[6 * i + 5]), t, q);
    double* o = pose + 7 * i;
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = q.v.x; o[4] = q.v.y; o[5] = q.v.z; o[6] = q.w;
  }
}
}
