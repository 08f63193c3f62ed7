// Regression head (fp32), SE(3) pose loss and the optimizer step tail.
//   head:  argus/models.py:58-64,88-90  (GELU -> Linear(2048,128) -> GELU -> Linear(128,128) -> GELU -> Linear(128,6))
//   loss:  argus/train.py:105-119       (|Log(Exp(pred) T^-1)|^2, forward + analytic gradient in one launch)
//   step:  argus/train.py:318-319       (clip_grad_norm_ + Adam)
// These operate on a few hundred KB: they are latency-bound, so they are written for few launches and exact
// fp32/fp64 arithmetic rather than for bandwidth.
#include "kernels.h"
#include "ptx.cuh"
#include "runtime.h"
#include "se3_math.cuh"

#include <algorithm>

namespace argus {

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__global__ void gelu_fwd_bf16_kernel(const bf16* __restrict__ x, float* __restrict__ y, int64_t n) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = gelu_f(__bfloat162float(x[i]));
}
void gelu_fwd_bf16(const bf16* x, float* y, int64_t n, cudaStream_t s) {
  ProfileScope prof("head", s, 0, 6.0 * n);
  launch_kernel(gelu_fwd_bf16_kernel, static_cast<int>(std::min<int64_t>((n + 255) / 256, 1184)), 256, 0, s, x, y, n);
  ARGUS_CUDA(cudaGetLastError());
}
__global__ void gelu_bwd_bf16_kernel(const float* __restrict__ dz, const bf16* __restrict__ x, bf16* __restrict__ dx,
                                     int64_t n) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dx[i] = __float2bfloat16(dz[i] * gelu_grad_f(__bfloat162float(x[i])));
}
void gelu_bwd_bf16(const float* dz, const bf16* x, bf16* dx, int64_t n, cudaStream_t s) {
  ProfileScope prof("head", s, 0, 8.0 * n);
  launch_kernel(gelu_bwd_bf16_kernel, static_cast<int>(std::min<int64_t>((n + 255) / 256, 1184)), 256, 0, s, dz, x, dx, n);
  ARGUS_CUDA(cudaGetLastError());
}

// y[b, o] = dot(x[b, :], w[o, :]) + bias[o]. A block stages kLinRows batch rows in shared memory; its warp w takes the
// outputs 8 * blockIdx.y + w, + 8 * gridDim.y, ...: one coalesced pass over the weight row serves all staged rows (the
// weights are read B / kLinRows times instead of B times). Per (b, o) the summation is lane-strided then a shuffle
// tree: a fixed order, the same bits for any grid. The outputs are spread over gridDim.y blocks so that a small batch
// (inference at batch 1: ONE block used to walk all 128 outputs of the 1024-wide layer, 37 us per layer) still fills
// the machine.
constexpr int kLinRows = 4;
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                  float* __restrict__ y, float* __restrict__ act, int B, int In, int Out) {
  pdl_prologue();
  extern __shared__ float s_x[];   // [kLinRows][In]
  const int b0 = blockIdx.x * kLinRows;
  const int rows = min(kLinRows, B - b0);
  for (int i = threadIdx.x; i < kLinRows * In; i += 256) {
    const int r = i / In;
    s_x[i] = r < rows ? x[static_cast<int64_t>(b0) * In + i] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = blockIdx.y * 8 + warp; o < Out; o += 8 * gridDim.y) {
    const float* wr = w + static_cast<int64_t>(o) * In;
    float acc[kLinRows];
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) acc[r] = 0.f;
    for (int i = lane; i < In; i += 32) {
      const float wv = __ldg(wr + i);
#pragma unroll
      for (int r = 0; r < kLinRows; ++r) acc[r] = fmaf(s_x[r * In + i], wv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      const float bias = b[o];
#pragma unroll
      for (int r = 0; r < kLinRows; ++r)
        if (r < rows) {
          const float v = acc[r] + bias;
          const int64_t idx = static_cast<int64_t>(b0 + r) * Out + o;
          y[idx] = v;
          if (act != nullptr) act[idx] = gelu_f(v);
        }
    }
  }
}
void linear_fwd(const float* x, const float* w, const float* b, float* y, float* act, int B, int In, int Out,
                cudaStream_t s) {
  ProfileScope prof("head", s, 2.0 * B * In * Out, 4.0 * (static_cast<double>(B) * In + static_cast<double>(In) * Out + static_cast<double>(B) * Out));
  ARGUS_CHECK(static_cast<size_t>(kLinRows) * In * sizeof(float) <= 48 * 1024, "linear_fwd: input width too large");
  const int gx = (B + kLinRows - 1) / kLinRows;
  const int gy = std::max(1, std::min((Out + 7) / 8, (2 * num_sms() + gx - 1) / gx));
  launch_kernel(linear_fwd_kernel, dim3(gx, gy), 256, static_cast<size_t>(kLinRows) * In * sizeof(float), s, x, w, b, y, act, B,
                In, Out);
  ARGUS_CUDA(cudaGetLastError());
}

__global__ void gelu_grad_inplace_kernel(float* dy, const float* __restrict__ pre, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dy[i] *= gelu_grad_f(pre[i]);
}
// dw[o, i] += sum_b dy[b, o] x[b, i]: thread = one input column i and four outputs o (x is read Out / 4 times instead of
// Out times; the dy loads are warp-wide broadcasts). The sum over b runs in index order: same bits for any grid.
__global__ void __launch_bounds__(128)
linear_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* dw, float* db, int B, int In,
                    int Out) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int o0 = blockIdx.y * 4;
  if (i >= In) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = 0; b < B; ++b) {
    const float xv = __ldg(x + static_cast<int64_t>(b) * In + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = (o0 + k < Out) ? __ldg(dy + static_cast<int64_t>(b) * Out + o0 + k) : 0.f;
      acc[k] = fmaf(d, xv, acc[k]);
      accb[k] += d;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (o0 + k < Out) {
      dw[static_cast<int64_t>(o0 + k) * In + i] += acc[k];
      if (i == 0) db[o0 + k] += accb[k];
    }
}
// dx[b, i] = sum_o dy[b, o] w[o, i]: thread = one input column i and four batch rows (w is read B / 4 times)
__global__ void __launch_bounds__(128)
linear_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int In,
                    int Out) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = blockIdx.y * 4;
  if (i >= In) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int o = 0; o < Out; ++o) {
    const float wv = __ldg(w + static_cast<int64_t>(o) * In + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = (b0 + k < B) ? __ldg(dy + static_cast<int64_t>(b0 + k) * Out + o) : 0.f;
      acc[k] = fmaf(d, wv, acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (b0 + k < B) dx[static_cast<int64_t>(b0 + k) * In + i] = acc[k];
}
// The backward of one linear layer in two halves, so that the caller can take the weight gradient off the critical path:
// linear_bwd_input turns the incoming gradient into the gradient of the layer output IN PLACE (GELU') and produces dx;
// linear_bwd_weights only reads that gradient and the saved input (dw += dy^T x, db += colsum dy).
void linear_bwd_input(float* dy, const float* pre, const float* w, float* dx, int B, int In, int Out, cudaStream_t s) {
  ProfileScope prof("head", s, 2.0 * B * In * Out, 4.0 * (static_cast<double>(B) * In + static_cast<double>(In) * Out + static_cast<double>(B) * Out));
  if (pre != nullptr) {
    launch_kernel(gelu_grad_inplace_kernel, (B * Out + 255) / 256, 256, 0, s, dy, pre, B * Out);
    ARGUS_CUDA(cudaGetLastError());
  }
  if (dx != nullptr) {
    launch_kernel(linear_bwd_x_kernel, dim3((In + 127) / 128, (B + 3) / 4), 128, 0, s, dy, w, dx, B, In, Out);
    ARGUS_CUDA(cudaGetLastError());
  }
}
void linear_bwd_weights(const float* dy, const float* x, float* dw, float* db, int B, int In, int Out, cudaStream_t s) {
  ProfileScope prof("head", s, 2.0 * B * In * Out, 4.0 * (static_cast<double>(B) * In + static_cast<double>(In) * Out + static_cast<double>(B) * Out));
  launch_kernel(linear_bwd_w_kernel, dim3((In + 127) / 128, (Out + 3) / 4), 128, 0, s, dy, x, dw, db, B, In, Out);
  ARGUS_CUDA(cudaGetLastError());
}
void linear_bwd(float* dy, const float* pre, const float* x, const float* w, float* dw, float* db, float* dx, int B,
                int In, int Out, cudaStream_t s) {
  ProfileScope prof("head", s, 4.0 * B * In * Out, 4.0 * (2.0 * B * In + 2.0 * In * Out + static_cast<double>(B) * Out));
  if (pre != nullptr) {
    launch_kernel(gelu_grad_inplace_kernel, (B * Out + 255) / 256, 256, 0, s, dy, pre, B * Out);
    ARGUS_CUDA(cudaGetLastError());
  }
  launch_kernel(linear_bwd_w_kernel, dim3((In + 127) / 128, (Out + 3) / 4), 128, 0, s, dy, x, dw, db, B, In, Out);
  ARGUS_CUDA(cudaGetLastError());
  if (dx != nullptr) {
    launch_kernel(linear_bwd_x_kernel, dim3((In + 127) / 128, (B + 3) / 4), 128, 0, s, dy, w, dx, B, In, Out);
    ARGUS_CUDA(cudaGetLastError());
  }
}

// ------------------------------------------------------------------------------------------------------------
// pose loss (forward + analytic backward) and pose exponential
// ------------------------------------------------------------------------------------------------------------
// One block: thread t handles samples t, t + blockDim, ...; the batch mean is reduced in a fixed order
// (deterministic) and ADDED to *loss_mean.
__global__ void __launch_bounds__(1024)
pose_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ loss,
                 float* loss_mean, float* __restrict__ grad, int B, float grad_scale) {
  pdl_prologue();
  __shared__ double red[32];
  double acc = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    double p[6], t[7], g[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) p[k] = pred[b * 6 + k];
#pragma unroll
    for (int k = 0; k < 7; ++k) t[k] = target[b * 7 + k];
    const double lv = se3::pose_loss_and_grad(p, t, g);
    acc += lv;
    if (loss != nullptr) loss[b] = static_cast<float>(lv);
    if (grad != nullptr) {
#pragma unroll
      for (int k = 0; k < 6; ++k) grad[b * 6 + k] = static_cast<float>(g[k] * grad_scale);
    }
  }
  if (loss_mean != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < (blockDim.x + 31) / 32; ++w) tot += red[w];
      *loss_mean += static_cast<float>(tot / B);
    }
  }
}
void pose_loss_fwd_bwd(const float* pred, const float* target, float* loss, float* loss_mean, float* grad, int B,
                       float grad_scale, cudaStream_t s) {
  ProfileScope prof("pose_loss", s, 0, 80.0 * B);
  if (B <= 0) return;
  const int threads = B >= 1024 ? 1024 : ((B + 31) / 32) * 32;
  launch_kernel(pose_loss_kernel, 1, threads, 0, s, pred, target, loss, loss_mean, grad, B, grad_scale);
  ARGUS_CUDA(cudaGetLastError());
}

__global__ void pose_exp_kernel(const float* __restrict__ pred, float* __restrict__ pose, int B, int wxyz) {
  pdl_prologue();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  se3::V3 t;
  se3::Quat q;
  se3::exp_se3(se3::v3(pred[b * 6 + 0], pred[b * 6 + 1], pred[b * 6 + 2]),
               se3::v3(pred[b * 6 + 3], pred[b * 6 + 4], pred[b * 6 + 5]), t, q);
  float* o = pose + b * 7;
  o[0] = static_cast<float>(t.x); o[1] = static_cast<float>(t.y); o[2] = static_cast<float>(t.z);
  if (wxyz) {
    o[3] = static_cast<float>(q.w); o[4] = static_cast<float>(q.v.x); o[5] = static_cast<float>(q.v.y); o[6] = static_cast<float>(q.v.z);
  } else {
    o[3] = static_cast<float>(q.v.x); o[4] = static_cast<float>(q.v.y); o[5] = static_cast<float>(q.v.z); o[6] = static_cast<float>(q.w);
  }
}
void pose_exp(const float* pred, float* pose, int B, int wxyz, cudaStream_t s) {
  ProfileScope prof("pose_loss", s, 0, 52.0 * B);
  if (B <= 0) return;
  launch_kernel(pose_exp_kernel, (B + 63) / 64, 64, 0, s, pred, pose, B, wxyz);
  ARGUS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// optimizer step tail: global grad norm (deterministic two-stage), clip, Adam
// ------------------------------------------------------------------------------------------------------------
constexpr int kNormBlocks = 592;  // 4 per SM

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ g, int64_t n, float* partial) {
  pdl_prologue();
  __shared__ float red[8];
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}
int grad_sqnorm_partials(const float* g, int64_t n, float* partial, cudaStream_t s) {
  ProfileScope prof("grad_norm", s, 0, 4.0 * n);
  launch_kernel(grad_sqnorm_kernel, kNormBlocks, 256, 0, s, g, n, partial);
  ARGUS_CUDA(cudaGetLastError());
  return kNormBlocks;
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, const float* __restrict__ partial, int n_partial, float gscale, float max_norm, float lr,
                 float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float* norm_out, int skip_nonfinite) {
  pdl_prologue();
  __shared__ float red[8];
  __shared__ float s_coef;
  __shared__ int s_skip;
  // every block reduces the partial sums in the same order -> identical clip coefficient everywhere
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float norm = sqrtf(tot) * gscale;
    float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
    coef = fminf(coef, 1.f);
    s_coef = coef * gscale;
    // GradScaler protocol (argus/train.py:316-320): a step whose gradients are not finite is skipped as a whole
    s_skip = (skip_nonfinite && !isfinite(norm)) ? 1 : 0;
    if (blockIdx.x == 0 && norm_out != nullptr) *norm_out = norm;
  }
  __syncthreads();
  if (s_skip) return;
  const float coef = s_coef;
  const float step_size = lr / bc1;
  // 16-byte accesses: 4 loads + 3 stores of 16 B per thread and iteration (28 B per parameter)
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= coef;
    mi = beta1 * mi + (1.f - beta1) * gi;
    vi = beta2 * vi + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
  };
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldg(g4 + i);
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    m4[i] = mm;
    v4[i] = vv;
    p4[i] = pp;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, g[i], mi, vi);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
void clip_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* partial, int n_partial,
                    float gscale, float max_norm, float lr, float beta1, float beta2, float eps, int step,
                    float* norm_out, bool skip_nonfinite, cudaStream_t s) {
  ProfileScope prof("clip_adam", s, 0, 28.0 * n);
  ARGUS_CHECK((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) % 16 == 0, "optimizer arenas must be 16-byte aligned");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  launch_kernel(clip_adam_kernel, 8 * num_sms(), 256, 0, s, p, g, m, v, n, partial, n_partial, gscale, max_norm, lr, beta1,
                beta2, eps, bc1, bc2_sqrt, norm_out, skip_nonfinite ? 1 : 0);
  ARGUS_CUDA(cudaGetLastError());
}

}  // namespace argus
