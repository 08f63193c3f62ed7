// Memory-bound kernels of the hot path: input/weight packing, batch-norm (finalize / apply / backward), pooling.
// Every kernel moves 16 bytes per thread per access (8 bf16 channels of one pixel), is coalesced along the
// channel-contiguous NHWC layout, and reduces with warp shuffles / shared memory before touching global atomics.
// Reference semantics: torch.nn.BatchNorm2d, ReLU, MaxPool2d(3,2,1), AdaptiveAvgPool2d(1) inside torchvision
// resnet50 as called from /root/reference/argus/models.py:84.
#include "kernels.h"
#include "ptx.cuh"
#include "runtime.h"

#include <algorithm>
#include <string>

namespace argus {

static inline int grid_for(int64_t work_items, int threads, int max_blocks_per_sm = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(num_sms()) * max_blocks_per_sm;
  return static_cast<int>(std::max<int64_t>(1, std::min(blocks, cap)));
}

struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 unpack8(const uint4& u) {
  F8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
  return r;
}
__device__ __forceinline__ uint4 pack8(const F8& f) {
  uint4 u;
  u.x = pack_bf16x2(f.v[0], f.v[1]);
  u.y = pack_bf16x2(f.v[2], f.v[3]);
  u.z = pack_bf16x2(f.v[4], f.v[5]);
  u.w = pack_bf16x2(f.v[6], f.v[7]);
  return u;
}
// bit k = (v[k] > 0): the ReLU mask of eight channels in one byte (1/16 of the bf16 activation bytes)
__device__ __forceinline__ uint8_t positive_bits(const F8& f) {
  uint32_t b = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) b |= (f.v[k] > 0.f ? 1u : 0u) << k;
  return static_cast<uint8_t>(b);
}
__device__ __forceinline__ F8 load8f(const float* p) {
  F8 r;
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ------------------------------------------------------------------------------------------------------------
// input packing: NCHW fp32 (or HWC u8) -> space-to-depth bf16 [n][H/2][W/2+4][16]
// ------------------------------------------------------------------------------------------------------------
__global__ void pack_input_f32_kernel(const float* __restrict__ x, uint4* __restrict__ out, int n_images, int H,
                                      int W) {
  pdl_prologue();
  const int Hs = H >> 1, Ws = W >> 1, Wp = Ws + 4;
  const int64_t total = static_cast<int64_t>(n_images) * Hs * Wp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int jp = static_cast<int>(i % Wp);
    const int64_t t = i / Wp;
    const int is = static_cast<int>(t % Hs);
    const int n = static_cast<int>(t / Hs);
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.f;
    const int j = jp - 2;
    if (j >= 0 && j < Ws) {
      const float* img = x + static_cast<int64_t>(n) * 3 * H * W;  // image n = channels [3n, 3n+3) of sample n / n_cams
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const float2 p = __ldg(reinterpret_cast<const float2*>(img + (static_cast<int64_t>(c) * H + 2 * is + a) * W + 2 * j));
          v[(a * 2 + 0) * 3 + c] = p.x;
          v[(a * 2 + 1) * 3 + c] = p.y;
        }
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]); o0.y = pack_bf16x2(v[2], v[3]); o0.z = pack_bf16x2(v[4], v[5]); o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]); o1.y = pack_bf16x2(v[10], v[11]); o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
    out[2 * i] = o0;
    out[2 * i + 1] = o1;
  }
}

__global__ void pack_input_u8_kernel(const uint8_t* __restrict__ x, uint4* __restrict__ out, int n_images, int H,
                                     int W) {
  pdl_prologue();
  const int Hs = H >> 1, Ws = W >> 1, Wp = Ws + 4;
  const int64_t total = static_cast<int64_t>(n_images) * Hs * Wp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int jp = static_cast<int>(i % Wp);
    const int64_t t = i / Wp;
    const int is = static_cast<int>(t % Hs);
    const int n = static_cast<int>(t / Hs);
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.f;
    const int j = jp - 2;
    if (j >= 0 && j < Ws) {
      const uint8_t* img = x + static_cast<int64_t>(n) * H * W * 3;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const uint8_t* px = img + (static_cast<int64_t>(2 * is + a) * W + 2 * j) * 3;  // 6 contiguous bytes, 2-byte aligned
        const uint16_t* p16 = reinterpret_cast<const uint16_t*>(px);
        const uint32_t w0 = p16[0], w1 = p16[1], w2 = p16[2];
        const float r0 = (w0 & 0xff), g0 = (w0 >> 8), b0 = (w1 & 0xff), r1 = (w1 >> 8), g1 = (w2 & 0xff), b1 = (w2 >> 8);
        const float k = 1.0f / 255.0f;
        v[(a * 2 + 0) * 3 + 0] = r0 * k; v[(a * 2 + 0) * 3 + 1] = g0 * k; v[(a * 2 + 0) * 3 + 2] = b0 * k;
        v[(a * 2 + 1) * 3 + 0] = r1 * k; v[(a * 2 + 1) * 3 + 1] = g1 * k; v[(a * 2 + 1) * 3 + 2] = b1 * k;
      }
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]); o0.y = pack_bf16x2(v[2], v[3]); o0.z = pack_bf16x2(v[4], v[5]); o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]); o1.y = pack_bf16x2(v[10], v[11]); o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
    out[2 * i] = o0;
    out[2 * i + 1] = o1;
  }
}

void pack_input_f32(const float* x, bf16* out, int n_images, int H, int W, cudaStream_t s) {
  ProfileScope prof("pack_input", s, 0, static_cast<double>(n_images) * H * W * 3 * 4 + static_cast<double>(n_images) * (H / 2) * (W / 2 + 4) * 32);
  const int64_t total = static_cast<int64_t>(n_images) * (H / 2) * (W / 2 + 4);
  launch_kernel(pack_input_f32_kernel, grid_for(total, 256), 256, 0, s, x, reinterpret_cast<uint4*>(out), n_images, H, W);
  ARGUS_CUDA(cudaGetLastError());
}
void pack_input_u8(const uint8_t* x, bf16* out, int n_images, int H, int W, cudaStream_t s) {
  ProfileScope prof("pack_input", s, 0, static_cast<double>(n_images) * H * W * 3 + static_cast<double>(n_images) * (H / 2) * (W / 2 + 4) * 32);
  const int64_t total = static_cast<int64_t>(n_images) * (H / 2) * (W / 2 + 4);
  launch_kernel(pack_input_u8_kernel, grid_for(total, 256), 256, 0, s, x, reinterpret_cast<uint4*>(out), n_images, H, W);
  ARGUS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// weight packing / gradient unpacking (table driven: one launch for all layers)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t packed_count(const WeightPackEntry& e) {
  return e.kind == 1 ? 64 * 256 : static_cast<int64_t>(e.cout) * e.cin * e.kk;
}
// maps a packed index to the PyTorch-layout index (or -1 for a zero-padding slot of the stem)
__device__ __forceinline__ int64_t packed_to_torch(const WeightPackEntry& e, int64_t i) {
  if (e.kind == 1) {
    const int ch = static_cast<int>(i & 15);
    const int q = static_cast<int>((i >> 4) & 3);
    const int p = static_cast<int>((i >> 6) & 3);
    const int co = static_cast<int>(i >> 8);
    if (ch >= 12) return -1;
    const int ab = ch / 3, c = ch - ab * 3;
    const int a = ab >> 1, b = ab & 1;
    const int kh = 2 * p + a - 1, kw = 2 * q + b - 1;
    if (kh < 0 || kh > 6 || kw < 0 || kw > 6) return -1;
    return ((static_cast<int64_t>(co) * 3 + c) * 7 + kh) * 7 + kw;
  }
  const int ci = static_cast<int>(i % e.cin);
  const int64_t r = i / e.cin;
  const int t = static_cast<int>(r % e.kk);
  const int co = static_cast<int>(r / e.kk);
  return (static_cast<int64_t>(co) * e.cin + ci) * e.kk + t;
}

// k x k entries are transposed one output channel at a time through shared memory ([ci][t] <-> [t][ci] inside the
// contiguous cin * kk block of that channel), so both the global reads and the global writes are coalesced; 1x1 / linear
// entries are already in the packed order; the stem (kind 1) keeps the per-element index map (it is 16 K elements).
constexpr int kPackTileMax = 512 * 9 + 9;   // largest cin * kk block (+ padding for the unpack direction)
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ params, bf16* __restrict__ packed,
                    const WeightPackEntry* __restrict__ table) {
  pdl_prologue();
  __shared__ float tile[kPackTileMax];
  const WeightPackEntry e = table[blockIdx.y];
  const int64_t n = packed_count(e);
  if (e.kind == 0 && e.kk > 1 && e.cin * e.kk <= kPackTileMax) {
    const int row = e.cin * e.kk;
    for (int co = blockIdx.x; co < e.cout; co += gridDim.x) {
      const float* src = params + e.src_off + static_cast<int64_t>(co) * row;
      for (int j = threadIdx.x; j < row; j += blockDim.x) tile[j] = src[j];   // [ci][t]
      __syncthreads();
      bf16* dst = packed + e.dst_off + static_cast<int64_t>(co) * row;
      for (int j = threadIdx.x; j < row; j += blockDim.x) {                   // j = t * cin + ci
        const int t = j / e.cin, ci = j - t * e.cin;
        dst[j] = __float2bfloat16(tile[ci * e.kk + t]);
      }
      __syncthreads();
    }
    return;
  }
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t src = packed_to_torch(e, i);
    packed[e.dst_off + i] = __float2bfloat16(src >= 0 ? params[e.src_off + src] : 0.f);
  }
}
__global__ void __launch_bounds__(256)
unpack_wgrads_kernel(const float* __restrict__ packed_grads, float* __restrict__ grads,
                     const WeightPackEntry* __restrict__ table) {
  pdl_prologue();
  __shared__ float tile[kPackTileMax];
  const WeightPackEntry e = table[blockIdx.y];
  if (e.kind == 0 && e.kk == 1) return;  // 1x1 / linear layers accumulate straight into the gradient arena
  const int64_t n = packed_count(e);
  if (e.kind == 0 && (e.cin + 1) * e.kk <= kPackTileMax) {
    const int row = e.cin * e.kk, pitch = e.cin + 1;   // padded rows: the transposed reads spread over the banks
    for (int co = blockIdx.x; co < e.cout; co += gridDim.x) {
      const float* src = packed_grads + e.dst_off + static_cast<int64_t>(co) * row;
      for (int j = threadIdx.x; j < row; j += blockDim.x) {                   // j = t * cin + ci
        const int t = j / e.cin, ci = j - t * e.cin;
        tile[t * pitch + ci] = src[j];
      }
      __syncthreads();
      float* dst = grads + e.src_off + static_cast<int64_t>(co) * row;
      for (int j = threadIdx.x; j < row; j += blockDim.x) {                   // j = ci * kk + t
        const int ci = j / e.kk, t = j - ci * e.kk;
        dst[j] += tile[t * pitch + ci];
      }
      __syncthreads();
    }
    return;
  }
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t dst = packed_to_torch(e, i);
    if (dst >= 0) grads[e.src_off + dst] += packed_grads[e.dst_off + i];
  }
}
void pack_weights(const float* params, bf16* packed, const WeightPackEntry* table_dev, int n_entries,
                  cudaStream_t s) {
  ProfileScope prof("pack_weights", s, 0, 0);
  // 256 blocks per entry: a 512-channel 3x3 layer is two output channels per block (one load / transpose / store round each)
  launch_kernel(pack_weights_kernel, dim3(256, n_entries), 256, 0, s, params, packed, table_dev);
  ARGUS_CUDA(cudaGetLastError());
}
void unpack_wgrads(const float* packed_grads, float* grads, const WeightPackEntry* table_dev, int n_entries,
                   cudaStream_t s) {
  ProfileScope prof("unpack_wgrads", s, 0, 0);
  launch_kernel(unpack_wgrads_kernel, dim3(256, n_entries), 256, 0, s, packed_grads, grads, table_dev);
  ARGUS_CUDA(cudaGetLastError());
}

// plain 4-byte global load as a volatile asm statement: a run of these is issued back to back (the compiler keeps
// their order and cannot fold the consumers in between), so N partial sums cost one memory round trip, not N
__device__ __forceinline__ float ldg_f32_issue(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// batch norm: finalize / fold
// ------------------------------------------------------------------------------------------------------------
// block = 8 channels x 32 slot lanes: lane sl adds slots sl, sl+32, ... in order, then lane 0 adds the 32 lane sums in
// order -> the statistics are bitwise reproducible and the slot loop is 32-way parallel.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ partial, int slots, double count, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                   float* save_mean, float* save_invstd, int C) {
  pdl_prologue();
  __shared__ double red[2][32][8];
  const int ch = threadIdx.x & 7, sl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + ch;
  double sum = 0.0, sqsum = 0.0;
  if (c < C) {
    // eight slots per batch: the loads are issued together (one L2 round trip instead of eight), the additions keep
    // the slot order (missing slots are skipped) -> same bits as the plain loop
    for (int k = sl; k < slots; k += 32 * 8) {
      float a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int kk = min(k + 32 * u, slots - 1);   // clamped: always a valid address, masked below
        a[u] = ldg_f32_issue(partial + (static_cast<size_t>(kk) * 2 + 0) * C + c);
        b[u] = ldg_f32_issue(partial + (static_cast<size_t>(kk) * 2 + 1) * C + c);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (k + 32 * u < slots) {
          sum += static_cast<double>(a[u]);
          sqsum += static_cast<double>(b[u]);
        }
      }
    }
  }
  red[0][sl][ch] = sum;
  red[1][sl][ch] = sqsum;
  __syncthreads();
  if (sl != 0 || c >= C) return;
  sum = 0.0;
  sqsum = 0.0;
  for (int k = 0; k < 32; ++k) { sum += red[0][k][ch]; sqsum += red[1][k][ch]; }
  const double mean = sum / count;
  double var = sqsum / count - mean * mean;
  if (var < 0) var = 0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mean) * sc;
  save_mean[c] = static_cast<float>(mean);
  save_invstd[c] = invstd;
  if (running_mean != nullptr) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}
void bn_finalize(const float* partial, int slots, double count, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                 float* save_mean, float* save_invstd, int C, cudaStream_t s) {
  ProfileScope prof("bn_finalize", s, 0, (8.0 * slots + 40.0) * C);
  launch_kernel(bn_finalize_kernel, (C + 7) / 8, 256, 0, s, partial, slots, count, gamma, beta, running_mean, running_var,
                                                     momentum, eps, scale, shift, save_mean, save_invstd, C);
  ARGUS_CUDA(cudaGetLastError());
}
__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                    float* scale, float* shift, int C) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}
void bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                  float eps, float* scale, float* shift, int C, cudaStream_t s) {
  ProfileScope prof("bn_finalize", s, 0, 24.0 * C);
  launch_kernel(bn_fold_eval_kernel, (C + 127) / 128, 128, 0, s, gamma, beta, running_mean, running_var, eps, scale, shift, C);
  ARGUS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// batch norm apply (+ residual, + ReLU)
// ------------------------------------------------------------------------------------------------------------
template <int RES, bool SUM>  // RES: 0 none, 1 plain residual, 2 residual with its own scale/shift (downsample branch)
__global__ void __launch_bounds__(256)
bn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                const uint4* __restrict__ res, const float* __restrict__ rscale, const float* __restrict__ rshift,
                int relu, uint4* __restrict__ y, uint8_t* __restrict__ bits, float* __restrict__ colsum_partial,
                int64_t nvec, int cvec) {
  pdl_prologue();
  // gridDim.x * 256 is a multiple of cvec (a power of two <= 256), so a thread always sees the same channel octet:
  // the per-channel constants live in registers for the whole grid-stride loop.
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int c0 = static_cast<int>(tid & (cvec - 1)) * 8;
  const F8 sc = load8f(scale + c0), sh = load8f(shift + c0);
  F8 rs, rb;
  if (RES == 2) { rs = load8f(rscale + c0); rb = load8f(rshift + c0); }
  F8 acc;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc.v[k] = 0.f;
  auto one = [&](int64_t i, const uint4& xq, const uint4& rq) {
    const F8 xv = unpack8(xq);
    F8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = fmaf(xv.v[k], sc.v[k], sh.v[k]);
    if (RES == 1) {
      const F8 rv = unpack8(rq);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] += rv.v[k];
    } else if (RES == 2) {
      const F8 rv = unpack8(rq);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] += fmaf(rv.v[k], rs.v[k], rb.v[k]);
    }
    if (bits != nullptr) bits[i] = positive_bits(o);
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = fmaxf(o.v[k], 0.f);
    }
    const uint4 packed = pack8(o);
    y[i] = packed;
    if (SUM) {
      const F8 r = unpack8(packed);   // sums of the STORED (bf16) values
#pragma unroll
      for (int k = 0; k < 8; ++k) acc.v[k] += r.v[k];
    }
  };
  int64_t i = tid;
  for (; i + stride < nvec; i += 2 * stride) {
    const uint4 xa = ldg_stream(x + i), xb = ldg_stream(x + i + stride);
    uint4 ra = make_uint4(0, 0, 0, 0), rbv = ra;
    if (RES != 0) { ra = ldg_stream(res + i); rbv = ldg_stream(res + i + stride); }
    one(i, xa, ra);
    one(i + stride, xb, rbv);
  }
  for (; i < nvec; i += stride) {
    uint4 ra = make_uint4(0, 0, 0, 0);
    if (RES != 0) ra = ldg_stream(res + i);
    one(i, ldg_stream(x + i), ra);
  }
  if (SUM) {
    // per-block column sums of the output: threads t, t + cvec, ... of the block share a channel octet
    __shared__ float red[8][256];
#pragma unroll
    for (int k = 0; k < 8; ++k) red[k][threadIdx.x] = acc.v[k];
    __syncthreads();
    const int lanes = cvec < 256 ? cvec : 256;
    if (threadIdx.x < lanes) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float s0 = 0.f;
        for (int r = threadIdx.x; r < 256; r += lanes) s0 += red[k][r];
        colsum_partial[static_cast<size_t>(blockIdx.x) * (cvec * 8) + threadIdx.x * 8 + k] = s0;
      }
    }
  }
}
// shared-memory ring version (defined below, next to the ring versions of the backward passes)
static bool bn_ring_enabled();
static int bn_ring_grid(int64_t nvec);
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static void launch_bn_apply_ring(const uint4* x, const float* scale, const float* shift, const uint4* res,
                                 const float* rscale, const float* rshift, int relu, uint4* y, uint8_t* bits,
                                 float* colsum_partial, int64_t nvec, int cvec, cudaStream_t s);
// number of per-block column-sum partials bn_apply writes (callers size and finalize the buffer with this)
int bn_apply_grid(int64_t rows, int C) {
  return bn_ring_enabled() ? bn_ring_grid(rows * (C / 8)) : grid_for(rows * (C / 8), 256);
}
void bn_apply(const bf16* x, const float* scale, const float* shift, const bf16* res, const float* rscale,
              const float* rshift, int relu, bf16* y, uint8_t* relu_bits, float* colsum_partial, int64_t rows, int C,
              cudaStream_t s) {
  ProfileScope prof("bn_apply", s, 0, static_cast<double>(rows) * C * (2 * (res ? 3 : 2) + (relu_bits ? 0.125 : 0.0)));
  ARGUS_CHECK(C % 8 == 0 && is_pow2(C / 8) && C <= 2048, "bn_apply: C/8 must be a power of two <= 256");
  const int64_t nvec = rows * (C / 8);
  const int grid = bn_apply_grid(rows, C);
  auto X = reinterpret_cast<const uint4*>(x);
  auto R = reinterpret_cast<const uint4*>(res);
  auto Y = reinterpret_cast<uint4*>(y);
  if (bn_ring_enabled()) {
    ARGUS_CHECK(aligned16(x) && aligned16(res) && aligned16(y), "bn_apply: tensors must be 16-byte aligned");
    ARGUS_CHECK(colsum_partial == nullptr || res == nullptr, "column sums are only produced by the plain (no residual) variant");
    launch_bn_apply_ring(X, scale, shift, R, rscale, rshift, relu, Y, relu_bits, colsum_partial, nvec, C / 8, s);
    ARGUS_CUDA(cudaGetLastError());
    return;
  }
  if (colsum_partial != nullptr) {
    ARGUS_CHECK(res == nullptr, "column sums are only produced by the plain (no residual) variant");
    launch_kernel(bn_apply_kernel<0, true>, grid, 256, 0, s, X, scale, shift, R, rscale, rshift, relu, Y, relu_bits, colsum_partial, nvec, C / 8);
  } else if (res == nullptr) {
    launch_kernel(bn_apply_kernel<0, false>, grid, 256, 0, s, X, scale, shift, R, rscale, rshift, relu, Y, relu_bits, nullptr, nvec, C / 8);
  } else if (rscale == nullptr) {
    launch_kernel(bn_apply_kernel<1, false>, grid, 256, 0, s, X, scale, shift, R, rscale, rshift, relu, Y, relu_bits, nullptr, nvec, C / 8);
  } else {
    launch_kernel(bn_apply_kernel<2, false>, grid, 256, 0, s, X, scale, shift, R, rscale, rshift, relu, Y, relu_bits, nullptr, nvec, C / 8);
  }
  ARGUS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// batch norm backward: per-channel reductions, then the elementwise input gradient
// ------------------------------------------------------------------------------------------------------------
template <int MASK>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ out,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ partial,
                     int64_t rows, int cvec) {
  pdl_prologue();
  __shared__ float red[16][256];
  const int lanes = cvec < 256 ? cvec : 256;   // threads along the channel dimension (cvec <= 256 for C <= 2048)
  const int row_lanes = 256 / lanes;           // rows processed concurrently by one block
  const int rl = threadIdx.x / lanes;
  const int oc = threadIdx.x % lanes;
  const int c0 = oc * 8;
  const F8 sc = load8f(scale + c0), sh = load8f(shift + c0), mu = load8f(mean + c0), is = load8f(invstd + c0);
  float a_dy[8], a_dyx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a_dy[k] = a_dyx[k] = 0.f;
  const uint8_t* bits = reinterpret_cast<const uint8_t*>(out);   // MASK == 3: one byte per channel octet
  auto body = [&](const uint4& dyv, const uint4& xv4, const uint4& ov4) {
    const F8 d = unpack8(dyv);
    const F8 xv = unpack8(xv4);
    F8 g = d;
    if (MASK == 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) g.v[k] = fmaf(xv.v[k], sc.v[k], sh.v[k]) > 0.f ? d.v[k] : 0.f;
    } else if (MASK == 2) {
      const F8 o = unpack8(ov4);
#pragma unroll
      for (int k = 0; k < 8; ++k) g.v[k] = o.v[k] > 0.f ? d.v[k] : 0.f;
    } else if (MASK == 3) {
#pragma unroll
      for (int k = 0; k < 8; ++k) g.v[k] = ((ov4.x >> k) & 1u) ? d.v[k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a_dy[k] += g.v[k];
      a_dyx[k] = fmaf(g.v[k], xv.v[k], a_dyx[k]);   // sum g*x; centred and scaled by invstd once at the end
    }
  };
  const int64_t rstride = static_cast<int64_t>(gridDim.x) * row_lanes;
  int64_t row = static_cast<int64_t>(blockIdx.x) * row_lanes + rl;
  for (; row + rstride < rows; row += 2 * rstride) {
    const int64_t i0 = row * cvec + oc, i1 = (row + rstride) * cvec + oc;
    const uint4 d0 = ldg_stream(dy + i0), d1 = ldg_stream(dy + i1);
    const uint4 x0 = ldg_stream(x + i0), x1 = ldg_stream(x + i1);
    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
    if (MASK == 2) { o0 = ldg_stream(out + i0); o1 = ldg_stream(out + i1); }
    if (MASK == 3) { o0.x = __ldg(bits + i0); o1.x = __ldg(bits + i1); }
    body(d0, x0, o0);
    body(d1, x1, o1);
  }
  for (; row < rows; row += rstride) {
    const int64_t i0 = row * cvec + oc;
    uint4 o0 = make_uint4(0, 0, 0, 0);
    if (MASK == 2) o0 = ldg_stream(out + i0);
    if (MASK == 3) o0.x = __ldg(bits + i0);
    body(ldg_stream(dy + i0), ldg_stream(x + i0), o0);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[k][threadIdx.x] = a_dy[k];
    red[8 + k][threadIdx.x] = (a_dyx[k] - mu.v[k] * a_dy[k]) * is.v[k];
  }
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s0 = 0.f, s1 = 0.f;
      for (int r = 0; r < row_lanes; ++r) {
        s0 += red[k][r * lanes + oc];
        s1 += red[8 + k][r * lanes + oc];
      }
      // partial[block][0][c] = sum g, partial[block][1][c] = sum g * xhat
      partial[(static_cast<size_t>(blockIdx.x) * 2 + 0) * (cvec * 8) + c0 + k] = s0;
      partial[(static_cast<size_t>(blockIdx.x) * 2 + 1) * (cvec * 8) + c0 + k] = s1;
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int blocks, float* dgamma, float* dbeta, int C,
                       const float* __restrict__ mean, const float* __restrict__ invstd) {
  pdl_prologue();
  __shared__ double red[2][32][8];
  const int ch = threadIdx.x & 7, sl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + ch;
  double sb = 0.0, sg = 0.0;
  if (c < C) {
    for (int k = sl; k < blocks; k += 32 * 8) {   // batched loads, ordered additions (see bn_finalize_kernel)
      float a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int kk = min(k + 32 * u, blocks - 1);
        a[u] = ldg_f32_issue(partial + (static_cast<size_t>(kk) * 2 + 0) * C + c);
        b[u] = ldg_f32_issue(partial + (static_cast<size_t>(kk) * 2 + 1) * C + c);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (k + 32 * u < blocks) {
          sb += static_cast<double>(a[u]);
          sg += static_cast<double>(b[u]);
        }
      }
    }
  }
  red[0][sl][ch] = sb;
  red[1][sl][ch] = sg;
  __syncthreads();
  if (sl != 0 || c >= C) return;
  sb = 0.0;
  sg = 0.0;
  for (int k = 0; k < 32; ++k) { sb += red[0][k][ch]; sg += red[1][k][ch]; }
  // mean != nullptr: the second sums are sum(g * x) of the RAW layer output (statistics slots of a dgrad epilogue,
  // bn_bwd_finalize_slots); centred and scaled here, once, in double
  if (mean != nullptr) sg = (sg - static_cast<double>(mean[c]) * sb) * static_cast<double>(invstd[c]);
  dbeta[c] += static_cast<float>(sb);
  dgamma[c] += static_cast<float>(sg);
}

// ------------------------------------------------------------------------------------------------------------
// Shared-memory ring versions of the two batch-norm backward passes (default; ARGUS_BN_RING=0 selects the register
// versions above). A producer warp streams 8 KB slices of dy / x (/ out or the bit mask) into a ring of stages with
// 1-D bulk copies (TMA) that complete on mbarriers; eight consumer warps read their 16-byte vectors from the landed
// stage, release it, and do the math. The bytes in flight per SM are set by the ring (2 CTAs x 6 stages x 16 KB)
// instead of by the registers the loads of one loop iteration can hold (1024 threads x 64 B), which is what kept the
// register versions at 55-70 % of the HBM copy bandwidth.
// ------------------------------------------------------------------------------------------------------------
constexpr int kRingVec = 512;   // uint4 per tensor per stage: two per consumer thread
template <int MASK>
struct BnRing {
  static constexpr int kTensorBytes = kRingVec * 16;
  static constexpr int kExtraBytes = MASK == 2 ? kTensorBytes : (MASK == 3 ? kRingVec : 0);
  static constexpr int kStageBytes = 2 * kTensorBytes + kExtraBytes;
  // 2 CTAs x 5 stages x 16 KB in flight per SM is far more than HBM latency needs; stopping at ~80 KB per CTA leaves
  // room for one CTA of the (ALU-bound, 41 KB) augmentation kernel of the next batch to share the SM
  static constexpr int kStages = MASK == 2 ? 3 : 5;
  static constexpr int kOffBars = kStages * kStageBytes;
  static constexpr int kTotal = kOffBars + 2 * kStages * 8;
  static_assert(kStages * kStageBytes >= 16 * 256 * 4, "the block reduction reuses the ring memory");
};
constexpr int kRingThreads = 288;   // 8 consumer warps + 1 producer warp

template <int MASK>
__device__ __forceinline__ void bn_ring_producer(uint32_t sbase, const uint4* dy, const uint4* x, const uint4* out,
                                                 int64_t nfull) {
  using R = BnRing<MASK>;
  int s = 0;
  uint32_t ph = 0;
  for (int64_t c = blockIdx.x; c < nfull; c += gridDim.x) {
    const uint32_t full = sbase + R::kOffBars + 8u * s, empty = sbase + R::kOffBars + 8u * (R::kStages + s);
    mbar_wait(empty, ph ^ 1);
    mbar_arrive_expect_tx(full, R::kStageBytes);
    const uint32_t dst = sbase + s * R::kStageBytes;
    bulk_load_1d(dst, dy + c * kRingVec, R::kTensorBytes, full);
    bulk_load_1d(dst + R::kTensorBytes, x + c * kRingVec, R::kTensorBytes, full);
    if (MASK == 2) bulk_load_1d(dst + 2 * R::kTensorBytes, out + c * kRingVec, R::kTensorBytes, full);
    if (MASK == 3)
      bulk_load_1d(dst + 2 * R::kTensorBytes, reinterpret_cast<const uint8_t*>(out) + c * kRingVec, kRingVec, full);
    if (++s == R::kStages) { s = 0; ph ^= 1; }
  }
}
__device__ __forceinline__ void bn_ring_init(uint32_t sbase, int off_bars, int stages) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(sbase + off_bars + 8u * s, 1);
      mbar_init(sbase + off_bars + 8u * (stages + s), 8);   // one arrive per consumer warp
    }
    fence_mbar_init();
  }
  __syncthreads();
}
// masked gradient of eight channels (see bn_bwd_reduce_kernel)
template <int MASK>
__device__ __forceinline__ F8 bn_masked_grad(const F8& d, const F8& xv, const uint4& ov4, const F8& sc, const F8& sh) {
  F8 g = d;
  if (MASK == 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = fmaf(xv.v[k], sc.v[k], sh.v[k]) > 0.f ? d.v[k] : 0.f;
  } else if (MASK == 2) {
    const F8 o = unpack8(ov4);
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = o.v[k] > 0.f ? d.v[k] : 0.f;
  } else if (MASK == 3) {
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = ((ov4.x >> k) & 1u) ? d.v[k] : 0.f;
  }
  return g;
}

template <int MASK>
__global__ void __launch_bounds__(kRingThreads, 2)
bn_bwd_reduce_ring_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ out,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          float* __restrict__ partial, int64_t nvec, int cvec) {
  pdl_prologue();
  using R = BnRing<MASK>;
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const uint32_t sbase = smem_u32(ring_smem);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  bn_ring_init(sbase, R::kOffBars, R::kStages);
  const int64_t nfull = nvec / kRingVec;
  const int lanes = cvec < 256 ? cvec : 256;
  const int oc = t & (lanes - 1);
  const int c0 = oc * 8;
  float a_dy[8], a_dyx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a_dy[k] = a_dyx[k] = 0.f;
  if (warp == 8) {
    if (lane == 0) bn_ring_producer<MASK>(sbase, dy, x, out, nfull);
  } else {
    F8 sc, sh;
    if (MASK == 1) { sc = load8f(scale + c0); sh = load8f(shift + c0); }
    auto body = [&](const uint4& dyv, const uint4& xv4, const uint4& ov4) {
      const F8 d = unpack8(dyv);
      const F8 xv = unpack8(xv4);
      const F8 g = bn_masked_grad<MASK>(d, xv, ov4, sc, sh);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a_dy[k] += g.v[k];
        a_dyx[k] = fmaf(g.v[k], xv.v[k], a_dyx[k]);
      }
    };
    int s = 0;
    uint32_t ph = 0;
    for (int64_t c = blockIdx.x; c < nfull; c += gridDim.x) {
      mbar_wait(sbase + R::kOffBars + 8u * s, ph);
      const uint8_t* st = ring_smem + s * R::kStageBytes;
      const uint4* sd = reinterpret_cast<const uint4*>(st);
      const uint4* sx = reinterpret_cast<const uint4*>(st + R::kTensorBytes);
      const uint4 d0 = sd[t], d1 = sd[t + 256], x0 = sx[t], x1 = sx[t + 256];
      uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
      if (MASK == 2) {
        const uint4* so = reinterpret_cast<const uint4*>(st + 2 * R::kTensorBytes);
        o0 = so[t]; o1 = so[t + 256];
      }
      if (MASK == 3) {
        const uint8_t* sb = st + 2 * R::kTensorBytes;
        o0.x = sb[t]; o1.x = sb[t + 256];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sbase + R::kOffBars + 8u * (R::kStages + s));   // stage free: data is in registers
      body(d0, x0, o0);
      body(d1, x1, o1);
      if (++s == R::kStages) { s = 0; ph ^= 1; }
    }
    if (blockIdx.x == 0) {
      // ragged tail (< one stage): straight from global memory
      const uint8_t* bits = reinterpret_cast<const uint8_t*>(out);
      for (int64_t i = nfull * kRingVec + t; i < nvec; i += 256) {
        uint4 o0 = make_uint4(0, 0, 0, 0);
        if (MASK == 2) o0 = ldg_stream(out + i);
        if (MASK == 3) o0.x = __ldg(bits + i);
        body(ldg_stream(dy + i), ldg_stream(x + i), o0);
      }
    }
  }
  // Block reduction over the threads that share a channel octet (t = oc mod lanes): warp shuffles across the lanes of
  // a warp that share it (lanes < 32), then one shared-memory pass across the eight warps. Fixed order -> deterministic.
  float v[16];
  {
    F8 mu, is;
    if (t < 256) { mu = load8f(mean + c0); is = load8f(invstd + c0); }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = a_dy[k];
      v[8 + k] = t < 256 ? (a_dyx[k] - mu.v[k] * a_dy[k]) * is.v[k] : 0.f;
    }
  }
  if (warp < 8) {
    for (int o = 16; o >= lanes; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
  }
  __syncthreads();   // every stage has been consumed: the ring memory becomes the reduction scratch
  float (*red)[16][32] = reinterpret_cast<float (*)[16][32]>(ring_smem);   // [warp][value][lane]
  if (warp < 8) {
#pragma unroll
    for (int k = 0; k < 16; ++k) red[warp][k][lane] = v[k];
  }
  __syncthreads();
  // thread t < lanes owns octet t; the warps holding it are w = (t / 32) + j * (lanes / 32) for lanes >= 32, all eight
  // warps (lane t) for lanes < 32
  if (t < lanes) {
    const int wstep = lanes >= 32 ? lanes / 32 : 1;
    const int w0 = lanes >= 32 ? t / 32 : 0;
    const int ln = t & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s0 = 0.f, s1 = 0.f;
      for (int w = w0; w < 8; w += wstep) {
        s0 += red[w][k][ln];
        s1 += red[w][8 + k][ln];
      }
      partial[(static_cast<size_t>(blockIdx.x) * 2 + 0) * (cvec * 8) + c0 + k] = s0;
      partial[(static_cast<size_t>(blockIdx.x) * 2 + 1) * (cvec * 8) + c0 + k] = s1;
    }
  }
}

template <int MASK>
__global__ void __launch_bounds__(kRingThreads, 2)
bn_bwd_apply_ring_kernel(uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ out,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ dgamma, const float* __restrict__ dbeta, uint4* __restrict__ dx,
                         int64_t nvec, int cvec, float inv_rows) {
  pdl_prologue();
  using R = BnRing<MASK>;
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const uint32_t sbase = smem_u32(ring_smem);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  bn_ring_init(sbase, R::kOffBars, R::kStages);
  const int64_t nfull = nvec / kRingVec;
  if (warp == 8) {
    if (lane == 0) bn_ring_producer<MASK>(sbase, dy, x, out, nfull);
    return;
  }
  const int c0 = (t & ((cvec < 256 ? cvec : 256) - 1)) * 8;
  const F8 sc = load8f(scale + c0);
  F8 sh;
  if (MASK == 1) sh = load8f(shift + c0);
  F8 k0, k1;
  {
    const F8 mu = load8f(mean + c0), is = load8f(invstd + c0), dg = load8f(dgamma + c0), db = load8f(dbeta + c0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float c2 = sc.v[k] * dg.v[k] * inv_rows * is.v[k];
      k1.v[k] = -c2;
      k0.v[k] = fmaf(c2, mu.v[k], -sc.v[k] * db.v[k] * inv_rows);
    }
  }
  auto body = [&](int64_t i, const uint4& dyv, const uint4& xv4, const uint4& ov4) {
    const F8 d = unpack8(dyv);
    const F8 xv = unpack8(xv4);
    const F8 g = bn_masked_grad<MASK>(d, xv, ov4, sc, sh);
    F8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = fmaf(sc.v[k], g.v[k], fmaf(k1.v[k], xv.v[k], k0.v[k]));
    dx[i] = pack8(r);
    if (MASK == 2) dy[i] = pack8(g);
  };
  int s = 0;
  uint32_t ph = 0;
  for (int64_t c = blockIdx.x; c < nfull; c += gridDim.x) {
    mbar_wait(sbase + R::kOffBars + 8u * s, ph);
    const uint8_t* st = ring_smem + s * R::kStageBytes;
    const uint4* sd = reinterpret_cast<const uint4*>(st);
    const uint4* sx = reinterpret_cast<const uint4*>(st + R::kTensorBytes);
    const uint4 d0 = sd[t], d1 = sd[t + 256], x0 = sx[t], x1 = sx[t + 256];
    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
    if (MASK == 2) {
      const uint4* so = reinterpret_cast<const uint4*>(st + 2 * R::kTensorBytes);
      o0 = so[t]; o1 = so[t + 256];
    }
    if (MASK == 3) {
      const uint8_t* sb = st + 2 * R::kTensorBytes;
      o0.x = sb[t]; o1.x = sb[t + 256];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sbase + R::kOffBars + 8u * (R::kStages + s));
    const int64_t i0 = c * kRingVec + t;
    body(i0, d0, x0, o0);
    body(i0 + 256, d1, x1, o1);
    if (++s == R::kStages) { s = 0; ph ^= 1; }
  }
  if (blockIdx.x == 0) {
    const uint8_t* bits = reinterpret_cast<const uint8_t*>(out);
    for (int64_t i = nfull * kRingVec + t; i < nvec; i += 256) {
      uint4 o0 = make_uint4(0, 0, 0, 0);
      if (MASK == 2) o0 = ldg_stream(out + i);
      if (MASK == 3) o0.x = __ldg(bits + i);
      body(i, dy[i], ldg_stream(x + i), o0);
    }
  }
}

// Forward batch-norm apply through the same ring: x (+ residual) stream in, y (+ ReLU bits, + per-block column sums) out.
template <int RES>
struct ApplyRing {
  static constexpr int kTensorBytes = kRingVec * 16;
  static constexpr int kStageBytes = RES ? 2 * kTensorBytes : kTensorBytes;
  static constexpr int kStages = RES ? 5 : 10;   // ~80 KB per CTA (see BnRing)
  static constexpr int kOffBars = kStages * kStageBytes;
  static constexpr int kTotal = kOffBars + 2 * kStages * 8;
  static_assert(kStages * kStageBytes >= 8 * 16 * 32 * 4, "the block reduction reuses the ring memory");
};
template <int RES, bool SUM>
__global__ void __launch_bounds__(kRingThreads, 2)
bn_apply_ring_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                     const uint4* __restrict__ res, const float* __restrict__ rscale, const float* __restrict__ rshift,
                     int relu, uint4* __restrict__ y, uint8_t* __restrict__ bits, float* __restrict__ colsum_partial,
                     int64_t nvec, int cvec) {
  pdl_prologue();
  using R = ApplyRing<RES>;
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const uint32_t sbase = smem_u32(ring_smem);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  bn_ring_init(sbase, R::kOffBars, R::kStages);
  const int64_t nfull = nvec / kRingVec;
  const int lanes = cvec < 256 ? cvec : 256;
  const int c0 = (t & (lanes - 1)) * 8;
  F8 acc;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc.v[k] = 0.f;
  if (warp == 8) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t c = blockIdx.x; c < nfull; c += gridDim.x) {
        const uint32_t full = sbase + R::kOffBars + 8u * s, empty = sbase + R::kOffBars + 8u * (R::kStages + s);
        mbar_wait(empty, ph ^ 1);
        mbar_arrive_expect_tx(full, R::kStageBytes);
        const uint32_t dst = sbase + s * R::kStageBytes;
        bulk_load_1d(dst, x + c * kRingVec, R::kTensorBytes, full);
        if (RES) bulk_load_1d(dst + R::kTensorBytes, res + c * kRingVec, R::kTensorBytes, full);
        if (++s == R::kStages) { s = 0; ph ^= 1; }
      }
    }
    if (!SUM) return;
  } else {
    const F8 sc = load8f(scale + c0), sh = load8f(shift + c0);
    F8 rs, rb;
    if (RES == 2) { rs = load8f(rscale + c0); rb = load8f(rshift + c0); }
    auto one = [&](int64_t i, const uint4& xq, const uint4& rq) {
      const F8 xv = unpack8(xq);
      F8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = fmaf(xv.v[k], sc.v[k], sh.v[k]);
      if (RES == 1) {
        const F8 rv = unpack8(rq);
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] += rv.v[k];
      } else if (RES == 2) {
        const F8 rv = unpack8(rq);
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] += fmaf(rv.v[k], rs.v[k], rb.v[k]);
      }
      if (bits != nullptr) bits[i] = positive_bits(o);
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = fmaxf(o.v[k], 0.f);
      }
      const uint4 packed = pack8(o);
      y[i] = packed;
      if (SUM) {
        const F8 r = unpack8(packed);   // sums of the STORED (bf16) values
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.v[k] += r.v[k];
      }
    };
    int s = 0;
    uint32_t ph = 0;
    for (int64_t c = blockIdx.x; c < nfull; c += gridDim.x) {
      mbar_wait(sbase + R::kOffBars + 8u * s, ph);
      const uint8_t* st = ring_smem + s * R::kStageBytes;
      const uint4* sx = reinterpret_cast<const uint4*>(st);
      const uint4 x0 = sx[t], x1 = sx[t + 256];
      uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
      if (RES) {
        const uint4* sr = reinterpret_cast<const uint4*>(st + R::kTensorBytes);
        r0 = sr[t]; r1 = sr[t + 256];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sbase + R::kOffBars + 8u * (R::kStages + s));
      const int64_t i0 = c * kRingVec + t;
      one(i0, x0, r0);
      one(i0 + 256, x1, r1);
      if (++s == R::kStages) { s = 0; ph ^= 1; }
    }
    if (blockIdx.x == 0) {
      for (int64_t i = nfull * kRingVec + t; i < nvec; i += 256) {
        uint4 r0 = make_uint4(0, 0, 0, 0);
        if (RES) r0 = ldg_stream(res + i);
        one(i, ldg_stream(x + i), r0);
      }
    }
  }
  if (SUM) {
    // per-block column sums: same reduction scheme as bn_bwd_reduce_ring_kernel
    if (warp < 8) {
      for (int o = 16; o >= lanes; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.v[k] += __shfl_xor_sync(0xffffffffu, acc.v[k], o);
      }
    }
    __syncthreads();
    float (*red)[8][32] = reinterpret_cast<float (*)[8][32]>(ring_smem);   // [warp][value][lane]
    if (warp < 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red[warp][k][lane] = acc.v[k];
    }
    __syncthreads();
    if (t < lanes) {
      const int wstep = lanes >= 32 ? lanes / 32 : 1;
      const int w0 = lanes >= 32 ? t / 32 : 0;
      const int ln = t & 31;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float s0 = 0.f;
        for (int w = w0; w < 8; w += wstep) s0 += red[w][k][ln];
        colsum_partial[static_cast<size_t>(blockIdx.x) * (cvec * 8) + t * 8 + k] = s0;
      }
    }
  }
}
template <int RES, bool SUM>
static void launch_bn_apply_ring_t(const uint4* x, const float* scale, const float* shift, const uint4* res,
                                   const float* rscale, const float* rshift, int relu, uint4* y, uint8_t* bits,
                                   float* colsum_partial, int64_t nvec, int cvec, cudaStream_t s) {
  static const bool once = [] {
    ARGUS_CUDA(cudaFuncSetAttribute(bn_apply_ring_kernel<RES, SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    ApplyRing<RES>::kTotal));
    return true;
  }();
  (void)once;
  launch_kernel(bn_apply_ring_kernel<RES, SUM>, bn_ring_grid(nvec), kRingThreads, ApplyRing<RES>::kTotal, s, 
      x, scale, shift, res, rscale, rshift, relu, y, bits, colsum_partial, nvec, cvec);
}
static void launch_bn_apply_ring(const uint4* x, const float* scale, const float* shift, const uint4* res,
                                 const float* rscale, const float* rshift, int relu, uint4* y, uint8_t* bits,
                                 float* colsum_partial, int64_t nvec, int cvec, cudaStream_t s) {
  if (colsum_partial != nullptr) launch_bn_apply_ring_t<0, true>(x, scale, shift, res, rscale, rshift, relu, y, bits, colsum_partial, nvec, cvec, s);
  else if (res == nullptr) launch_bn_apply_ring_t<0, false>(x, scale, shift, res, rscale, rshift, relu, y, bits, nullptr, nvec, cvec, s);
  else if (rscale == nullptr) launch_bn_apply_ring_t<1, false>(x, scale, shift, res, rscale, rshift, relu, y, bits, nullptr, nvec, cvec, s);
  else launch_bn_apply_ring_t<2, false>(x, scale, shift, res, rscale, rshift, relu, y, bits, nullptr, nvec, cvec, s);
}

static bool bn_ring_enabled() {
  static const bool on = [] { const char* e = getenv("ARGUS_BN_RING"); return !(e && e[0] == '0'); }();
  return on;
}
static int bn_ring_grid(int64_t nvec) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nvec / kRingVec, 2LL * num_sms())));
}
template <int MASK>
static void launch_bn_bwd_reduce_ring(const uint4* dy, const uint4* x, const uint4* out, const float* scale,
                                      const float* shift, const float* mean, const float* invstd, float* partial,
                                      int64_t nvec, int cvec, int grid, cudaStream_t s) {
  static const bool once = [] {
    ARGUS_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_ring_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    BnRing<MASK>::kTotal));
    return true;
  }();
  (void)once;
  launch_kernel(bn_bwd_reduce_ring_kernel<MASK>, grid, kRingThreads, BnRing<MASK>::kTotal, s, dy, x, out, scale, shift, mean, invstd,
                                                                                  partial, nvec, cvec);
}
template <int MASK>
static void launch_bn_bwd_apply_ring(uint4* dy, const uint4* x, const uint4* out, const float* scale,
                                     const float* shift, const float* mean, const float* invstd, const float* dgamma,
                                     const float* dbeta, uint4* dx, int64_t nvec, int cvec, float inv_rows,
                                     cudaStream_t s) {
  static const bool once = [] {
    ARGUS_CUDA(cudaFuncSetAttribute(bn_bwd_apply_ring_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    BnRing<MASK>::kTotal));
    return true;
  }();
  (void)once;
  launch_kernel(bn_bwd_apply_ring_kernel<MASK>, bn_ring_grid(nvec), kRingThreads, BnRing<MASK>::kTotal, s, 
      dy, x, out, scale, shift, mean, invstd, dgamma, dbeta, dx, nvec, cvec, inv_rows);
}

int64_t bn_bwd_scratch_elems() { return 4LL * num_sms() * 2 * 2048; }

void bn_bwd_reduce(const bf16* dy, const bf16* x, const bf16* out, const float* scale, const float* shift,
                   const float* mean, const float* invstd, float* dgamma, float* dbeta, int64_t rows, int C,
                   int mask_mode, float* scratch, cudaStream_t s) {
  ARGUS_CHECK(scratch != nullptr, "bn_bwd_reduce needs a scratch buffer");
  std::string fam = "bn_bwd_reduce";
  if (profile_enabled() && profile_detailed()) fam += ":R" + std::to_string(rows) + "_C" + std::to_string(C) + "_m" + std::to_string(mask_mode);
  ProfileScope prof(fam, s, 0, static_cast<double>(rows) * C * (mask_mode == 2 ? 6.0 : (mask_mode == 3 ? 4.125 : 4.0)));
  ARGUS_CHECK(C % 8 == 0 && is_pow2(C / 8) && C <= 2048, "bn_bwd_reduce: C/8 must be a power of two <= 256");
  const int cvec = C / 8;
  const int lanes = std::min(cvec, 256);
  const int row_lanes = 256 / lanes;
  const int64_t row_groups = (rows + row_lanes - 1) / row_lanes;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(row_groups, 4LL * num_sms())));
  auto DY = reinterpret_cast<const uint4*>(dy);
  auto X = reinterpret_cast<const uint4*>(x);
  auto O = reinterpret_cast<const uint4*>(out);
  if (bn_ring_enabled() && aligned16(dy) && aligned16(x) && aligned16(out)) {
    const int64_t nvec = rows * cvec;
    const int rgrid = bn_ring_grid(nvec);
    switch (mask_mode) {
      case 0: launch_bn_bwd_reduce_ring<0>(DY, X, O, scale, shift, mean, invstd, scratch, nvec, cvec, rgrid, s); break;
      case 1: launch_bn_bwd_reduce_ring<1>(DY, X, O, scale, shift, mean, invstd, scratch, nvec, cvec, rgrid, s); break;
      case 2:
        ARGUS_CHECK(out != nullptr, "mask_mode 2 needs the block output");
        launch_bn_bwd_reduce_ring<2>(DY, X, O, scale, shift, mean, invstd, scratch, nvec, cvec, rgrid, s);
        break;
      case 3:
        ARGUS_CHECK(out != nullptr, "mask_mode 3 needs the ReLU bit mask");
        launch_bn_bwd_reduce_ring<3>(DY, X, O, scale, shift, mean, invstd, scratch, nvec, cvec, rgrid, s);
        break;
      default: throw Error("bad mask_mode");
    }
    ARGUS_CUDA(cudaGetLastError());
    launch_kernel(bn_bwd_finalize_kernel, (C + 7) / 8, 256, 0, s, scratch, rgrid, dgamma, dbeta, C, nullptr, nullptr);
    ARGUS_CUDA(cudaGetLastError());
    return;
  }
  switch (mask_mode) {
    case 0: launch_kernel(bn_bwd_reduce_kernel<0>, grid, 256, 0, s, DY, X, O, scale, shift, mean, invstd, scratch, rows, cvec); break;
    case 1: launch_kernel(bn_bwd_reduce_kernel<1>, grid, 256, 0, s, DY, X, O, scale, shift, mean, invstd, scratch, rows, cvec); break;
    case 2:
      ARGUS_CHECK(out != nullptr, "mask_mode 2 needs the block output");
      launch_kernel(bn_bwd_reduce_kernel<2>, grid, 256, 0, s, DY, X, O, scale, shift, mean, invstd, scratch, rows, cvec);
      break;
    case 3:
      ARGUS_CHECK(out != nullptr, "mask_mode 3 needs the ReLU bit mask");
      launch_kernel(bn_bwd_reduce_kernel<3>, grid, 256, 0, s, DY, X, O, scale, shift, mean, invstd, scratch, rows, cvec);
      break;
    default: throw Error("bad mask_mode");
  }
  ARGUS_CUDA(cudaGetLastError());
  launch_kernel(bn_bwd_finalize_kernel, (C + 7) / 8, 256, 0, s, scratch, grid, dgamma, dbeta, C, nullptr, nullptr);
  ARGUS_CUDA(cudaGetLastError());
}

void bn_bwd_finalize_slots(const float* partial, int slots, const float* mean, const float* invstd, float* dgamma,
                           float* dbeta, int C, cudaStream_t s) {
  ProfileScope prof("bn_bwd_reduce", s, 0, static_cast<double>(slots) * 2 * C * 4);
  launch_kernel(bn_bwd_finalize_kernel, (C + 7) / 8, 256, 0, s, partial, slots, dgamma, dbeta, C, mean, invstd);
  ARGUS_CUDA(cudaGetLastError());
}

template <int MASK>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ out,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, uint4* __restrict__ dx,
                    int64_t nvec, int cvec, float inv_rows) {
  pdl_prologue();
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;   // multiple of cvec
  const int c0 = static_cast<int>(tid & (cvec - 1)) * 8;
  const F8 sc = load8f(scale + c0), sh = load8f(shift + c0);
  // dx = sc * (g - db/M - xhat * dg/M) with xhat = (x - mu) * is   ==   sc*g + k1*x + k0
  F8 k0, k1;
  {
    const F8 mu = load8f(mean + c0), is = load8f(invstd + c0), dg = load8f(dgamma + c0), db = load8f(dbeta + c0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float c2 = sc.v[k] * dg.v[k] * inv_rows * is.v[k];
      k1.v[k] = -c2;
      k0.v[k] = fmaf(c2, mu.v[k], -sc.v[k] * db.v[k] * inv_rows);
    }
  }
  const uint8_t* bits = reinterpret_cast<const uint8_t*>(out);   // MASK == 3: one byte per channel octet
  auto body = [&](int64_t i, const uint4& dyv, const uint4& xv4, const uint4& ov4) {
    const F8 d = unpack8(dyv);
    const F8 xv = unpack8(xv4);
    F8 g = d;
    if (MASK == 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) g.v[k] = fmaf(xv.v[k], sc.v[k], sh.v[k]) > 0.f ? d.v[k] : 0.f;
    } else if (MASK == 2) {
      const F8 o = unpack8(ov4);
#pragma unroll
      for (int k = 0; k < 8; ++k) g.v[k] = o.v[k] > 0.f ? d.v[k] : 0.f;
    } else if (MASK == 3) {
#pragma unroll
      for (int k = 0; k < 8; ++k) g.v[k] = ((ov4.x >> k) & 1u) ? d.v[k] : 0.f;
    }
    F8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = fmaf(sc.v[k], g.v[k], fmaf(k1.v[k], xv.v[k], k0.v[k]));
    dx[i] = pack8(r);
    if (MASK == 2) dy[i] = pack8(g);
  };
  int64_t i = tid;
  for (; i + stride < nvec; i += 2 * stride) {
    const uint4 da = dy[i], dbv = dy[i + stride];
    const uint4 xa = ldg_stream(x + i), xb = ldg_stream(x + i + stride);
    uint4 oa = make_uint4(0, 0, 0, 0), ob = oa;
    if (MASK == 2) { oa = ldg_stream(out + i); ob = ldg_stream(out + i + stride); }
    if (MASK == 3) { oa.x = __ldg(bits + i); ob.x = __ldg(bits + i + stride); }
    body(i, da, xa, oa);
    body(i + stride, dbv, xb, ob);
  }
  for (; i < nvec; i += stride) {
    uint4 oa = make_uint4(0, 0, 0, 0);
    if (MASK == 2) oa = ldg_stream(out + i);
    if (MASK == 3) oa.x = __ldg(bits + i);
    body(i, dy[i], ldg_stream(x + i), oa);
  }
}

void bn_bwd_apply(bf16* dy, const bf16* x, const bf16* out, const float* scale, const float* shift,
                  const float* mean, const float* invstd, const float* dgamma, const float* dbeta, bf16* dx,
                  int64_t rows, int C, int mask_mode, cudaStream_t s) {
  ProfileScope prof("bn_bwd_apply", s, 0, static_cast<double>(rows) * C * (mask_mode == 2 ? 10.0 : (mask_mode == 3 ? 6.125 : 6.0)));
  ARGUS_CHECK(C % 8 == 0 && is_pow2(C / 8) && C <= 2048, "bn_bwd_apply: C/8 must be a power of two <= 256");
  const int cvec = C / 8;
  const int64_t nvec = rows * cvec;
  const int grid = grid_for(nvec, 256);
  const float inv_rows = static_cast<float>(1.0 / static_cast<double>(rows));
  auto DY = reinterpret_cast<uint4*>(dy);
  auto X = reinterpret_cast<const uint4*>(x);
  auto O = reinterpret_cast<const uint4*>(out);
  auto DX = reinterpret_cast<uint4*>(dx);
  if (bn_ring_enabled() && aligned16(dy) && aligned16(x) && aligned16(out) && aligned16(dx)) {
    switch (mask_mode) {
      case 0: launch_bn_bwd_apply_ring<0>(DY, X, O, scale, shift, mean, invstd, dgamma, dbeta, DX, nvec, cvec, inv_rows, s); break;
      case 1: launch_bn_bwd_apply_ring<1>(DY, X, O, scale, shift, mean, invstd, dgamma, dbeta, DX, nvec, cvec, inv_rows, s); break;
      case 2: launch_bn_bwd_apply_ring<2>(DY, X, O, scale, shift, mean, invstd, dgamma, dbeta, DX, nvec, cvec, inv_rows, s); break;
      case 3: launch_bn_bwd_apply_ring<3>(DY, X, O, scale, shift, mean, invstd, dgamma, dbeta, DX, nvec, cvec, inv_rows, s); break;
      default: throw Error("bad mask_mode");
    }
    ARGUS_CUDA(cudaGetLastError());
    return;
  }
  switch (mask_mode) {
    case 0: launch_kernel(bn_bwd_apply_kernel<0>, grid, 256, 0, s, DY, X, O, scale, shift, mean, invstd, dgamma, dbeta, DX, nvec, cvec, inv_rows); break;
    case 1: launch_kernel(bn_bwd_apply_kernel<1>, grid, 256, 0, s, DY, X, O, scale, shift, mean, invstd, dgamma, dbeta, DX, nvec, cvec, inv_rows); break;
    case 2: launch_kernel(bn_bwd_apply_kernel<2>, grid, 256, 0, s, DY, X, O, scale, shift, mean, invstd, dgamma, dbeta, DX, nvec, cvec, inv_rows); break;
    case 3: launch_kernel(bn_bwd_apply_kernel<3>, grid, 256, 0, s, DY, X, O, scale, shift, mean, invstd, dgamma, dbeta, DX, nvec, cvec, inv_rows); break;
    default: throw Error("bad mask_mode");
  }
  ARGUS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// pooling
// ------------------------------------------------------------------------------------------------------------
// Packed formulation: relu(x * sc + sh) is monotone in x (increasing for sc > 0, decreasing for sc < 0), so the window
// maximum of the activated values is the activation of the window maximum of sign(sc) * x. The nine taps are compared
// as raw bf16 pairs (sign flip = one XOR per pair, compare mask + max + index select: four packed instructions per tap and
// channel pair instead of seven scalar ones per channel), and the batch norm + ReLU are applied once per output. The
// arg-max differs from "first maximum of the activated values" only where the whole window is <= 0 after the ReLU,
// and there the gradient is zero anyway (the ReLU mask of the backward pass removes it).
__device__ __forceinline__ uint32_t bf16x2_gt_mask(uint32_t a, uint32_t b) {   // 0xffff per lane where a > b
  uint32_t m;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(b));
  return m;
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  uint32_t m;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(b));
  return m;
}
// SHIFT: W / 2, H / 2 and C / 8 are powers of two (every size the bf16 path accepts): the output index is split with
// shifts; otherwise with 32-bit divisions. All nine taps are loaded unconditionally from clamped coordinates (nine
// independent 16-byte loads in flight per thread) and the out-of-bounds ones replaced by -inf afterwards.
template <bool SHIFT>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                   uint4* __restrict__ y, uint2* __restrict__ idx, int N, int H, int W, int cvec, int lg_c, int lg_wo,
                   int lg_ho) {
  pdl_prologue();
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * cvec;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int cv, pw, ph, n;
    if (SHIFT) {
      const uint32_t u = static_cast<uint32_t>(i);   // total < 2^32 is checked by the launcher
      cv = u & (cvec - 1);
      pw = (u >> lg_c) & (Wo - 1);
      ph = (u >> (lg_c + lg_wo)) & (Ho - 1);
      n = u >> (lg_c + lg_wo + lg_ho);
    } else {
      cv = static_cast<int>(i % cvec);
      int64_t t = i / cvec;
      pw = static_cast<int>(t % Wo);
      t /= Wo;
      ph = static_cast<int>(t % Ho);
      n = static_cast<int>(t / Ho);
    }
    F8 sc, sh;
    uint32_t flip[4] = {0u, 0u, 0u, 0u};
    if (scale != nullptr) {
      sc = load8f(scale + cv * 8);
      sh = load8f(shift + cv * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        flip[j] = (sc.v[2 * j] < 0.f ? 0x00008000u : 0u) | (sc.v[2 * j + 1] < 0.f ? 0x80000000u : 0u);
    }
    // window rows 2ph-1 .. 2ph+1, columns 2pw-1 .. 2pw+1: only the first row / column can be outside (H, W even)
    const bool top_ok = ph > 0, left_ok = pw > 0;
    const uint4* base = x + (static_cast<int64_t>(n) * H * W) * cvec + cv;
    uint4 tap[3][3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = max(2 * ph - 1 + kh, 0);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = max(2 * pw - 1 + kw, 0);
        tap[kh][kw] = __ldg(base + (h * W + w) * cvec);
      }
    }
    uint32_t best[4], bi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { best[j] = 0xff80ff80u; bi[j] = 0u; }   // -inf, tap 0
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const bool ok = (kh > 0 || top_ok) && (kw > 0 || left_ok);
        const uint4 u = tap[kh][kw];
        const uint32_t wv[4] = {ok ? (u.x ^ flip[0]) : 0xff80ff80u, ok ? (u.y ^ flip[1]) : 0xff80ff80u,
                                ok ? (u.z ^ flip[2]) : 0xff80ff80u, ok ? (u.w ^ flip[3]) : 0xff80ff80u};
        const uint32_t code = static_cast<uint32_t>(kh * 3 + kw) * 0x00010001u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t m = bf16x2_gt_mask(wv[j], best[j]);   // strict: the first maximum keeps its index
          best[j] = bf16x2_max(wv[j], best[j]);
          bi[j] = (bi[j] & ~m) | (code & m);
        }
      }
    F8 o = unpack8(make_uint4(best[0] ^ flip[0], best[1] ^ flip[1], best[2] ^ flip[2], best[3] ^ flip[3]));
    if (scale != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = fmaxf(fmaf(o.v[k], sc.v[k], sh.v[k]), 0.f);
    }
    y[i] = pack8(o);
    if (idx != nullptr) {
      uint2 p;
      p.x = (bi[0] & 0xffu) | ((bi[0] >> 8) & 0xff00u) | ((bi[1] & 0xffu) << 16) | ((bi[1] >> 16) << 24);
      p.y = (bi[2] & 0xffu) | ((bi[2] >> 8) & 0xff00u) | ((bi[3] & 0xffu) << 16) | ((bi[3] >> 16) << 24);
      idx[i] = p;
    }
  }
}
static int log2_exact(int v) {   // log2 of a power of two, -1 otherwise
  if (v <= 0 || (v & (v - 1))) return -1;
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
void maxpool_fwd(const bf16* x, const float* scale, const float* shift, bf16* y, uint8_t* idx, int N, int H, int W,
                 int C, cudaStream_t s) {
  ProfileScope prof("maxpool", s, 0, static_cast<double>(N) * H * W * C * 2 * 1.25 + (idx ? static_cast<double>(N) * H * W * C / 4 : 0.0));
  const int64_t total = static_cast<int64_t>(N) * (H / 2) * (W / 2) * (C / 8);
  const int lg_c = log2_exact(C / 8), lg_wo = log2_exact(W / 2), lg_ho = log2_exact(H / 2);
  const bool shift_ok = lg_c >= 0 && lg_wo >= 0 && lg_ho >= 0 && total < (1LL << 32) &&
                        static_cast<int64_t>(H) * W * (C / 8) < (1LL << 31);
  if (shift_ok)
    launch_kernel(maxpool_fwd_kernel<true>, grid_for(total, 256), 256, 0, s, reinterpret_cast<const uint4*>(x), scale, shift,
                  reinterpret_cast<uint4*>(y), reinterpret_cast<uint2*>(idx), N, H, W, C / 8, lg_c, lg_wo, lg_ho);
  else
    launch_kernel(maxpool_fwd_kernel<false>, grid_for(total, 256), 256, 0, s, reinterpret_cast<const uint4*>(x), scale, shift,
                  reinterpret_cast<uint4*>(y), reinterpret_cast<uint2*>(idx), N, H, W, C / 8, 0, 0, 0);
  ARGUS_CUDA(cudaGetLastError());
}

__global__ void maxpool_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx, uint4* __restrict__ dx,
                                   int N, int H, int W, int cvec) {
  pdl_prologue();
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t total = static_cast<int64_t>(N) * H * W * cvec;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % cvec);
    int64_t t = i / cvec;
    const int w = static_cast<int>(t % W);
    t /= W;
    const int h = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    F8 g;
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = 0.f;
    // pooled windows containing (h, w): ph in {h/2, (h+1)/2}, pw likewise
    for (int ph = h >> 1; ph <= ((h + 1) >> 1); ++ph) {
      if (ph >= Ho) continue;
      const int kh = h - (2 * ph - 1);
      for (int pw = w >> 1; pw <= ((w + 1) >> 1); ++pw) {
        if (pw >= Wo) continue;
        const int kw = w - (2 * pw - 1);
        const int code = kh * 3 + kw;
        const int64_t j = ((static_cast<int64_t>(n) * Ho + ph) * Wo + pw) * cvec + cv;
        const uint2 id = __ldg(idx + j);
        const F8 d = unpack8(__ldg(dy + j));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int b = (k < 4 ? (id.x >> (8 * k)) : (id.y >> (8 * (k - 4)))) & 0xff;
          if (b == code) g.v[k] += d.v[k];
        }
      }
    }
    dx[i] = pack8(g);
  }
}
void maxpool_bwd(const bf16* dy, const uint8_t* idx, bf16* dx, int N, int H, int W, int C, cudaStream_t s) {
  ProfileScope prof("maxpool", s, 0, static_cast<double>(N) * H * W * C * 2 * 1.25 + static_cast<double>(N) * H * W * C / 4);
  const int64_t total = static_cast<int64_t>(N) * H * W * (C / 8);
  launch_kernel(maxpool_bwd_kernel, grid_for(total, 256), 256, 0, s, reinterpret_cast<const uint4*>(dy),
                                                          reinterpret_cast<const uint2*>(idx),
                                                          reinterpret_cast<uint4*>(dx), N, H, W, C / 8);
  ARGUS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// stem: max-pool backward fused into the batch-norm (+ReLU) backward of the stem convolution
// ------------------------------------------------------------------------------------------------------------
// The gradient wrt the stem activation is never materialised: both passes (per-channel sums, then dx) rebuild it from
// the pooled gradient and the arg-max bytes. Thread = one 2x2 patch of stem pixels x 8 channels: the patch touches
// exactly four pooling windows -- (a,b), (a,b+1), (a+1,b), (a+1,b+1) -- so 4 window loads serve 4 pixels (the plain
// max-pool backward kernel needs up to 4 per pixel), and the 1.07 GB intermediate (write + two reads) disappears.
template <int APPLY>
__global__ void __launch_bounds__(256)   // (256, 3) forces 80 registers and spills: measured 0.60 -> 0.83 ms
stem_pool_bn_bwd_kernel(const uint4* __restrict__ dpool, const uint2* __restrict__ idx, const uint4* __restrict__ raw,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ partial,
                        uint4* __restrict__ dx, int N, int H, int W, float inv_rows, int lg_wo, int lg_ho) {
  pdl_prologue();
  constexpr int cvec = 8;   // C = 64
  __shared__ float red[16][256];
  const int Ho = H >> 1, Wo = W >> 1;
  const int cv = threadIdx.x & 7;
  const int c0 = cv * 8;
  const F8 sc = load8f(scale + c0), sh = load8f(shift + c0);
  F8 k0, k1;
  if (APPLY) {
    const F8 mu = load8f(mean + c0), is = load8f(invstd + c0), dg = load8f(dgamma + c0), db = load8f(dbeta + c0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float c2 = sc.v[k] * dg.v[k] * inv_rows * is.v[k];
      k1.v[k] = -c2;
      k0.v[k] = fmaf(c2, mu.v[k], -sc.v[k] * db.v[k] * inv_rows);
    }
  }
  float a_dy[8], a_dyx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a_dy[k] = a_dyx[k] = 0.f;
  const int64_t patches = static_cast<int64_t>(N) * Ho * Wo;
  for (int64_t pi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 3; pi < patches;
       pi += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 3) {
    int a, b, n;
    if (lg_wo >= 0) {   // Wo, Ho powers of two (every size the bf16 path accepts): shifts instead of 64-bit divisions
      const uint32_t u = static_cast<uint32_t>(pi);
      b = u & (Wo - 1);
      a = (u >> lg_wo) & (Ho - 1);
      n = u >> (lg_wo + lg_ho);
    } else {
      b = static_cast<int>(pi % Wo);
      const int64_t t = pi / Wo;
      a = static_cast<int>(t % Ho);
      n = static_cast<int>(t / Ho);
    }
    // the four windows of this patch: w[dy][dx] = window (a + dy, b + dx)
    F8 wd[2][2];
    uint2 wi[2][2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        const bool ok = (a + dy < Ho) && (b + dxx < Wo);
        if (ok) {
          const int64_t j = ((static_cast<int64_t>(n) * Ho + a + dy) * Wo + b + dxx) * cvec + cv;
          wi[dy][dxx] = __ldg(idx + j);
          wd[dy][dxx] = unpack8(__ldg(dpool + j));
        } else {
          wi[dy][dxx] = make_uint2(0xffffffffu, 0xffffffffu);   // code 255 never matches
#pragma unroll
          for (int k = 0; k < 8; ++k) wd[dy][dxx].v[k] = 0.f;
        }
      }
    auto code_of = [](const uint2& id, int k) { return ((k < 4 ? (id.x >> (8 * k)) : (id.y >> (8 * (k - 4)))) & 0xffu); };
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const int64_t i = ((static_cast<int64_t>(n) * H + 2 * a + py) * W + 2 * b + px) * cvec + cv;
        const F8 xv = unpack8(ldg_stream(raw + i));
        F8 g;
#pragma unroll
        for (int k = 0; k < 8; ++k) g.v[k] = 0.f;
        // windows containing pixel (2a+py, 2b+px): rows {a} (py == 0) or {a, a+1} (py == 1); tap kh = 1 + py - 2*dy
#pragma unroll
        for (int dy = 0; dy <= py; ++dy)
#pragma unroll
          for (int dxx = 0; dxx <= px; ++dxx) {
            const uint32_t code = static_cast<uint32_t>((1 + py - 2 * dy) * 3 + (1 + px - 2 * dxx));
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (code_of(wi[dy][dxx], k) == code) g.v[k] += wd[dy][dxx].v[k];
          }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (!(fmaf(xv.v[k], sc.v[k], sh.v[k]) > 0.f)) g.v[k] = 0.f;
        if (APPLY) {
          F8 r;
#pragma unroll
          for (int k = 0; k < 8; ++k) r.v[k] = fmaf(sc.v[k], g.v[k], fmaf(k1.v[k], xv.v[k], k0.v[k]));
          dx[i] = pack8(r);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a_dy[k] += g.v[k];
            a_dyx[k] = fmaf(g.v[k], xv.v[k], a_dyx[k]);
          }
        }
      }
  }
  if (!APPLY) {
    const F8 mu = load8f(mean + c0), is = load8f(invstd + c0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[k][threadIdx.x] = a_dy[k];
      red[8 + k][threadIdx.x] = (a_dyx[k] - mu.v[k] * a_dy[k]) * is.v[k];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float s0 = 0.f, s1 = 0.f;
        for (int r = 0; r < 32; ++r) {
          s0 += red[k][r * 8 + cv];
          s1 += red[8 + k][r * 8 + cv];
        }
        partial[(static_cast<size_t>(blockIdx.x) * 2 + 0) * 64 + c0 + k] = s0;
        partial[(static_cast<size_t>(blockIdx.x) * 2 + 1) * 64 + c0 + k] = s1;
      }
    }
  }
}

void stem_pool_bn_backward(const bf16* dpool, const uint8_t* idx, const bf16* raw, const float* scale,
                           const float* shift, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                           bf16* dx, int N, int H, int W, int C, float* scratch, cudaStream_t s) {
  ARGUS_CHECK(C == 64, "the fused stem backward is written for 64 channels");
  ARGUS_CHECK(H % 2 == 0 && W % 2 == 0, "stem output must have even height and width");
  ARGUS_CHECK(scratch != nullptr, "stem_pool_bn_backward needs a scratch buffer");
  const int64_t rows = static_cast<int64_t>(N) * H * W;
  const int64_t threads = rows / 4 * 8;
  const float inv_rows = static_cast<float>(1.0 / static_cast<double>(rows));
  int lg_wo = log2_exact(W / 2), lg_ho = log2_exact(H / 2);
  if (lg_wo < 0 || lg_ho < 0 || rows / 4 >= (1LL << 32)) lg_wo = lg_ho = -1;
  auto DP = reinterpret_cast<const uint4*>(dpool);
  auto ID = reinterpret_cast<const uint2*>(idx);
  auto RW = reinterpret_cast<const uint4*>(raw);
  {
    ProfileScope prof("bn_bwd_reduce", s, 0, static_cast<double>(rows) * C * (2.0 + 0.75));
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((threads + 255) / 256, 4LL * num_sms())));
    launch_kernel(stem_pool_bn_bwd_kernel<0>, grid, 256, 0, s, DP, ID, RW, scale, shift, mean, invstd, nullptr, nullptr, scratch,
                                                    nullptr, N, H, W, inv_rows, lg_wo, lg_ho);
    ARGUS_CUDA(cudaGetLastError());
    launch_kernel(bn_bwd_finalize_kernel, (C + 7) / 8, 256, 0, s, scratch, grid, dgamma, dbeta, C, nullptr, nullptr);
    ARGUS_CUDA(cudaGetLastError());
  }
  {
    ProfileScope prof("bn_bwd_apply", s, 0, static_cast<double>(rows) * C * (4.0 + 0.75));
    const int grid = grid_for(threads, 256);
    launch_kernel(stem_pool_bn_bwd_kernel<1>, grid, 256, 0, s, DP, ID, RW, scale, shift, mean, invstd, dgamma, dbeta, nullptr,
                                                    reinterpret_cast<uint4*>(dx), N, H, W, inv_rows, lg_wo, lg_ho);
    ARGUS_CUDA(cudaGetLastError());
  }
}

__global__ void avgpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int HW, int cvec) {
  pdl_prologue();
  const int64_t total = static_cast<int64_t>(N) * cvec;
  const float inv = 1.0f / HW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % cvec);
    const int n = static_cast<int>(i / cvec);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    // eight independent 16-byte loads in flight, added in pixel order (same bits as a plain loop, an eighth of the
    // latency chain: the batch-1 inference forward spent 23 us here)
    const uint4* src = x + static_cast<int64_t>(n) * HW * cvec + cv;
    int p = 0;
    for (; p + 8 <= HW; p += 8) {
      uint4 u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = ldg_stream(src + static_cast<int64_t>(p + j) * cvec);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const F8 v = unpack8(u[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v.v[k];
      }
    }
    for (; p < HW; ++p) {
      const F8 v = unpack8(ldg_stream(src + static_cast<int64_t>(p) * cvec));
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v.v[k];
    }
    F8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = acc[k] * inv;
    y[i] = pack8(o);
  }
}
void avgpool_fwd(const bf16* x, bf16* y, int N, int HW, int C, cudaStream_t s) {
  ProfileScope prof("avgpool", s, 0, static_cast<double>(N) * (HW + 1) * C * 2);
  const int64_t total = static_cast<int64_t>(N) * (C / 8);
  launch_kernel(avgpool_fwd_kernel, grid_for(total, 128), 128, 0, s, reinterpret_cast<const uint4*>(x),
                                                          reinterpret_cast<uint4*>(y), N, HW, C / 8);
  ARGUS_CUDA(cudaGetLastError());
}
__global__ void avgpool_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, const uint8_t* __restrict__ bits,
                                   int N, int HW, int cvec) {
  pdl_prologue();
  const int64_t total = static_cast<int64_t>(N) * HW * cvec;
  const float inv = 1.0f / HW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % cvec);
    const int n = static_cast<int>(i / (static_cast<int64_t>(HW) * cvec));
    F8 v = unpack8(__ldg(dy + static_cast<int64_t>(n) * cvec + cv));
    const uint32_t b = bits != nullptr ? bits[i] : 0xffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = ((b >> k) & 1u) ? v.v[k] * inv : 0.f;
    dx[i] = pack8(v);
  }
}
void avgpool_bwd(const bf16* dy, bf16* dx, const uint8_t* relu_bits, int N, int HW, int C, cudaStream_t s) {
  ProfileScope prof("avgpool", s, 0, static_cast<double>(N) * (HW + 1) * C * 2);
  const int64_t total = static_cast<int64_t>(N) * HW * (C / 8);
  launch_kernel(avgpool_bwd_kernel, grid_for(total, 256), 256, 0, s, reinterpret_cast<const uint4*>(dy),
                                                          reinterpret_cast<uint4*>(dx), relu_bits, N, HW, C / 8);
  ARGUS_CUDA(cudaGetLastError());
}

}  // namespace argus
