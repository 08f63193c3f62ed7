// Fused GPU image augmentation (reference: the kornia chain built in /root/reference/argus/data.py:41-103 and applied
// per sample at data.py:213-225 — planckian jitter -> colour jiggle -> gaussian blur -> motion blur -> plasma shadow).
//
// One kernel per batch: uint8 HWC (or fp32 NCHW) in -> /255 -> colour ops in registers -> 5-tap separable gaussian
// and 3x3 motion kernel from a shared-memory halo tile -> plasma shadow -> clamp -> bf16 space-to-depth layout that
// the stem convolution's TMA reads directly (or fp32 NCHW for the drop-in Augmentation.forward()).
// Parameters are a pure function of (seed, step, image, field) through a splitmix64 hash, so the numpy oracle
// (oracle/augment.py) reproduces them bit for bit. The arithmetic spec is documented there.
#include "kernels.h"
#include "ptx.cuh"
#include "runtime.h"

namespace argus {

constexpr int kAugParams = 24;

__constant__ float c_planck_r[25] = {1.67361629f, 1.48101866f, 1.35384262f, 1.26161098f, 1.19069767f, 1.13347971f,
                                     1.08616924f, 1.04604697f, 1.01147437f, 0.980993509f, 0.954290628f, 0.930232584f,
                                     0.908562839f, 0.889056146f, 0.871265829f, 0.855178356f, 0.840171874f,
                                     0.826245725f, 0.813359082f, 0.801546395f, 0.790547788f, 0.780236304f,
                                     0.770519972f, 0.761659145f, 0.752847612f};
__constant__ float c_planck_b[25] = {0.00322660711f, 0.392596096f, 0.574794173f, 0.7076509f, 0.813289046f,
                                     0.900768399f, 0.97469461f, 1.03866208f, 1.09395969f, 1.14233255f, 1.18520916f,
                                     1.22329891f, 1.25731492f, 1.28789508f, 1.31549537f, 1.34056723f, 1.36326528f,
                                     1.38402057f, 1.40292096f, 1.42031789f, 1.43673468f, 1.4513427f, 1.46540606f,
                                     1.47818613f, 1.49000645f};
// the 24 permutations of (brightness, contrast, saturation, hue), lexicographic, 2 bits per slot
__constant__ uint8_t c_orders[24] = {0x1B, 0x1E, 0x27, 0x2D, 0x36, 0x39, 0x4B, 0x4E, 0x63, 0x6C, 0x72, 0x78,
                                     0x87, 0x8D, 0x93, 0x9C, 0xB1, 0xB4, 0xC6, 0xC9, 0xD2, 0xD8, 0xE1, 0xE4};

__host__ __device__ __forceinline__ uint64_t hash_u64(uint64_t seed, uint64_t step, uint64_t image, uint64_t field) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + step * 0xBF58476D1CE4E5B9ull + image * 0x94D049BB133111EBull + field;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t step, uint64_t image, uint64_t field) {
  return static_cast<float>(hash_u64(seed, step, image, field) >> 40) * 5.9604644775390625e-08f;  // 2^-24
}
__device__ __forceinline__ float lerp_rn(float u, float lo, float span) { return __fadd_rn(lo, __fmul_rn(u, span)); }

// ------------------------------------------------------------------------------------------------------------
// parameter sampling
// ------------------------------------------------------------------------------------------------------------
__global__ void aug_params_kernel(float* __restrict__ params, int n_images, int n_cams, uint64_t seed, uint64_t step,
                                  AugConfig cfg) {
  pdl_prologue();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_images) return;
  float* P = params + static_cast<size_t>(n) * kAugParams;
  const uint64_t img = n;
  const uint64_t pair = static_cast<uint64_t>(n / n_cams) + (1ull << 32);
  auto U = [&](uint64_t who, uint64_t field) { return hash_uniform(seed, step, who, field); };
  // planckian jitter
  float pr = 1.f, pb = 1.f;
  if (cfg.planckian_jitter) {
    const bool apply = U(img, 0) < 0.5f;
    int idx = static_cast<int>(__fmul_rn(U(img, 1), 25.f));
    idx = idx > 24 ? 24 : idx;
    if (apply) { pr = c_planck_r[idx]; pb = c_planck_b[idx]; }
  }
  P[0] = pr; P[1] = pb;
  // colour jiggle (per pair)
  if (cfg.color_jiggle) {
    P[2] = __fadd_rn(lerp_rn(U(pair, 2), cfg.brightness_lo, cfg.brightness_span), -1.f);
    P[3] = lerp_rn(U(pair, 3), cfg.contrast_lo, cfg.contrast_span);
    P[4] = lerp_rn(U(pair, 4), cfg.saturation_lo, cfg.saturation_span);
    P[5] = __fmul_rn(lerp_rn(U(pair, 5), cfg.hue_lo, cfg.hue_span), 6.28318530717958647692f);
    int o = static_cast<int>(__fmul_rn(U(pair, 6), 24.f));
    P[6] = static_cast<float>(o > 23 ? 23 : o);
  } else {
    P[2] = 0.f; P[3] = 1.f; P[4] = 1.f; P[5] = 0.f; P[6] = -1.f;
  }
  // gaussian blur
  P[7] = (cfg.blur && U(img, 7) < 0.5f) ? lerp_rn(U(img, 8), 3.f, 5.f) : 0.f;
  // motion blur kernel
  float k[9] = {0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f};
  if (cfg.motion_blur && U(img, 9) < 0.7f) {
    const float angle = __fmul_rn(lerp_rn(U(img, 10), -35.f, 70.f), 0.01745329251994329577f);
    float d = lerp_rn(U(img, 11), -0.5f, 1.f);
    d = fminf(fmaxf(d, -1.f), 1.f);
    d = __fmul_rn(__fadd_rn(d, 1.f), 0.5f);
    const float row[3] = {d, 0.5f, __fadd_rn(1.f, -d)};
    const float ca = cosf(angle), sa = sinf(angle);
    float sum = 0.f;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const float x = static_cast<float>(j - 1), y = static_cast<float>(i - 1);
        const int xs = static_cast<int>(rintf(__fadd_rn(__fmul_rn(ca, x), -__fmul_rn(sa, y))));
        const int ys = static_cast<int>(rintf(__fadd_rn(__fmul_rn(sa, x), __fmul_rn(ca, y))));
        const float v = (ys == 0 && xs >= -1 && xs <= 1) ? row[xs + 1] : 0.f;
        k[i * 3 + j] = v;
        sum = __fadd_rn(sum, v);
      }
    for (int i = 0; i < 9; ++i) k[i] = __fdiv_rn(k[i], sum);
  }
  for (int i = 0; i < 9; ++i) P[8 + i] = k[i];
  // plasma shadow
  if (cfg.plasma_shadow) {
    P[17] = lerp_rn(U(img, 12), 0.1f, 0.30000000000000004f);
    P[18] = lerp_rn(U(img, 13), -0.6f, 0.6f);
    P[19] = lerp_rn(U(img, 14), 0.f, 0.5f);
  } else {
    P[17] = 0.25f; P[18] = 0.f; P[19] = 0.f;
  }
  P[20] = U(img, 15);
  P[21] = 0.f; P[22] = 1.f; P[23] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// plasma fractal: 6 octaves of smooth value noise, octave l has 2^(l+1) cells per side, amplitude roughness^l
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lattice(uint64_t seed_bits, int octave, int iy, int ix) {
  const uint64_t key = (seed_bits << 40) | (static_cast<uint64_t>(octave) << 32) | (static_cast<uint64_t>(iy) << 16) |
                       static_cast<uint64_t>(ix);
  return hash_uniform(0x504C41534D41ull, 0, key, 0);
}
template <typename Lattice>
__device__ __forceinline__ float plasma_eval(int y, int x, float inv_h, float inv_w, float roughness, Lattice lat) {
  const float ys = __fmul_rn(static_cast<float>(y) + 0.5f, inv_h);
  const float xs = __fmul_rn(static_cast<float>(x) + 0.5f, inv_w);
  float field = 0.f, amp = 1.f;
#pragma unroll
  for (int l = 0; l < 6; ++l) {
    const float cells = static_cast<float>(2 << l);
    const float fy = __fmul_rn(ys, cells), fx = __fmul_rn(xs, cells);
    const float flo_y = floorf(fy), flo_x = floorf(fx);
    const int iy = static_cast<int>(flo_y), ix = static_cast<int>(flo_x);
    float ty = __fadd_rn(fy, -flo_y), tx = __fadd_rn(fx, -flo_x);
    ty = __fmul_rn(__fmul_rn(ty, ty), __fadd_rn(3.f, -__fmul_rn(2.f, ty)));
    tx = __fmul_rn(__fmul_rn(tx, tx), __fadd_rn(3.f, -__fmul_rn(2.f, tx)));
    const float v00 = lat(l, iy, ix), v01 = lat(l, iy, ix + 1);
    const float v10 = lat(l, iy + 1, ix), v11 = lat(l, iy + 1, ix + 1);
    const float top = __fadd_rn(v00, __fmul_rn(__fadd_rn(v01, -v00), tx));
    const float bot = __fadd_rn(v10, __fmul_rn(__fadd_rn(v11, -v10), tx));
    field = __fadd_rn(field, __fmul_rn(amp, __fadd_rn(top, __fmul_rn(__fadd_rn(bot, -top), ty))));
    amp = __fmul_rn(amp, roughness);
  }
  return field;
}
struct HashLattice {
  uint64_t seed_bits;
  __device__ __forceinline__ float operator()(int l, int iy, int ix) const { return lattice(seed_bits, l, iy, ix); }
};
// whole-image lattice in shared memory: octave l holds (2^(l+1) + 1)^2 values at offset c_lat_off[l]
constexpr int kLatTotal = 9 + 25 + 81 + 289 + 1089 + 4225;  // 5718
__constant__ int c_lat_off[6] = {0, 9, 34, 115, 404, 1493};
struct FullTable {
  const float* t;
  __device__ __forceinline__ float operator()(int l, int iy, int ix) const {
    return t[c_lat_off[l] + iy * ((2 << l) + 1) + ix];
  }
};
// per-tile lattice window: octave l covers lattice rows [iy0[l], iy0[l] + 17) x cols [ix0[l], ix0[l] + 17)
constexpr int kWin = 17;
struct TileTable {
  const float* t;
  const int* iy0;
  const int* ix0;
  __device__ __forceinline__ float operator()(int l, int iy, int ix) const {
    return t[(l * kWin + (iy - iy0[l])) * kWin + (ix - ix0[l])];
  }
};

// one block per image: min / max of the un-normalised field -> params[21], params[22]
__global__ void __launch_bounds__(256) plasma_minmax_kernel(float* __restrict__ params, int H, int W) {
  pdl_prologue();
  __shared__ float s_lo[8], s_hi[8];
  __shared__ float s_lat[kLatTotal];
  float* P = params + static_cast<size_t>(blockIdx.x) * kAugParams;
  const float roughness = P[17];
  const uint64_t seed_bits = static_cast<uint64_t>(rintf(P[20] * 16777216.f));
  const float inv_h = 1.f / H, inv_w = 1.f / W;
  float lo = INFINITY, hi = -INFINITY;
  if (P[18] != 0.f) {
    for (int l = 0; l < 6; ++l) {
      const int side = (2 << l) + 1;
      for (int i = threadIdx.x; i < side * side; i += blockDim.x)
        s_lat[c_lat_off[l] + i] = lattice(seed_bits, l, i / side, i % side);
    }
    __syncthreads();
    FullTable lat{s_lat};
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
      const float f = plasma_eval(i / W, i % W, inv_h, inv_w, roughness, lat);
      lo = fminf(lo, f);
      hi = fmaxf(hi, f);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { lo = fminf(lo, s_lo[i]); hi = fmaxf(hi, s_hi[i]); }
    lo = fminf(lo, s_lo[0]); hi = fmaxf(hi, s_hi[0]);
    P[21] = lo;
    P[22] = hi;
  }
}

// ------------------------------------------------------------------------------------------------------------
// colour operations (kornia.enhance semantics, see oracle/augment.py)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
  const float mx = fmaxf(fmaxf(r, g), b), mn = fminf(fminf(r, g), b);
  v = mx;
  const float delta = mx - mn;
  s = __fdiv_rn(delta, mx + 1e-8f);
  const float dz = delta == 0.f ? 1.f : delta;
  const float rc = mx - r, gc = mx - g, bc = mx - b;
  float hh = (r == mx) ? (bc - gc) : ((g == mx) ? __fadd_rn(rc - bc, __fmul_rn(2.f, dz)) : __fadd_rn(gc - rc, __fmul_rn(4.f, dz)));
  hh = __fdiv_rn(hh, dz);
  hh = __fdiv_rn(hh, 6.f);
  hh = hh - floorf(hh);
  h = __fmul_rn(6.28318530717958647692f, hh);
}
__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float& r, float& g, float& b) {
  const float h6 = __fmul_rn(__fdiv_rn(h, 6.28318530717958647692f), 6.f);
  const float fl = floorf(h6);
  int hi = static_cast<int>(fl) % 6;
  if (hi < 0) hi += 6;
  const float f = h6 - fl;
  const float p = __fmul_rn(v, 1.f - s);
  const float q = __fmul_rn(v, __fadd_rn(1.f, -__fmul_rn(f, s)));
  const float t = __fmul_rn(v, __fadd_rn(1.f, -__fmul_rn(1.f - f, s)));
  switch (hi) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}
__device__ __forceinline__ void color_ops(float& r, float& g, float& b, const float* __restrict__ P) {
  r = fminf(__fmul_rn(r, P[0]), 1.f);
  b = fminf(__fmul_rn(b, P[1]), 1.f);
  const int order = static_cast<int>(P[6]);
  if (order < 0) return;
  const uint32_t code = c_orders[order];
#pragma unroll
  for (int slot = 0; slot < 4; ++slot) {
    const int op = (code >> (6 - 2 * slot)) & 3;
    if (op == 0) {
      r = clamp01(r + P[2]); g = clamp01(g + P[2]); b = clamp01(b + P[2]);
    } else if (op == 1) {
      r = clamp01(__fmul_rn(r, P[3])); g = clamp01(__fmul_rn(g, P[3])); b = clamp01(__fmul_rn(b, P[3]));
    } else if (op == 2) {
      float h, s, v;
      rgb_to_hsv(r, g, b, h, s, v);
      hsv_to_rgb(h, clamp01(__fmul_rn(s, P[4])), v, r, g, b);
    } else {
      float h, s, v;
      rgb_to_hsv(r, g, b, h, s, v);
      h = h + P[5];
      h = __fadd_rn(h, -__fmul_rn(6.28318530717958647692f, floorf(__fdiv_rn(h, 6.28318530717958647692f))));
      hsv_to_rgb(h, s, v, r, g, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// the fused kernel: one 32x32 output tile of one image per block
// ------------------------------------------------------------------------------------------------------------
constexpr int kTile = 32;
constexpr int kHalo = 3;
constexpr int kIn = kTile + 2 * kHalo;   // 38
constexpr int kMid = kTile + 2;          // 34

template <bool IN_U8, bool OUT_S2D>
__global__ void __launch_bounds__(256)
augment_kernel(const void* __restrict__ in, void* __restrict__ out, const float* __restrict__ params, int H, int W,
               int apply) {
  pdl_prologue();
  __shared__ float sA[3][kIn][kIn + 1];    // colour-jittered input with halo; later reused for the blurred tile
  __shared__ float sB[3][kIn][kMid + 1];   // after the horizontal gaussian pass
  __shared__ float sP[kAugParams];
  __shared__ float sLat[6 * kWin * kWin];
  __shared__ int sIy0[6], sIx0[6];
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
  // (apply == 0 is plain u8 -> bf16 staging: the parameter table may be NULL and is never read)
  if (threadIdx.x < kAugParams) sP[threadIdx.x] = apply ? params[static_cast<size_t>(n) * kAugParams + threadIdx.x] : 0.f;
  __syncthreads();
  const float sigma = apply ? sP[7] : 0.f;
  // plasma lattice window of this tile (a 32-pixel tile spans at most 16 cells of the finest octave when the image
  // side is >= 128; smaller images fall back to hashing per pixel)
  const bool use_table = (H >= 128 && W >= 128);
  const uint64_t seed_bits = static_cast<uint64_t>(rintf(sP[20] * 16777216.f));
  if (apply && sP[18] != 0.f && use_table) {
    if (threadIdx.x < 6) {
      const int l = threadIdx.x;
      const float cells = static_cast<float>(2 << l);
      sIy0[l] = static_cast<int>(floorf(__fmul_rn(__fmul_rn(static_cast<float>(y0) + 0.5f, 1.f / H), cells)));
      sIx0[l] = static_cast<int>(floorf(__fmul_rn(__fmul_rn(static_cast<float>(x0) + 0.5f, 1.f / W), cells)));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 6 * kWin * kWin; i += 256) {
      const int l = i / (kWin * kWin);
      const int rem = i - l * kWin * kWin;
      const int ry = rem / kWin, rx = rem - ry * kWin;
      const int iy = sIy0[l] + ry, ix = sIx0[l] + rx;
      const int side = (2 << l) + 1;
      sLat[i] = (iy < side && ix < side) ? lattice(seed_bits, l, iy, ix) : 0.f;
    }
  }

  // ---- stage 1: load (+ reflect), /255, colour ops
  for (int i = threadIdx.x; i < kIn * kIn; i += 256) {
    const int ty = i / kIn, tx = i - ty * kIn;
    int gy = y0 - kHalo + ty, gx = x0 - kHalo + tx;
    gy = gy < 0 ? -gy : (gy >= H ? 2 * (H - 1) - gy : gy);
    gx = gx < 0 ? -gx : (gx >= W ? 2 * (W - 1) - gx : gx);
    float r, g, b;
    if (IN_U8) {
      const uint8_t* p = static_cast<const uint8_t*>(in) + (static_cast<size_t>(n) * H * W + static_cast<size_t>(gy) * W + gx) * 3;
      r = __fmul_rn(static_cast<float>(p[0]), 1.0f / 255.0f);
      g = __fmul_rn(static_cast<float>(p[1]), 1.0f / 255.0f);
      b = __fmul_rn(static_cast<float>(p[2]), 1.0f / 255.0f);
    } else {
      const float* p = static_cast<const float*>(in) + static_cast<size_t>(n) * 3 * H * W + static_cast<size_t>(gy) * W + gx;
      r = p[0]; g = p[static_cast<size_t>(H) * W]; b = p[2 * static_cast<size_t>(H) * W];
    }
    if (apply) color_ops(r, g, b, sP);
    sA[0][ty][tx] = r; sA[1][ty][tx] = g; sA[2][ty][tx] = b;
  }
  __syncthreads();

  // ---- stage 2: separable 5-tap gaussian (reflect already materialised in the halo)
  if (sigma > 0.f) {
    float k[5];
    float ks = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float d = static_cast<float>(i - 2);
      k[i] = expf(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(__fmul_rn(2.f, sigma), sigma)));
      ks = __fadd_rn(ks, k[i]);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) k[i] = __fdiv_rn(k[i], ks);
    for (int i = threadIdx.x; i < 3 * kIn * kMid; i += 256) {
      const int c = i / (kIn * kMid);
      const int rem = i - c * kIn * kMid;
      const int ty = rem / kMid, ox = rem - ty * kMid;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 5; ++t) acc = __fadd_rn(acc, __fmul_rn(k[t], sA[c][ty][ox + t]));
      sB[c][ty][ox] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * kMid * kMid; i += 256) {
      const int c = i / (kMid * kMid);
      const int rem = i - c * kMid * kMid;
      const int oy = rem / kMid, ox = rem - oy * kMid;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 5; ++t) acc = __fadd_rn(acc, __fmul_rn(k[t], sB[c][oy + t][ox]));
      sA[c][oy][ox] = acc;   // (oy, ox) <-> pixel (y0 - 1 + oy, x0 - 1 + ox)
    }
  }
  __syncthreads();
  // motion blur uses a zero ('constant') border: blank the window positions that fall outside the image
  // (without the gaussian pass the window is simply the centre of the halo tile: index offset 2)
  const int off = sigma > 0.f ? 0 : 2;
  for (int i = threadIdx.x; i < kMid * kMid; i += 256) {
    const int oy = i / kMid, ox = i - oy * kMid;
    const int gy = y0 - 1 + oy, gx = x0 - 1 + ox;
    if (gy < 0 || gy >= H || gx < 0 || gx >= W) {
      sA[0][oy + off][ox + off] = 0.f; sA[1][oy + off][ox + off] = 0.f; sA[2][oy + off][ox + off] = 0.f;
    }
  }
  __syncthreads();

  // ---- stage 3: 3x3 motion kernel, plasma shadow, clamp, store. Thread = one 2x2 pixel quad (one s2d pixel).
  const int sy = threadIdx.x >> 4, sx = threadIdx.x & 15;
  const float inv_h = 1.f / H, inv_w = 1.f / W;
  const float roughness = sP[17], intensity = apply ? sP[18] : 0.f, quantity = sP[19];
  const float p_lo = sP[21], p_den = fmaxf(sP[22] - sP[21], 1e-12f);
  float v[2][2][3];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int py = 2 * sy + a, px = 2 * sx + b;   // position inside the tile; window index = +1
      float shade = 0.f;
      if (intensity != 0.f) {
        const float f = use_table
                            ? plasma_eval(y0 + py, x0 + px, inv_h, inv_w, roughness, TileTable{sLat, sIy0, sIx0})
                            : plasma_eval(y0 + py, x0 + px, inv_h, inv_w, roughness, HashLattice{seed_bits});
        const float fn = __fdiv_rn(f - p_lo, p_den);
        shade = fn < quantity ? intensity : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float acc;
        if (apply) {
          acc = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = __fadd_rn(acc, __fmul_rn(sP[8 + i * 3 + j], sA[c][py + i + off][px + j + off]));
        } else {
          acc = sA[c][py + 1 + off][px + 1 + off];
        }
        v[a][b][c] = clamp01(acc + shade);
      }
    }
  if (OUT_S2D) {
    // [n][H/2][W/2 + 4][16] bf16; channel = (a*2 + b)*3 + c
    const int Hs = H >> 1, Ws = W >> 1, Wp = Ws + 4;
    const int is = (y0 >> 1) + sy, js = (x0 >> 1) + sx;
    uint4* row = reinterpret_cast<uint4*>(out) + (static_cast<size_t>(n) * Hs + is) * Wp * 2;
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0][0][0], v[0][0][1]); o0.y = pack_bf16x2(v[0][0][2], v[0][1][0]);
    o0.z = pack_bf16x2(v[0][1][1], v[0][1][2]); o0.w = pack_bf16x2(v[1][0][0], v[1][0][1]);
    o1.x = pack_bf16x2(v[1][0][2], v[1][1][0]); o1.y = pack_bf16x2(v[1][1][1], v[1][1][2]);
    o1.z = 0u; o1.w = 0u;
    row[(js + 2) * 2] = o0;
    row[(js + 2) * 2 + 1] = o1;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (js == 0) { row[0] = z; row[1] = z; row[2] = z; row[3] = z; }
    if (js == Ws - 1) { row[(Ws + 2) * 2] = z; row[(Ws + 2) * 2 + 1] = z; row[(Ws + 3) * 2] = z; row[(Ws + 3) * 2 + 1] = z; }
  } else {
    float* o = static_cast<float*>(out) + static_cast<size_t>(n) * 3 * H * W;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        float2 w2 = make_float2(v[a][0][c], v[a][1][c]);
        *reinterpret_cast<float2*>(o + (static_cast<size_t>(c) * H + y0 + 2 * sy + a) * W + x0 + 2 * sx) = w2;
      }
  }
}

// ------------------------------------------------------------------------------------------------------------
// spaghetti arcs (reference: argus/utils.py:252-275 draw_spaghetti, applied to the decoded image at
// argus/data.py:212-215 before the kornia chain; PIL on the loader's CPU workers there). Rasterisation rule and
// sampling: oracle/augment.py (calibrated against Pillow's ImageDraw.arc, IoU 0.92; bit-exact against the oracle).
// ------------------------------------------------------------------------------------------------------------
constexpr int kArcFields = 10;   // cx, cy, rx, ry, cos0, sin0, cos1, sin1, width, sweep_deg
constexpr uint64_t kArcFieldBase = 1000;

__global__ void spaghetti_params_kernel(float* __restrict__ arcs, int n_images, int n_arcs, int H, int W, uint64_t seed,
                                        uint64_t step) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_images * n_arcs) return;
  const uint64_t img = t / n_arcs, a = t % n_arcs;
  auto U = [&](uint64_t k) { return hash_uniform(seed, step, img, kArcFieldBase + a * 8 + k); };
  auto randint = [](float u, int lo, int hi) {
    const int v = lo + static_cast<int>(floorf(__fmul_rn(u, static_cast<float>(hi - lo))));
    return v < hi - 1 ? v : hi - 1;
  };
  const int x0 = randint(U(0), 0, W), y0 = randint(U(1), 0, H);
  const int x1 = randint(U(2), x0, W), y1 = randint(U(3), y0, H);
  const int a0 = randint(U(4), 0, 360), a1 = randint(U(5), 0, 360);
  float* A = arcs + static_cast<size_t>(t) * kArcFields;
  A[0] = __fmul_rn(static_cast<float>(x0 + x1), 0.5f);
  A[1] = __fmul_rn(static_cast<float>(y0 + y1), 0.5f);
  A[2] = __fadd_rn(__fmul_rn(static_cast<float>(x1 - x0), 0.5f), 0.5f);
  A[3] = __fadd_rn(__fmul_rn(static_cast<float>(y1 - y0), 0.5f), 0.5f);
  const double r0 = static_cast<double>(a0) * 0.017453292519943295, r1 = static_cast<double>(a1) * 0.017453292519943295;
  A[4] = static_cast<float>(cos(r0)); A[5] = static_cast<float>(sin(r0));
  A[6] = static_cast<float>(cos(r1)); A[7] = static_cast<float>(sin(r1));
  A[8] = floorf(__fadd_rn(1.0f, __fmul_rn(U(6), 4.0f)));
  A[9] = static_cast<float>(((a1 - a0) % 360 + 360) % 360);
}

// thread = one pixel; black where any arc covers it (n_arcs <= 16 cached in shared memory per image row block)
__global__ void __launch_bounds__(256)
spaghetti_draw_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const float* __restrict__ arcs,
                      int n_arcs, int H, int W) {
  pdl_prologue();
  __shared__ float sArc[16 * kArcFields];
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < n_arcs * kArcFields; i += blockDim.x)
    sArc[i] = arcs[static_cast<size_t>(n) * n_arcs * kArcFields + i];
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix - y * W;
  bool hit = false;
  for (int k = 0; k < n_arcs; ++k) {
    const float* A = sArc + k * kArcFields;
    const float dx = __fadd_rn(static_cast<float>(x), -A[0]), dy = __fadd_rn(static_cast<float>(y), -A[1]);
    const float u = __fdiv_rn(dx, A[2]), v = __fdiv_rn(dy, A[3]);
    if (!(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)) <= 1.0f)) continue;
    const float irx = __fadd_rn(A[2], -A[8]), iry = __fadd_rn(A[3], -A[8]);
    if (irx > 0.f && iry > 0.f) {
      const float ui = __fdiv_rn(dx, irx), vi = __fdiv_rn(dy, iry);
      if (__fadd_rn(__fmul_rn(ui, ui), __fmul_rn(vi, vi)) < 1.0f) continue;
    }
    const float a = __fadd_rn(__fmul_rn(A[4], v), -__fmul_rn(A[5], u));
    const float b = __fadd_rn(__fmul_rn(A[7], u), -__fmul_rn(A[6], v));
    const bool sector = (A[9] <= 180.f) ? (a >= 0.f && b >= 0.f) : !(a < 0.f && b < 0.f);
    if (sector) { hit = true; break; }
  }
  const size_t o = (static_cast<size_t>(n) * H * W + pix) * 3;
  out[o] = hit ? 0 : in[o];
  out[o + 1] = hit ? 0 : in[o + 1];
  out[o + 2] = hit ? 0 : in[o + 2];
}

void spaghetti_sample_params(float* arcs, int n_images, int n_arcs, int H, int W, uint64_t seed, uint64_t step,
                             cudaStream_t s) {
  ARGUS_CHECK(n_arcs >= 0 && n_arcs <= 16, "at most 16 arcs per image");
  if (n_images * n_arcs <= 0) return;
  ProfileScope prof("augment_params", s, 0, 40.0 * n_images * n_arcs);
  launch_kernel(spaghetti_params_kernel, (n_images * n_arcs + 127) / 128, 128, 0, s, arcs, n_images, n_arcs, H, W, seed, step);
  ARGUS_CUDA(cudaGetLastError());
}
void spaghetti_draw(const uint8_t* in, uint8_t* out, const float* arcs, int n_images, int n_arcs, int H, int W,
                    cudaStream_t s) {
  ARGUS_CHECK(n_arcs >= 0 && n_arcs <= 16, "at most 16 arcs per image");
  if (n_images <= 0) return;
  ProfileScope prof("augment", s, 0, 6.0 * n_images * H * W);
  dim3 grid((H * W + 255) / 256, n_images);
  launch_kernel(spaghetti_draw_kernel, grid, 256, 0, s, in, out, arcs, n_arcs, H, W);
  ARGUS_CUDA(cudaGetLastError());
}

void augment_sample_params(float* params, int n_images, int n_cams, uint64_t seed, uint64_t step, const AugConfig& cfg,
                           cudaStream_t s) {
  ProfileScope prof("augment_params", s, 0, 96.0 * n_images);
  if (n_images <= 0) return;
  launch_kernel(aug_params_kernel, (n_images + 127) / 128, 128, 0, s, params, n_images, n_cams, seed, step, cfg);
  ARGUS_CUDA(cudaGetLastError());
}

void augment_images(const void* in, bool in_u8, void* out, bool out_s2d, float* params, int n_images, int H, int W,
                    bool apply, cudaStream_t s) {
  ARGUS_CHECK(H % kTile == 0 && W % kTile == 0, "augmentation needs H and W to be multiples of 32");
  if (n_images <= 0) return;
  if (apply) {
    ProfileScope prof("augment_params", s, 0, 8.0 * n_images);
    launch_kernel(plasma_minmax_kernel, n_images, 256, 0, s, params, H, W);
    ARGUS_CUDA(cudaGetLastError());
  }
  // algorithmic bytes: u8 RGB in + bf16 RGB out = 9 B per pixel
  const double bytes = static_cast<double>(n_images) * H * W * (in_u8 ? 3.0 : 12.0) +
                       static_cast<double>(n_images) * H * W * (out_s2d ? 6.0 : 12.0);
  ProfileScope prof("augment", s, 0, bytes);
  dim3 grid(W / kTile, H / kTile, n_images);
  const int ap = apply ? 1 : 0;
  if (in_u8 && out_s2d) launch_kernel(augment_kernel<true, true>, grid, 256, 0, s, in, out, params, H, W, ap);
  else if (in_u8 && !out_s2d) launch_kernel(augment_kernel<true, false>, grid, 256, 0, s, in, out, params, H, W, ap);
  else if (!in_u8 && out_s2d) launch_kernel(augment_kernel<false, true>, grid, 256, 0, s, in, out, params, H, W, ap);
  else launch_kernel(augment_kernel<false, false>, grid, 256, 0, s, in, out, params, H, W, ap);
  ARGUS_CUDA(cudaGetLastError());
}

}  // namespace argus
