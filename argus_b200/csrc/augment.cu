// Fused GPU image augmentation (reference: the kornia chain built in /root/reference/argus/data.py:41-103 and applied
// per sample at data.py:213-225 -- [random erasing x2] -> planckian jitter -> colour jiggle -> gaussian blur -> motion
// blur -> plasma shadow -> [salt & pepper]) and the spaghetti arcs the dataset draws before it (argus/utils.py:252-275,
// data.py:212-215).
//
// Per batch:  arc_paint_kernel      one warp per arc: Pillow's ImageDraw.arc rasteriser, bit for bit (oracle/pil_arc.py)
//                                   -> 1 bit per pixel
//             plasma_mask_kernel    one block per image: kornia's diamond-square plasma fractal, built level by level in
//                                   shared memory, thresholded by shade_quantity -> 1 bit per pixel
//             augment_kernel        uint8 HWC (or fp32 NCHW) in -> arcs -> /255 -> erasing -> colour ops in registers ->
//                                   5-tap separable gaussian and 3x3 motion kernel from a shared-memory halo tile -> shadow
//                                   -> clamp -> salt & pepper -> bf16 space-to-depth layout that the stem convolution's TMA
//                                   reads directly (or fp32 NCHW for the drop-in Augmentation.forward()).
// Every random number is a pure function of (seed, step, image, field) through integer hashes, so the numpy oracle
// (oracle/augment.py) reproduces parameters, masks and arcs bit for bit; the arithmetic spec is documented there.
#include "kernels.h"
#include "ptx.cuh"
#include "runtime.h"

namespace argus {

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

// per-pixel random numbers (plasma fractal, salt & pepper): 32-bit hash of (image seed, level, y, x) -> 24-bit uniform
__device__ __forceinline__ float pixel_uniform(uint32_t seed32, uint32_t level, uint32_t y, uint32_t x) {
  uint32_t h = seed32 ^ (level * 0x9E3779B9u);
  h = (h ^ y) * 0x85EBCA6Bu;
  h ^= h >> 15;
  h = (h ^ x) * 0xC2B2AE35u;
  h ^= h >> 13;
  h *= 0x27D4EB2Fu;
  h ^= h >> 16;
  return static_cast<float>(h >> 8) * 5.9604644775390625e-08f;
}

// ------------------------------------------------------------------------------------------------------------
// parameter sampling (layout: oracle/augment.py::sample_params)
// ------------------------------------------------------------------------------------------------------------
// kornia RectangleEraseGenerator: area = U(scale) * H * W, aspect = h / w, h = round(sqrt(area * aspect)) clamped to
// [1, H], w likewise, top-left = floor(U * (size - extent + 1)).
__device__ void erase_rect(float u_area, float u_ratio_a, float u_ratio_b, float u_pick, float u_x, float u_y, float s_lo,
                           float s_span, bool two_sided, float r_lo, float r_span, float r2_span, int H, int W, float* out) {
  // (spans are passed as the float32 literals the oracle uses, not recomputed from the bounds in float32)
  const float area = __fmul_rn(lerp_rn(u_area, s_lo, s_span), static_cast<float>(H * W));
  float ratio;
  if (two_sided) {   // ratio range straddles 1: one sampler below, one above, picked by a coin
    const float r1 = lerp_rn(u_ratio_a, r_lo, r_span), r2 = lerp_rn(u_ratio_b, 1.f, r2_span);
    ratio = (rintf(u_pick) != 0.f) ? r1 : r2;
  } else {
    ratio = lerp_rn(u_ratio_a, r_lo, r_span);
  }
  float h = rintf(__fsqrt_rn(__fmul_rn(area, ratio)));
  float w = rintf(__fsqrt_rn(__fdiv_rn(area, ratio)));
  h = fminf(fmaxf(h, 1.f), static_cast<float>(H));
  w = fminf(fmaxf(w, 1.f), static_cast<float>(W));
  out[0] = floorf(__fmul_rn(u_x, __fadd_rn(__fadd_rn(static_cast<float>(W), -w), 1.f)));
  out[1] = floorf(__fmul_rn(u_y, __fadd_rn(__fadd_rn(static_cast<float>(H), -h), 1.f)));
  out[2] = w;
  out[3] = h;
}

__global__ void aug_params_kernel(float* __restrict__ params, int n_images, int n_cams, int H, int W, uint64_t seed,
                                  uint64_t step, AugConfig cfg) {
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
  // motion blur: kornia get_motion_kernel2d(3, angle, direction, 'nearest') = the row [d, .5, 1-d] through the centre,
  // rotated by warp_affine(align_corners=True, nearest, zeros): output (x, y) samples the source at R(angle) (x, y)
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
  P[21] = 0.f; P[22] = 0.f; P[23] = 0.f;
  // random erasing (two rectangles: black wide ones, white square-ish ones; data.py:52-64)
  for (int e = 0; e < 2; ++e) {
    float* R = P + 24 + 5 * e;
    R[0] = 0.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 0.f;
    if (cfg.random_erasing && U(img, 16 + 8 * e) < 0.5f) {
      R[0] = 1.f;
      const uint64_t f = 17 + 8 * e;
      // scale (0.02, 0.1), ratio (2, 3)  |  scale (0.02, 0.05), ratio (0.8, 1.2)
      if (e == 0) erase_rect(U(img, f), U(img, f + 1), U(img, f + 2), U(img, f + 3), U(img, f + 4), U(img, f + 5), 0.02f,
                             0.08f, false, 2.f, 1.f, 0.f, H, W, R + 1);
      else erase_rect(U(img, f), U(img, f + 1), U(img, f + 2), U(img, f + 3), U(img, f + 4), U(img, f + 5), 0.02f, 0.03f,
                      true, 0.8f, 0.2f, 0.2f, H, W, R + 1);
    }
  }
  // salt & pepper (data.py:94-95: p = 0.7, kornia defaults amount (0.01, 0.06), salt_vs_pepper (0.4, 0.6))
  P[34] = 0.f; P[35] = 0.f; P[36] = 0.f;
  if (cfg.salt_and_pepper && U(img, 32) < 0.7f) {
    P[34] = 1.f;
    P[35] = lerp_rn(U(img, 33), 0.01f, 0.05f);
    P[36] = lerp_rn(U(img, 34), 0.4f, 0.2f);
  }
  P[37] = U(img, 35);
  P[38] = 0.f; P[39] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// plasma shadow mask: kornia.contrib.diamond_square (RandomPlasmaShadow, data.py:87-92) per image.
//   seed grid (sh x sw, all uniform) -> `depth` doublings; at level k (scale = roughness^k):
//     diamond centres (odd, odd)   = (1 - scale) * mean of the 4 diagonal parents + scale * u
//     square centres (one odd)     = (1 - scale) * (sum of the 4 axial neighbours / 4, x 4/3 on the border) + scale * u
//   final grid (2^ceil(log2(H-1)) + 1 per side) cropped to H x W; shadow where the field < shade_quantity.
// All levels but the last live in shared memory (ping-pong); the last one is evaluated per pixel on the fly.
// ------------------------------------------------------------------------------------------------------------
constexpr int kPlasmaMaxA = 129 * 129;   // largest second-to-last level (256 x 256 images)
constexpr int kPlasmaMaxB = 65 * 65;
constexpr int kPlasmaMaxD = 128 * 128;   // diamond centres of the last level
constexpr int kPlasmaSmemBytes = (kPlasmaMaxA + kPlasmaMaxB + kPlasmaMaxD) * 4;
constexpr int kPlasmaThreads = 1024;

struct PlasmaGeom {
  int depth, sh, sw;   // number of doublings, seed grid
};
__host__ __device__ inline int ceil_log2_int(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
__host__ __device__ inline PlasmaGeom plasma_geom(int H, int W) {
  const int lh = ceil_log2_int(H - 1), lw = ceil_log2_int(W - 1);
  PlasmaGeom g;
  g.depth = (lh < lw ? lh : lw) - 1;
  g.sh = (1 << (lh - g.depth)) + 1;
  g.sw = (1 << (lw - g.depth)) + 1;
  return g;
}

__device__ __forceinline__ float ds_mix(float scale, float region, float u) {
  return __fadd_rn(__fmul_rn(__fadd_rn(1.f, -scale), region), __fmul_rn(scale, u));
}
__device__ __forceinline__ float ds_diamond(const float* __restrict__ prev, int pw, int y, int x, float scale,
                                            uint32_t seed32, uint32_t level) {
  const float* p = prev + (y >> 1) * pw + (x >> 1);
  const float m = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[pw]), p[pw + 1]));
  return ds_mix(scale, m, pixel_uniform(seed32, level, y, x));
}

__global__ void __launch_bounds__(kPlasmaThreads)
plasma_mask_kernel(const float* __restrict__ params, uint32_t* __restrict__ mask, int H, int W) {
  pdl_prologue();
  extern __shared__ float s_plasma[];
  float* bufA = s_plasma;
  float* bufB = s_plasma + kPlasmaMaxA;
  const int n = blockIdx.x;
  const float* P = params + static_cast<size_t>(n) * kAugParams;
  const float roughness = P[17], intensity = P[18], quantity = P[19];
  uint32_t* out = mask + static_cast<size_t>(n) * H * (W >> 5);
  const int words = H * (W >> 5);
  if (intensity == 0.f) {
    for (int i = threadIdx.x; i < words; i += kPlasmaThreads) out[i] = 0u;
    return;
  }
  const uint32_t seed32 = static_cast<uint32_t>(rintf(P[20] * 16777216.f));
  const PlasmaGeom g = plasma_geom(H, W);
  // level `depth - 1` must land in bufA: alternate backwards
  float* cur = ((g.depth - 1) & 1) ? bufB : bufA;
  float* other = (cur == bufA) ? bufB : bufA;
  int h = g.sh, w = g.sw;
  for (int i = threadIdx.x; i < h * w; i += kPlasmaThreads) cur[i] = pixel_uniform(seed32, 0u, i / w, i % w);
  __syncthreads();
  float scale = 1.f;
  for (int level = 1; level < g.depth; ++level) {
    scale = __fmul_rn(scale, roughness);
    const int nh = 2 * h - 1, nw = 2 * w - 1;
    float* nxt = other;
    for (int i = threadIdx.x; i < nh * nw; i += kPlasmaThreads) {
      const int y = i / nw, x = i - y * nw;
      if (!((y | x) & 1)) nxt[i] = cur[(y >> 1) * w + (x >> 1)];
      else if (y & x & 1) nxt[i] = ds_diamond(cur, w, y, x, scale, seed32, level);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nh * nw; i += kPlasmaThreads) {
      const int y = i / nw, x = i - y * nw;
      if (((y ^ x) & 1) == 0) continue;
      const float up = y > 0 ? nxt[i - nw] : 0.f, down = y < nh - 1 ? nxt[i + nw] : 0.f;
      const float left = x > 0 ? nxt[i - 1] : 0.f, right = x < nw - 1 ? nxt[i + 1] : 0.f;
      float r = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(up, left), right), down));
      if (y == 0 || x == 0 || y == nh - 1 || x == nw - 1) r = __fmul_rn(r, 1.33333337306976318359375f);
      nxt[i] = ds_mix(scale, r, pixel_uniform(seed32, level, y, x));
    }
    __syncthreads();
    other = cur; cur = nxt; h = nh; w = nw;
  }
  // last level from `cur` (h x w): final grid (2h-1) x (2w-1), cropped to H x W. Its diamond centres (odd, odd) go to
  // shared memory once; the square centres are evaluated per pixel on the fly, nothing else is stored.
  scale = __fmul_rn(scale, roughness);
  const uint32_t level = g.depth;
  const int fh = 2 * h - 1, fw = 2 * w - 1;
  float* dia = s_plasma + kPlasmaMaxA + kPlasmaMaxB;   // [(h - 1) x (w - 1)]: centre (2i+1, 2j+1) at [i][j]
  const int dw = w - 1;
  for (int i = threadIdx.x; i < (h - 1) * dw; i += kPlasmaThreads) {
    const int dy = i / dw, dx = i - dy * dw;
    dia[i] = ds_diamond(cur, w, 2 * dy + 1, 2 * dx + 1, scale, seed32, level);
  }
  __syncthreads();
  auto value_at = [&](int y, int x) -> float {   // even-even or odd-odd positions only
    if (!((y | x) & 1)) return cur[(y >> 1) * w + (x >> 1)];
    return dia[(y >> 1) * dw + (x >> 1)];
  };
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wshift = 31 - __clz(W >> 5);   // W / 32 is a power of two for W in {32, 64, 128, 256}; else divide
  const int wpr = W >> 5;                  // words per row
  const bool pow2 = (wpr & (wpr - 1)) == 0;
  for (int wi = warp; wi < words; wi += kPlasmaThreads / 32) {
    const int y = pow2 ? (wi >> wshift) : wi / wpr;
    const int x = (wi - y * wpr) * 32 + lane;
    float f;
    if (((y ^ x) & 1) == 0) {
      f = value_at(y, x);
    } else {
      const float up = y > 0 ? value_at(y - 1, x) : 0.f, down = y < fh - 1 ? value_at(y + 1, x) : 0.f;
      const float left = x > 0 ? value_at(y, x - 1) : 0.f, right = x < fw - 1 ? value_at(y, x + 1) : 0.f;
      float r = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(up, left), right), down));
      if (y == 0 || x == 0 || y == fh - 1 || x == fw - 1) r = __fmul_rn(r, 1.33333337306976318359375f);
      f = ds_mix(scale, r, pixel_uniform(seed32, level, y, x));
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, f < quantity);
    if (lane == 0) out[wi] = bits;
  }
}

// ------------------------------------------------------------------------------------------------------------
// colour operations (kornia.enhance semantics, see oracle/augment.py). Fast-math divisions: results agree with the
// float32 oracle to a few ulp (asserted at 1e-5 absolute on [0, 1] images); the output is bf16 anyway.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp01(float v) { return __saturatef(v); }

__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h6, float& s, float& v) {
  // h6 = hue in sixths of a turn, in [0, 6)
  const float mx = fmaxf(fmaxf(r, g), b), mn = fminf(fminf(r, g), b);
  v = mx;
  const float delta = mx - mn;
  s = __fdividef(delta, mx + 1e-8f);
  const float inv = __fdividef(1.f, delta == 0.f ? 1.f : delta);
  float hh = (r == mx) ? (g - b) * inv : ((g == mx) ? fmaf(b - r, inv, 2.f) : fmaf(r - g, inv, 4.f));
  h6 = hh < 0.f ? hh + 6.f : hh;
}
__device__ __forceinline__ void hsv_to_rgb(float h6, float s, float v, float& r, float& g, float& b) {
  // branch-free form of the sector table (no divergence between pixels of different hue):
  //   channel = v - v s clamp(min(k, 4 - k), 0, 1),  k = (n + h6) mod 6,  n = 5, 3, 1 for R, G, B
  // e.g. sector 0 (h6 = f): R -> v, G -> v (1 - (1 - f) s) = t, B -> v (1 - s) = p, exactly kornia's (v, t, p)
  const float vs = v * s;
  float kr = h6 + 5.f, kg = h6 + 3.f, kb = h6 + 1.f;
  kr = kr >= 6.f ? kr - 6.f : kr;
  kg = kg >= 6.f ? kg - 6.f : kg;
  kb = kb >= 6.f ? kb - 6.f : kb;
  r = fmaf(-vs, __saturatef(fminf(kr, 4.f - kr)), v);
  g = fmaf(-vs, __saturatef(fminf(kg, 4.f - kg)), v);
  b = fmaf(-vs, __saturatef(fminf(kb, 4.f - kb)), v);
}
// planckian gain, then the four jiggle ops in the sampled order. Saturation and hue act on different HSV components, so
// when they are adjacent in the order they share one RGB <-> HSV round trip.
__device__ __forceinline__ void color_ops(float& r, float& g, float& b, const float* __restrict__ P, uint32_t code,
                                          float hue6) {
  r = fminf(r * P[0], 1.f);
  b = fminf(b * P[1], 1.f);
  if (code == 0xFFFFFFFFu) return;
#pragma unroll
  for (int slot = 0; slot < 4; ++slot) {
    const int op = (code >> (6 - 2 * slot)) & 3;
    if (op == 0) {
      r = clamp01(r + P[2]); g = clamp01(g + P[2]); b = clamp01(b + P[2]);
    } else if (op == 1) {
      r = clamp01(r * P[3]); g = clamp01(g * P[3]); b = clamp01(b * P[3]);
    } else {
      const int nxt = slot < 3 ? static_cast<int>((code >> (4 - 2 * slot)) & 3) : -1;
      const int prv = slot > 0 ? static_cast<int>((code >> (8 - 2 * slot)) & 3) : -1;
      if (prv >= 2) continue;   // already applied together with the previous (saturation / hue) slot
      const bool both = nxt >= 2;
      float h6, s, v;
      rgb_to_hsv(r, g, b, h6, s, v);
      if (op == 2 || both) s = clamp01(s * P[4]);
      if (op == 3 || both) {
        h6 += hue6;
        h6 = h6 < 0.f ? h6 + 6.f : (h6 >= 6.f ? h6 - 6.f : h6);
      }
      hsv_to_rgb(h6, s, v, r, g, b);
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
augment_kernel(const void* __restrict__ in, void* __restrict__ out, const float* __restrict__ params,
               const uint32_t* __restrict__ arc_mask, const uint32_t* __restrict__ plasma_mask, int H, int W, int apply) {
  pdl_prologue();
  __shared__ float sA[3][kIn][kIn + 1];    // colour-jittered input with halo; later reused for the blurred tile
  __shared__ float sB[3][kIn][kMid + 1];   // after the horizontal gaussian pass
  __shared__ float sP[kAugParams];
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
  // (apply == 0 is plain u8 -> bf16 staging: the parameter table may be NULL and is never read)
  if (threadIdx.x < kAugParams) sP[threadIdx.x] = apply ? params[static_cast<size_t>(n) * kAugParams + threadIdx.x] : 0.f;
  __syncthreads();
  const float sigma = apply ? sP[7] : 0.f;
  const bool motion = apply && !(sP[12] == 1.f);           // identity kernel = off
  // halo actually needed by this image: 2 for the gaussian, 1 for the motion kernel
  const int halo = (sigma > 0.f ? 2 : 0) + (motion ? 1 : 0);
  const int lo = kHalo - halo, span = kTile + 2 * halo;   // tile-local index range [lo, lo + span) per axis
  const uint32_t code = (apply && sP[6] >= 0.f) ? c_orders[static_cast<int>(sP[6])] : 0xFFFFFFFFu;
  const float hue6 = sP[5] * 0.95492965855137201461f;   // radians -> sixths of a turn
  const int e1 = apply && sP[24] != 0.f, e2 = apply && sP[29] != 0.f;
  const int e1x = static_cast<int>(sP[25]), e1y = static_cast<int>(sP[26]), e1w = static_cast<int>(sP[27]), e1h = static_cast<int>(sP[28]);
  const int e2x = static_cast<int>(sP[30]), e2y = static_cast<int>(sP[31]), e2w = static_cast<int>(sP[32]), e2h = static_cast<int>(sP[33]);
  const size_t img_words = static_cast<size_t>(H) * (W >> 5);

  // ---- stage 1: load (+ reflect), arcs, /255, erasing, colour ops. Tile-local pixel (ty, tx), both in [lo, lo + span).
  const bool border_tile = (y0 == 0) | (x0 == 0) | (y0 + kTile == H) | (x0 + kTile == W);
  const bool erasing = (e1 | e2) != 0;
  auto load_pixel = [&](int ty, int tx) {
    int gy = y0 - kHalo + ty, gx = x0 - kHalo + tx;
    if (border_tile) {   // reflect (only tiles on the image border can reach outside)
      gy = gy < 0 ? -gy : (gy >= H ? 2 * (H - 1) - gy : gy);
      gx = gx < 0 ? -gx : (gx >= W ? 2 * (W - 1) - gx : gx);
    }
    float r, g, b;
    if (IN_U8) {
      const uint8_t* p = static_cast<const uint8_t*>(in) + (static_cast<size_t>(n) * H * W + static_cast<size_t>(gy) * W + gx) * 3;
      r = static_cast<float>(p[0]) * (1.0f / 255.0f);
      g = static_cast<float>(p[1]) * (1.0f / 255.0f);
      b = static_cast<float>(p[2]) * (1.0f / 255.0f);
      if (arc_mask != nullptr) {
        const uint32_t wbits = __ldg(arc_mask + n * img_words + static_cast<size_t>(gy) * (W >> 5) + (gx >> 5));
        if ((wbits >> (gx & 31)) & 1u) { r = 0.f; g = 0.f; b = 0.f; }
      }
    } else {
      const float* p = static_cast<const float*>(in) + static_cast<size_t>(n) * 3 * H * W + static_cast<size_t>(gy) * W + gx;
      r = p[0]; g = p[static_cast<size_t>(H) * W]; b = p[2 * static_cast<size_t>(H) * W];
    }
    if (apply) {
      if (erasing) {
        if (e1 && gx >= e1x && gx < e1x + e1w && gy >= e1y && gy < e1y + e1h) { r = 0.f; g = 0.f; b = 0.f; }
        if (e2 && gx >= e2x && gx < e2x + e2w && gy >= e2y && gy < e2y + e2h) { r = 1.f; g = 1.f; b = 1.f; }
      }
      color_ops(r, g, b, sP, code, hue6);
    }
    sA[0][ty][tx] = r; sA[1][ty][tx] = g; sA[2][ty][tx] = b;
  };
  {
    // no integer division anywhere: warp w takes the rows w, w + 8, ... with one lane per column for the first 32
    // columns; the (span - 32) <= 6 halo columns left over are a second, flattened pass (one pixel per thread)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < span; r += 8) load_pixel(lo + r, lo + lane);
    const int extra = span - 32;                      // 0, 2, 4 or 6 columns
    if (extra > 0) {
      // thread t -> row t / extra, column 32 + t % extra, with extra in {2, 4, 6}: multiply-shift instead of a division
      const int t = threadIdx.x;
      const int row = extra == 2 ? (t >> 1) : (extra == 4 ? (t >> 2) : (t * 10923) >> 16);   // t / 6 for t < 256
      const int col = t - row * extra;
      if (row < span) load_pixel(lo + row, lo + 32 + col);
    }
  }
  __syncthreads();

  // ---- stage 2: separable 5-tap gaussian (reflect already materialised in the halo). One thread = one run of 16 / 17
  // consecutive outputs of one (channel, line): a sliding window in registers, 1.25 shared-memory loads per output.
  int off = 2;   // index offset of the (blurred or not) image window inside sA: window (oy, ox) <-> pixel (y0-1+oy, ..)
  if (sigma > 0.f) {
    float k[5];
    float ks = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float d = static_cast<float>(i - 2);
      k[i] = __expf(-d * d / (2.f * sigma * sigma));
      ks += k[i];
    }
    const float inv = 1.f / ks;
#pragma unroll
    for (int i = 0; i < 5; ++i) k[i] *= inv;
    // window = tile + 1 pixel of motion halo when the motion kernel is on (mlo = 0), else just the tile (mlo = 1)
    const int mlo = motion ? 0 : 1, mspan = kMid - 2 * mlo;   // 34 or 32 outputs per line, in two runs of 17 / 16
    const int run = mspan >> 1;
    {
      // horizontal: item = (channel, half, line); lanes walk the lines (row stride 39 floats: conflict free)
      const int items = 3 * 2 * span;
      for (int it = threadIdx.x; it < items; it += 256) {
        const int c = it >= 4 * span ? 2 : (it >= 2 * span ? 1 : 0);
        const int rem = it - c * 2 * span;
        const int half = rem >= span ? 1 : 0;
        const int ty = lo + rem - half * span;
        const int ox0 = mlo + half * run;
        const float* src = &sA[c][ty][ox0];
        float* dst = &sB[c][ty][ox0];
        float w0 = src[0], w1 = src[1], w2 = src[2], w3 = src[3];
#pragma unroll
        for (int j = 0; j < 17; ++j) {
          if (j < run) {
            const float w4 = src[j + 4];
            dst[j] = fmaf(k[4], w4, fmaf(k[3], w3, fmaf(k[2], w2, fmaf(k[1], w1, k[0] * w0))));
            w0 = w1; w1 = w2; w2 = w3; w3 = w4;
          }
        }
      }
    }
    __syncthreads();
    {
      // vertical: item = (channel, half, column); lanes walk the columns
      const int items = 3 * 2 * mspan;
      for (int it = threadIdx.x; it < items; it += 256) {
        const int c = it >= 4 * mspan ? 2 : (it >= 2 * mspan ? 1 : 0);
        const int rem = it - c * 2 * mspan;
        const int half = rem >= mspan ? 1 : 0;
        const int ox = mlo + rem - half * mspan;
        const int oy0 = mlo + half * run;
        float w0 = sB[c][oy0][ox], w1 = sB[c][oy0 + 1][ox], w2 = sB[c][oy0 + 2][ox], w3 = sB[c][oy0 + 3][ox];
#pragma unroll
        for (int j = 0; j < 17; ++j) {
          if (j < run) {
            const float w4 = sB[c][oy0 + j + 4][ox];
            sA[c][oy0 + j][ox] = fmaf(k[4], w4, fmaf(k[3], w3, fmaf(k[2], w2, fmaf(k[1], w1, k[0] * w0))));
            w0 = w1; w1 = w2; w2 = w3; w3 = w4;   // (oy, ox) <-> pixel (y0 - 1 + oy, x0 - 1 + ox)
          }
        }
      }
    }
    off = 0;
    __syncthreads();
  }
  // motion blur uses a zero ('constant') border: blank the window positions that fall outside the image
  if (motion && (y0 == 0 || x0 == 0 || y0 + kTile == H || x0 + kTile == W)) {
    for (int i = threadIdx.x; i < kMid * kMid; i += 256) {
      const int oy = i / kMid, ox = i - oy * kMid;
      const int gy = y0 - 1 + oy, gx = x0 - 1 + ox;
      if (gy < 0 || gy >= H || gx < 0 || gx >= W) {
        sA[0][oy + off][ox + off] = 0.f; sA[1][oy + off][ox + off] = 0.f; sA[2][oy + off][ox + off] = 0.f;
      }
    }
    __syncthreads();
  }

  // ---- stage 3: 3x3 motion kernel, plasma shadow, clamp, salt & pepper, store. Thread = one 2x2 quad (one s2d pixel).
  const int sy = threadIdx.x >> 4, sx = threadIdx.x & 15;
  const float intensity = apply ? sP[18] : 0.f;
  uint32_t shadow[2] = {0u, 0u};
  if (intensity != 0.f) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
      shadow[a] = __ldg(plasma_mask + n * img_words + static_cast<size_t>(y0 + 2 * sy + a) * (W >> 5) + (x0 >> 5)) >> (2 * sx);
  }
  const bool snp = apply && sP[34] != 0.f;
  const uint32_t snp_seed = static_cast<uint32_t>(rintf(sP[37] * 16777216.f));
  float v[2][2][3];
  float km[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) km[i] = sP[8 + i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float w[4][4];
    if (motion) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) w[i][j] = sA[c][2 * sy + i + off][2 * sx + j + off];
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float acc;
        if (motion) {
          acc = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = fmaf(km[i * 3 + j], w[a + i][b + j], acc);
        } else {
          acc = sA[c][2 * sy + a + 1 + off][2 * sx + b + 1 + off];
        }
        const float shade = ((shadow[a] >> b) & 1u) ? intensity : 0.f;
        v[a][b][c] = clamp01(acc + shade);
      }
  }
  if (snp) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const uint32_t gy = y0 + 2 * sy + a, gx = x0 + 2 * sx + b;
        if (pixel_uniform(snp_seed, 100u, gy, gx) < sP[35]) {
          const float val = pixel_uniform(snp_seed, 101u, gy, gx) < sP[36] ? 1.f : 0.f;
          v[a][b][0] = val; v[a][b][1] = val; v[a][b][2] = val;
        }
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
// spaghetti arcs (reference: argus/utils.py:252-275 draw_spaghetti, applied to the decoded image at argus/data.py:212-215
// before the kornia chain; PIL's ImageDraw.arc on the loader's CPU workers there). The rasteriser below restates Pillow's
// algorithm (integer quadrant walk of the outer / inner ellipse + clipping by the ellipse normals at the start / end
// angles) and is pinned pixel for pixel against the real Pillow through oracle/pil_arc.py.
// ------------------------------------------------------------------------------------------------------------
constexpr uint64_t kArcFieldBase = 1000;

__global__ void spaghetti_params_kernel(float* __restrict__ arcs, int n_images, int n_arcs, int H, int W, uint64_t seed,
                                        uint64_t step) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_images * n_arcs) return;
  const uint64_t img = t / n_arcs, a = t % n_arcs;
  auto U = [&](uint64_t k) { return hash_uniform(seed, step, img, kArcFieldBase + a * 8 + k); };
  auto randint = [](float u, int lo, int hi) {   // np.random.randint(lo, hi)
    const int v = lo + static_cast<int>(floorf(__fmul_rn(u, static_cast<float>(hi - lo))));
    return v < hi - 1 ? v : hi - 1;
  };
  const int x0 = randint(U(0), 0, W), y0 = randint(U(1), 0, H);
  const int x1 = randint(U(2), x0, W), y1 = randint(U(3), y0, H);
  const int a0 = randint(U(4), 0, 360), a1 = randint(U(5), 0, 360);
  float* A = arcs + static_cast<size_t>(t) * kArcFields;
  A[0] = static_cast<float>(x0); A[1] = static_cast<float>(y0); A[2] = static_cast<float>(x1); A[3] = static_cast<float>(y1);
  A[4] = static_cast<float>(a0); A[5] = static_cast<float>(a1);
  A[6] = floorf(__fadd_rn(1.0f, __fmul_rn(U(6), 4.0f)));   // int(np.random.uniform(1, 5))
  A[7] = 0.f;
}

// Pillow's quarter_state: one quadrant of the ellipse with semi-axes (a, b) in doubled coordinates, walked from
// (a, b % 2) to (a % 2, b) by steps (0,+2), (-2,+2), (-2,0), each time to the candidate closest to the ellipse.
struct QuarterWalk {
  int a, b, cx, cy, ex, ey;
  long long a2, b2, a2b2;
  bool finished;
  __device__ void init(int a_, int b_) {
    finished = a_ < 0 || b_ < 0;
    a = a_; b = b_;
    cx = a_; cy = b_ & 1; ex = a_ & 1; ey = b_;
    a2 = static_cast<long long>(a_) * a_; b2 = static_cast<long long>(b_) * b_; a2b2 = a2 * b2;
  }
  __device__ long long delta(long long x, long long y) const { return llabs(a2 * y * y + b2 * x * x - a2b2); }
  __device__ bool next(int& rx, int& ry) {
    if (finished) return false;
    rx = cx; ry = cy;
    if (cx == ex && cy == ey) {
      finished = true;
    } else {
      int nx = cx, ny = cy + 2;
      long long nd = delta(nx, ny);
      if (nx > 1) {
        long long d = delta(cx - 2, cy + 2);
        if (nd > d) { nx = cx - 2; ny = cy + 2; nd = d; }
        d = delta(cx - 2, cy);
        if (nd > d) { nx = cx - 2; ny = cy; }
      }
      cx = nx; cy = ny;
    }
    return true;
  }
};

struct ArcInterval { int lo, hi; bool ok; };
__device__ __forceinline__ ArcInterval iv_isect(ArcInterval p, ArcInterval q) {
  ArcInterval r;
  r.lo = max(p.lo, q.lo); r.hi = min(p.hi, q.hi);
  r.ok = p.ok && q.ok && r.lo <= r.hi;
  return r;
}
__device__ __forceinline__ int pil_round_up(double f) {
  return f >= 0.0 ? static_cast<int>(floor(f + 0.5)) : -static_cast<int>(floor(fabs(f) + 0.5));
}
__device__ __forceinline__ int pil_round_down(double f) {
  return f >= 0.0 ? static_cast<int>(ceil(f - 0.5)) : -static_cast<int>(ceil(fabs(f) - 0.5));
}
constexpr int kArcInf = 1 << 30;
// integer interval of the scan coordinate x with A x + B y + C >= 0
__device__ __forceinline__ ArcInterval arc_halfplane(double A, double B, double C, int y) {
  ArcInterval r;
  r.lo = -kArcInf; r.hi = kArcInf; r.ok = true;
  if (A > 1e-9) r.lo = pil_round_up(-(B * y + C) / A);
  else if (A < -1e-9) r.hi = pil_round_down(-(B * y + C) / A);
  else r.ok = (B * y + C >= -1e-9);
  return r;
}
// what the arc [al, ar] leaves of the half ellipse k (0: angles 0..180, 1: 180..360):
// 0 none, 1 all, 2 from the start cap on, 3 up to the end cap, 4 between the caps, 5 everything but the gap
__device__ int arc_half_rule(double al, double ar, int k) {
  const double e = ar < 360.0 ? ar : ar - 360.0;
  const double q0 = 180.0 * k, q1 = q0 + 180.0;
  const bool has_start = q0 <= al && al < q1, has_end = q0 < e && e <= q1;
  if (has_start && has_end) return (ar - al) < 180.0 ? 4 : 5;
  if (has_start) return 2;
  if (has_end) return 3;
  const double mid = q0 + 90.0;
  return ((al <= mid && mid <= ar) || (al <= mid + 360.0 && mid + 360.0 <= ar)) ? 1 : 0;
}
__device__ void arc_normalize(double& al, double& ar) {
  if (ar - al >= 360.0) { al = 0.0; ar = 360.0; return; }
  al = fmod(al, 360.0);
  if (al < 0.0) al += 360.0;
  double d = fmod(ar - al, 360.0);
  if (d < 0.0) d += 360.0;
  ar = al + d;
}

constexpr int kArcMaxRows = 260;   // rows of the quadrant walk: Y = b % 2, b % 2 + 2, ..., b  (b <= 511)

// one warp per arc: lanes 0 / 1 walk the outer / inner ellipse, then all lanes clip and paint the rows
__global__ void __launch_bounds__(32)
arc_paint_kernel(const float* __restrict__ arcs, uint32_t* __restrict__ mask, int n_arcs, int H, int W) {
  pdl_prologue();
  __shared__ int16_t s_r[kArcMaxRows], s_l[kArcMaxRows];
  const int n = blockIdx.x / n_arcs;
  const float* A = arcs + static_cast<size_t>(blockIdx.x) * kArcFields;
  const int x0 = static_cast<int>(A[0]), y0 = static_cast<int>(A[1]), x1 = static_cast<int>(A[2]), y1 = static_cast<int>(A[3]);
  const int wd = static_cast<int>(A[6]);
  const int a = x1 - x0, b = y1 - y0;
  double al = A[4], ar = A[5];
  arc_normalize(al, ar);
  if (ar == al || a < 0 || b < 0 || wd < 1 || (b >> 1) + 1 > kArcMaxRows) return;
  const bool full = (ar == al + 360.0);
  const int lane = threadIdx.x;
  const int par = b & 1, leftmost = a & 1;
  const int nrows = (b >> 1) + 1;   // index j <-> Y = par + 2 j
  for (int j = lane; j < nrows; j += 32) { s_r[j] = -1; s_l[j] = static_cast<int16_t>(leftmost); }
  __syncwarp();
  if (lane == 0) {
    // r(Y) = x of the FIRST point of the outer walk on row Y
    QuarterWalk q;
    q.init(a, b);
    int cx, cy;
    while (q.next(cx, cy)) {
      const int j = (cy - par) >> 1;
      if (s_r[j] < 0) s_r[j] = static_cast<int16_t>(cx);
    }
  } else if (lane == 1) {
    // l(Y) = x of the LAST point of the inner walk on row Y (rows the inner ellipse does not reach keep `leftmost`)
    QuarterWalk q;
    q.init(a - 2 * (wd - 1), b - 2 * (wd - 1));
    int cx, cy;
    while (q.next(cx, cy)) s_l[(cy - par) >> 1] = static_cast<int16_t>(cx);
  }
  __syncwarp();
  // clip tree (wide frame; a tall ellipse is handled transposed and the tree transposed back)
  double lcA = 0, lcB = 0, lcC = 0, rcA = 0, rcB = 0, rcC = 0;
  int rule[2] = {1, 1};
  const bool transposed = a < b;
  if (!full) {
    double A_ = a, B_ = b, al2 = al, ar2 = ar;
    if (transposed) {
      A_ = b; B_ = a;
      al2 = 90.0 - ar; ar2 = 90.0 - al;
      arc_normalize(al2, ar2);
    }
    const double kPi = 3.14159265358979323846;
    lcA = -A_ * sin(al2 * kPi / 180.0); lcB = B_ * cos(al2 * kPi / 180.0);
    lcC = (A_ * A_ - B_ * B_) * sin(al2 * kPi / 90.0) / 2.0;
    rcA = A_ * sin(ar2 * kPi / 180.0); rcB = -B_ * cos(ar2 * kPi / 180.0);
    rcC = (B_ * B_ - A_ * A_) * sin(ar2 * kPi / 90.0) / 2.0;
    rule[0] = arc_half_rule(al2, ar2, 0);
    rule[1] = arc_half_rule(al2, ar2, 1);
    if (transposed) {
      double t = lcA; lcA = lcB; lcB = t;
      t = rcA; rcA = rcB; rcB = t;
    }
  }
  uint32_t* img = mask + static_cast<size_t>(n) * H * (W >> 5);
  auto paint = [&](int py, ArcInterval o) {
    if (!o.ok) return;
    int p0 = (o.lo + a) >> 1, p1 = (o.hi + a) >> 1;   // o.lo + a >= 0 whenever the interval survives the segment clip
    p0 = max(p0, 0); p1 = min(p1, a);
    const int y = y0 + py;
    if (p0 > p1 || y < 0 || y >= H) return;
    int xa = max(x0 + p0, 0), xb = min(x0 + p1, W - 1);
    for (int w0 = xa >> 5; w0 <= (xb >> 5); ++w0) {
      const int lo = max(xa, w0 << 5) & 31, hi = min(xb, (w0 << 5) + 31) & 31;
      const uint32_t bits = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
      atomicOr(img + static_cast<size_t>(y) * (W >> 5) + w0, bits);
    }
  };
  for (int py = lane; py <= b; py += 32) {
    const int sy = 2 * py - b;
    const int Y = sy < 0 ? -sy : sy;
    const int j = (Y - par) >> 1;
    const int l = s_l[j], r = s_r[j];
    ArcInterval segs[2];
    int nseg;
    if (!(l > 0 || l < r)) {
      segs[0].lo = -r; segs[0].hi = r; segs[0].ok = true; nseg = 1;
    } else {
      // solid row of an even-width ellipse (l == 0): the centre pixel belongs to the left segment only
      segs[0].lo = -r; segs[0].hi = -l; segs[0].ok = true;
      segs[1].lo = l > 0 ? l : 2; segs[1].hi = r; segs[1].ok = segs[1].lo <= r;
      nseg = 2;
    }
    for (int s = 0; s < nseg; ++s) {
      if (!segs[s].ok) continue;
      if (full) { paint(py, segs[s]); continue; }
      const ArcInterval hl = arc_halfplane(lcA, lcB, lcC, sy), hr = arc_halfplane(rcA, rcB, rcC, sy);
      for (int k = 0; k < 2; ++k) {
        // half-plane node of half k: Y >= 0 / Y <= 0 in the wide frame, X >= 0 / X <= 0 once transposed back
        const double sgn = k == 0 ? 1.0 : -1.0;
        const ArcInterval hk = transposed ? arc_halfplane(sgn, 0.0, 0.0, sy) : arc_halfplane(0.0, sgn, 0.0, sy);
        const ArcInterval base = iv_isect(segs[s], hk);
        if (!base.ok || rule[k] == 0) continue;
        switch (rule[k]) {
          case 1: paint(py, base); break;
          case 2: paint(py, iv_isect(base, hl)); break;
          case 3: paint(py, iv_isect(base, hr)); break;
          case 4: paint(py, iv_isect(iv_isect(base, hl), hr)); break;
          default: paint(py, iv_isect(base, hl)); paint(py, iv_isect(base, hr)); break;
        }
      }
    }
  }
}

// out = in with the masked pixels painted black (uint8 HWC)
__global__ void __launch_bounds__(256)
spaghetti_apply_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const uint32_t* __restrict__ mask,
                       int64_t n_pixels) {
  pdl_prologue();
  const int64_t pix = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (pix >= n_pixels) return;
  const bool hit = (mask[pix >> 5] >> (pix & 31)) & 1u;
  const size_t o = static_cast<size_t>(pix) * 3;
  out[o] = hit ? 0 : in[o];
  out[o + 1] = hit ? 0 : in[o + 1];
  out[o + 2] = hit ? 0 : in[o + 2];
}

void spaghetti_sample_params(float* arcs, int n_images, int n_arcs, int H, int W, uint64_t seed, uint64_t step,
                             cudaStream_t s) {
  ARGUS_CHECK(n_arcs >= 0, "negative arc count");
  if (n_images * n_arcs <= 0) return;
  ProfileScope prof("augment_params", s, 0, 4.0 * kArcFields * n_images * n_arcs);
  launch_kernel(spaghetti_params_kernel, (n_images * n_arcs + 127) / 128, 128, 0, s, arcs, n_images, n_arcs, H, W, seed, step);
  ARGUS_CUDA(cudaGetLastError());
}
void spaghetti_mask(const float* arcs, uint32_t* mask, int n_images, int n_arcs, int H, int W, cudaStream_t s) {
  ARGUS_CHECK(W % 32 == 0, "the arc mask needs W to be a multiple of 32");
  ARGUS_CHECK(H <= 512 && W <= 512, "arc rasteriser: images up to 512 x 512");
  if (n_images <= 0) return;
  ProfileScope prof("spaghetti", s, 0, static_cast<double>(n_images) * H * W / 4.0);
  ARGUS_CUDA(cudaMemsetAsync(mask, 0, static_cast<size_t>(n_images) * H * (W / 32) * sizeof(uint32_t), s));
  pdl_break(s, kPdlAfterMemop);
  if (n_arcs > 0) {
    launch_kernel(arc_paint_kernel, n_images * n_arcs, 32, 0, s, arcs, mask, n_arcs, H, W);
    ARGUS_CUDA(cudaGetLastError());
  }
}
void spaghetti_draw(const uint8_t* in, uint8_t* out, const float* arcs, uint32_t* mask_ws, int n_images, int n_arcs, int H,
                    int W, cudaStream_t s) {
  if (n_images <= 0) return;
  spaghetti_mask(arcs, mask_ws, n_images, n_arcs, H, W, s);
  ProfileScope prof("spaghetti", s, 0, 6.0 * n_images * H * W);
  const int64_t n_pixels = static_cast<int64_t>(n_images) * H * W;
  launch_kernel(spaghetti_apply_kernel, static_cast<unsigned>((n_pixels + 255) / 256), 256, 0, s, in, out, mask_ws, n_pixels);
  ARGUS_CUDA(cudaGetLastError());
}

void augment_sample_params(float* params, int n_images, int n_cams, int H, int W, uint64_t seed, uint64_t step,
                           const AugConfig& cfg, cudaStream_t s) {
  ProfileScope prof("augment_params", s, 0, 4.0 * kAugParams * n_images);
  if (n_images <= 0) return;
  launch_kernel(aug_params_kernel, (n_images + 127) / 128, 128, 0, s, params, n_images, n_cams, H, W, seed, step, cfg);
  ARGUS_CUDA(cudaGetLastError());
}

void augment_images(const void* in, bool in_u8, void* out, bool out_s2d, const float* params, const uint32_t* arc_mask,
                    uint32_t* plasma_mask, int n_images, int H, int W, bool apply, cudaStream_t s) {
  ARGUS_CHECK(H % kTile == 0 && W % kTile == 0, "augmentation needs H and W to be multiples of 32");
  ARGUS_CHECK(!(arc_mask != nullptr && !in_u8), "the arc mask applies to uint8 input");
  if (n_images <= 0) return;
  if (apply) {
    ARGUS_CHECK(plasma_mask != nullptr, "augmentation needs the plasma mask workspace (n * H * W / 8 bytes)");
    ARGUS_CHECK(H <= 256 && W <= 256, "plasma shadow: images up to 256 x 256 (the fractal is built in shared memory)");
    static bool attr_set = false;
    if (!attr_set) {
      ARGUS_CUDA(cudaFuncSetAttribute(plasma_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlasmaSmemBytes));
      attr_set = true;
    }
    ProfileScope prof("plasma_mask", s, 0, static_cast<double>(n_images) * H * W / 8.0);
    launch_kernel(plasma_mask_kernel, n_images, kPlasmaThreads, kPlasmaSmemBytes, s, params, plasma_mask, H, W);
    ARGUS_CUDA(cudaGetLastError());
  }
  // algorithmic bytes: u8 RGB in + bf16 RGB out = 9 B per pixel
  const double bytes = static_cast<double>(n_images) * H * W * (in_u8 ? 3.0 : 12.0) +
                       static_cast<double>(n_images) * H * W * (out_s2d ? 6.0 : 12.0);
  ProfileScope prof("augment", s, 0, bytes);
  dim3 grid(W / kTile, H / kTile, n_images);
  const int ap = apply ? 1 : 0;
  if (in_u8 && out_s2d) launch_kernel(augment_kernel<true, true>, grid, 256, 0, s, in, out, params, arc_mask, plasma_mask, H, W, ap);
  else if (in_u8 && !out_s2d) launch_kernel(augment_kernel<true, false>, grid, 256, 0, s, in, out, params, arc_mask, plasma_mask, H, W, ap);
  else if (!in_u8 && out_s2d) launch_kernel(augment_kernel<false, true>, grid, 256, 0, s, in, out, params, arc_mask, plasma_mask, H, W, ap);
  else launch_kernel(augment_kernel<false, false>, grid, 256, 0, s, in, out, params, arc_mask, plasma_mask, H, W, ap);
  ARGUS_CUDA(cudaGetLastError());
}

}  // namespace argus
