/*
 * argus_b200 — C ABI of libargus_b200.so: the B200 (sm_100a) implementation of the training / inference hot path
 * of pculbertson/argus.
 *
 * The reference is pure Python and has no FFI of its own; each entry point below names the reference call site it
 * replaces (paths relative to the reference tree). INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C types only: device pointers are `void*` / `float*`, sizes are `int` / `int64_t`, streams are passed
 *     as `void*` holding a `cudaStream_t` (NULL = legacy default stream);
 *   - every function returns 0 on success and non-zero on failure; `argus_last_error_string()` returns the
 *     thread-local message of the last failure; C++ exceptions never cross the boundary;
 *   - all work is enqueued asynchronously on the caller's stream; the caller owns every buffer it passes in;
 *   - there is NO CPU fallback: a missing GPU or a non-sm_100 GPU is an error;
 *   - activations are NHWC bf16, convolution weights are [Cout][kh][kw][Cin] bf16 ("packed" layout) unless noted;
 *     spatial sizes must be powers of two and channel counts multiples of 64.
 */
#ifndef ARGUS_B200_H_
#define ARGUS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------- */
const char* argus_last_error_string(void);
int argus_version(void);
/* Fails unless the current CUDA device is compute capability 10.x. */
int argus_require_device(void);

/* ---- convolution primitives (torch.nn.Conv2d / nn.Linear inside torchvision resnet50, called from
 *      argus/models.py:84; cuDNN / cuBLAS in the reference) ------------------------------------------------- */
/* kind 0: k in {1,3}, stride in {1,2}, pad = k/2.  kind 1: the 7x7/2 stem on the space-to-depth input produced by
 * argus_pack_input (x is [N][H/2][W/2+4][16] bf16, w is the stem weight repacked to [64][256] by
 * argus_pack_stem_weight); N, H, W always describe the conv input image. */
int argus_conv2d_forward(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int k,
                         int stride, int kind, const float* scale, const float* shift, const void* residual, int relu,
                         float* stat_sum, float* stat_sqsum, void* stream);
/* dx = conv_transpose(dy, w) (+ residual). For stride 2 the caller zero-fills dx first when k == 1. */
int argus_conv2d_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int Cin, int Cout, int k,
                       int stride, const void* residual, void* stream);
/* dw[Cout][k*k*Cin] (fp32) += dy^T * im2col(x); the caller zero-fills dw. */
int argus_conv2d_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, int k,
                       int stride, int kind, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARGUS_B200_H_ */
