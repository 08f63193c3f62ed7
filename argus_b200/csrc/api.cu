// extern "C" entry points of libargus_b200.so (declared in include/argus_b200.h).
#include "../../include/argus_b200.h"

#include "conv_ops.h"
#include "runtime.h"

using namespace argus;

static ConvShape make_shape(int N, int H, int W, int Cin, int Cout, int k, int stride, int kind) {
  ConvShape s;
  s.N = N; s.H = H; s.W = W; s.Cin = Cin; s.Cout = Cout; s.k = k; s.stride = stride; s.kind = kind;
  return s;
}

extern "C" {

const char* argus_last_error_string(void) { return get_last_error(); }
int argus_version(void) { return 100; }

int argus_require_device(void) {
  ARGUS_API_BEGIN
  require_sm100();
  ARGUS_API_END
}

int argus_conv2d_forward(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int k,
                         int stride, int kind, const float* scale, const float* shift, const void* residual, int relu,
                         float* stat_sum, float* stat_sqsum, void* stream) {
  ARGUS_API_BEGIN
  require_sm100();
  ConvShape s = make_shape(N, H, W, Cin, Cout, k, stride, kind);
  ConvLaunch l = plan_conv_forward(s, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w),
                                   static_cast<__nv_bfloat16*>(y));
  Epilogue e;
  e.scale = scale; e.shift = shift; e.residual = static_cast<const __nv_bfloat16*>(residual); e.relu = relu;
  e.stat_sum = stat_sum; e.stat_sqsum = stat_sqsum;
  launch_conv(l, e, static_cast<cudaStream_t>(stream));
  ARGUS_API_END
}

int argus_conv2d_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int Cin, int Cout, int k,
                       int stride, const void* residual, void* stream) {
  ARGUS_API_BEGIN
  require_sm100();
  ConvShape s = make_shape(N, H, W, Cin, Cout, k, stride, 0);
  ARGUS_CHECK(residual == nullptr || stride == 1, "residual add is only supported for stride-1 dgrad");
  auto ls = plan_conv_dgrad(s, static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(w),
                            static_cast<__nv_bfloat16*>(dx));
  Epilogue e;
  e.residual = static_cast<const __nv_bfloat16*>(residual);
  for (auto& l : ls) launch_conv(l, e, static_cast<cudaStream_t>(stream));
  ARGUS_API_END
}

int argus_conv2d_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, int k,
                       int stride, int kind, void* stream) {
  ARGUS_API_BEGIN
  require_sm100();
  ConvShape s = make_shape(N, H, W, Cin, Cout, k, stride, kind);
  WgradLaunch l = plan_conv_wgrad(s, static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), dw);
  launch_wgrad(l, static_cast<cudaStream_t>(stream));
  ARGUS_API_END
}

}  // extern "C"
