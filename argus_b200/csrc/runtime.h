// Host-side runtime helpers shared by every translation unit of libargus_b200.so:
// error reporting across the C ABI, TMA tensor-map encoding, device buffers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace argus {

// Every C-ABI entry point wraps its body in ARGUS_API_BEGIN/END: C++ exceptions never cross the boundary,
// they become a non-zero return code plus a thread-local message (argus_last_error_string()).
void set_last_error(const std::string& msg);
const char* get_last_error();

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define ARGUS_CHECK(cond, msg)                                                                        \
  do {                                                                                                \
    if (!(cond)) throw ::argus::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + \
                                      std::string(msg));                                              \
  } while (0)

#define ARGUS_CUDA(expr)                                                                                \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw ::argus::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + #expr + ": " + \
                           cudaGetErrorString(_e));                                                     \
  } while (0)

#define ARGUS_API_BEGIN try { ::argus::pdl_break_all();
#define ARGUS_API_END                          \
  return 0;                                    \
  }                                            \
  catch (const std::exception& e) {            \
    ::argus::set_last_error(e.what());         \
    return 1;                                  \
  }                                            \
  catch (...) {                                \
    ::argus::set_last_error("unknown error");  \
    return 2;                                  \
  }

// bf16 tiled tensor map with SWIZZLE_128B. dims/strides innermost first; strides[i] is the byte stride of
// dimension i+1 (dimension 0 is contiguous). Out-of-bounds elements read as zero.
CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box);

int num_sms();
void require_sm100();

// ---- launch accounting and optional per-kernel-family timing (used by bench.py for the roofline figures) ----
// Every kernel launcher opens a ProfileScope: it always bumps the launch counter and, while profiling is enabled,
// brackets the launch with CUDA events on the launching stream and records algorithmic FLOPs / bytes.
void profile_enable(bool on);                    // clears previous records when switching on
bool profile_enabled();
int64_t launch_count();                          // kernels launched by this library since load
std::string profile_report_json();               // synchronises the recorded events and aggregates per family
bool profile_detailed();                         // ARGUS_PROFILE_DETAIL=1: conv families are split per layer shape
struct ProfileScope {
  ProfileScope(const char* family, cudaStream_t stream, double flops, double bytes);
  ProfileScope(const std::string& family, cudaStream_t stream, double flops, double bytes);
  ~ProfileScope();
  int slot = -1;
  cudaStream_t stream = nullptr;
};

// Kernel launch with a programmatic-stream-serialization edge to the previous kernel of the stream (PDL, see
// pdl_prologue() in ptx.cuh) when ARGUS_PDL=1; by default launches carry no attribute (plain stream order, see
// pdl_enabled() in conv_ops.cu for the measurement behind the default).
bool pdl_enabled();
// Every C-ABI entry point starts with pdl_break_all(): the first kernel of a call is launched without a programmatic
// edge (the caller may have enqueued anything in between). pdl_break() marks the event records / event waits / memsets /
// memcpys the library itself enqueues: the kernel that follows one of them is launched without the attribute, so
// programmatic edges only ever connect two back-to-back kernels of the library.
enum PdlBreakKind { kPdlAfterWait = 1, kPdlAfterRecord = 2, kPdlAfterMemop = 4 };
void pdl_break(cudaStream_t stream, int kind);
void pdl_break_all();
bool pdl_chain_ok(cudaStream_t stream);     // true when the next kernel on `stream` may carry the attribute
void pdl_mark_kernel(cudaStream_t stream);
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && pdl_chain_ok(stream)) ? 1 : 0;
  ARGUS_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  pdl_mark_kernel(stream);
}
#endif

inline int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace argus
