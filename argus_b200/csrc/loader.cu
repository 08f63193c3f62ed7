// Double-buffered pinned-memory batch loader (host C++; reference: the DataLoader / DistributedSampler plumbing of
// /root/reference/argus/train.py:147-192 and the per-sample `.to(device)` at train.py:302-303).
//
// The reference decodes two PNGs per sample on forked CPU workers and ships fp32 tensors (1.57 MB per pair) with a
// synchronous H2D copy. Here samples are pre-decoded into a raw uint8 shard file (393 KB per pair); a worker thread
// gathers the next batch of the rank's index slice into pinned host memory while the GPU trains on the previous one,
// and the H2D copy runs on a side stream, ordered against the consumer's stream with events.
//
// Shard file layout (little endian): 64-byte header {char magic[8] = "ARGUSRAW", u32 version = 1, u32 n_cams,
// u32 H, u32 W, u64 n_samples, u64 pose_offset, u64 image_offset, pad}, then n_samples x 7 float32 poses
// [x, y, z, qx, qy, qz, qw], then n_samples x (n_cams x H x W x 3) uint8 images.
#include "../../include/argus_b200.h"
#include "runtime.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace argus {

struct ShardHeader {
  char magic[8];
  uint32_t version, n_cams, H, W;
  uint64_t n_samples, pose_offset, image_offset;
  uint8_t pad[16];
};
static_assert(sizeof(ShardHeader) == 64, "header is 64 bytes");

static inline uint64_t splitmix64(uint64_t& x) {
  uint64_t z = (x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

class Loader {
 public:
  Loader(const char* path, int batch, int rank, int world, uint64_t seed, int shuffle, int drop_last)
      : batch_(batch), rank_(rank), world_(world), seed_(seed), shuffle_(shuffle != 0), drop_last_(drop_last != 0) {
    ARGUS_CHECK(batch > 0 && world > 0 && rank >= 0 && rank < world, "bad loader arguments");
    fd_ = open(path, O_RDONLY);
    ARGUS_CHECK(fd_ >= 0, std::string("cannot open shard file ") + path);
    struct stat st;
    ARGUS_CHECK(fstat(fd_, &st) == 0, "fstat failed");
    size_ = static_cast<size_t>(st.st_size);
    ARGUS_CHECK(size_ >= sizeof(ShardHeader), "shard file too small");
    map_ = static_cast<const uint8_t*>(mmap(nullptr, size_, PROT_READ, MAP_SHARED, fd_, 0));
    ARGUS_CHECK(map_ != MAP_FAILED, "mmap failed");
    std::memcpy(&hdr_, map_, sizeof(hdr_));
    ARGUS_CHECK(std::memcmp(hdr_.magic, "ARGUSRAW", 8) == 0 && hdr_.version == 1, "not an ARGUSRAW v1 shard");
    sample_bytes_ = static_cast<size_t>(hdr_.n_cams) * hdr_.H * hdr_.W * 3;
    ARGUS_CHECK(hdr_.image_offset + hdr_.n_samples * sample_bytes_ <= size_, "truncated shard file");
    madvise(const_cast<uint8_t*>(map_), size_, MADV_SEQUENTIAL);
    ARGUS_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      ARGUS_CUDA(cudaEventCreateWithFlags(&copied_[i], cudaEventDisableTiming));
      ARGUS_CUDA(cudaEventCreateWithFlags(&consumed_[i], cudaEventDisableTiming));
    }
  }

  ~Loader() {
    stop();
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    for (int i = 0; i < 2; ++i) {
      if (copied_[i]) cudaEventDestroy(copied_[i]);
      if (consumed_[i]) cudaEventDestroy(consumed_[i]);
    }
    if (map_ && map_ != MAP_FAILED) munmap(const_cast<uint8_t*>(map_), size_);
    if (fd_ >= 0) close(fd_);
  }

  const ShardHeader& header() const { return hdr_; }

  // DistributedSampler semantics (train.py:154-166): one permutation per epoch shared by all ranks, padded by
  // wrapping so that it divides evenly, rank r takes positions r, r + world, ...
  int64_t samples_per_rank() const { return static_cast<int64_t>((hdr_.n_samples + world_ - 1) / world_); }
  int64_t batches_per_epoch() const {
    const int64_t n = samples_per_rank();
    return drop_last_ ? n / batch_ : (n + batch_ - 1) / batch_;
  }

  // Caller-owned staging: two pinned host buffers and two device buffers for images (batch*sample_bytes) and poses.
  void bind(uint8_t* host_img[2], float* host_pose[2], uint8_t* dev_img[2], float* dev_pose[2]) {
    for (int i = 0; i < 2; ++i) {
      host_img_[i] = host_img[i]; host_pose_[i] = host_pose[i];
      dev_img_[i] = dev_img[i]; dev_pose_[i] = dev_pose[i];
    }
    bound_ = true;
  }

  void start_epoch(int epoch) {
    ARGUS_CHECK(bound_, "loader buffers are not bound");
    stop();
    build_indices(epoch);
    next_fill_ = 0;
    next_take_ = 0;
    for (int i = 0; i < 2; ++i) { filled_[i] = false; fill_batch_[i] = -1; }
    consumed_recorded_[0] = consumed_recorded_[1] = false;
    copied_recorded_[0] = copied_recorded_[1] = false;
    quit_ = false;
    worker_ = std::thread([this] { this->run(); });
  }

  // When is a device buffer free to be overwritten by the copy of batch k + 2? Once everything that reads batch k has run.
  //  * fetch-then-step loops (fetch k, step k, fetch k + 1, ...): at call k + 1 all work on batch k is already
  //    enqueued on `stream`, so the "consumed" event of buffer k is recorded THEN and the copy of batch k + 2 overlaps
  //    step k + 1 (default);
  //  * look-ahead loops (fetch k + 1, step k, prefetch k + 1, fetch k + 2, ... as argus_b200/train.py runs): at call
  //    k + 1 step k has NOT been enqueued yet, so the event is recorded at call k + 2, right before the copy that reuses
  //    the buffer (by then step k and the side-stream staging of batch k, which step k waited for, are enqueued).
  void set_lookahead(bool on) { lookahead_ = on; }

  // Returns the number of samples in the batch (0 at the end of the epoch). The device buffers are valid on
  // `stream` after this call; a buffer is overwritten by the call after next (see set_lookahead for what that call
  // waits for).
  int next(cudaStream_t stream, int* buf_index) {
    const int64_t nb = batches_per_epoch();
    if (next_take_ >= nb) return 0;
    const int buf = static_cast<int>(next_take_ & 1);
    int count = 0;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return filled_[buf] && fill_batch_[buf] == next_take_; });
      count = fill_count_[buf];
    }
    // the device buffer may still be read by the step that consumed it two batches ago
    if (lookahead_ && next_take_ >= 2) {
      ARGUS_CUDA(cudaEventRecord(consumed_[buf], stream)); pdl_break(stream, kPdlAfterRecord);
      consumed_recorded_[buf] = true;
    }
    if (consumed_recorded_[buf]) ARGUS_CUDA(cudaStreamWaitEvent(copy_stream_, consumed_[buf], 0)); pdl_break(copy_stream_, kPdlAfterWait);
    ARGUS_CUDA(cudaMemcpyAsync(dev_img_[buf], host_img_[buf], static_cast<size_t>(count) * sample_bytes_,
                               cudaMemcpyHostToDevice, copy_stream_)); pdl_break(copy_stream_, kPdlAfterMemop);
    ARGUS_CUDA(cudaMemcpyAsync(dev_pose_[buf], host_pose_[buf], static_cast<size_t>(count) * 7 * sizeof(float),
                               cudaMemcpyHostToDevice, copy_stream_)); pdl_break(copy_stream_, kPdlAfterMemop);
    ARGUS_CUDA(cudaEventRecord(copied_[buf], copy_stream_)); pdl_break(copy_stream_, kPdlAfterRecord);
    copied_recorded_[buf] = true;
    ARGUS_CUDA(cudaStreamWaitEvent(stream, copied_[buf], 0)); pdl_break(stream, kPdlAfterWait);
    // everything the caller enqueues on `stream` until the next call consumes this buffer
    const int prev = buf ^ 1;
    if (!lookahead_ && next_take_ > 0) {
      ARGUS_CUDA(cudaEventRecord(consumed_[prev], stream)); pdl_break(stream, kPdlAfterRecord);
      consumed_recorded_[prev] = true;
    }
    {
      // hand the pinned buffer back to the worker once its H2D copy has been issued; the worker waits on the
      // copy event before overwriting it
      std::lock_guard<std::mutex> lk(mu_);
      filled_[buf] = false;
      host_busy_[buf] = true;
    }
    cv_.notify_all();
    *buf_index = buf;
    ++next_take_;
    return count;
  }

  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_.notify_all();
    if (worker_.joinable()) worker_.join();
  }

 private:
  void build_indices(int epoch) {
    const uint64_t n = hdr_.n_samples;
    std::vector<uint64_t> perm(n);
    for (uint64_t i = 0; i < n; ++i) perm[i] = i;
    if (shuffle_) {
      uint64_t st = seed_ * 0x2545F4914F6CDD1Dull + static_cast<uint64_t>(epoch) + 1;
      for (uint64_t i = n; i > 1; --i) {
        const uint64_t j = splitmix64(st) % i;
        std::swap(perm[i - 1], perm[j]);
      }
    }
    const uint64_t per = static_cast<uint64_t>(samples_per_rank());
    indices_.resize(per);
    for (uint64_t k = 0; k < per; ++k) indices_[k] = perm[(k * world_ + rank_) % n];
  }

  void run() {
    const int64_t nb = batches_per_epoch();
    const float* poses = reinterpret_cast<const float*>(map_ + hdr_.pose_offset);
    const uint8_t* images = map_ + hdr_.image_offset;
    for (int64_t b = 0; b < nb; ++b) {
      const int buf = static_cast<int>(b & 1);
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return quit_ || !filled_[buf]; });
        if (quit_) return;
      }
      if (host_busy_[buf]) {
        // the previous H2D copy out of this pinned buffer must have finished
        cudaEventSynchronize(copied_[buf]);
        host_busy_[buf] = false;
      }
      const int64_t lo = b * batch_;
      const int64_t hi = std::min<int64_t>(lo + batch_, static_cast<int64_t>(indices_.size()));
      for (int64_t k = lo; k < hi; ++k) {
        const uint64_t idx = indices_[k];
        std::memcpy(host_img_[buf] + static_cast<size_t>(k - lo) * sample_bytes_, images + idx * sample_bytes_,
                    sample_bytes_);
        std::memcpy(host_pose_[buf] + static_cast<size_t>(k - lo) * 7, poses + idx * 7, 7 * sizeof(float));
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        filled_[buf] = true;
        fill_batch_[buf] = b;
        fill_count_[buf] = static_cast<int>(hi - lo);
      }
      cv_.notify_all();
    }
  }

  int batch_, rank_, world_;
  uint64_t seed_;
  bool shuffle_, drop_last_;
  int fd_ = -1;
  size_t size_ = 0, sample_bytes_ = 0;
  const uint8_t* map_ = nullptr;
  ShardHeader hdr_{};
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t copied_[2] = {nullptr, nullptr}, consumed_[2] = {nullptr, nullptr};
  bool consumed_recorded_[2] = {false, false}, copied_recorded_[2] = {false, false};
  bool lookahead_ = false;
  uint8_t* host_img_[2] = {nullptr, nullptr};
  float* host_pose_[2] = {nullptr, nullptr};
  uint8_t* dev_img_[2] = {nullptr, nullptr};
  float* dev_pose_[2] = {nullptr, nullptr};
  bool bound_ = false;
  std::vector<uint64_t> indices_;
  std::thread worker_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool filled_[2] = {false, false};
  std::atomic<bool> host_busy_[2] = {{false}, {false}};
  int64_t fill_batch_[2] = {-1, -1};
  int fill_count_[2] = {0, 0};
  int64_t next_fill_ = 0, next_take_ = 0;
  bool quit_ = false;
};

}  // namespace argus

using namespace argus;

struct argus_loader {
  Loader impl;
  argus_loader(const char* path, int batch, int rank, int world, uint64_t seed, int shuffle, int drop_last)
      : impl(path, batch, rank, world, seed, shuffle, drop_last) {}
};

extern "C" {

int argus_loader_create(argus_loader** out, const char* path, int batch, int rank, int world, uint64_t seed,
                        int shuffle, int drop_last) {
  ARGUS_API_BEGIN
  ARGUS_CHECK(out != nullptr && path != nullptr, "null argument");
  *out = new argus_loader(path, batch, rank, world, seed, shuffle, drop_last);
  ARGUS_API_END
}
int argus_loader_destroy(argus_loader* l) {
  ARGUS_API_BEGIN
  delete l;
  ARGUS_API_END
}
int argus_loader_info(argus_loader* l, int64_t* n_samples, int* n_cams, int* H, int* W, int64_t* samples_per_rank,
                      int64_t* batches_per_epoch) {
  ARGUS_API_BEGIN
  ARGUS_CHECK(l != nullptr, "null loader");
  const ShardHeader& h = l->impl.header();
  if (n_samples) *n_samples = static_cast<int64_t>(h.n_samples);
  if (n_cams) *n_cams = static_cast<int>(h.n_cams);
  if (H) *H = static_cast<int>(h.H);
  if (W) *W = static_cast<int>(h.W);
  if (samples_per_rank) *samples_per_rank = l->impl.samples_per_rank();
  if (batches_per_epoch) *batches_per_epoch = l->impl.batches_per_epoch();
  ARGUS_API_END
}
int argus_loader_bind(argus_loader* l, void* host_img0, void* host_img1, float* host_pose0, float* host_pose1,
                      void* dev_img0, void* dev_img1, float* dev_pose0, float* dev_pose1) {
  ARGUS_API_BEGIN
  ARGUS_CHECK(l != nullptr, "null loader");
  uint8_t* hi[2] = {static_cast<uint8_t*>(host_img0), static_cast<uint8_t*>(host_img1)};
  float* hp[2] = {host_pose0, host_pose1};
  uint8_t* di[2] = {static_cast<uint8_t*>(dev_img0), static_cast<uint8_t*>(dev_img1)};
  float* dp[2] = {dev_pose0, dev_pose1};
  for (int i = 0; i < 2; ++i) ARGUS_CHECK(hi[i] && hp[i] && di[i] && dp[i], "null staging buffer");
  l->impl.bind(hi, hp, di, dp);
  ARGUS_API_END
}
int argus_loader_set_lookahead(argus_loader* l, int on) {
  ARGUS_API_BEGIN
  ARGUS_CHECK(l != nullptr, "null loader");
  l->impl.set_lookahead(on != 0);
  ARGUS_API_END
}
int argus_loader_start_epoch(argus_loader* l, int epoch) {
  ARGUS_API_BEGIN
  ARGUS_CHECK(l != nullptr, "null loader");
  l->impl.start_epoch(epoch);
  ARGUS_API_END
}
int argus_loader_next(argus_loader* l, void* stream, int* count, int* buffer_index) {
  ARGUS_API_BEGIN
  ARGUS_CHECK(l != nullptr && count != nullptr, "null argument");
  int buf = -1;
  *count = l->impl.next(static_cast<cudaStream_t>(stream), &buf);
  if (buffer_index) *buffer_index = buf;
  ARGUS_API_END
}

}  // extern "C"
