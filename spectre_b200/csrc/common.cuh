// Internal plumbing shared by the translation units of libspectre_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "curve.cuh"
#include "../../include/spectre_b200.h"

namespace spb {


// Grow-only device buffer (workspace slots live for the lifetime of the context: no malloc in the hot path).
struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
};

struct NttTables {
  Fr omega;
  uint32_t k, h;
  Fr* tw_lo = nullptr;  // 2^h
  Fr* tw_hi = nullptr;  // 2^(k-h)
  Fr* tw_full = nullptr;  // 2^k (optional)
};

// One MSM lane of a device: own stream, workspace slots (named "...#lane"), events and pinned result area. Consecutive MSMs of a
// batch cycle over the lanes so the latency-bound tail of one overlaps the sort and accumulation of the next (msm.cu).
struct MsmLane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void* pinned = nullptr;
  bool ready = false;
};
static const int kMaxMsmLanes = 4;

struct DeviceState {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t dep_ev = nullptr;  // recorded on `stream` when other streams (MSM lanes, peer devices) must wait for it
  cudaEvent_t stage_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // MSM stage boundaries
  std::map<std::string, DevBuf> slots;
  MsmLane lanes[kMaxMsmLanes];   // created on first use, released with the context (all access under the context lock)
  std::vector<NttTables> ntt_tables;
  void* pinned = nullptr;  // small pinned staging area for results
  size_t pinned_cap = 0;
  void* stage[2] = {nullptr, nullptr};              // two pinned 16 MiB buffers for file <-> device streaming (lazily allocated)
  cudaEvent_t stage_done[2] = {nullptr, nullptr};
};

}  // namespace spb

struct spb_ctx {
  std::vector<spb::DeviceState> dev;
  std::mutex mu;
  std::string last_error;
  bool peer_access = false;  // every device of the context can load from every other one (NVLink P2P enabled)
  // counters (per-call instrumentation behind the C ABI)
  uint64_t n_kernel_launches = 0;
  float last_kernel_ms = 0.f;
  uint64_t last_msm_adds = 0;  // G1 additions of the last MSM call / batch
  bool shplonk_slots_busy = false;  // the context's SHPLONK workspace slots are held by an open spb_shplonk handle
  float msm_stage_ms[7] = {0, 0, 0, 0, 0, 0, 0};  // count, scan, scatter, accumulate, stitch, segment, window (device 0)
};

namespace spb {

int set_error(spb_ctx* ctx, int code, const char* fmt, ...);
// returns nullptr (and sets the error) on failure
void* slot(spb_ctx* ctx, DeviceState& d, const char* name, size_t bytes);

#define SPB_CUDA(ctx, call)                                                                      \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return spb::set_error(ctx, SPB_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define SPB_TRY(expr)            \
  do {                           \
    int rc_ = (expr);            \
    if (rc_ != 0) return rc_;    \
  } while (0)

// ---- capi.cu: file <-> device streaming through two pinned staging buffers (read of chunk i+1 overlaps the DMA of chunk i) ----
int stream_file_to_device(spb_ctx* ctx, DeviceState& d, FILE* f, void* d_dst, size_t bytes, const char* what);
int stream_device_to_file(spb_ctx* ctx, DeviceState& d, FILE* f, const void* d_src, size_t bytes, const char* what);

// ---- msm.cu ----
void msm_release_ctx(spb_ctx* ctx);

// ---- ntt.cu ----
struct NttOpts {
  uint64_t n_in = 0, n_out = 0;  // 0 = n
  const Fr* pre3 = nullptr;      // host pointers to 3 factors, or nullptr
  const Fr* post3 = nullptr;
};
// d_src/d_dst device pointers (may alias); omega host value (Montgomery)
int ntt_device(spb_ctx* ctx, DeviceState& d, const Fr* d_src, Fr* d_dst, uint32_t log_n, const Fr& omega, const NttOpts& opts);
// several devices in the context: six-step across devices with one all-to-all (host buffers)
bool ntt_multi_applicable(spb_ctx* ctx, uint32_t log_n);
int ntt_multi_host(spb_ctx* ctx, const Fr* in, Fr* out, uint32_t log_n, const Fr& omega, const NttOpts& opts, float* ev_ms);

// ---- poly.cu: device-resident cores (pointers on device d, work enqueued on d.stream, no synchronisation unless noted) ----
int dev_grand_product(spb_ctx* ctx, DeviceState& d, const Fr* da, size_t n, Fr* dz, const Fr& init);  // z[i] = init * prod_{j<i} a[j]
int dev_kate_division(spb_ctx* ctx, DeviceState& d, const Fr* da, size_t n, const Fr& b, Fr* dq);
int dev_batch_invert(spb_ctx* ctx, DeviceState& d, Fr* da, size_t n);
int dev_product_enqueue(spb_ctx* ctx, DeviceState& d, const Fr* da, size_t n, Fr** d_total);  // *d_total: one Fr on device d, valid after the enqueued work
int dev_eval_polynomial(spb_ctx* ctx, DeviceState& d, const Fr* dp, size_t n, const Fr& x, Fr* out_host);  // synchronises

// ---- host field helpers (64-bit path) ----
inline Fr fr_from_u64(uint64_t v) {
  Fr a = fp_zero<FrParams>(); a.l[0] = (uint32_t)v; a.l[1] = (uint32_t)(v >> 32);
  return fp_to_mont(a);
}

}  // namespace spb
