// C ABI of libspectre_b200.so: context, NTT, EvaluationDomain. (MSM entry points are in msm.cu, batch
// polynomial ops in poly.cu.) See include/spectre_b200.h for the contract of every function.
#include "common.cuh"
#include "ntt.cuh"
#include "../../include/spectre_b200.h"
#include <stdarg.h>
#include <stdio.h>
#include <sys/types.h>
#include <stdlib.h>
#include <string.h>

using namespace spb;

static_assert(sizeof(spb_fr) == sizeof(Fr) && sizeof(spb_g1_affine) == sizeof(G1Affine) && sizeof(spb_g1) == sizeof(G1Jac),
              "C ABI structs must match the device structs byte for byte");

namespace spb {

int set_error(spb_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (ctx) ctx->last_error = buf;
  return code;
}

void* slot(spb_ctx* ctx, DeviceState& d, const char* name, size_t bytes) {
  DevBuf& b = d.slots[name];
  if (b.cap >= bytes && b.ptr) return b.ptr;
  if (b.ptr) { cudaStreamSynchronize(d.stream); cudaFree(b.ptr); b.ptr = nullptr; b.cap = 0; }
  size_t want = bytes + bytes / 8;  // a little headroom so growing sizes do not realloc every call
  cudaError_t e = cudaMalloc(&b.ptr, want);
  if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&b.ptr, want); }
  if (e != cudaSuccess) { set_error(ctx, SPB_ERR_OOM, "cudaMalloc(%zu) for slot %s: %s", want, name, cudaGetErrorString(e)); b.ptr = nullptr; return nullptr; }
  b.cap = want;
  return b.ptr;
}

// Double-buffered staging: two pinned 16 MiB buffers per device (slot-like, allocated once). While the DMA of one buffer is
// in flight (cudaMemcpyAsync + an event), the host fills / drains the other, so the file system and the PCIe copy overlap
// instead of alternating as a synchronous cudaMemcpy loop does. (cuFile / GDS would remove the bounce buffer altogether; the
// pool's boxes expose no nvidia-fs, where cuFile itself falls back to exactly this scheme.)
static const size_t kStageBytes = (size_t)16 << 20;
static int stage_buffers(spb_ctx* ctx, DeviceState& d, char** a, char** b, cudaEvent_t* ea, cudaEvent_t* eb) {
  if (!d.stage[0]) {
    SPB_CUDA(ctx, cudaMallocHost(&d.stage[0], kStageBytes));
    SPB_CUDA(ctx, cudaMallocHost(&d.stage[1], kStageBytes));
    SPB_CUDA(ctx, cudaEventCreateWithFlags(&d.stage_done[0], cudaEventDisableTiming));
    SPB_CUDA(ctx, cudaEventCreateWithFlags(&d.stage_done[1], cudaEventDisableTiming));
  }
  *a = (char*)d.stage[0]; *b = (char*)d.stage[1]; *ea = d.stage_done[0]; *eb = d.stage_done[1];
  return 0;
}
int stream_file_to_device(spb_ctx* ctx, DeviceState& d, FILE* f, void* d_dst, size_t bytes, const char* what) {
  char* buf[2]; cudaEvent_t ev[2];
  SPB_TRY(stage_buffers(ctx, d, &buf[0], &buf[1], &ev[0], &ev[1]));
  bool busy[2] = {false, false};
  int cur = 0;
  for (size_t off = 0; off < bytes; off += kStageBytes, cur ^= 1) {
    const size_t cnt = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    if (busy[cur]) SPB_CUDA(ctx, cudaEventSynchronize(ev[cur]));          // the DMA that last used this buffer has drained it
    if (fread(buf[cur], 1, cnt, f) != cnt) return set_error(ctx, SPB_ERR_ARG, "%s: file is truncated", what);
    SPB_CUDA(ctx, cudaMemcpyAsync((char*)d_dst + off, buf[cur], cnt, cudaMemcpyHostToDevice, d.stream));
    SPB_CUDA(ctx, cudaEventRecord(ev[cur], d.stream));
    busy[cur] = true;
  }
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}
int stream_device_to_file(spb_ctx* ctx, DeviceState& d, FILE* f, const void* d_src, size_t bytes, const char* what) {
  char* buf[2]; cudaEvent_t ev[2];
  SPB_TRY(stage_buffers(ctx, d, &buf[0], &buf[1], &ev[0], &ev[1]));
  size_t pending_cnt[2] = {0, 0};
  int cur = 0;
  for (size_t off = 0; off < bytes || pending_cnt[0] || pending_cnt[1]; cur ^= 1) {
    if (pending_cnt[cur]) {                                               // drain the buffer whose D2H was issued two steps ago
      SPB_CUDA(ctx, cudaEventSynchronize(ev[cur]));
      if (fwrite(buf[cur], 1, pending_cnt[cur], f) != pending_cnt[cur]) return set_error(ctx, SPB_ERR_ARG, "%s: short write", what);
      pending_cnt[cur] = 0;
    }
    if (off < bytes) {
      const size_t cnt = bytes - off < kStageBytes ? bytes - off : kStageBytes;
      SPB_CUDA(ctx, cudaMemcpyAsync(buf[cur], (const char*)d_src + off, cnt, cudaMemcpyDeviceToHost, d.stream));
      SPB_CUDA(ctx, cudaEventRecord(ev[cur], d.stream));
      pending_cnt[cur] = cnt; off += cnt;
    }
  }
  return 0;
}

}  // namespace spb

struct spb_domain {
  uint32_t j, k, extended_k, quotient_poly_degree;
  Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor;
  uint32_t t_len;
  Fr* d_t_evaluations;  // device, t_len values
  int device;
};

template <class P>
__global__ void field_op_kernel(int op, const Fp<P>* a, const Fp<P>* b, Fp<P>* o, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<P> x = a[i], y = b[i], r;
  if (op == 0) r = fp_mul(x, y);
  else if (op == 1) r = fp_add(x, y);
  else r = fp_sub(x, y);
  o[i] = r;
}

// ILP independent multiply chains per thread; result folded and written so nothing is optimised away
template <class P, int ILP, bool SQR = false>
__global__ void modmul_bench_kernel(Fp<P>* out, uint32_t iters) {
  Fp<P> x[ILP], y;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int j = 0; j < ILP; j++) { x[j] = fp_one<P>(); x[j].l[0] ^= tid * 2654435761u + j; x[j].l[7] &= 0x0fffffffu; }
  y = x[0]; y.l[1] ^= 0x9e3779b9u;
  for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = SQR ? fp_sqr(x[j]) : fp_mul(x[j], y);
  }
  Fp<P> acc = x[0];
#pragma unroll
  for (int j = 1; j < ILP; j++) acc = fp_add(acc, x[j]);
  out[tid] = acc;
}

// Raw pipe-rate probes (8 independent chains per thread): 0 = IMAD.WIDE.U32, 1 = IMAD (lo), 2 = DFMA,
// 3 = IMAD.WIDE + DFMA interleaved 1:1, 4 = IADD3, 5 = IMAD.WIDE + IADD3 interleaved 1:1
template <int KIND>
__global__ void pipe_probe_kernel(unsigned long long* out, uint32_t iters, uint32_t seed) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long w[8]; uint32_t v[8]; double f[8];
  uint32_t a = tid * 2654435761u + seed, b = a ^ 0x9e3779b9u;
  double fa = 1.0 + (double)(tid & 1023) * 1e-9, fb = 0.999999 + (double)(seed & 7) * 1e-9;
#pragma unroll
  for (int j = 0; j < 8; j++) { w[j] = a + j; v[j] = b + j; f[j] = fa + j; }
  for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      // one multiplicand is the chain's own low word so the product is not loop-invariant
      if (KIND == 0 || KIND == 3 || KIND == 5) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(w[j]) : "r"(b));
      if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(a), "r"(b));
      if (KIND == 2 || KIND == 3) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(f[j]) : "d"(fb), "d"(fa));
      if (KIND == 4 || KIND == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(b));
    }
  }
  unsigned long long acc = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) acc += w[j] + v[j] + (unsigned long long)__double_as_longlong(f[j]);
  out[tid] = acc;
}

extern "C" {

int spb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

spb_ctx* spb_init(const int* device_ids, int n_dev) {
  int avail = spb_device_count();
  if (n_dev <= 0) n_dev = 1;
  if (avail <= 0) { fprintf(stderr, "spectre_b200: no CUDA device visible; this library has no CPU fallback\n"); return nullptr; }
  spb_ctx* ctx = new spb_ctx();
  for (int i = 0; i < n_dev; i++) {
    int id = device_ids ? device_ids[i] : i;
    if (id < 0 || id >= avail) { fprintf(stderr, "spectre_b200: device id %d out of range (have %d)\n", id, avail); delete ctx; return nullptr; }
    DeviceState d; d.device = id;
    if (cudaSetDevice(id) != cudaSuccess) { delete ctx; return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, id) != cudaSuccess) { delete ctx; return nullptr; }
    d.sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return nullptr; }
    cudaEventCreate(&d.ev0); cudaEventCreate(&d.ev1);
    cudaEventCreateWithFlags(&d.dep_ev, cudaEventDisableTiming);
    for (int e = 0; e < 8; e++) cudaEventCreate(&d.stage_ev[e]);
    d.pinned_cap = 1 << 20;
    if (cudaMallocHost(&d.pinned, d.pinned_cap) != cudaSuccess) { delete ctx; return nullptr; }
    ctx->dev.push_back(d);
  }
  // several devices: direct NVLink peer copies for the NTT all-to-all and the sharded-MSM scalar scatter
  ctx->peer_access = ctx->dev.size() > 1;
  for (size_t i = 0; i < ctx->dev.size(); i++)
    for (size_t j = 0; j < ctx->dev.size(); j++) {
      if (i == j) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, ctx->dev[i].device, ctx->dev[j].device);
      if (!can) { ctx->peer_access = false; continue; }
      cudaSetDevice(ctx->dev[i].device);
      cudaError_t e = cudaDeviceEnablePeerAccess(ctx->dev[j].device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ctx->peer_access = false;
      cudaGetLastError();
    }
  if (!ctx->dev.empty()) cudaSetDevice(ctx->dev[0].device);
  return ctx;
}

void spb_shutdown(spb_ctx* ctx) {
  if (!ctx) return;
  msm_release_ctx(ctx);
  for (auto& d : ctx->dev) {
    cudaSetDevice(d.device);
    cudaStreamSynchronize(d.stream);
    for (auto& kv : d.slots) if (kv.second.ptr) cudaFree(kv.second.ptr);
    for (auto& t : d.ntt_tables) { cudaFree(t.tw_lo); cudaFree(t.tw_hi); if (t.tw_full) cudaFree(t.tw_full); }
    if (d.pinned) cudaFreeHost(d.pinned);
    for (int i = 0; i < 2; i++) { if (d.stage[i]) cudaFreeHost(d.stage[i]); if (d.stage_done[i]) cudaEventDestroy(d.stage_done[i]); }
    cudaEventDestroy(d.ev0); cudaEventDestroy(d.ev1); cudaEventDestroy(d.dep_ev);
    for (int e = 0; e < 8; e++) cudaEventDestroy(d.stage_ev[e]);
    cudaStreamDestroy(d.stream);
  }
  delete ctx;
}

const char* spb_last_error(spb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
uint64_t spb_kernel_launches(spb_ctx* ctx) { return ctx ? ctx->n_kernel_launches : 0; }
float spb_last_device_ms(spb_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.f; }

void* spb_stream(spb_ctx* ctx, int dev_index) {
  if (!ctx || dev_index < 0 || (size_t)dev_index >= ctx->dev.size()) return nullptr;
  return (void*)ctx->dev[dev_index].stream;
}

// ---- file <-> device (params / proving-key files: SURVEY.md 8f rank 4) -----------------------------------------------------
int spb_read_file_dev(spb_ctx* ctx, const char* path, uint64_t offset, void* d_dst, size_t bytes) {
  if (!ctx || !path || (bytes && !d_dst)) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  FILE* f = fopen(path, "rb");
  if (!f) return set_error(ctx, SPB_ERR_ARG, "spb_read_file_dev: cannot open %s", path);
  int rc = fseeko(f, (off_t)offset, SEEK_SET) == 0 ? stream_file_to_device(ctx, d, f, d_dst, bytes, "spb_read_file_dev") : set_error(ctx, SPB_ERR_ARG, "spb_read_file_dev: seek failed");
  fclose(f);
  return rc;
}
int spb_write_file_dev(spb_ctx* ctx, const char* path, int append, const void* d_src, size_t bytes) {
  if (!ctx || !path || (bytes && !d_src)) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  FILE* f = fopen(path, append ? "ab" : "wb");
  if (!f) return set_error(ctx, SPB_ERR_ARG, "spb_write_file_dev: cannot open %s", path);
  int rc = stream_device_to_file(ctx, d, f, d_src, bytes, "spb_write_file_dev");
  if (fclose(f) != 0 && rc == 0) rc = set_error(ctx, SPB_ERR_ARG, "spb_write_file_dev: close failed");
  return rc;
}

int spb_host_register(spb_ctx* ctx, void* ptr, size_t bytes) {
  SPB_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return 0;
}
int spb_host_unregister(spb_ctx* ctx, void* ptr) {
  SPB_CUDA(ctx, cudaHostUnregister(ptr));
  return 0;
}

// host-only fold of partial MSM results (multi-rank all-gather + local add; EC addition is not an NCCL op)
int spb_g1_sum(const spb_g1* pts, size_t n, spb_g1* out) {
  if (!out || (n && !pts)) return SPB_ERR_ARG;
  G1Xyzz acc = xyzz_identity();
  for (size_t i = 0; i < n; i++) { G1Jac j; memcpy(&j, &pts[i], sizeof j); xyzz_add(acc, xyzz_from_jac(j)); }
  G1Jac r = jac_from_affine(xyzz_to_affine(acc));
  memcpy(out, &r, sizeof r);
  return 0;
}

// out[i] = sum_g pts[g * count + i]: the fold of a whole batch of sharded MSMs after ONE all-gather (groups = ranks)
int spb_g1_sum_batch(const spb_g1* pts, size_t groups, size_t count, spb_g1* out) {
  if (!out || (groups && count && !pts)) return SPB_ERR_ARG;
  for (size_t i = 0; i < count; i++) {
    G1Xyzz acc = xyzz_identity();
    for (size_t g = 0; g < groups; g++) { G1Jac j; memcpy(&j, &pts[g * count + i], sizeof j); xyzz_add(acc, xyzz_from_jac(j)); }
    G1Jac r = jac_from_affine(xyzz_to_affine(acc));
    memcpy(&out[i], &r, sizeof r);
  }
  return 0;
}

// ---- NTT ---------------------------------------------------------------------------------------------------
static int ntt_timed(spb_ctx* ctx, DeviceState& d, const Fr* src, Fr* dst, uint32_t k, const Fr& omega, const NttOpts& o) {
  SPB_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
  SPB_TRY(ntt_device(ctx, d, src, dst, k, omega, o));
  SPB_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  SPB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, d.ev0, d.ev1));
  return 0;
}

int spb_ntt_dev(spb_ctx* ctx, spb_fr* d_a, uint32_t log_n, const spb_fr* omega) {
  if (!ctx || !d_a || !omega) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  Fr w; memcpy(&w, omega, 32);
  return ntt_timed(ctx, d, (const Fr*)d_a, (Fr*)d_a, log_n, w, NttOpts());
}

// host-buffer transform through a device staging slot
static int ntt_host(spb_ctx* ctx, const Fr* in, size_t n_in_copy, Fr* out, size_t n_out_copy, uint32_t k, const Fr& omega, const NttOpts& o) {
  if (ntt_multi_applicable(ctx, k)) {
    NttOpts o2 = o;
    if (!o2.n_in) o2.n_in = n_in_copy;
    if (!o2.n_out) o2.n_out = n_out_copy;
    return ntt_multi_host(ctx, in, out, k, omega, o2, nullptr);
  }
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  size_t n = (size_t)1 << k;
  Fr* buf = (Fr*)slot(ctx, d, "ntt_io", n * sizeof(Fr));
  if (!buf) return SPB_ERR_OOM;
  SPB_CUDA(ctx, cudaMemcpyAsync(buf, in, n_in_copy * sizeof(Fr), cudaMemcpyHostToDevice, d.stream));
  SPB_TRY(ntt_timed(ctx, d, buf, buf, k, omega, o));
  SPB_CUDA(ctx, cudaMemcpyAsync(out, buf, n_out_copy * sizeof(Fr), cudaMemcpyDeviceToHost, d.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}

int spb_ntt(spb_ctx* ctx, spb_fr* a, uint32_t log_n, const spb_fr* omega) {
  if (!ctx || !a || !omega) return SPB_ERR_ARG;
  if (log_n > 28) return set_error(ctx, SPB_ERR_ARG, "spb_ntt: log_n %u > 28", log_n);
  std::lock_guard<std::mutex> lk(ctx->mu);
  Fr w; memcpy(&w, omega, 32);
  size_t n = (size_t)1 << log_n;
  return ntt_host(ctx, (const Fr*)a, n, (Fr*)a, n, log_n, w, NttOpts());
}

// ---- EvaluationDomain --------------------------------------------------------------------------------------
__global__ void vanishing_table_kernel(Fr* t, Fr cur0, Fr step, uint32_t len) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  Fr cur = fp_mul(cur0, fp_pow_u64(step, i));
  t[i] = fp_inv(fp_sub(cur, fp_one<FrParams>()));
}
__global__ void mul_periodic_kernel(Fr* a, const Fr* t, uint64_t n, uint32_t mask) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v = ntt_ld_stream(a + i);
  ntt_stg(a + i, fp_mul(v, ntt_ldg(t + (i & mask))));
}

int spb_domain_new(spb_ctx* ctx, uint32_t j, uint32_t k, spb_domain** out) {
  if (!ctx || !out || j < 2) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  spb_domain* dm = new spb_domain();
  dm->j = j; dm->k = k; dm->quotient_poly_degree = j - 1;
  uint64_t n = 1ull << k;
  uint32_t ek = k;
  while ((1ull << ek) < n * dm->quotient_poly_degree) ek++;
  if (ek > 28) { delete dm; return set_error(ctx, SPB_ERR_ARG, "domain: extended_k %u > 28", ek); }
  dm->extended_k = ek;
  Fr w; { constexpr uint32_t v[8] = SPB_FR_ROOT_OF_UNITY_MONT; for (int i = 0; i < 8; i++) w.l[i] = v[i]; }
  for (uint32_t i = ek; i < 28; i++) w = fp_sqr(w);
  dm->extended_omega = w; dm->extended_omega_inv = fp_inv(w);
  for (uint32_t i = k; i < ek; i++) w = fp_sqr(w);
  dm->omega = w; dm->omega_inv = fp_inv(w);
  { constexpr uint32_t v[8] = SPB_FR_ZETA_MONT; for (int i = 0; i < 8; i++) dm->g_coset.l[i] = v[i]; }
  dm->g_coset_inv = fp_sqr(dm->g_coset);
  dm->ifft_divisor = fp_inv(fr_from_u64(n));
  dm->extended_ifft_divisor = fp_inv(fr_from_u64(1ull << ek));
  dm->t_len = 1u << (ek - k);
  DeviceState& d = ctx->dev[0];
  dm->device = d.device;
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  SPB_CUDA(ctx, cudaMalloc(&dm->d_t_evaluations, dm->t_len * sizeof(Fr)));
  Fr cur0 = fp_pow_u64(dm->g_coset, n), step = fp_pow_u64(dm->extended_omega, n);
  vanishing_table_kernel<<<(dm->t_len + 63) / 64, 64, 0, d.stream>>>(dm->d_t_evaluations, cur0, step, dm->t_len);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  *out = dm;
  return 0;
}

void spb_domain_free(spb_ctx* ctx, spb_domain* dm) {
  if (!dm) return;
  if (ctx) { std::lock_guard<std::mutex> lk(ctx->mu); cudaSetDevice(dm->device); cudaFree(dm->d_t_evaluations); }
  delete dm;
}
uint32_t spb_domain_extended_k(const spb_domain* d) { return d ? d->extended_k : 0; }
void spb_domain_constants(const spb_domain* d, spb_fr out[8]) {
  if (!d || !out) return;
  const Fr* src[8] = {&d->omega, &d->omega_inv, &d->extended_omega, &d->extended_omega_inv, &d->g_coset, &d->g_coset_inv, &d->ifft_divisor, &d->extended_ifft_divisor};
  for (int i = 0; i < 8; i++) memcpy(&out[i], src[i], 32);
}

static void l2c_opts(const spb_domain* dm, Fr post[3], NttOpts& o) { for (int i = 0; i < 3; i++) post[i] = dm->ifft_divisor; o.post3 = post; }
static void c2e_opts(const spb_domain* dm, Fr pre[3], NttOpts& o) {
  pre[0] = fp_one<FrParams>(); pre[1] = dm->g_coset; pre[2] = dm->g_coset_inv;
  o.pre3 = pre; o.n_in = 1ull << dm->k;
}
static void e2c_opts(const spb_domain* dm, Fr post[3], NttOpts& o) {
  post[0] = dm->extended_ifft_divisor;
  post[1] = fp_mul(dm->extended_ifft_divisor, dm->g_coset_inv);
  post[2] = fp_mul(dm->extended_ifft_divisor, dm->g_coset);
  o.post3 = post; o.n_out = (1ull << dm->k) * dm->quotient_poly_degree;
}

int spb_lagrange_to_coeff(spb_ctx* ctx, const spb_domain* dm, spb_fr* a) {
  if (!ctx || !dm || !a) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Fr post[3]; NttOpts o; l2c_opts(dm, post, o);
  size_t n = (size_t)1 << dm->k;
  return ntt_host(ctx, (const Fr*)a, n, (Fr*)a, n, dm->k, dm->omega_inv, o);
}
int spb_coeff_to_lagrange(spb_ctx* ctx, const spb_domain* dm, spb_fr* a) {
  if (!ctx || !dm || !a) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  size_t n = (size_t)1 << dm->k;
  return ntt_host(ctx, (const Fr*)a, n, (Fr*)a, n, dm->k, dm->omega, NttOpts());
}
int spb_coeff_to_extended(spb_ctx* ctx, const spb_domain* dm, const spb_fr* in, spb_fr* out) {
  if (!ctx || !dm || !in || !out) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Fr pre[3]; NttOpts o; c2e_opts(dm, pre, o);
  return ntt_host(ctx, (const Fr*)in, (size_t)1 << dm->k, (Fr*)out, (size_t)1 << dm->extended_k, dm->extended_k, dm->extended_omega, o);
}
int spb_extended_to_coeff(spb_ctx* ctx, const spb_domain* dm, const spb_fr* in, spb_fr* out) {
  if (!ctx || !dm || !in || !out) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Fr post[3]; NttOpts o; e2c_opts(dm, post, o);
  return ntt_host(ctx, (const Fr*)in, (size_t)1 << dm->extended_k, (Fr*)out, ((size_t)1 << dm->k) * dm->quotient_poly_degree, dm->extended_k, dm->extended_omega_inv, o);
}
static int div_vanishing_device(spb_ctx* ctx, DeviceState& d, const spb_domain* dm, Fr* d_a) {
  uint64_t e = 1ull << dm->extended_k;
  mul_periodic_kernel<<<(unsigned)((e + 255) / 256), 256, 0, d.stream>>>(d_a, dm->d_t_evaluations, e, dm->t_len - 1);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  return 0;
}
int spb_divide_by_vanishing(spb_ctx* ctx, const spb_domain* dm, spb_fr* a) {
  if (!ctx || !dm || !a) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  size_t e = (size_t)1 << dm->extended_k;
  Fr* buf = (Fr*)slot(ctx, d, "ntt_io", e * sizeof(Fr));
  if (!buf) return SPB_ERR_OOM;
  SPB_CUDA(ctx, cudaMemcpyAsync(buf, a, e * sizeof(Fr), cudaMemcpyHostToDevice, d.stream));
  SPB_TRY(div_vanishing_device(ctx, d, dm, buf));
  SPB_CUDA(ctx, cudaMemcpyAsync(a, buf, e * sizeof(Fr), cudaMemcpyDeviceToHost, d.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}
int spb_lagrange_to_coeff_dev(spb_ctx* ctx, const spb_domain* dm, spb_fr* d_a) {
  if (!ctx || !dm || !d_a) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0]; SPB_CUDA(ctx, cudaSetDevice(d.device));
  Fr post[3]; NttOpts o; l2c_opts(dm, post, o);
  return ntt_timed(ctx, d, (const Fr*)d_a, (Fr*)d_a, dm->k, dm->omega_inv, o);
}
int spb_coeff_to_extended_dev(spb_ctx* ctx, const spb_domain* dm, const spb_fr* d_in, spb_fr* d_out) {
  if (!ctx || !dm || !d_in || !d_out) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0]; SPB_CUDA(ctx, cudaSetDevice(d.device));
  Fr pre[3]; NttOpts o; c2e_opts(dm, pre, o);
  return ntt_timed(ctx, d, (const Fr*)d_in, (Fr*)d_out, dm->extended_k, dm->extended_omega, o);
}
int spb_extended_to_coeff_dev(spb_ctx* ctx, const spb_domain* dm, const spb_fr* d_in, spb_fr* d_out) {
  if (!ctx || !dm || !d_in || !d_out) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0]; SPB_CUDA(ctx, cudaSetDevice(d.device));
  Fr post[3]; NttOpts o; e2c_opts(dm, post, o);
  return ntt_timed(ctx, d, (const Fr*)d_in, (Fr*)d_out, dm->extended_k, dm->extended_omega_inv, o);
}
// `count` transforms with the same options, polynomial i on device i mod D of the context ("shard by polynomial
// for the NTTs"): the buffers stay on the first device; the other devices read the first pass's input and write the last
// pass's output through NVLink peer access, intermediate passes run in their own HBM.
static int ntt_batch_devices(spb_ctx* ctx, const spb_fr* const* d_in, spb_fr* const* d_out, size_t count, uint32_t k, const Fr& omega, const NttOpts& o) {
  DeviceState& d0 = ctx->dev[0];
  uint32_t min_k = 16;   // smaller transforms are launch-bound: first device only (tests lower it: SPB_SHARD_MIN_LOGN)
  if (const char* e = getenv("SPB_SHARD_MIN_LOGN")) { int v = atoi(e); if (v >= 1) min_k = (uint32_t)v; }
  const size_t D = (ctx->peer_access && ctx->dev.size() > 1 && k >= min_k) ? ctx->dev.size() : 1;
  SPB_CUDA(ctx, cudaSetDevice(d0.device));
  SPB_CUDA(ctx, cudaEventRecord(d0.ev0, d0.stream));
  if (D > 1) SPB_CUDA(ctx, cudaEventRecord(d0.dep_ev, d0.stream));
  for (size_t i = 0; i < count; i++) {
    if (!d_in[i] || !d_out[i]) return SPB_ERR_ARG;
    DeviceState& d = ctx->dev[i % D];
    SPB_CUDA(ctx, cudaSetDevice(d.device));
    if (i < D && i > 0) SPB_CUDA(ctx, cudaStreamWaitEvent(d.stream, d0.dep_ev, 0));
    SPB_TRY(ntt_device(ctx, d, (const Fr*)d_in[i], (Fr*)d_out[i], k, omega, o));
  }
  for (size_t i = 1; i < D && i < count; i++) {
    DeviceState& d = ctx->dev[i];
    SPB_CUDA(ctx, cudaSetDevice(d.device));
    SPB_CUDA(ctx, cudaEventRecord(d.dep_ev, d.stream));
    SPB_CUDA(ctx, cudaStreamWaitEvent(d0.stream, d.dep_ev, 0));
  }
  SPB_CUDA(ctx, cudaSetDevice(d0.device));
  SPB_CUDA(ctx, cudaEventRecord(d0.ev1, d0.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d0.stream));
  SPB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, d0.ev0, d0.ev1));
  return 0;
}
int spb_lagrange_to_coeff_batch_dev(spb_ctx* ctx, const spb_domain* dm, spb_fr* const* d_a, size_t count) {
  if (!ctx || !dm || (count && !d_a)) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Fr post[3]; NttOpts o; l2c_opts(dm, post, o);
  return ntt_batch_devices(ctx, (const spb_fr* const*)d_a, d_a, count, dm->k, dm->omega_inv, o);
}
int spb_coeff_to_extended_batch_dev(spb_ctx* ctx, const spb_domain* dm, const spb_fr* const* d_in, spb_fr* const* d_out, size_t count) {
  if (!ctx || !dm || (count && (!d_in || !d_out))) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Fr pre[3]; NttOpts o; c2e_opts(dm, pre, o);
  return ntt_batch_devices(ctx, d_in, d_out, count, dm->extended_k, dm->extended_omega, o);
}
int spb_divide_by_vanishing_dev(spb_ctx* ctx, const spb_domain* dm, spb_fr* d_a) {
  if (!ctx || !dm || !d_a) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0]; SPB_CUDA(ctx, cudaSetDevice(d.device));
  SPB_TRY(div_vanishing_device(ctx, d, dm, (Fr*)d_a));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}

// ---- test utilities ----------------------------------------------------------------------------------------
int spb_test_field_op(spb_ctx* ctx, int field, int op, const spb_fr* a, const spb_fr* b, spb_fr* out, size_t n) {
  if (!ctx || !a || !b || !out) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  Fr* buf = (Fr*)slot(ctx, d, "test_io", 3 * n * sizeof(Fr));
  if (!buf) return SPB_ERR_OOM;
  SPB_CUDA(ctx, cudaMemcpyAsync(buf, a, n * 32, cudaMemcpyHostToDevice, d.stream));
  SPB_CUDA(ctx, cudaMemcpyAsync(buf + n, b, n * 32, cudaMemcpyHostToDevice, d.stream));
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (field == 0) field_op_kernel<FrParams><<<blocks, 128, 0, d.stream>>>(op, (const Fr*)buf, (const Fr*)(buf + n), (Fr*)(buf + 2 * n), n);
  else field_op_kernel<FqParams><<<blocks, 128, 0, d.stream>>>(op, (const Fq*)buf, (const Fq*)(buf + n), (Fq*)(buf + 2 * n), n);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaMemcpyAsync(out, buf + 2 * n, n * 32, cudaMemcpyDeviceToHost, d.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}

int spb_bench_pipe(spb_ctx* ctx, int kind, uint32_t threads, uint32_t iters, float* ms) {
  if (!ctx || !ms) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  threads = (threads + 255) / 256 * 256;
  unsigned long long* buf = (unsigned long long*)slot(ctx, d, "test_io", (size_t)threads * 8);
  if (!buf) return SPB_ERR_OOM;
  unsigned blocks = threads / 256;
  SPB_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
  switch (kind) {
    case 0: pipe_probe_kernel<0><<<blocks, 256, 0, d.stream>>>(buf, iters, 1); break;
    case 1: pipe_probe_kernel<1><<<blocks, 256, 0, d.stream>>>(buf, iters, 1); break;
    case 2: pipe_probe_kernel<2><<<blocks, 256, 0, d.stream>>>(buf, iters, 1); break;
    case 3: pipe_probe_kernel<3><<<blocks, 256, 0, d.stream>>>(buf, iters, 1); break;
    case 4: pipe_probe_kernel<4><<<blocks, 256, 0, d.stream>>>(buf, iters, 1); break;
    default: pipe_probe_kernel<5><<<blocks, 256, 0, d.stream>>>(buf, iters, 1); break;
  }
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  SPB_CUDA(ctx, cudaEventElapsedTime(ms, d.ev0, d.ev1));
  return 0;
}

int spb_bench_modmul(spb_ctx* ctx, int field, uint32_t threads, uint32_t iters, int ilp, float* ms) {
  if (!ctx || !ms) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  threads = (threads + 255) / 256 * 256;
  Fr* buf = (Fr*)slot(ctx, d, "test_io", (size_t)threads * sizeof(Fr));
  if (!buf) return SPB_ERR_OOM;
  SPB_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
  unsigned blocks = threads / 256;
  if (ilp & 0x100) {  // squaring chains (two independent ones per thread)
    if (field == 0) modmul_bench_kernel<FrParams, 2, true><<<blocks, 256, 0, d.stream>>>((Fr*)buf, iters);
    else modmul_bench_kernel<FqParams, 2, true><<<blocks, 256, 0, d.stream>>>((Fq*)buf, iters);
  } else if (field == 0) {
    if (ilp == 1) modmul_bench_kernel<FrParams, 1><<<blocks, 256, 0, d.stream>>>((Fr*)buf, iters);
    else if (ilp == 2) modmul_bench_kernel<FrParams, 2><<<blocks, 256, 0, d.stream>>>((Fr*)buf, iters);
    else modmul_bench_kernel<FrParams, 4><<<blocks, 256, 0, d.stream>>>((Fr*)buf, iters);
  } else {
    if (ilp == 1) modmul_bench_kernel<FqParams, 1><<<blocks, 256, 0, d.stream>>>((Fq*)buf, iters);
    else if (ilp == 2) modmul_bench_kernel<FqParams, 2><<<blocks, 256, 0, d.stream>>>((Fq*)buf, iters);
    else modmul_bench_kernel<FqParams, 4><<<blocks, 256, 0, d.stream>>>((Fq*)buf, iters);
  }
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  SPB_CUDA(ctx, cudaEventElapsedTime(ms, d.ev0, d.ev1));
  return 0;
}

}  // extern "C"
