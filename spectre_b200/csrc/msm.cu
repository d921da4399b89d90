// Host side of the MSM and of the device-resident ParamsKZG (SRS) handle. Kernels: msm.cuh.
#define SPB_MSM_KERNELS 1
#include "common.cuh"
#include "msm.cuh"
#include <cub/device/device_scan.cuh>
#include <string.h>

using namespace spb;

struct SrsShard {
  int dev_index = 0;
  size_t start = 0, count = 0;
  G1Affine* g = nullptr;
  G1Affine* g_lagrange = nullptr;
};
struct spb_srs {
  uint32_t k = 0;
  size_t n = 0;
  std::vector<SrsShard> shards;  // one per device of the context, contiguous point ranges
};

static uint64_t g_last_adds = 0;

namespace spb {

// Enqueue one MSM on device `d` (no host synchronisation). Window sums land in d.pinned, followed by the
// 32-bit number of sorted entries.
static int msm_enqueue(spb_ctx* ctx, DeviceState& d, const Fr* d_scalars, const G1Affine* d_bases, uint64_t n, MsmGeom g) {
  const uint64_t nb = (uint64_t)g.W * g.B;
  const uint64_t cap = n * g.W;                 // upper bound on entries
  const uint64_t Tmax = (cap + g.L - 1) / g.L;  // upper bound on chunks
  if (cap >= 0xffffffffull) return set_error(ctx, SPB_ERR_ARG, "msm: %llu entries exceed the 32-bit sort index", (unsigned long long)cap);
  uint32_t* counts = (uint32_t*)slot(ctx, d, "msm_counts", (nb + 1) * 4);
  uint32_t* offsets = (uint32_t*)slot(ctx, d, "msm_offsets", (nb + 1) * 4);
  uint32_t* ent_key = (uint32_t*)slot(ctx, d, "msm_ent_key", (cap + 1) * 4);
  uint32_t* ent_val = (uint32_t*)slot(ctx, d, "msm_ent_val", (cap + 1) * 4);
  G1Xyzz* buckets = (G1Xyzz*)slot(ctx, d, "msm_buckets", nb * sizeof(G1Xyzz));
  uint32_t* head_key = (uint32_t*)slot(ctx, d, "msm_head_key", (Tmax + 1) * 4);
  uint32_t* tail_key = (uint32_t*)slot(ctx, d, "msm_tail_key", (Tmax + 1) * 4);
  G1Xyzz* head = (G1Xyzz*)slot(ctx, d, "msm_head", (Tmax + 1) * sizeof(G1Xyzz));
  G1Xyzz* tail = (G1Xyzz*)slot(ctx, d, "msm_tail", (Tmax + 1) * sizeof(G1Xyzz));
  uint32_t* giant = (uint32_t*)slot(ctx, d, "msm_giant", (Tmax + 2) * 4);  // [0] = count, [1..] = queue
  const uint32_t s = g.B < 16 ? g.B : 16;
  const uint32_t segs = g.B / s;
  G1Xyzz* seg_out = (G1Xyzz*)slot(ctx, d, "msm_seg", (uint64_t)g.W * segs * sizeof(G1Xyzz));
  G1Xyzz* win_out = (G1Xyzz*)slot(ctx, d, "msm_win", (uint64_t)g.W * sizeof(G1Xyzz));
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts, offsets, (int)(nb + 1), d.stream);
  void* scan_tmp = slot(ctx, d, "msm_scan_tmp", scan_bytes ? scan_bytes : 16);
  if (!counts || !offsets || !ent_key || !ent_val || !buckets || !head_key || !tail_key || !head || !tail || !giant || !seg_out || !win_out || !scan_tmp)
    return SPB_ERR_OOM;
  if ((size_t)g.W * sizeof(G1Xyzz) + 16 > d.pinned_cap) return set_error(ctx, SPB_ERR_STATE, "msm: pinned staging too small");

  SPB_CUDA(ctx, cudaMemsetAsync(counts, 0, (nb + 1) * 4, d.stream));
  SPB_CUDA(ctx, cudaMemsetAsync(buckets, 0, nb * sizeof(G1Xyzz), d.stream));
  SPB_CUDA(ctx, cudaMemsetAsync(giant, 0, 4, d.stream));
  const unsigned tb = 256;
  cudaEventRecord(d.stage_ev[0], d.stream);
  msm_count_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, d.stream>>>(n, d_scalars, g, counts);
  cudaEventRecord(d.stage_ev[1], d.stream);
  SPB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, counts, offsets, (int)(nb + 1), d.stream));
  // cursor := offsets (the counters are dead after the scan; reuse their storage)
  SPB_CUDA(ctx, cudaMemcpyAsync(counts, offsets, (nb + 1) * 4, cudaMemcpyDeviceToDevice, d.stream));
  cudaEventRecord(d.stage_ev[2], d.stream);
  msm_scatter_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, d.stream>>>(n, d_scalars, g, counts, ent_key, ent_val);
  const uint32_t* total = offsets + nb;  // number of entries M, resident on the device
  cudaEventRecord(d.stage_ev[3], d.stream);
  msm_accumulate_kernel<<<(unsigned)((Tmax + 127) / 128), 128, 0, d.stream>>>(total, g, ent_key, ent_val, d_bases, buckets, head_key, head, tail_key, tail);
  cudaEventRecord(d.stage_ev[4], d.stream);
  msm_stitch_kernel<<<(unsigned)((Tmax + 127) / 128), 128, 0, d.stream>>>(total, g.L, 24, head_key, head, tail_key, tail, buckets, giant, giant + 1);
  msm_giant_kernel<<<256, 128, 0, d.stream>>>(total, g.L, giant, giant + 1, head_key, head, tail_key, tail, buckets);
  cudaEventRecord(d.stage_ev[5], d.stream);
  msm_segment_kernel<<<(unsigned)(((uint64_t)g.W * segs + 127) / 128), 128, 0, d.stream>>>(g, s, buckets, seg_out);
  cudaEventRecord(d.stage_ev[6], d.stream);
  msm_window_kernel<<<g.W, 128, 0, d.stream>>>(segs, seg_out, win_out);
  cudaEventRecord(d.stage_ev[7], d.stream);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches += 7;
  SPB_CUDA(ctx, cudaMemcpyAsync(d.pinned, win_out, (size_t)g.W * sizeof(G1Xyzz), cudaMemcpyDeviceToHost, d.stream));
  SPB_CUDA(ctx, cudaMemcpyAsync((char*)d.pinned + (size_t)g.W * sizeof(G1Xyzz), total, 4, cudaMemcpyDeviceToHost, d.stream));
  return 0;
}

static G1Xyzz msm_finish(DeviceState& d, MsmGeom g) {
  const G1Xyzz* S = (const G1Xyzz*)d.pinned;
  uint32_t M; memcpy(&M, (const char*)d.pinned + (size_t)g.W * sizeof(G1Xyzz), 4);
  g_last_adds += (uint64_t)M + 2ull * g.W * g.B;
  return msm_combine_windows(S, g.W, g.c);
}

static void write_result(const G1Xyzz& r, spb_g1* out) {
  G1Affine a = xyzz_to_affine(r);
  G1Jac j = jac_from_affine(a);
  memcpy(out, &j, sizeof(G1Jac));
}

struct MsmPart {
  int dev_index;
  const Fr* d_scalars;      // device pointer on that device
  const G1Affine* d_bases;  // device pointer on that device
  uint64_t n;
};

// Run the parts (one per device) concurrently and fold the partial sums on the host.
static int msm_run_parts(spb_ctx* ctx, const std::vector<MsmPart>& parts, spb_g1* out) {
  g_last_adds = 0;
  std::vector<MsmGeom> geoms(parts.size());
  for (size_t i = 0; i < parts.size(); i++) {
    if (!parts[i].n) continue;
    DeviceState& d = ctx->dev[parts[i].dev_index];
    SPB_CUDA(ctx, cudaSetDevice(d.device));
    geoms[i] = msm_choose_geometry(parts[i].n);
    SPB_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
    SPB_TRY(msm_enqueue(ctx, d, parts[i].d_scalars, parts[i].d_bases, parts[i].n, geoms[i]));
    SPB_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
  }
  G1Xyzz acc = xyzz_identity();
  float worst = 0.f;
  for (size_t i = 0; i < parts.size(); i++) {
    if (!parts[i].n) continue;
    DeviceState& d = ctx->dev[parts[i].dev_index];
    SPB_CUDA(ctx, cudaSetDevice(d.device));
    SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
    float ms = 0.f;
    SPB_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
    if (ms > worst) worst = ms;
    if (parts[i].dev_index == 0) for (int e = 0; e < 7; e++) cudaEventElapsedTime(&ctx->msm_stage_ms[e], d.stage_ev[e], d.stage_ev[e + 1]);
    G1Xyzz r = msm_finish(d, geoms[i]);
    xyzz_add(acc, r);
  }
  ctx->last_kernel_ms = worst;
  write_result(acc, out);
  return 0;
}

}  // namespace spb

extern "C" {

uint64_t spb_last_msm_adds(spb_ctx* ctx) { (void)ctx; return g_last_adds; }
void spb_last_msm_stage_ms(spb_ctx* ctx, float out[7]) { for (int i = 0; i < 7; i++) out[i] = ctx ? ctx->msm_stage_ms[i] : 0.f; }
void spb_msm_geometry(size_t n, uint32_t* c, uint32_t* windows) { MsmGeom g = msm_choose_geometry(n ? n : 1); *c = g.c; *windows = g.W; }

// ---- ParamsKZG ---------------------------------------------------------------------------------------------
static spb_srs* srs_alloc(spb_ctx* ctx, uint32_t k) {
  spb_srs* s = new spb_srs();
  s->k = k; s->n = (size_t)1 << k;
  size_t D = ctx->dev.size();
  for (size_t i = 0; i < D; i++) {
    SrsShard sh; sh.dev_index = (int)i;
    sh.start = s->n * i / D; sh.count = s->n * (i + 1) / D - sh.start;
    s->shards.push_back(sh);
  }
  return s;
}

void spb_srs_free(spb_ctx* ctx, spb_srs* srs) {
  if (!srs) return;
  if (ctx) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto& sh : srs->shards) {
      cudaSetDevice(ctx->dev[sh.dev_index].device);
      if (sh.g) cudaFree(sh.g);
      if (sh.g_lagrange) cudaFree(sh.g_lagrange);
    }
  }
  delete srs;
}

int spb_srs_upload(spb_ctx* ctx, uint32_t k, const spb_g1_affine* g, const spb_g1_affine* g_lagrange, spb_srs** out) {
  if (!ctx || !out || k > 28) return SPB_ERR_ARG;
  spb_srs* s;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    s = srs_alloc(ctx, k);
    for (auto& sh : s->shards) {
      DeviceState& d = ctx->dev[sh.dev_index];
      cudaError_t e = cudaSetDevice(d.device);
      if (e == cudaSuccess && g && sh.count) {
        e = cudaMalloc(&sh.g, sh.count * sizeof(G1Affine));
        if (e == cudaSuccess) e = cudaMemcpyAsync(sh.g, (const G1Affine*)g + sh.start, sh.count * sizeof(G1Affine), cudaMemcpyHostToDevice, d.stream);
      }
      if (e == cudaSuccess && g_lagrange && sh.count) {
        e = cudaMalloc(&sh.g_lagrange, sh.count * sizeof(G1Affine));
        if (e == cudaSuccess) e = cudaMemcpyAsync(sh.g_lagrange, (const G1Affine*)g_lagrange + sh.start, sh.count * sizeof(G1Affine), cudaMemcpyHostToDevice, d.stream);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
      if (e != cudaSuccess) { set_error(ctx, SPB_ERR_CUDA, "spb_srs_upload: %s", cudaGetErrorString(e)); goto fail; }
    }
    *out = s;
    return 0;
  }
fail:
  spb_srs_free(ctx, s);
  return SPB_ERR_CUDA;
}

int spb_srs_setup(spb_ctx* ctx, uint32_t k, const spb_fr* secret, spb_srs** out) {
  if (!ctx || !out || !secret || k > 28) return SPB_ERR_ARG;
  spb_srs* s;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    s = srs_alloc(ctx, k);
    Fr tau; memcpy(&tau, secret, 32);
    const uint64_t n = 1ull << k;
    Fr w; { constexpr uint32_t v[8] = SPB_FR_ROOT_OF_UNITY_MONT; for (int i = 0; i < 8; i++) w.l[i] = v[i]; }
    for (uint32_t i = k; i < 28; i++) w = fp_sqr(w);
    Fr coef = fp_mul(fp_sub(fp_pow_u64(tau, n), fp_one<FrParams>()), fp_inv(fr_from_u64(n)));
    for (auto& sh : s->shards) {
      if (!sh.count) continue;
      DeviceState& d = ctx->dev[sh.dev_index];
      cudaError_t e = cudaSetDevice(d.device);
      Fr* sc = (Fr*)slot(ctx, d, "srs_scalars", sh.count * sizeof(Fr));
      if (!sc) goto fail;
      if (e == cudaSuccess) e = cudaMalloc(&sh.g, sh.count * sizeof(G1Affine));
      if (e == cudaSuccess) e = cudaMalloc(&sh.g_lagrange, sh.count * sizeof(G1Affine));
      if (e != cudaSuccess) { set_error(ctx, SPB_ERR_CUDA, "spb_srs_setup: %s", cudaGetErrorString(e)); goto fail; }
      unsigned blocks = (unsigned)((sh.count + 127) / 128);
      srs_scalars_kernel<<<blocks, 128, 0, d.stream>>>(0, tau, w, coef, sh.start, sh.count, sc);
      g1_fixed_base_mul_kernel<<<blocks, 128, 0, d.stream>>>(sc, sh.count, sh.g);
      srs_scalars_kernel<<<blocks, 128, 0, d.stream>>>(1, tau, w, coef, sh.start, sh.count, sc);
      g1_fixed_base_mul_kernel<<<blocks, 128, 0, d.stream>>>(sc, sh.count, sh.g_lagrange);
      ctx->n_kernel_launches += 4;
      e = cudaStreamSynchronize(d.stream);
      if (e == cudaSuccess) e = cudaGetLastError();
      if (e != cudaSuccess) { set_error(ctx, SPB_ERR_CUDA, "spb_srs_setup: %s", cudaGetErrorString(e)); goto fail; }
    }
    *out = s;
    return 0;
  }
fail:
  spb_srs_free(ctx, s);
  return SPB_ERR_CUDA;
}

int spb_srs_download(spb_ctx* ctx, const spb_srs* srs, int basis, size_t start, size_t count, spb_g1_affine* out) {
  if (!ctx || !srs || !out || start + count > srs->n) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (auto& sh : srs->shards) {
    size_t lo = start > sh.start ? start : sh.start;
    size_t hi = (start + count) < (sh.start + sh.count) ? (start + count) : (sh.start + sh.count);
    if (lo >= hi) continue;
    const G1Affine* src = basis == SPB_BASIS_G ? sh.g : sh.g_lagrange;
    if (!src) return set_error(ctx, SPB_ERR_STATE, "spb_srs_download: basis %d not resident", basis);
    DeviceState& d = ctx->dev[sh.dev_index];
    SPB_CUDA(ctx, cudaSetDevice(d.device));
    SPB_CUDA(ctx, cudaMemcpy((G1Affine*)out + (lo - start), src + (lo - sh.start), (hi - lo) * sizeof(G1Affine), cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ---- MSM ---------------------------------------------------------------------------------------------------
int spb_msm_raw(spb_ctx* ctx, const spb_fr* scalars, const spb_g1_affine* bases, size_t n, spb_g1* out) {
  if (!ctx || !out || (n && (!scalars || !bases))) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::vector<MsmPart> parts;
  size_t D = ctx->dev.size();
  for (size_t i = 0; i < D; i++) {
    size_t lo = n * i / D, cnt = n * (i + 1) / D - lo;
    if (!cnt) continue;
    DeviceState& d = ctx->dev[i];
    SPB_CUDA(ctx, cudaSetDevice(d.device));
    Fr* ds = (Fr*)slot(ctx, d, "msm_scalars", cnt * sizeof(Fr));
    G1Affine* db = (G1Affine*)slot(ctx, d, "msm_bases", cnt * sizeof(G1Affine));
    if (!ds || !db) return SPB_ERR_OOM;
    SPB_CUDA(ctx, cudaMemcpyAsync(ds, (const Fr*)scalars + lo, cnt * sizeof(Fr), cudaMemcpyHostToDevice, d.stream));
    SPB_CUDA(ctx, cudaMemcpyAsync(db, (const G1Affine*)bases + lo, cnt * sizeof(G1Affine), cudaMemcpyHostToDevice, d.stream));
    parts.push_back(MsmPart{(int)i, ds, db, cnt});
  }
  return msm_run_parts(ctx, parts, out);
}

static int msm_srs_common(spb_ctx* ctx, const spb_srs* srs, int basis, const spb_fr* scalars, bool scalars_on_device, size_t n, spb_g1* out) {
  if (!ctx || !srs || !out || (n && !scalars)) return SPB_ERR_ARG;
  if (n > srs->n) return set_error(ctx, SPB_ERR_ARG, "spb_msm: %zu scalars but the SRS has %zu points", n, srs->n);
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::vector<MsmPart> parts;
  for (auto& sh : srs->shards) {
    if (sh.start >= n) continue;
    size_t cnt = (n - sh.start) < sh.count ? (n - sh.start) : sh.count;
    const G1Affine* b = basis == SPB_BASIS_G ? sh.g : sh.g_lagrange;
    if (!b) return set_error(ctx, SPB_ERR_STATE, "spb_msm: basis %d not resident", basis);
    DeviceState& d = ctx->dev[sh.dev_index];
    SPB_CUDA(ctx, cudaSetDevice(d.device));
    const Fr* ds;
    if (scalars_on_device && sh.dev_index == 0) {
      ds = (const Fr*)scalars + sh.start;
    } else {
      Fr* buf = (Fr*)slot(ctx, d, "msm_scalars", cnt * sizeof(Fr));
      if (!buf) return SPB_ERR_OOM;
      // device-resident scalars live on device 0: peer copy for the other shards
      if (scalars_on_device) SPB_CUDA(ctx, cudaMemcpyPeerAsync(buf, d.device, (const Fr*)scalars + sh.start, ctx->dev[0].device, cnt * sizeof(Fr), d.stream));
      else SPB_CUDA(ctx, cudaMemcpyAsync(buf, (const Fr*)scalars + sh.start, cnt * sizeof(Fr), cudaMemcpyHostToDevice, d.stream));
      ds = buf;
    }
    parts.push_back(MsmPart{sh.dev_index, ds, b, cnt});
  }
  return msm_run_parts(ctx, parts, out);
}

int spb_msm(spb_ctx* ctx, const spb_srs* srs, int basis, const spb_fr* scalars, size_t n, spb_g1* out) {
  return msm_srs_common(ctx, srs, basis, scalars, false, n, out);
}
int spb_msm_dev(spb_ctx* ctx, const spb_srs* srs, int basis, const spb_fr* d_scalars, size_t n, spb_g1* out) {
  return msm_srs_common(ctx, srs, basis, d_scalars, true, n, out);
}

int spb_g1_fixed_base_mul(spb_ctx* ctx, const spb_fr* scalars, size_t n, spb_g1_affine* out) {
  if (!ctx || !scalars || !out) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  Fr* ds = (Fr*)slot(ctx, d, "srs_scalars", n * sizeof(Fr));
  G1Affine* dp = (G1Affine*)slot(ctx, d, "fbm_out", n * sizeof(G1Affine));
  if (!ds || !dp) return SPB_ERR_OOM;
  SPB_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, d.stream));
  g1_fixed_base_mul_kernel<<<(unsigned)((n + 127) / 128), 128, 0, d.stream>>>(ds, n, dp);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaMemcpyAsync(out, dp, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, d.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}

}  // extern "C"
