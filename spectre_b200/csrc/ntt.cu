// Host side of the NTT: power-table cache, digit plan, pass launches. Kernels are in ntt.cuh.
#define SPB_NTT_KERNELS 1
#include "common.cuh"
#include "ntt.cuh"
#include <stdlib.h>
#include <string.h>

namespace spb {

static const uint32_t kTileElemsLog = 12;     // 4096 elements (~135 KB of limb planes) per tile
// largest sub-NTT held in one shared-memory tile (11: two passes up to 2^22; the tile is then 2048 x 2 columns)
static uint32_t max_digit_bits() {
  static uint32_t v = 0;
  if (!v) { const char* e = getenv("SPB_NTT_MAX_DIGIT"); v = e ? (uint32_t)atoi(e) : 11; if (v < 4) v = 4; if (v > 12) v = 12; }
  return v;
}
// full omega^i tables are kept while their total stays under this many bytes per device (else two-level tables)
static size_t full_table_budget() {
  static size_t v = 0;
  if (!v) { const char* e = getenv("SPB_NTT_FULL_TABLE_MB"); v = (e ? (size_t)atoll(e) : 6144) << 20; if (!v) v = 1; }
  return v;
}

struct NttPlan {
  uint32_t npass;
  uint32_t s[3];
};

static NttPlan make_plan(uint32_t k) {
  NttPlan p; memset(&p, 0, sizeof p);
  const uint32_t kMaxDigitBits = max_digit_bits();
  if (k <= kMaxDigitBits) { p.npass = 1; p.s[0] = k; return p; }
  p.npass = (k + kMaxDigitBits - 1) / kMaxDigitBits;
  uint32_t rem = k;
  for (uint32_t i = 0; i < p.npass; i++) {
    uint32_t left = p.npass - i;
    p.s[i] = (rem + left - 1) / left;  // larger digits first
    rem -= p.s[i];
  }
  return p;
}

static int get_tables(spb_ctx* ctx, DeviceState& d, uint32_t k, const Fr& omega, uint32_t h, NttTables** out) {
  for (auto& t : d.ntt_tables)
    if (t.k == k && t.h == h && fp_eq(t.omega, omega)) { *out = &t; return 0; }
  NttTables t; t.omega = omega; t.k = k; t.h = h;
  size_t nlo = (size_t)1 << h, nhi = (size_t)1 << (k - h);
  SPB_CUDA(ctx, cudaMalloc(&t.tw_lo, nlo * sizeof(Fr)));
  SPB_CUDA(ctx, cudaMalloc(&t.tw_hi, nhi * sizeof(Fr)));
  fr_pow_table_kernel<<<(unsigned)((nlo + 127) / 128), 128, 0, d.stream>>>(t.tw_lo, omega, nlo, 0);
  fr_pow_table_kernel<<<(unsigned)((nhi + 127) / 128), 128, 0, d.stream>>>(t.tw_hi, omega, nhi, h);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches += 2;
  {
    size_t used = 0;
    for (auto& o : d.ntt_tables) if (o.tw_full) used += ((size_t)1 << o.k) * sizeof(Fr);
    size_t need = ((size_t)1 << k) * sizeof(Fr);
    if (k >= 12 && used + need <= full_table_budget() && cudaMalloc(&t.tw_full, need) == cudaSuccess) {
      fr_pow_table_kernel<<<(unsigned)((((size_t)1 << k) + 127) / 128), 128, 0, d.stream>>>(t.tw_full, omega, (uint64_t)1 << k, 0);
      ctx->n_kernel_launches++;
    } else {
      t.tw_full = nullptr;
      cudaGetLastError();
    }
  }
  // a long-lived prover touches a handful of (k, omega) pairs; cap the cache anyway
  if (d.ntt_tables.size() >= 32) {
    cudaStreamSynchronize(d.stream);
    cudaFree(d.ntt_tables.front().tw_lo); cudaFree(d.ntt_tables.front().tw_hi); if (d.ntt_tables.front().tw_full) cudaFree(d.ntt_tables.front().tw_full);
    d.ntt_tables.erase(d.ntt_tables.begin());
  }
  d.ntt_tables.push_back(t);
  *out = &d.ntt_tables.back();
  return 0;
}

int ntt_device(spb_ctx* ctx, DeviceState& d, const Fr* d_src, Fr* d_dst, uint32_t k, const Fr& omega, const NttOpts& opts) {
  if (k > 28) return set_error(ctx, SPB_ERR_ARG, "ntt: log_n = %u exceeds the two-adicity (28) of Fr", k);
  const uint64_t n = 1ull << k;
  NttPlan plan = make_plan(k);
  uint32_t smax = plan.s[0];
  uint32_t h = k - smax;
  NttTables* tb = nullptr;
  SPB_TRY(get_tables(ctx, d, k, omega, h, &tb));

  Fr* tmp = nullptr;
  if (plan.npass > 1) {
    tmp = (Fr*)slot(ctx, d, "ntt_tmp", n * sizeof(Fr));
    if (!tmp) return SPB_ERR_OOM;
  }
  static bool attr_set = false;
  if (!attr_set) {
    SPB_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }

  uint32_t a = 0;
  for (uint32_t pi = 0; pi < plan.npass; pi++) {
    NttPassParams p; memset(&p, 0, sizeof p);
    p.k = k; p.h = h; p.tw_lo = tb->tw_lo; p.tw_hi = tb->tw_hi; p.tw_full = tb->tw_full;
    p.s = plan.s[pi]; p.a = a; p.b = k - a - p.s; p.s1 = plan.s[0];
    p.first = (pi == 0); p.last = (pi == plan.npass - 1);
    p.b_next = p.last ? 0 : p.b - plan.s[pi + 1];
    p.n_in = opts.n_in ? opts.n_in : n;
    p.n_out = opts.n_out ? opts.n_out : n;
    if (p.first && opts.pre3) { p.use_pre = 1; for (int i = 0; i < 3; i++) p.pre[i] = opts.pre3[i]; }
    if (p.last && opts.post3) { p.use_post = 1; for (int i = 0; i < 3; i++) p.post[i] = opts.post3[i]; }
    // columns per tile: as many as fit, bounded by what the direction offers
    uint32_t avail = p.last ? (p.a ? p.s1 : 0) : p.b;
    uint32_t logc = kTileElemsLog > p.s ? kTileElemsLog - p.s : 0;
    if (logc > avail) logc = avail;
    if (logc > 5) logc = 5;
    p.logc = logc;
    p.src = (pi == 0) ? d_src : tmp;
    p.dst = p.last ? d_dst : tmp;
    uint64_t tiles = n >> (p.s + logc);
    uint32_t S = 1u << p.s, C = 1u << logc;
    uint32_t quads = (S * C) / 4; if (quads < 32) quads = 32;
    uint32_t threads = quads < 512 ? quads : 512;
    size_t smem = (size_t)8 * 4 * ((size_t)ntt_col_stride(S, C) * C + ntt_tw_words(S));
    if (smem > 227 * 1024) return set_error(ctx, SPB_ERR_STATE, "ntt: tile needs %zu B of shared memory", smem);
    ntt_pass_kernel<<<(unsigned)tiles, threads, smem, d.stream>>>(p);
    SPB_CUDA(ctx, cudaGetLastError());
    ctx->n_kernel_launches++;
    a += p.s;
  }
  return 0;
}

}  // namespace spb
