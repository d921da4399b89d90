// Quotient numerator on the extended coset: the three row loops of halo2's Evaluator::evaluate_h
// ([UPSTREAM] halo2_proofs/src/plonk/evaluation.rs; SURVEY.md 8a row a6, stage 8 of create_proof):
//   * spb_graph_evaluate_dev      -- GraphEvaluator::evaluate for every extended row (custom gates, and the
//                                    compressed-expression product each lookup needs),
//   * spb_permutation_constraints_dev -- the permutation argument terms folded with powers of y,
//   * spb_lookup_constraints_dev  -- the five lookup-argument terms of one lookup.
// Each is one streaming pass over the extended polynomials it reads: algorithmic bytes = 32 B x (#input columns + 1
// read + 1 write of `values`) per extended row. The gate graph arrives in the flat encoding documented at
// include/spectre_b200.h (spb_graph) -- what the Rust shim produces from GraphEvaluator's `calculations`.
// Intermediates live in a device scratch laid out [intermediate][thread slot] (coalesced), rows are grid-strided.
// The per-row bodies are in quotient.cuh (host+device: tests/hostemu runs them on the CPU).
// Parity: bit-exact against the CPU restatement on synthetic constraint systems (tests/test_gpu_quotient.py);
// not pinned by any reference-owned vector (none exists for this row).
#include "common.cuh"
#include "quotient.cuh"
#include <stdlib.h>
#include <string.h>

using namespace spb;

__global__ void __launch_bounds__(256) graph_evaluate_kernel(GraphArgs a) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, nslots = gridDim.x * blockDim.x;
  for (uint64_t row = a.row_lo + slot; row < a.row_hi; row += nslots) graph_evaluate_row(a, row, slot, nslots);
}

// one extended_omega power per block (thread 0: 2 log2(idx) products) times a 256-entry table entry per thread, instead of
// a full exponentiation per row
__global__ void __launch_bounds__(256) permutation_constraints_kernel(PermArgs a) {
  __shared__ Fr base;
  const uint64_t block_row = a.row_lo + blockIdx.x * (uint64_t)blockDim.x;   // row_lo is a multiple of 256
  if (threadIdx.x == 0) base = fp_pow_u64(a.extended_omega, block_row);
  __syncthreads();
  const uint64_t idx = block_row + threadIdx.x;
  if (idx < a.row_hi) permutation_constraints_row(a, idx, fp_mul(base, ntt_ldg(a.omega_pow + threadIdx.x)));
}

__global__ void __launch_bounds__(256) lookup_constraints_kernel(LookupArgs a) {
  uint64_t idx = a.row_lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx < a.row_hi) lookup_constraints_row(a, idx);
}

namespace {

// Row-range shards of one pass over `size` extended rows (SURVEY.md 8e "evaluate_h / pointwise": embarrassingly parallel by
// extended-row range). The polynomials stay where the caller put them -- on the first device of the context; the other
// devices run the same kernel on their row range and read the inputs (rotations included: any row of any polynomial) and
// write their slice of `values` straight through NVLink peer access, so no halo is exchanged and no staging copy exists.
struct RowShard { int dev_index; uint64_t lo, hi; };
std::vector<RowShard> row_shards(spb_ctx* ctx, uint64_t size) {
  std::vector<RowShard> v;
  const size_t D = ctx->dev.size();
  uint64_t min_rows = (uint64_t)1 << 16;   // below this a pass is launch-bound: one device (tests lower it: SPB_SHARD_MIN_ROWS)
  if (const char* e = getenv("SPB_SHARD_MIN_ROWS")) { long long v = atoll(e); if (v >= 256) min_rows = (uint64_t)v; }
  if (D > 1 && ctx->peer_access && size >= min_rows) {
    const uint64_t per = ((size + D - 1) / D + 255) / 256 * 256;
    for (size_t i = 0; i < D; i++) {
      uint64_t lo = per * i, hi = lo + per < size ? lo + per : size;
      if (lo < hi) v.push_back(RowShard{(int)i, lo, hi});
    }
  } else {
    v.push_back(RowShard{0, 0, size});
  }
  return v;
}
// the shard's device: current device set, its stream ordered after everything queued on the first device's stream so far
int shard_begin(spb_ctx* ctx, const RowShard& sh) {
  DeviceState& d0 = ctx->dev[0];
  DeviceState& d = ctx->dev[sh.dev_index];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  // d0.dep_ev was recorded at entry (SPB_ENTER0), i.e. BEFORE the first device's own shard was enqueued: the other devices
  // wait for the caller's inputs only, not for the first device's share of this pass
  if (sh.dev_index != 0) SPB_CUDA(ctx, cudaStreamWaitEvent(d.stream, d0.dep_ev, 0));
  return 0;
}
// wait for every shard; last_kernel_ms = device time of the pass on the first device's clock
int shards_finish(spb_ctx* ctx, const std::vector<RowShard>& shards) {
  DeviceState& d0 = ctx->dev[0];
  for (auto& sh : shards) {
    if (sh.dev_index == 0) continue;
    DeviceState& d = ctx->dev[sh.dev_index];
    SPB_CUDA(ctx, cudaSetDevice(d.device));
    SPB_CUDA(ctx, cudaEventRecord(d.dep_ev, d.stream));
    SPB_CUDA(ctx, cudaStreamWaitEvent(d0.stream, d.dep_ev, 0));
  }
  SPB_CUDA(ctx, cudaSetDevice(d0.device));
  SPB_CUDA(ctx, cudaEventRecord(d0.ev1, d0.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d0.stream));
  SPB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, d0.ev0, d0.ev1));
  for (auto& sh : shards) if (sh.dev_index != 0) { SPB_CUDA(ctx, cudaSetDevice(ctx->dev[sh.dev_index].device)); SPB_CUDA(ctx, cudaStreamSynchronize(ctx->dev[sh.dev_index].stream)); }
  SPB_CUDA(ctx, cudaSetDevice(d0.device));
  return 0;
}

// copy a host array of device pointers to the device (slot `name`)
const Fr* const* upload_ptrs(spb_ctx* ctx, DeviceState& d, const char* name, const spb_fr* const* p, uint32_t n) {
  void* dst = slot(ctx, d, name, (n ? n : 1) * sizeof(void*));
  if (!dst) return nullptr;
  if (n && cudaMemcpyAsync(dst, p, n * sizeof(void*), cudaMemcpyHostToDevice, d.stream) != cudaSuccess) return nullptr;
  return (const Fr* const*)dst;
}

}  // namespace

extern "C" {

#define SPB_ENTER0(ctx)                         \
  std::lock_guard<std::mutex> lk((ctx)->mu);    \
  DeviceState& d0 = (ctx)->dev[0];              \
  SPB_CUDA(ctx, cudaSetDevice(d0.device));      \
  SPB_CUDA(ctx, cudaEventRecord(d0.ev0, d0.stream));  \
  if ((ctx)->dev.size() > 1) SPB_CUDA(ctx, cudaEventRecord(d0.dep_ev, d0.stream));

int spb_graph_evaluate_dev(spb_ctx* ctx, const spb_graph* g, const spb_fr* const* d_fixed, uint32_t n_fixed, const spb_fr* const* d_advice, uint32_t n_advice,
                           const spb_fr* const* d_instance, uint32_t n_instance, const spb_fr* challenges, uint32_t n_challenges, const spb_fr* beta,
                           const spb_fr* gamma, const spb_fr* theta, const spb_fr* y, spb_fr* d_values, uint64_t size, int32_t rot_scale) {
  if (!ctx || !g || !d_values || !beta || !gamma || !theta || !y || (g->program_words && !g->program)) return SPB_ERR_ARG;
  if (g->num_intermediates > 0xffff || g->num_constants > 0x10000 || g->num_rotations > 0xffff) return set_error(ctx, SPB_ERR_ARG, "graph: index fields are 16 bits");
  SPB_ENTER0(ctx);
  std::vector<Fr> sc(4 + n_challenges);
  memcpy(&sc[0], beta, 32); memcpy(&sc[1], gamma, 32); memcpy(&sc[2], theta, 32); memcpy(&sc[3], y, 32);
  if (n_challenges) memcpy(&sc[4], challenges, (size_t)n_challenges * 32);
  const std::vector<RowShard> shards = row_shards(ctx, size);
  for (auto& sh : shards) {
    DeviceState& d = ctx->dev[sh.dev_index];
    SPB_TRY(shard_begin(ctx, sh));
    const uint32_t threads = 256, blocks = (uint32_t)d.sm_count * 2;   // grid-stride: 2 x 256 threads per SM
    const uint64_t nslots = (uint64_t)threads * blocks;
    GraphArgs a; memset(&a, 0, sizeof a);
    uint32_t* dprog = (uint32_t*)slot(ctx, d, "q_prog", (g->program_words ? g->program_words : 1) * 4);
    Fr* dconst = (Fr*)slot(ctx, d, "q_const", (g->num_constants ? g->num_constants : 1) * sizeof(Fr));
    int32_t* drot = (int32_t*)slot(ctx, d, "q_rot", (g->num_rotations ? g->num_rotations : 1) * 4);
    Fr* dscal = (Fr*)slot(ctx, d, "q_scalars", (4 + (size_t)n_challenges) * sizeof(Fr));
    Fr* scratch = (Fr*)slot(ctx, d, "q_scratch", (g->num_intermediates ? g->num_intermediates : 1) * nslots * sizeof(Fr));
    if (!dprog || !dconst || !drot || !dscal || !scratch) return SPB_ERR_OOM;
    SPB_CUDA(ctx, cudaMemcpyAsync(dprog, g->program, g->program_words * 4, cudaMemcpyHostToDevice, d.stream));
    if (g->num_constants) SPB_CUDA(ctx, cudaMemcpyAsync(dconst, g->constants, (size_t)g->num_constants * 32, cudaMemcpyHostToDevice, d.stream));
    if (g->num_rotations) SPB_CUDA(ctx, cudaMemcpyAsync(drot, g->rotations, (size_t)g->num_rotations * 4, cudaMemcpyHostToDevice, d.stream));
    SPB_CUDA(ctx, cudaMemcpyAsync(dscal, sc.data(), sc.size() * 32, cudaMemcpyHostToDevice, d.stream));
    a.fixed = upload_ptrs(ctx, d, "q_fixed", d_fixed, n_fixed);
    a.advice = upload_ptrs(ctx, d, "q_advice", d_advice, n_advice);
    a.instance = upload_ptrs(ctx, d, "q_instance", d_instance, n_instance);
    if (!a.fixed || !a.advice || !a.instance) return set_error(ctx, SPB_ERR_CUDA, "graph: pointer table upload failed");
    a.prog = dprog; a.ncalc = g->num_calculations; a.constants = dconst; a.rotations = drot; a.scalars = dscal;
    a.values = (Fr*)d_values; a.scratch = scratch; a.size = size; a.rot_scale = rot_scale; a.row_lo = sh.lo; a.row_hi = sh.hi;
    graph_evaluate_kernel<<<blocks, threads, 0, d.stream>>>(a);
    SPB_CUDA(ctx, cudaGetLastError());
    ctx->n_kernel_launches++;
  }
  return shards_finish(ctx, shards);  // synchronises: `sc` and the caller's arrays outlive the copies
}

int spb_permutation_constraints_dev(spb_ctx* ctx, spb_fr* d_values, uint64_t size, int32_t rot_scale, int32_t last_rotation, uint32_t n_sets, uint32_t chunk_len,
                                    const spb_fr* const* d_z, uint32_t n_cols, const spb_fr* const* d_col_values, const spb_fr* const* d_sigma,
                                    const spb_fr* d_l0, const spb_fr* d_l_last, const spb_fr* d_l_active, const spb_fr* beta, const spb_fr* gamma,
                                    const spb_fr* y, const spb_fr* extended_omega) {
  if (!ctx || !d_values || !beta || !gamma || !y || !extended_omega || !d_l0 || !d_l_last || !d_l_active) return SPB_ERR_ARG;
  if (!n_sets) return 0;
  if (!d_z || !chunk_len || (n_cols && (!d_col_values || !d_sigma))) return SPB_ERR_ARG;
  SPB_ENTER0(ctx);
  PermArgs a; memset(&a, 0, sizeof a);
  a.values = (Fr*)d_values; a.size = size; a.rot_scale = rot_scale; a.last_rotation = last_rotation;
  a.n_sets = n_sets; a.chunk_len = chunk_len; a.n_cols = n_cols;
  a.l0 = (const Fr*)d_l0; a.l_last = (const Fr*)d_l_last; a.l_active = (const Fr*)d_l_active;
  memcpy(&a.beta, beta, 32); memcpy(&a.gamma, gamma, 32); memcpy(&a.y, y, 32); memcpy(&a.extended_omega, extended_omega, 32);
  Fr zeta; { constexpr uint32_t v[8] = SPB_FR_ZETA_MONT; for (int i = 0; i < 8; i++) zeta.l[i] = v[i]; }
  { constexpr uint32_t v[8] = SPB_FR_DELTA_MONT; for (int i = 0; i < 8; i++) a.delta.l[i] = v[i]; }
  a.delta_start = fp_mul(a.beta, zeta);
  std::vector<Fr> pw(256);
  pw[0] = fp_one<FrParams>();
  for (int j = 1; j < 256; j++) pw[j] = fp_mul(pw[j - 1], a.extended_omega);
  const std::vector<RowShard> shards = row_shards(ctx, size);
  for (auto& sh : shards) {
    DeviceState& d = ctx->dev[sh.dev_index];
    SPB_TRY(shard_begin(ctx, sh));
    a.z = upload_ptrs(ctx, d, "q_z", d_z, n_sets);
    a.col_values = upload_ptrs(ctx, d, "q_cols", d_col_values, n_cols);
    a.sigma = upload_ptrs(ctx, d, "q_sigma", d_sigma, n_cols);
    Fr* dpw = (Fr*)slot(ctx, d, "q_omega_pow", 256 * sizeof(Fr));
    if (!a.z || !a.col_values || !a.sigma || !dpw) return set_error(ctx, SPB_ERR_CUDA, "permutation: table upload failed");
    SPB_CUDA(ctx, cudaMemcpyAsync(dpw, pw.data(), 256 * sizeof(Fr), cudaMemcpyHostToDevice, d.stream));
    a.omega_pow = dpw; a.row_lo = sh.lo; a.row_hi = sh.hi;
    permutation_constraints_kernel<<<(unsigned)((sh.hi - sh.lo + 255) / 256), 256, 0, d.stream>>>(a);
    SPB_CUDA(ctx, cudaGetLastError());
    ctx->n_kernel_launches++;
  }
  return shards_finish(ctx, shards);
}

int spb_lookup_constraints_dev(spb_ctx* ctx, spb_fr* d_values, uint64_t size, int32_t rot_scale, const spb_fr* d_product, const spb_fr* d_permuted_input,
                               const spb_fr* d_permuted_table, const spb_fr* d_table_value, const spb_fr* d_l0, const spb_fr* d_l_last,
                               const spb_fr* d_l_active, const spb_fr* beta, const spb_fr* gamma, const spb_fr* y) {
  if (!ctx || !d_values || !d_product || !d_permuted_input || !d_permuted_table || !d_table_value || !d_l0 || !d_l_last || !d_l_active || !beta || !gamma || !y)
    return SPB_ERR_ARG;
  SPB_ENTER0(ctx);
  LookupArgs a;
  a.values = (Fr*)d_values; a.size = size; a.rot_scale = rot_scale;
  a.product = (const Fr*)d_product; a.permuted_input = (const Fr*)d_permuted_input; a.permuted_table = (const Fr*)d_permuted_table;
  a.table_value = (const Fr*)d_table_value; a.l0 = (const Fr*)d_l0; a.l_last = (const Fr*)d_l_last; a.l_active = (const Fr*)d_l_active;
  memcpy(&a.beta, beta, 32); memcpy(&a.gamma, gamma, 32); memcpy(&a.y, y, 32);
  const std::vector<RowShard> shards = row_shards(ctx, size);
  for (auto& sh : shards) {
    DeviceState& d = ctx->dev[sh.dev_index];
    SPB_TRY(shard_begin(ctx, sh));
    a.row_lo = sh.lo; a.row_hi = sh.hi;
    lookup_constraints_kernel<<<(unsigned)((sh.hi - sh.lo + 255) / 256), 256, 0, d.stream>>>(a);
    SPB_CUDA(ctx, cudaGetLastError());
    ctx->n_kernel_launches++;
  }
  return shards_finish(ctx, shards);
}

}  // extern "C"
