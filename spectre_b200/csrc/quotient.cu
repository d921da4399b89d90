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
#include <string.h>

using namespace spb;

__global__ void __launch_bounds__(256) graph_evaluate_kernel(GraphArgs a) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, nslots = gridDim.x * blockDim.x;
  for (uint64_t row = slot; row < a.size; row += nslots) graph_evaluate_row(a, row, slot, nslots);
}

__global__ void __launch_bounds__(256) permutation_constraints_kernel(PermArgs a) {
  uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx < a.size) permutation_constraints_row(a, idx);
}

__global__ void __launch_bounds__(256) lookup_constraints_kernel(LookupArgs a) {
  uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx < a.size) lookup_constraints_row(a, idx);
}

extern "C" {

#define SPB_ENTER(ctx)                          \
  std::lock_guard<std::mutex> lk((ctx)->mu);    \
  DeviceState& d = (ctx)->dev[0];               \
  SPB_CUDA(ctx, cudaSetDevice(d.device));

// copy a host array of device pointers to the device (slot `name`)
static const Fr* const* upload_ptrs(spb_ctx* ctx, DeviceState& d, const char* name, const spb_fr* const* p, uint32_t n) {
  void* dst = slot(ctx, d, name, (n ? n : 1) * sizeof(void*));
  if (!dst) return nullptr;
  if (n && cudaMemcpyAsync(dst, p, n * sizeof(void*), cudaMemcpyHostToDevice, d.stream) != cudaSuccess) return nullptr;
  return (const Fr* const*)dst;
}

int spb_graph_evaluate_dev(spb_ctx* ctx, const spb_graph* g, const spb_fr* const* d_fixed, uint32_t n_fixed, const spb_fr* const* d_advice, uint32_t n_advice,
                           const spb_fr* const* d_instance, uint32_t n_instance, const spb_fr* challenges, uint32_t n_challenges, const spb_fr* beta,
                           const spb_fr* gamma, const spb_fr* theta, const spb_fr* y, spb_fr* d_values, uint64_t size, int32_t rot_scale) {
  if (!ctx || !g || !d_values || !beta || !gamma || !theta || !y || (g->program_words && !g->program)) return SPB_ERR_ARG;
  if (g->num_intermediates > 0xffff || g->num_constants > 0x10000 || g->num_rotations > 0xffff) return set_error(ctx, SPB_ERR_ARG, "graph: index fields are 16 bits");
  SPB_ENTER(ctx);
  const uint32_t threads = 256, blocks = (uint32_t)d.sm_count * 2;   // grid-stride: 2 x 256 threads per SM
  const uint64_t nslots = (uint64_t)threads * blocks;
  GraphArgs a; memset(&a, 0, sizeof a);
  uint32_t* dprog = (uint32_t*)slot(ctx, d, "q_prog", (g->program_words ? g->program_words : 1) * 4);
  Fr* dconst = (Fr*)slot(ctx, d, "q_const", (g->num_constants ? g->num_constants : 1) * sizeof(Fr));
  int32_t* drot = (int32_t*)slot(ctx, d, "q_rot", (g->num_rotations ? g->num_rotations : 1) * 4);
  Fr* dscal = (Fr*)slot(ctx, d, "q_scalars", (4 + (size_t)n_challenges) * sizeof(Fr));
  Fr* scratch = (Fr*)slot(ctx, d, "q_scratch", (g->num_intermediates ? g->num_intermediates : 1) * nslots * sizeof(Fr));
  if (!dprog || !dconst || !drot || !dscal || !scratch) return SPB_ERR_OOM;
  std::vector<Fr> sc(4 + n_challenges);
  memcpy(&sc[0], beta, 32); memcpy(&sc[1], gamma, 32); memcpy(&sc[2], theta, 32); memcpy(&sc[3], y, 32);
  if (n_challenges) memcpy(&sc[4], challenges, (size_t)n_challenges * 32);
  SPB_CUDA(ctx, cudaMemcpyAsync(dprog, g->program, g->program_words * 4, cudaMemcpyHostToDevice, d.stream));
  if (g->num_constants) SPB_CUDA(ctx, cudaMemcpyAsync(dconst, g->constants, (size_t)g->num_constants * 32, cudaMemcpyHostToDevice, d.stream));
  if (g->num_rotations) SPB_CUDA(ctx, cudaMemcpyAsync(drot, g->rotations, (size_t)g->num_rotations * 4, cudaMemcpyHostToDevice, d.stream));
  SPB_CUDA(ctx, cudaMemcpyAsync(dscal, sc.data(), sc.size() * 32, cudaMemcpyHostToDevice, d.stream));
  a.fixed = upload_ptrs(ctx, d, "q_fixed", d_fixed, n_fixed);
  a.advice = upload_ptrs(ctx, d, "q_advice", d_advice, n_advice);
  a.instance = upload_ptrs(ctx, d, "q_instance", d_instance, n_instance);
  if (!a.fixed || !a.advice || !a.instance) return set_error(ctx, SPB_ERR_CUDA, "graph: pointer table upload failed");
  a.prog = dprog; a.ncalc = g->num_calculations; a.constants = dconst; a.rotations = drot; a.scalars = dscal;
  a.values = (Fr*)d_values; a.scratch = scratch; a.size = size; a.rot_scale = rot_scale;
  graph_evaluate_kernel<<<blocks, threads, 0, d.stream>>>(a);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));  // `sc` and the caller's arrays must outlive the copies
  return 0;
}

int spb_permutation_constraints_dev(spb_ctx* ctx, spb_fr* d_values, uint64_t size, int32_t rot_scale, int32_t last_rotation, uint32_t n_sets, uint32_t chunk_len,
                                    const spb_fr* const* d_z, uint32_t n_cols, const spb_fr* const* d_col_values, const spb_fr* const* d_sigma,
                                    const spb_fr* d_l0, const spb_fr* d_l_last, const spb_fr* d_l_active, const spb_fr* beta, const spb_fr* gamma,
                                    const spb_fr* y, const spb_fr* extended_omega) {
  if (!ctx || !d_values || !beta || !gamma || !y || !extended_omega || !d_l0 || !d_l_last || !d_l_active) return SPB_ERR_ARG;
  if (!n_sets) return 0;
  if (!d_z || !chunk_len || (n_cols && (!d_col_values || !d_sigma))) return SPB_ERR_ARG;
  SPB_ENTER(ctx);
  PermArgs a; memset(&a, 0, sizeof a);
  a.z = upload_ptrs(ctx, d, "q_z", d_z, n_sets);
  a.col_values = upload_ptrs(ctx, d, "q_cols", d_col_values, n_cols);
  a.sigma = upload_ptrs(ctx, d, "q_sigma", d_sigma, n_cols);
  if (!a.z || !a.col_values || !a.sigma) return set_error(ctx, SPB_ERR_CUDA, "permutation: pointer table upload failed");
  a.values = (Fr*)d_values; a.size = size; a.rot_scale = rot_scale; a.last_rotation = last_rotation;
  a.n_sets = n_sets; a.chunk_len = chunk_len; a.n_cols = n_cols;
  a.l0 = (const Fr*)d_l0; a.l_last = (const Fr*)d_l_last; a.l_active = (const Fr*)d_l_active;
  memcpy(&a.beta, beta, 32); memcpy(&a.gamma, gamma, 32); memcpy(&a.y, y, 32); memcpy(&a.extended_omega, extended_omega, 32);
  Fr zeta; { constexpr uint32_t v[8] = SPB_FR_ZETA_MONT; for (int i = 0; i < 8; i++) zeta.l[i] = v[i]; }
  { constexpr uint32_t v[8] = SPB_FR_DELTA_MONT; for (int i = 0; i < 8; i++) a.delta.l[i] = v[i]; }
  a.delta_start = fp_mul(a.beta, zeta);
  permutation_constraints_kernel<<<(unsigned)((size + 255) / 256), 256, 0, d.stream>>>(a);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}

int spb_lookup_constraints_dev(spb_ctx* ctx, spb_fr* d_values, uint64_t size, int32_t rot_scale, const spb_fr* d_product, const spb_fr* d_permuted_input,
                               const spb_fr* d_permuted_table, const spb_fr* d_table_value, const spb_fr* d_l0, const spb_fr* d_l_last,
                               const spb_fr* d_l_active, const spb_fr* beta, const spb_fr* gamma, const spb_fr* y) {
  if (!ctx || !d_values || !d_product || !d_permuted_input || !d_permuted_table || !d_table_value || !d_l0 || !d_l_last || !d_l_active || !beta || !gamma || !y)
    return SPB_ERR_ARG;
  SPB_ENTER(ctx);
  LookupArgs a;
  a.values = (Fr*)d_values; a.size = size; a.rot_scale = rot_scale;
  a.product = (const Fr*)d_product; a.permuted_input = (const Fr*)d_permuted_input; a.permuted_table = (const Fr*)d_permuted_table;
  a.table_value = (const Fr*)d_table_value; a.l0 = (const Fr*)d_l0; a.l_last = (const Fr*)d_l_last; a.l_active = (const Fr*)d_l_active;
  memcpy(&a.beta, beta, 32); memcpy(&a.gamma, gamma, 32); memcpy(&a.y, y, 32);
  lookup_constraints_kernel<<<(unsigned)((size + 255) / 256), 256, 0, d.stream>>>(a);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}

}  // extern "C"
