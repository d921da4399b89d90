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
// Parity: bit-exact against the CPU restatement on synthetic constraint systems (tests/test_gpu_quotient.py);
// not pinned by any reference-owned vector (none exists for this row).
#include "common.cuh"
#include "ntt.cuh"
#include <string.h>

using namespace spb;

struct GraphArgs {
  const uint32_t* prog; uint32_t ncalc;
  const Fr* constants; const int32_t* rotations;
  const Fr* const* fixed; const Fr* const* advice; const Fr* const* instance;
  const Fr* scalars;   // [beta, gamma, theta, y, challenges...]
  Fr* values; Fr* scratch; uint64_t size; int32_t rot_scale;
};

__device__ __forceinline__ uint64_t rotation_idx(uint64_t idx, int32_t rot, int32_t rot_scale, uint64_t size) {
  long long v = ((long long)idx + (long long)rot * rot_scale) % (long long)size;
  if (v < 0) v += (long long)size;
  return (uint64_t)v;
}

__device__ __forceinline__ Fr graph_src(const GraphArgs& a, const uint32_t* w, uint64_t row, uint32_t slot, uint32_t nslots, const Fr& previous) {
  const uint32_t kind = w[0], idx = w[1] & 0xffffu, rot = w[1] >> 16;
  switch (kind) {
    case 0: return ntt_ldg(a.constants + idx);
    case 1: return a.scratch[(uint64_t)idx * nslots + slot];
    case 2: return ntt_ldg(a.fixed[idx] + rotation_idx(row, a.rotations[rot], a.rot_scale, a.size));
    case 3: return ntt_ldg(a.advice[idx] + rotation_idx(row, a.rotations[rot], a.rot_scale, a.size));
    case 4: return ntt_ldg(a.instance[idx] + rotation_idx(row, a.rotations[rot], a.rot_scale, a.size));
    case 5: return ntt_ldg(a.scalars + 4 + idx);
    case 6: return ntt_ldg(a.scalars + 0);
    case 7: return ntt_ldg(a.scalars + 1);
    case 8: return ntt_ldg(a.scalars + 2);
    case 9: return ntt_ldg(a.scalars + 3);
    default: return previous;
  }
}

__global__ void __launch_bounds__(256) graph_evaluate_kernel(GraphArgs a) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, nslots = gridDim.x * blockDim.x;
  for (uint64_t row = slot; row < a.size; row += nslots) {
    const Fr previous = a.values[row];
    const uint32_t* w = a.prog;
    Fr last = fp_zero<FrParams>();
    for (uint32_t c = 0; c < a.ncalc; c++) {
      const uint32_t op = w[0] & 0xffu, nparts = w[0] >> 8, target = w[1];
      Fr r;
      if (op <= 2) {
        Fr x = graph_src(a, w + 2, row, slot, nslots, previous), y = graph_src(a, w + 4, row, slot, nslots, previous);
        r = op == 0 ? fp_add(x, y) : op == 1 ? fp_sub(x, y) : fp_mul(x, y);
        w += 6;
      } else if (op == 6) {
        Fr acc = graph_src(a, w + 2, row, slot, nslots, previous), factor = graph_src(a, w + 4, row, slot, nslots, previous);
        for (uint32_t p = 0; p < nparts; p++) acc = fp_add(fp_mul(acc, factor), graph_src(a, w + 6 + 2 * p, row, slot, nslots, previous));
        r = acc; w += 6 + 2 * nparts;
      } else {
        Fr x = graph_src(a, w + 2, row, slot, nslots, previous);
        r = op == 3 ? fp_sqr(x) : op == 4 ? fp_dbl(x) : op == 5 ? fp_neg(x) : x;
        w += 4;
      }
      a.scratch[(uint64_t)target * nslots + slot] = r;
      last = r;
    }
    a.values[row] = last;
  }
}

struct PermArgs {
  Fr* values; uint64_t size; int32_t rot_scale, last_rotation; uint32_t n_sets, chunk_len, n_cols;
  const Fr* const* z; const Fr* const* col_values; const Fr* const* sigma;
  const Fr* l0; const Fr* l_last; const Fr* l_active;
  Fr beta, gamma, y, delta_start, delta, extended_omega;
};

__global__ void __launch_bounds__(256) permutation_constraints_kernel(PermArgs a) {
  uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx >= a.size) return;
  const uint64_t r_next = rotation_idx(idx, 1, a.rot_scale, a.size), r_last = rotation_idx(idx, a.last_rotation, a.rot_scale, a.size);
  const Fr one = fp_one<FrParams>();
  Fr v = ntt_ld_stream(a.values + idx);
  const Fr l0 = ntt_ldg(a.l0 + idx), l_last = ntt_ldg(a.l_last + idx), l_active = ntt_ldg(a.l_active + idx);
  v = fp_add(fp_mul(v, a.y), fp_mul(fp_sub(one, ntt_ldg(a.z[0] + idx)), l0));
  { Fr zl = ntt_ldg(a.z[a.n_sets - 1] + idx); v = fp_add(fp_mul(v, a.y), fp_mul(fp_sub(fp_sqr(zl), zl), l_last)); }
  for (uint32_t s = 1; s < a.n_sets; s++)
    v = fp_add(fp_mul(v, a.y), fp_mul(fp_sub(ntt_ldg(a.z[s] + idx), ntt_ldg(a.z[s - 1] + r_last)), l0));
  Fr current_delta = fp_mul(a.delta_start, fp_pow_u64(a.extended_omega, idx));
  for (uint32_t s = 0; s < a.n_sets; s++) {
    const uint32_t lo = s * a.chunk_len, hi = lo + a.chunk_len < a.n_cols ? lo + a.chunk_len : a.n_cols;
    Fr left = ntt_ldg(a.z[s] + r_next), right = ntt_ldg(a.z[s] + idx);
    for (uint32_t c = lo; c < hi; c++) {
      Fr val = ntt_ldg(a.col_values[c] + idx);
      left = fp_mul(left, fp_add(fp_add(val, fp_mul(a.beta, ntt_ldg(a.sigma[c] + idx))), a.gamma));
      right = fp_mul(right, fp_add(fp_add(val, current_delta), a.gamma));
      current_delta = fp_mul(current_delta, a.delta);
    }
    v = fp_add(fp_mul(v, a.y), fp_mul(fp_sub(left, right), l_active));
  }
  ntt_stg(a.values + idx, v);
}

__global__ void __launch_bounds__(256) lookup_constraints_kernel(Fr* values, uint64_t size, int32_t rot_scale, const Fr* product, const Fr* permuted_input,
                                                                 const Fr* permuted_table, const Fr* table_value, const Fr* l0p, const Fr* l_lastp,
                                                                 const Fr* l_activep, Fr beta, Fr gamma, Fr y) {
  uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx >= size) return;
  const uint64_t r_next = rotation_idx(idx, 1, rot_scale, size), r_prev = rotation_idx(idx, -1, rot_scale, size);
  const Fr one = fp_one<FrParams>();
  const Fr l0 = ntt_ldg(l0p + idx), l_last = ntt_ldg(l_lastp + idx), l_active = ntt_ldg(l_activep + idx);
  const Fr a_in = ntt_ldg(permuted_input + idx), s_tb = ntt_ldg(permuted_table + idx), zp = ntt_ldg(product + idx);
  const Fr a_minus_s = fp_sub(a_in, s_tb);
  Fr v = ntt_ld_stream(values + idx);
  v = fp_add(fp_mul(v, y), fp_mul(fp_sub(one, zp), l0));
  v = fp_add(fp_mul(v, y), fp_mul(fp_sub(fp_sqr(zp), zp), l_last));
  Fr lhs = fp_mul(fp_mul(ntt_ldg(product + r_next), fp_add(a_in, beta)), fp_add(s_tb, gamma));
  v = fp_add(fp_mul(v, y), fp_mul(fp_sub(lhs, fp_mul(zp, ntt_ldg(table_value + idx))), l_active));
  v = fp_add(fp_mul(v, y), fp_mul(a_minus_s, l0));
  v = fp_add(fp_mul(v, y), fp_mul(fp_mul(a_minus_s, fp_sub(a_in, ntt_ldg(permuted_input + r_prev))), l_active));
  ntt_stg(values + idx, v);
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
  Fr b, g, yy; memcpy(&b, beta, 32); memcpy(&g, gamma, 32); memcpy(&yy, y, 32);
  lookup_constraints_kernel<<<(unsigned)((size + 255) / 256), 256, 0, d.stream>>>((Fr*)d_values, size, rot_scale, (const Fr*)d_product, (const Fr*)d_permuted_input,
                                                                                 (const Fr*)d_permuted_table, (const Fr*)d_table_value, (const Fr*)d_l0,
                                                                                 (const Fr*)d_l_last, (const Fr*)d_l_active, b, g, yy);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches++;
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}

}  // extern "C"
