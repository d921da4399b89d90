// Lookup argument, stage 4 of create_proof: permute_expression_pair on the device.
// Replaces [UPSTREAM] halo2_proofs/src/plonk/lookup/prover.rs::permute_expression_pair (SURVEY.md 8a row a8 / 8f rank 3:
// after MSM and NTT move to the GPU this sort + BTreeMap walk is the serial CPU tail of a proof).
// Same result as the CPU algorithm, element for element:
//   permuted_input  = the input values sorted by canonical integer (halo2curves' Ord for Fr),
//   permuted_table  = at the first row of each distinct input value that value; the remaining table elements
//                     (ascending) handed to the repeated rows from the LAST repeated row backwards.
// Built from data-parallel primitives: Montgomery -> canonical, LSD radix sort over the four 64-bit limbs
// (cub::DeviceRadixSort::SortPairs carrying row indices; four stable passes), adjacent-difference flags, a binary
// search of every distinct input value in the sorted table, two exclusive scans and a scatter.
// Algorithmic bytes: 4 x 32 B per row (two columns in, two out).
#include "common.cuh"
#include "ntt.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <string.h>

using namespace spb;

__global__ void lk_canon_kernel(const Fr* in, Fr* out, uint32_t* idx, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  ntt_stg(out + i, fp_from_mont(ntt_ldg(in + i)));
  idx[i] = (uint32_t)i;
}
__global__ void lk_gather_limb_kernel(const Fr* canon, const uint32_t* idx, uint32_t limb, unsigned long long* keys, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr& v = canon[idx[i]];
  keys[i] = ((unsigned long long)v.l[2 * limb + 1] << 32) | v.l[2 * limb];
}
__global__ void lk_gather_kernel(const Fr* canon, const uint32_t* idx, Fr* sorted, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) ntt_stg(sorted + i, ntt_ldg(canon + idx[i]));
}
__device__ __forceinline__ int lk_cmp(const Fr& a, const Fr& b) {
  for (int i = 7; i >= 0; i--) { if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1; }
  return 0;
}
// first[i] = 1 on the first row of each distinct sorted input value; those rows look their value up in the sorted table
__global__ void lk_match_kernel(const Fr* sin, const Fr* stb, uint64_t n, uint32_t* repeated_flag, uint32_t* used, int* error) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v = ntt_ldg(sin + i);
  bool first = i == 0 || lk_cmp(v, ntt_ldg(sin + i - 1)) != 0;
  repeated_flag[i] = first ? 0u : 1u;
  if (!first) return;
  uint64_t lo = 0, hi = n;
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (lk_cmp(ntt_ldg(stb + mid), v) < 0) lo = mid + 1; else hi = mid; }
  if (lo >= n || lk_cmp(ntt_ldg(stb + lo), v) != 0) { atomicExch(error, 1); return; }
  used[lo] = 1u;   // distinct values hit distinct first occurrences
}
__global__ void lk_leftover_flag_kernel(const uint32_t* used, uint32_t* left_flag, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) left_flag[i] = used[i] ? 0u : 1u;
}
// rep_rows[rank] = row for repeated rows; first rows get permuted_table = value
__global__ void lk_emit_input_kernel(const Fr* sin, const uint32_t* repeated_flag, const uint32_t* rep_rank, uint32_t* rep_rows, Fr* pin, Fr* ptab, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr m = fp_to_mont(ntt_ldg(sin + i));
  ntt_stg(pin + i, m);
  if (repeated_flag[i]) rep_rows[rep_rank[i]] = (uint32_t)i;
  else ntt_stg(ptab + i, m);
}
// leftover element of ascending rank r goes to the repeated row of rank m-1-r
__global__ void lk_emit_leftover_kernel(const Fr* stb, const uint32_t* left_flag, const uint32_t* left_rank, const uint32_t* rep_rows, uint32_t m, Fr* ptab, uint64_t n) {
  uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j >= n || !left_flag[j]) return;
  uint32_t r = left_rank[j];
  if (r >= m) return;
  ntt_stg(ptab + rep_rows[m - 1 - r], fp_to_mont(ntt_ldg(stb + j)));
}

namespace {
inline unsigned nb(uint64_t n) { return (unsigned)((n + 255) / 256); }

// sorted[i] = canonical values of src in ascending order (idx/keys/scratch are n-sized work arrays)
int sort_canonical(spb_ctx* ctx, DeviceState& d, const Fr* src, Fr* canon, Fr* sorted, uint32_t* idx_a, uint32_t* idx_b, unsigned long long* keys_a,
                   unsigned long long* keys_b, void* tmp, size_t tmp_bytes, uint64_t n) {
  lk_canon_kernel<<<nb(n), 256, 0, d.stream>>>(src, canon, idx_a, n);
  for (uint32_t limb = 0; limb < 4; limb++) {
    lk_gather_limb_kernel<<<nb(n), 256, 0, d.stream>>>(canon, idx_a, limb, keys_a, n);
    SPB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 64, d.stream));
    uint32_t* t = idx_a; idx_a = idx_b; idx_b = t;
  }
  lk_gather_kernel<<<nb(n), 256, 0, d.stream>>>(canon, idx_a, sorted, n);   // after four swaps idx_a is the caller's idx_a again
  ctx->n_kernel_launches += 6;
  return 0;
}
}  // namespace

extern "C" {

int spb_permute_expression_pair_dev(spb_ctx* ctx, const spb_fr* d_input, const spb_fr* d_table, size_t usable, spb_fr* d_permuted_input, spb_fr* d_permuted_table) {
  if (!ctx || (usable && (!d_input || !d_table || !d_permuted_input || !d_permuted_table))) return SPB_ERR_ARG;
  if (!usable) return 0;
  if (usable >= 0x7fffffffull) return SPB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceState& d = ctx->dev[0];
  SPB_CUDA(ctx, cudaSetDevice(d.device));
  const uint64_t n = usable;
  Fr* canon = (Fr*)slot(ctx, d, "lk_canon", n * 32);
  Fr* sin = (Fr*)slot(ctx, d, "lk_sin", n * 32);
  Fr* stb = (Fr*)slot(ctx, d, "lk_stb", n * 32);
  uint32_t* idx_a = (uint32_t*)slot(ctx, d, "lk_idx_a", n * 4);
  uint32_t* idx_b = (uint32_t*)slot(ctx, d, "lk_idx_b", n * 4);
  unsigned long long* keys_a = (unsigned long long*)slot(ctx, d, "lk_keys_a", n * 8);
  unsigned long long* keys_b = (unsigned long long*)slot(ctx, d, "lk_keys_b", n * 8);
  uint32_t* flags = (uint32_t*)slot(ctx, d, "lk_flags", (4 * n + 8) * 4);   // repeated_flag | rep_rank | used->left_flag | left_rank
  int* err = (int*)slot(ctx, d, "lk_err", 16);
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 64, d.stream);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags, flags, (int)n + 1, d.stream);
  size_t tmp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  void* tmp = slot(ctx, d, "lk_tmp", tmp_bytes ? tmp_bytes : 16);
  if (!canon || !sin || !stb || !idx_a || !idx_b || !keys_a || !keys_b || !flags || !err || !tmp) return SPB_ERR_OOM;
  uint32_t* repeated_flag = flags, *rep_rank = flags + (n + 1), *used = flags + 2 * (n + 1), *left_rank = flags + 3 * (n + 1);

  SPB_TRY(sort_canonical(ctx, d, (const Fr*)d_input, canon, sin, idx_a, idx_b, keys_a, keys_b, tmp, tmp_bytes, n));
  SPB_TRY(sort_canonical(ctx, d, (const Fr*)d_table, canon, stb, idx_a, idx_b, keys_a, keys_b, tmp, tmp_bytes, n));
  SPB_CUDA(ctx, cudaMemsetAsync(flags, 0, (4 * n + 8) * 4, d.stream));
  SPB_CUDA(ctx, cudaMemsetAsync(err, 0, 4, d.stream));
  lk_match_kernel<<<nb(n), 256, 0, d.stream>>>(sin, stb, n, repeated_flag, used, err);
  SPB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, repeated_flag, rep_rank, (int)n + 1, d.stream));   // rep_rank[n] = #repeated
  lk_leftover_flag_kernel<<<nb(n), 256, 0, d.stream>>>(used, used, n);                                            // in place: used -> left_flag
  SPB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, used, left_rank, (int)n + 1, d.stream));            // left_rank[n] = #leftover
  uint32_t counts[2] = {0, 0}; int herr = 0;
  SPB_CUDA(ctx, cudaMemcpyAsync(&counts[0], rep_rank + n, 4, cudaMemcpyDeviceToHost, d.stream));
  SPB_CUDA(ctx, cudaMemcpyAsync(&counts[1], left_rank + n, 4, cudaMemcpyDeviceToHost, d.stream));
  SPB_CUDA(ctx, cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, d.stream));
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  if (herr || counts[0] != counts[1]) return set_error(ctx, SPB_ERR_CONSTRAINT, "permute_expression_pair: an input value does not occur in the table (ConstraintSystemFailure)");
  uint32_t* rep_rows = idx_b;   // free again
  lk_emit_input_kernel<<<nb(n), 256, 0, d.stream>>>(sin, repeated_flag, rep_rank, rep_rows, (Fr*)d_permuted_input, (Fr*)d_permuted_table, n);
  lk_emit_leftover_kernel<<<nb(n), 256, 0, d.stream>>>(stb, used, left_rank, rep_rows, counts[0], (Fr*)d_permuted_table, n);
  SPB_CUDA(ctx, cudaGetLastError());
  ctx->n_kernel_launches += 4;
  SPB_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return 0;
}

}  // extern "C"
