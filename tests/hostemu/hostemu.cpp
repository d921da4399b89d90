// Host build of the DEVICE arithmetic headers, for CPU-only unit tests (tests/test_hostemu.py).
// Compiled twice: with -DSPB_EMULATE_PTX (the 32-bit-limb PTX carry-chain algorithms, instruction for
// instruction, with an emulated carry flag) and without (the 64-bit host-glue path).
#include "../../spectre_b200/csrc/curve.cuh"
#include <string.h>
using namespace spb;
extern "C" {
void he_fr_mul(Fr* o, const Fr* a, const Fr* b, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_mul(a[i], b[i]); }
void he_fr_add(Fr* o, const Fr* a, const Fr* b, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_add(a[i], b[i]); }
void he_fr_sub(Fr* o, const Fr* a, const Fr* b, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_sub(a[i], b[i]); }
void he_fr_neg(Fr* o, const Fr* a, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_neg(a[i]); }
void he_fr_inv(Fr* o, const Fr* a, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_inv(a[i]); }
void he_fq_mul(Fq* o, const Fq* a, const Fq* b, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_mul(a[i], b[i]); }
void he_fq_add(Fq* o, const Fq* a, const Fq* b, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_add(a[i], b[i]); }
void he_fq_sub(Fq* o, const Fq* a, const Fq* b, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_sub(a[i], b[i]); }
void he_fq_inv(Fq* o, const Fq* a, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_inv(a[i]); }
void he_fr_from_mont(Fr* o, const Fr* a, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fp_from_mont(a[i]); }
// acc (XYZZ, starts at identity) += each affine point in order; returns affine
void he_sum_mixed(G1Affine* out, const G1Affine* pts, size_t n) {
  G1Xyzz acc = xyzz_identity();
  for (size_t i = 0; i < n; i++) xyzz_add_mixed(acc, pts[i]);
  *out = xyzz_to_affine(acc);
}
// pairwise tree through the general XYZZ+XYZZ adder
void he_sum_general(G1Affine* out, const G1Affine* pts, size_t n) {
  G1Xyzz acc = xyzz_identity();
  for (size_t i = 0; i < n; i++) { G1Xyzz t = xyzz_from_affine(pts[i]); xyzz_add(acc, t); }
  *out = xyzz_to_affine(acc);
}
void he_mul_u32(G1Affine* out, const G1Affine* p, uint32_t k) { *out = xyzz_to_affine(xyzz_mul_u32(xyzz_from_affine(*p), k)); }
void he_dbl(G1Affine* out, const G1Affine* p) { *out = xyzz_to_affine(xyzz_dbl(xyzz_from_affine(*p))); }
int he_on_curve(const G1Affine* p) { return affine_on_curve(*p) ? 1 : 0; }
void he_jac_roundtrip(G1Affine* out, const G1Affine* p) { G1Jac j = jac_from_affine(*p); G1Xyzz x = xyzz_from_jac(j); *out = jac_to_affine(jac_from_affine(xyzz_to_affine(x))); }
}

// ---- MSM pipeline, emulated serially with the same per-thread bodies the CUDA kernels call -----------------
#include "../../spectre_b200/csrc/msm.cuh"
#include <vector>
extern "C" {
// c = 0: library's own geometry choice. precomp = 1: build the 2^(c*j) tables first and use one bucket set.
// Returns the number of sorted entries; *giants = number of chains that took the block path.
uint64_t he_msm(G1Affine* out, const Fr* scalars, const G1Affine* bases, size_t n, uint32_t c, uint32_t L, uint32_t cap, uint32_t* giants, int precomp) {
  if (!c) c = msm_choose_c(n ? n : 1, precomp != 0);
  MsmGeom g = msm_make_geometry(c, precomp != 0, (uint32_t)n);
  if (L) g.L = L;
  std::vector<G1Affine> table;
  const G1Affine* pts = bases;
  if (precomp) {
    table.resize((size_t)g.W * n);
    for (size_t i = 0; i < n; i++) table[i] = bases[i];
    for (uint64_t t = 0; t < n; t++) msm_precompute_thread(t, n, g.c, g.W, table.data());
    pts = table.data();
  }
  uint64_t nb = (uint64_t)g.BW * g.B;
  std::vector<uint32_t> counts(nb + 1, 0), offsets(nb + 1, 0);
  for (uint64_t t = 0; t < n; t++) msm_count_thread(t, n, scalars, g, counts.data());
  uint64_t run = 0;
  for (uint64_t i = 0; i <= nb; i++) { offsets[i] = (uint32_t)run; run += counts[i]; }
  uint64_t M = offsets[nb];
  std::vector<uint32_t> cursor(offsets);
  std::vector<MsmEntry> ent(M + 1);
  for (uint64_t t = 0; t < n; t++) msm_scatter_thread(t, n, scalars, g, cursor.data(), ent.data());
  std::vector<G1Xyzz> buckets(nb);
  memset(buckets.data(), 0, nb * sizeof(G1Xyzz));
  uint64_t T = (M + g.L - 1) / g.L;
  std::vector<uint32_t> hk(T + 1, 0x12345678u), tk(T + 1, 0x12345678u), glist(T + 2);
  std::vector<G1Xyzz> head(T + 1), tail(T + 1);
  for (uint64_t t = 0; t < T; t++) msm_accumulate_thread(t, M, g, ent.data(), pts, buckets.data(), hk.data(), head.data(), tk.data(), tail.data());
  uint32_t gcount = 0;
  for (uint64_t t = 0; t < T; t++) msm_stitch_thread(t, T, cap, hk.data(), head.data(), tk.data(), tail.data(), buckets.data(), &gcount, glist.data());
  for (uint32_t gi = 0; gi < gcount; gi++) {  // what msm_giant_kernel does, serially
    uint64_t t0 = glist[gi]; uint32_t key = tk[t0];
    G1Xyzz acc = xyzz_identity();
    for (uint64_t j = t0 + 1; j < T && hk[j] == key; j++) xyzz_add(acc, head[j]);
    xyzz_add(acc, tail[t0]);
    buckets[key] = acc;
  }
  if (giants) *giants = gcount;
  MsmTail tl = msm_tail_shape(g.c);
  uint32_t per = msm_tail_partials(tl);
  std::vector<G1Xyzz> partials((uint64_t)g.BW * per), win(g.BW);
  msm_tail_host(g, buckets.data(), partials.data());
  for (uint32_t w = 0; w < g.BW; w++) win[w] = msm_tail_finish(g, partials.data() + (uint64_t)w * per);
  *out = xyzz_to_affine(msm_combine_windows(win.data(), g.BW, g.c));
  return M;
}
void he_geometry(uint64_t n, int precomp, uint32_t* c, uint32_t* W) { MsmGeom g = msm_make_geometry(msm_choose_c(n, precomp != 0), precomp != 0, 0); *c = g.c; *W = g.W; }
}
