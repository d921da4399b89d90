// Host build of the DEVICE arithmetic headers, for CPU-only unit tests (tests/test_hostemu.py).
// Compiled twice: with -DSPB_EMULATE_PTX (the 32-bit-limb PTX carry-chain algorithms, instruction for
// instruction, with an emulated carry flag) and without (the 64-bit host-glue path).
#include "../../spectre_b200/csrc/curve.cuh"
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
