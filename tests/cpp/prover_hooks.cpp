// ctypes hooks into include/spectre_b200_prover.hpp for the CPU tests (tests/test_cpp_prover.py).
#include "../../include/spectre_b200_prover.hpp"
using namespace halo2;
using hostfield::U256;
static const hostfield::Params& P(int fq) { return fq ? hostfield::fq_params() : hostfield::fr_params(); }
static U256 ld(const uint64_t* a) { return {a[0], a[1], a[2], a[3]}; }
static void st(const U256& v, uint64_t* o) { for (int i = 0; i < 4; i++) o[i] = v[i]; }
extern "C" {
// canonical in, canonical out, through the Montgomery routines
void ph_mul(int fq, const uint64_t* a, const uint64_t* b, uint64_t* o) { st(hostfield::from_mont(P(fq), hostfield::mul(P(fq), hostfield::to_mont(P(fq), ld(a)), hostfield::to_mont(P(fq), ld(b)))), o); }
void ph_add(int fq, const uint64_t* a, const uint64_t* b, uint64_t* o) { st(hostfield::add(P(fq), ld(a), ld(b)), o); }
void ph_sub(int fq, const uint64_t* a, const uint64_t* b, uint64_t* o) { st(hostfield::sub(P(fq), ld(a), ld(b)), o); }
void ph_inv(int fq, const uint64_t* a, uint64_t* o) { st(hostfield::from_mont(P(fq), hostfield::inv(P(fq), hostfield::to_mont(P(fq), ld(a)))), o); }
void ph_pow(int fq, const uint64_t* a, uint64_t e, uint64_t* o) { st(hostfield::from_mont(P(fq), hostfield::pow_u64(P(fq), hostfield::to_mont(P(fq), ld(a)), e)), o); }
void ph_keccak(const uint8_t* d, size_t n, uint8_t* out) { auto h = keccak256(d, n); memcpy(out, h.data(), 32); }
// script: ops[i] = 0 common_scalar, 1 write_scalar, 2 write_point (two values), 3 squeeze. Returns proof length; challenges -> chal
size_t ph_transcript(const uint64_t* digest, const int* ops, size_t n_ops, const uint64_t* vals, uint64_t* chal, uint8_t* proof, size_t* absorbed) {
  EvmTranscriptWrite t(ld(digest));
  size_t v = 0, c = 0;
  for (size_t i = 0; i < n_ops; i++) {
    if (ops[i] == 0) t.common_scalar(ld(vals + 4 * v++));
    else if (ops[i] == 1) t.write_scalar(ld(vals + 4 * v++));
    else if (ops[i] == 2) { t.write_ec_point(ld(vals + 4 * v), ld(vals + 4 * v + 4)); v += 2; }
    else st(t.squeeze_challenge(), chal + 4 * c++);
  }
  memcpy(proof, t.proof().data(), t.proof().size());
  for (size_t i = 0; i < t.absorbed().size(); i++) absorbed[i] = t.absorbed()[i];
  return t.proof().size();
}
// Poseidon: permutation of `t` canonical elements in place; the transcript script as above (ops 0..3), over PoseidonTranscriptWrite
void ph_poseidon_permute(uint32_t t, uint32_t r_f, uint32_t r_p, uint64_t* state) {
  PoseidonSpec spec(t, r_f, r_p);
  std::vector<U256> s;
  for (uint32_t i = 0; i < t; i++) s.push_back(hostfield::to_mont(hostfield::fr_params(), ld(state + 4 * i)));
  spec.permute(s);
  for (uint32_t i = 0; i < t; i++) st(hostfield::from_mont(hostfield::fr_params(), s[i]), state + 4 * i);
}
size_t ph_poseidon_transcript(const uint64_t* digest, const int* ops, size_t n_ops, const uint64_t* vals, uint64_t* chal, uint8_t* proof) {
  PoseidonTranscriptWrite t(ld(digest));
  size_t v = 0, c = 0;
  for (size_t i = 0; i < n_ops; i++) {
    if (ops[i] == 0) t.common_scalar(ld(vals + 4 * v++));
    else if (ops[i] == 1) t.write_scalar(ld(vals + 4 * v++));
    else if (ops[i] == 2) { t.write_ec_point(ld(vals + 4 * v), ld(vals + 4 * v + 4)); v += 2; }
    else st(t.squeeze_challenge(), chal + 4 * c++);
  }
  memcpy(proof, t.proof().data(), t.proof().size());
  return t.proof().size();
}
}
