// spectre_b200_prover.hpp -- compiled (C++17) host side of the device-resident proof pipeline: the counterpart of
// spectre_b200/plonk.py and transcript.py, i.e. the part of halo2's `create_proof` that stays on the host when every
// polynomial lives in HBM (protocol order, transcript, RNG draws, rotation-set bookkeeping, expression flattening).
//
//   [UPSTREAM] halo2_proofs/src/plonk/{keygen.rs, prover.rs}, poly/kzg/multiopen/shplonk.rs (construct_intermediate_sets),
//   [UPSTREAM] snark-verifier/src/system/halo2/transcript/evm.rs (EvmTranscript),
//   reached in the reference from lightclient-circuits/src/util/circuit.rs:131,158,211.
//
// Header-only over the C ABI of spectre_b200.h. Device memory is reached through the small `DeviceMemory` interface (a
// CUDA-runtime implementation is provided when the including translation unit defines SPB_PROVER_WITH_CUDART and links
// cudart); the CPU tests bind it to host memory and to a test-only shim of the C ABI (tests/abi_shim), so the driver's
// logic is exercised without a GPU and must reproduce the Python driver's proof bytes exactly.
//
// STATUS (round 1): validated on the CPU against the Python driver through the shim; not yet run against libspectre_b200.so
// on a GPU (the Python driver is the GPU-validated one).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "spectre_b200.h"

namespace halo2 {
namespace hostfield {

// ---- 256-bit Montgomery arithmetic on the host (challenge / point bookkeeping only: a few hundred operations per proof) ----
struct Params { uint64_t mod[4], r[4], r2[4], inv; };
inline const Params& fr_params() {
  static const Params p = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                           {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full},
                           {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}, 0xc2e1f593efffffffull};
  return p;
}
inline const Params& fq_params() {
  static const Params p = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                           {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full},
                           {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}, 0x87d20782e4866389ull};
  return p;
}
using U256 = std::array<uint64_t, 4>;

inline bool geq(const U256& a, const uint64_t* m) {
  for (int i = 3; i >= 0; i--) { if (a[i] != m[i]) return a[i] > m[i]; }
  return true;
}
inline void sub_in_place(U256& a, const uint64_t* m) {
  unsigned __int128 borrow = 0;
  for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)a[i] - m[i] - (uint64_t)borrow; a[i] = (uint64_t)d; borrow = (d >> 64) & 1; }
}
inline U256 add(const Params& P, const U256& a, const U256& b) {
  U256 r; unsigned __int128 c = 0;
  for (int i = 0; i < 4; i++) { c += (unsigned __int128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  if (c || geq(r, P.mod)) sub_in_place(r, P.mod);
  return r;
}
inline U256 neg(const Params& P, const U256& a) {
  if (!(a[0] | a[1] | a[2] | a[3])) return a;
  U256 r = {P.mod[0], P.mod[1], P.mod[2], P.mod[3]};
  sub_in_place(r, a.data());
  return r;
}
inline U256 sub(const Params& P, const U256& a, const U256& b) { return add(P, a, neg(P, b)); }
// Montgomery product a * b * 2^-256 mod m (CIOS)
inline U256 mul(const Params& P, const U256& a, const U256& b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    unsigned __int128 c = 0;
    for (int j = 0; j < 4; j++) { c += (unsigned __int128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * P.inv;
    c = (unsigned __int128)m * P.mod[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) { c += (unsigned __int128)m * P.mod[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  U256 r = {t[0], t[1], t[2], t[3]};
  if (t[4] || geq(r, P.mod)) sub_in_place(r, P.mod);
  return r;
}
inline U256 one(const Params& P) { return {P.r[0], P.r[1], P.r[2], P.r[3]}; }
inline U256 to_mont(const Params& P, const U256& canonical) { return mul(P, canonical, {P.r2[0], P.r2[1], P.r2[2], P.r2[3]}); }
inline U256 from_mont(const Params& P, const U256& a) { return mul(P, a, {1, 0, 0, 0}); }
inline U256 pow(const Params& P, U256 base, const U256& e) {
  U256 acc = one(P);
  for (int i = 0; i < 256; i++) {
    if ((e[i / 64] >> (i % 64)) & 1) acc = mul(P, acc, base);
    base = mul(P, base, base);
  }
  return acc;
}
inline U256 pow_u64(const Params& P, const U256& base, uint64_t e) { return pow(P, base, {e, 0, 0, 0}); }
inline U256 inv(const Params& P, const U256& a) {
  U256 e = {P.mod[0], P.mod[1], P.mod[2], P.mod[3]};
  e[0] -= 2;                                          // both moduli end in ...01 / ...47: no borrow
  return pow(P, a, e);
}
inline U256 from_u64(const Params& P, uint64_t v) { return to_mont(P, {v, 0, 0, 0}); }
// 32 big-endian bytes -> canonical integer (not reduced) and back
inline U256 from_be(const uint8_t* b) {
  U256 r = {0, 0, 0, 0};
  for (int i = 0; i < 32; i++) r[3 - i / 8] |= (uint64_t)b[i] << (8 * (7 - i % 8));
  return r;
}
inline void to_be(const U256& a, uint8_t* b) {
  for (int i = 0; i < 32; i++) b[i] = (uint8_t)(a[3 - i / 8] >> (8 * (7 - i % 8)));
}
// reduce a 256-bit integer mod m (at most a few subtractions: inputs are < 2^256 < 6 m)
inline U256 reduce(const Params& P, U256 a) {
  while (geq(a, P.mod)) sub_in_place(a, P.mod);
  return a;
}
inline bool less(const U256& a, const U256& b) {      // canonical integers
  for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] < b[i]; }
  return false;
}

}  // namespace hostfield

// ---- Keccak-256 (the EVM's KECCAK256: original 0x01 padding) ------------------------------------------------------
inline std::array<uint8_t, 32> keccak256(const uint8_t* data, size_t len) {
  static const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull, 0x0000000080000001ull,
                                  0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000Aull,
                                  0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull, 0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull,
                                  0x000000000000800Aull, 0x800000008000000Aull, 0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
  auto rol = [](uint64_t v, int n) { n %= 64; return n ? (v << n) | (v >> (64 - n)) : v; };
  const size_t rate = 136;
  std::vector<uint8_t> msg(data, data + len);
  msg.push_back(0x01);
  while (msg.size() % rate) msg.push_back(0);
  msg.back() |= 0x80;
  uint64_t a[5][5] = {};
  for (size_t off = 0; off < msg.size(); off += rate) {
    for (size_t i = 0; i < rate / 8; i++) { uint64_t w; memcpy(&w, &msg[off + 8 * i], 8); a[i % 5][i / 5] ^= w; }   // little-endian host
    for (int round = 0; round < 24; round++) {
      uint64_t c[5], d[5], b[5][5];
      for (int x = 0; x < 5; x++) c[x] = a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4];
      for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
      for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) a[x][y] ^= d[x];
      for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) b[y][(2 * x + 3 * y) % 5] = rol(a[x][y], ROT[x][y]);
      for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) a[x][y] = b[x][y] ^ (~b[(x + 1) % 5][y] & b[(x + 2) % 5][y]);
      a[0][0] ^= RC[round];
    }
  }
  std::array<uint8_t, 32> out;
  for (int i = 0; i < 4; i++) memcpy(&out[8 * i], &a[i % 5][i / 5], 8);
  return out;
}

// ---- EvmTranscript (prover side) -------------------------------------------------------------------------------------
// Scalars cross this class as canonical integers (U256 < r); points as canonical affine coordinates.
class EvmTranscriptWrite {
 public:
  using U256 = hostfield::U256;
  explicit EvmTranscriptWrite(const U256& vk_digest) { buf_.resize(32); hostfield::to_be(vk_digest, buf_.data()); }
  void common_scalar(const U256& v) { append(v, buf_); }
  void common_ec_point(const U256& x, const U256& y) {
    if (!(x[0] | x[1] | x[2] | x[3] | y[0] | y[1] | y[2] | y[3])) throw std::runtime_error("EvmTranscript cannot absorb the point at infinity");
    append(x, buf_); append(y, buf_);
  }
  void write_scalar(const U256& v) { common_scalar(v); append(v, proof_); }
  void write_ec_point(const U256& x, const U256& y) { common_ec_point(x, y); append(x, proof_); append(y, proof_); }
  U256 squeeze_challenge() {
    std::vector<uint8_t> data = buf_;
    if (buf_.size() == 32) data.push_back(1);
    absorbed_.push_back(data.size());
    auto h = keccak256(data.data(), data.size());
    buf_.assign(h.begin(), h.end());
    return hostfield::reduce(hostfield::fr_params(), hostfield::from_be(h.data()));
  }
  const std::vector<uint8_t>& proof() const { return proof_; }
  const std::vector<size_t>& absorbed() const { return absorbed_; }

 private:
  static void append(const U256& v, std::vector<uint8_t>& dst) { uint8_t b[32]; hostfield::to_be(v, b); dst.insert(dst.end(), b, b + 32); }
  std::vector<uint8_t> buf_, proof_;
  std::vector<size_t> absorbed_;
};

}  // namespace halo2
