// CPU test driver for include/spectre_b200_prover.hpp: reads a circuit instance dumped by tests/test_cpp_prover.py, runs
// keygen + create_proof through whatever implements the C ABI at link time (the test links tests/abi_shim) and writes the
// proof bytes. usage: prover_main <dir>
#include <cstdio>
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>

#ifdef SPB_PROVER_WITH_CUDART
#include <cuda_runtime.h>
#endif
#include "../../include/spectre_b200_prover.hpp"

using namespace halo2;
using namespace halo2::plonk;

static std::vector<char> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot read " + path);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static U256 parse_hex(const std::string& s) {
  U256 v = {0, 0, 0, 0};
  std::string h = s.substr(0, 2) == "0x" ? s.substr(2) : s;
  while (h.size() < 64) h = "0" + h;
  for (int i = 0; i < 4; i++) v[3 - i] = std::stoull(h.substr(16 * i, 16), nullptr, 16);
  return v;
}

// spectre_b200/circuits.py::aggregation_shape
static ConstraintSystem aggregation_shape() {
  ConstraintSystem cs; cs.num_fixed = 4; cs.num_advice = 1; cs.num_instance = 1;
  ExprP a[4]; for (int r = 0; r < 4; r++) a[r] = Advice(0, r);
  cs.gates = {Prod(Sum(Sum(a[0], Prod(a[1], a[2])), Neg(a[3])), Fixed(3))};
  cs.lookups = {Lookup{{Prod(Advice(0), Fixed(2))}, {Fixed(1)}}};
  cs.permutation = {{Col::Fixed, 0}, {Col::Advice, 0}, {Col::Instance, 0}};
  cs.fixed_queries = {{0, 0}, {1, 0}, {2, 0}, {3, 0}};
  cs.finalize();
  return cs;
}
// spectre_b200/circuits.py::halo2lib_shape(G, L, spread=True)
static ConstraintSystem halo2lib_shape(uint32_t G, uint32_t L) {
  ConstraintSystem cs; const uint32_t A = G + L + 2;
  cs.num_fixed = G + 4; cs.num_advice = A; cs.num_instance = 1;
  for (uint32_t c = 0; c < G; c++) {
    ExprP a[4]; for (int r = 0; r < 4; r++) a[r] = Advice(c, r);
    cs.gates.push_back(Prod(Fixed(c), Sum(Sum(a[0], Prod(a[1], a[2])), Neg(a[3]))));
  }
  for (uint32_t l = 0; l < L; l++) cs.lookups.push_back(Lookup{{Advice(G + l)}, {Fixed(G + 1)}});
  cs.lookups.push_back(Lookup{{Advice(G + L), Advice(G + L + 1)}, {Fixed(G + 2), Fixed(G + 3)}});
  for (uint32_t c = 0; c < A; c++) cs.permutation.push_back({Col::Advice, c});
  cs.permutation.push_back({Col::Fixed, G}); cs.permutation.push_back({Col::Instance, 0});
  cs.finalize();
  return cs;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: prover_main <dir>\n"); return 2; }
  try {
    const std::string dir = argv[1];
    std::ifstream meta(dir + "/meta.txt");
    std::string shape; uint32_t G = 0, L = 0, k = 0; U256 digest{}; std::vector<U256> inst; std::vector<std::pair<Cell, Cell>> copies; std::vector<size_t> rng_counts;
    std::string line;
    bool poseidon = false;                                       // "transcript poseidon": the inner snark's transcript instead of the EVM one
    uint8_t chacha_seed[32] = {0}; bool have_chacha = false;   // the random polynomial from the device ChaCha20 stream instead of rng.bin
    while (std::getline(meta, line)) {
      std::istringstream is(line); std::string key; is >> key;
      if (key == "shape") { is >> shape; if (shape == "halo2lib") is >> G >> L; }
      else if (key == "k") is >> k;
      else if (key == "digest") { std::string h; is >> h; digest = parse_hex(h); }
      else if (key == "instances") { std::string h; while (is >> h) inst.push_back(parse_hex(h)); }
      else if (key == "copy") { uint32_t c1, c2; uint64_t r1, r2; is >> c1 >> r1 >> c2 >> r2; copies.push_back({{c1, r1}, {c2, r2}}); }
      else if (key == "rng") { size_t c; while (is >> c) rng_counts.push_back(c); }
      else if (key == "transcript") { std::string name; is >> name; poseidon = name == "poseidon"; }
      else if (key == "chacha_poly") { std::string h; is >> h; if (h.size() != 64) throw std::runtime_error("chacha_poly: 64 hex digits expected"); for (int i = 0; i < 32; i++) chacha_seed[i] = (uint8_t)std::stoul(h.substr(2 * i, 2), nullptr, 16); have_chacha = true; }
    }
    ConstraintSystem cs = shape == "aggregation" ? aggregation_shape() : halo2lib_shape(G, L);
    const size_t n = (size_t)1 << k;
    auto fixed_raw = slurp(dir + "/fixed.bin"), advice_raw = slurp(dir + "/advice.bin"), rng_raw = slurp(dir + "/rng.bin");
    if (fixed_raw.size() != (size_t)cs.num_fixed * n * 32 || advice_raw.size() != (size_t)cs.num_advice * n * 32) throw std::runtime_error("column files have the wrong size");
    std::vector<const Fr*> fixed, advice;
    for (uint32_t c = 0; c < cs.num_fixed; c++) fixed.push_back((const Fr*)fixed_raw.data() + (size_t)c * n);
    for (uint32_t c = 0; c < cs.num_advice; c++) advice.push_back((const Fr*)advice_raw.data() + (size_t)c * n);

    spb_ctx* ctx = spb_init(nullptr, 1);
    if (!ctx) throw std::runtime_error("spb_init failed");
    std::ifstream tf(dir + "/tau.bin", std::ios::binary); Fr tau; tf.read((char*)&tau, 32);
    spb_srs* srs = nullptr;
    if (spb_srs_setup(ctx, k, &tau, &srs) != 0) throw std::runtime_error(std::string("spb_srs_setup: ") + spb_last_error(ctx));
#ifdef SPB_PROVER_WITH_CUDART
    if (getenv("SPB_MAIN_TABLES") && spb_srs_precompute(ctx, srs) != 0) throw std::runtime_error(std::string("spb_srs_precompute: ") + spb_last_error(ctx));
    CudaMemory mem(ctx); // the real library: device buffers through the CUDA runtime, on the context's stream
    // a prover keeps its synthesis buffers pinned: register the column / rng files once so uploads are plain DMA
    spb_host_register(ctx, fixed_raw.data(), fixed_raw.size());
    spb_host_register(ctx, advice_raw.data(), advice_raw.size());
    if (!rng_raw.empty()) spb_host_register(ctx, rng_raw.data(), rng_raw.size());
#else
    HostMemory mem;      // the CPU shim: "device" pointers are host pointers
#endif
    {
      Engine E(ctx, mem, srs, k, (uint32_t)cs.degree());
      auto t_k0 = std::chrono::steady_clock::now();
      ProvingKey pk = keygen(E, cs, fixed, copies, &digest);
      auto t_k1 = std::chrono::steady_clock::now();
      size_t call = 0, pos = 0;
      Rng rng = [&](size_t count, Fr* out) {
        if (call >= rng_counts.size() || rng_counts[call] != count) throw std::runtime_error("rng stream out of step at call " + std::to_string(call));
        memcpy(out, rng_raw.data() + pos * 32, count * 32); pos += count; call++;
      };
      // SPB_MAIN_REPEAT=r: prove r times from the same RNG stream (identical bytes each time); the last pass is the warm one
      const int repeat = getenv("SPB_MAIN_REPEAT") ? std::max(1, atoi(getenv("SPB_MAIN_REPEAT"))) : 1;
      std::vector<uint8_t> proof;
      std::printf("keygen_ms %.3f\n", std::chrono::duration<double, std::milli>(t_k1 - t_k0).count());
      for (int rep = 0; rep < repeat; rep++) {
        call = 0; pos = 0;
        auto t0 = std::chrono::steady_clock::now();
        BulkRng bulk = nullptr;
        if (have_chacha) bulk = [&](Engine& e, size_t count) { return e.random_chacha(chacha_seed, 0, count); };
        std::vector<uint8_t> pr;
        if (poseidon) { PoseidonTranscriptWrite T(pk.vk_digest); pr = create_proof(E, pk, {inst}, advice, rng, T, bulk); }
        else { EvmTranscriptWrite T(pk.vk_digest); pr = create_proof(E, pk, {inst}, advice, rng, T, bulk); }
        std::printf("create_proof_ms %.3f\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        if (rep && pr != proof) throw std::runtime_error("repeated proof differs");
        proof = pr;
      }
      std::ofstream out(dir + "/proof.bin", std::ios::binary);
      out.write((const char*)proof.data(), proof.size());
      std::ofstream vk(dir + "/vk.txt");
      for (auto* v : {&pk.fixed_commitments, &pk.sigma_commitments}) for (auto& p : *v) { uint8_t b[64]; hostfield::to_be(p.x, b); hostfield::to_be(p.y, b + 32); for (int i = 0; i < 64; i++) { char h[3]; std::snprintf(h, 3, "%02x", b[i]); vk << h; } vk << "\n"; }
      std::printf("proof %zu bytes, rng calls %zu/%zu\n", proof.size(), call, rng_counts.size());
    }
#ifdef SPB_PROVER_WITH_CUDART
    mem.trim();          // return the cached device blocks while the context (and its stream) still exist
    spb_host_unregister(ctx, fixed_raw.data()); spb_host_unregister(ctx, advice_raw.data());
    if (!rng_raw.empty()) spb_host_unregister(ctx, rng_raw.data());
#endif
    spb_srs_free(ctx, srs);
    spb_shutdown(ctx);
    return 0;
  } catch (const std::exception& e) { std::fprintf(stderr, "prover_main: %s\n", e.what()); return 1; }
}
