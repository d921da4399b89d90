// Compile/link check and (on a GPU box) a tiny run of the C++ host mirror include/spectre_b200.hpp.
// Usage: host_mirror [run]   -- without "run" it only proves the header compiles and every symbol links.
#include <cstdio>
#include <cstring>
#include "../../include/spectre_b200.hpp"

int main(int argc, char** argv) {
  if (argc < 2 || std::strcmp(argv[1], "run") != 0) { std::printf("linked\n"); return 0; }
  try {
    halo2::Backend be({0});
    const uint32_t k = 8;
    halo2::Fr s{};  s.l[0] = 5;                         // some Montgomery residue as the "secret"
    auto params = halo2::poly::kzg::ParamsKZG::setup(be, k, s);
    std::vector<halo2::Fr> poly(1u << k);
    for (size_t i = 0; i < poly.size(); i++) { poly[i] = halo2::Fr{}; poly[i].l[0] = i * 7 + 1; }
    halo2::G1 c1 = params.commit_lagrange(poly);
    auto g = params.get_g(SPB_BASIS_G_LAGRANGE);
    halo2::G1 c2 = halo2::arithmetic::best_multiexp(be, poly, g);
    if (std::memcmp(&c1, &c2, sizeof c1) != 0) { std::printf("MISMATCH commit_lagrange vs best_multiexp\n"); return 1; }
    halo2::poly::EvaluationDomain dom(be, 4, k);
    auto coeff = poly; dom.lagrange_to_coeff(coeff); dom.coeff_to_lagrange(coeff);
    if (std::memcmp(coeff.data(), poly.data(), poly.size() * sizeof(halo2::Fr)) != 0) { std::printf("MISMATCH lagrange round trip\n"); return 1; }
    auto ext = dom.coeff_to_extended(poly);
    auto back = dom.extended_to_coeff(ext);
    if (std::memcmp(back.data(), poly.data(), poly.size() * sizeof(halo2::Fr)) != 0) { std::printf("MISMATCH extended round trip\n"); return 1; }
    bool threw = false;
    try { std::vector<halo2::G1Affine> two(2); halo2::arithmetic::best_multiexp(be, poly, two); } catch (const std::invalid_argument&) { threw = true; }
    if (!threw) { std::printf("length mismatch did not throw\n"); return 1; }
    std::printf("host mirror ok\n");
    return 0;
  } catch (const std::exception& e) { std::printf("exception: %s\n", e.what()); return 2; }
}
