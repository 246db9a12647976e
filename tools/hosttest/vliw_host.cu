// CPU unit test of the cooperative pairing engine (vliw.cuh) and of fp_inv_bingcd: the same verdicts as the
// scalar path on the reference's verify_kzg_proof vectors.  Built and run by tests/test_host_cuda_logic.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "verify.cuh"
#include "vliw.cuh"
using namespace kzgb200;
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); std::vector<uint8_t> s(12 + 4096*48 + 65*96); if (fread(s.data(), 1, s.size(), f) != s.size()) return 2; fclose(f);
  G2Affine tau, gen;
  if (!g2_from_compressed_unchecked(gen, s.data() + 12 + 4096*48) || !g2_from_compressed_unchecked(tau, s.data() + 12 + 4096*48 + 96)) return 3;
  static PairingTables T; prepare_g2(T.g2_gen, gen); prepare_g2(T.tau_g2, tau);
  // inversion self-check
  Fp x = Fp::from_u32(123456789u);
  for (int i = 0; i < 50; i++) { x = x * x + Fp::from_u32(i + 3); Fp a = vliw::fp_inv_bingcd(x), b = fp_inv(x); if (!(a == b)) { puts("bingcd mismatch"); return 4; } }
  std::vector<Fp> regs(vliw::kTotalRegs);
  vliw::Lanes L{0, 1, vliw::default_tables()};
  uint8_t rec[160];
  while (fread(rec, 1, 160, stdin) == 160) {
    Fr z, y; G1Affine C, pi; int v;
    if (!scalar_from_be32_checked(z, rec + 48) || !scalar_from_be32_checked(y, rec + 80) || !g1_from_compressed(C, rec, true) || !g1_from_compressed(pi, rec + 112, true)) v = 2;
    else {
      G1Affine X = kzg_lhs_point(C, z, y, pi), npi = pi; if (!npi.inf) npi.y = npi.y.neg();
      v = vliw::coop_pairing_product_is_one(regs.data(), X, T.g2_gen, npi, T.tau_g2, L) ? 1 : 0;
    }
    putchar('0' + v);
  }
  putchar('\n');
}
