// CPU unit test of the cooperative pairing engine (vliw.cuh) and of fp_inv_bingcd: the same verdicts as the
// scalar path on the reference's verify_kzg_proof vectors.  Built and run by tests/test_host_cuda_logic.py.
#include <cstdio>
#include <cstring>
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
  static G1Affine gen_table[64][15];
  for (int w = 0; w < 64; w++) for (int d = 1; d <= 15; d++) { uint32_t k[8] = {0,0,0,0,0,0,0,0}; k[w / 8] = (uint32_t)d << (4 * (w % 8)); gen_table[w][d - 1] = g1_to_affine(scalar_mul_affine(g1_generator(), k, 256)); }
  std::vector<Fp> regs(vliw::kTotalRegs);
  vliw::Lanes L{0, 1, vliw::default_tables()};
  uint8_t rec[160];
  std::vector<G1Affine> mx, mp; std::vector<uint8_t> mv;
  while (fread(rec, 1, 160, stdin) == 160) {
    Fr z, y; G1Affine C, pi; int v;
    if (!scalar_from_be32_checked(z, rec + 48) || !scalar_from_be32_checked(y, rec + 80) || !g1_from_compressed(C, rec, true) || !g1_from_compressed(pi, rec + 112, true)) v = 2;
    else {
      G1Affine X = kzg_lhs_point(C, z, y, pi), npi = pi; if (!npi.inf) npi.y = npi.y.neg();
      // the table / GLV form of the same point (per-tuple batched path)
      G1Affine Xf = g1_to_affine(kzg_lhs_point_fast(C, z, y, pi, gen_table));
      if (Xf.inf != X.inf || (!X.inf && (!(Xf.x == X.x) || !(Xf.y == X.y)))) { puts("lhs mismatch"); return 5; }
      v = vliw::coop_pairing_product_is_one(regs.data(), X, T.g2_gen, npi, T.tau_g2, L) ? 1 : 0;
      if (!X.inf && !npi.inf) { mx.push_back(X); mp.push_back(npi); mv.push_back((uint8_t)v); }
    }
    putchar('0' + v);
  }
  // the same checks three at a time in lockstep (multi-group form used by the many-tuple kernel)
  const int G = 3;
  std::vector<Fp> mregs(G * (vliw::kTotalRegs + 1));
  static vliw::SharedTables mt;      // the multi-group form executes from a remapped copy of the throughput tables (shared registers)
  {
    vliw::Tables src = vliw::throughput_tables();
    memcpy(mt.mul, src.mul, sizeof(uint16_t) * 19 * vliw::thr::kNumMul); memcpy(mt.lin, src.lin, sizeof(uint32_t) * 3 * vliw::thr::kNumLin);
    memcpy(mt.term, src.term, sizeof(uint16_t) * vliw::thr::kNumTerm); memcpy(mt.level, src.level, sizeof(vliw::Level) * vliw::thr::kNumLevel);
    memcpy(mt.prog, src.prog, sizeof(vliw::Program) * vliw::kNumPrograms);
    vliw::remap_tables_multi(&mt, 0, 1);
  }
  static Fp mshared[vliw::kNumSharedRegs];
  vliw::Lanes LM{0, 1, vliw::Tables{mt.mul, mt.lin, mt.term, mt.level, mt.prog, false}};
  LM.groups = G; LM.stride = vliw::kTotalRegsThr * 12 + 1; LM.shared = mshared;
  for (size_t i = 0; i + G <= mx.size() && i < 12; i += G) {
    uint8_t ok[G];
    vliw::coop_pairing_multi(mregs.data(), &mx[i], T.g2_gen, &mp[i], T.tau_g2, LM, ok);
    for (int g = 0; g < G; g++) if (ok[g] != mv[i + g]) { puts("multi mismatch"); return 6; }
  }
  putchar('\n');
}
