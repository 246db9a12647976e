// CPU unit test of the 29-bit-limb cooperative pairing engine (fp29.cuh, vliw29.cuh): Montgomery products against the
// 12 x 32 field layer, conversions, inversion, the sequential reference executors of the two engine instructions on signed
// representatives, and the same verdicts as the scalar path on the reference's verify_kzg_proof vectors (the engine programs
// run on the reference executors here; the 16-lane device forms are compared with them on the GPU).
// Built and run by tests/test_host_cuda_logic.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "verify.cuh"
#include "vliw29.cuh"
using namespace kzgb200;
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); std::vector<uint8_t> s(12 + 4096*48 + 65*96); if (fread(s.data(), 1, s.size(), f) != s.size()) return 2; fclose(f);
  G2Affine tau, gen;
  if (!g2_from_compressed_unchecked(gen, s.data() + 12 + 4096*48) || !g2_from_compressed_unchecked(tau, s.data() + 12 + 4096*48 + 96)) return 3;
  static PairingTables T; prepare_g2(T.g2_gen, gen); prepare_g2(T.tau_g2, tau);
  static vliw29::LineCoeffs29 L1[kMillerSteps], L2[kMillerSteps];
  for (int k = 0; k < kMillerSteps; k++)
    for (int e = 0; e < 6; e++) {
      const LineCoeffs& a = T.g2_gen[k]; const LineCoeffs& b = T.tau_g2[k];
      const Fp2& fa = e < 2 ? a.A : (e < 4 ? a.B : a.C); const Fp2& fb = e < 2 ? b.A : (e < 4 ? b.B : b.C);
      L1[k].v[e] = f29::from_fp((e & 1) ? fa.c1 : fa.c0); L2[k].v[e] = f29::from_fp((e & 1) ? fb.c1 : fb.c0);
    }
  // field self-check: products, subtracted dual products, round trips, inversion
  Fp x = Fp::from_u32(123456789u), y = Fp::from_u32(987654321u);
  for (int i = 0; i < 200; i++) {
    x = x * x + Fp::from_u32(i + 3); y = y * x + Fp::from_u32(7 * i + 1);
    f29::F29 a = f29::from_fp(x), b = f29::from_fp(y), r;
    Fp back = f29::canonical(a);
    if (!(back == x.to_raw())) { puts("round trip mismatch"); return 4; }
    if (!(f29::canonical(f29::mul29(a, b)) == (x * y).to_raw())) { puts("mul mismatch"); return 4; }
    f29::mont_mul29(r.l, a.l, b.l, b.l, b.l, true, false, 0); r.l[14] = r.l[15] = 0;
    if (!(f29::canonical(r) == (x * y + y * y).to_raw())) { puts("dual mismatch"); return 4; }
    f29::mont_mul29(r.l, a.l, b.l, b.l, b.l, true, true, 2); r.l[14] = r.l[15] = 0;
    if (!(f29::canonical(r) == (x * y - y * y).to_raw())) { puts("neg dual mismatch"); return 4; }
    if (!(f29::canonical(vliw29::inv29(a)) == vliw::fp_inv_bingcd(x).to_raw())) { puts("inverse mismatch"); return 4; }
    // reference executors on signed representatives: r3 = a b - b b ; r4 = r3 - 2 a + b (reduced)
    f29::F29 file[8]; file[0] = a; file[1] = b;
    const uint32_t mul_ins[4] = {3u | (0u << 16), 1u | (1u << 16), 1u | (3u << 16), 0u};
    vliw29::exec_mul_ref(file, mul_ins);
    if (!(vliw29::canonical_signed(file[3]) == (x * y - y * y).to_raw())) { puts("exec_mul_ref mismatch"); return 4; }
    const uint32_t terms[3] = {3u * 64u | (1u << 16), 0u * 64u | (0xfffeu << 16), 1u * 64u | (1u << 16)};
    for (uint32_t red = 0; red < 2; red++) {
      const uint32_t lin_ins[4] = {4u, 0u, 3u, red};
      vliw29::exec_lin_ref(file, lin_ins, terms);
      if (!(vliw29::canonical_signed(file[4]) == (x * y - y * y - x - x + y).to_raw())) { puts("exec_lin_ref mismatch"); return 4; }
    }
  }
  std::vector<f29::F29> regs(vliw29::kTotalRegs);
  vliw29::Lanes L{0, 1, vliw29::default_tables()};
  uint8_t rec[160];
  while (fread(rec, 1, 160, stdin) == 160) {
    Fr z, yy; G1Affine C, pi; int v;
    if (!scalar_from_be32_checked(z, rec + 48) || !scalar_from_be32_checked(yy, rec + 80) || !g1_from_compressed(C, rec, true) || !g1_from_compressed(pi, rec + 112, true)) v = 2;
    else {
      G1Affine X = kzg_lhs_point(C, z, yy, pi), npi = pi; if (!npi.inf) npi.y = npi.y.neg();
      v = vliw29::coop_pairing_product_is_one(regs.data(), X, L1, npi, L2, L) ? 1 : 0;
    }
    putchar('0' + v);
  }
  putchar('\n');
}
