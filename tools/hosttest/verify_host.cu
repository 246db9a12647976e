#include <cstdio>
#include <cstdlib>
#include <vector>
#include "verify.cuh"
using namespace kzgb200;
// argv[1] = setup bin ; stdin: records of 160 bytes (c48 z32 y32 p48) ; stdout: one digit per record
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); std::vector<uint8_t> s(12 + 4096*48 + 65*96); fread(s.data(), 1, s.size(), f); fclose(f);
  G2Affine tau, gen;
  if (!g2_from_compressed_unchecked(gen, s.data() + 12 + 4096*48)) { puts("bad g2[0]"); return 1; }
  if (!g2_from_compressed_unchecked(tau, s.data() + 12 + 4096*48 + 96)) { puts("bad g2[1]"); return 1; }
  G2Affine gg = g2_generator();
  if (!(gg.x == gen.x) || !(gg.y == gen.y)) { puts("generator mismatch"); return 1; }
  static PairingTables T; prepare_g2(T.g2_gen, gen); prepare_g2(T.tau_g2, tau);
  uint8_t rec[160];
  while (fread(rec, 1, 160, stdin) == 160) { putchar('0' + verify_kzg_proof_one(rec, rec+48, rec+80, rec+112, &T)); }
  putchar('\n');
}
