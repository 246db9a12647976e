// CPU unit test of the GLV scalar split and endomorphism (glv.cuh): k1 + k2 x^2 == k with k1 < x^2, k2 < 2^128, and
// [k]P == [k1]P + [k2](-phi(P)) on the G1 generator.  stdin: hex scalars (64 digits), one per line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "glv.cuh"
using namespace kzgb200;
int main() {
    char h[200];
    while (scanf("%s", h) == 1) {
        uint32_t k[8], k1[4], k2[4];
        for (int i = 0; i < 8; i++) { char b[9]; memcpy(b, h + (7 - i) * 8, 8); b[8] = 0; k[i] = strtoul(b, 0, 16); }
        glv_split(k, k1, k2);
        for (int i = 3; i >= 0; i--) printf("%08x", k1[i]);
        printf(" ");
        for (int i = 3; i >= 0; i--) printf("%08x", k2[i]);
        G1Affine g = g1_generator();
        G1 full = scalar_mul_affine(g, k, 256);
        G1 split = scalar_mul_affine(g, k1, 128).add(scalar_mul_affine(glv_endo_neg(g), k2, 128));
        printf(" %d\n", full.equals(split) ? 1 : 0);
    }
}
