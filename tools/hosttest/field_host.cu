#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "field.cuh"
using namespace kzgb200;
// usage: reads lines "F a b c d" hex from stdin; prints mont_mul(a,b), dual(a,b,c,d), a+b, a-b, a^2, c^2
template<class T> void rd(T& x, const char* h){ int N=T::N; for(int i=0;i<N;i++){ char buf[9]; memcpy(buf,h+(N-1-i)*8,8); buf[8]=0; x.l[i]=strtoul(buf,0,16);} }
template<class T> void pr(const T& x){ for(int i=T::N-1;i>=0;i--) printf("%08x",x.l[i]); printf(" "); }
template<class T> void run(char* a,char* b,char* c,char* d){ T A,B,C,D; rd(A,a);rd(B,b);rd(C,c);rd(D,d); pr(A*B); pr(T::mul_dual(A,B,C,D)); pr(A+B); pr(A-B); pr(A.sqr()); pr(C.sqr_inl()); printf("\n"); }
#include <cstring>
int main(){ char f[8]; static char a[200],b[200],c[200],d[200];
  while(scanf("%s %s %s %s %s",f,a,b,c,d)==5){ if(f[0]=='r') run<Fr>(a,b,c,d); else run<Fp>(a,b,c,d);} }
