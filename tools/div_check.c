// Empirical check of the activation quantiser's division: fma(fma(-q0, s, x), inv, q0) with inv = fl(1/s), q0 = fl(x*inv)
// against the IEEE quotient x / s on 2e9 random pairs (bf16- and fp32-mantissa x).  gcc -O2 -ffp-contract=off div_check.c -lm
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static uint64_t st=88172645463325252ULL;
static inline uint64_t rnd(){st^=st<<13;st^=st>>7;st^=st<<17;return st;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
int main(){
  long bad=0,badr=0,n=0;
  for(long i=0;i<2000000000L;i++){
    uint32_t a=(uint32_t)rnd(), b=(uint32_t)rnd();
    // x: random sign, exponent in [-20,20], bf16 mantissa (7 bits) or full
    int full = i&1;
    uint32_t xm = full ? (a&0x7FFFFF) : (a&0x7F0000);
    uint32_t xe = 127 + ((a>>23)%20) - 10;
    float x=u2f(((a>>31)<<31)|(xe<<23)|xm);
    uint32_t se = 127 + ((b>>23)%12) - 10;
    float s=u2f((se<<23)|(b&0x7FFFFF));
    float inv=1.0f/s;
    float q0=x*inv;
    float e=fmaf(-q0,s,x);
    float q1=fmaf(e,inv,q0);
    float qt=x/s;
    n++;
    if(q1!=qt){bad++; if(rintf(q1)!=rintf(qt)) badr++; if(bad<5) printf("x=%a s=%a q1=%a qt=%a\n",x,s,q1,qt);}
  }
  printf("n=%ld bad=%ld badr=%ld\n",n,bad,badr);
}
