#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
static inline uint64_t rng(uint64_t *s){ uint64_t x=*s; x^=x<<13; x^=x>>7; x^=x<<17; return *s=x; }
static inline double fdiv(double x,double b,double y){ double q=x*y; double r=fma(-q,b,x); return fma(r,y,q); }
static inline uint64_t bits(double d){uint64_t u; memcpy(&u,&d,8); return u;}
int main(int argc,char**argv){
  uint64_t s=0x9E3779B97F4A7C15ULL ^ (argc>1?strtoull(argv[1],0,10):0); long bad=0; const double B=1e9, Y=1e-9;
  long n = argc>2?atol(argv[2]):400000000L;
  /* 1. integers a < 2^53 (ns values), uniform in log-ish scale and uniform */
  for(long i=0;i<n;i++){ uint64_t r=rng(&s); int sh=r&63; uint64_t a=(rng(&s)>>11)>> (sh%53); double x=(double)a;
     if(bits(fdiv(x,B,Y))!=bits(x/B)){ if(bad<5) printf("int bad %llu\n",(unsigned long long)a); bad++; } }
  printf("ints done bad=%ld\n",bad);
  /* 2. structured: multiples of 1e9 +- k, powers of two +- k */
  for(uint64_t m=0;m<9000000ULL;m+=1){ for(int k=-3;k<=3;k++){ uint64_t a=m*1000000000ULL+k; if((int64_t)a<0||a>=(1ULL<<53)) continue; double x=(double)a; if(bits(fdiv(x,B,Y))!=bits(x/B)) bad++; } }
  for(int e=1;e<53;e++) for(int k=-200000;k<=200000;k++){ int64_t a=(1LL<<e)+k; if(a<0) continue; double x=(double)a; if(bits(fdiv(x,B,Y))!=bits(x/B)) bad++; }
  printf("structured done bad=%ld\n",bad);
  /* 3. random doubles x in [1e-20,1e6] / random b in [1e-3, 1e12] with y=1/b */
  for(long i=0;i<n;i++){ double x=ldexp((double)(rng(&s)>>11), -53 - (int)(rng(&s)%70) + 20); double b=ldexp((double)((rng(&s)>>11)|(1ULL<<52)), -52 - 10 + (int)(rng(&s)%50)); double y=1.0/b;
     if(bits(fdiv(x,b,y))!=bits(x/b)){ if(bad<10) printf("gen bad %a / %a\n",x,b); bad++; } }
  printf("generic done bad=%ld\n",bad);
  /* 4. b with few significant bits (typical rates: 8, 10, 9.5, 0.1 recip...) and x = -log(1-u) style */
  double bs[]={8.0,10.0,9.5,9.7,0.1,1.0/0.1,7.0,50.0,512.0,8192.0,3e8,20.0,32.0,500.0,300.0,25.0,5.0,200.0,1.0/0.05,2e9};
  for(unsigned j=0;j<sizeof bs/sizeof*bs;j++){ double b=bs[j], y=1.0/b; for(long i=0;i<n/8;i++){ double u=(double)(rng(&s)>>11)*0x1p-53; double x=-log1p(-u); if(bits(fdiv(x,b,y))!=bits(x/b)) {bad++; if(bad<10) printf("rate bad %a / %a\n",x,b);} } }
  printf("rates done bad=%ld\n",bad);
  return bad!=0;
}
