#include <cstdio>
#include <cstring>
#include <random>
#include <cmath>
#include <cstdint>
extern "C" int old_format(double x, char *buf, int cap);
extern "C" int new_format(double x, char *buf, int cap);
int main(){
  std::mt19937_64 rng(7); char a[64], b[64]; long n=0, bad=0;
  auto chk=[&](double x){ int la=old_format(x,a,64), lb=new_format(x,b,64); n++; if(la!=lb||memcmp(a,b,la)){ if(bad<10) printf("DIFF %.17g: %s vs %s\n",x,a,b); bad++; } };
  std::uniform_real_distribution<double> u(0,1); std::gamma_distribution<double> g(0.5,3.0);
  for(long i=0;i<4000000;i++){ chk(g(rng)); }
  for(long i=0;i<4000000;i++){ double m=u(rng)+0.5; int e=(int)(rng()%640)-320; chk(m*std::pow(10.0,e)); chk(-m*std::pow(10.0,e)); }
  for(long i=0;i<3000000;i++){ uint64_t bits=rng(); double x; memcpy(&x,&bits,8); if(std::isfinite(x)) chk(x); }
  for(long i=0;i<2000000;i++){ double x=(double)(rng()%100000000)/ (double)(1+rng()%1000); chk(x); chk(std::round(x*1000)/1000); }
  for(int e=-323;e<=308;e++){ double p=std::pow(10.0,e); chk(p); chk(std::nextafter(p,0)); chk(std::nextafter(p,1e309)); chk(5*p); chk(9.99999999999999*p); chk(9.999999999999995*p);}
  chk(5e-324); chk(1.7976931348623157e308); chk(2.2250738585072014e-308);
  printf("compared %ld values, %ld differences\n", n, bad);
  return bad!=0;
}
