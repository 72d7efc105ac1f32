// Sanitizer harness for csrc/delaunay_host.cuh (host-only code), built by tests/test_delaunay.py with
// g++ -fsanitize=address,undefined: random, duplicate, collinear, huge, denormal, NaN and Inf point sets must neither
// touch memory out of bounds nor hang, and every triangle must index a real point.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <limits>
#include "../homography.js_b200/csrc/delaunay_host.cuh"
static uint64_t st = 0x1234567ull;
static uint32_t rnd(){ st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); }
static double u(){ return rnd() / 4294967296.0; }
int main(int argc,char**argv){
  int iters = argc>1?atoi(argv[1]):1000; long tot=0;
  for(int it=0; it<iters; ++it){
    int n = rnd()%40; if (it%50==0) n = 500 + rnd()%500;
    std::vector<double> p(2*n);
    int mode = rnd()%8;
    for(int i=0;i<2*n;++i){
      double v;
      switch(mode){
        case 0: v = u()*1000; break;
        case 1: v = (double)(rnd()%5); break;            // many duplicates / collinear
        case 2: v = (i&1)? 3.0 : u()*10; break;           // all collinear
        case 3: v = u()*1e300; break;
        case 4: v = (rnd()%7==0)? NAN : u()*100; break;
        case 5: v = (rnd()%7==0)? INFINITY : u()*100; break;
        case 6: v = (double)(rnd()%3) * 1e-320; break;     // denormals
        default: v = std::floor(u()*8)/8; break;
      }
      p[i]=v;
    }
    std::vector<uint32_t> t = hg_delaunay_detail::triangulate(p.data(), (size_t)n);
    for (uint32_t id : t) if (id >= (uint32_t)n) { printf("bad id\n"); return 3; }
    if (t.size()%3) { printf("bad size\n"); return 4; }
    tot += t.size()/3;
  }
  printf("triangles %ld\n", tot); return 0; }
