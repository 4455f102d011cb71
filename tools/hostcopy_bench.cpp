// 64 copies of 512 KiB through host_copy (vkhel_b200/csrc/hostcopy.cu), the staging copy of
// vkhel_vector_copy_from_host, alone: make build/bin/hostcopy_bench; VKHEL_COPY_THREADS=1..8 build/bin/hostcopy_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <time.h>
extern "C" void host_copy(void *dst, const void *src, size_t bytes);
static double now(){timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec*1e6+t.tv_nsec*1e-3;}
int main(){ size_t n=512<<10; char*src=(char*)malloc(64*n); char*dst=(char*)malloc(16*n); memset(src,1,64*n); memset(dst,2,16*n);
 for(int r=0;r<6;r++){double t0=now(); for(int v=0;v<64;v++) host_copy(dst+(v%16)*n, src+v*n, n); double dt=now()-t0; printf("%.0f us per 64 x 512 KiB = %.1f GB/s\n",dt,64.0*n/dt*1e-3);} }
