// Dependent-chain latencies of the fp64 operations the SOR sweep is made of (one warp, one thread active).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void chain(double *out, double x0, double y, int n, long long *cycles) {
  double x = x0;
  const long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    if (OP == 0) x = __fma_rn(x, y, y);
    if (OP == 1) x = __dadd_rn(x, y);
    if (OP == 2) x = __dmul_rn(x, y);
    if (OP == 3) x = __ddiv_rn(y, x);
    if (OP == 4) x = exp(-x);
    if (OP == 5) x = __ddiv_rn(y, exp(-x));
    if (OP == 6) x = log(x + 2.0);
    if (OP == 7) { float f = (float)x; f = __fmaf_rn(f, 1.0001f, 0.5f); x = f; }
    if (OP == 8) x = __drcp_rn(x) + 1.0;
    if (OP == 9) x = sqrt(x) + 1.0;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const char *names[] = {"dfma", "dadd", "dmul", "ddiv", "exp", "exp+div", "log", "cvt+ffma+cvt", "drcp+add", "dsqrt+add"};
  const int n = 2000;
#define RUN(OP) for (int threads : {1, 32, 256, 1024}) { chain<OP><<<1, threads>>>(out, 0.7, 1.0000001, n, cyc); chain<OP><<<1, threads>>>(out, 0.7, 1.0000001, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-14s threads %4d : %7.1f cycles per op\n", names[OP], threads, double(h) / n); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
  return 0;
}
