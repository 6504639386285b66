// Developer micro-benchmark (not part of the product): which HBM bandwidth can a
// multi-stream read-modify-write of the SoA ensemble reach on this GPU, as a
// function of layout, occupancy and cache hints?  Used to set the target for K1a.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/membench tools/membench.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

struct Streams {
  double *in[8];
  double *out[8];
  unsigned *win, *wout;
};

template <int HINT> __device__ __forceinline__ double2 ld2(const double *p) {
  if (HINT) return __ldcs(reinterpret_cast<const double2 *>(p));
  return *reinterpret_cast<const double2 *>(p);
}
template <int HINT> __device__ __forceinline__ void st2(double *p, double2 v) {
  if (HINT) __stcs(reinterpret_cast<double2 *>(p), v);
  else *reinterpret_cast<double2 *>(p) = v;
}

// NS fp64 streams (+ one u32 stream if W), 2 elements per lane and iteration
template <int NS, int W, int HINT>
__global__ void __launch_bounds__(256) soaKernel(Streams S, long long n) {
  const long long groups = n / 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    double2 v[NS];
    uint2 w;
#pragma unroll
    for (int c = 0; c < NS; c++) v[c] = ld2<HINT>(S.in[c] + 2 * g);
    if (W) w = __ldcs(reinterpret_cast<const uint2 *>(S.win + 2 * g));
#pragma unroll
    for (int c = 0; c < NS; c++) {
      v[c].x = fma(v[c].x, 1.0000001, 1e-9);
      v[c].y = fma(v[c].y, 1.0000001, 1e-9);
    }
#pragma unroll
    for (int c = 0; c < NS; c++) st2<HINT>(S.out[c] + 2 * g, v[c]);
    if (W) __stcs(reinterpret_cast<uint2 *>(S.wout + 2 * g), w);
  }
}

// Array-of-tiles layout: tile t holds NS fields x T particles contiguously
template <int NS, int T>
__global__ void __launch_bounds__(256) tileKernel(double *in, double *out, long long nTiles) {
  for (long long t = blockIdx.x; t < nTiles; t += gridDim.x) {
    double *src = in + t * (long long)(NS * T);
    double *dst = out + t * (long long)(NS * T);
    for (int j = threadIdx.x; j < T / 2; j += blockDim.x) {
      double2 v[NS];
#pragma unroll
      for (int c = 0; c < NS; c++) v[c] = __ldcs(reinterpret_cast<const double2 *>(src + c * T + 2 * j));
#pragma unroll
      for (int c = 0; c < NS; c++) {
        v[c].x = fma(v[c].x, 1.0000001, 1e-9);
        v[c].y = fma(v[c].y, 1.0000001, 1e-9);
      }
#pragma unroll
      for (int c = 0; c < NS; c++) __stcs(reinterpret_cast<double2 *>(dst + c * T + 2 * j), v[c]);
    }
  }
}

template <typename F> float timeIt(F f, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char **argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 100000000LL;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, n = %lld\n", prop.name, sms, n);
  const size_t strideB = (n * 8 + 255) & ~size_t(255);
  double *A, *B;
  unsigned *WA, *WB;
  CK(cudaMalloc(&A, strideB * 8));
  CK(cudaMalloc(&B, strideB * 8));
  CK(cudaMalloc(&WA, n * 4));
  CK(cudaMalloc(&WB, n * 4));
  CK(cudaMemset(A, 0, strideB * 8));
  CK(cudaMemset(B, 0, strideB * 8));
  CK(cudaMemset(WA, 0, n * 4));
  Streams inplace{}, outplace{};
  for (int c = 0; c < 8; c++) {
    inplace.in[c] = inplace.out[c] = (double *)((char *)A + strideB * c);
    outplace.in[c] = (double *)((char *)A + strideB * c);
    outplace.out[c] = (double *)((char *)B + strideB * c);
  }
  inplace.win = inplace.wout = WA;
  outplace.win = WA;
  outplace.wout = WB;

  auto report = [&](const char *name, float ms, double bytes) {
    printf("%-58s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9);
    fflush(stdout);
  };
  for (int perSm : {2, 4, 8}) {
    const int grid = sms * perSm;
    char buf[128];
    snprintf(buf, sizeof buf, "1 stream out-of-place, %d CTA/SM, cs hints", perSm);
    report(buf, timeIt([&] { soaKernel<1, 0, 1><<<grid, 256>>>(outplace, n); }), 16.0 * n);
    snprintf(buf, sizeof buf, "1 stream in-place, %d CTA/SM, cs hints", perSm);
    report(buf, timeIt([&] { soaKernel<1, 0, 1><<<grid, 256>>>(inplace, n); }), 16.0 * n);
    snprintf(buf, sizeof buf, "8 f64 + 1 u32 streams in-place, %d CTA/SM, cs hints", perSm);
    report(buf, timeIt([&] { soaKernel<8, 1, 1><<<grid, 256>>>(inplace, n); }), 136.0 * n);
    snprintf(buf, sizeof buf, "8 f64 + 1 u32 streams in-place, %d CTA/SM, default", perSm);
    report(buf, timeIt([&] { soaKernel<8, 1, 0><<<grid, 256>>>(inplace, n); }), 136.0 * n);
    snprintf(buf, sizeof buf, "8 f64 + 1 u32 streams out-of-place, %d CTA/SM, cs hints", perSm);
    report(buf, timeIt([&] { soaKernel<8, 1, 1><<<grid, 256>>>(outplace, n); }), 136.0 * n);
    snprintf(buf, sizeof buf, "8 f64 tiles of 512 in-place, %d CTA/SM", perSm);
    report(buf, timeIt([&] { tileKernel<8, 512><<<grid, 256>>>(A, A, n / 512); }), 128.0 * (n / 512 * 512));
    snprintf(buf, sizeof buf, "8 f64 tiles of 2048 in-place, %d CTA/SM", perSm);
    report(buf, timeIt([&] { tileKernel<8, 2048><<<grid, 256>>>(A, A, n / 2048); }), 128.0 * (n / 2048 * 2048));
    snprintf(buf, sizeof buf, "8 f64 tiles of 2048 out-of-place, %d CTA/SM", perSm);
    report(buf, timeIt([&] { tileKernel<8, 2048><<<grid, 256>>>(A, B, n / 2048); }), 128.0 * (n / 2048 * 2048));
  }
  // reference: cudaMemcpy device to device of the same volume
  report("cudaMemcpyAsync D2D 6.4 GB", timeIt([&] { cudaMemcpyAsync(B, A, strideB * 8, cudaMemcpyDeviceToDevice); }),
         2.0 * strideB * 8);
  CK(cudaDeviceSynchronize());
  return 0;
}
