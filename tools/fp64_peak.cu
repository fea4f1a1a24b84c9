// fp64_peak.cu -- measured fp64 FMA peak of the device (DFMA chains), the second ceiling of the VPFP kernels
// (the FFT operators and the Fokker-Planck solve sit on the fp64 ridge of the machine).
// Measurement utility only: built by __graft_entry__.build() into vlapy_b200/lib/libfp64_peak.so, called by
// tools/fp64_peak.py; not part of the product library.
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double a, double b) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = (double)threadIdx.x * 1e-9 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

template <int CHAINS>
static double run(int blocks, int threads, int iters, double* scratch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma_kernel<CHAINS><<<blocks, threads>>>(scratch, iters / 8, 0.999999, 1e-7);   // warm-up
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<CHAINS><<<blocks, threads>>>(scratch, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return best;
}

// returns TFLOP/s (2 flops per FMA) for `warps_per_sm` resident warps per SM and `chains` independent chains per thread
extern "C" double fp64_peak_tflops(int warps_per_sm, int chains, int iters) {
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int threads = (warps_per_sm >= 32 ? 1024 : warps_per_sm * 32);
  const int per_sm = (warps_per_sm >= 32 ? warps_per_sm / 32 : 1);
  const int blocks = nsm * per_sm;
  double* scratch = nullptr;
  cudaMalloc(&scratch, sizeof(double) * (size_t)blocks * threads);
  double ms;
  switch (chains) {
    case 1: ms = run<1>(blocks, threads, iters, scratch); break;
    case 2: ms = run<2>(blocks, threads, iters, scratch); break;
    case 4: ms = run<4>(blocks, threads, iters, scratch); break;
    default: chains = 8; ms = run<8>(blocks, threads, iters, scratch); break;
  }
  cudaFree(scratch);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  const double flops = 2.0 * (double)blocks * threads * chains * 8.0 * (double)iters;
  return flops / (ms * 1e-3) / 1e12;
}
