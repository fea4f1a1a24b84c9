// wirebench.cu -- what does a kernel that is bound by NVLink (stores into a peer GPU's memory) do to a kernel that runs
// beside it on the same GPU?  Background: the multi-GPU layout changes are fused into the last pass of the operators as
// peer stores and run at wire speed with the SMs mostly idle; running the NEXT chunk's arithmetic beside them was
// measured slower (profiles/ab_r02_vdfdx_pipelined_pass3_4gpu.txt).  This benchmark separates two explanations:
// the co-runner loses because it shares SMs (load/store units) with CTAs whose remote stores are backed up, or because
// the backed-up stores clog a chip-wide resource (crossbar / L2).  One process, two GPUs (cudaDeviceEnablePeerAccess):
//   wire kernel : persistent CTAs (4 per SM) copy local tiles into the peer's memory; CTAs on SMs >= sm_limit exit at once,
//                 so sm_limit confines it to the first SMs (with all their registers: nothing else fits there)
//   local kernel: grid-stride copy in local HBM (the co-runner)
// Measurement utility (tools/gpu_session.sh stage "wirebench", needs 2 GPUs); not part of the product library.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
  } while (0)

constexpr size_t BYTES = 1ull << 30;          // 1 GiB per buffer
constexpr int TILE = 32 * 1024;               // bytes per tile (one CTA iteration)
// dynamic shared memory nobody touches: 4 wire CTAs fill an SM (4 x 55 KB), so that no CTA of the co-runner (32 KB each,
// like passes 1 and 2 of v df/dx) fits on an SM that runs the wire kernel
constexpr int WIRE_SMEM = 55 * 1024, LOCAL_SMEM = 32 * 1024;

__global__ void __launch_bounds__(128, 4) wire_kernel(const double2* __restrict__ src, double2* __restrict__ peer,
                                                      unsigned* counter, unsigned ntiles, unsigned sm_limit) {
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (smid >= sm_limit) return;
  __shared__ unsigned s_tile;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    __syncthreads();
    if (tile >= ntiles) return;
    const size_t base = (size_t)tile * (TILE / 16);
    double2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = src[base + threadIdx.x + 128 * i];
#pragma unroll
    for (int i = 0; i < 16; ++i) peer[base + threadIdx.x + 128 * i] = v[i];
  }
}

__global__ void __launch_bounds__(256) local_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}

int main() {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("wirebench needs 2 GPUs\n"); return 0; }
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, 0, 1));
  if (!can) { printf("no peer access 0 -> 1\n"); return 0; }
  double2 *peer, *src, *dst, *src2;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&peer, BYTES));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&src, BYTES)); CK(cudaMalloc(&dst, BYTES)); CK(cudaMalloc(&src2, BYTES));
  CK(cudaMemset(src, 1, BYTES)); CK(cudaMemset(src2, 2, BYTES));
  unsigned* counter;
  CK(cudaMalloc(&counter, 4));
  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(wire_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WIRE_SMEM));
  cudaStream_t s1, s2;
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t a1, b1, a2, b2;
  CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b1)); CK(cudaEventCreate(&a2)); CK(cudaEventCreate(&b2));
  const unsigned ntiles = (unsigned)(BYTES / TILE);
  const size_t n16 = BYTES / 16;
  auto wire = [&](unsigned sm_limit, int ctas_per_sm = 4) {
    CK(cudaMemsetAsync(counter, 0, 4, s1));
    CK(cudaEventRecord(a1, s1));
    wire_kernel<<<nsm * ctas_per_sm, 128, WIRE_SMEM, s1>>>(src, peer, counter, ntiles, sm_limit);
    CK(cudaEventRecord(b1, s1));
  };
  auto local = [&](int reps) {
    CK(cudaEventRecord(a2, s2));
    for (int r = 0; r < reps; ++r) local_kernel<<<nsm * 16, 256, LOCAL_SMEM, s2>>>(src2, dst, n16);
    CK(cudaEventRecord(b2, s2));
  };
  float t1, t2;
  printf("1 GiB per transfer; %d SMs; wire kernel: 4 CTAs of 128 threads per SM on the first sm_limit SMs\n", nsm);
  for (int rep = 0; rep < 2; ++rep) {                                     // second round = warm
    local(2); CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&t2, a2, b2));
    if (rep) printf("local copy alone (2 x 1 GiB read + write)         : %.3f ms  (%.0f GB/s)\n", t2, 4.0 * BYTES / t2 / 1e6);
    const unsigned limits[5] = {(unsigned)nsm, 64, 48, 32, 16};
    for (unsigned lim : limits) {
      wire(lim); CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&t1, a1, b1));
      if (rep) printf("wire alone, sm_limit %3u                          : %.3f ms  (%.0f GB/s over NVLink)\n", lim, t1, BYTES / t1 / 1e6);
    }
    for (int cps = 1; cps <= 2; ++cps) {                                  // SHARED SMs: 1 or 2 wire CTAs on every SM
      wire((unsigned)nsm, cps); CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&t1, a1, b1));
      if (rep) printf("wire alone, %d CTA(s) on every SM                  : %.3f ms  (%.0f GB/s over NVLink)\n", cps, t1, BYTES / t1 / 1e6);
      wire((unsigned)nsm, cps); local(2); CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&t1, a1, b1)); CK(cudaEventElapsedTime(&t2, a2, b2));
      if (rep) printf("together, %d wire CTA(s) on every SM (shared SMs): wire %.3f ms (%.0f GB/s), local copy %.3f ms (%.0f GB/s)\n",
                      cps, t1, BYTES / t1 / 1e6, t2, 4.0 * BYTES / t2 / 1e6);
    }
    for (unsigned lim : limits) {
      if (lim == (unsigned)nsm) continue;                                 // (4 CTAs on every SM leave no room for the co-runner)
      wire(lim); local(2); CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&t1, a1, b1)); CK(cudaEventElapsedTime(&t2, a2, b2));
      if (rep) printf("together, sm_limit %3u: wire %.3f ms (%.0f GB/s), local copy %.3f ms (%.0f GB/s)\n", lim, t1,
                      BYTES / t1 / 1e6, t2, 4.0 * BYTES / t2 / 1e6);
    }
  }
  return 0;
}
