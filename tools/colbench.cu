// colbench.cu -- micro-benchmarks behind the design of the single-pass column kernel (v df/dx with a whole column
// tile on chip across a thread-block cluster): how fast can NARROW column tiles of a row-major fp64 matrix be read
// and written (32 / 64 / 128 bytes per row, rows 8 apart per CTA), with plain 16-byte loads and with TMA 3-D box
// loads + L2 promotion, and how fast is the cluster's distributed shared memory (remote 16-byte stores).
// Measurement utility (tools/gpu_session.sh stage "colbench"); not part of the product library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
  } while (0)

constexpr int NX = 16384, NV = 16384, C = 8, L = NX / C;   // CTA r of a tile touches rows r + 8 n', n' < 2048

// ---- plain loads / stores: SEG bytes per row, 256 threads, 16 bytes per thread and access
template <int SEG, int MODE>   // MODE 0 read, 1 write, 2 copy (read everything, then write)
__global__ void __launch_bounds__(256) narrow_kernel(const double* __restrict__ in, double* __restrict__ out, double* sink) {
  constexpr int LPR = SEG / 16;             // lanes per row
  constexpr int RPI = 256 / LPR;            // rows per instruction of the CTA
  constexpr int NI = L / RPI;               // instructions per thread
  const int tile = blockIdx.x / C, r = blockIdx.x % C;
  const int lane_col = threadIdx.x % LPR, row0 = threadIdx.x / LPR;
  const size_t col = (size_t)tile * (SEG / 8) + 2 * lane_col;
  double2 v[NI];
  double acc = 0.0;
  if (MODE != 1) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const size_t n = (size_t)r + (size_t)C * (row0 + (size_t)i * RPI);
      v[i] = *reinterpret_cast<const double2*>(in + n * NV + col);
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) acc += v[i].x + v[i].y;
  } else {
#pragma unroll
    for (int i = 0; i < NI; ++i) v[i] = make_double2((double)i, (double)threadIdx.x);
  }
  if (MODE != 0) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const size_t n = (size_t)r + (size_t)C * (row0 + (size_t)i * RPI);
      *reinterpret_cast<double2*>(out + n * NV + col) = v[i];
    }
  }
  if (acc == 123.456) sink[blockIdx.x] = acc;
}

// ---- TMA: 3-D view (col, n % 8, n / 8) of the matrix, boxes of (SEG/8 cols, 1, 256 rows), 8 boxes per CTA tile
template <int SEG>
__global__ void __launch_bounds__(256) tma_read_kernel(const __grid_constant__ CUtensorMap tmap, double* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  const int tile = blockIdx.x / C, r = blockIdx.x % C;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(L * SEG)) : "memory");
    for (int b = 0; b < L / 256; ++b) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(smem + (size_t)b * 256 * SEG);
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"(dst), "l"(&tmap), "r"(tile * (SEG / 8)), "r"(r), "r"(b * 256), "r"(bar) : "memory");
    }
  }
  asm volatile(
      "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(bar) : "memory");
  // touch the tile so the copy cannot be skipped
  const double* s = reinterpret_cast<const double*>(smem);
  double acc = 0.0;
  for (int i = threadIdx.x; i < L * SEG / 8; i += 256 * 16) acc += s[i];
  if (acc == 123.456) sink[blockIdx.x] = acc;
}

// ---- distributed shared memory: every thread stores 16-byte values into the other CTAs of its cluster
template <int CS>
__global__ void __launch_bounds__(256) dsmem_kernel(int rounds, long long* cycles, double* sink) {
  extern __shared__ __align__(16) unsigned char smem[];          // 64 KB receive buffer
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned me = cluster.block_rank();
  double2* mine = reinterpret_cast<double2*>(smem);
  cluster.sync();
  const long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
    // 4096 values per CTA and round (64 KB), value i goes to CTA (i / 512 ... ) i.e. 1/CS of the data per peer
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = k * 256 + threadIdx.x;
      const unsigned peer = (me + 1 + (i * CS) / 4096) % CS;       // block of 4096/CS consecutive values per peer
      double2* remote = cluster.map_shared_rank(mine, peer);
      remote[i] = make_double2((double)it, (double)i);
    }
    cluster.sync();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (mine[threadIdx.x].x == 123.456) sink[blockIdx.x] = mine[threadIdx.x].y;
}

template <class F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

template <int SEG>
static void run_narrow(const double* in, double* out, double* sink) {
  const int grid = (NV * 8 / SEG) * C;
  const double gb = (double)NX * NV * 8 / 1e9;
  float t0 = time_ms([&] { narrow_kernel<SEG, 0><<<grid, 256>>>(in, out, sink); });
  float t1 = time_ms([&] { narrow_kernel<SEG, 1><<<grid, 256>>>(in, out, sink); });
  float t2 = time_ms([&] { narrow_kernel<SEG, 2><<<grid, 256>>>(in, out, sink); });
  printf("plain  %3d B/row: read %.3f ms %7.0f GB/s | write %.3f ms %7.0f GB/s | read+write %.3f ms %7.0f GB/s\n", SEG, t0,
         gb / t0 * 1e3, t1, gb / t1 * 1e3, t2, 2 * gb / t2 * 1e3);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int SEG>
static void run_tma(EncodeFn enc, double* in, double* sink, CUtensorMapL2promotion promo, const char* pname) {
  CUtensorMap tmap;
  cuuint64_t dims[3] = {(cuuint64_t)NV, (cuuint64_t)C, (cuuint64_t)L};
  cuuint64_t strides[2] = {(cuuint64_t)NV * 8, (cuuint64_t)NV * 8 * C};       // bytes, dims 1 and 2
  cuuint32_t box[3] = {SEG / 8, 1, 256};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc); return; }
  const int grid = (NV * 8 / SEG) * C;
  const size_t smem = (size_t)L * SEG;
  CK(cudaFuncSetAttribute(tma_read_kernel<SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const double gb = (double)NX * NV * 8 / 1e9;
  float t = time_ms([&] { tma_read_kernel<SEG><<<grid, 256, smem>>>(tmap, sink); });
  printf("TMA    %3d B/row (L2 promotion %s): read %.3f ms %7.0f GB/s\n", SEG, pname, t, gb / t * 1e3);
}

template <int CS>
static void run_dsmem(long long* cycles, double* sink) {
  const int rounds = 64;
  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(dsmem_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  if (CS > 8) CK(cudaFuncSetAttribute(dsmem_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int per_sm = 1; per_sm <= 2; ++per_sm) {
    cudaLaunchConfig_t cfg{};
    int grid = (nsm * per_sm / CS) * CS;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 65536; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    cudaOccupancyMaxActiveClusters(&nclusters, dsmem_kernel<CS>, &cfg);
    float t = time_ms([&] { CK(cudaLaunchKernelEx(&cfg, dsmem_kernel<CS>, rounds, cycles, sink)); }, 3);
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (long long c : h) mean += (double)c;
    mean /= grid;
    const double bytes_cta = (double)rounds * 65536.0 * (CS - 1) / CS;   // remote part
    printf("DSMEM cluster %2d, %d CTA/SM (grid %d, max co-resident clusters %d): %.3f ms, %.1f remote B/clk per CTA, "
           "%.0f GB/s aggregate (remote bytes only)\n", CS, per_sm, grid, nclusters, t, bytes_cta / mean,
           bytes_cta * grid / (t * 1e-3) / 1e9);
  }
}

int main() {
  double *in, *out, *sink;
  long long* cycles;
  CK(cudaMalloc(&in, sizeof(double) * (size_t)NX * NV));
  CK(cudaMalloc(&out, sizeof(double) * (size_t)NX * NV));
  CK(cudaMalloc(&sink, sizeof(double) * 65536 * 8));
  CK(cudaMalloc(&cycles, sizeof(long long) * 4096));
  CK(cudaMemset(in, 0, sizeof(double) * (size_t)NX * NV));
  {
    float t = time_ms([&] { CK(cudaMemcpyAsync(out, in, sizeof(double) * (size_t)NX * NV, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy D2D 2.147 GB: %.3f ms %7.0f GB/s (read+write)\n", t, 2.0 * NX * NV * 8 / 1e9 / t * 1e3);
  }
  run_narrow<32>(in, out, sink);
  run_narrow<64>(in, out, sink);
  run_narrow<128>(in, out, sink);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn) {
    EncodeFn enc = (EncodeFn)fn;
    run_tma<32>(enc, in, sink, CU_TENSOR_MAP_L2_PROMOTION_NONE, "none");
    run_tma<32>(enc, in, sink, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "128B");
    run_tma<32>(enc, in, sink, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "256B");
    run_tma<64>(enc, in, sink, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "256B");
  }
  run_dsmem<4>(cycles, sink);
  run_dsmem<8>(cycles, sink);
  run_dsmem<16>(cycles, sink);
  return 0;
}
