// vpfp_cuda.cu -- sm_100a kernels, launchers and the C ABI of libvpfp_b200.so (include/vpfp_b200.h).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
// There is no CPU path in this file: every entry point enqueues CUDA kernels on the caller's
// stream and returns; failures are reported through the return code + vpfp_last_error().
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vpfp_b200.h"
#include "advect.h"
#include "advect_fast.cuh"
#include "fp_fast.cuh"
#include "midfft.cuh"
#include "tinyfft.cuh"
#include "fp_reg.cuh"
#include "rowfft.cuh"
#include "rowops.h"
#include "tridiag.h"
#include "spline.h"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(VPFP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
  } while (0)

static bool is_pow2(long n) { return n > 0 && (n & (n - 1)) == 0; }

// ------------------------------------------------------------------------------------------
// optional per-launch timing (vpfp_profile_*): CUDA events recorded on the launching stream around
// every kernel of this library, read back by bench.py for the roofline block.  Off by default.
// ------------------------------------------------------------------------------------------
struct ProfRec { const char* label; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::atomic<long> g_launches{0};   // kernels enqueued by this library (vpfp_launch_count)
struct ProfScope {
  cudaStream_t st; bool on; size_t idx;
  ProfScope(const char* label, cudaStream_t s, int nlaunch = 1) : st(s), on(g_prof_on), idx(0) {
    g_launches += nlaunch;
    if (!on) return;
    ProfRec r; r.label = label;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    idx = g_prof.size(); g_prof.push_back(r);
  }
  ~ProfScope() { if (on) cudaEventRecord(g_prof[idx].b, st); }
};

// ------------------------------------------------------------------------------------------
// generic phase-program driver: phases separated by CTA barriers (see vpfp_common.h)
// ------------------------------------------------------------------------------------------
template <class Prog>
__global__ void __launch_bounds__(1024) prog_kernel(const Prog prog, const int nph) {
  extern __shared__ __align__(16) unsigned char smem[];
  for (int ph = 0; ph < nph; ++ph) {
    prog.phase(ph, (long)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, smem);
    __syncthreads();
  }
}

// the same driver for one-phase programs launched with 128 threads: without the 1024-thread bound of prog_kernel the
// compiler may use up to 255 registers (x-mode sums: more rows in flight)
template <class Prog>
__global__ void __launch_bounds__(128) prog128_kernel(const Prog prog) {
  prog.phase(0, (long)blockIdx.x, (int)threadIdx.x, 128, nullptr);
}

template <class Prog>
static int launch_prog(const Prog& prog, long nblocks, int threads, long smem, int nph,
                       cudaStream_t st, const char* label = "generic") {
  if (nblocks <= 0) return VPFP_OK;
  if (nblocks > 2147483647L) return fail(VPFP_ERR_UNSUPPORTED, "grid too large");
  if (smem > 227 * 1024) return fail(VPFP_ERR_UNSUPPORTED, "tile does not fit shared memory");
  static std::map<int, long> configured;  // per device: largest smem opted in for this Prog
  if (smem > 48 * 1024) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (configured[dev] < smem) {
      CUDA_TRY(cudaFuncSetAttribute(prog_kernel<Prog>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
      configured[dev] = 227 * 1024;
    }
  }
  {
    ProfScope ps(label, st);
    prog_kernel<Prog><<<(unsigned)nblocks, threads, smem, st>>>(prog, nph);
  }
  CUDA_TRY(cudaGetLastError());
  return VPFP_OK;
}

// ------------------------------------------------------------------------------------------
// per-device caches: twiddle tables exp(-2 pi i m / N) and a grow-only scratch buffer
// ------------------------------------------------------------------------------------------
struct DeviceCache {
  std::map<int, cplx*> tw;
  void* scratch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // grow-only, one per purpose
  size_t scratch_bytes[5] = {0, 0, 0, 0, 0};
  std::map<int, double*> spline_tab;   // per n: elimination factors of the (1,4,1) system, then n ones, then n fours
  std::vector<void*> retired;   // outgrown scratch blocks: kept until vpfp_shutdown (a captured CUDA graph may hold them)
};
enum { SCR_XMODES = 0, SCR_PHANTOM = 1, SCR_DENSITY = 2, SCR_TRIDIAG = 3, SCR_SPLINE = 4 };
static std::map<int, DeviceCache> g_cache;
static std::mutex g_cache_mu;

static int get_twiddles(int N, const cplx** out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_cache_mu);
  DeviceCache& c = g_cache[dev];
  auto it = c.tw.find(N);
  if (it != c.tw.end()) {
    *out = it->second;
    return VPFP_OK;
  }
  std::vector<cplx> h((size_t)N);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int m = 0; m < N; ++m) {
    long double a = two_pi * (long double)m / (long double)N;
    h[m].x = (double)cosl(a);
    h[m].y = (double)(-sinl(a));
  }
  // exact values on the axes and diagonals
  h[0].x = 1.0; h[0].y = 0.0;
  if (N % 4 == 0) { h[N / 4].x = 0.0; h[N / 4].y = -1.0; h[3 * N / 4].x = 0.0; h[3 * N / 4].y = 1.0; }
  if (N % 2 == 0) { h[N / 2].x = -1.0; h[N / 2].y = 0.0; }
  cplx* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(cplx) * (size_t)N));
  CUDA_TRY(cudaMemcpy(d, h.data(), sizeof(cplx) * (size_t)N, cudaMemcpyHostToDevice));
  c.tw[N] = d;
  *out = d;
  return VPFP_OK;
}

// Reduction scratch: one grow-only block per device and purpose.  A block that has become too small is RETIRED, not
// freed (no cudaFree, hence no implicit device synchronisation and no dangling pointer): a CUDA graph captured
// earlier keeps writing to the block it was captured with, which nobody else uses any more.  Growing needs a
// cudaMalloc, which is illegal during stream capture: callers that capture a graph run the same calls once before
// (vlapy_b200/outer_loop.py _GraphStep.capture does).  The blocks are shared by all streams of a device: calls that
// use scratch (x-modes, fused density, odd row counts) must not run concurrently on two streams of one device.
static unsigned long g_scratch_generation = 0;
static int get_scratch(int slot, size_t bytes, void** out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_cache_mu);
  DeviceCache& c = g_cache[dev];
  if (c.scratch_bytes[slot] < bytes) {
    void* fresh = nullptr;
    CUDA_TRY(cudaMalloc(&fresh, bytes));
    if (c.scratch[slot]) c.retired.push_back(c.scratch[slot]);
    c.scratch[slot] = fresh;
    c.scratch_bytes[slot] = bytes;
    ++g_scratch_generation;
  }
  *out = c.scratch[slot];
  return VPFP_OK;
}

// tables of the spline operators (spline.h), cached per device and length n: [0, n) the elimination factors
// c'_k = 1 / (4 - c'_{k-1}) of the constant (1, 4, 1) system, [n, 2n) ones, [2n, 3n) fours (broadcast diagonals)
static int get_spline_tab(int n, const double** out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_cache_mu);
  DeviceCache& c = g_cache[dev];
  auto it = c.spline_tab.find(n);
  if (it != c.spline_tab.end()) { *out = it->second; return VPFP_OK; }
  std::vector<double> h(3 * (size_t)n);
  double cp = 0.0;
  for (int k = 0; k < n; ++k) { cp = 1.0 / (4.0 - cp); h[k] = cp; h[n + k] = 1.0; h[2 * (size_t)n + k] = 4.0; }
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(double) * h.size()));
  CUDA_TRY(cudaMemcpy(d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
  c.spline_tab[n] = d;
  *out = d;
  return VPFP_OK;
}

// ------------------------------------------------------------------------------------------
// advection / Poisson launcher
// ------------------------------------------------------------------------------------------
// ---- fast register-resident path (advect_fast.cuh): N1, N2 in {64, 128}
// dynamic shared memory above 48 KB is an opt-in per kernel AND per device (function attributes belong to
// the device's context): remembered per (kernel, device)
template <class Kern>
static int opt_in_smem(Kern kern, size_t smem) {
  if (smem <= 48 * 1024) return VPFP_OK;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  std::lock_guard<std::mutex> lk(mu);
  size_t& have = done[{(const void*)kern, dev}];
  if (have < smem) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
  return VPFP_OK;
}

template <int L, int MODE, int INV>
static int launch_pass13(const fast::FastArgs& fa, cudaStream_t st) {
  constexpr int CB = (L == 128) ? 16 : (L == 64 ? 32 : 64);
  constexpr int threads = CB * fast::Geo<L>::TPC;
  const size_t smem = sizeof(cplx) * (L * CB + L);     // exchange buffer + twiddle table
  long grid;
  if (MODE == ADV_COLS) grid = (long)fa.nsim * fa.N2 * ((fa.seq_cnt + CB - 1) / CB);
  else grid = ((long)fa.seq_cnt * fa.N2 + CB - 1) / CB;
  if (grid > 2147483647L) return fail(VPFP_ERR_UNSUPPORTED, "grid too large");
  {
    ProfScope ps(MODE == ADV_COLS ? (INV ? "vdfdx.pass3" : "vdfdx.pass1") : (INV ? "edfdv.pass3" : "edfdv.pass1"), st);
    fast::pass13_kernel<L, MODE, INV, CB><<<(unsigned)grid, threads, smem, st>>>(fa);
  }
  CUDA_TRY(cudaGetLastError());
  return VPFP_OK;
}

// pass 2 prefetches its next tile with cp.async into the exchange buffer (PFM 2, advect_fast.cuh); a CTA walks
// 8 consecutive group-pair tiles (measured: 4 and 16 are slower)
template <int L, int MODE, bool EX>
static int launch_pass2_pf(const fast::FastArgs& fa, cudaStream_t st) {
  constexpr int PFM = 2;
#ifndef PASS2_CB128
#define PASS2_CB128 8     // packed column pairs per CTA at L = 128 (measured: 16 = 256-byte rows, 256 threads, 2 CTAs/SM: no gain)
#endif
  constexpr int CB = (L == 128) ? PASS2_CB128 : (L == 64 ? 16 : 32);
  constexpr int threads = CB * 2 * fast::Geo<L>::TPC;
  const size_t smem = fast::pass2_smem<L, CB>(MODE, PFM);
  int rc = opt_in_smem(fast::pass2_kernel<L, MODE, CB, PFM, EX>, smem);
  if (rc) return rc;
  const int T1 = fa.N1 / 2;
  const int t1_chunk = 8;
  const int nchunks = (T1 + t1_chunk - 1) / t1_chunk;
  const long grid = (long)(MODE == ADV_COLS ? fa.nsim : 1) * ((fa.seq_cnt + CB - 1) / CB) * nchunks;
  if (grid > 2147483647L) return fail(VPFP_ERR_UNSUPPORTED, "grid too large");
  {
    ProfScope ps(MODE == ADV_COLS ? "vdfdx.pass2" : "edfdv.pass2", st);
    fast::pass2_kernel<L, MODE, CB, PFM, EX><<<(unsigned)grid, threads, smem, st>>>(fa, t1_chunk);
  }
  CUDA_TRY(cudaGetLastError());
  return VPFP_OK;
}

template <int L, int MODE>
static int launch_pass2(const fast::FastArgs& fa, cudaStream_t st) {
  return fa.exact ? launch_pass2_pf<L, MODE, true>(fa, st) : launch_pass2_pf<L, MODE, false>(fa, st);
}

template <int MODE>
static int run_three_passes(const fast::FastArgs& fa, cudaStream_t st) {
  int rc;
  switch (fa.N1) {
    case 16: rc = launch_pass13<16, MODE, 0>(fa, st); break;
    case 32: rc = launch_pass13<32, MODE, 0>(fa, st); break;
    case 64: rc = launch_pass13<64, MODE, 0>(fa, st); break;
    default: rc = launch_pass13<128, MODE, 0>(fa, st); break;
  }
  if (rc) return rc;
  switch (fa.N2) {
    case 16: rc = launch_pass2<16, MODE>(fa, st); break;
    case 32: rc = launch_pass2<32, MODE>(fa, st); break;
    case 64: rc = launch_pass2<64, MODE>(fa, st); break;
    default: rc = launch_pass2<128, MODE>(fa, st); break;
  }
  if (rc) return rc;
  switch (fa.N1) {
    case 16: rc = launch_pass13<16, MODE, 1>(fa, st); break;
    case 32: rc = launch_pass13<32, MODE, 1>(fa, st); break;
    case 64: rc = launch_pass13<64, MODE, 1>(fa, st); break;
    default: rc = launch_pass13<128, MODE, 1>(fa, st); break;
  }
  return rc;
}

template <int MODE>
static int run_fast_mode(fast::FastArgs fa, cudaStream_t st) {
  // (slab-by-slab execution of the three passes out of L2 was measured slower and is gone: DESIGN.md section 4)
  fa.seq_off = 0; fa.seq_cnt = fa.nseq;
  return run_three_passes<MODE>(fa, st);
}

// ---- mid-size single-pass kernels (midfft.cuh): N in {256, 512, 1024, 2048}, uniform wavenumber grids
template <class P>
static int launch_midfft(const midfft::Args& ma, cudaStream_t st, const char* label) {
  P prog;
  prog.a = ma;
  int rc = opt_in_smem(midfft::midfft_kernel<P>, (size_t)P::SMEM_BYTES);
  if (rc) return rc;
  // persistent: SMs x resident CTAs, every CTA walks tiles blockIdx.x, + grid, ... and prefetches its next one
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  static std::mutex mu;
  static std::map<int, int> resident;      // per device (and per instantiation P)
  int per_dev;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!resident.count(dev)) {
      int nsm = 0, occ = 0;
      CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, midfft::midfft_kernel<P>, P::NT, P::SMEM_BYTES));
      if (occ < 1) return fail(VPFP_ERR_CUDA, "midfft kernel does not fit an SM");
      resident[dev] = nsm * occ;
    }
    per_dev = resident[dev];
  }
  long grid = prog.ntiles();
  const long cap = P::PREFETCH ? per_dev : 16L * per_dev;    // without prefetch: many CTAs hide each other's loads
  if (grid > cap) grid = cap;
  {
    ProfScope ps(label, st);
    midfft::midfft_kernel<P><<<(unsigned)grid, P::NT, P::SMEM_BYTES, st>>>(prog);
  }
  CUDA_TRY(cudaGetLastError());
  return VPFP_OK;
}

static bool midfft_eligible(const AdvectProg& a, int flags) {
  if (a.op != OP_PHASE || !(flags & VPFP_PHASE_TABLE) || (flags & (VPFP_FORCE_GENERIC | VPFP_FORCE_THREE_PASS))) return false;
  return a.N == 256 || a.N == 512 || a.N == 1024 || a.N == 2048;
}

struct DensityReq;
static int midfft_density(midfft::Args& ma, const AdvectProg& a, const DensityReq* dens, int CB, cudaStream_t st, bool finish);

static int run_midfft(const AdvectProg& a, cudaStream_t st, const DensityReq* dens = nullptr, bool* dens_done = nullptr) {
  midfft::Args ma;
  ma.nsim = a.nsim; ma.nseq = a.nseq; ma.nrows = a.nrows;
  ma.fin = a.fin; ma.ld_in = a.ld_in; ma.fout = a.fout; ma.ld_out = a.ld_out;
  ma.kvec = a.kvec; ma.cvec = a.cvec; ma.dt = a.dt;
  ma.dens_partial = nullptr; ma.dv = 0.0; ma.edge_flags = 3;
  int rc = get_twiddles(a.N, &ma.tw);
  if (rc) return rc;
  if (a.mode == ADV_COLS && dens && a.N <= 1024) {
    // charge density fused into the store phase (one partial row per column tile) + a small reduce kernel
    const int CB = (a.N == 1024) ? 4 : 8;
    rc = midfft_density(ma, a, dens, CB, st, false);
    if (rc) return rc;
    switch (a.N) {
      case 256: rc = launch_midfft<midfft::Prog<256, 8, 4, ADV_COLS, 8, true>>(ma, st, "vdfdx.mid"); break;
      case 512: rc = launch_midfft<midfft::Prog<512, 8, 8, ADV_COLS, 8, true>>(ma, st, "vdfdx.mid"); break;
      default: rc = launch_midfft<midfft::Prog<1024, 16, 8, ADV_COLS, 4, true>>(ma, st, "vdfdx.mid"); break;
    }
    if (rc) return rc;
    rc = midfft_density(ma, a, dens, CB, st, true);
    if (rc) return rc;
    if (dens_done) *dens_done = true;
    return VPFP_OK;
  }
  if (a.mode == ADV_COLS) {
    switch (a.N) {
      case 256: return launch_midfft<midfft::Prog<256, 8, 4, ADV_COLS, 8>>(ma, st, "vdfdx.mid");
      case 512: return launch_midfft<midfft::Prog<512, 8, 8, ADV_COLS, 8>>(ma, st, "vdfdx.mid");
      case 1024: return launch_midfft<midfft::Prog<1024, 16, 8, ADV_COLS, 4>>(ma, st, "vdfdx.mid");
      default: return launch_midfft<midfft::Prog<2048, 16, 16, ADV_COLS, 4>>(ma, st, "vdfdx.mid");
    }
  }
  switch (a.N) {
#ifndef MIDFFT_RCB
#define MIDFFT_RCB 4      // row pairs per CTA at nv = 512 (measured: 8 -> 1.070 ms, 4 -> 1.005 ms, 2 -> 1.016 ms at C4)
#endif
    case 256: return launch_midfft<midfft::Prog<256, 8, 4, ADV_ROWS, 8>>(ma, st, "edfdv.mid");
    case 512: return launch_midfft<midfft::Prog<512, 8, 8, ADV_ROWS, MIDFFT_RCB>>(ma, st, "edfdv.mid");
    case 1024: return launch_midfft<midfft::Prog<1024, 16, 8, ADV_ROWS, 4>>(ma, st, "edfdv.mid");
    default: return launch_midfft<midfft::Prog<2048, 16, 16, ADV_ROWS, 2>>(ma, st, "edfdv.mid");
  }
}

static bool fast_eligible(const AdvectProg& a, const AdvectPlan& pl) {
  if (a.op != OP_PHASE || pl.N1 == 1) return false;
  auto ok = [](int n) { return n == 16 || n == 32 || n == 64 || n == 128; };
  return ok(pl.N1) && ok(pl.N2);
}

struct DensityReq {   // fused charge density request (vdfdx only)
  double* out = nullptr;
  double dv = 0.0;
  int edge_flags = 3;
};

// before the launch (finish = false): scratch for the partial rows; after it: the sum over the column tiles
static int midfft_density(midfft::Args& ma, const AdvectProg& a, const DensityReq* dens, int CB, cudaStream_t st, bool finish) {
  const int tiles = (a.nseq + CB - 1) / CB;
  const long n = (long)a.nsim * a.N;
  if (!finish) {
    void* scr = nullptr;
    int rc = get_scratch(SCR_DENSITY, sizeof(double) * (size_t)tiles * n, &scr);
    if (rc) return rc;
    ma.dens_partial = (double*)scr; ma.dv = dens->dv; ma.edge_flags = dens->edge_flags;
    return VPFP_OK;
  }
  {
    // many tiles, few cells (C2: 128 tiles, 256 cells): the serial sum of one thread per cell is a chain of dependent
    // loads (12.7 us); eight groups in a first launch, their sums in a second (as for the three passes)
    const int groups = (tiles >= 64 && tiles % 8 == 0) ? 8 : 1;
    ProfScope ps("vdfdx.density_reduce", st, groups > 1 ? 2 : 1);
    const unsigned gx = (unsigned)((n + 255) / 256);
    if (groups > 1) {
      fast::dens_reduce_kernel<<<dim3(gx, groups), 256, 0, st>>>(ma.dens_partial, tiles / groups, 1, n, nullptr, 1, a.N);
      fast::dens_reduce_kernel<<<gx, 256, 0, st>>>(ma.dens_partial, groups, tiles / groups, n, dens->out, 1, a.N);
    } else {
      fast::dens_reduce_kernel<<<gx, 256, 0, st>>>(ma.dens_partial, tiles, 1, n, dens->out, 1, a.N);
    }
  }
  CUDA_TRY(cudaGetLastError());
  return VPFP_OK;
}

struct ScatterReq {   // last pass stores into peer shards (multi-GPU), see advect_fast.cuh FastArgs
  int mode = 0, nparts = 1, my_rank = 0;
  double* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

static int run_advect(AdvectProg a, cudaStream_t st, int flags = VPFP_PHASE_EXACT,
                      const DensityReq* dens = nullptr, bool* dens_done = nullptr,
                      const ScatterReq* scat = nullptr) {
  // Small problems (< 4M cells) are latency bound: one generic kernel with the whole sequence in
  // shared memory beats three dependent launches.  Otherwise 256 <= N <= 16384 take the
  // register-resident three passes.
  const long cells = (long)a.N * 2 * a.nseq * (a.mode == ADV_COLS ? a.nsim : 1);
  const bool small = cells < (1L << 22) && a.N <= 2048;
  const AdvectPlan pl = ((flags & VPFP_FORCE_GENERIC) || small) ? make_advect_plan(a.mode, a.N, 2048, 2048)
                                                                : make_advect_plan(a.mode, a.N, 128, 128);
  if (!(scat && scat->mode) && a.op == OP_PHASE && a.mode == ADV_COLS && (a.N == 16 || a.N == 32) &&
      (flags & VPFP_PHASE_TABLE) && !(flags & (VPFP_FORCE_GENERIC | VPFP_FORCE_THREE_PASS)) && !(a.ld_in & 1) && !(a.ld_out & 1) &&
      !(((uintptr_t)a.fin | (uintptr_t)a.fout) & 15)) {
    // nx = 16 / 32 (the reference's Landau grid): the whole x transform in the registers of one thread (tinyfft.cuh);
    // density: the caller's fallback
    const long work = (long)a.nsim * a.nseq;
    const unsigned grid = (unsigned)((work + 127) / 128);
    ProfScope ps("vdfdx.tiny", st);
    if (a.N == 32) {
      tiny::ColsProg<32> p;
      p.nsim = a.nsim; p.nseq = a.nseq; p.fin = a.fin; p.ld_in = a.ld_in; p.fout = a.fout; p.ld_out = a.ld_out;
      p.kvec = a.kvec; p.cvec = a.cvec; p.dt = a.dt;
      prog128_kernel<tiny::ColsProg<32>><<<grid, 128, 0, st>>>(p);
    } else {
      tiny::ColsProg<16> p;
      p.nsim = a.nsim; p.nseq = a.nseq; p.fin = a.fin; p.ld_in = a.ld_in; p.fout = a.fout; p.ld_out = a.ld_out;
      p.kvec = a.kvec; p.cvec = a.cvec; p.dt = a.dt;
      prog128_kernel<tiny::ColsProg<16>><<<grid, 128, 0, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return VPFP_OK;
  }
  if (!(scat && scat->mode) && midfft_eligible(a, flags))     // density at N = 2048: the caller's fallback
    return run_midfft(a, st, (dens && dens->out) ? dens : nullptr, dens_done);
  int rc = get_twiddles(a.N, &a.tw);
  if (rc) return rc;
  if (scat && scat->mode && !(fast_eligible(a, pl) && !(flags & VPFP_FORCE_GENERIC)))
    return fail(VPFP_ERR_UNSUPPORTED, "peer scatter needs the register-resident kernels (256 <= N <= 16384, >= 4M cells)");
  if (pl.N1 == 1) {
    advect_set_pass(a, pl, 0);
    return launch_prog(a, a.ntiles(), pl.threads[0], a.smem_bytes(), a.nphases(), st,
                       a.op == OP_POISSON ? "poisson" : (a.mode == ADV_COLS ? "vdfdx.single" : "edfdv.single"));
  }
  if (a.mode == ADV_ROWS && (a.nrows & 1)) {
    void* ph = nullptr;
    rc = get_scratch(SCR_PHANTOM, sizeof(double) * (size_t)a.N, &ph);
    if (rc) return rc;
    a.phantom = (double*)ph;
  }
  if (fast_eligible(a, pl) && !(flags & VPFP_FORCE_GENERIC)) {
    fast::FastArgs fa;
    fa.mode = a.mode; fa.exact = (flags & VPFP_PHASE_TABLE) ? 0 : 1;
    fa.N = a.N; fa.N1 = pl.N1; fa.N2 = pl.N2; fa.lN2 = ilog2(pl.N2);
    fa.nsim = a.nsim; fa.nseq = a.nseq; fa.nrows = a.nrows;
    fa.seq_off = 0; fa.seq_cnt = a.nseq;
    fa.fin = a.fin; fa.ld_in = a.ld_in; fa.fout = a.fout; fa.ld_out = a.ld_out;
    fa.kvec = a.kvec; fa.cvec = a.cvec; fa.dt = a.dt; fa.phantom = a.phantom;
    fa.twN = a.tw;
    fa.dens_partial = nullptr; fa.dv = 0.0; fa.edge_flags = 3;
    fa.peer_mode = 0; fa.nparts = 1; fa.my_rank = 0; fa.lpart = 0;
    for (int i = 0; i < 8; ++i) fa.peer[i] = nullptr;
    if (scat && scat->mode) {
      fa.peer_mode = scat->mode; fa.nparts = scat->nparts; fa.my_rank = scat->my_rank;
      fa.lpart = ilog2(a.N / scat->nparts);
      for (int i = 0; i < scat->nparts; ++i) fa.peer[i] = scat->peer[i];
    }
    int dens_tiles = 0;
    if (dens && dens->out && a.mode == ADV_COLS) {
      const int CB = (pl.N1 == 128) ? 16 : (pl.N1 == 64 ? 32 : 64);
      dens_tiles = (a.nseq + CB - 1) / CB;
      void* scr = nullptr;
      rc = get_scratch(SCR_DENSITY, sizeof(double) * (size_t)dens_tiles * a.nsim * a.N, &scr);
      if (rc) return rc;
      fa.dens_partial = (double*)scr; fa.dv = dens->dv; fa.edge_flags = dens->edge_flags;
    }
    rc = get_twiddles(pl.N1, &fa.twL1);
    if (rc) return rc;
    rc = get_twiddles(pl.N2, &fa.twL2);
    if (rc) return rc;
    rc = (a.mode == ADV_COLS) ? run_fast_mode<ADV_COLS>(fa, st) : run_fast_mode<ADV_ROWS>(fa, st);
    if (rc) return rc;
    if (fa.dens_partial) {
      const long n = (long)a.nsim * a.N;
      {
        const int groups = (dens_tiles >= 64 && dens_tiles % 8 == 0) ? 8 : 1;
        ProfScope ps("vdfdx.density_reduce", st, groups > 1 ? 2 : 1);
        const unsigned gx = (unsigned)((n + 255) / 256);
        if (groups > 1) {
          fast::dens_reduce_kernel<<<dim3(gx, groups), 256, 0, st>>>(fa.dens_partial, dens_tiles / groups, 1, n, nullptr, fa.N1, fa.N2);
          fast::dens_reduce_kernel<<<gx, 256, 0, st>>>(fa.dens_partial, groups, dens_tiles / groups, n, dens->out, fa.N1, fa.N2);
        } else {
          fast::dens_reduce_kernel<<<gx, 256, 0, st>>>(fa.dens_partial, dens_tiles, 1, n, dens->out, fa.N1, fa.N2);
        }
      }
      CUDA_TRY(cudaGetLastError());
      if (dens_done) *dens_done = true;
    }
    return VPFP_OK;
  }
  for (int pass = 1; pass <= 3; ++pass) {
    advect_set_pass(a, pl, pass);
    rc = launch_prog(a, a.ntiles(), pl.threads[pass], a.smem_bytes(), a.nphases(), st,
                     a.op == OP_POISSON ? "poisson" : (a.mode == ADV_COLS ? "vdfdx.generic" : "edfdv.generic"));
    if (rc) return rc;
  }
  return VPFP_OK;
}

// Short rows (ncols <= 2048: ensembles of small grids, C1/C2/C4): ONE WARP per row instead of one CTA with a
// shared-memory tree -- every lane takes 16-byte pieces of the row, the eight sums are folded with shuffles in a
// fixed order (deterministic).  262144 rows of 512 cells (C4) took 3.7 ms with the CTA-per-row program (292 GB/s):
// the barriers of its tree dominate rows this short.  The f ln f sum uses the table logarithm of the Fokker-Planck
// kernels (fp_fast.cuh log_sum; the library log() made this kernel fp64-bound: 0.59 ms at C4).
template <int NMOM>
__global__ void __launch_bounds__(256) moments_warp_kernel(const MomentsProg p, const double2* __restrict__ logtab) {
  __shared__ double2 LT[128];
  if (NMOM > 7) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) LT[i] = logtab[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const double* src = p.f + row * p.ld;
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0;
  // (measured and rejected at C4, 0.51 ms for this version: unit weights with end-cell corrections after the
  // reduction and two accumulator sets, 0.57 ms; eight 16-byte pieces per lane loaded up front, 0.63 ms)
  auto add = [&](int j, double fv, double vv) {
    const double t = trapz_w(j, p.ncols, p.dv, p.edge_flags) * fv;
    acc[0] += t;
    if (NMOM > 1) {
      double q = t * vv;
      acc[1] += q;
      q *= vv; acc[2] += q;
      q *= vv; acc[3] += q;
      q *= vv; acc[4] += q;
      q *= vv; acc[5] += q;
    }
    if (NMOM > 6) {
      acc[6] = fma(t, fv, acc[6]);
      acc[7] = fma(t, fpfast::log_sum(fv, LT), acc[7]);
    }
  };
  const bool vec = ((p.ld & 1) == 0) && ((p.ncols & 1) == 0) &&
                   (((reinterpret_cast<uintptr_t>(p.f) | reinterpret_cast<uintptr_t>(p.v)) & 15) == 0);
  if (vec) {
    for (int j = 2 * lane; j < p.ncols; j += 64) {
      const double2 f2 = *reinterpret_cast<const double2*>(src + j);
      const double2 v2 = *reinterpret_cast<const double2*>(p.v + j);
      add(j, f2.x, v2.x);
      add(j + 1, f2.y, v2.y);
    }
  } else {
    for (int j = lane; j < p.ncols; j += 32) add(j, src[j], p.v[j]);
  }
#pragma unroll
  for (int k = 0; k < NMOM; ++k) {
    double y = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
    if (lane == 0 && k < p.nmom) p.out[(long)k * p.out_ld + row] = y;
  }
}

// Direct O(N^2) DFT Poisson solve for lengths that are not powers of two (the reference's own
// field-solver test uses nx = 96).  One CTA per density row.
struct PoissonDftProg {
  const double* n;
  const double* ook;
  const double* driver;
  double* e;
  int N;
  __host__ __device__ long smem_bytes() const { return (long)3 * N * sizeof(double); }
  __device__ void phase(int ph, long blk, int tid, int nthr, unsigned char* smem) const {
    double* rho = reinterpret_cast<double*>(smem);
    double* re = rho + N;
    double* im = re + N;
    const double w = -6.283185307179586476925286766559 / (double)N;
    if (ph == 0) {
      for (int x = tid; x < N; x += nthr) rho[x] = 1.0 - n[blk * N + x];
    } else if (ph == 1) {
      for (int k = tid; k < N; k += nthr) {
        double sr = 0.0, si = 0.0;
        for (int x = 0; x < N; ++x) {
          long kx = ((long)k * x) % N;
          double s, c;
          sincos(w * (double)kx, &s, &c);
          sr += rho[x] * c;
          si += rho[x] * s;
        }
        // multiply by i * one_over_kx[k]
        double o = ook[blk * N + k];
        re[k] = -o * si;
        im[k] = o * sr;
      }
    } else {
      for (int x = tid; x < N; x += nthr) {
        double acc = 0.0;
        for (int k = 0; k < N; ++k) {
          long kx = ((long)k * x) % N;
          double s, c;
          sincos(-w * (double)kx, &s, &c);
          acc += re[k] * c - im[k] * s;
        }
        double val = acc / (double)N;
        if (driver) val += driver[blk * N + x];
        e[blk * N + x] = val;
      }
    }
  }
};

// tables of the f ln f logarithm: n x (1/c_i, ln c_i), c_i = 1 + (i + 1/2)/n, cached per device
// (n = 128: fp_fast.cuh log_sum, n = 64: fp_reg.cuh log_split)
static std::map<std::pair<int, int>, double2*> g_logtab;
static int get_logtab(int n, const double2** out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_cache_mu);
  auto it = g_logtab.find({dev, n});
  if (it != g_logtab.end()) { *out = it->second; return VPFP_OK; }
  std::vector<double2> h(n);
  for (int i = 0; i < n; ++i) {
    const long double c = 1.0L + ((long double)i + 0.5L) / (long double)n;
    h[i].x = (double)(1.0L / c);
    h[i].y = (double)(-logl((long double)h[i].x));     // consistent with the rounded reciprocal
  }
  double2* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(double2) * n));
  CUDA_TRY(cudaMemcpy(d, h.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
  g_logtab[{dev, n}] = d;
  *out = d;
  return VPFP_OK;
}

template <int M, int T>
static int launch_fp_fast(const fpfast::Args& a, cudaStream_t st) {
  const size_t smem = fpfast::smem_bytes<M, T>();
  {
    int rc = opt_in_smem(fpfast::fp_kernel<M, T>, smem);
    if (rc) return rc;
  }
  int per_sm = (int)((227 * 1024) / smem);
  if (per_sm > 2048 / T) per_sm = 2048 / T;
  if (per_sm < 1) per_sm = 1;
  long grid = 148L * per_sm * 4;
  if (grid > a.rows) grid = a.rows;
  {
    ProfScope ps("fp_step", st);
    fpfast::fp_kernel<M, T><<<(unsigned)grid, T, smem, st>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  return VPFP_OK;
}

// register-resident Fokker-Planck kernel (fp_reg.cuh): nv = 32 T, next row prefetched with cp.async
static bool fp_reg_eligible(const fpfast::Args& a) {
  if (a.nv != 4096 && a.nv != 8192 && a.nv != 16384) return false;
  if ((a.ld_in & 1) || (a.ld_out & 1)) return false;                      // 16-byte row alignment
  return (((uintptr_t)a.fin | (uintptr_t)a.fout) & 15) == 0;
}

template <int M, int T, int OP>
static int launch_fp_reg_op(const fpfast::Args& a, cudaStream_t st) {
  const size_t smem = fpreg::Geo<M, T>::SMEM;
  {
    int rc = opt_in_smem(fpreg::fp_reg_kernel<M, T, OP>, smem);
    if (rc) return rc;
  }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 512 / T) per_sm = 512 / T;
  if (per_sm < 1) per_sm = 1;
  long grid = 148L * per_sm;                      // persistent: every CTA walks rows blockIdx.x, + grid, ...
  if (grid > a.rows) grid = a.rows;
  {
    ProfScope ps("fp_step", st);
    fpreg::fp_reg_kernel<M, T, OP><<<(unsigned)grid, T, smem, st>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  return VPFP_OK;
}

template <int M, int T>
static int launch_fp_reg(const fpfast::Args& a, cudaStream_t st) {
  return a.op == 0 ? launch_fp_reg_op<M, T, 0>(a, st) : launch_fp_reg_op<M, T, 1>(a, st);
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
// ---- single-pass row kernel (rowfft.cuh): e df/dv with one HBM read and one write of f
static bool rowfft_eligible(const double* f_in, long ld_in, const double* f_out, long ld_out, int rows, int nv,
                            int flags) {
  if (!(flags & VPFP_PHASE_TABLE) || (flags & (VPFP_FORCE_GENERIC | VPFP_FORCE_THREE_PASS))) return false;
  if (nv != 4096 && nv != 8192 && nv != 16384) return false;
  if ((long)rows * nv < (1L << 22)) return false;
  if ((ld_in & 1) || (ld_out & 1) || ((uintptr_t)f_in & 15) || ((uintptr_t)f_out & 15)) return false;
  return true;
}

template <class P>
static int launch_rowfft(const rowfft::Args& ra, cudaStream_t st, const char* label = "edfdv.row") {
  P prog;
  prog.a = ra;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  static std::map<int, int> grid_for;   // per device: SMs x resident CTAs
  {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!grid_for.count(dev)) {
      CUDA_TRY(cudaFuncSetAttribute(rowfft::rowfft_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)P::SMEM_BYTES));
      int nsm = 0, occ = 0;
      CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rowfft::rowfft_kernel<P>, P::T, P::SMEM_BYTES));
      if (occ < 1) return fail(VPFP_ERR_CUDA, "rowfft kernel does not fit an SM");
      grid_for[dev] = nsm * occ;
    }
  }
  int grid = grid_for[dev];
  if (grid > ra.nrows) grid = ra.nrows;
  {
    ProfScope ps(label, st);
    rowfft::rowfft_kernel<P><<<grid, P::T, P::SMEM_BYTES, st>>>(prog);
  }
  CUDA_TRY(cudaGetLastError());
  return VPFP_OK;
}

static int run_rowfft(const double* f_in, long ld_in, double* f_out, long ld_out, const double* e, const double* kv,
                      double dt, int rows, int nv, const ScatterReq* scat, cudaStream_t st, int flags = 0) {
  rowfft::Args ra;
  memset(&ra, 0, sizeof(ra));
  ra.fin = f_in; ra.ld_in = ld_in; ra.fout = f_out; ra.ld_out = ld_out; ra.kvec = kv; ra.cvec = e; ra.dt = dt;
  ra.nrows = rows;
  int rc = get_twiddles(nv, &ra.twN);
  if (rc) return rc;
  if (scat && scat->mode) {
    ra.peer_mode = scat->mode; ra.nparts = scat->nparts; ra.my_rank = scat->my_rank;
    ra.lpart = ilog2(nv / scat->nparts);
    for (int i = 0; i < scat->nparts; ++i) ra.peer[i] = scat->peer[i];
  }
  if (ra.peer_mode) {
    switch (nv) {
      case 16384: return launch_rowfft<rowfft::Prog<32, 16, true>>(ra, st);
      case 8192: return launch_rowfft<rowfft::Prog<16, 16, true>>(ra, st);
      default: return launch_rowfft<rowfft::Prog<8, 16, true>>(ra, st);
    }
  }
  switch (nv) {
    case 16384: return launch_rowfft<rowfft::Prog<32, 16>>(ra, st);
    case 8192: return launch_rowfft<rowfft::Prog<16, 16>>(ra, st);
    default: return launch_rowfft<rowfft::Prog<8, 16>>(ra, st);
  }
}

extern "C" {

int vpfp_moments(const double* f, long ld, const double* v, double dv, double* out, long out_ld,
                 int nmom, int rows, int ncols, int edge_flags, void* stream);

int vpfp_abi_version(void) { return VPFP_ABI_VERSION; }
const char* vpfp_last_error(void) { return g_err.c_str(); }

int vpfp_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  for (auto& kv : g_cache) {
    cudaSetDevice(kv.first);
    for (auto& t : kv.second.tw) cudaFree(t.second);
    for (int n : {128, 64})
      if (g_logtab.count({kv.first, n})) { cudaFree(g_logtab[{kv.first, n}]); g_logtab.erase({kv.first, n}); }
    for (int i = 0; i < 5; ++i)
      if (kv.second.scratch[i]) cudaFree(kv.second.scratch[i]);
    for (auto& t : kv.second.spline_tab) cudaFree(t.second);
    for (void* r : kv.second.retired) cudaFree(r);
  }
  g_cache.clear();
  return VPFP_OK;
}

long vpfp_launch_count(int reset) {
  const long n = g_launches.load();
  if (reset) g_launches = 0;
  return n;
}

unsigned long vpfp_scratch_generation(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  return g_scratch_generation;
}

int vpfp_profile_enable(int on) {
  g_prof_on = on != 0;
  return VPFP_OK;
}

// Synchronises the device, writes "label count total_ms\n" lines into buf and clears the records.
int vpfp_profile_report(char* buf, int buflen) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.label];
      e.first += 1; e.second += ms;
    }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  std::string out;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if ((int)out.size() + 1 > buflen) return fail(VPFP_ERR_ARG, "vpfp_profile_report: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return VPFP_OK;
}

int vpfp_edfdv_exp(const double* f_in, long ld_in, double* f_out, long ld_out, const double* e,
                   const double* kv, double dt, int rows, int nv, int flags, void* stream) {
  if (!f_in || !f_out || !e || !kv || rows <= 0 || nv <= 0 || ld_in < nv || ld_out < nv)
    return fail(VPFP_ERR_ARG, "vpfp_edfdv_exp: bad argument");
  if (!is_pow2(nv) || nv < 4 || nv > (1 << 24))
    return fail(VPFP_ERR_UNSUPPORTED, "e df/dv: <exponential> needs nv = 2^k >= 4 on the b200 backend");
  if (rowfft_eligible(f_in, ld_in, f_out, ld_out, rows, nv, flags))
    return run_rowfft(f_in, ld_in, f_out, ld_out, e, kv, dt, rows, nv, nullptr, (cudaStream_t)stream, flags);
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_ROWS; a.op = OP_PHASE; a.N = nv;
  a.nsim = 1; a.nrows = rows; a.nseq = (rows + 1) / 2;
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out;
  a.kvec = kv; a.cvec = e; a.addv = nullptr; a.dt = dt;
  return run_advect(a, (cudaStream_t)stream, flags);
}

int vpfp_vdfdx_exp(const double* f_in, long ld_in, double* f_out, long ld_out, const double* kx,
                   const double* v, double dt, int batch, int nx, int ncols, int flags,
                   void* stream) {
  if (!f_in || !f_out || !kx || !v || batch <= 0 || nx <= 0 || ncols <= 0 || ld_in < ncols ||
      ld_out < ncols)
    return fail(VPFP_ERR_ARG, "vpfp_vdfdx_exp: bad argument");
  if (!is_pow2(nx) || nx < 2 || nx > (1 << 24))
    return fail(VPFP_ERR_UNSUPPORTED, "v df/dx: <exponential> needs nx = 2^k >= 2 on the b200 backend");
  if ((ncols & 1) || (ld_in & 1) || (ld_out & 1) || ((uintptr_t)f_in & 15) || ((uintptr_t)f_out & 15))
    return fail(VPFP_ERR_UNSUPPORTED, "v df/dx: column count and row pitch must be even, f 16-byte aligned");
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_COLS; a.op = OP_PHASE; a.N = nx;
  a.nsim = batch; a.nrows = nx; a.nseq = ncols / 2;
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out;
  a.kvec = kx; a.cvec = v; a.addv = nullptr; a.dt = dt;
  return run_advect(a, (cudaStream_t)stream, flags);
}

int vpfp_vdfdx_exp_density(const double* f_in, long ld_in, double* f_out, long ld_out, const double* kx,
                           const double* v, double dt, int batch, int nx, int ncols, int flags,
                           double* n_out, double dv, int edge_flags, void* stream) {
  if (!f_in || !f_out || !kx || !v || !n_out || batch <= 0 || nx <= 0 || ncols <= 0 || ld_in < ncols ||
      ld_out < ncols)
    return fail(VPFP_ERR_ARG, "vpfp_vdfdx_exp_density: bad argument");
  if (!is_pow2(nx) || nx < 2 || nx > (1 << 24))
    return fail(VPFP_ERR_UNSUPPORTED, "v df/dx: <exponential> needs nx = 2^k >= 2 on the b200 backend");
  if ((ncols & 1) || (ld_in & 1) || (ld_out & 1) || ((uintptr_t)f_in & 15) || ((uintptr_t)f_out & 15))
    return fail(VPFP_ERR_UNSUPPORTED, "v df/dx: column count and row pitch must be even, f 16-byte aligned");
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_COLS; a.op = OP_PHASE; a.N = nx;
  a.nsim = batch; a.nrows = nx; a.nseq = ncols / 2;
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out;
  a.kvec = kx; a.cvec = v; a.addv = nullptr; a.dt = dt;
  DensityReq dr; dr.out = n_out; dr.dv = dv; dr.edge_flags = edge_flags;
  bool done = false;
  int rc = run_advect(a, (cudaStream_t)stream, flags, &dr, &done);
  if (rc) return rc;
  if (!done) {
    // sizes served by the generic kernels: density as a separate row reduction of the result
    return vpfp_moments(f_out, ld_out, v, dv, n_out, (long)batch * nx, 1, batch * nx, ncols, edge_flags, stream);
  }
  return VPFP_OK;
}

// ---- multi-GPU: peer-mapped shards and advection with the layout change fused into the last store
int vpfp_ipc_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return fail(VPFP_ERR_ARG, "vpfp_ipc_alloc: bad argument");
  CUDA_TRY(cudaMalloc(ptr, bytes));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, *ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return VPFP_OK;
}

int vpfp_ipc_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(VPFP_ERR_ARG, "vpfp_ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VPFP_OK;
}

int vpfp_ipc_close(void* ptr) {
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return VPFP_OK;
}

int vpfp_ipc_free(void* ptr) {
  CUDA_TRY(cudaFree(ptr));
  return VPFP_OK;
}

int vpfp_edfdv_exp_scatter(const double* f_in, long ld_in, double* scratch, long ld_scratch, const double* e,
                           const double* kv, double dt, int rows, int nv, int flags, void* const* peer_fv,
                           int nparts, int my_rank, void* stream) {
  if (!f_in || !scratch || !e || !kv || !peer_fv || rows <= 0 || nv <= 0 || nparts < 1 || nparts > 8 ||
      my_rank < 0 || my_rank >= nparts || nv % nparts)
    return fail(VPFP_ERR_ARG, "vpfp_edfdv_exp_scatter: bad argument");
  if (!is_pow2(nv) || !is_pow2(nparts))
    return fail(VPFP_ERR_UNSUPPORTED, "e df/dv scatter: nv and the number of ranks must be powers of two");
  if (rowfft_eligible(f_in, ld_in, scratch, ld_scratch, rows, nv, flags)) {
    ScatterReq sr1; sr1.mode = 1; sr1.nparts = nparts; sr1.my_rank = my_rank;
    for (int i = 0; i < nparts; ++i) sr1.peer[i] = (double*)peer_fv[i];
    return run_rowfft(f_in, ld_in, scratch, ld_scratch, e, kv, dt, rows, nv, &sr1, (cudaStream_t)stream);
  }
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_ROWS; a.op = OP_PHASE; a.N = nv;
  a.nsim = 1; a.nrows = rows; a.nseq = (rows + 1) / 2;
  a.fin = f_in; a.ld_in = ld_in; a.fout = scratch; a.ld_out = ld_scratch;
  a.kvec = kv; a.cvec = e; a.addv = nullptr; a.dt = dt;
  ScatterReq sr; sr.mode = 1; sr.nparts = nparts; sr.my_rank = my_rank;
  for (int i = 0; i < nparts; ++i) sr.peer[i] = (double*)peer_fv[i];
  return run_advect(a, (cudaStream_t)stream, flags, nullptr, nullptr, &sr);
}

int vpfp_vdfdx_exp_scatter(const double* f_in, long ld_in, double* scratch, long ld_scratch, const double* kx,
                           const double* v, double dt, int nx, int ncols, int flags, double* n_out, double dv,
                           int edge_flags, void* const* peer_fx, int nparts, int my_rank, void* stream) {
  if (!f_in || !scratch || !kx || !v || !peer_fx || nx <= 0 || ncols <= 0 || nparts < 1 || nparts > 8 ||
      my_rank < 0 || my_rank >= nparts || nx % nparts)
    return fail(VPFP_ERR_ARG, "vpfp_vdfdx_exp_scatter: bad argument");
  if (!is_pow2(nx) || !is_pow2(nparts) || (ncols & 1) || (ld_in & 1) || (ld_scratch & 1))
    return fail(VPFP_ERR_UNSUPPORTED, "v df/dx scatter: nx and ranks powers of two, even column counts");
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_COLS; a.op = OP_PHASE; a.N = nx;
  a.nsim = 1; a.nrows = nx; a.nseq = ncols / 2;
  a.fin = f_in; a.ld_in = ld_in; a.fout = scratch; a.ld_out = ld_scratch;
  a.kvec = kx; a.cvec = v; a.addv = nullptr; a.dt = dt;
  ScatterReq sr; sr.mode = 2; sr.nparts = nparts; sr.my_rank = my_rank;
  for (int i = 0; i < nparts; ++i) sr.peer[i] = (double*)peer_fx[i];
  DensityReq dr; dr.out = n_out; dr.dv = dv; dr.edge_flags = edge_flags;
  bool done = false;
  int rc = run_advect(a, (cudaStream_t)stream, flags, n_out ? &dr : nullptr, &done, &sr);
  if (rc) return rc;
  if (n_out && !done) return fail(VPFP_ERR_UNSUPPORTED, "v df/dx scatter: fused density unavailable at this size");
  return VPFP_OK;
}

int vpfp_edfdv_cd2(const double* f_in, long ld_in, double* f_out, long ld_out, const double* e,
                   double dt, double dv, int rows, int nv, void* stream) {
  if (!f_in || !f_out || !e || rows <= 0 || nv < 3) return fail(VPFP_ERR_ARG, "vpfp_edfdv_cd2: bad argument");
  Cd2Prog p;
  p.fin = f_in; p.ld_in = ld_in; p.fout = f_out; p.ld_out = ld_out; p.e = e;
  p.dt = dt; p.dv = dv; p.rows = rows; p.nv = nv;
  const int threads = 256;
  p.cblocks = (nv + threads * 4 - 1) / (threads * 4);
  return launch_prog(p, (long)rows * p.cblocks, threads, 0, 1, (cudaStream_t)stream, "edfdv.cd2");
}

int vpfp_moments(const double* f, long ld, const double* v, double dv, double* out, long out_ld,
                 int nmom, int rows, int ncols, int edge_flags, void* stream) {
  if (!f || !v || !out || rows <= 0 || ncols < 2 || nmom < 1 || nmom > 8)
    return fail(VPFP_ERR_ARG, "vpfp_moments: bad argument");
  MomentsProg p;
  p.f = f; p.ld = ld; p.v = v; p.dv = dv; p.out = out; p.out_ld = out_ld;
  p.nmom = nmom; p.rows = rows; p.ncols = ncols; p.edge_flags = edge_flags;
  if ((ncols <= 2048 && rows >= 1024) || ncols <= 512) {
    // many short rows, or very short rows (C1: 32 x 512): one warp per row (moments_warp_kernel)
    const unsigned grid = (unsigned)((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    const double2* lt = nullptr;
    if (nmom > 6) {
      int rc = get_logtab(128, &lt);
      if (rc) return rc;
    }
    {
      ProfScope ps("moments", st);
      if (nmom == 1) moments_warp_kernel<1><<<grid, 256, 0, st>>>(p, lt);
      else if (nmom <= 6) moments_warp_kernel<6><<<grid, 256, 0, st>>>(p, lt);
      else moments_warp_kernel<8><<<grid, 256, 0, st>>>(p, lt);
    }
    CUDA_TRY(cudaGetLastError());
    return VPFP_OK;
  }
  int threads = 256;
  while (threads > 32 && threads * 2 > ncols) threads >>= 1;
  return launch_prog(p, rows, threads, p.smem_bytes(threads), p.nphases(threads), (cudaStream_t)stream, "moments");
}

int vpfp_poisson(const double* n, const double* one_over_kx, const double* driver, double* e,
                 int batch, int nx, void* stream) {
  if (!n || !one_over_kx || !e || batch <= 0 || nx < 2) return fail(VPFP_ERR_ARG, "vpfp_poisson: bad argument");
  if (!is_pow2(nx)) {
    if (nx > 4096) return fail(VPFP_ERR_UNSUPPORTED, "spectral Poisson: non power-of-two nx > 4096");
    PoissonDftProg p;
    p.n = n; p.ook = one_over_kx; p.driver = driver; p.e = e; p.N = nx;
    return launch_prog(p, batch, 256, p.smem_bytes(), 3, (cudaStream_t)stream, "poisson");
  }
  if ((nx == 4096 || nx == 8192 || nx == 16384) && !(((uintptr_t)n | (uintptr_t)e | (uintptr_t)(driver ? driver : n)) & 15)) {
    // one CTA per density row, one launch (rowfft.cuh in Poisson mode); the generic path below takes three launches
    rowfft::Args ra;
    memset(&ra, 0, sizeof(ra));
    ra.fin = n; ra.ld_in = nx; ra.fout = e; ra.ld_out = nx; ra.kvec = one_over_kx; ra.cvec = nullptr; ra.dt = 0.0;
    ra.nrows = batch; ra.addv = driver;
    int rc = get_twiddles(nx, &ra.twN);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (nx) {
      case 16384: return launch_rowfft<rowfft::Prog<32, 16, false, true>>(ra, st, "poisson");
      case 8192: return launch_rowfft<rowfft::Prog<16, 16, false, true>>(ra, st, "poisson");
      default: return launch_rowfft<rowfft::Prog<8, 16, false, true>>(ra, st, "poisson");
    }
  }
  if (nx == 16 || nx == 32) {
    // the reference's Landau grid: two density rows in the registers of one thread (tinyfft.cuh)
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)(((batch + 1) / 2 + 127) / 128);
    ProfScope ps("poisson", st);
    if (nx == 32) {
      tiny::PoissonProg<32> p;
      p.nrows = batch; p.n = n; p.ook = one_over_kx; p.driver = driver; p.e = e;
      prog128_kernel<tiny::PoissonProg<32>><<<grid, 128, 0, st>>>(p);
    } else {
      tiny::PoissonProg<16> p;
      p.nrows = batch; p.n = n; p.ook = one_over_kx; p.driver = driver; p.e = e;
      prog128_kernel<tiny::PoissonProg<16>><<<grid, 128, 0, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return VPFP_OK;
  }
  if (nx == 256 || nx == 512 || nx == 1024 || nx == 2048) {
    // mid-size single-pass kernel in Poisson mode (midfft.cuh): two density rows per sequence, one launch
    midfft::Args ma;
    memset(&ma, 0, sizeof(ma));
    ma.nsim = 1; ma.nrows = batch; ma.nseq = (batch + 1) / 2;
    ma.fin = n; ma.ld_in = nx; ma.fout = e; ma.ld_out = nx; ma.kvec = one_over_kx; ma.addv = driver;
    int rc = get_twiddles(nx, &ma.tw);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (nx) {
      case 256: return launch_midfft<midfft::Prog<256, 8, 4, ADV_ROWS, 8, false, true>>(ma, st, "poisson");
      case 512: return launch_midfft<midfft::Prog<512, 8, 8, ADV_ROWS, 4, false, true>>(ma, st, "poisson");
      case 1024: return launch_midfft<midfft::Prog<1024, 16, 8, ADV_ROWS, 4, false, true>>(ma, st, "poisson");
      default: return launch_midfft<midfft::Prog<2048, 16, 16, ADV_ROWS, 2, false, true>>(ma, st, "poisson");
    }
  }
  AdvectProg a;
  memset(&a, 0, sizeof(a));
  a.mode = ADV_ROWS; a.op = OP_POISSON; a.N = nx;
  a.nsim = 1; a.nrows = batch; a.nseq = (batch + 1) / 2;
  a.fin = n; a.ld_in = nx; a.fout = e; a.ld_out = nx;
  a.kvec = one_over_kx; a.cvec = nullptr; a.addv = driver; a.dt = 0.0;
  return run_advect(a, (cudaStream_t)stream);
}

int vpfp_fp_step(const double* f_in, long ld_in, double* f_out, long ld_out, const double* v,
                 double nu, double dt, double dv, int op, double* moments_out, long mom_ld,
                 int rows, int nv, void* stream) {
  if (!f_in || !f_out || !v || rows <= 0 || nv <= 0) return fail(VPFP_ERR_ARG, "vpfp_fp_step: bad argument");
  if (op != VPFP_FP_LB && op != VPFP_FP_DG)
    return fail(VPFP_ERR_UNSUPPORTED, "Collision Operator: unknown operator id");
  if (nv < 8 || nv > 16384)
    return fail(VPFP_ERR_UNSUPPORTED, "Fokker-Planck step needs 8 <= nv <= 16384 on the b200 backend");
  FpProg p;
  p.fin = f_in; p.ld_in = ld_in; p.fout = f_out; p.ld_out = ld_out; p.v = v;
  p.nu = nu; p.dt = dt; p.dv = dv; p.op = op; p.mom_out = moments_out; p.mom_ld = mom_ld;
  p.rows = rows; p.nv = nv;
  int m = nv / 256;
  if (m < 4) m = 4;
  if (m > 16) m = 16;
  p.m = m;
  p.P = nv / m;
  int threads = ((p.P + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  return launch_prog(p, rows, threads, p.smem_bytes(threads), p.nphases(threads), (cudaStream_t)stream);
}

int vpfp_fp_step_linspace(const double* f_in, long ld_in, double* f_out, long ld_out, double v0,
                          double vstep, double vlast, double nu, double dt, double dv, int op,
                          double* moments_out, long mom_ld, int rows, int nv, void* stream) {
  if (!f_in || !f_out || rows <= 0 || nv <= 0) return fail(VPFP_ERR_ARG, "vpfp_fp_step_linspace: bad argument");
  if (op != VPFP_FP_LB && op != VPFP_FP_DG)
    return fail(VPFP_ERR_UNSUPPORTED, "Collision Operator: unknown operator id");
  fpfast::Args a;
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out;
  a.v0 = v0; a.vstep = vstep; a.vlast = vlast; a.nu = nu; a.dt = dt; a.dv = dv; a.op = op;
  a.mom_out = moments_out; a.mom_ld = mom_ld; a.rows = rows; a.nv = nv;
  int rc = get_logtab(128, &a.logtab);
  if (rc) return rc;
  rc = get_logtab(64, &a.logtab64);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (fp_reg_eligible(a)) {
    if (nv == 16384) return launch_fp_reg<32, 512>(a, st);
    if (nv == 8192) return launch_fp_reg<32, 256>(a, st);
    return launch_fp_reg<32, 128>(a, st);
  }
  switch (nv) {
    case 16384: return launch_fp_fast<32, 512>(a, st);
    case 8192: return launch_fp_fast<16, 512>(a, st);
    case 4096: return launch_fp_fast<16, 256>(a, st);
    case 2048: return launch_fp_fast<8, 256>(a, st);
    case 1024: return launch_fp_fast<8, 128>(a, st);
    case 512: return launch_fp_fast<4, 128>(a, st);
    case 256: return launch_fp_fast<4, 64>(a, st);
    case 128: return launch_fp_fast<4, 32>(a, st);
    default: return fail(VPFP_ERR_UNSUPPORTED, "vpfp_fp_step_linspace: nv must be a power of two in [128, 16384]");
  }
}

int vpfp_fp_diagonals(const double* f, long ld, const double* v, double nu, double dt, double dv, int op,
                      double* a, long lda, double* b, long ldb, double* c, long ldc, int rows, int nv, void* stream) {
  if (!f || !v || !a || !b || !c || rows <= 0 || nv < 2 || ld < nv || lda < nv - 1 || ldb < nv || ldc < nv - 1)
    return fail(VPFP_ERR_ARG, "vpfp_fp_diagonals: bad argument");
  if (op != VPFP_FP_LB && op != VPFP_FP_DG)
    return fail(VPFP_ERR_UNSUPPORTED, "Collision Operator: unknown operator id");
  DiagProg p;
  p.f = f; p.ld = ld; p.v = v; p.nu = nu; p.dt = dt; p.dv = dv; p.op = op;
  p.a = a; p.lda = lda; p.b = b; p.ldb = ldb; p.c = c; p.ldc = ldc; p.rows = rows; p.nv = nv;
  int threads = 256;
  while (threads > 32 && threads * 2 > nv) threads >>= 1;
  return launch_prog(p, rows, threads, p.smem_bytes(threads), p.nphases(threads), (cudaStream_t)stream, "fp_diagonals");
}

int vpfp_tridiag_solve(const double* a, long lda, const double* b, long ldb, const double* c, long ldc,
                       const double* d, long ldd, double* x, long ldx, int rows, int nv, void* stream) {
  if (!a || !b || !c || !d || !x || rows <= 0 || nv <= 0 || (lda != 0 && lda < nv - 1) || (ldb != 0 && ldb < nv) ||
      (ldc != 0 && ldc < nv - 1) || ldd < nv || ldx < nv)
    return fail(VPFP_ERR_ARG, "vpfp_tridiag_solve: bad argument");
  if (nv < 8 || nv > 16384)
    return fail(VPFP_ERR_UNSUPPORTED, "batched tridiagonal solver needs 8 <= nv <= 16384 on the b200 backend");
  TridiagProg p;
  p.a = a; p.lda = lda; p.b = b; p.ldb = ldb; p.c = c; p.ldc = ldc; p.d = d; p.ldd = ldd; p.x = x; p.ldx = ldx;
  p.rows = rows; p.nv = nv;
  int m = nv / 256;
  if (m < 4) m = 4;
  if (m > 16) m = 16;
  p.m = m;
  p.P = nv / m;
  int threads = ((p.P + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  return launch_prog(p, rows, threads, p.smem_bytes(threads), p.nphases(), (cudaStream_t)stream, "tridiag_solve");
}

// ---- semi-Lagrangian operators (spline.h)
int vpfp_vdfdx_sl(const double* f_in, long ld_in, double* f_out, long ld_out, const double* x, const double* v,
                  double dt, double dx, int nx, int nv, void* stream) {
  if (!f_in || !f_out || !x || !v || nx < 4 || nv < 1 || ld_in < nv || ld_out < nv || !(dx > 0.0))
    return fail(VPFP_ERR_ARG, "vpfp_vdfdx_sl: bad argument (nx >= 4)");
  cudaStream_t st = (cudaStream_t)stream;
  void* scr = nullptr;
  int rc = get_scratch(SCR_SPLINE, sizeof(double) * (size_t)(nx + 2) * nv, &scr);
  if (rc) return rc;
  const double* tab = nullptr;
  rc = get_spline_tab(nx, &tab);
  if (rc) return rc;
  SplineColSweepProg sw;
  sw.f = f_in; sw.ld = ld_in; sw.M = (double*)scr; sw.ldm = nv; sw.cp = tab; sw.h = dx; sw.nx = nx; sw.nv = nv;
  rc = launch_prog(sw, (nv + 63) / 64, 64, 0, 1, st, "vdfdx.sl");
  if (rc) return rc;
  SplineEvalProg<1> ev;
  ev.f = f_in; ev.ld = ld_in; ev.M = (const double*)scr; ev.ldm = nv; ev.out = f_out; ev.ld_out = ld_out;
  ev.ax = x; ev.c = v; ev.dt = dt; ev.nx = nx; ev.nv = nv; ev.cblocks = (nv + 255) / 256;
  return launch_prog(ev, (long)nx * ev.cblocks, 256, 0, 1, st, "vdfdx.sl");
}

int vpfp_tridiag_solve(const double* a, long lda, const double* b, long ldb, const double* c, long ldc,
                       const double* d, long ldd, double* x, long ldx, int rows, int nv, void* stream);

int vpfp_edfdv_sl(const double* f_in, long ld_in, double* f_out, long ld_out, const double* e, const double* v,
                  double dt, double dv, int nx, int nv, void* stream) {
  if (!f_in || !f_out || !e || !v || nx < 1 || nv < 4 || ld_in < nv || ld_out < nv || !(dv > 0.0))
    return fail(VPFP_ERR_ARG, "vpfp_edfdv_sl: bad argument (nv >= 4)");
  if (nv - 2 < 8 || nv - 2 > 16384)
    return fail(VPFP_ERR_UNSUPPORTED, "e df/dv: <sl> needs 10 <= nv <= 16386 on the b200 backend");
  cudaStream_t st = (cudaStream_t)stream;
  void* scr = nullptr;
  int rc = get_scratch(SCR_SPLINE, sizeof(double) * (size_t)nx * (nv + 2), &scr);
  if (rc) return rc;
  const double* tab = nullptr;
  rc = get_spline_tab(nv, &tab);
  if (rc) return rc;
  double* M = (double*)scr;
  const long ldm = nv + 2;
  SplineRowRhsProg rh;
  rh.f = f_in; rh.ld = ld_in; rh.M = M; rh.ldm = ldm; rh.h = dv; rh.nx = nx; rh.nv = nv; rh.cblocks = (nv + 255) / 256;
  rc = launch_prog(rh, (long)nx * rh.cblocks, 256, 0, 1, st, "edfdv.sl");
  if (rc) return rc;
  // (1, 4, 1) systems of all rows in place: unknowns M_2 .. M_{nv-1}; the diagonals are broadcast (pitch 0)
  rc = vpfp_tridiag_solve(tab + nv, 0, tab + 2 * (long)nv, 0, tab + nv, 0, M + 2, ldm, M + 2, ldm, nx, nv - 2, stream);
  if (rc) return rc;
  SplineEvalProg<0> ev;
  ev.f = f_in; ev.ld = ld_in; ev.M = M; ev.ldm = ldm; ev.out = f_out; ev.ld_out = ld_out;
  ev.ax = v; ev.c = e; ev.dt = dt; ev.nx = nx; ev.nv = nv; ev.cblocks = (nv + 255) / 256;
  return launch_prog(ev, (long)nx * ev.cblocks, 256, 0, 1, st, "edfdv.sl");
}

int vpfp_xmodes_partial(const double* f, long ld, double* out, int nmodes, int batch, int nx, int ncols,
                        int x_offset, int nx_total, void* stream);

int vpfp_xmodes(const double* f, long ld, double* out, int nmodes, int batch, int nx, int ncols,
                void* stream) {
  return vpfp_xmodes_partial(f, ld, out, nmodes, batch, nx, ncols, 0, nx, stream);
}

int vpfp_xmodes_partial(const double* f, long ld, double* out, int nmodes, int batch, int nx, int ncols,
                        int x_offset, int nx_total, void* stream) {
  if (!f || !out || nmodes < 1 || nmodes > 5 || batch <= 0 || nx <= 0 || ncols <= 0 || nx_total < nx || x_offset < 0)
    return fail(VPFP_ERR_ARG, "vpfp_xmodes: bad argument (1 <= nmodes <= 5)");
  XmodesProg p;
  p.f = f; p.ld = ld; p.nmodes = nmodes; p.batch = batch; p.nx = nx; p.ncols = ncols;
  p.x_offset = x_offset; p.nx_total = nx_total;
  // measured configuration: a CTA of 128 threads reads 128 x 16 contiguous bytes of every row; chunks of 128 rows, at
  // most 32 row chunks.  Small grids (C1, C2: a handful of CTAs, each walking its rows one group of eight after the
  // other -- 24 us at 256 x 2048) get shorter chunks, down to 8 rows, until there is a CTA per SM.
  const int threads = 128, env_xch = 32;
  p.cblocks = (ncols + threads - 1) / threads;
  int rows_per_chunk = 128;
  {
    const long col_ctas = (long)batch * ((ncols / 2 + threads - 1) / threads);
    while (rows_per_chunk > 8 && col_ctas * ((nx + rows_per_chunk - 1) / rows_per_chunk) < 148) rows_per_chunk >>= 1;
  }
  int xch = nx / rows_per_chunk;
  if (xch < 1) xch = 1;
  if (xch > env_xch) xch = env_xch;
  // (measured and rejected at 4096 x 4096: 128 row chunks instead of 32 -- the first stage gains 8 us, the reduction
  // of four times as many partials loses 22 us)
  p.xchunks = xch;
  void* scratch = nullptr;
  size_t bytes = (size_t)batch * xch * nmodes * ncols * 2 * sizeof(double);
  int rc = get_scratch(SCR_XMODES, bytes, &scratch);
  if (rc) return rc;
  p.partial = (double*)scratch;
  if (nmodes == 2 && (ncols & 1) == 0 && (ld & 1) == 0 && ((uintptr_t)f & 15) == 0) {
    Xmodes2Prog p2;                       // two columns per thread, eight rows in flight
    p2.f = f; p2.ld = ld; p2.partial = p.partial; p2.batch = batch; p2.nx = nx; p2.ncols = ncols;
    p2.xchunks = xch; p2.cblocks = (ncols / 2 + threads - 1) / threads;
    p2.x_offset = x_offset; p2.nx_total = nx_total;
    {
      ProfScope ps("xmodes", (cudaStream_t)stream);
      prog128_kernel<Xmodes2Prog><<<(unsigned)((long)batch * xch * p2.cblocks), 128, 0, (cudaStream_t)stream>>>(p2);
    }
    CUDA_TRY(cudaGetLastError());
    rc = VPFP_OK;
  } else {
    rc = launch_prog(p, (long)batch * xch * p.cblocks, threads, 0, 1, (cudaStream_t)stream, "xmodes");
  }
  if (rc) return rc;
  XmodesReduceProg r;
  r.partial = p.partial; r.out = out; r.nmodes = nmodes; r.batch = batch; r.ncols = ncols; r.xchunks = xch;
  long total = (long)batch * nmodes * ncols * 2;
  return launch_prog(r, (total + 255) / 256, 256, 0, 1, (cudaStream_t)stream, "xmodes.reduce");
}

int vpfp_driver(const double* x, double t, const double* pulses, int npulse, double* out, int nx,
                void* stream) {
  if (!x || !out || nx <= 0 || npulse < 0 || (npulse > 0 && !pulses))
    return fail(VPFP_ERR_ARG, "vpfp_driver: bad argument");
  if (npulse > DRIVER_MAX_PULSES) return fail(VPFP_ERR_UNSUPPORTED, "vpfp_driver: too many pulses");
  DriverProg p;
  p.x = x; p.out = out; p.t = t; p.t_dev = nullptr; p.ninc = 0; p.nx = nx; p.npulse = npulse;
  for (int i = 0; i < npulse * 7; ++i) p.pulses[i] = pulses[i];
  return launch_prog(p, (nx + 255) / 256, 256, 0, 1, (cudaStream_t)stream, "driver");
}

int vpfp_driver_dev(const double* x, const double* t_dev, const double* incs, int ninc, const double* pulses,
                    int npulse, double* out, int nx, void* stream) {
  if (!x || !out || !t_dev || nx <= 0 || npulse < 0 || ninc < 0 || ninc > 6 || (npulse > 0 && !pulses))
    return fail(VPFP_ERR_ARG, "vpfp_driver_dev: bad argument");
  if (npulse > DRIVER_MAX_PULSES) return fail(VPFP_ERR_UNSUPPORTED, "vpfp_driver: too many pulses");
  DriverProg p;
  p.x = x; p.out = out; p.t = 0.0; p.t_dev = t_dev; p.ninc = ninc; p.nx = nx; p.npulse = npulse;
  for (int i = 0; i < ninc; ++i) p.inc[i] = incs[i];
  for (int i = 0; i < npulse * 7; ++i) p.pulses[i] = pulses[i];
  return launch_prog(p, (nx + 255) / 256, 256, 0, 1, (cudaStream_t)stream, "driver");
}

int vpfp_series_batch(const double* moments, long mom_ld, const double* e, const double* de, double* out,
                      int nx, int batch, void* stream) {
  if (!moments || !e || !out || nx <= 0 || batch <= 0) return fail(VPFP_ERR_ARG, "vpfp_series: bad argument");
  SeriesProg p;
  p.mom = moments; p.mom_ld = mom_ld; p.e = e; p.de = de; p.out = out; p.nx = nx;
  int threads = (nx >= 4096) ? 1024 : 256;
  while (threads > 32 && threads > nx) threads >>= 1;
  return launch_prog(p, batch, threads, p.smem_bytes(threads), p.nphases(threads), (cudaStream_t)stream, "series");
}

int vpfp_series(const double* moments, long mom_ld, const double* e, const double* de, double* out,
                int nx, void* stream) {
  return vpfp_series_batch(moments, mom_ld, e, de, out, nx, 1, stream);
}

int vpfp_driver_batch(const double* x, double t, const double* t_dev, const double* incs, int ninc,
                      const double* pulses_dev, int npulse, double* out, int nx, int batch, void* stream) {
  if (!x || !out || !pulses_dev || nx <= 0 || batch <= 0 || npulse < 1 || ninc < 0 || ninc > 6)
    return fail(VPFP_ERR_ARG, "vpfp_driver_batch: bad argument");
  DriverBatchProg p;
  p.x = x; p.pulses = pulses_dev; p.out = out; p.t = t; p.t_dev = t_dev; p.ninc = t_dev ? ninc : 0;
  for (int i = 0; i < p.ninc; ++i) p.inc[i] = incs[i];
  p.nx = nx; p.npulse = npulse; p.batch = batch;
  const long n = (long)nx * batch;
  return launch_prog(p, (n + 255) / 256, 256, 0, 1, (cudaStream_t)stream, "driver");
}

}  // extern "C"
