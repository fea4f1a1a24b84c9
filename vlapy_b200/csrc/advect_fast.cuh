// advect_fast.cuh -- register-resident versions of the three advection passes (see advect.h for
// the algorithm and the meaning of N = N1*N2, COLS/ROWS packing and the pointwise un-mixing).
//
// Differences from the generic phase program in advect.h:
//  * every thread owns 16 complex values in registers; a length-L transform (L = 64 or 128) is
//    two register butterflies (radix L/8, then radix 8) with ONE shared-memory exchange between;
//  * the inverse runs the same two steps backwards (radix 8, exchange, radix L/8), so between the
//    forward and the inverse transform of pass 2 the data never leaves registers;
//  * in pass 2 a thread owns sub-transform m of one group AND sub-transform R1-1-m of the partner
//    group (or R1-m inside a self-paired group): both members of every conjugate pair (k, N-k)
//    sit in the same thread, so un-mixing the two packed real channels costs two complex
//    multiplies per PAIR:  U = Z + conj(Zp), V = Z - conj(Zp), X1 = Pa U, X2 = Pb V,
//    Y[k] = (X1 + X2)/2, Y[N-k] = conj(X1 - X2)/2;
//  * phase factors come from geometric tables (VPFP_PHASE_TABLE): P(k) = base(k1) * G^(k2) with
//    G = exp(-i N1 phi), phi = (K[1] dt) c, built once per column tile from 17 sincos per
//    sequence and reused for all group-pair tiles the CTA walks through; VPFP_PHASE_EXACT keeps the
//    per-bin sincos of theta = (K[k] dt) c (the reference's own rounding);
//  * intermediates are stored in NATURAL k1 order (row k1*N2 + n2): every thread knows the bins
//    it holds, so no bit reversal is needed anywhere.
#pragma once
#include "advect.h"
#include "butterflies.h"

namespace fast {

// ------------------------------------------------------------------------------------------
// geometry of a length-L transform held as 16 values per thread
// ------------------------------------------------------------------------------------------
template <int L>
struct Geo {
  static constexpr int R1 = L / 8;    // first radix (8 or 16)
  static constexpr int TPC = L / 16;  // threads per sequence
  static constexpr int NA = 16 / R1;  // radix-R1 butterflies per thread in step A
};

struct FastArgs {
  int mode, exact;
  int N, N1, N2, lN2;       // N2 = 1 << lN2
  int nsim, nseq, nrows;
  int seq_off, seq_cnt;     // this launch handles packed sequences [seq_off, seq_off + seq_cnt) (L2 slab)
  const double* fin; long ld_in;
  double* fout; long ld_out;
  const double* kvec;  // [nsim][N] (COLS) or [N] (ROWS)
  const double* cvec;  // v[ncols] (COLS) or e[nrows] (ROWS)
  double dt;
  double* phantom;
  const cplx* twN;  // exp(-2 pi i m/N)
  const cplx* twL1; // exp(-2 pi i m/N1), N1 entries
  const cplx* twL2; // exp(-2 pi i m/N2), N2 entries
  // optional fused charge density (COLS pass 3 epilogue): partial[tile_b][sim*N + x] = sum over the
  // tile's columns of w_j f[x][j] (np.trapz weights: dv, dv/2 on a global end column)
  double* dens_partial;
  double dv;
  int edge_flags;
  // optional scatter of the LAST pass to peer GPUs (multi-GPU layout change fused into the store):
  //  peer_mode 1 (ROWS, e df/dv of an x-shard): element (row, n) goes to the v-shard of rank
  //    q = n / part, at [my_rank * nrows + row][n % part], pitch part          (part = nv / P)
  //  peer_mode 2 (COLS, v df/dx of a v-shard): element (x, col) goes to the x-shard of rank
  //    q = x / part, at [x % part][my_rank * ncols + col], pitch ncols * P     (part = nx / P)
  int peer_mode, nparts, my_rank, lpart;   // part = 1 << lpart
  double* peer[8];
};

__device__ __forceinline__ cplx ldg_plain(const double* p) {
  const double2 v = *reinterpret_cast<const double2*>(p);
  return cmake(v.x, v.y);
}
__device__ __forceinline__ cplx ldg_c(const cplx* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cmake(v.x, v.y);
}

// global access of element n of packed sequence `seq` (COLS: column pair; ROWS: row pair)
template <int MODE>
__device__ __forceinline__ cplx gload(const FastArgs& a, const double* base, long ld, int sim, int seq, long n,
                                      bool first_pass) {
  if (MODE == ADV_COLS) {
    const double2 v = *reinterpret_cast<const double2*>(base + ((long)sim * a.N + n) * ld + 2 * (long)seq);
    return cmake(v.x, v.y);
  }
  const long ra = 2 * (long)seq, rb = ra + 1;
  double re = base[ra * ld + n];
  double im = (rb < a.nrows) ? base[rb * ld + n] : ((!first_pass && a.phantom) ? a.phantom[n] : 0.0);
  return cmake(re, im);
}

template <int MODE>
__device__ __forceinline__ void gstore(const FastArgs& a, int sim, int seq, long n, cplx val, bool last_pass) {
  if (last_pass && a.peer_mode) {
    // stores over NVLink straight into the target rank's shard (P2P-mapped pointers)
    if (MODE == ADV_COLS) {
      const int q = (int)(n >> a.lpart);
      const long xl = n & ((1L << a.lpart) - 1);
      const long ncols = 2L * a.nseq;
      double* dst = a.peer[q] + xl * (ncols * a.nparts) + (long)a.my_rank * ncols + 2 * (long)seq;
      *reinterpret_cast<double2*>(dst) = make_double2(val.x, val.y);
    } else {
      const int q = (int)(n >> a.lpart);
      const long part = 1L << a.lpart;
      const long c = n & (part - 1);
      const long ra = 2 * (long)seq, rb = ra + 1;
      double* base = a.peer[q] + ((long)a.my_rank * a.nrows) * part + c;
      base[ra * part] = val.x;
      if (rb < a.nrows) base[rb * part] = val.y;
    }
    return;
  }
  if (MODE == ADV_COLS) {
    *reinterpret_cast<double2*>(a.fout + ((long)sim * a.N + n) * a.ld_out + 2 * (long)seq) = make_double2(val.x, val.y);
    return;
  }
  const long ra = 2 * (long)seq, rb = ra + 1;
  a.fout[ra * a.ld_out + n] = val.x;
  if (rb < a.nrows) a.fout[rb * a.ld_out + n] = val.y;
  else if (!last_pass && a.phantom) a.phantom[n] = val.y;
}

// ------------------------------------------------------------------------------------------
// pass 1 (INV = 0): forward length-L transform over n1 for fixed n2 (the four-step twiddle
//                    W_N^(n2 k1) is applied by pass 2 when it loads, where it is a broadcast)
// pass 3 (INV = 1): inverse length-L transform over k1 for fixed n2
// Tile: CB batch lanes (COLS: packed columns of one n2; ROWS: consecutive n2 of one row pair).
// Block: CB * TPC threads, lanes run along the batch.
// ------------------------------------------------------------------------------------------
template <int L, int MODE, int INV, int CB>
__global__ void __launch_bounds__(CB* Geo<L>::TPC) pass13_kernel(const FastArgs a) {
  using G = Geo<L>;
  constexpr int R1 = G::R1, TPC = G::TPC, NA = G::NA;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* S = reinterpret_cast<cplx*>(smem_raw);  // [L][CB]
  const int b = threadIdx.x % CB;
  const int t = threadIdx.x / CB;  // 0..TPC-1
  // tile decode
  int sim = 0, seq, n2;
  bool valid;
  if (MODE == ADV_COLS) {
    const int tiles_b = (a.seq_cnt + CB - 1) / CB;
    const int bt = blockIdx.x % tiles_b;
    const long r = blockIdx.x / tiles_b;
    n2 = (int)(r % a.N2);
    sim = (int)(r / a.N2);
    seq = a.seq_off + bt * CB + b;
    valid = seq < a.seq_off + a.seq_cnt;
  } else {
    const long gl = (long)blockIdx.x * CB + b;        // lanes enumerate (sequence, n2)
    seq = a.seq_off + (int)(gl >> a.lN2);
    n2 = (int)(gl & (a.N2 - 1));
    valid = seq < a.seq_off + a.seq_cnt;
  }
  const long N2 = a.N2;
  // twiddles exp(-2 pi i m / L) of this length-L transform: one copy per CTA in shared memory (read
  // through the cache by every thread they were 15 dependent L2 round trips in the middle of the tile)
  cplx* TWS = S + L * CB;
  for (int w = threadIdx.x; w < L; w += CB * TPC) TWS[w] = ldg_c(a.twL1 + w);
  // COLS: element (l, seq) sits at base + l * rs; one base per thread and compile-time multiples of rs,
  // so neither loads nor stores wait for address registers of their predecessors
  const long rs_in = N2 * a.ld_in, rs_out = N2 * a.ld_out;
  const double* in0 = a.fin + ((long)sim * a.N + n2) * a.ld_in + 2L * seq;
  double* out0 = a.fout + ((long)sim * a.N + n2) * a.ld_out + 2L * seq;
  auto ldc = [](const double* p) { const double2 v = *reinterpret_cast<const double2*>(p); return cmake(v.x, v.y); };
  auto stc = [](double* p, cplx v) { *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y); };
  cplx x[16];
  if (!INV) {
    // ---- step A: radix-R1 over l = r + 8 j
    const double* src = a.fin;
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int r = t + TPC * q;
      const double* pr = in0 + (long)r * rs_in;
#pragma unroll
      for (int j = 0; j < R1; ++j) {
        if (MODE == ADV_COLS)
          x[q * R1 + j] = valid ? ldc(pr + (long)(8 * j) * rs_in) : cmake(0.0, 0.0);
        else
          x[q * R1 + j] = valid ? gload<MODE>(a, src, a.ld_in, sim, seq, (long)(r + 8 * j) * N2 + n2, true)
                                : cmake(0.0, 0.0);
      }
    }
    __syncthreads();                                     // TWS
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int r = t + TPC * q;
      fftR<R1, -1>(x + q * R1);
#pragma unroll
      for (int m = 1; m < R1; ++m) x[q * R1 + m] = cmul(x[q * R1 + m], TWS[r * m]);
#pragma unroll
      for (int m = 0; m < R1; ++m) S[(m * 8 + r) * CB + b] = x[q * R1 + m];
    }
    __syncthreads();
    // ---- step B: two radix-8 sub-transforms m = t, t + TPC
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int m = t + TPC * s;
#pragma unroll
      for (int r = 0; r < 8; ++r) x[s * 8 + r] = S[(m * 8 + r) * CB + b];
      fft8<-1>(x + s * 8);
      double* pm = out0 + (long)m * rs_out;
#pragma unroll
      for (int kp = 0; kp < 8; ++kp) {
        const int k1 = m + R1 * kp;
        if (!valid) continue;
        if (MODE == ADV_COLS) stc(pm + (long)(R1 * kp) * rs_out, x[s * 8 + kp]);
        else gstore<MODE>(a, sim, seq, (long)k1 * N2 + n2, x[s * 8 + kp], false);
      }
    }
  } else {
    // ---- step B': inverse radix-8 over k' for sub-transforms m = t, t + TPC
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int m = t + TPC * s;
      const double* pm = out0 + (long)m * rs_out;
#pragma unroll
      for (int kp = 0; kp < 8; ++kp) {
        if (MODE == ADV_COLS)
          x[s * 8 + kp] = valid ? ldc(pm + (long)(R1 * kp) * rs_out) : cmake(0.0, 0.0);
        else
          x[s * 8 + kp] = valid ? gload<MODE>(a, a.fout, a.ld_out, sim, seq, (long)(m + R1 * kp) * N2 + n2, false)
                                : cmake(0.0, 0.0);
      }
    }
    __syncthreads();                                     // TWS
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int m = t + TPC * s;
      fft8<1>(x + s * 8);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        cplx val = x[s * 8 + r];
        if (r * m != 0) val = cmulc(val, TWS[r * m]);
        S[(m * 8 + r) * CB + b] = val;
      }
    }
    __syncthreads();
    // ---- step A': inverse radix-R1 over m for r = t (+ TPC)
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int r = t + TPC * q;
#pragma unroll
      for (int m = 0; m < R1; ++m) x[q * R1 + m] = S[(m * 8 + r) * CB + b];
      fftR<R1, 1>(x + q * R1);
      if (MODE == ADV_COLS && !a.peer_mode) {
        double* pr = out0 + (long)r * rs_out;
#pragma unroll
        for (int j = 0; j < R1; ++j)
          if (valid) stc(pr + (long)(8 * j) * rs_out, x[q * R1 + j]);
      } else {
#pragma unroll
        for (int j = 0; j < R1; ++j)
          if (valid) gstore<MODE>(a, sim, seq, (long)(r + 8 * j) * N2 + n2, x[q * R1 + j], true);
      }
    }
    if (MODE == ADV_COLS && a.dens_partial != nullptr) {
      // charge density of the new f (vlapy/core/field.py:27-36): weighted sum over this tile's
      // 2*CB real columns.  Every thread parks its 16 weighted pairs in the (now free) exchange
      // buffer, D[l][b] with an odd pitch; thread l then adds the CB entries of row l in a fixed
      // order: one partial row per column tile.  (A recursive-halving shuffle reduction without shared
      // memory was measured slower: 0.97 ms against 0.83 ms for pass 3.)
      double* D = reinterpret_cast<double*>(smem_raw);
      constexpr int DP = CB + 1;
      const int ncols = 2 * a.nseq;
      const double wa = (valid ? ((2 * seq == 0 && (a.edge_flags & 1)) ? 0.5 * a.dv : a.dv) : 0.0);
      const double wb = (valid ? ((2 * seq + 1 == ncols - 1 && (a.edge_flags & 2)) ? 0.5 * a.dv : a.dv) : 0.0);
      const int tiles_b = (a.seq_cnt + CB - 1) / CB;
      const int bt = a.seq_off / CB + blockIdx.x % tiles_b;
      __syncthreads();                                   // every thread has read its part of S
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const int r = t + TPC * q;
#pragma unroll
        for (int j = 0; j < R1; ++j) D[(r + 8 * j) * DP + b] = wa * x[q * R1 + j].x + wb * x[q * R1 + j].y;
      }
      __syncthreads();
      const int l = threadIdx.x;
      if (l < L) {
        double d = 0.0;
#pragma unroll 8
        for (int bb = 0; bb < CB; ++bb) d += D[l * DP + bb];
        // cell x = l N2 + n2 is parked at n2 L + l: the CTA writes 1 KB in one piece (x order: 128 lone 8-byte writes,
        // 1 KB apart); dens_reduce_kernel puts the sums back in x order
        a.dens_partial[((long)bt * a.nsim + sim) * a.N + (long)n2 * L + l] = d;
      }
    }
  }
}

// second stage of the fused density: n[x] = sum over column tiles (fixed order, deterministic)
// Two stages when there are many tiles (the kernel has only n threads): blockIdx.y = group g sums its tiles
// [g gs, (g+1) gs) in order and leaves the result IN PLACE in the group's first tile row (read and written by the same
// thread); the second launch (stride = gs) adds the group sums in order.  Deterministic.  (Measured and rejected: eight
// interleaved accumulators per thread, 90 us against 59 us for 512 x 16384 partials.)
// The partial rows hold cell x = l N2 + n2 of a simulation at n2 L + l (pass 3 writes them that way); the final stage
// stores in x order.
__global__ void dens_reduce_kernel(double* __restrict__ partial, int ntiles, int stride, long n, double* __restrict__ out,
                                   int L, int N2) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t0 = blockIdx.y * ntiles * stride;
  double s = 0.0;
  for (int t = 0; t < ntiles; ++t) s += partial[(long)(t0 + t * stride) * n + i];
  if (out) {
    const long N = (long)L * N2, p = i % N;
    out[i - p + (p % L) * N2 + p / L] = s;
  } else {
    partial[(long)t0 * n + i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// pass 2: for the group pair {k1, N1-k1}: four-step twiddle, forward length-L transform over n2,
// pointwise phase multiply with un-mixing, inverse transform, conjugate twiddle.  L = N2.
// Block: CB * 2*TPC threads.  The CTA keeps its CB packed sequences and walks over t1_chunk
// consecutive group-pair tiles: phase tables are built once, and the next tile is prefetched
// with cp.async into a staging buffer while the current one is transformed.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int L, int MODE, int CB>
struct P2Layout {
  using G = Geo<L>;
  static constexpr int R1 = G::R1;
  // exchange buffer index of (group g, sub-transform m, position r, sequence b)
  __device__ static __forceinline__ int sidx(int g, int m, int r, int b) {
    if (MODE == ADV_COLS) return (g * L + m * 8 + r) * CB + b;
    return (b * 2 + g) * (R1 * 9) + m * 9 + r;   // 9: odd pitch, conflict free for lanes along r or m
  }
  static constexpr int S_ELEMS = (MODE == ADV_COLS) ? 2 * L * CB : CB * 2 * R1 * 9;
};

// The next tile is prefetched with cp.async INTO THE EXCHANGE BUFFER: a thread reads last (inverse step A') and first
// (step A of the next tile) the same 16 slots of S, so it prefetches exactly those and needs no barrier for them
// and no second buffer.  (Round 1 also measured direct loads at 4 CTAs/SM and a staged second buffer at 2 CTAs/SM:
// both slower; PFM stays in the signature as the tag of the surviving variant.)
// EX: per-bin sincos of the reference's own rounding (VPFP_PHASE_EXACT) instead of the geometric tables -- a
// template parameter, so that the table kernel does not carry the registers and code of the sincos path.
template <int L, int MODE, int CB, int PFM, bool EX>
__global__ void __launch_bounds__(CB * 2 * Geo<L>::TPC, (PFM == 1 || CB * 2 * Geo<L>::TPC > 128) ? 2 : 4) pass2_kernel(const FastArgs a, const int t1_chunk) {
  static_assert(PFM == 2, "only the own-slot prefetch variant is built");
  using G = Geo<L>;
  using LY = P2Layout<L, MODE, CB>;
  constexpr int R1 = G::R1, TPC = G::TPC, NA = G::NA;
  constexpr int NT = CB * 2 * TPC;
  constexpr int HALF = L / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NPT = 8 + HALF / 8 + 1;               // Lo[0..7] = G^j, Hi[0..HALF/8] = G^(8j)
  cplx* S = reinterpret_cast<cplx*>(smem_raw);        // exchange
  cplx* PT = S + LY::S_ELEMS;                         // [2 chan][CB][NPT]
  cplx* BASE0 = PT + 2 * CB * NPT;                    // [2 tiles (parity)][2 groups][2 chan][CB]
  cplx* TWL = BASE0 + 8 * CB;                         // [L]     exp(-2 pi i m / L)
  cplx* TWT = TWL + L;                                // [2][L]  four-step twiddles W_N^(n2 k1) of the staged tile
  cplx* TWC = TWT + 2 * L;                            // [2][L]  second buffer of the four-step twiddles
  double* PHI = reinterpret_cast<double*>(TWC + 2 * L);                       // [2 chan][CB]

  // thread roles: b = packed sequence, u = 0..R1-1 (step A: group u / TPC, r-slot u % TPC)
  int b, u;
  if (MODE == ADV_COLS) { b = threadIdx.x % CB; u = threadIdx.x / CB; }
  else { const int ta = threadIdx.x % TPC; b = (threadIdx.x / TPC) % CB; u = (threadIdx.x / (TPC * CB)) * TPC + ta; }
  const int T1 = a.N1 / 2;
  const int nchunks = (T1 + t1_chunk - 1) / t1_chunk;
  const int tiles_b = (a.seq_cnt + CB - 1) / CB;
  int sim = 0;
  const int chunk = blockIdx.x % nchunks;
  const long rest = blockIdx.x / nchunks;
  const int bt = (int)(rest % tiles_b);
  if (MODE == ADV_COLS) sim = (int)(rest / tiles_b);
  const int seq0 = a.seq_off + bt * CB;                  // first packed sequence of this CTA
  const int seq_end = a.seq_off + a.seq_cnt;
  const int seq = seq0 + b;
  const bool valid = seq < seq_end;
  const long N = a.N, N1 = a.N1;
  const double inv_n = 1.0 / (double)N;

  double ca = 0.0, cb = 0.0;
  if (valid) {
    const long ra = 2 * (long)seq, rb = ra + 1;
    ca = a.cvec[ra];
    cb = (MODE == ADV_COLS || rb < a.nrows) ? a.cvec[rb] : 0.0;
  }
  const double* K = a.kvec + (MODE == ADV_COLS ? (long)sim * N : 0);

  // this thread's 16 elements of the tile of group pair t1 -> its own slots of S, and its share
  // of the tile's four-step twiddles -> twbuf
  // COLS: column base of this thread's packed sequence (element of row n at col0 + n * ld_out)
  double* const col0 = a.fout + (long)sim * N * a.ld_out + 2 * (long)seq;
  auto prefetch_own = [&](int t1, cplx* twbuf) {
    const int k1g0 = (t1 == 0) ? 0 : t1, k1g1 = (t1 == 0) ? (int)(N1 / 2) : (int)(N1 - t1);
    const int gq = u / TPC, taq = u % TPC;
    const long k1 = gq ? k1g1 : k1g0;
#pragma unroll
    for (int q = 0; q < NA; ++q)
#pragma unroll
      for (int j = 0; j < R1; ++j) {
        const int r = taq + TPC * q;
        cplx* dst = S + LY::sidx(gq, j, r, b);
        const long n = k1 * L + r + 8 * j;
        if (!valid) { *dst = cmake(0.0, 0.0); continue; }
        if (MODE == ADV_COLS) {
          // one base per (tile, thread), compile-time multiples of the row pitch
          cp_async16(dst, col0 + (k1 * L + taq) * a.ld_out + (long)(TPC * q + 8 * j) * a.ld_out);
        } else {
          const long ra = 2 * (long)seq, rb = ra + 1;
          cp_async8(&dst->x, a.fout + ra * a.ld_out + n);
          if (rb < a.nrows) cp_async8(&dst->y, a.fout + rb * a.ld_out + n);
          else dst->y = a.phantom ? a.phantom[n] : 0.0;
        }
      }
    for (int w = threadIdx.x; w < 2 * L; w += NT) {
      const int n2 = w % L, gg = w / L;
      cp_async16(twbuf + w, a.twN + (long)n2 * (gg ? k1g1 : k1g0));
    }
    cp_async_commit();
  };

  for (int w = threadIdx.x; w < L; w += NT) TWL[w] = ldg_c(a.twL2 + w);
  // tiles t1 = t1_begin + it, it = 0 .. t1_chunk - 1 (t1 < T1): the loop runs on `it`, whose bound is a kernel
  // parameter, so that no per-thread loop bounds stay live across the tile (they were spilled, and their reloads
  // queued behind the cp.async / LDS bursts: 12 % of the kernel in the ncu source view)
  const int t1_begin = chunk * t1_chunk;
  prefetch_own(t1_begin, TWT);

  if (!EX) {
    // phi = (K[1] dt) c ; G = exp(-i N1 phi); PT[j] = G^j = H[j>>3] * Lo[j&7]
    for (int w = threadIdx.x; w < 2 * CB; w += NT) {
      const int ch = w / CB, bb = w % CB;
      const int sq = seq0 + bb;
      double c = 0.0;
      if (sq < seq_end) {
        const long rr = 2 * (long)sq + ch;
        c = (MODE == ADV_COLS || rr < a.nrows) ? a.cvec[rr] : 0.0;
      }
      PHI[w] = mul_rn(mul_rn(K[1], a.dt), c);
    }
    __syncthreads();
    for (int w = threadIdx.x; w < 2 * CB * NPT; w += NT) {
      const int sq = w / NPT, i = w % NPT;
      const int j = (i < 8) ? i : 8 * (i - 8);          // Lo: j = 0..7, Hi: j = 0, 8, .., HALF
      const double th = (double)N1 * PHI[sq] * (double)j;
      double sn, cs;
      sincos(th, &sn, &cs);
      PT[sq * NPT + i] = cmake(cs, -sn);
    }
  }

#pragma unroll 1
  for (int it = 0; it < t1_chunk; ++it) {
    const int t1 = chunk * t1_chunk + it;
    if (t1 >= T1) break;
    const bool has_next = (it + 1 < t1_chunk) && (t1 + 1 < T1);
    const bool self = (t1 == 0);
    const int k1g0 = self ? 0 : t1, k1g1 = self ? (int)(N1 / 2) : (int)(N1 - t1);
#define K1G(g) ((g) ? k1g1 : k1g0)
    // BASE alternates between two buffers: the writers of tile t1 cannot overtake the pointwise readers
    // of tile t1-2 (barriers of tile t1-1 lie between), so no barrier is needed here
    cplx* BASE = BASE0 + (it & 1) * 4 * CB;
    if (!EX) {
      for (int w = threadIdx.x; w < 4 * CB; w += NT) {     // base(k1) per group/channel, scaled by 1/(2N)
        const int bb = w % CB, gc = w / CB, g = gc >> 1, ch = gc & 1;
        double sn, cs;
        sincos((double)K1G(g) * PHI[ch * CB + bb], &sn, &cs);
        BASE[(g * 2 + ch) * CB + bb] = cmake(cs * 0.5 * inv_n, -sn * 0.5 * inv_n);
      }
    }
    cplx x[16];
    const int g = u / TPC, ta = u % TPC;
    const cplx* TWI;                                   // twiddles for the inverse at the end of this tile
    {
      // own slots of S and this tile's twiddle buffer were filled by prefetch_own one tile earlier
      cplx* TWcur = (it & 1) ? TWC : TWT;
      TWI = TWcur;
      cp_async_wait_all();
      __syncthreads();                                 // twiddles (and BASE) visible; previous tile done with BASE
#pragma unroll
      for (int q = 0; q < NA; ++q)
#pragma unroll
        for (int j = 0; j < R1; ++j) {
          const int r = ta + TPC * q;
          x[q * R1 + j] = cmul(S[LY::sidx(g, j, r, b)], TWcur[g * L + r + 8 * j]);
        }
    }
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int r = ta + TPC * q;
      fftR<R1, -1>(x + q * R1);
#pragma unroll
      for (int m = 1; m < R1; ++m) x[q * R1 + m] = cmul(x[q * R1 + m], TWL[r * m]);
#pragma unroll
      for (int m = 0; m < R1; ++m) S[LY::sidx(g, m, r, b)] = x[q * R1 + m];
    }
    __syncthreads();
    // ---------------- step B: sub-transforms (gA, mA) in x[0..7], (gB, mB) in x[8..15]
    int gA, mA, gB, mB;
    bool special = false;  // the thread holding m = 0 and m = R1/2 of the k1 = 0 group
    if (!self) {
      gA = 0; mA = u; gB = 1; mB = R1 - 1 - u;
    } else if (u < R1 / 2) {
      gA = gB = 0;
      if (u == 0) { mA = 0; mB = R1 / 2; special = true; }
      else { mA = u; mB = R1 - u; }
    } else {
      gA = gB = 1; mA = u - R1 / 2; mB = R1 - 1 - mA;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      x[r] = S[LY::sidx(gA, mA, r, b)];
      x[8 + r] = S[LY::sidx(gB, mB, r, b)];
    }
    fft8<-1>(x);
    fft8<-1>(x + 8);
    // ---------------- pointwise: pairs (Z at bin k, Zp at bin N - k), both in this thread
    {
      const long k1A = K1G(gA);
      // table phases: P = base(k1) * Hi[j >> 3] * Lo[j & 7]; within a thread j & 7 takes one value for
      // the bins below Nyquist and one above (k2 = m + R1 k', R1 a multiple of 8), base one per group
      cplx baseA_a, baseA_b, loP_a, loP_b, loN_a, loN_b;
      if (!EX) {
        baseA_a = BASE[(gA * 2 + 0) * CB + b]; baseA_b = BASE[(gA * 2 + 1) * CB + b];
        const int mlo = (special ? 0 : mA) & 7;
        loP_a = PT[b * NPT + mlo]; loP_b = PT[(CB + b) * NPT + mlo];
        loN_a = PT[b * NPT + ((8 - mlo) & 7)]; loN_b = PT[(CB + b) * NPT + ((8 - mlo) & 7)];
      }
      auto pair_op = [&](cplx& Zr, cplx& Zpr, const long kbin, const long k1, const int gsel, const bool selfpair) {
        const bool neg = (2 * kbin > N);
        const bool nyq = (2 * kbin == N);
        cplx Pa, Pb;  // scaled by 1/(2N)
        if (EX) {
          const long kr = neg ? N - kbin : kbin;
          const double kdt = mul_rn(K[kr], a.dt);
          double sn, cs;
          sincos(mul_rn(kdt, ca), &sn, &cs);
          Pa = cmake(cs * 0.5 * inv_n, (nyq ? 0.0 : (neg ? sn : -sn)) * 0.5 * inv_n);
          sincos(mul_rn(kdt, cb), &sn, &cs);
          Pb = cmake(cs * 0.5 * inv_n, (nyq ? 0.0 : (neg ? sn : -sn)) * 0.5 * inv_n);
        } else {
          const int k2 = (int)((kbin - k1) / N1);           // 0..L-1
          const int j = neg ? (L - k2) : k2;                // |signed k2| in 0..HALF
          const bool hoisted = (R1 % 8 == 0) && ((!special) || (k1 == 0 && ((k2 & 7) == 0)));
          cplx la = hoisted ? (neg ? loN_a : loP_a) : PT[b * NPT + (j & 7)];
          cplx lb = hoisted ? (neg ? loN_b : loP_b) : PT[(CB + b) * NPT + (j & 7)];
          cplx ta_ = cmul(PT[b * NPT + 8 + (j >> 3)], la);
          cplx tb_ = cmul(PT[(CB + b) * NPT + 8 + (j >> 3)], lb);
          if (neg) { ta_ = cconj(ta_); tb_ = cconj(tb_); }
          (void)gsel;
          Pa = cmul(baseA_a, ta_);
          Pb = cmul(baseA_b, tb_);
          if (nyq) { Pa.y = 0.0; Pb.y = 0.0; }
        }
        const cplx Z = Zr, Zp = Zpr;
        const cplx U = cadd(Z, cconj(Zp)), V = csub(Z, cconj(Zp));
        const cplx X1 = cmul(Pa, U), X2 = cmul(Pb, V);
        Zr = cadd(X1, X2);
        if (!selfpair) Zpr = cconj(csub(X1, X2));
      };
      if (!special) {
        // cross pattern: A[k'] (bin k1A + N1 (mA + R1 k')) pairs with B[7 - k']
#pragma unroll
        for (int pr = 0; pr < 8; ++pr) pair_op(x[pr], x[15 - pr], k1A + N1 * (mA + R1 * pr), k1A, gA, false);
      } else {
        // k1 = 0 group, sub-transforms m = 0 (x[0..7]) and m = R1/2 (x[8..15]), paired inside themselves
        pair_op(x[0], x[0], 0, 0, 0, true);                              // DC
        pair_op(x[4], x[4], N1 * (long)(R1 * 4), 0, 0, true);            // Nyquist
#pragma unroll
        for (int pr = 1; pr < 4; ++pr) pair_op(x[pr], x[8 - pr], N1 * (long)(R1 * pr), 0, 0, false);
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) pair_op(x[8 + pr], x[15 - pr], N1 * (long)(R1 / 2 + R1 * pr), 0, 0, false);
      }
    }
    // ---------------- step B': inverse radix-8, conj twiddle, exchange
    fft8<1>(x);
    fft8<1>(x + 8);
    // (in place: the slots written here are the slots this thread read for step B, no barrier)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      cplx va = x[r], vb = x[8 + r];
      va = cmulc(va, TWL[r * mA]);
      vb = cmulc(vb, TWL[r * mB]);
      S[LY::sidx(gA, mA, r, b)] = va;
      S[LY::sidx(gB, mB, r, b)] = vb;
    }
    __syncthreads();
    // ---------------- step A': inverse radix-R1, conj four-step twiddle, store
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int r = ta + TPC * q;
#pragma unroll
      for (int m = 0; m < R1; ++m) x[q * R1 + m] = S[LY::sidx(g, m, r, b)];
    }
    if (has_next) prefetch_own(t1 + 1, ((it + 1) & 1) ? TWC : TWT);
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int r = ta + TPC * q;
      fftR<R1, 1>(x + q * R1);
      double* const pr = col0 + ((long)K1G(g) * L + r) * a.ld_out;
#pragma unroll
      for (int j = 0; j < R1; ++j) {
        const int n2 = r + 8 * j;
        const cplx val = cmulc(x[q * R1 + j], TWI[g * L + n2]);
        if (!valid) continue;
        if (MODE == ADV_COLS) *reinterpret_cast<double2*>(pr + (long)(8 * j) * a.ld_out) = make_double2(val.x, val.y);
        else gstore<MODE>(a, sim, seq, (long)K1G(g) * L + n2, val, false);
      }
    }
#undef K1G
  }
}

template <int L, int CB>
constexpr size_t pass2_smem(int mode, int pfm) {
  (void)pfm;
  return sizeof(cplx) * (size_t)(((mode == ADV_ROWS) ? CB * 2 * (L / 8) * 9 : 2 * L * CB) +
                                 2 * CB * (8 + L / 16 + 1) + 8 * CB + 3 * L + 2 * L) +
         sizeof(double) * 2 * CB;
}

}  // namespace fast
