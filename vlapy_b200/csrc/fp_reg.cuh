// fp_reg.cuh -- implicit Fokker-Planck step with the row held in REGISTERS and the next row
// streaming into shared memory behind it.  Same decomposition as fp_fast.cuh (chunk interiors
// eliminated exactly, tridiagonal separator system solved by cyclic reduction, interiors
// back-substituted); replaces vlapy/core/collisions.py:44-81, 104-158, 232-263 via
// vlapy/core/step.py:102-108, and the row moments of vlapy/core/step.py:153-173, 189-226.
//
// What is different from fp_fast.cuh (nv = 32 T, T = 128 / 256 / 512 threads, one row per CTA):
//  * a thread owns one chunk of M = 32 consecutive cells in registers (64 of its 128 registers);
//    every sweep is fully unrolled, so the chunk is addressed statically and shared memory is not
//    touched between the load and the store of a row;
//  * shared memory is therefore free to receive the NEXT row of the CTA (cp.async, issued as soon
//    as every thread has pulled its chunk out): the HBM read of row r+1 overlaps all arithmetic of
//    row r, and the stores of row r (straight from registers, 16 bytes per instruction, every
//    32-byte sector completed by the same thread) overlap its moment sums;
//  * the eliminations are DIVISION FREE.  The chunk system is scaled by the power of two that brings
//    its diagonal into [1, 2) (b' = sc b, a'_i = sc A_i, c'_i = sc C_i: exact, so the matrix entries
//    are the reference's own roundings) and its pivots are carried as ratios of leading-minor
//    determinants, p_i = N_i / N_{i-1}:
//        N_i = b' N_{i-1} - (a'_i c'_{i-1}) N_{i-2}       (N_{-1} = 1, N_0 = b',  2^-i <= N_i <= 2^(i+1))
//        Z_i = d_i N_{i-1} - a'_i Z_{i-1}                 (eliminated right-hand side  z_i = Z_i / N_{i-1})
//        G_i = -a'_i G_{i-1}                              (eliminated unit spike       g_i = G_i / N_{i-1})
//    one FMA on the critical path per cell instead of multiply, FMA and a Newton reciprocal.  The
//    spike sweeps (LU downwards, UL upwards, both in ONE unrolled loop: independent recurrences
//    hide each other's latency) need a single reciprocal each, 1 / det; the back-substitution
//        x_i = (Z_i - c'_i N_{i-1} x_{i+1}) / N_i
//    takes 1/N_i from MUFU.RCP64H and two Newton steps, off the critical path;
//  * determinants needed by the back-substitution are parked in shared memory (12 per thread); the
//    back-substitution runs in three groups, the determinants of the two lower groups are
//    regenerated from a two-value checkpoint (4 flops per cell).
// Diagonal dominance (|a'| + |c'| < b') keeps every pivot in [b'/2, b']: no overflow or underflow for
// any nu, dt.  The staging layout gives every chunk a 16-byte pad (pitch 272 bytes): the coalesced
// 16-byte cp.async writes and the per-thread 16-byte chunk reads are both conflict free.
#pragma once
#include "fp_fast.cuh"

namespace fpreg {

using fpfast::Args;
using fpfast::rcp_near;
using fpfast::warp_sum;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
#if defined(__CUDA_ARCH__)
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
#else
  memcpy(smem, gmem, 16);
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}
__device__ __forceinline__ double2 ld2(const double* p) {     // 16-byte aligned pair
#if defined(__CUDA_ARCH__)
  return *reinterpret_cast<const double2*>(p);
#else
  double2 v;
  v.x = p[0]; v.y = p[1];
  return v;
#endif
}
// the same, never merged with an earlier load of the same address (keeps the values of the earlier
// load from staying live in registers until this one)
__device__ __forceinline__ double2 ld2_fresh(const double* p) {
#if defined(__CUDA_ARCH__)
  double2 v;
  const unsigned sa = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(sa));
  return v;
#else
  return ld2(p);
#endif
}
__device__ __forceinline__ void store2(double* p, double x, double y) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<double2*>(p) = make_double2(x, y);
#else
  p[0] = x; p[1] = y;
#endif
}

// 1/p for a well-scaled positive p: hardware seed (MUFU.RCP64H, ~2^-23) + two Newton steps (< 1 ulp)
__device__ __forceinline__ double rcp_fast(double p) {
  double r;
#if defined(__CUDA_ARCH__)
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
#else
  r = 1.0 / p;
  long long b;
  memcpy(&b, &r, 8);
  b &= ~0xffffffffLL;                                   // what the hardware seed keeps (tests/emul)
  memcpy(&r, &b, 8);
#endif
  double e = fma(-p, r, 1.0);
  r = fma(r, e, r);
  e = fma(-p, r, 1.0);
  return fma(r, e, r);
}

constexpr int imax(int a, int b) { return a > b ? a : b; }

template <int M, int T>
struct Geo {
  static constexpr int MI = M - 1;                       // interior cells of a chunk
  static constexpr int PITCH = M + 2;                    // doubles per chunk in the staging buffer
  static constexpr int G1 = MI / 3, G2 = (2 * MI) / 3;   // back-substitution groups [0,G1) [G1,G2) [G2,MI)
  static constexpr int PCAP = imax(imax(G1, G2 - G1), MI - G2) + 1;   // N_{lo-1} .. N_{hi-1} of a group
  static constexpr size_t SMEM = sizeof(double) * (size_t)(T * PITCH + 64 + 8 * T + PCAP * T + 1024);
};

// ln x = e ln2 + ln c_i + log1p(r) with a 64-entry table (c_i = 1 + (i + 1/2)/64, |r| < 2^-7) and a
// degree-6 series: absolute error < 3e-16.  The table is REPLICATED 8 times in shared memory, entry i
// of replica g at 16-byte slot 8 i + g, and a lane reads replica (lane & 7): the eight lanes that
// share a 128-bit shared-memory transaction hit eight different bank groups whatever their
// arguments are (one shared copy costs ~3x the transactions in bank conflicts).
// Returns ln c_i + log1p(r) and the exponent e separately (the caller sums f * e and multiplies by
// ln2 once).  BRANCH FREE, so that the 32 logarithms of a chunk interleave: non-positive, subnormal
// or non-finite arguments only raise `bad`, and the caller then redoes the two logarithm sums of the
// chunk on the library path (numpy's NaN / -inf semantics, vlapy/core/step.py:222-224).
__device__ __forceinline__ double log_split(double x, const double2* __restrict__ tab_lane, double* e_out, int* bad) {
  const long long bits = __double_as_longlong(x);
  const int hi = (int)(bits >> 32);
  *bad |= ((unsigned)(hi - 0x00100000) >= (unsigned)(0x7ff00000 - 0x00100000)) ? 1 : 0;
  const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const double2 tb = tab_lane[((hi >> 14) & 63) * 8];     // (1/c_i, ln c_i), this lane's replica
  const double r = fma(m, tb.x, -1.0);
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(r, p, -0.25);
  p = fma(r, p, 1.0 / 3.0);
  p = fma(r, p, -0.5);
  p = fma(r * r, p, r);
  // biased exponent -> double without a conversion instruction: 2^52 + ex as a bit pattern
  const double eb = __longlong_as_double(0x4330000000000000LL | (long long)((hi >> 20) & 0x7ff));
  *e_out = eb - 4503599627371519.0;                       // 2^52 + 1023
  return tb.y + p;
}

// (i - c)^q for the local monomial sums (exact in double for the chunk sizes used here)
__host__ __device__ constexpr double kpow(double d, int q) { return q == 0 ? 1.0 : d * kpow(d, q - 1); }

// sums of one chunk: mu[q] = sum_i x_i (i - I0)^q (q = 0..5), s2 = sum x^2, sl = sum x (ln c + log1p r),
// se = sum x e.  W = 1: whole cell; W = -1/2: np.trapz correction of an end cell.
struct ChunkSums {
  double mu[6], s2, sl, se;
  int bad;
};
template <int M, int I, int HALF>
__device__ __forceinline__ void chunk_terms(ChunkSums& S, double x, const double2* LT) {
  constexpr double d = (double)I - 0.5 * (double)(M - 1);
  const double xs = HALF ? -0.5 * x : x;
  double e;
  const double lx = log_split(x, LT, &e, &S.bad);
  S.mu[0] += xs;
  S.mu[1] = fma(xs, kpow(d, 1), S.mu[1]);
  S.mu[2] = fma(xs, kpow(d, 2), S.mu[2]);
  S.mu[3] = fma(xs, kpow(d, 3), S.mu[3]);
  S.mu[4] = fma(xs, kpow(d, 4), S.mu[4]);
  S.mu[5] = fma(xs, kpow(d, 5), S.mu[5]);
  S.s2 = fma(xs, x, S.s2);
  S.sl = fma(xs, lx, S.sl);
  S.se = fma(xs, e, S.se);
}
// the same with the cell index as an argument (constant after unrolling): for the loop that interleaves the moment
// sums with the transposed stores
// the same with the cell index as an argument (constant after unrolling): for the loop that interleaves the moment
// sums with the transposed stores.  (Measured and rejected: monomial sums on the mirror pairs (i, M-1-i), 8
// instructions per pair instead of 12 -- the longer live ranges cost more in spills than the arithmetic saved:
// 1.61 ms against 1.55 ms.)
template <int M>
__device__ __forceinline__ void chunk_terms_at(ChunkSums& S, double x, const double2* LT, const int i) {
  const double d = (double)i - 0.5 * (double)(M - 1);
  const double d2 = d * d;
  double e;
  const double lx = log_split(x, LT, &e, &S.bad);
  S.mu[0] += x;
  S.mu[1] = fma(x, d, S.mu[1]);
  S.mu[2] = fma(x, d2, S.mu[2]);
  S.mu[3] = fma(x, d2 * d, S.mu[3]);
  S.mu[4] = fma(x, d2 * d2, S.mu[4]);
  S.mu[5] = fma(x, d2 * d2 * d, S.mu[5]);
  S.s2 = fma(x, x, S.s2);
  S.sl = fma(x, lx, S.sl);
  S.se = fma(x, e, S.se);
}
template <int M, int I>
struct ChunkLoop {
  __device__ __forceinline__ static void run(ChunkSums& S, const double (&c)[M], const double2* LT) {
    chunk_terms<M, I, 0>(S, c[I], LT);
    ChunkLoop<M, I + 1>::run(S, c, LT);
  }
  // library path of the two logarithm sums (a chunk holds a cell that log_split cannot take)
  __device__ __forceinline__ static void rare(double& sl, const double (&c)[M]) {
    sl = fma(c[I], fpfast::log_rare(c[I]), sl);
    ChunkLoop<M, I + 1>::rare(sl, c);
  }
};
template <int M>
struct ChunkLoop<M, M> {
  __device__ __forceinline__ static void run(ChunkSums&, const double (&)[M], const double2*) {}
  __device__ __forceinline__ static void rare(double&, const double (&)[M]) {}
};

// OP: 0 Lenard-Bernstein, 1 Dougherty (a template parameter: the kernel of one operator does not carry the
// registers and code of the other)
template <int M, int T, int OP>
__global__ void __launch_bounds__(T, (M >= 64 ? 256 : 512) / T) fp_reg_kernel(const Args a) {
  using G = Geo<M, T>;
  static_assert(M >= 16 && G::G1 >= 3 && (M % 2) == 0 && T % 32 == 0 && T <= 512, "chunk geometry");
  constexpr int NW = T / 32;
  constexpr int MI = G::MI, PITCH = G::PITCH, G1 = G::G1, G2 = G::G2, PCAP = G::PCAP;
  constexpr int U = M / 2;                               // 16-byte units per chunk
  constexpr int WP = 18;                                 // doubles per lane-row of a warp's transposition slice
  static_assert(NW * 32 * WP <= (8 + PCAP) * T, "transposition slices fit the scratch area");
  VPFP_DYN_SMEM(smem_raw);
  double* stage = reinterpret_cast<double*>(smem_raw);   // T * PITCH: the row being loaded
  double* red = stage + T * PITCH;                       // 64
  double* X = red + 64;                                  // 8 * T scratch (separator system / moments)
  volatile double* PV = X + 8 * T;                       // PCAP * T parked pivots (private to a thread)
  double2* LT = reinterpret_cast<double2*>(X + 8 * T + PCAP * T);   // 64 x 8 entries of the log table
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  for (int i = t; i < 512; i += T) LT[i] = a.logtab64[i >> 3];      // 8 replicas, see log_split
  const double2* const LTl = LT + (t & 7);                          // this lane's replica
  const int s = t * M;
  const double vs0 = fma((double)s, a.vstep, a.v0);      // velocity of the chunk's first cell
  const bool first_thread = (t == 0), last_thread = (t == T - 1);
  const double zero = (a.rows < 0) ? 1.0 : 0.0;          // 0.0 the compiler cannot see through (TIE below)

  // 16-byte unit u = k T + t of the row goes to chunk u / U at offset u % U: one base per thread,
  // compile-time strides (T is a multiple of U)
  static_assert(T % (M / 2) == 0, "staging layout");
  double* const stage_t = stage + (t / U) * PITCH + 2 * (t % U);
  auto prefetch = [&](long r) {
    const double* src = a.fin + r * a.ld_in + 2 * t;
#pragma unroll
    for (int k = 0; k < U; ++k) cp_async16(stage_t + k * (T / U) * PITCH, src + k * 2 * T);
    cp_async_commit();
  };

  long r = blockIdx.x;
  if (r < a.rows) prefetch(r);
  for (; r < a.rows; r += gridDim.x) {
    cp_async_wait_all();
    __syncthreads();
    // ---------------- first moment straight from the staged row (unit weights, np.trapz ends fixed below)
    const double* const sp = stage + t * PITCH;           // this thread's chunk in the staging buffer
    // (+ 0 * r: keeps the 32 cell velocities from being hoisted out of the row loop and spilled)
    const double vs = fma(zero, (double)r, vs0);
    const double v_lo = vs, v_hi = fma((double)(M - 1), a.vstep, vs);
    double acc0 = 0.0;
    if (OP == 0) {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const double2 f2 = ld2(sp + 2 * j);
        const double va = fma((double)(2 * j), a.vstep, vs), vb = fma((double)(2 * j + 1), a.vstep, vs);
        acc0 = fma(f2.x * va, va, acc0);
        acc0 = fma(f2.y * vb, vb, acc0);
      }
      if (first_thread) acc0 = fma(-0.5 * sp[0] * v_lo, v_lo, acc0);
      if (last_thread) acc0 = fma(-0.5 * sp[M - 1] * v_hi, v_hi, acc0);
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const double2 f2 = ld2(sp + 2 * j);
        acc0 = fma(f2.x, fma((double)(2 * j), a.vstep, vs), acc0);
        acc0 = fma(f2.y, fma((double)(2 * j + 1), a.vstep, vs), acc0);
      }
      if (first_thread) acc0 = fma(-0.5 * sp[0], v_lo, acc0);
      if (last_thread) acc0 = fma(-0.5 * sp[M - 1], v_hi, acc0);
    }
    // (red was last read before the barrier at the top of this row)
    double Tm = fpfast::block_sum<T, false>(acc0 * a.dv, red), vbar = 0.0;
    if (OP == 1) {
      vbar = Tm;
      double acc1 = 0.0;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const double2 f2 = ld2(sp + 2 * j);
        const double da = fma((double)(2 * j), a.vstep, vs) - vbar, db = fma((double)(2 * j + 1), a.vstep, vs) - vbar;
        acc1 = fma(f2.x * da, da, acc1);
        acc1 = fma(f2.y * db, db, acc1);
      }
      if (first_thread) acc1 = fma(-0.5 * sp[0] * (v_lo - vbar), v_lo - vbar, acc1);
      if (last_thread) acc1 = fma(-0.5 * sp[M - 1] * (v_hi - vbar), v_hi - vbar, acc1);
      Tm = fpfast::block_sum<T>(acc1 * a.dv, red);
    }
    // diagonals, affine in the cell index i (see fp_fast.cuh):
    //   A_i = nudt(tdv + (v_{i-1} - vbar)/2/dv) = A0 + i dA,  C_i = nudt(tdv - (v_{i+1} - vbar)/2/dv) = C0 - i dA
    const double nudt = a.nu * a.dt;
    const double tdv = -Tm / (a.dv * a.dv);
    const double bd = 1.0 + nudt * (2.0 * Tm / (a.dv * a.dv));
    // exact power-of-two scaling of the system: sc = 2^-k with bt = sc * bd in [1, 2)
    const double sc = __longlong_as_double((2046LL - ((__double_as_longlong(bd) >> 52) & 0x7ff)) << 52);
    const double bt = bd * sc;
    const double hb = nudt / (2.0 * a.dv);
    const double dA = hb * a.vstep;
    const double As0 = fma(hb, fma((double)(s - 1), a.vstep, a.v0) - vbar, nudt * tdv);   // A_s
    const double Cs0 = fma(-hb, fma((double)(s + 1), a.vstep, a.v0) - vbar, nudt * tdv);  // C_s
    const double dAn = dA * sc, As0n = As0 * sc, Cs0n = Cs0 * sc;                         // exact
#define CA(i) fma(dA, (double)(i), As0)        /* A_{s+i} */
#define CC(i) fma(-dA, (double)(i), Cs0)       /* C_{s+i} */
#define CAN(i) fma(dAn, (double)(i), As0n)     /* a'_{s+i} = sc A_{s+i} */
#define CCN(i) fma(-dAn, (double)(i), Cs0n)    /* c'_{s+i} = sc C_{s+i} */
    constexpr int L = MI - 1;                  // last interior cell
    // The coefficients are pure functions of the cell index: left alone, the scheduler computes them
    // far ahead of the recurrences that consume them and spills the chunk.  TIE makes the first
    // coefficient of every block of TB cells depend (exactly: + 0 * x) on the running determinant.
#ifndef FPREG_TB
#define FPREG_TB 4
#endif
    constexpr int TB = FPREG_TB;
#define TIE(val, dep) fma((dep), zero, (val))

    // ---------------- chunk interior [0, MI) -> six spike end values; both sweeps in one loop
    double u_first, u_last, w_first, w_last, y_first, y_last;
    {
      // the staged chunk is read 16 bytes (two cells) at a time, one pair ahead of its use, with
      // loads that keep their place in the instruction stream (see ld2_fresh)
      double2 lp = ld2_fresh(sp), up = ld2_fresh(sp + L);
      double2 lq = ld2_fresh(sp + 2), uq = ld2_fresh(sp + L - 2);
      double n2 = 1.0, n1 = bt, z = lp.x, g = 1.0;        // LU sweep down: N_{k-2}, N_{k-1}, Z_{k-1}, G_{k-1}
      double m2 = 1.0, m1 = bt, h = 1.0, tt = up.x;       // UL sweep up:   NU_{i+2}, NU_{i+1}, T_{i+1}, H_{i+1}
      static_assert((L & 1) == 0, "pairs");
#ifndef FPREG_FUSE
#define FPREG_FUSE 1
#endif
// Z_i = d_i N_{i-1} - a_i Z_{i-1}: with the product d_i N_{i-1} formed first (off the chain: N comes from its own
// recurrence) the recurrence is ONE dependent FMA per cell; fma(d, N, -(a Z)) puts a multiply and an FMA on the chain
// (the forward sweep waited on fixed-latency dependencies for 46 % of its samples).  Two roundings either way.
// Used in the forward sweep only: the spike sweeps are fp64-pipe bound, and there the early products lengthen live
// ranges until the chunk spills (376 bytes).
#ifndef FPREG_ZFMA
#define FPREG_ZFMA 1
#endif
#if FPREG_ZFMA
#define FPREG_ZSTEP(d, n, a_, zprev) fma(-(a_), (zprev), (d) * (n))
#else
#define FPREG_ZSTEP(d, n, a_, zprev) fma((d), (n), -((a_) * (zprev)))
#endif
// FPREG_ZFMA_SPIKE (candidate for the next A/B, off): the same one-FMA recurrence in the two spike sweeps, coefficients
// tied to the right-hand-side chains z / tt instead of the determinant chains.
#ifndef FPREG_ZFMA_SPIKE
#define FPREG_ZFMA_SPIKE 0
#endif
#if FPREG_ZFMA_SPIKE
#define FPREG_SPIKE_ZSTEP(d, n, a_, zprev) fma(-(a_), (zprev), (d) * (n))
#define FPREG_SPIKE_TIE(nchain, zchain) (zchain)
#else
#define FPREG_SPIKE_ZSTEP(d, n, a_, zprev) fma((d), (n), -((a_) * (zprev)))
#define FPREG_SPIKE_TIE(nchain, zchain) (nchain)
#endif
#define LU_STEP(k)                                                                        \
  {                                                                                       \
    const int o = ((k) - 1) % TB;                                                         \
    if (o == 0) { ab = TIE(CAN(k), FPREG_SPIKE_TIE(n1, z)); cb = TIE(CCN((k) - 1), FPREG_SPIKE_TIE(n1, z)); } \
    const double ak = (o == 0) ? ab : fma(dAn, (double)o, ab);  /* a'_k */                \
    const double ck = (o == 0) ? cb : fma(-dAn, (double)o, cb); /* c'_{k-1} */            \
    const double nk = fma(bt, n1, -((ak * ck) * n2));                                     \
    if (((k) & 1) == 0) { lp = lq; if ((k) + 2 <= L) lq = ld2_fresh(sp + (k) + 2); }      \
    z = FPREG_SPIKE_ZSTEP(((k) & 1) ? lp.y : lp.x, n1, ak, z);                             \
    g = -ak * g;                                                                          \
    n2 = n1; n1 = nk;                                                                     \
  }
#define UL_STEP(k)                                                                        \
  {                                                                                       \
    const int o = ((k) - 1) % TB;                                                         \
    const int i = L - (k);                                                                \
    if (o == 0) { ub = TIE(CCN(i), FPREG_SPIKE_TIE(m1, tt)); vb = TIE(CAN(i + 1), FPREG_SPIKE_TIE(m1, tt)); } \
    const double ci = (o == 0) ? ub : fma(dAn, (double)o, ub);  /* c'_i */                \
    const double ai = (o == 0) ? vb : fma(-dAn, (double)o, vb); /* a'_{i+1} */            \
    const double mi = fma(bt, m1, -((ai * ci) * m2));                                     \
    if ((i & 1) == 1) { up = uq; if (i - 3 >= 0) uq = ld2_fresh(sp + i - 3); }            \
    tt = FPREG_SPIKE_ZSTEP((i & 1) ? up.y : up.x, m1, ci, tt);                             \
    h = -ci * h;                                                                          \
    m2 = m1; m1 = mi;                                                                     \
  }
      double ab = 0.0, cb = 0.0, ub = 0.0, vb = 0.0;      // coefficient bases of the current block of TB cells
#if FPREG_FUSE
#pragma unroll
      for (int k = 1; k <= L; ++k) {
        LU_STEP(k)
        UL_STEP(k)
      }
#else
#pragma unroll
      for (int k = 1; k <= L; ++k) LU_STEP(k)
#pragma unroll
      for (int k = 1; k <= L; ++k) UL_STEP(k)
#endif
#undef LU_STEP
#undef UL_STEP
      const double rl = sc * rcp_fast(n1);                // sc / det, det = N_L
      const double ru = sc * rcp_fast(m1);                // the same determinant from the other side
      y_last = z * rl; w_last = n2 * rl; u_last = g * rl;
      u_first = m2 * ru; w_first = h * ru; y_first = tt * ru;
    }
    // ---------------- separator equation of this chunk (needs the next chunk's first-spikes)
    const double As = (t > 0) ? CA(0) : 0.0;
    const double Ce1 = CC(MI - 1);
    const double Ae = CA(MI);
    const double Ce = (t < T - 1) ? CC(MI) : 0.0;
    // (X was last read by the previous row's moment reduction / transposition, i.e. before the barrier at the top
    // of this row: no barrier needed before it is written)
    X[t] = u_first; X[T + t] = w_first; X[2 * T + t] = y_first;
    __syncthreads();
    double ra, rb, rc, rd;
    {
      ra = -Ae * As * u_last;
      rb = fma(-Ae * Ce1, w_last, bd);
      rc = 0.0;
      rd = fma(-Ae, y_last, sp[M - 1]);
      if (t < T - 1) {
        rb = fma(-Ce * CA(M), X[t + 1], rb);              // A of the next chunk's first row
        rc = -Ce * CC(M + MI - 1) * X[T + t + 1];         // C of the next chunk's last interior row
        rd = fma(-Ce, X[2 * T + t + 1], rd);
      }
    }
    // ---------------- cyclic reduction over the T separators (see fp_fast.cuh)
    {
      double* cur = X;
      double* nxt = X + 4 * T;
      {
        const double ib = 1.0 / rb;
        ra *= ib; rc *= ib; rd *= ib;
      }
      __syncthreads();
      cur[t] = ra; cur[T + t] = rc; cur[2 * T + t] = rd;
      int more = __syncthreads_or((fabs(ra) + fabs(rc)) > 1e-18);
#pragma unroll 1
      for (int st = 1; st < T && more; st <<= 1) {
        double nb = 1.0, na = 0.0, nc = 0.0;
        const int im = t - st, ip = t + st;
        if (im >= 0) {
          na = -ra * cur[im];
          nb = fma(-ra, cur[T + im], nb);
          rd = fma(-ra, cur[2 * T + im], rd);
        }
        if (ip < T) {
          nc = -rc * cur[T + ip];
          nb = fma(-rc, cur[ip], nb);
          rd = fma(-rc, cur[2 * T + ip], rd);
        }
        const double ib = 1.0 / nb;
        ra = na * ib; rc = nc * ib; rd *= ib;
        nxt[t] = ra; nxt[T + t] = rc; nxt[2 * T + t] = rd;
        more = __syncthreads_or((fabs(ra) + fabs(rc)) > 1e-18);
        double* tmp = cur; cur = nxt; nxt = tmp;
      }
      __syncthreads();
      X[t] = rd;
      __syncthreads();
    }
    // ---------------- interior with known neighbours.  The forward sweep reads the staged chunk (pairs,
    // one ahead, like the spike sweeps) and leaves the eliminated right-hand sides Z_i in REGISTERS:
    // from here on the chunk lives in c[], and once every thread is through, the staging buffer
    // receives the CTA's next row.
    // The back-substitution  x_i = (Z_i - c'_i N_{i-1} x_{i+1}) / N_i  is split into a PREPARE step per
    // cell, off the critical path (R_i = 1/N_i, c[i] <- Z_i R_i, E_i = c'_i N_{i-1} R_i parked in shared
    // memory), and a one-FMA recurrence  x_i = c[i] - E_i x_{i+1}.  Three groups of cells
    // [0,G1) [G1,G2) [G2,MI): the determinants of the top group are parked by the forward sweep, those of
    // the two lower groups are regenerated from a two-value checkpoint.  The cp.async of the next row
    // are issued one per prepared cell (a burst of 16 stalls the load/store queue).
    double c[M];
    const double xe = X[t];
    {
      const double xl = (t > 0) ? X[t - 1] : 0.0;
      double2 lp = ld2_fresh(sp), lq = ld2_fresh(sp + 2);
      double z = sc * fma(-As, xl, lp.x);                 // right-hand side scaled like the matrix
      c[0] = z;
      double n2 = 1.0, n1 = bt, k2 = 1.0, k1 = bt;        // N_{i-2}, N_{i-1}; checkpoint N_{G1-2}, N_{G1-1}
      double ab = 0.0, cb = 0.0;
#pragma unroll
      for (int i = 1; i <= L; ++i) {
        if ((i & 1) == 0) { lp = lq; if (i + 2 <= L) lq = ld2_fresh(sp + i + 2); }
        const int o = (i - 1) % TB;
        // (with the one-FMA Z recurrence the coefficients are tied to z: the determinant chain must not run ahead of it)
        if (o == 0) { ab = TIE(CAN(i), FPREG_ZFMA ? z : n1); cb = TIE(CCN(i - 1), FPREG_ZFMA ? z : n1); }
        const double ai = (o == 0) ? ab : fma(dAn, (double)o, ab);            // a'_i
        const double ck = (o == 0) ? cb : fma(-dAn, (double)o, cb);           // c'_{i-1}
        const double ni = fma(bt, n1, -((ai * ck) * n2));
        double di = (i & 1) ? lp.y : lp.x;
        if (i == L) di = fma(-Ce1, xe, di);
        z = FPREG_ZSTEP(di, sc * n1, ai, z);
        c[i] = z;
        if (i == G2) PV[t] = n1;                          // N_{G2-1}
        if (i >= G2) PV[(i - G2 + 1) * T + t] = ni;       // slot j of the top group: N_{G2-1+j}
        if (i == G1 - 1) { k2 = n1; k1 = ni; }
        n2 = n1; n1 = ni;
      }
      __syncthreads();                                    // every thread is done with the staged row
      const long rn = r + gridDim.x;
      const bool has_next = rn < a.rows;
      const double* nsrc = a.fin + (has_next ? rn : r) * a.ld_in + 2 * t;
      const bool spread = has_next;
      int pf = 0;                                         // compile-time after unrolling: cp.async issued so far
#define PF_ONE()                                                                                   \
  if (pf < U) {                                                                                    \
    if (spread) cp_async16(stage_t + pf * (T / U) * PITCH, nsrc + pf * 2 * T);                     \
    ++pf;                                                                                          \
  }
#pragma unroll
      for (int i = G2; i <= L; ++i) {                     // prepare the top group from its parked determinants
        PF_ONE()
        const double R = rcp_fast(PV[(i - G2 + 1) * T + t]);
        const double e = (CCN(i) * PV[(i - G2) * T + t]) * R;
        c[i] *= R;
        if (i < L) PV[(i - G2) * T + t] = e;
      }
      double x = c[L];
#pragma unroll
      for (int i = L - 1; i >= G2; --i) {
        x = fma(-PV[(i - G2) * T + t], x, c[i]);
        c[i] = x;
      }
      n2 = k2; n1 = k1;                                   // regenerate N_{G1} .. N_{G2-1}
#pragma unroll
      for (int i = G1; i < G2; ++i) {
        PF_ONE()
        const double ni = fma(bt, n1, -((TIE(CAN(i), n1) * CCN(i - 1)) * n2));
        const double R = rcp_fast(ni);
        c[i] *= R;
        PV[(i - G1) * T + t] = (CCN(i) * n1) * R;
        n2 = n1; n1 = ni;
      }
#pragma unroll
      for (int i = G2 - 1; i >= G1; --i) {
        x = fma(-PV[(i - G1) * T + t], x, c[i]);
        c[i] = x;
      }
      {                                                   // cell 0: N_{-1} = 1, N_0 = b'
        const double R = rcp_fast(bt);
        c[0] *= R;
        PV[t] = CCN(0) * R;
      }
      n2 = 1.0; n1 = bt;                                  // regenerate N_1 .. N_{G1-1}
#pragma unroll
      for (int i = 1; i < G1; ++i) {
        PF_ONE()
        const double ni = fma(bt, n1, -((TIE(CAN(i), n1) * CCN(i - 1)) * n2));
        const double R = rcp_fast(ni);
        c[i] *= R;
        PV[i * T + t] = (CCN(i) * n1) * R;
        n2 = n1; n1 = ni;
      }
      static_assert(MI - 1 >= U, "one cp.async per prepared cell covers the row");
#undef PF_ONE
      cp_async_commit();
#pragma unroll
      for (int i = G1 - 1; i >= 0; --i) {
        x = fma(-PV[i * T + t], x, c[i]);
        c[i] = x;
      }
      c[M - 1] = xe;
    }
#undef CA
#undef CC
#undef CAN
#undef TIE
#undef CCN
    // ---------------- store.  A warp owns 32 consecutive chunks = one contiguous piece of the row, but a
    // thread's chunk is 256 bytes away from its neighbour's: stored straight from registers, every
    // 16-byte store instruction would touch 32 different 128-byte lines.  The warp transposes 16 cells
    // per lane at a time through its own slice of the (now idle) scratch area and writes whole lines.
    __syncthreads();                                      // every warp is done with X and its parked values
    // The moment sums of the stored cells are interleaved with the stores (FPREG_INTERLEAVE): transposition and
    // stores are load/store-queue bound (mio / lg throttle: 13 % of the kernel in the ncu source view), the
    // sums are fp64 bound, and between the two __syncwarp the scheduler is free to mix them.
    ChunkSums S;
#pragma unroll
    for (int q = 0; q < 6; ++q) S.mu[q] = 0.0;
    S.s2 = 0.0; S.sl = 0.0; S.se = 0.0; S.bad = 0;
#ifndef FPREG_INTERLEAVE
#define FPREG_INTERLEAVE 1
#endif
    const bool mom_inline = FPREG_INTERLEAVE && (a.mom_out != nullptr);
    {
      double* wbuf = X + warp * (32 * WP);
      double* drow = a.fout + r * a.ld_out + (long)warp * 32 * M;
#pragma unroll
      for (int hh = 0; hh < M / 16; ++hh) {
#pragma unroll
        for (int j = 0; j < 8; ++j) store2(wbuf + lane * WP + 2 * j, c[16 * hh + 2 * j], c[16 * hh + 2 * j + 1]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int pc = k * 4 + (lane >> 3), o = (lane & 7) * 2;     // lane-row (chunk of the warp), offset
          const double2 v2 = ld2(wbuf + pc * WP + o);
          store2(drow + pc * M + 16 * hh + o, v2.x, v2.y);
          if (mom_inline) {
            chunk_terms_at<M>(S, c[16 * hh + 2 * k], LTl, 16 * hh + 2 * k);
            chunk_terms_at<M>(S, c[16 * hh + 2 * k + 1], LTl, 16 * hh + 2 * k + 1);
          }
        }
        __syncwarp();
      }
    }
    if (a.mom_out) {
      // v^p moments from local monomial sums: v_i = vc + (i - I0) vstep, so
      //   sum_i x_i v_i^p = sum_q C(p,q) vc^(p-q) vstep^q mu_q        (6 FMAs per cell instead of 11 flops)
      if (!FPREG_INTERLEAVE) ChunkLoop<M, 0>::run(S, c, LTl);
      if (first_thread) chunk_terms<M, 0, 1>(S, c[0], LTl);             // np.trapz: half weight at both ends
      if (last_thread) chunk_terms<M, M - 1, 1>(S, c[M - 1], LTl);
      if (S.bad) {
        double sl = 0.0;
        ChunkLoop<M, 0>::rare(sl, c);
        if (first_thread) sl = fma(-0.5 * c[0], fpfast::log_rare(c[0]), sl);
        if (last_thread) sl = fma(-0.5 * c[M - 1], fpfast::log_rare(c[M - 1]), sl);
        S.sl = sl; S.se = 0.0;
      }
      const double dl = a.vstep, vc = fma(0.5 * (double)(M - 1), dl, vs);
      const double dl2 = dl * dl;
      const double n0 = S.mu[0], n1 = S.mu[1] * dl, n2 = S.mu[2] * dl2, n3 = S.mu[3] * (dl2 * dl),
                   n4 = S.mu[4] * (dl2 * dl2), n5 = S.mu[5] * (dl2 * dl2 * dl);
      double acc[8];
      acc[0] = n0;
      acc[1] = fma(vc, n0, n1);
      acc[2] = fma(vc, fma(vc, n0, 2.0 * n1), n2);
      acc[3] = fma(vc, fma(vc, fma(vc, n0, 3.0 * n1), 3.0 * n2), n3);
      acc[4] = fma(vc, fma(vc, fma(vc, fma(vc, n0, 4.0 * n1), 6.0 * n2), 4.0 * n3), n4);
      acc[5] = fma(vc, fma(vc, fma(vc, fma(vc, fma(vc, n0, 5.0 * n1), 10.0 * n2), 10.0 * n3), 5.0 * n4), n5);
      acc[6] = S.s2;
      acc[7] = fma(S.se, 6.93147180369123816490e-01, fma(S.se, 1.90821492927058770002e-10, S.sl));
      // reduction over the CTA: the eight sums of every thread through shared memory (warp w adds
      // sum w in a fixed order: 8 x 10 shuffles per CTA instead of 80 per thread)
      __syncthreads();                                    // the transposition slices are idle
#pragma unroll
      for (int k = 0; k < 8; ++k) X[k * T + t] = acc[k] * a.dv;
      __syncthreads();
      for (int k = warp; k < 8; k += NW) {
        double y = 0.0;
#pragma unroll
        for (int j = 0; j < T / 32; ++j) y += X[k * T + lane + 32 * j];
        y = warp_sum(y);
        if (lane == 0) a.mom_out[(long)k * a.mom_ld + r] = y;
      }
    }
    // (no barrier here: the next row starts with cp.async.wait + __syncthreads before anything is written)
  }
}

}  // namespace fpreg
