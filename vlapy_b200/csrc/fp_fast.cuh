// fp_fast.cuh -- implicit Fokker-Planck step, one CTA per x-row, tuned version of FpProg
// (rowops.h; same algorithm: chunk interiors eliminated exactly, tridiagonal separator system
// solved by cyclic reduction, interiors back-substituted).  Replaces
// vlapy/core/collisions.py:44-81, 104-158, 232-263 via vlapy/core/step.py:102-108.
//
//  * the row sits in shared memory with one padding word per chunk, so both the coalesced
//    row-order accesses and the per-thread chunk sweeps (stride M+1 doubles) are conflict free;
//  * the reduction sweeps keep only running scalars; the back-substitution keeps the pivots of
//    the upper half of a thread's chunk in registers and recomputes the lower half's;
//    reciprocals of consecutive pivots are refined from one another by two Newton steps;
//  * diagonals are affine in the cell index (the velocity grid is np.linspace: v_i = v0 + i*step,
//    checked bit-for-bit on the host before this kernel is chosen): A_i = a0 + a1 i,
//    C_i = c0 + c1 i, one FMA each, nothing is loaded;
//  * reductions use warp shuffles + one shared-memory hop; the separator system (one unknown per
//    thread) is reduced by shuffle-PCR inside each warp and shared-memory PCR across warps.
#pragma once
#include "vpfp_common.h"

namespace fpfast {

struct Args {
  const double* fin; long ld_in;
  double* fout; long ld_out;
  double v0, vstep, vlast;   // v_i = v0 + i*vstep (i < nv-1), v_{nv-1} = vlast  (np.linspace)
  double nu, dt, dv;
  int op;                    // 0 lb, 1 dg
  double* mom_out; long mom_ld;
  int rows, nv;
  const double2* logtab;     // 128 x (1/c_i, ln c_i), c_i = 1 + (i + 1/2)/128
  const double2* logtab64;   // 64 x (1/c_i, ln c_i), c_i = 1 + (i + 1/2)/64 (fp_reg.cuh)
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// sum over the CTA; every thread gets the result. red: >= 32 doubles of shared scratch.
// PROTECT = false: the caller knows that a barrier already lies between the previous use of red and this call.
template <int T, bool PROTECT = true>
__device__ __forceinline__ double block_sum(double x, double* red) {
  constexpr int NW = T / 32;
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (PROTECT) __syncthreads();  // protect red from the previous use
  if (l == 0) red[w] = x;
  __syncthreads();
  double y = (l < NW) ? red[l] : 0.0;
  return warp_sum(y);
}

// reciprocal of p given the reciprocal r of a nearby value (the pivots of a diagonally dominant
// tridiagonal matrix with slowly varying coefficients converge geometrically): two Newton steps
// when the first residual is below 2^-14 (then the result is good to < 1 ulp), else a division.
__device__ __forceinline__ double rcp_near(double p, double r) {
  const double e = fma(-p, r, 1.0);                    // 1/p = r (1 + e + e^2 + ...)
  if (fabs(e) < 3.0e-6) return fma(r, fma(e, e, e), r);  // truncation e^3 < 3e-17
  return 1.0 / p;
}

// natural logarithm for the f ln f moment: x = 2^e m, m in [1,2) split by its top 7 mantissa bits
// into c_i (1 + r); ln x = e ln2 + ln c_i + log1p(r), |r| < 2^-8, degree-6 series.  Absolute error
// ~1e-16 (what a sum of f ln f needs); non-positive, subnormal or non-finite arguments take the
// library path so that NaN / -inf semantics match numpy (vlapy/core/step.py:222-224).
#if defined(__CUDACC__)
__device__ __noinline__ double log_rare(double x) { return log(x); }
#else
inline double log_rare(double x) { return log(x); }
#endif
__device__ __forceinline__ double log_sum(double x, const double2* __restrict__ tab) {
  const long long bits = __double_as_longlong(x);
  const int ex = (int)((bits >> 52) & 0x7ff);
  if (bits <= 0 || ex == 0 || ex == 0x7ff) return log_rare(x);   // out of line: 32 inlined copies bloat the kernel
  const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const int idx = (int)((bits >> 45) & 127);
  const double2 t = tab[idx];                          // (1/c_i, ln c_i)
  const double r = fma(m, t.x, -1.0);
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(r, p, -0.25);
  p = fma(r, p, 1.0 / 3.0);
  p = fma(r, p, -0.5);
  p = fma(r * r, p, r);
  const double e = (double)(ex - 1023);
  return fma(e, 6.93147180369123816490e-01, fma(e, 1.90821492927058770002e-10, t.y + p));
}

template <int M, int T>
__global__ void __launch_bounds__(T, 1) fp_kernel(const Args a) {
  constexpr int NV = M * T;
  constexpr int NW = T / 32;
  constexpr int MI = M - 1;          // interior cells of a chunk
  constexpr int H = MI / 2;          // pivots of cells [H, MI) are kept, [0, H) are recomputed
  VPFP_DYN_SMEM(smem_raw);
  double* row = reinterpret_cast<double*>(smem_raw);  // NV + T (one pad per chunk)
  double* red = row + NV + T;                          // 64
  double* X = red + 64;                                // 8 * T scratch (separator system / moments)
  double2* LT = reinterpret_cast<double2*>(X + 8 * T); // 128 entries of the log table
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  for (int i = t; i < 128; i += T) LT[i] = a.logtab[i];
  auto sk = [](int i) { return i + i / M; };           // skewed index
  // velocity of cell t + k*T (row-order loops) -- affine in k; the reference grid is np.linspace
  const double vt = fma((double)t, a.vstep, a.v0);
  const double hdv = 0.5 * a.dv;

  for (long r = blockIdx.x; r < a.rows; r += gridDim.x) {
    // ---------------- load + first moment
    const double* src = a.fin + r * a.ld_in;
    double acc0 = 0.0;
    const double vinc = (double)T * a.vstep;
    {
      double vi = vt;
#pragma unroll 8
      for (int k = 0; k < M; ++k) {
        const int i = t + k * T;
        const double fv = src[i];
        row[sk(i)] = fv;
        const double w = ((k == 0 && t == 0) || (k == M - 1 && t == T - 1)) ? hdv : a.dv;
        acc0 += (a.op == 0) ? w * fv * vi * vi : w * fv * vi;
        vi += vinc;
      }
    }
    const double first = block_sum<T>(acc0, red);   // (contains the syncs that publish `row`)
    double Tm = first, vbar = 0.0;
    if (a.op == 1) {
      vbar = first;
      double acc1 = 0.0;
      double vi = vt;
#pragma unroll 8
      for (int k = 0; k < M; ++k) {
        const int i = t + k * T;
        const double d = vi - vbar;
        const double w = ((k == 0 && t == 0) || (k == M - 1 && t == T - 1)) ? hdv : a.dv;
        acc1 += w * row[sk(i)] * d * d;
        vi += vinc;
      }
      Tm = block_sum<T>(acc1, red);
    }
    // diagonals, affine in the cell index i:
    //   A_i = nudt(tdv + (v_{i-1} - vbar)/2/dv) = A0 + i dA,  C_i = nudt(tdv - (v_{i+1} - vbar)/2/dv) = C0 - i dA
    const double nudt = a.nu * a.dt;
    const double tdv = -Tm / (a.dv * a.dv);
    const double bd = 1.0 + nudt * (2.0 * Tm / (a.dv * a.dv));
    const double rbd = 1.0 / bd;
    const double hb = nudt / (2.0 * a.dv);
    const double dA = hb * a.vstep;
    const int s = t * M, e = s + M - 1;
    const double As0 = fma(hb, fma((double)(s - 1), a.vstep, a.v0) - vbar, nudt * tdv);   // A_s
    const double Cs0 = fma(-hb, fma((double)(s + 1), a.vstep, a.v0) - vbar, nudt * tdv);  // C_s
#define CA(i) fma(dA, (double)(i), As0)       /* A_{s+i} */
#define CC(i) fma(-dA, (double)(i), Cs0)      /* C_{s+i} */

    // ---------------- chunk interior [s, e-1] -> six spike end values (no per-cell storage)
    double u_first, u_last, w_first, w_last, y_first, y_last;
    {
      double rq = rbd, tt = row[sk(e - 1)], h = 1.0;      // UL sweep up
      {
        double cc = CC(MI - 2), ca = CA(MI - 1);          // C_{s+i}, A_{s+i+1}, running down in i
#pragma unroll 4
        for (int i = MI - 2; i >= 0; --i) {
          const double rr = cc * rq;
          rq = rcp_near(fma(-rr, ca, bd), rq);
          tt = fma(-rr, tt, row[sk(s + i)]);
          h = -rr * h;
          cc += dA; ca -= dA;
        }
      }
      u_first = rq; w_first = h * rq; y_first = tt * rq;
      double rp = rbd, z = row[sk(s)], g = 1.0;           // LU sweep down
      {
        double ca = CA(1), cc = CC(0);                    // A_{s+i}, C_{s+i-1}, running up in i
#pragma unroll 4
        for (int i = 1; i <= MI - 1; ++i) {
          const double l = ca * rp;
          rp = rcp_near(fma(-l, cc, bd), rp);
          z = fma(-l, z, row[sk(s + i)]);
          g = -l * g;
          ca += dA; cc -= dA;
        }
      }
      y_last = z * rp; w_last = rp; u_last = g * rp;
    }
    // ---------------- separator equation of this chunk (needs the next chunk's first-spikes)
    const double As = (t > 0) ? CA(0) : 0.0;
    const double Ce1 = CC(MI - 1);
    const double Ae = CA(MI);
    const double Ce = (t < T - 1) ? CC(MI) : 0.0;
    __syncthreads();
    X[t] = u_first; X[T + t] = w_first; X[2 * T + t] = y_first;
    __syncthreads();
    double ra, rb, rc, rd;
    {
      ra = -Ae * As * u_last;
      rb = fma(-Ae * Ce1, w_last, bd);
      rc = 0.0;
      rd = fma(-Ae, y_last, row[sk(e)]);
      if (t < T - 1) {
        rb = fma(-Ce * CA(M), X[t + 1], rb);              // A of the next chunk's first row
        rc = -Ce * CC(M + MI - 1) * X[T + t + 1];         // C of the next chunk's last interior row
        rd = fma(-Ce, X[2 * T + t + 1], rd);
      }
    }
    // ---------------- cyclic reduction over the T separators (shared memory ping-pong).
    // Rows are kept normalised (unit diagonal: a, c, d divided by b), so a step needs one
    // reciprocal.  The matrix is diagonally dominant, the couplings shrink quadratically per
    // step; once every |a| + |c| is below 1e-18 the rows are decoupled and x = d.
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
    // ---------------- interior with known neighbours, in place.  Pivots of the upper half of the
    // chunk are kept in registers during the sweep down; the lower half's are recomputed.
    {
      const double xe = X[t];
      const double xl = (t > 0) ? X[t - 1] : 0.0;
      double rpv[MI - H];
      double z = fma(-As, xl, row[sk(s)]);
      if (MI == 1) z = fma(-Ce1, xe, z);
      row[sk(s)] = z;
      double rp = rbd;
      if (H == 0) rpv[0] = rp;
#pragma unroll
      for (int i = 1; i <= MI - 1; ++i) {
        const double l = CA(i) * rp;
        rp = rcp_near(fma(-l, CC(i - 1), bd), rp);
        double di = row[sk(s + i)];
        if (i == MI - 1) di = fma(-Ce1, xe, di);
        z = fma(-l, z, di);
        row[sk(s + i)] = z;
        if (i >= H) rpv[i - H] = rp;
        if ((i & 3) == 3) asm volatile("" ::: "memory");
      }
      double x = z * rp;
      row[sk(s + MI - 1)] = x;
#pragma unroll
      for (int i = MI - 2; i >= H; --i) {
        x = fma(-CC(i), x, row[sk(s + i)]) * rpv[i - H];
        row[sk(s + i)] = x;
        if ((i & 3) == 0) asm volatile("" ::: "memory");
      }
      if (H > 0) {
        rp = rbd;
        rpv[0] = rp;
#pragma unroll
        for (int i = 1; i <= H - 1; ++i) {
          const double l = CA(i) * rp;
          rp = rcp_near(fma(-l, CC(i - 1), bd), rp);
          rpv[i] = rp;
        }
#pragma unroll
        for (int i = H - 1; i >= 0; --i) {
          x = fma(-CC(i), x, row[sk(s + i)]) * rpv[i];
          row[sk(s + i)] = x;
          if ((i & 3) == 0) asm volatile("" ::: "memory");
        }
      }
      row[sk(e)] = xe;
    }
#undef CA
#undef CC
    __syncthreads();
    // ---------------- store + moments of the new row
    double* dst = a.fout + r * a.ld_out;
    if (a.mom_out) {
      double acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.0;
      double vi = vt;
#pragma unroll 4
      for (int k = 0; k < M; ++k, vi += vinc) {
        const int i = t + k * T;
        const double x = row[sk(i)];
        dst[i] = x;
        const double w = ((k == 0 && t == 0) || (k == M - 1 && t == T - 1)) ? hdv : a.dv;
        const double tw = w * x;
        acc[0] += tw;
        double p = tw * vi; acc[1] += p;
        p *= vi; acc[2] += p;
        p *= vi; acc[3] += p;
        p *= vi; acc[4] += p;
        p *= vi; acc[5] += p;
        acc[6] = fma(tw, x, acc[6]);
        acc[7] = fma(tw, log_sum(x, LT), acc[7]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = warp_sum(acc[k]);
      __syncthreads();
      if (lane == 0)
#pragma unroll
        for (int k = 0; k < 8; ++k) X[k * NW + warp] = acc[k];
      __syncthreads();
      for (int k = warp; k < 8; k += NW) {
        double y = (lane < NW) ? X[k * NW + lane] : 0.0;
        y = warp_sum(y);
        if (lane == 0) a.mom_out[(long)k * a.mom_ld + r] = y;
      }
    } else {
#pragma unroll 8
      for (int k = 0; k < M; ++k) {
        const int i = t + k * T;
        dst[i] = row[sk(i)];
      }
    }
    __syncthreads();
  }
}

template <int M, int T>
constexpr size_t smem_bytes() {
  return sizeof(double) * (size_t)(M * T + T + 64 + 8 * T + 256);
}

}  // namespace fpfast
