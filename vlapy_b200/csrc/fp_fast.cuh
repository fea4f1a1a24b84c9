// fp_fast.cuh -- implicit Fokker-Planck step, one CTA per x-row, tuned version of FpProg
// (rowops.h; same algorithm: chunk interiors eliminated exactly, tridiagonal separator system
// solved by cyclic reduction, interiors back-substituted).  Replaces
// vlapy/core/collisions.py:44-81, 104-158, 232-263 via vlapy/core/step.py:102-108.
//
//  * the row sits in shared memory with one padding word per chunk, so both the coalesced
//    row-order accesses and the per-thread chunk sweeps (stride M+1 doubles) are conflict free;
//  * the LU pivots 1/p_i of a thread's chunk stay in registers between the reduction sweep and the
//    back-substitution: one reciprocal per element per sweep direction, none in the final solve;
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
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// sum over the CTA; every thread gets the result. red: >= 32 doubles of shared scratch.
template <int T>
__device__ __forceinline__ double block_sum(double x, double* red) {
  constexpr int NW = T / 32;
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect red from the previous use
  if (l == 0) red[w] = x;
  __syncthreads();
  double y = (l < NW) ? red[l] : 0.0;
  return warp_sum(y);
}

template <int M, int T>
__global__ void __launch_bounds__(T, 1) fp_kernel(const Args a) {
  constexpr int NV = M * T;
  constexpr int NW = T / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* row = reinterpret_cast<double*>(smem_raw);  // NV + T (one pad per chunk)
  double* red = row + NV + T;                          // 64
  double* X = red + 64;                                // 8 * T scratch (separator system / moments)
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  auto sk = [](int i) { return i + i / M; };           // skewed index
  auto vel = [&](int i) { return (i == NV - 1) ? a.vlast : __dadd_rn(a.v0, __dmul_rn((double)i, a.vstep)); };
  auto wgt = [&](int i) { return (i == 0 || i == NV - 1) ? 0.5 * a.dv : a.dv; };

  for (long r = blockIdx.x; r < a.rows; r += gridDim.x) {
    // ---------------- load + first moment
    const double* src = a.fin + r * a.ld_in;
    double acc0 = 0.0;
#pragma unroll 4
    for (int k = 0; k < M; ++k) {
      const int i = t + k * T;
      const double fv = src[i];
      row[sk(i)] = fv;
      const double vi = vel(i);
      acc0 += (a.op == 0) ? wgt(i) * fv * vi * vi : wgt(i) * fv * vi;
    }
    const double first = block_sum<T>(acc0, red);   // (contains the syncs that publish `row`)
    double Tm = first, vbar = 0.0;
    if (a.op == 1) {
      vbar = first;
      double acc1 = 0.0;
#pragma unroll 4
      for (int k = 0; k < M; ++k) {
        const int i = t + k * T;
        const double d = vel(i) - vbar;
        acc1 += wgt(i) * row[sk(i)] * d * d;
      }
      Tm = block_sum<T>(acc1, red);
    }
    // diagonals: A_i = nudt(tdv + (v_{i-1} - vbar)/2/dv), C_i = nudt(tdv - (v_{i+1} - vbar)/2/dv)
    const double nudt = a.nu * a.dt;
    const double tdv = -Tm / (a.dv * a.dv);
    const double bd = 1.0 + nudt * (2.0 * Tm / (a.dv * a.dv));
    const double hb = nudt / (2.0 * a.dv);
    auto cA = [&](int i) { return nudt * tdv + hb * (vel(i - 1) - vbar); };
    auto cC = [&](int i) { return nudt * tdv - hb * (vel(i + 1) - vbar); };

    // ---------------- chunk interior [s, e-1], separator e
    const int s = t * M, e = s + M - 1;
    double rpv[M];                    // LU pivots' reciprocals of the interior
    double u_first, u_last, w_first, w_last, y_first, y_last;
    {
      // UL sweep up (reads the untouched right-hand side)
      double rq = 1.0 / bd, tt = row[sk(e - 1)], h = 1.0;
#pragma unroll
      for (int i = M - 3; i >= 0; --i) {
        const double rr = cC(s + i) * rq;
        rq = 1.0 / (bd - rr * cA(s + i + 1));
        tt = row[sk(s + i)] - rr * tt;
        h = -rr * h;
      }
      u_first = rq; w_first = h * rq; y_first = tt * rq;
      // LU sweep down
      double rp = 1.0 / bd, z = row[sk(s)], g = 1.0;
      rpv[0] = rp;
#pragma unroll
      for (int i = 1; i <= M - 2; ++i) {
        const double l = cA(s + i) * rp;
        rp = 1.0 / (bd - l * cC(s + i - 1));
        z = row[sk(s + i)] - l * z;
        g = -l * g;
        rpv[i] = rp;
      }
      y_last = z * rp; w_last = rp; u_last = g * rp;
    }
    // ---------------- separator equation of this chunk (needs the next chunk's first-spikes)
    const double As = (t > 0) ? cA(s) : 0.0;
    const double Ce1 = cC(e - 1);
    const double Ae = cA(e);
    const double Ce = (t < T - 1) ? cC(e) : 0.0;
    // neighbour chunk j+1: u_first, w_first, y_first, and its As, Ce1
    __syncthreads();
    X[t] = u_first; X[T + t] = w_first; X[2 * T + t] = y_first;
    __syncthreads();
    double ra, rb, rc, rd;
    {
      const double As2 = cA(s + M);                    // A of next chunk's first row (= cA(e+1))
      const double Ce12 = cC(e + M - 1);               // C of next chunk's last interior row
      ra = -Ae * As * u_last;
      rb = bd - Ae * Ce1 * w_last;
      rc = 0.0;
      rd = row[sk(e)] - Ae * y_last;
      if (t < T - 1) {
        rb -= Ce * As2 * X[t + 1];
        rc = -Ce * Ce12 * X[T + t + 1];
        rd -= Ce * X[2 * T + t + 1];
      }
    }
    // ---------------- cyclic reduction over the T separators
    // Shared-memory PCR (all steps): ping-pong between X[0,4T) and X[4T,8T)
    {
      double* cur = X;
      double* nxt = X + 4 * T;
      __syncthreads();
      cur[t] = ra; cur[T + t] = rb; cur[2 * T + t] = rc; cur[3 * T + t] = rd;
      __syncthreads();
#pragma unroll 1
      for (int st = 1; st < T; st <<= 1) {
        double na = 0.0, nc = 0.0, nb = cur[T + t], nd = cur[3 * T + t];
        const double a_ = cur[t], c_ = cur[2 * T + t];
        const int im = t - st, ip = t + st;
        if (im >= 0) {
          const double al = -a_ / cur[T + im];
          na = al * cur[im];
          nb += al * cur[2 * T + im];
          nd += al * cur[3 * T + im];
        }
        if (ip < T) {
          const double ga = -c_ / cur[T + ip];
          nc = ga * cur[2 * T + ip];
          nb += ga * cur[ip];
          nd += ga * cur[3 * T + ip];
        }
        nxt[t] = na; nxt[T + t] = nb; nxt[2 * T + t] = nc; nxt[3 * T + t] = nd;
        __syncthreads();
        double* tmp = cur; cur = nxt; nxt = tmp;
      }
      const double xe = cur[3 * T + t] / cur[T + t];
      __syncthreads();
      X[t] = xe;
      __syncthreads();
    }
    // ---------------- interior with known neighbours, in place
    {
      const double xe = X[t];
      const double xl = (t > 0) ? X[t - 1] : 0.0;
      double z = row[sk(s)] - As * xl;
      if (M == 2) z -= Ce1 * xe;
      row[sk(s)] = z;
#pragma unroll
      for (int i = 1; i <= M - 2; ++i) {
        const double l = cA(s + i) * rpv[i - 1];
        double di = row[sk(s + i)];
        if (i == M - 2) di -= Ce1 * xe;
        z = di - l * z;
        row[sk(s + i)] = z;
      }
      double x = z * rpv[M - 2];
      row[sk(e - 1)] = x;
#pragma unroll
      for (int i = M - 3; i >= 0; --i) {
        x = (row[sk(s + i)] - cC(s + i) * x) * rpv[i];
        row[sk(s + i)] = x;
      }
      row[sk(e)] = xe;
    }
    __syncthreads();
    // ---------------- store + moments of the new row
    double* dst = a.fout + r * a.ld_out;
    if (a.mom_out) {
      double acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.0;
#pragma unroll 2
      for (int k = 0; k < M; ++k) {
        const int i = t + k * T;
        const double x = row[sk(i)];
        dst[i] = x;
        const double vi = vel(i);
        const double tw = wgt(i) * x;
        acc[0] += tw;
        double p = tw * vi; acc[1] += p;
        p *= vi; acc[2] += p;
        p *= vi; acc[3] += p;
        p *= vi; acc[4] += p;
        p *= vi; acc[5] += p;
        acc[6] += tw * x;
        acc[7] += tw * log(x);
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
#pragma unroll 4
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
  return sizeof(double) * (size_t)(M * T + T + 64 + 8 * T);
}

}  // namespace fpfast
