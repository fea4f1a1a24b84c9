// rowfft2.cuh -- e df/dv row kernel for nv = 16384 with TWO independent CTAs per SM.
// Same operator and same arithmetic decomposition as rowfft.cuh (vlapy/core/vlasov.py:123-138,
//     f_new[x, :] = Re ifft_v( exp(-i kv dt e[x]) fft_v f[x, :] ),
// real row as one complex sequence of M = 8192 points, M = 32 x 16 x 16, pairs (k, M-k) un-mixed, phase
// multiplied and re-mixed in registers), but organised so that a CTA needs only HALF the registers and
// HALF the shared memory of rowfft.cuh:
//
// rowfft.cuh keeps a whole row in the registers of one 256-thread CTA, which fills the register file: one
// CTA per SM, every warp of the SM in the same barrier-separated phase, so its shared-memory exchanges
// (0.63 ms), its fp64 butterflies (0.78 ms) and its HBM traffic (0.65 ms) are paid one after the other
// (1.73 ms measured).  Here a CTA of 128 threads works on one PARITY of the spectrum at a time: the first
// radix-32 stage is split into its radix-2 step and two radix-16 transforms, the even outputs k1 = 2 kappa
// (parity h = 0) and the odd outputs k1 = 2 kappa + 1 (h = 1) are two independent 4096-point problems
//     y_h[j] = (z[j] + (-1)^h z[j + M/2]) W_32^(j h)   ->   16 x 16 x 16 transform,
// and the bins k and M - k have the same parity, so the whole pointwise step stays inside a parity.  The
// CTA runs parity 0 (row read from HBM), parks the 4096 inverse-transformed values E in shared memory
// (a thread parks and later reads only its own slots), runs parity 1 (row re-read: an L2 hit, the row was
// loaded a few microseconds earlier) and finishes with the last radix-2 step
//     out[j] = E[j] + conj(W_32^j) O[j],   out[j + 16 L2] = E[j] - conj(W_32^j) O[j].
// 32 complex values per thread (half of them in flight through the exchange buffer, which therefore holds
// half a parity: 34 KB), 64 KB of parking space: 103 KB of shared memory and 128 x 255 registers per CTA,
// so TWO CTAs are resident per SM and run their phases independently -- one CTA's shared-memory and HBM
// phases overlap the other's fp64 phases.  DRAM traffic is unchanged (16 B/cell), L2 -> SM reads double.
//
// Index sets (thread t = 0..127, parity h):
//   stage 1: columns rr = t, t + 128 (z[m1 L2 + rr], m1 = 0..31; L2 = 256), outputs kappa = 0..15, twiddle
//            W_M^(rr (2 kappa + h));
//   stage 2: items (kappa, m3) = (t >> 4 [+ 8], t & 15), radix 16 over m2 (rr = 16 m2 + m3), twiddle
//            W_L2^(m3 k2);
//   stage 3: sub-transforms sigma = kappa + 16 k2 (s = h + 2 sigma in rowfft.cuh's numbering, bins
//            k = s + 512 k3); the thread owns sigma_A = t and its partner sigma_B = 256 - h - t (thread 0 of
//            parity 0: the self-paired 0 and 128), i.e. both members of every pair (k, M - k).
//   Exchanges move half of the values at a time (the exchange buffer holds 16 of a thread's 32 values):
//   exchange 1 by kappa < 8 / >= 8 (= the consumer's two items), exchange 2 by k2 < 8 / >= 8 (sigma_A has
//   k2 < 8, sigma_B has k2 >= 8); a half is written from, and read back into, the same 16 registers.
//
// The kernel body is a phase program (vpfp_common.h): 16 barrier-separated phases per parity; tests/emul
// runs the same source thread by thread on the host.
#pragma once
#include "rowfft.cuh"

namespace rowfft2 {

using fast::fft16;
using rowfft::Args;
using rowfft::cos32;
using rowfft::pair_op;
using rowfft::sin32;

// HINTS (cache hints, A/B): bit 0: the finished row is stored with evict-first (st.global.cs), so that
// it does not push the rows that wait for their parity-1 re-read out of L2; bit 1: the parity-1 re-read is
// a last-use load (ld.global.lu).
template <int HINTS>
struct ProgT {
  static constexpr int T = 128, V = 32;
  static constexpr int M = 8192, N = 16384, L2 = 256, S = 512;
  static constexpr int NPH = 32;                     // 16 phases per parity
  VPFP_HD static constexpr bool sync_after(int) { return true; }
  static constexpr int X_ELEMS = 128 * 17;           // half a parity, rows of 16 padded to 17 (>= 8 * L2)
  static constexpr int NHI = 16, NTAB = 16 + 32 + NHI;
  static constexpr int PARK = 4096;
  static constexpr long SMEM_BYTES = (long)sizeof(cplx) * (X_ELEMS + L2 + NTAB + 1 + PARK);

  struct Regs {
    cplx x[V];
  };

  Args a;

  VPFP_HD static cplx* xbuf(unsigned char* smem) { return reinterpret_cast<cplx*>(smem); }
  VPFP_HD static cplx* tw2(unsigned char* smem) { return xbuf(smem) + X_ELEMS; }
  VPFP_HD static cplx* tabs(unsigned char* smem) { return tw2(smem) + L2; }   // G[16], Lo[32], Hi[16]
  VPFP_HD static double* cosM(unsigned char* smem) { return reinterpret_cast<double*>(tabs(smem) + NTAB); }
  VPFP_HD static cplx* park(unsigned char* smem) { return tabs(smem) + NTAB + 1; }

  VPFP_HD static cplx ldc(const cplx* p) {
#if defined(__CUDA_ARCH__)
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return cmake(v.x, v.y);
#else
    return *p;
#endif
  }
  // a point of the row: parity 0 reads it with the default policy (it is needed again), parity 1 for the last time
  template <int H>
  VPFP_HD static cplx ld_row(const cplx* p) {
#if defined(__CUDA_ARCH__)
    if (H == 1 && (HINTS & 2)) {
      const double2 v = __ldlu(reinterpret_cast<const double2*>(p));
      return cmake(v.x, v.y);
    }
#endif
    return *p;
  }
  VPFP_HD static void st_row(cplx* p, const cplx v) {
#if defined(__CUDA_ARCH__)
    if (HINTS & 1) {
      __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
      return;
    }
#endif
    *p = v;
  }

  // once per CTA: the stage-2 twiddle table W_L2^j
  VPFP_HD void init(int tid, Regs&, unsigned char* smem) const {
    cplx* TW2 = tw2(smem);
    for (int j = tid; j < L2; j += T) TW2[j] = a.twN[(long)j * (N / L2)];
  }
  VPFP_HD void prefetch_row(long, int, unsigned char*) const {}   // interface of the emulation driver

  // phase tables of the row: G[j] = exp(-i phi S j), Lo[j] = exp(-i phi j), Hi[j] = exp(-i phi 32 j)/(4M),
  // cos(phi M); phi = (K[1] dt) e[row].  One sincos per thread for the first 65 (lane-major) threads.
  VPFP_HD void row_tables(long row, int tid, unsigned char* smem) const {
    constexpr int NWARP = T / 32;
    const int w = (tid & 31) * NWARP + (tid >> 5);
    if (w < NTAB + 1) {
      const double phi_pi = mul_rn(mul_rn(a.kvec[1], a.dt), a.cvec[row]) * 0.31830988618379067154;
      double k, sc = 1.0;
      if (w < 16) k = (double)(S * w);
      else if (w < 48) k = (double)(w - 16);
      else if (w < NTAB) { k = (double)(32 * (w - 48)); sc = 0.25 / (double)M; }
      else k = (double)M;
      double sn, cs;
      rowfft::Prog<32, 16>::sincospi_hd(phi_pi * k, &sn, &cs);
      if (w < NTAB) tabs(smem)[w] = cmake(cs * sc, -sn * sc);
      else *cosM(smem) = cs;
    }
  }

  // y[kappa] *= w^(2 kappa + H)  (CONJ: conj(w)^(2 kappa + H)), kappa = 0..15; powers as u^(4 g) (w^H u^i), u = w^2
  template <int H, bool CONJ>
  VPFP_HD static void twiddle1(cplx* y, const cplx w) {
    const cplx u = cmul(w, w), u2 = cmul(u, u), u4 = cmul(u2, u2);
    const cplx b0 = H ? w : cmake(1.0, 0.0);
    const cplx b1 = H ? cmul(w, u) : u;
    const cplx b2 = H ? cmul(b1, u) : u2;
    const cplx b3 = cmul(b2, u);
    cplx A = u4;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (g == 0 && i == 0 && !H) continue;                 // w^0
        const cplx b = (i == 0) ? b0 : (i == 1) ? b1 : (i == 2) ? b2 : b3;
        const cplx tw = (g == 0) ? b : ((i == 0 && !H) ? A : cmul(A, b));
        y[4 * g + i] = CONJ ? cmulc(y[4 * g + i], tw) : cmul(y[4 * g + i], tw);
      }
      if (g > 0 && g < 3) A = cmul(A, u4);
    }
  }

  // registers of exchange-2 half g that hold the point m3 of the thread's sub-transform
  VPFP_HD static constexpr int slot2(int g, int m3) { return 16 * (m3 >> 3) + 8 * g + (m3 & 7); }

  template <int H>
  VPFP_HD void phase_h(int p, long row, long nextrow, int tid, Regs& r, unsigned char* smem) const {
    cplx* X = xbuf(smem);
    cplx* TW2 = tw2(smem);
    cplx* G = tabs(smem);
    cplx* LO = G + 16;
    cplx* HI = LO + 32;
    cplx* PK = park(smem);
    cplx* x = r.x;
    const int kq = tid >> 4, m3 = tid & 15;      // stage-2 item (kappa = kq + 8 q', m3)
    switch (p) {
      case 0: {
        // ---- row tables (once per row); stage 1 of this parity straight from global memory
        if (H == 0) row_tables(row, tid, smem);
        if (H == 1 && tid == 0 && a.l2_prefetch > 0 && nextrow >= 0) {
#if defined(__CUDA_ARCH__)
          const double* pn = a.fin + nextrow * a.ld_in;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pn), "r"((unsigned)(N * 8)) : "memory");
#endif
        }
        const cplx* src = reinterpret_cast<const cplx*>(a.fin + row * a.ld_in);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int rr = tid + T * q;
          cplx y[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const cplx za = ld_row<H>(src + j * L2 + rr), zb = ld_row<H>(src + (j + 16) * L2 + rr);
            if (H == 0) y[j] = cadd(za, zb);
            else {
              const cplx d = csub(za, zb);
              y[j] = (j == 0) ? d : (j == 8) ? fast::rot_i<-1>(d) : fast::mul_w16<-1>(d, cos32(j), sin32(j));
            }
          }
          fft16<-1>(y);
          twiddle1<H, false>(y, ldc(a.twN + 2 * rr));
#pragma unroll
          for (int k = 0; k < 16; ++k) x[16 * (k >> 3) + 8 * q + (k & 7)] = y[k];
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int kl = 0; kl < 8; ++kl) X[kl * L2 + tid + T * q] = x[8 * q + kl];
      } break;
      case 1:
      case 3: {
        // ---- stage 2 of item q' = (p - 1) / 2: radix 16 over m2, twiddle W_L2^(m3 k2)
        const int o = (p == 1) ? 0 : 16;
        cplx y[16];
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) y[m2] = X[kq * L2 + m2 * 16 + m3];
        fft16<-1>(y);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) x[o + k2] = (k2 == 0) ? y[0] : cmul(y[k2], TW2[m3 * k2]);
      } break;
      case 2: {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int kl = 0; kl < 8; ++kl) X[kl * L2 + tid + T * q] = x[16 + 8 * q + kl];
      } break;
      case 4:
      case 6: {
        // ---- exchange 2, half g: the values with k2 = 8 g + j of both items
        const int g = (p == 4) ? 0 : 1;
#pragma unroll
        for (int qq = 0; qq < 2; ++qq)
#pragma unroll
          for (int j = 0; j < 8; ++j) X[(kq + 8 * qq + 16 * j) * 17 + m3] = x[16 * qq + 8 * g + j];
      } break;
      case 5: {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[slot2(0, i)] = X[tid * 17 + i];                // sigma_A = tid
      } break;
      case 7: {
        // ---- sigma_B; stage 3 of both sub-transforms; pointwise on pairs; inverse stage 3
        const bool special = (H == 0) && (tid == 0);
        {
          const int lamB = (128 - H - tid) & 127;
#pragma unroll
          for (int i = 0; i < 16; ++i) x[slot2(1, i)] = X[lamB * 17 + i];
        }
        cplx z[V];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          z[i] = x[slot2(0, i)];
          z[16 + i] = x[slot2(1, i)];
        }
        fft16<-1>(z);
        fft16<-1>(z + 16);
        const int sA = special ? 0 : H + 2 * tid, sB = special ? S / 2 : S - H - 2 * tid;   // s = h + 2 sigma
        const cplx wA = ldc(a.twN + sA), wB = ldc(a.twN + sB);
        const cplx bA = cmul(LO[sA & 31], HI[sA >> 5]);
        const cplx bB = cmul(LO[sB & 31], HI[sB >> 5]);
        // Pair slots j = 0..15 hold (z[j], z[31-j]): bin sA + S j with its partner sB + S (15 - j).
        // Thread 0 of parity 0 owns the self-paired sub-transforms s = 0 and s = S/2; its registers are
        // permuted so that the same slots hold  j = 0: (A[8], A[8])  1..7: (A[j], A[16-j])
        // 8..15: (B[j-8], B[23-j]),  and bin 0 (X[0] and X[M], both real) is finished separately
        // (rowfft.cuh, phase 3).
        cplx dc = z[0];
        if (H == 0 && special) {
          cplx y[V];
          y[0] = z[8]; y[31] = z[8];
#pragma unroll
          for (int i = 1; i < 8; ++i) y[i] = z[i];
#pragma unroll
          for (int i = 8; i < 16; ++i) y[i] = z[i + 8];
#pragma unroll
          for (int i = 16; i < 24; ++i) y[i] = z[i + 8];
#pragma unroll
          for (int i = 24; i < 31; ++i) y[i] = z[i - 15];
#pragma unroll
          for (int i = 0; i < V; ++i) z[i] = y[i];
        }
        const cplx wHi = special ? wB : wA;
        const cplx pHi = special ? bB : bA;
        const cplx qLo = special ? bA : bB;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int jw = (j < 8) ? j : j - 8;
          const int gk = special ? (j == 0 ? 8 : jw) : j;
          const int gm = special ? (j == 0 ? 8 : (j < 8 ? 16 - j : 23 - j)) : 15 - j;
          cplx w32 = (j < 8) ? cmake(cos32(j), -sin32(j))
                             : cmake(special ? cos32(jw) : cos32(j), special ? -sin32(jw) : -sin32(j));
          if (j == 0 && special) w32 = cmake(0.0, -1.0);
          const cplx Wk = cmul((j < 8) ? wA : wHi, w32);
          const cplx Pk = cmul((j < 8) ? bA : pHi, G[gk]);
          const cplx Pm = cmul((j < 8) ? qLo : bB, G[gm]);
          pair_op(z[j], z[31 - j], Wk, Pk, Pm);
        }
        if (H == 0 && special) {
          cplx y[V];
          {
            const double sc = 0.5 / (double)M;
            const double y0 = (dc.x + dc.y) * sc, ym = (dc.x - dc.y) * (*cosM(smem)) * sc;
            y[0] = cmake(y0 + ym, y0 - ym);
          }
          y[8] = z[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) y[i] = z[i];
#pragma unroll
          for (int i = 8; i < 16; ++i) y[i + 8] = z[i];
#pragma unroll
          for (int i = 16; i < 24; ++i) y[i + 8] = z[i];
#pragma unroll
          for (int i = 24; i < 31; ++i) y[i - 15] = z[i];
#pragma unroll
          for (int i = 0; i < V; ++i) z[i] = y[i];
        }
        fft16<1>(z);
        fft16<1>(z + 16);
#pragma unroll
        for (int i = 0; i < V; ++i) x[i] = z[i];          // x[0..15]: sigma_A by m3, x[16..31]: sigma_B
      } break;
      case 8: {
#pragma unroll
        for (int i = 0; i < 16; ++i) X[tid * 17 + i] = x[i];
      } break;
      case 9: {
        // ---- inverse exchange 2, half 0: k2 = j of both items into the registers sigma_A left
#pragma unroll
        for (int qq = 0; qq < 2; ++qq)
#pragma unroll
          for (int j = 0; j < 8; ++j) x[8 * qq + j] = X[(kq + 8 * qq + 16 * j) * 17 + m3];
      } break;
      case 10: {
        const int lamB = (128 - H - tid) & 127;
#pragma unroll
        for (int i = 0; i < 16; ++i) X[lamB * 17 + i] = x[16 + i];
      } break;
      case 11: {
        // ---- half 1 (k2 = 8 + j); inverse stage 2 of both items
#pragma unroll
        for (int qq = 0; qq < 2; ++qq)
#pragma unroll
          for (int j = 0; j < 8; ++j) x[16 + 8 * qq + j] = X[(kq + 8 * qq + 16 * j) * 17 + m3];
        cplx y[V];
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const cplx val = x[16 * (k2 >> 3) + 8 * qq + (k2 & 7)];
            y[16 * qq + k2] = (k2 == 0) ? val : cmulc(val, TW2[m3 * k2]);
          }
          fft16<1>(y + 16 * qq);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) x[i] = y[i];          // x[16 q' + m2]
      } break;
      case 12:
      case 14: {
        const int o = (p == 12) ? 0 : 16;
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) X[kq * L2 + m2 * 16 + m3] = x[o + m2];
      } break;
      case 13: {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int kl = 0; kl < 8; ++kl) x[8 * q + kl] = X[kl * L2 + tid + T * q];
      } break;
      default: {
        // ---- kappa >= 8; inverse stage 1; parity 0 parks its result, parity 1 finishes the row
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int kl = 0; kl < 8; ++kl) x[16 + 8 * q + kl] = X[kl * L2 + tid + T * q];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int rr = tid + T * q;
          cplx y[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) y[k] = x[16 * (k >> 3) + 8 * q + (k & 7)];
          twiddle1<H, true>(y, ldc(a.twN + 2 * rr));
          fft16<1>(y);
          if (H == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) PK[(16 * q + j) * T + tid] = y[j];
          } else {
            cplx* dst = reinterpret_cast<cplx*>(a.fout + row * a.ld_out) + rr;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const cplx o = (j == 0) ? y[0] : (j == 8) ? fast::rot_i<1>(y[8]) : fast::mul_w16<1>(y[j], cos32(j), sin32(j));
              const cplx ev = PK[(16 * q + j) * T + tid];
              st_row(dst + j * L2, cadd(ev, o));
              st_row(dst + (j + 16) * L2, csub(ev, o));
            }
          }
        }
      } break;
    }
  }

  VPFP_HD void phase(int ph, long row, long nextrow, int tid, Regs& r, unsigned char* smem) const {
    if (ph < 16) phase_h<0>(ph, row, nextrow, tid, r, smem);
    else phase_h<1>(ph - 16, row, nextrow, tid, r, smem);
  }
};

using Prog = ProgT<0>;

#if defined(__CUDACC__)
// the 32 phases as straight-line code (compile-time phase index: the registers of Regs never become an array in
// local memory)
template <int PH, class P>
__device__ __forceinline__ void run_phases(const P& prog, long row, long nxt, int tid, typename P::Regs& r,
                                           unsigned char* smem) {
  if constexpr (PH < P::NPH) {
    if constexpr (PH < 16) prog.template phase_h<0>(PH, row, nxt, tid, r, smem);
    else prog.template phase_h<1>(PH - 16, row, nxt, tid, r, smem);
    __syncthreads();
    run_phases<PH + 1, P>(prog, row, nxt, tid, r, smem);
  }
}

template <class P>
__global__ void __launch_bounds__(P::T, 2) rowfft2_kernel(const P prog) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typename P::Regs r;
  const int tid = (int)threadIdx.x;
  prog.init(tid, r, smem_raw);
  __syncthreads();
  for (long row = blockIdx.x; row < prog.a.nrows; row += gridDim.x) {
    long nxt = row + gridDim.x;
    if (nxt >= prog.a.nrows) nxt = -1;
    run_phases<0, P>(prog, row, nxt, tid, r, smem_raw);
  }
}
#endif

}  // namespace rowfft2
