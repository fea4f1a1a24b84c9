// rowfft4.cuh -- single-pass e df/dv for nv = 16384 with SIXTEEN warps per SM.
// Same operator and algebra as rowfft.cuh (vlapy/core/vlasov.py:123-138; the real row as one
// complex sequence of M = nv/2 = 8192 points, pairs (k, M-k) un-mixed, phase-multiplied and
// re-mixed in registers by rowfft::pair_op), different shape:
//   rowfft.cuh : 256 threads x 32 points, radix 32 x 16 x 16, 255 registers  ->  8 warps per SM
//   this file  : 512 threads x 16 points, radix 16 x 8 x 8 x 8, 128 registers -> 16 warps per SM
// The row kernels are bound by dependent-issue latency (two warps per scheduler cannot cover the
// fp64 chains of a butterfly), not by HBM or the fp64 pipe: twice the warps is worth one more
// exchange each way.  The M points live in shared memory between stages, in ONE layout
//     slot(k1, u) = 513 k1 + u,      u = 64 j2 + 8 j3 + j4,
// that every stage reads and writes in place (a thread writes the slots it has just read), so a
// stage is one phase and a row costs seven barriers.  Lanes run along j4/j3 in stages 1-3 and along
// k1 in stage 4: the odd pitch of k1 keeps every 128-bit shared-memory transaction conflict free.
// Index bookkeeping (decimation in frequency, natural order in and out of every butterfly):
//   m = 512 m1 + r,  r = 64 m2 + r2,  r2 = 8 m3 + m4        k = k1 + 16 k2 + 128 k3 + 1024 k4
//   stage 1: radix 16 over m1, twiddle W_M^(r k1)           (thread t owns r = t)
//   stage 2: radix 8 over m2,  twiddle W_512^(r2 k2)        (two butterflies per thread)
//   stage 3: radix 8 over m3,  twiddle W_64^(m4 k3)
//   stage 4: radix 8 over m4 for sub-transform s = k1 + 16 k2 + 128 k3; a thread owns s and its partner
//            S - s (S = 1024), i.e. BOTH members of every pair: bin s + S k4 pairs with (S-s) + S (7-k4).
// The next row is prefetched with cp.async into the slots the thread itself reads last (inverse stage
// 1) and first (stage 1), so the copy needs no barrier (as in rowfft.cuh).
#pragma once
#include "rowfft.cuh"

namespace rowfft4 {

using rowfft::Args;
using rowfft::cos32;
using rowfft::sin32;
using fast::fft16;
using fast::fft8;

struct Prog {
  static constexpr int M = 8192, N = 16384, T = 512, S = 1024, V = 16;
  static constexpr int PA = 513;                 // pitch of k1 in the exchange buffer
  static constexpr int NHI = S / 32;
  static constexpr int NPH = 7;
  VPFP_HD static constexpr bool sync_after(int) { return true; }
  static constexpr int X_ELEMS = 16 * PA;
  static constexpr int NTAB = 8 + 32 + NHI;      // G[8], Lo[32], Hi[NHI]
  static constexpr long SMEM_BYTES = (long)sizeof(cplx) * (X_ELEMS + NTAB) + 16;

  struct Regs {
    cplx x[V];
    cplx w1, w4;      // W_M^t and its fourth power (stage-1 twiddles are powers of w1)
    cplx wA, wB;      // W_N^sA, W_N^sB
    cplx v2, v3;      // W_512^(t & 63), W_64^(t & 7): bases of the stage-2 / stage-3 twiddles (both butterflies of a thread share them)
    double phi_pi;    // phase slope of the row: (K[1] dt e[row]) / pi
  };

  Args a;

  VPFP_HD static cplx* xbuf(unsigned char* smem) { return reinterpret_cast<cplx*>(smem); }
  VPFP_HD static cplx* tabs(unsigned char* smem) { return xbuf(smem) + X_ELEMS; }    // G[8], Lo[32], Hi[NHI]
  VPFP_HD static double* cosM(unsigned char* smem) { return reinterpret_cast<double*>(tabs(smem) + NTAB); }
  VPFP_HD static int slot(int k1, int u) { return k1 * PA + u; }

  VPFP_HD void init(int tid, Regs& r, unsigned char* smem) const {
    const cplx w = a.twN[2 * tid];
    const cplx w2 = cmul(w, w);
    r.w1 = w;
    r.w4 = cmul(w2, w2);
    const int sA = (tid == 0) ? 0 : tid, sB = (tid == 0) ? S / 2 : S - tid;
    r.wA = a.twN[sA];
    r.wB = a.twN[sB];
    r.v2 = a.twN[(long)(tid & 63) * (N / 512)];
    r.v3 = a.twN[(long)(tid & 7) * (N / 64)];
    (void)smem;
  }

  // w^1 .. w^7 (six complex multiplies; cheaper than seven table reads per butterfly, which made the
  // kernel shared-memory bound)
  VPFP_HD static void powers7(const cplx w, cplx* p) {
    p[1] = w;
    p[2] = cmul(w, w);
    p[3] = cmul(p[2], w);
    p[4] = cmul(p[2], p[2]);
    p[5] = cmul(p[4], w);
    p[6] = cmul(p[3], p[3]);
    p[7] = cmul(p[6], w);
  }

  VPFP_HD void prefetch_row(long row, int tid, unsigned char* smem) const {
    cplx* X = xbuf(smem);
    const double* src = a.fin + row * a.ld_in;
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) rowfft::cp_async16(X + slot(m1, tid), src + 2L * (m1 * 512 + tid));
    rowfft::cp_async_commit_wait(false);
  }

  // phase tables of the row: G[j] = exp(-i phi S j) (j < 8), Lo[j] = exp(-i phi j), Hi[j] = exp(-i phi 32 j)/(4M),
  // cos(phi M).  Entry w is computed by lane w / 16 of warp w % 16 (every warp pays for a few sincos).
  VPFP_HD void row_tables(int tid, const Regs& r, unsigned char* smem) const {
    constexpr int NWARP = T / 32;
    cplx* G = tabs(smem);
    const int w = (tid & 31) * NWARP + (tid >> 5);
    if ((tid & 31) * NWARP < NTAB + 1 && w < NTAB + 1) {
      double k, sc = 1.0;
      if (w < 8) k = (double)(S * w);
      else if (w < 40) k = (double)(w - 8);
      else if (w < NTAB) { k = (double)(32 * (w - 40)); sc = 0.25 / (double)M; }
      else k = (double)M;
      double sn, cs;
      rowfft::Prog<16, 16>::sincospi_hd(r.phi_pi * k, &sn, &cs);
      if (w < NTAB) G[w] = cmake(cs * sc, -sn * sc);
      else *cosM(smem) = cs;
    }
  }

  VPFP_HD void store_pair(long row, int m, cplx val) const {
    const long n = 2L * m;
    if (a.peer_mode) {
      const int q = (int)(n >> a.lpart);
      const long part = 1L << a.lpart;
      *reinterpret_cast<cplx*>(a.peer[q] + ((long)a.my_rank * a.nrows + row) * part + (n & (part - 1))) = val;
      return;
    }
    *reinterpret_cast<cplx*>(a.fout + row * a.ld_out + n) = val;
  }

  VPFP_HD void phase(int ph, long row, long nextrow, int tid, Regs& r, unsigned char* smem) const {
    cplx* X = xbuf(smem);
    cplx* G = tabs(smem);
    cplx* LO = G + 8;
    cplx* HI = LO + 32;
    cplx* x = r.x;
    switch (ph) {
      case 0: {
        // ---- phase tables; stage 1 (radix 16 over m1) on the row that prefetch_row brought in
        r.phi_pi = mul_rn(mul_rn(a.kvec[1], a.dt), a.cvec[row]) * 0.31830988618379067154;
        row_tables(tid, r, smem);
        rowfft::cp_async_commit_wait(true);
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) x[m1] = X[slot(m1, tid)];
        fft16<-1>(x);
        rowfft::Prog<16, 16>::twiddle1<false>(x, r.w1, r.w4);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) X[slot(k1, tid)] = x[k1];
      } break;
      case 1: {
        // ---- stage 2: radix 8 over m2 for (k1, r2), in place
        cplx tw[8];
        powers7(r.v2, tw);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = tid + T * q, k1 = c >> 6, r2 = c & 63;
#pragma unroll
          for (int m2 = 0; m2 < 8; ++m2) x[q * 8 + m2] = X[slot(k1, m2 * 64 + r2)];
          fft8<-1>(x + q * 8);
#pragma unroll
          for (int k2 = 1; k2 < 8; ++k2) x[q * 8 + k2] = cmul(x[q * 8 + k2], tw[k2]);
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) X[slot(k1, k2 * 64 + r2)] = x[q * 8 + k2];
        }
      } break;
      case 2: {
        // ---- stage 3: radix 8 over m3 for (k1, k2, m4), in place
        cplx tw[8];
        powers7(r.v3, tw);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = tid + T * q, k1 = c >> 6, k2 = (c >> 3) & 7, m4 = c & 7;
#pragma unroll
          for (int m3 = 0; m3 < 8; ++m3) x[q * 8 + m3] = X[slot(k1, k2 * 64 + m3 * 8 + m4)];
          fft8<-1>(x + q * 8);
#pragma unroll
          for (int k3 = 1; k3 < 8; ++k3) x[q * 8 + k3] = cmul(x[q * 8 + k3], tw[k3]);
#pragma unroll
          for (int k3 = 0; k3 < 8; ++k3) X[slot(k1, k2 * 64 + k3 * 8 + m4)] = x[q * 8 + k3];
        }
      } break;
      case 3: {
        // ---- stage 4 for sub-transforms sA, sB; pointwise on pairs; inverse stage 4
        const bool special = (tid == 0);
        const int sA = special ? 0 : tid, sB = special ? S / 2 : S - tid;
        const int uA = ((sA >> 4) & 7) * 64 + (sA >> 7) * 8, uB = ((sB >> 4) & 7) * 64 + (sB >> 7) * 8;
#pragma unroll
        for (int m4 = 0; m4 < 8; ++m4) {
          x[m4] = X[slot(sA & 15, uA + m4)];
          x[8 + m4] = X[slot(sB & 15, uB + m4)];
        }
        fft8<-1>(x);
        fft8<-1>(x + 8);
        const cplx bA = cmul(LO[sA & 31], HI[sA >> 5]);
        const cplx bB = cmul(LO[sB & 31], HI[sB >> 5]);
        // Pair slots j = 0..7 hold (x[j], x[15-j]).  Ordinary threads: bin sA + S j with its partner
        // sB + S (7-j).  Thread 0 owns the two self-paired sub-transforms s = 0 (a[j] = x[j]) and s = S/2
        // (b[j] = x[8+j]); its registers are permuted so that the same slots hold
        //   j = 0: (a[4], a[4])   1..3: (a[j], a[8-j])   4..7: (b[j-4], b[11-j]),
        // and bin 0 (X[0] and X[M], both real) is finished separately.
        const cplx dc = x[0];
        if (special) {
          cplx y[V];
          y[0] = x[4]; y[15] = x[4];
          y[1] = x[1]; y[2] = x[2]; y[3] = x[3];
          y[14] = x[7]; y[13] = x[6]; y[12] = x[5];
          y[4] = x[8]; y[5] = x[9]; y[6] = x[10]; y[7] = x[11];
          y[11] = x[15]; y[10] = x[14]; y[9] = x[13]; y[8] = x[12];
#pragma unroll
          for (int i = 0; i < V; ++i) x[i] = y[i];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int jw = (j < 4) ? j : j - 4;
          // ordinary thread: W_N^(sA + S j), P(sA + S j), P(sB + S (7-j))
          int wj = j, gk = j, gm = 7 - j;
          cplx wbase = r.wA, pbase = bA, qbase = bB;
          if (special) {
            if (j == 0) { wj = 4; gk = 4; gm = 4; qbase = bA; }             // bin M/2: W_N^(M/2) = W_16^4 = -i
            else if (j < 4) { wj = j; gk = j; gm = 8 - j; qbase = bA; }
            else { wj = jw; gk = jw; gm = 7 - jw; wbase = r.wB; pbase = bB; qbase = bB; }
          }
          const cplx w16 = special ? cmake(j == 0 ? 0.0 : cos32(2 * jw), j == 0 ? -1.0 : -sin32(2 * jw))
                                   : cmake(cos32(2 * j), -sin32(2 * j));
          (void)wj;
          const cplx Wk = cmul(wbase, w16);
          const cplx Pk = cmul(pbase, G[gk]);
          const cplx Pm = cmul(qbase, G[gm]);
          rowfft::pair_op(x[j], x[15 - j], Wk, Pk, Pm);
        }
        if (special) {
          cplx y[V];
          {  // bin 0: Y[0] = X[0], Y[M] = Re(P_M) X[M]
            const double sc = 0.5 / (double)M;
            const double y0 = (dc.x + dc.y) * sc, ym = (dc.x - dc.y) * (*cosM(smem)) * sc;
            y[0] = cmake(y0 + ym, y0 - ym);
          }
          y[4] = x[0];
          y[1] = x[1]; y[2] = x[2]; y[3] = x[3];
          y[7] = x[14]; y[6] = x[13]; y[5] = x[12];
          y[8] = x[4]; y[9] = x[5]; y[10] = x[6]; y[11] = x[7];
          y[15] = x[11]; y[14] = x[10]; y[13] = x[9]; y[12] = x[8];
#pragma unroll
          for (int i = 0; i < V; ++i) x[i] = y[i];
        }
        fft8<1>(x);
        fft8<1>(x + 8);
#pragma unroll
        for (int m4 = 0; m4 < 8; ++m4) {
          X[slot(sA & 15, uA + m4)] = x[m4];
          X[slot(sB & 15, uB + m4)] = x[8 + m4];
        }
      } break;
      case 4: {
        // ---- inverse stage 3, in place
        cplx tw[8];
        powers7(r.v3, tw);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = tid + T * q, k1 = c >> 6, k2 = (c >> 3) & 7, m4 = c & 7;
#pragma unroll
          for (int k3 = 0; k3 < 8; ++k3) {
            cplx val = X[slot(k1, k2 * 64 + k3 * 8 + m4)];
            if (k3 > 0) val = cmulc(val, tw[k3]);
            x[q * 8 + k3] = val;
          }
          fft8<1>(x + q * 8);
#pragma unroll
          for (int m3 = 0; m3 < 8; ++m3) X[slot(k1, k2 * 64 + m3 * 8 + m4)] = x[q * 8 + m3];
        }
      } break;
      case 5: {
        // ---- inverse stage 2, in place
        cplx tw[8];
        powers7(r.v2, tw);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = tid + T * q, k1 = c >> 6, r2 = c & 63;
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) {
            cplx val = X[slot(k1, k2 * 64 + r2)];
            if (k2 > 0) val = cmulc(val, tw[k2]);
            x[q * 8 + k2] = val;
          }
          fft8<1>(x + q * 8);
#pragma unroll
          for (int m2 = 0; m2 < 8; ++m2) X[slot(k1, m2 * 64 + r2)] = x[q * 8 + m2];
        }
      } break;
      default: {
        // ---- inverse stage 1, store; the thread's slots take the next row as soon as they are read
        const double* nsrc = a.fin + (nextrow >= 0 ? nextrow : row) * a.ld_in;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
          cplx* sl = X + slot(k1, tid);
          x[k1] = *sl;
          if (nextrow >= 0) rowfft::cp_async16(sl, nsrc + 2L * (k1 * 512 + tid));
        }
        rowfft::cp_async_commit_wait(false);
        rowfft::Prog<16, 16>::twiddle1<true>(x, r.w1, r.w4);
        fft16<1>(x);
        if (!a.peer_mode) {
          cplx* dst = reinterpret_cast<cplx*>(a.fout + row * a.ld_out) + tid;
#pragma unroll
          for (int m1 = 0; m1 < 16; ++m1) dst[m1 * 512] = x[m1];
        } else {
#pragma unroll
          for (int m1 = 0; m1 < 16; ++m1) store_pair(row, m1 * 512 + tid, x[m1]);
        }
      } break;
    }
  }
};

#if defined(__CUDACC__)
__global__ void __launch_bounds__(Prog::T, 1) rowfft4_kernel(const Prog prog) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Prog::Regs r;
  const int tid = (int)threadIdx.x;
  prog.init(tid, r, smem_raw);
  if (blockIdx.x < prog.a.nrows) prog.prefetch_row(blockIdx.x, tid, smem_raw);
  __syncthreads();
  for (long row = blockIdx.x; row < prog.a.nrows; row += gridDim.x) {
    long nxt = row + gridDim.x;
    if (nxt >= prog.a.nrows) nxt = -1;
#pragma unroll
    for (int ph = 0; ph < Prog::NPH; ++ph) {
      prog.phase(ph, row, nxt, tid, r, smem_raw);
      __syncthreads();
    }
  }
}
#endif

}  // namespace rowfft4
