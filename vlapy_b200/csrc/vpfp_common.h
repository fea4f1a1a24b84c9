// vpfp_common.h -- shared host/device helpers for the VPFP kernels.
//
// Every kernel body in this library is written as a "program": a struct with
//   int  nphases() const;
//   void phase(int ph, int blk, int tid, int nthr, unsigned char* smem) const;
// where consecutive phases are separated by a CTA-wide barrier.  On the GPU a generic
// __global__ driver runs the phases with __syncthreads() between them (vpfp_cuda.cu); the very
// same source compiles with g++ for tests/emul/, where a host driver runs phase(ph) for every
// tid sequentially.  That lets the index arithmetic and numerics be checked against the oracle
// on a machine without a GPU.  The emulation is test infrastructure only; the product library
// contains no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VPFP_HD __host__ __device__ __forceinline__
#define VPFP_ALIGN16 __align__(16)
#else
#define VPFP_HD inline
#define VPFP_ALIGN16 __attribute__((aligned(16)))
#endif

// dynamic shared memory of a __global__ kernel (tests/emul/simt.h provides the host stand-in)
#if defined(__CUDACC__)
#define VPFP_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#define VPFP_DYN_SMEM(name) unsigned char* name = simt_dyn_smem()
#endif

struct VPFP_ALIGN16 cplx {
  double x, y;
};

// Rounded-once multiply/add so that phase arguments follow the reference's two roundings
// ((k*dt)*c) no matter how the compiler would like to contract them.
VPFP_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}

VPFP_HD cplx cmake(double x, double y) {
  cplx r;
  r.x = x;
  r.y = y;
  return r;
}
VPFP_HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
VPFP_HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
VPFP_HD cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
VPFP_HD cplx cmulc(cplx a, cplx b) {  // a * conj(b)
  return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
VPFP_HD cplx cconj(cplx a) { return cmake(a.x, -a.y); }
VPFP_HD cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
VPFP_HD cplx cmul_i(cplx a) { return cmake(-a.y, a.x); }    // a * (+i)
VPFP_HD cplx cmul_mi(cplx a) { return cmake(a.y, -a.x); }   // a * (-i)

VPFP_HD unsigned brev_bits(unsigned p, int bits) {
  if (bits == 0) return 0u;
#if defined(__CUDA_ARCH__)
  return __brev(p) >> (32 - bits);
#else
  unsigned r = 0;
  for (int i = 0; i < bits; ++i) r |= ((p >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}

VPFP_HD int ilog2(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}

VPFP_HD void sincos_hd(double t, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(t, s, c);
#else
  *s = sin(t);
  *c = cos(t);
#endif
}

VPFP_HD double ldg(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
