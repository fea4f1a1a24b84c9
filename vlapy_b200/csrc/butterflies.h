// butterflies.h -- radix-4/8/16 register butterflies shared by the fast advection kernels
// (advect_fast.cuh).  Host/device so that tests/emul can check them against numpy.
#pragma once
#include "vpfp_common.h"

namespace fast {

// ------------------------------------------------------------------------------------------
// register butterflies (natural order in and out). DIR = -1 forward (W = e^{-2 pi i/n}), +1 inverse.
// ------------------------------------------------------------------------------------------
template <int DIR>
VPFP_HD cplx rot_i(cplx t) {  // t * (DIR * i)
  return (DIR < 0) ? cmake(t.y, -t.x) : cmake(-t.y, t.x);
}

template <int DIR>
VPFP_HD void fft4(cplx& c0, cplx& c1, cplx& c2, cplx& c3) {
  cplx s0 = cadd(c0, c2), s1 = csub(c0, c2), s2 = cadd(c1, c3), t = rot_i<DIR>(csub(c1, c3));
  c0 = cadd(s0, s2);
  c2 = csub(s0, s2);
  c1 = cadd(s1, t);
  c3 = csub(s1, t);
}

template <int DIR>
VPFP_HD cplx mul_w8_1(cplx d) {  // d * W8^1
  const double h = 0.70710678118654752440;
  return (DIR < 0) ? cmake((d.x + d.y) * h, (d.y - d.x) * h) : cmake((d.x - d.y) * h, (d.x + d.y) * h);
}
template <int DIR>
VPFP_HD cplx mul_w8_3(cplx d) {  // d * W8^3
  const double h = 0.70710678118654752440;
  return (DIR < 0) ? cmake((d.y - d.x) * h, -(d.x + d.y) * h) : cmake(-(d.x + d.y) * h, (d.x - d.y) * h);
}

template <int DIR>
VPFP_HD void fft8(cplx* x) {
  cplx a0 = cadd(x[0], x[4]), a1 = cadd(x[1], x[5]), a2 = cadd(x[2], x[6]), a3 = cadd(x[3], x[7]);
  cplx b0 = csub(x[0], x[4]);
  cplx b1 = mul_w8_1<DIR>(csub(x[1], x[5]));
  cplx b2 = rot_i<DIR>(csub(x[2], x[6]));
  cplx b3 = mul_w8_3<DIR>(csub(x[3], x[7]));
  fft4<DIR>(a0, a1, a2, a3);
  fft4<DIR>(b0, b1, b2, b3);
  x[0] = a0; x[2] = a1; x[4] = a2; x[6] = a3;
  x[1] = b0; x[3] = b1; x[5] = b2; x[7] = b3;
}

template <int DIR>
VPFP_HD cplx mul_w16(cplx d, double c, double s) {  // d * (c - i s) fwd, (c + i s) inv
  return (DIR < 0) ? cmake(d.x * c + d.y * s, d.y * c - d.x * s) : cmake(d.x * c - d.y * s, d.y * c + d.x * s);
}

template <int DIR>
VPFP_HD void fft16(cplx* x) {
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // cos, sin(pi/8)
  const double c3 = 0.38268343236508977173, s3 = 0.92387953251128675613;  // cos, sin(3pi/8)
#pragma unroll
  for (int r = 0; r < 4; ++r) fft4<DIR>(x[r], x[r + 4], x[r + 8], x[r + 12]);
  // y[r][m] sits at x[r + 4m]; twiddle W16^(r m)
  x[1 + 4] = mul_w16<DIR>(x[1 + 4], c1, s1);        // r=1, m=1: W16^1
  x[1 + 8] = mul_w8_1<DIR>(x[1 + 8]);               // r=1, m=2: W16^2
  x[1 + 12] = mul_w16<DIR>(x[1 + 12], c3, s3);      // r=1, m=3: W16^3
  x[2 + 4] = mul_w8_1<DIR>(x[2 + 4]);               // r=2, m=1: W16^2
  x[2 + 8] = rot_i<DIR>(x[2 + 8]);                  // r=2, m=2: W16^4
  x[2 + 12] = mul_w8_3<DIR>(x[2 + 12]);             // r=2, m=3: W16^6
  x[3 + 4] = mul_w16<DIR>(x[3 + 4], c3, s3);        // r=3, m=1: W16^3
  x[3 + 8] = mul_w8_3<DIR>(x[3 + 8]);               // r=3, m=2: W16^6
  x[3 + 12] = mul_w16<DIR>(x[3 + 12], -c1, -s1);    // r=3, m=3: W16^9 = -W16^1
#pragma unroll
  for (int m = 0; m < 4; ++m) fft4<DIR>(x[4 * m], x[4 * m + 1], x[4 * m + 2], x[4 * m + 3]);
  // X[m + 4 k'] sits at x[4m + k'] -> natural order
  cplx t[16];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int k = 0; k < 4; ++k) t[m + 4 * k] = x[4 * m + k];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = t[i];
}

template <int R, int DIR>
VPFP_HD void fftR(cplx* x) {
  if (R == 2) {
    const cplx a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  } else if (R == 4) {
    fft4<DIR>(x[0], x[1], x[2], x[3]);
  } else if (R == 8) {
    fft8<DIR>(x);
  } else {
    fft16<DIR>(x);
  }
}

}  // namespace fast
