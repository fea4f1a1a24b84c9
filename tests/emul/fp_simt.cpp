// fp_simt.cpp -- TEST-ONLY host execution (simt.h: one OS thread per CUDA thread) of the
// Fokker-Planck kernels fp_fast.cuh and fp_reg.cuh.  g++ -O2 -ffp-contract=off -pthread.
#include "simt.h"

#include "../../vlapy_b200/csrc/fp_reg.cuh"

static void fill_logtab(std::vector<double2>& h, int n) {
  h.resize(n);
  for (int i = 0; i < n; ++i) {
    const long double c = 1.0L + ((long double)i + 0.5L) / (long double)n;
    h[i].x = (double)(1.0L / c);
    h[i].y = (double)(-logl((long double)h[i].x));
  }
}

template <int M, int T>
static void run_reg(const fpfast::Args& a, int grid) {
  if (a.op == 0) simt::launch((unsigned)grid, T, [&] { fpreg::fp_reg_kernel<M, T, 0>(a); });
  else simt::launch((unsigned)grid, T, [&] { fpreg::fp_reg_kernel<M, T, 1>(a); });
}
template <int M, int T>
static void run_fast(const fpfast::Args& a, int grid) {
  simt::launch((unsigned)grid, T, [&] { fpfast::fp_kernel<M, T>(a); });
}

// which: 0 = fp_fast.cuh, 1 = fp_reg.cuh.  moments_out: [8][mom_ld] or null.
extern "C" int emul_fp_simt(int which, const double* f_in, long ld_in, double* f_out, long ld_out, double v0,
                            double vstep, double vlast, double nu, double dt, double dv, int op,
                            double* moments_out, long mom_ld, int rows, int nv, int grid) {
  std::vector<double2> lt, lt256;
  fill_logtab(lt, 128);
  fill_logtab(lt256, 64);
  fpfast::Args a;
  a.fin = f_in; a.ld_in = ld_in; a.fout = f_out; a.ld_out = ld_out;
  a.v0 = v0; a.vstep = vstep; a.vlast = vlast; a.nu = nu; a.dt = dt; a.dv = dv; a.op = op;
  a.mom_out = moments_out; a.mom_ld = mom_ld; a.rows = rows; a.nv = nv; a.logtab = lt.data(); a.logtab64 = lt256.data();
  if (which == 1) {
    if (nv == 16384) run_reg<32, 512>(a, grid);
    else if (nv == 8192) run_reg<32, 256>(a, grid);
    else if (nv == 4096) run_reg<32, 128>(a, grid);
    else return 1;
    return 0;
  }
  if (nv == 16384) run_fast<32, 512>(a, grid);
  else if (nv == 4096) run_fast<16, 256>(a, grid);
  else if (nv == 1024) run_fast<8, 128>(a, grid);
  else return 1;
  return 0;
}
