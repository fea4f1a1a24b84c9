// simt.h -- TEST-ONLY host execution of CUDA kernels that are not written as phase programs
// (fp_fast.cuh, fp_reg.cuh): every CUDA thread of a CTA is an OS thread, __syncthreads() and warp
// shuffles are real barriers, CTAs run one after another.  Slow (use a handful of rows), but it
// runs the kernel SOURCE itself, so index arithmetic and numerics can be checked against the
// oracle on a machine without a GPU.  Never part of the product library.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace simt {

struct Dim { unsigned x = 0, y = 0, z = 0; };
inline thread_local Dim threadIdx_, blockIdx_;
inline Dim blockDim_, gridDim_;

class Barrier {
 public:
  void reset(int n) { n_ = n; count_ = 0; gen_ = 0; acc_ = 0; }
  int wait(int pred = 0) {
    std::unique_lock<std::mutex> lk(m_);
    acc_ |= (pred != 0);
    const int g = gen_;
    if (++count_ == n_) {
      res_[g & 1] = acc_;
      acc_ = 0; count_ = 0; ++gen_;
      cv_.notify_all();
      return res_[g & 1];
    }
    cv_.wait(lk, [&] { return gen_ != g; });
    return res_[g & 1];
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_ = 0, count_ = 0, gen_ = 0, acc_ = 0, res_[2] = {0, 0};
};

struct Warp {
  Barrier bar;
  double slot[32];
};

inline Barrier cta_bar;
inline std::vector<Warp>* warps = nullptr;
alignas(16) inline unsigned char smem[232448];

inline void syncthreads() { cta_bar.wait(0); }
inline int syncthreads_or(int p) { return cta_bar.wait(p); }
inline void syncwarp() { (*warps)[threadIdx_.x >> 5].bar.wait(); }
inline double shfl_xor(double x, int o) {
  Warp& w = (*warps)[threadIdx_.x >> 5];
  const int lane = threadIdx_.x & 31;
  w.slot[lane] = x;
  w.bar.wait();
  const double r = w.slot[lane ^ o];
  w.bar.wait();
  return r;
}

// run `body` as a grid of `grid` CTAs of `block` threads (block a multiple of 32)
inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
  gridDim_.x = grid; blockDim_.x = block;
  std::vector<Warp> w(block / 32);
  for (auto& x : w) x.bar.reset(32);
  warps = &w;
  for (unsigned b = 0; b < grid; ++b) {
    cta_bar.reset((int)block);
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([&, t, b] {
        threadIdx_.x = t; blockIdx_.x = b;
        body();
      });
    for (auto& x : th) x.join();
  }
  warps = nullptr;
}

}  // namespace simt

inline unsigned char* simt_dyn_smem() { return simt::smem; }

#if !defined(__CUDACC__)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__
#define __align__(n) alignas(n)
#define threadIdx simt::threadIdx_
#define blockIdx simt::blockIdx_
#define blockDim simt::blockDim_
#define gridDim simt::gridDim_
#define __syncthreads() simt::syncthreads()
#define __syncthreads_or(p) simt::syncthreads_or(p)
#define __syncwarp() simt::syncwarp()
#define __shfl_xor_sync(mask, x, o) simt::shfl_xor((x), (o))
struct double2 { double x, y; };
inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
#endif
