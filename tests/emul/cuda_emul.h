// tests/emul/cuda_emul.h -- TEST INFRASTRUCTURE: a minimal host emulation of the CUDA execution model, enough to
// run the warp-synchronous kernels of hh-suite_b200/csrc/hhg_mac.cuh on the CPU, bit for bit (IEEE float/double,
// no contraction: compile with -ffp-contract=off).  One OS thread per CUDA thread of a block, blocks run one after
// another; __syncwarp/__syncthreads are barriers, shuffles go through an exchange buffer.  This lets kernel changes
// be checked against the oracle in the authoring container (no GPU) before they are spent GPU minutes on.
#pragma once
#include <algorithm>
#include <barrier>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct emul_dim3 { unsigned x = 1, y = 1, z = 1; emul_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline thread_local emul_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct EmulBlock {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<unsigned long long> xchg;     // one 64-bit slot per thread
};
inline EmulBlock* g_emul_block = nullptr;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

inline void __syncthreads() { g_emul_block->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { g_emul_block->bar->arrive_and_wait(); }   // blocks are single warps here

template <typename T>
inline T emul_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  g_emul_block->xchg[threadIdx.x] = raw;
  g_emul_block->bar->arrive_and_wait();
  const unsigned long long got = g_emul_block->xchg[(threadIdx.x & ~31u) | (unsigned)(src_lane & 31)];
  g_emul_block->bar->arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int lane) { return emul_shfl(v, lane); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask) { return emul_shfl(v, (int)(threadIdx.x & 31) ^ mask); }

inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline long long clock64() { return 0; }
using std::max;
using std::min;

// launch<<<grid, block, smem>>> replacement: run every block of the grid with `threads` OS threads
template <typename Kernel, typename... Args>
void emul_launch(emul_dim3 grid, unsigned threads, Kernel k, Args... args) {
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      EmulBlock blk;
      blk.bar.reset(new std::barrier<>((std::ptrdiff_t)threads));
      blk.xchg.assign(threads, 0);
      g_emul_block = &blk;
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([=]() {
          threadIdx = emul_dim3(t); blockIdx = emul_dim3(bx, by); blockDim = emul_dim3(threads); gridDim = grid;
          k(args...);
        });
      for (auto& th : pool) th.join();
      g_emul_block = nullptr;
    }
}
