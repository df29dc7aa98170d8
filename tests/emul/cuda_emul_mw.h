// tests/emul/cuda_emul_mw.h -- TEST INFRASTRUCTURE: host emulation of the CUDA execution model for MULTI-WARP blocks
// (the alignment -> HMM kernels of hh-suite_b200/csrc/hhg_msa.cuh).  One OS thread per CUDA thread, blocks run one
// after another; __syncthreads is a block barrier, shuffles use a per-warp barrier and exchange buffer, __shared__
// variables are statics (one block is alive at a time), atomics are the compiler's.  IEEE arithmetic, no contraction
// (-ffp-contract=off), so results can be compared with the reference bit for bit in a container without a GPU.
#pragma once
#include <algorithm>
#include <barrier>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct emul_dim3 { unsigned x = 1, y = 1, z = 1; emul_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline thread_local emul_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct EmulBlock {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<std::barrier<>>> wbar;
  std::vector<unsigned long long> xchg;
};
inline EmulBlock* g_emul_block = nullptr;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

inline void __syncthreads() { g_emul_block->bar->arrive_and_wait(); }
inline void emul_syncwarp() { g_emul_block->wbar[threadIdx.x >> 5]->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emul_syncwarp(); }

template <typename T>
inline T emul_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  g_emul_block->xchg[threadIdx.x] = raw;
  emul_syncwarp();
  const unsigned long long got = g_emul_block->xchg[(threadIdx.x & ~31u) | (unsigned)(src_lane & 31)];
  emul_syncwarp();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int lane) { return emul_shfl(v, lane); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask) { return emul_shfl(v, (int)(threadIdx.x & 31) ^ mask); }

inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __double2float_rn(double a) { return (float)a; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline uint32_t __vcmpltu4(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) if (((a >> (8 * k)) & 255u) < ((b >> (8 * k)) & 255u)) r |= 255u << (8 * k);
  return r;
}
inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) if (((a >> (8 * k)) & 255u) == ((b >> (8 * k)) & 255u)) r |= 255u << (8 * k);
  return r;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMax(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
inline int atomicMin(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
struct alignas(16) float4 { float x, y, z, w; };
using std::max;
using std::min;

template <typename Kernel, typename... Args>
void emul_launch2(unsigned grid, unsigned grid_y, unsigned threads, Kernel k, Args... args);

template <typename Kernel, typename... Args>
void emul_launch(unsigned grid, unsigned threads, Kernel k, Args... args) { emul_launch2(grid, 1u, threads, k, args...); }

template <typename Kernel, typename... Args>
void emul_launch2(unsigned grid, unsigned grid_y, unsigned threads, Kernel k, Args... args) {
  for (unsigned by = 0; by < grid_y; ++by)
  for (unsigned bx = 0; bx < grid; ++bx) {
    EmulBlock blk;
    blk.bar.reset(new std::barrier<>((std::ptrdiff_t)threads));
    for (unsigned w = 0; w * 32 < threads; ++w)
      blk.wbar.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(32u, threads - w * 32)));
    blk.xchg.assign(threads, 0);
    g_emul_block = &blk;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t)
      pool.emplace_back([=]() {
        threadIdx = emul_dim3(t); blockIdx = emul_dim3(bx, by); blockDim = emul_dim3(threads); gridDim = emul_dim3(grid, grid_y);
        k(args...);
      });
    for (auto& th : pool) th.join();
    g_emul_block = nullptr;
  }
}
