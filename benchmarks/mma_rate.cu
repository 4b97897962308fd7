// Rate probe for tcgen05.mma on sm_100a: one CTA per SM, one elected lane of a converged warp issues `iters`
// back-to-back MMAs of one shape (operands are whatever the memory holds), clock64 around issue + commit + wait.
// (Issued from inside `if (threadIdx.x == 32)` instead, every shape costs ~120 cycles: the compiler wraps each
// instruction in an ELECT / R2UR / branch waterfall because the operands are not provably warp-uniform.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../torch-geometric-pool_b200/csrc -I../include \
//        mma_rate.cu -o mma_rate.bin && ./mma_rate.bin
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace tgp::tc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, bool tf32) {
  if (tf32)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                 "r"(a), "l"(db), "r"(idesc)
                 : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                 "r"(a), "l"(db), "r"(idesc)
                 : "memory");
}

// mode: 0 = SS (A, B from smem), 1 = TS (A from TMEM)
__global__ void __launch_bounds__(128, 1) k_rate(int iters, int N, int tf32, int mode, int b_mn, int n_acc, int M, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {  // whole warp runs the loop; one elected lane issues (uniform operands, no waterfall)
    const uint32_t fmt = tf32 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(smem), 16, 1024, 2);
    const uint64_t db = b_mn ? make_desc(smem_u32(smem) + 16384, 2048, tf32 ? 512 : 1024, tf32 ? 1 : 2)
                             : make_desc(smem_u32(smem) + 16384, 16, 1024, 2);
    const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
    const uint32_t amask = (uint32_t)(n_acc - 1), nn = (uint32_t)N, ta = tmu + 448;
    long long t0 = clock64();
    if (elect_one()) {
#pragma unroll 16
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmu + ((uint32_t)i & amask) * nn;  // n_acc is a power of two
        if (mode == 1) mma_ts(d, ta, db, idesc, tf32 != 0);
        else if (tf32) umma<true>(d, da, db, idesc, 1u);
        else umma<false>(d, da, db, idesc, 1u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// The MMA warp's loop of the engine kernels in isolation: per iteration an mbarrier wait on an already complete phase,
// tcgen05.fence, n_mma MMAs (TS, tf32, N = 64 / 128 alternating when mix), n_commit commits, __syncwarp.
__global__ void __launch_bounds__(128, 1) k_loop(int iters, int n_mma, int do_wait, int n_commit, int mix, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar_ready, bar_c[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar_ready), 1);
    mbar_init(smem_u32(&bar_c[0]), 1);
    mbar_init(smem_u32(&bar_c[1]), 1);
    fence_barrier_init();
    mbar_arrive(smem_u32(&bar_ready));  // phase 0 complete for good
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (idesc & ~(0x3fu << 17)) | ((uint32_t)(128 >> 3) << 17);
    const uint64_t db = make_desc(smem_u32(smem) + 16384, 2048, 512, 1);
    const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
    const uint32_t ta = tmu + 448;
    long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < iters; ++i) {
        if (do_wait) mbar_wait(smem_u32(&bar_ready), 0);
        tc_fence_after();
        if (elect_one()) {
          for (int j = 0; j < n_mma; ++j) mma_ts(tmu, ta, db, (mix && !(j & 1)) ? idesc2 : idesc, true);
          if (n_commit > 0) umma_commit(smem_u32(&bar_c[0]));
          if (n_commit > 1) umma_commit(smem_u32(&bar_c[1]));
        }
        __syncwarp();
      }
    } else if (mode == 1) {  // one elected thread runs the whole loop
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          if (do_wait) mbar_wait(smem_u32(&bar_ready), 0);
          tc_fence_after();
          for (int j = 0; j < n_mma; ++j) mma_ts(tmu, ta, db, (mix && !(j & 1)) ? idesc2 : idesc, true);
          if (n_commit > 0) umma_commit(smem_u32(&bar_c[0]));
          if (n_commit > 1) umma_commit(smem_u32(&bar_c[1]));
        }
      }
      __syncwarp();
    } else if (mode == 2) {  // no tcgen05.fence
      for (int i = 0; i < iters; ++i) {
        if (do_wait) mbar_wait(smem_u32(&bar_ready), 0);
        if (elect_one()) {
          for (int j = 0; j < n_mma; ++j) mma_ts(tmu, ta, db, (mix && !(j & 1)) ? idesc2 : idesc, true);
          if (n_commit > 0) umma_commit(smem_u32(&bar_c[0]));
          if (n_commit > 1) umma_commit(smem_u32(&bar_c[1]));
        }
        __syncwarp();
      }
    } else if (mode == 3) {  // no __syncwarp
      for (int i = 0; i < iters; ++i) {
        if (do_wait) mbar_wait(smem_u32(&bar_ready), 0);
        tc_fence_after();
        if (elect_one()) {
          for (int j = 0; j < n_mma; ++j) mma_ts(tmu, ta, db, (mix && !(j & 1)) ? idesc2 : idesc, true);
          if (n_commit > 0) umma_commit(smem_u32(&bar_c[0]));
          if (n_commit > 1) umma_commit(smem_u32(&bar_c[1]));
        }
      }
      __syncwarp();
    } else {  // 12 MMAs fully unrolled (as the engine issues them)
      for (int i = 0; i < iters; ++i) {
        if (do_wait) mbar_wait(smem_u32(&bar_ready), 0);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 12; ++j) mma_ts(tmu, ta, db, idesc, true);
          if (n_commit > 0) umma_commit(smem_u32(&bar_c[0]));
          if (n_commit > 1) umma_commit(smem_u32(&bar_c[1]));
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  {
    long long* out;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    struct L { int n_mma, do_wait, n_commit, mix, mode = 0; };
    const L ls[] = {{12, 0, 0, 0}, {12, 1, 0, 0}, {12, 0, 1, 0}, {12, 0, 2, 0}, {12, 1, 2, 0}, {8, 1, 2, 1}, {8, 0, 0, 1},
                    {24, 1, 2, 0}, {4, 1, 2, 0}, {1, 1, 2, 0}, {1, 0, 0, 0}, {1, 0, 1, 0},
                    {12, 1, 2, 0, 1}, {1, 0, 0, 0, 1}, {1, 1, 2, 0, 1}, {12, 1, 2, 0, 2}, {1, 0, 0, 0, 2}, {12, 1, 2, 0, 3}, {1, 0, 0, 0, 3},
                    {12, 1, 2, 0, 4}, {12, 0, 0, 0, 4}};
    for (const L& l : ls) {
      k_loop<<<148, 128, 64 * 1024>>>(1024, l.n_mma, l.do_wait, l.n_commit, l.mix, l.mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      printf("loop mode %d: %2d MMAs (%s) wait=%d commits=%d : %7.1f cycles/iteration %s\n", l.mode, l.n_mma, l.mix ? "N=128/64 mix" : "N=64", l.do_wait,
             l.n_commit, (double)cyc / 1024, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }

  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  struct Cfg { int N, tf32, mode, b_mn, n_acc; const char* name; int M = 128; };
  const Cfg cfgs[] = {
      {64, 1, 0, 0, 1, "tf32 SS N=64  B K-major   1 acc"}, {64, 1, 0, 1, 1, "tf32 SS N=64  B MN-major  1 acc"},
      {64, 1, 1, 1, 1, "tf32 TS N=64  B MN-major  1 acc"}, {64, 1, 1, 1, 4, "tf32 TS N=64  B MN-major  4 acc"},
      {64, 1, 1, 0, 1, "tf32 TS N=64  B K-major   1 acc"}, {128, 1, 0, 0, 1, "tf32 SS N=128 B K-major   1 acc"},
      {256, 1, 0, 0, 1, "tf32 SS N=256 B K-major   1 acc"}, {64, 0, 0, 0, 1, "bf16 SS N=64  B K-major   1 acc"},
      {64, 0, 1, 1, 1, "bf16 TS N=64  B MN-major  1 acc"}, {64, 0, 1, 0, 1, "bf16 TS N=64  B K-major   1 acc"},
      {256, 0, 0, 0, 1, "bf16 SS N=256 B K-major   1 acc"}, {256, 0, 0, 1, 1, "bf16 SS N=256 B MN-major  1 acc"},
      {128, 1, 1, 1, 1, "tf32 TS N=128 B MN-major  1 acc"}, {128, 1, 1, 1, 2, "tf32 TS N=128 B MN-major  2 acc"},
      {192, 1, 1, 1, 1, "tf32 TS N=192 B MN-major  1 acc"}, {256, 1, 1, 1, 1, "tf32 TS N=256 B MN-major  1 acc"},
      {32, 1, 1, 1, 1, "tf32 TS N=32  B MN-major  1 acc"}, {16, 1, 1, 1, 1, "tf32 TS N=16  B MN-major  1 acc"},
      {64, 1, 1, 1, 1, "tf32 TS N=64  M=64 MN-major 1 acc", 64}, {64, 1, 1, 1, 4, "tf32 TS N=64  M=64 MN-major 4 acc", 64},
      {128, 1, 1, 1, 1, "tf32 TS N=128 M=64 MN-major 1 acc", 64}, {256, 1, 1, 1, 1, "tf32 TS N=256 M=64 MN-major 1 acc", 64},
      {256, 1, 0, 1, 1, "tf32 SS N=256 M=64 MN-major 1 acc", 64}, {64, 1, 0, 1, 4, "tf32 SS N=64  B MN-major  4 acc"},
      {128, 0, 1, 1, 1, "bf16 TS N=128 B MN-major  1 acc"}, {256, 0, 1, 1, 1, "bf16 TS N=256 B MN-major  1 acc"},
      {256, 0, 1, 0, 1, "bf16 TS N=256 B K-major   1 acc"}, {128, 0, 1, 0, 2, "bf16 TS N=128 B K-major   2 acc"},
  };
  for (const Cfg& c : cfgs) {
    for (int grid : {1, 148}) {
      k_rate<<<grid, 128, 64 * 1024>>>(iters, c.N, c.tf32, c.mode, c.b_mn, c.n_acc, c.M, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      const int K = c.tf32 ? 8 : 16;
      printf("%s grid=%3d: %7.1f cycles/MMA  (%5.0f flop/clk/SM)  %s\n", c.name, grid, (double)cyc / iters,
             2.0 * c.M * c.N * K * iters / (double)cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
