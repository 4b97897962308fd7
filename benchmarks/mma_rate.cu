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
__global__ void __launch_bounds__(128, 1) k_rate(int iters, int N, int tf32, int mode, int b_mn, int n_acc, long long* out) {
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
                           ((uint32_t)(128 >> 4) << 24);
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

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  struct Cfg { int N, tf32, mode, b_mn, n_acc; const char* name; };
  const Cfg cfgs[] = {
      {64, 1, 0, 0, 1, "tf32 SS N=64  B K-major   1 acc"}, {64, 1, 0, 1, 1, "tf32 SS N=64  B MN-major  1 acc"},
      {64, 1, 1, 1, 1, "tf32 TS N=64  B MN-major  1 acc"}, {64, 1, 1, 1, 4, "tf32 TS N=64  B MN-major  4 acc"},
      {64, 1, 1, 0, 1, "tf32 TS N=64  B K-major   1 acc"}, {128, 1, 0, 0, 1, "tf32 SS N=128 B K-major   1 acc"},
      {256, 1, 0, 0, 1, "tf32 SS N=256 B K-major   1 acc"}, {64, 0, 0, 0, 1, "bf16 SS N=64  B K-major   1 acc"},
      {64, 0, 1, 1, 1, "bf16 TS N=64  B MN-major  1 acc"}, {64, 0, 1, 0, 1, "bf16 TS N=64  B K-major   1 acc"},
      {256, 0, 0, 0, 1, "bf16 SS N=256 B K-major   1 acc"}, {256, 0, 0, 1, 1, "bf16 SS N=256 B MN-major  1 acc"},
  };
  for (const Cfg& c : cfgs) {
    for (int grid : {1, 148}) {
      k_rate<<<grid, 128, 64 * 1024>>>(iters, c.N, c.tf32, c.mode, c.b_mn, c.n_acc, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      const int K = c.tf32 ? 8 : 16;
      printf("%s grid=%3d: %7.1f cycles/MMA  (%5.0f flop/clk/SM)  %s\n", c.name, grid, (double)cyc / iters,
             2.0 * 128 * c.N * K * iters / (double)cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
