// Dense Reduce + Connect + auxiliary losses (MinCut / DiffPool), forward and backward.
// Reference: tgp/reduce/base_reduce.py:158-161 (S^T X), tgp/connect/dense_conn.py:112-122 ((S^T A) S),
// tgp/utils/ops.py:282-335 (post-processing), tgp/utils/losses.py:39-123,476-500,644-708 (losses),
// call order from tgp/poolers/mincut.py:219-237 and tgp/poolers/diffpool.py:208-218.
//
// This translation unit holds the shape-general path: a strided batched GEMM on the FP32 pipe plus the
// fused statistics / post-processing / gradient-assembly kernels.  tc_gemm.cu (engine) and dense_fused.cu (fused
// forward) provide the tcgen05 paths that replace `bgemm` for the supported shapes.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "dense.cuh"
#include "tc_gemm.cuh"

namespace cg = cooperative_groups;

namespace tgp {

// ------------------------------------------------------------------------------------------
// Strided batched GEMM  C[b] (+)= alpha * A[b] (MxKd) * B[b] (KdxN), arbitrary element strides.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, FP32 accumulate.
// ------------------------------------------------------------------------------------------
template <typename TA, typename TB, typename TC>
static __global__ void __launch_bounds__(256)
    k_bgemm(const TA* __restrict__ A, const TB* __restrict__ Bm, TC* __restrict__ C, int M, int N, int Kd, int64_t sAb,
            int64_t sAm, int64_t sAk, int64_t sBb, int64_t sBk, int64_t sBn, int64_t sCb, int64_t sCm, float alpha,
            int accumulate) {
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  int b = blockIdx.z;
  int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const TA* Ab = A + (int64_t)b * sAb;
  const TB* Bb = Bm + (int64_t)b * sBb;
  int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  bool a_kc = (sAk == 1), b_nc = (sBn == 1);
  for (int k0 = 0; k0 < Kd; k0 += 16) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int idx = t + r * 256;
      int m, k;
      if (a_kc) { m = idx >> 4; k = idx & 15; } else { m = idx & 63; k = idx >> 6; }
      float v = 0.f;
      if (m0 + m < M && k0 + k < Kd) v = to_f32<TA>(Ab[(int64_t)(m0 + m) * sAm + (int64_t)(k0 + k) * sAk]);
      As[k][m] = v;
      int n, kk;
      if (b_nc) { n = idx & 63; kk = idx >> 6; } else { n = idx >> 4; kk = idx & 15; }
      float u = 0.f;
      if (n0 + n < N && k0 + kk < Kd) u = to_f32<TB>(Bb[(int64_t)(k0 + kk) * sBk + (int64_t)(n0 + n) * sBn]);
      Bs[kk][n] = u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  TC* Cb = C + (int64_t)b * sCb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      TC* p = Cb + (int64_t)m * sCm + n;
      if (accumulate) v += to_f32<TC>(*p);
      *p = from_f32<TC>(v);
    }
  }
}

template <typename TA, typename TB, typename TC>
int bgemm(const TA* A, const TB* B, TC* C, int batch, int M, int N, int Kd, int64_t sAb, int64_t sAm, int64_t sAk,
          int64_t sBb, int64_t sBk, int64_t sBn, int64_t sCb, int64_t sCm, float alpha, bool accumulate,
          cudaStream_t st) {
  if (batch == 0 || M == 0 || N == 0) return TGPB200_OK;
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64), (unsigned)batch);
  launch("k_bgemm", k_bgemm<TA, TB, TC>, grid, 256, 0, st, A, B, C, M, N, Kd, sAb, sAm, sAk, sBb, sBk, sBn, sCb, sCm, alpha,
                                            accumulate ? 1 : 0);
  return launch_status();
}

// ------------------------------------------------------------------------------------------
// Row statistics of one pass over A and S:  d[b,i] = sum_j A[b,i,j];  ss[b,i] = sum_k S^2;
// a2[b,i] = sum_j A^2;  ent[b,i] = -sum_k S log(S + eps).   One warp per (b, i).
// ------------------------------------------------------------------------------------------
// V consecutive elements <-> registers (V = 4: one 128-bit fp32 / 64-bit bf16 access; V = 1: scalar)
template <int V, typename T>
__device__ __forceinline__ void ldv(const T* p, float (&a)[V]) {
  if constexpr (V == 8 && sizeof(T) == 2) {
    const uint4 x = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) a[2 * q] = __uint_as_float(w[q] << 16), a[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  } else if constexpr (V == 4 && sizeof(T) == 4) {
    const float4 x = *reinterpret_cast<const float4*>(p);
    a[0] = x.x, a[1] = x.y, a[2] = x.z, a[3] = x.w;
  } else if constexpr (V == 4) {
    const uint2 x = *reinterpret_cast<const uint2*>(p);
    a[0] = __uint_as_float(x.x << 16), a[1] = __uint_as_float(x.x & 0xffff0000u);
    a[2] = __uint_as_float(x.y << 16), a[3] = __uint_as_float(x.y & 0xffff0000u);
  } else {
    a[0] = to_f32<T>(p[0]);
  }
}
template <int V, typename T>
__device__ __forceinline__ void stv(T* p, const float (&a)[V]) {
  if constexpr (V == 4 && sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
  } else if constexpr (V == 4) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a[0], a[1]), hi = __floats2bfloat162_rn(a[2], a[3]);
    uint2 x;
    x.x = *reinterpret_cast<uint32_t*>(&lo), x.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = x;
  } else {
    p[0] = from_f32<T>(a[0]);
  }
}
// strip copy global -> shared as fp32, several 128-bit loads in flight per thread
template <int V, typename T>
__device__ __forceinline__ void stage_strip(float* dst, const T* src, int cnt) {
#pragma unroll 4
  for (int i = threadIdx.x * V; i < cnt; i += blockDim.x * V) {
    float a[V];
    ldv<V, T>(src + i, a);
    stv<V, float>(dst + i, a);
  }
}

template <typename T, int V>
static __global__ void k_row_stats(const T* __restrict__ A, const T* __restrict__ S, int64_t rows, int N, int K,
                                   float eps, float* __restrict__ d, float* __restrict__ ss, float* __restrict__ a2,
                                   float* __restrict__ ent) {
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float sd = 0.f, sa2 = 0.f, s2 = 0.f, se = 0.f;
  if (A) {
    const T* a = A + r * N;
#pragma unroll 4
    for (int j = lane * V; j < N; j += 32 * V) {
      float v[V];
      ldv<V, T>(a + j, v);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        sd += v[q];
        sa2 += v[q] * v[q];
      }
    }
  }
  const T* s = S + r * K;
#pragma unroll 2
  for (int k = lane * V; k < K; k += 32 * V) {
    float v[V];
    ldv<V, T>(s + k, v);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      s2 += v[q] * v[q];
      // bf16 inputs carry 8 significant bits: the MUFU-based log is exact enough and frees the issue slots
      se -= v[q] * (sizeof(T) == 2 ? __logf(v[q] + eps) : logf(v[q] + eps));
    }
  }
  sd = warp_sum(sd), sa2 = warp_sum(sa2), s2 = warp_sum(s2), se = warp_sum(se);
  if (lane == 0) {
    d[r] = sd, ss[r] = s2, a2[r] = sa2, ent[r] = se;
  }
}

// ------------------------------------------------------------------------------------------
// Per-graph kernels run as a thread-block CLUSTER per graph: CTA `rank` owns the row strip
// [rank * rows_per, (rank + 1) * rows_per) of the [K, K] matrices (staged once in its shared memory), and the
// graph-wide reductions are combined through distributed shared memory in a fixed order (bitwise reproducible).
// A cluster of one CTA is the small-K case.
// ------------------------------------------------------------------------------------------
struct GraphCluster {
  cg::cluster_group g;
  unsigned rank, size;
  float* slots;  // kClusterSlots floats in every CTA's shared memory, one per reduction
  int next;
  __device__ GraphCluster(float* s) : g(cg::this_cluster()), slots(s), next(0) {
    rank = g.block_rank();
    size = g.num_blocks();
  }
  __device__ __forceinline__ void sync() {
    if (size > 1) g.sync(); else __syncthreads();
  }
  // combine a per-CTA value (identical in all threads of the CTA) over the cluster
  template <typename Op>
  __device__ __forceinline__ float combine(float v, Op op) {
    if (size == 1) return v;
    const int s = next++;
    if (threadIdx.x == 0) slots[s] = v;
    g.sync();
    float r = *g.map_shared_rank(&slots[s], 0);
    for (unsigned k = 1; k < size; ++k) r = op(r, *g.map_shared_rank(&slots[s], k));
    return r;
  }
  __device__ __forceinline__ float sum(float v, float* red) {
    return combine(block_sum(v, red), [](float a, float b) { return a + b; });
  }
  // n sums with ONE cluster round trip (v[] in/out)
  template <int NV>
  __device__ __forceinline__ void sum_n(float (&v)[NV], float* red) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = block_sum(v[j], red);
    if (size == 1) return;
    const int s = next;
    next += NV;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < NV; ++j) slots[s + j] = v[j];
    }
    g.sync();
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float r = *g.map_shared_rank(&slots[s + j], 0);
      for (unsigned k = 1; k < size; ++k) r += *g.map_shared_rank(&slots[s + j], k);
      v[j] = r;
    }
  }
  __device__ __forceinline__ float max(float v, float* red) {
    return combine(block_max(v, red), [](float a, float b) { return fmaxf(a, b); });
  }
  __device__ __forceinline__ int min_int(int v, int* redi) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v = ::min(v, __shfl_xor_sync(kFull, v, o));
    __syncthreads();
    if (lane == 0) redi[w] = v;
    __syncthreads();
    v = lane < nw ? redi[lane] : INT_MAX;
    for (int o = 16; o > 0; o >>= 1) v = ::min(v, __shfl_xor_sync(kFull, v, o));
    return __float_as_int(
        combine(__int_as_float(v), [](float a, float b) { return __int_as_float(::min(__float_as_int(a), __float_as_int(b))); }));
  }
  // full[v] = sum over CTAs of part[v]  (part lives at the same shared-memory offset in every CTA)
  template <typename O>
  __device__ __forceinline__ void gather_sum(const float* part, int K, O out) {
    sync();
    for (int v = threadIdx.x; v < K; v += blockDim.x) {
      float s = part[v];
      if (size > 1) {
        s = *g.map_shared_rank(part + v, 0);
        for (unsigned k = 1; k < size; ++k) s += *g.map_shared_rank(part + v, k);
      }
      out(v, s);
    }
    __syncthreads();
  }
};
constexpr int kClusterSlots = 24;

// Column sums over the row strip [u0, u1) with all warps busy and 128-byte row reads: warp w accumulates rows
// u0 + w, u0 + w + nw, ... of a 32-column strip, the partials are summed over warps through shared memory
// (fixed order).  `colred` must hold 33 * 32 floats.  f(u, v) is the addend; out(v, sum) consumes the result.
template <typename F, typename O>
__device__ __forceinline__ void block_col_reduce(int K, int u0, int u1, float* colred, F f, O out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c0 = 0; c0 < K; c0 += 32) {
    const int v = c0 + lane;
    float acc = 0.f;
    if (v < K)
      for (int u = u0 + w; u < u1; u += nw) acc += f(u, v);
    colred[w * 33 + lane] = acc;
    __syncthreads();
    if (w == 0 && v < K) {
      float s = 0.f;
      for (int k = 0; k < nw; ++k) s += colred[k * 33 + lane];
      out(v, s);
    }
    __syncthreads();
  }
}

// One pass over a row strip with 128-bit accesses: thread t owns columns 4 * (t % TPR) .. + 3 (TPR = K / 4 threads
// per row, a power of two <= blockDim) and walks rows r0 + t / TPR, + G, ...   f(i, r, c0, p) yields the four addends
// of the strip-local element i = (r - r0) * K + c0.   colout[v] = sum over the strip's rows (all v), rowout[r] = sum
// over the columns (rows r0..r1 only); both in a fixed order.
// colred: 4 * blockDim floats, rowp: (r1 - r0) * max(1, K / 128) floats.
template <bool COLS, bool ROWS, typename F>
__device__ __forceinline__ void strip_sums4(int K, int r0, int r1, float* colred, float* rowp, float* colout,
                                            float* rowout, F f) {
  const int t = threadIdx.x, nt = blockDim.x, lane = t & 31;
  const int TPR = K >> 2, tshift = __ffs(TPR) - 1;
  const int G = nt >> tshift;  // rows in flight
  const int cx = t & (TPR - 1), g = t >> tshift, c0 = cx << 2;
  const int wpr = TPR >> 5;  // warps per row when a row spans whole warps
  float ca[4] = {0.f, 0.f, 0.f, 0.f};
  for (int rb = r0; rb < r1; rb += G) {
    const int r = rb + g;
    const bool on = g < G && r < r1;
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    if (on) f((r - r0) * K + c0, r, c0, p);
    if (COLS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ca[j] += p[j];
    }
    if (ROWS) {
      float v = (p[0] + p[1]) + (p[2] + p[3]);
      if (TPR >= 32) {
        v = warp_sum(v);
        if (lane == 0 && on) rowp[(r - r0) * wpr + (cx >> 5)] = v;
      } else {
        for (int o = TPR >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (cx == 0 && on) rowout[r] = v;
      }
    }
  }
  if (COLS) {
    if (g < G) *reinterpret_cast<float4*>(colred + g * K + c0) = make_float4(ca[0], ca[1], ca[2], ca[3]);
    __syncthreads();
    for (int v = t; v < K; v += nt) {
      float s = 0.f;
      for (int k = 0; k < G; ++k) s += colred[k * K + v];
      colout[v] = s;
    }
  }
  if (ROWS && TPR >= 32) {
    __syncthreads();
    for (int r = r0 + t; r < r1; r += nt) {
      float s = 0.f;
      for (int k = 0; k < wpr; ++k) s += rowp[(r - r0) * wpr + k];
      rowout[r] = s;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Per-graph epilogue (one cluster per graph): loss statistics from the RAW S^T A S and S^T S,
// then post-processing (diag zero -> D^-1/2 A D^-1/2 -> / max|A|) exactly in ops.py:307-333 order.
// stats[b] = {num, den, ||M||_F^2, ortho_b, ||A||_F^2, ent_b, maxnorm m, unused}
// ------------------------------------------------------------------------------------------
template <typename T, int V, bool STAGED>
static __global__ void __launch_bounds__(512)
    k_graph_epilogue(const float* __restrict__ Araw, const float* __restrict__ M, const float* __restrict__ d,
                     const float* __restrict__ ss, const float* __restrict__ a2, const float* __restrict__ ent, int N,
                     int K, uint32_t flags, float eps, T* __restrict__ Apool, float* __restrict__ dvec,
                     float* __restrict__ stats, int32_t* __restrict__ argmax, int rows_per, int rowp_floats) {
  // dq[K] (degree scale, all columns), part[K], rowp[rowp_floats]  (+ the two staged row strips)
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[32];
  __shared__ int redi[32];
  __shared__ __align__(16) float colred[4 * 512];
  __shared__ float slots[kClusterSlots];
  GraphCluster cl(slots);
  float* dq = sm;
  float* part = sm + K;
  float* rowp = sm + 2 * K;
  const int b = blockIdx.x / cl.size, t = threadIdx.x, nt = blockDim.x;
  const int r0 = min(K, (int)cl.rank * rows_per), r1 = min(K, r0 + rows_per);
  const int i0 = r0 * K, cnt = (r1 - r0) * K;  // this CTA's elements are i0 + [0, cnt)
  const bool hasA = Araw != nullptr, hasM = M != nullptr;
  const float* Ar = hasA ? Araw + (int64_t)b * K * K + i0 : nullptr;
  const float* Mg = hasM ? M + (int64_t)b * K * K + i0 : nullptr;
  float* st = stats + (int64_t)b * 8;
  if constexpr (STAGED) {  // the kernel is a chain of short passes over the same strip: keep it on chip
    float* sA = sm + 2 * K + rowp_floats;
    float* sM = sA + rows_per * K;
    if (hasA) stage_strip<V, float>(sA, Ar, cnt);
    if (hasM) stage_strip<V, float>(sM, Mg, cnt);
    Ar = sA;  // unconditional: the later accesses compile to shared-memory loads
    Mg = sM;
    __syncthreads();
  }

  const int kshift = (K & (K - 1)) == 0 ? __ffs(K) - 1 : -1;
  auto row_of = [&](int i) { return r0 + (kshift >= 0 ? (i >> kshift) : (i / K)); };  // i is strip-local
  auto col_of = [&](int i, int r) { return kshift >= 0 ? (i & (K - 1)) : (i - (r - r0) * K); };
  // den, ||A||^2, entropy, trace, ||M||^2: fixed-order reductions, one cluster round trip
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = cl.rank * nt + t; i < N; i += nt * cl.size) {
    int64_t r = (int64_t)b * N + i;
    acc[0] += d[r] * ss[r];
    acc[1] += a2[r];
    acc[2] += ent[r];
  }
  if (hasA)
    for (int r = r0 + t; r < r1; r += nt) acc[3] += Ar[(int64_t)(r - r0) * K + r];
  if (hasM)
    for (int i = t * V; i < cnt; i += nt * V) {
      float mv[V];
      ldv<V, float>(Mg + i, mv);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[4] += mv[j] * mv[j];
    }
  cl.sum_n(acc, red);
  const float den = acc[0], sa2 = acc[1], se = acc[2], num = acc[3], m2 = acc[4];
  float ortho = 0.f;
  if (hasM) {
    const float nM = sqrtf(m2), isk = 1.0f / sqrtf((float)K), inM = 1.0f / nM;
    float u2 = 0.f;
    for (int i = t * V; i < cnt; i += nt * V) {
      float mv[V];
      ldv<V, float>(Mg + i, mv);
      const int r = row_of(i), c = col_of(i, r);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float u = mv[j] * inM - ((r == c + j) ? isk : 0.f);
        u2 += u * u;
      }
    }
    ortho = sqrtf(cl.sum(u2, red));
  }
  if (t == 0 && cl.rank == 0) {
    st[0] = num, st[1] = den, st[2] = m2, st[3] = ortho, st[4] = sa2, st[5] = se, st[6] = 1.f, st[7] = 0.f;
  }
  if (!hasA || !Apool) {
    cl.sync();  // no CTA may retire while a peer can still read its slots
    return;
  }

  bool rsl = flags & TGPB200_REMOVE_SELF_LOOPS, dn = flags & TGPB200_DEGREE_NORM;
  bool tr = flags & TGPB200_ADJ_TRANSPOSE, wn = flags & TGPB200_EDGE_WEIGHT_NORM;
  // Multiply by 1/d_r * 1/d_c instead of dividing twice (ops.py:318-326 divides): ~1.5 ulp from the divided form,
  // far inside the fp32 parity tolerance, and the IEEE divisions were a quarter of this kernel's instructions.
  constexpr bool kRecip = true;
  // degree vector over the diag-zeroed matrix: s_v = column sum (adj_transpose) or row sum
  if (dn) {
    for (int v = t; v < K; v += nt) part[v] = 0.f;
    __syncthreads();
    if (V == 4 && kshift >= 2 && (K >> 2) <= nt) {  // one vectorised pass
      auto addends = [&](int i, int r, int c0, float (&p)[4]) {
        const float4 x = *reinterpret_cast<const float4*>(Ar + i);
        p[0] = x.x, p[1] = x.y, p[2] = x.z, p[3] = x.w;
        if (rsl) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (r == c0 + j) p[j] = 0.f;
        }
      };
      if (tr) strip_sums4<true, false>(K, r0, r1, colred, rowp, part, nullptr, addends);
      else strip_sums4<false, true>(K, r0, r1, colred, rowp, nullptr, part, addends);
    } else if (tr) {  // column sums of the strip (tiled: every warp reads 128-byte row segments)
      block_col_reduce(
          K, r0, r1, colred, [&](int u, int v) { return (rsl && u == v) ? 0.f : Ar[(int64_t)(u - r0) * K + v]; },
          [&](int v, float s) { part[v] = s; });
    } else {  // row sums: one warp per row (rows of other CTAs stay zero)
      int lane = t & 31, w = t >> 5, nw = nt >> 5;
      for (int v = r0 + w; v < r1; v += nw) {
        float s = 0.f;
        for (int u = lane; u < K; u += 32)
          if (!(rsl && u == v)) s += Ar[(int64_t)(v - r0) * K + u];
        s = warp_sum(s);
        if (lane == 0) part[v] = s;
      }
    }
    cl.gather_sum(part, K, [&](int v, float s) {
      const float dv = sqrtf(fmaxf(s, eps));
      dq[v] = kRecip ? 1.0f / dv : dv;
      if (cl.rank == 0) dvec[(int64_t)b * K + v] = s;
    });
  }
  __syncthreads();
  // the V post-processed values starting at strip-local element i (same row)
  auto values = [&](int i, float (&v)[V]) {
    ldv<V, float>(Ar + i, v);
    const int r = row_of(i), c = col_of(i, r);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (rsl && r == c + j) v[j] = 0.f;
      if (dn) {
        if (kRecip) v[j] = v[j] * (dq[r] * dq[c + j]);
        else v[j] = tr ? __fdiv_rn(__fdiv_rn(v[j], dq[c + j]), dq[r]) : __fdiv_rn(__fdiv_rn(v[j], dq[r]), dq[c + j]);
      }
    }
  };
  T* Ap = Apool + (int64_t)b * K * K + i0;
  if (!wn) {
    for (int i = t * V; i < cnt; i += nt * V) {
      float v[V];
      values(i, v);
      stv<V, T>(Ap + i, v);
    }
  } else {
    float mx = 0.f;
    for (int i = t * V; i < cnt; i += nt * V) {
      float v[V];
      values(i, v);
#pragma unroll
      for (int j = 0; j < V; ++j) mx = fmaxf(mx, fabsf(v[j]));
    }
    const float bm = cl.max(mx, red);
    // Arg of the max for the backward: the FIRST index whose magnitude reaches the max (torch.max rule).
    // A symmetric adjacency ties (r,c) with (c,r) up to rounding noise of the GEMM, so "reaches" is taken
    // with a 2e-6 relative slack; the reference then picks the upper-triangle element, and so do we.
    const float thr = bm * (1.f - 2e-6f);
    int cand = INT_MAX;
    for (int i = t * V; i < cnt && cand == INT_MAX; i += nt * V) {
      float v[V];
      values(i, v);
#pragma unroll
      for (int j = V - 1; j >= 0; --j)
        if (fabsf(v[j]) >= thr) cand = i0 + i + j;
    }
    cand = cl.min_int(cand, redi);
    const float m = bm == 0.f ? 1.f : bm;
    if (t == 0 && cl.rank == 0) {
      st[6] = m;
      argmax[b] = bm == 0.f ? -1 : cand;
    }
    for (int i = t * V; i < cnt; i += nt * V) {
      float v[V];
      values(i, v);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = __fdiv_rn(v[j], m);
      stv<V, T>(Ap + i, v);
    }
  }
  cl.sync();
}

// losses[0..3] = {cut (mean_b), ortho (mean_b), link, entropy}; one block, fixed order.
// The link loss ||A - S S^T||_F^2 comes from the identity q = ||A||^2 - 2 tr(S^T A S) + ||S^T S||^2 (no [N, N]
// temporary).  The identity cancels when the residual is small against ||A||: losses[5] = 1 then asks the direct
// residual pass (k_link_residual / k_link_fix) to replace q; `link_tau` is the q / ||A||^2 ratio below which the
// identity no longer holds the tolerance of the operand dtype.
static __global__ void k_finalize_losses(const float* __restrict__ stats, int B, float eps, float link_div,
                                         float ent_div, float link_tau, float* __restrict__ losses) {
  __shared__ float red[32];
  float cut = 0.f, ortho = 0.f, q = 0.f, ent = 0.f, a2 = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* st = stats + (int64_t)b * 8;
    cut += -(st[0] / (st[1] + eps));
    ortho += st[3];
    q += st[4] - 2.f * st[0] + st[2];
    a2 += st[4];
    ent += st[5];
  }
  cut = block_sum(cut, red), ortho = block_sum(ortho, red), q = block_sum(q, red), ent = block_sum(ent, red);
  a2 = block_sum(a2, red);
  if (threadIdx.x == 0) {
    losses[0] = cut / (float)B;
    losses[1] = ortho / (float)B;
    losses[2] = sqrtf(fmaxf(q, 0.f)) / link_div;
    losses[3] = ent / ent_div;
    losses[4] = q;
    losses[5] = (link_tau > 0.f && q < link_tau * a2) ? 1.f : 0.f;
  }
}

// Direct residual of the link loss, only when the identity cancels (losses[5] != 0; every CTA returns at once
// otherwise): partial[b * tiles + t] = sum over a 64 x 64 tile of (A - S S^T)^2 on the FP32 pipe.  Rare path (a
// trained assignment that reproduces A), so shape generality and a fixed reduction order matter, not speed.
template <typename T>
static __global__ void __launch_bounds__(256)
    k_link_residual(const T* __restrict__ A, const T* __restrict__ S, int N, int K, int tiles_n, int64_t total,
                    const float* __restrict__ losses, float* __restrict__ partial) {
  if (losses[5] == 0.f) return;
  __shared__ float si[16][65], sj[16][65];
  __shared__ float red[32];
  const int tiles = tiles_n * tiles_n;
  // grid-stride over the (graph, tile) pairs: the launch is a few CTAs per SM, so the common case (flag clear, every
  // CTA returns at once) costs microseconds instead of the scheduling of B * tiles CTAs (61 us per C3 step)
  for (int64_t wid = blockIdx.x; wid < total; wid += gridDim.x) {
  const int tile = (int)(wid % tiles), b = (int)(wid / tiles);
  const int i0 = (tile / tiles_n) * 64, j0 = (tile % tiles_n) * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 4 x 4 micro-tile per thread
  const T* Sb = S + (int64_t)b * N * K;
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, k = e & 15;
      const bool kin = k0 + k < K;
      si[k][r] = (kin && i0 + r < N) ? to_f32<T>(Sb[(int64_t)(i0 + r) * K + k0 + k]) : 0.f;
      sj[k][r] = (kin && j0 + r < N) ? to_f32<T>(Sb[(int64_t)(j0 + r) * K + k0 + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = si[k][ty * 4 + u], c[u] = sj[k][tx * 4 + u];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], c[v], acc[u][v]);
    }
    __syncthreads();
  }
  const T* Ab = A + (int64_t)b * N * N;
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < N && j < N) {
        const float r = to_f32<T>(Ab[(int64_t)i * N + j]) - acc[u][v];
        s = fmaf(r, r, s);
      }
    }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[wid] = s;
  __syncthreads();
  }
}
static __global__ void k_link_fix(const float* __restrict__ partial, int64_t n, float link_div,
                                  float* __restrict__ losses) {
  if (losses[5] == 0.f) return;
  __shared__ float red[32];
  float q = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) q += partial[i];
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    losses[2] = sqrtf(q) / link_div;
    losses[4] = q;
  }
}

// ------------------------------------------------------------------------------------------
// Backward assembly (one cluster per graph, row strips as in the epilogue):
//   Graw = d L / d (S^T A S) raw   from  Gpool (through max-norm, degree-norm, diag-zero) + trace terms
//   P    = d L / d (S^T S) symmetrised:  dS += S P        (ortho + link ||M||^2 term)
//   coef[b] = {c_den, c_a2, c_ent}  for the element-wise terms
// gl = upstream grads of {cut, ortho, link, entropy} (device floats, already x coefficient).
// ------------------------------------------------------------------------------------------
template <typename T, int V, bool STAGED>
static __global__ void __launch_bounds__(512)
    k_graph_bwd(const float* __restrict__ Araw, const float* __restrict__ M, const T* __restrict__ Gpool,
                const float* __restrict__ dvec, const float* __restrict__ stats, const int32_t* __restrict__ argmax,
                const float* __restrict__ gl, const float* __restrict__ losses, int B, int K, uint32_t flags,
                int loss_kind, float eps, float link_div, float ent_div, float* __restrict__ Graw,
                float* __restrict__ P, float* __restrict__ coef, int rows_per, int rowp_floats, T* __restrict__ Gt,
                T* __restrict__ Pt) {
  // dsq[K] = 1/d_v, part[K], tot[K] = rowdot + coldot, rowv[K], rowp[rowp_floats]  (+ three staged strips)
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[32];
  __shared__ __align__(16) float colred[4 * 512];
  __shared__ float slots[kClusterSlots];
  GraphCluster cl(slots);
  float* dsq = sm;
  float* part = sm + K;
  float* tot = sm + 2 * K;
  float* rowv = sm + 3 * K;
  float* rowp = sm + 4 * K;
  const int b = blockIdx.x / cl.size, t = threadIdx.x, nt = blockDim.x;
  const int r0 = min(K, (int)cl.rank * rows_per), r1 = min(K, r0 + rows_per);
  const int i0 = r0 * K, cnt = (r1 - r0) * K;
  const float* st = stats + (int64_t)b * 8;
  const bool hasM = M != nullptr, hasG = Gpool != nullptr;
  const float* Ar = Araw + (int64_t)b * K * K + i0;
  const float* Mb = hasM ? M + (int64_t)b * K * K + i0 : nullptr;
  const T* Gp = hasG ? Gpool + (int64_t)b * K * K + i0 : nullptr;
  const float* Gs = nullptr;  // staged fp32 copy of the upstream gradient strip
  if constexpr (STAGED) {
    float* sA = sm + 4 * K + rowp_floats;
    float* sM = sA + rows_per * K;
    float* sG = sM + rows_per * K;
    stage_strip<V, float>(sA, Ar, cnt);
    if (hasM) stage_strip<V, float>(sM, Mb, cnt);
    if (hasG) stage_strip<V, T>(sG, Gp, cnt);
    Ar = sA;  // unconditional: the later accesses compile to shared-memory loads
    Mb = sM;
    Gs = sG;
    __syncthreads();
  }
  auto gp = [&](int i) {
    if constexpr (STAGED) return Gs[i]; else return to_f32<T>(Gp[i]);
  };
  auto gpv = [&](int i, float (&g)[V]) {
    if constexpr (STAGED) ldv<V, float>(Gs + i, g); else ldv<V, T>(Gp + i, g);
  };
  auto outv = [&](T* ot, float* of, int i, const float (&v)[V]) {  // operand-dtype copy (bf16 runs) or fp32
    if (ot) stv<V, T>(ot + i, v); else stv<V, float>(of + i, v);
  };
  const int64_t o0 = (int64_t)b * K * K + i0;  // output offset of this strip
  bool rsl = flags & TGPB200_REMOVE_SELF_LOOPS, dn = flags & TGPB200_DEGREE_NORM;
  bool tr = flags & TGPB200_ADJ_TRANSPOSE, wn = flags & TGPB200_EDGE_WEIGHT_NORM;
  float m = st[6];

  // --- loss coefficients
  float g_cut = gl ? gl[0] : 0.f, g_ortho = gl ? gl[1] : 0.f, g_link = gl ? gl[2] : 0.f, g_ent = gl ? gl[3] : 0.f;
  float c_num = 0.f, c_den = 0.f, c_a2 = 0.f, c_m2 = 0.f, c_ent = 0.f;
  if (loss_kind == 1) {
    float dd = st[1] + eps;
    c_num = -g_cut / ((float)B * dd);
    c_den = g_cut * st[0] / ((float)B * dd * dd);
  } else if (loss_kind == 2) {
    float q = losses[4];
    float L = sqrtf(fmaxf(q, 0.f));
    float cq = L > 0.f ? g_link / (2.f * L * link_div) : 0.f;  // dL/dq
    c_a2 = cq, c_num = -2.f * cq, c_m2 = cq;
    c_ent = g_ent / ent_div;
  }
  if (t == 0 && cl.rank == 0) {
    coef[b * 4 + 0] = c_den, coef[b * 4 + 1] = c_a2, coef[b * 4 + 2] = c_ent, coef[b * 4 + 3] = 0.f;
  }

  // index split without an integer division when K is a power of two (the common case); i is strip-local
  const int kshift = (K & (K - 1)) == 0 ? __ffs(K) - 1 : -1;
  auto row_of = [&](int i) { return r0 + (kshift >= 0 ? (i >> kshift) : (i / K)); };
  auto col_of = [&](int i, int r) { return kshift >= 0 ? (i & (K - 1)) : (i - (r - r0) * K); };

  // --- P = dL/dM + transpose  (M symmetric so both terms are symmetric)
  if (P && hasM) {
    if (loss_kind == 1) {
      const float nM = sqrtf(st[2]), rr = st[3], isk = 1.0f / sqrtf((float)K), inM = 1.0f / nM;
      // With U = M/nM - I/sqrt(K) and r = ||U||:  <U, M> = nM * <U, M/nM> = nM * r^2 / 2  (no extra reduction,
      // and no cancellation near perfect orthogonality).
      const float um = 0.5f * nM * rr * rr;
      const float go = g_ortho / (float)B;
      // dM = go * (U/r - <U,M>/r * M/nM^2) / nM
      const float c1 = rr > 0.f ? 2.f * go / (rr * nM) : 0.f, c2 = rr > 0.f ? 2.f * go * um / (rr * nM * nM * nM) : 0.f;
      for (int i = t * V; i < cnt; i += nt * V) {
        float mv[V], pv[V];
        ldv<V, float>(Mb + i, mv);
        const int r = row_of(i), c = col_of(i, r);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float u = mv[j] * inM - ((r == c + j) ? isk : 0.f);
          pv[j] = c1 * u - c2 * mv[j];
        }
        outv(Pt ? Pt + o0 : nullptr, P + o0, i, pv);
      }
    } else {
      for (int i = t * V; i < cnt; i += nt * V) {
        float mv[V], pv[V];
        ldv<V, float>(Mb + i, mv);
#pragma unroll
        for (int j = 0; j < V; ++j) pv[j] = 4.f * c_m2 * mv[j];
        outv(Pt ? Pt + o0 : nullptr, P + o0, i, pv);
      }
    }
  }

  // --- Graw from Gpool
  if (!hasG) {
    for (int i = t * V; i < cnt; i += nt * V) {
      const int r = row_of(i), c = col_of(i, r);
      float gv[V];
#pragma unroll
      for (int j = 0; j < V; ++j) gv[j] = (r == c + j) ? c_num : 0.f;
      outv(Gt ? Gt + o0 : nullptr, Graw + o0, i, gv);
    }
    return;  // no cluster reduction was started on this path
  }
  // dsq holds 1 / d_v = 1 / sqrt(clamp(s_v, eps)): the gradient assembly multiplies by reciprocals (the kernel
  // was bound by IEEE divisions); results agree with the divided form to ~1 ulp.
  if (dn)
    for (int v = t; v < K; v += nt) dsq[v] = 1.0f / sqrtf(fmaxf(dvec[(int64_t)b * K + v], eps));
  __syncthreads();
  const float im = 1.0f / m;
  // A2 = normalised matrix before max-norm; G2 = grad wrt A2
  float corr = 0.f;  // sum G * A2 (for the max-norm arg term)
  if (wn && m != 0.f) {
    for (int i = t * V; i < cnt; i += nt * V) {
      float av[V], gv[V];
      ldv<V, float>(Ar + i, av);
      gpv(i, gv);
      const int r = row_of(i), c = col_of(i, r);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float v = (rsl && r == c + j) ? 0.f : av[j];
        if (dn) v = v * (dsq[r] * dsq[c + j]);
        corr += gv[j] * v;
      }
    }
  }
  if (wn) corr = cl.sum(corr, red);
  const int am = wn ? argmax[b] - i0 : -1;  // strip-local index of the arg-max element (out of range elsewhere)
  const float argterm = -corr * im * im;
  // tot[v] = sum_j G2[v,j] A2[v,j] + sum_i G2[i,v] A2[i,v]   (row dot + column dot)
  if (dn) {
    for (int v = t; v < K; v += nt) part[v] = 0.f;
    __syncthreads();
    if (V == 4 && kshift >= 2 && (K >> 2) <= nt) {  // row and column dots in one vectorised pass
      strip_sums4<true, true>(K, r0, r1, colred, rowp, part, rowv, [&](int i, int r, int c0, float (&p)[4]) {
        float av[4], gv[4];
        ldv<4, float>(Ar + i, av);
        if constexpr (STAGED) ldv<4, float>(Gs + i, gv); else ldv<4, T>(Gp + i, gv);
        const float ir = dsq[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = (rsl && r == c0 + j) ? 0.f : av[j] * (ir * dsq[c0 + j]);
          float g = gv[j] * im;
          if (i + j == am) g += (a < 0.f ? -argterm : argterm);
          p[j] = g * a;
        }
      });
      for (int v = r0 + t; v < r1; v += nt) part[v] += rowv[v];
    } else {
      block_col_reduce(  // column dots over the strip (tiled)
          K, r0, r1, colred,
          [&](int u, int v) {
            int i = (u - r0) * K + v;
            float a = (rsl && u == v) ? 0.f : Ar[i] * (dsq[v] * dsq[u]);
            float g = gp(i) * im;
            if (i == am) g += (a < 0.f ? -argterm : argterm);
            return g * a;
          },
          [&](int v, float sc) { part[v] = sc; });
      int lane = t & 31, w = t >> 5, nw = nt >> 5;
      for (int v = r0 + w; v < r1; v += nw) {  // row dots: one warp per row of the strip
        float sr = 0.f;
        const float iv = dsq[v];
        for (int u = lane; u < K; u += 32) {
          int i = (v - r0) * K + u;
          float a = (rsl && u == v) ? 0.f : Ar[i] * (iv * dsq[u]);
          float g = gp(i) * im;
          if (i == am) g += (a < 0.f ? -argterm : argterm);
          sr += g * a;
        }
        sr = warp_sum(sr);
        if (lane == 0) part[v] += sr;
      }
    }
    cl.gather_sum(part, K, [&](int v, float s) { tot[v] = s; });
  }
  __syncthreads();
  for (int i = t * V; i < cnt; i += nt * V) {
    float av[V], gv[V], ov[V];
    ldv<V, float>(Ar + i, av);
    gpv(i, gv);
    const int r = row_of(i), c0 = col_of(i, r);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = c0 + j;
      float g = gv[j] * im;
      if (wn && i + j == am) {
        float a = (rsl && r == c) ? 0.f : av[j];
        if (dn) a = a * (dsq[r] * dsq[c]);
        g += (a < 0.f ? -argterm : argterm);
      }
      float out = g;
      if (dn) {
        out = g * (dsq[r] * dsq[c]);
        int v = tr ? c : r;  // s_v is a column sum (transpose) or a row sum
        float sv = dvec[(int64_t)b * K + v];
        // dL/d d_v = -(rdot + cdot) / d_v ;  d d_v / d s_v = 1 / (2 d_v)   (clamp passes the gradient for s_v >= eps)
        if (sv >= eps) out += -0.5f * tot[v] * dsq[v] * dsq[v];
      }
      if (rsl && r == c) out = 0.f;
      if (r == c) out += c_num;
      ov[j] = out;
    }
    outv(Gt ? Gt + o0 : nullptr, Graw + o0, i, ov);
  }
  cl.sync();  // no CTA may retire while a peer can still read its shared memory
}

// dS += element-wise terms:  c_den * 2 d_i S_ik  +  c_ent * (-log(S+eps) - S/(S+eps)).
template <typename T>
static __global__ void k_ds_elementwise(const T* __restrict__ S, const float* __restrict__ d,
                                        const float* __restrict__ coef, int64_t total, int N, int K, float eps,
                                        T* __restrict__ dS) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t row = i / K;
  int b = (int)(row / N);
  float c_den = coef[b * 4 + 0], c_ent = coef[b * 4 + 2];
  if (c_den == 0.f && c_ent == 0.f) return;
  float s = to_f32<T>(S[i]);
  float v = to_f32<T>(dS[i]);
  if (c_den != 0.f) v += c_den * 2.f * d[row] * s;
  if (c_ent != 0.f) v += c_ent * (-logf(s + eps) - s / (s + eps));
  dS[i] = from_f32<T>(v);
}

// dA += c_den * ss_i (row broadcast) + c_a2 * 2 A_ij
template <typename T>
static __global__ void k_da_elementwise(const T* __restrict__ A, const float* __restrict__ ss,
                                        const float* __restrict__ coef, int64_t total, int N, T* __restrict__ dA) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t row = i / N;
  int b = (int)(row / N);
  float c_den = coef[b * 4 + 0], c_a2 = coef[b * 4 + 1];
  if (c_den == 0.f && c_a2 == 0.f) return;
  float v = to_f32<T>(dA[i]);
  if (c_den != 0.f) v += c_den * ss[row];
  if (c_a2 != 0.f) v += c_a2 * 2.f * to_f32<T>(A[i]);
  dA[i] = from_f32<T>(v);
}

// 128-bit form of the above (N a multiple of the vector width, 16-byte aligned pointers): one row / graph lookup per
// vector, streaming loads and stores (dA and A are touched once)
template <typename T>
static __global__ void k_da_elementwise_vec(const T* __restrict__ A, const float* __restrict__ ss,
                                            const float* __restrict__ coef, int64_t total_vec, int N,
                                            T* __restrict__ dA) {
  constexpr int V = 16 / sizeof(T);
  int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= total_vec) return;
  const int64_t i = iv * V;
  const int64_t row = i / N;
  const int b = (int)(row / N);
  const float c_den = coef[b * 4 + 0], c_a2 = coef[b * 4 + 1];
  if (c_den == 0.f && c_a2 == 0.f) return;
  const float add = c_den != 0.f ? c_den * ss[row] : 0.f;
  uint4 gv = __ldcs(reinterpret_cast<const uint4*>(dA + i));
  uint4 av = c_a2 != 0.f ? __ldcs(reinterpret_cast<const uint4*>(A + i)) : make_uint4(0, 0, 0, 0);
  T* g = reinterpret_cast<T*>(&gv);
  const T* a = reinterpret_cast<const T*>(&av);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    float v = to_f32<T>(g[k]) + add;
    if (c_a2 != 0.f) v += c_a2 * 2.f * to_f32<T>(a[k]);
    g[k] = from_f32<T>(v);
  }
  __stcs(reinterpret_cast<uint4*>(dA + i), gv);
}

template <typename T>
static __global__ void k_cast_out(const float* __restrict__ in, T* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = from_f32<T>(in[i]);
}

// ------------------------------------------------------------------------------------------
// Host orchestration
// ------------------------------------------------------------------------------------------
// A GEMM operand as stored in memory (per batch item, cols contiguous):
//   K-major : [MN extent rows][K extent cols];   MN-major: [K extent rows][MN extent cols]
struct Mat {
  const void* ptr;
  int64_t bs, rs;
  int mn;
};

// D[b] (M x N) = sum_p A_p B_p, written with (row, col) strides.  Tensor-core engine when the layout
// satisfies the TMA constraints, FP32-pipe batched GEMM otherwise (shape generality, not a second backend:
// both paths live in this library and compute the same product).
struct EwHook {  // fused element-wise terms of dS (tc::GemmProblem::ew_*); *applied is set when the engine took it
  const void* S = nullptr;
  const float* d = nullptr;
  const float* coef = nullptr;
  float eps = 0.f;
  bool* applied = nullptr;
};

template <typename T, typename TC>
static int mm(int batch, int M, int N, int npairs, const int* kd, const Mat* A, const Mat* Bm, TC* out, int64_t obs,
              int64_t ors, int64_t ocs, cudaStream_t st, EwHook hook = EwHook(), const char* tag = nullptr) {
  tc::GemmProblem p;
  memset(&p, 0, sizeof(p));
  p.tag = tag;
  p.batch = batch, p.M = M, p.N = N, p.num_pairs = npairs;
  for (int i = 0; i < npairs; ++i) {
    p.kd[i] = kd[i];
    p.a[i] = {A[i].ptr, A[i].bs, A[i].rs, A[i].mn};
    p.b[i] = {Bm[i].ptr, Bm[i].bs, Bm[i].rs, Bm[i].mn};
  }
  p.out = out, p.out_batch_stride = obs, p.out_row_stride = ors, p.out_col_stride = ocs;
  p.alpha = 1.f, p.accumulate = 0;
  p.out_bf16 = std::is_same<TC, __nv_bfloat16>::value;
  p.in_bf16 = std::is_same<T, __nv_bfloat16>::value;
  if (tc::gemm_supported(p)) {
    if (hook.S && ocs == 1 && ors == N) {
      p.ew_S = hook.S, p.ew_d = hook.d, p.ew_coef = hook.coef, p.ew_eps = hook.eps;
      if (hook.applied) *hook.applied = true;
    }
    return tc::gemm(p, st);
  }
  if (ocs != 1) {  // transposed output: D^T = B^T A^T has unit column stride
    if (ors != 1) return TGPB200_ERR_UNSUPPORTED;
    return mm<T, TC>(batch, N, M, npairs, kd, Bm, A, out, obs, ocs, ors, st, EwHook(), tag);
  }
  for (int i = 0; i < npairs; ++i) {
    int64_t sAm = A[i].mn ? 1 : A[i].rs, sAk = A[i].mn ? A[i].rs : 1;
    int64_t sBk = Bm[i].mn ? Bm[i].rs : 1, sBn = Bm[i].mn ? 1 : Bm[i].rs;
    int rc = bgemm<T, T, TC>((const T*)A[i].ptr, (const T*)Bm[i].ptr, out, batch, M, N, kd[i], A[i].bs, sAm, sAk,
                             Bm[i].bs, sBk, sBn, obs, ors, 1.f, i > 0, st);
    if (rc) return rc;
  }
  return TGPB200_OK;
}

template <typename T, typename TC>
static int mm1(int batch, int M, int N, int kd, Mat A, Mat Bm, TC* out, int64_t obs, int64_t ors, int64_t ocs,
               cudaStream_t st, const char* tag = nullptr) {
  return mm<T, TC>(batch, M, N, 1, &kd, &A, &Bm, out, obs, ors, ocs, st, EwHook(), tag);
}

// How a per-graph kernel splits a [K, K] matrix over a cluster: R CTAs of `threads` threads, `rows_per` rows each;
// the strips are staged in shared memory when `nvec` K-vectors plus `nstrips` strips fit.
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
struct StripPlan {
  int R, rows_per, threads, stage, rowp_floats;
  size_t smem;
};
static StripPlan strip_plan(int K, int nvec, int nstrips, int64_t elems_for_512 = 2048) {
  // tuning knobs (read per call; benchmarks/per_graph_kernels.py sweeps them)
  const char* e = getenv("TGPB200_STRIP_ELEMS");
  const int ev = e ? atoi(e) : 0;
  const int target = ev > 0 ? ev : 32768;  // measured best on B200 (K = 256: two CTAs per graph)
  const char* es = getenv("TGPB200_STRIP_STAGE");
  const bool allow_stage = es && es[0] == '1';  // staging measured no faster than L2-served re-reads
  StripPlan p;
  p.R = 1;
  while (p.R < 8 && ceil_div(K, p.R) * (int64_t)K > target) p.R *= 2;
  p.rows_per = (int)ceil_div(K, p.R);
  p.rowp_floats = (int)(ceil_div((int64_t)p.rows_per * (K >= 128 ? K / 128 : 1), 4) * 4);
  const size_t vec = ((size_t)nvec * K + p.rowp_floats) * sizeof(float);
  const size_t strips = (size_t)nstrips * p.rows_per * K * sizeof(float);
  p.stage = allow_stage && vec + strips <= 160 * 1024;
  p.smem = vec + (p.stage ? strips : 0);
  {
    const char* et = getenv("TGPB200_STRIP_THREADS");
    const int tv = et ? atoi(et) : 0;
    p.threads = (tv == 128 || tv == 256 || tv == 512) ? tv : ((int64_t)p.rows_per * K >= elems_for_512 ? 512 : 256);
  }
  return p;
}

template <typename T>
struct DensePlan {
  T* Tt;  // T = S^T A  [B, K, N]
  float *Araw, *M, *d, *ss, *a2, *ent, *dvec, *stats, *losses;
  int32_t* argmax;
  bool ok;
  DensePlan(Workspace& ws, int64_t B, int64_t N, int64_t K) {
    Tt = ws.take<T>((size_t)B * N * K);
    Araw = ws.take<float>((size_t)B * K * K);
    M = ws.take<float>((size_t)B * K * K);
    d = ws.take<float>((size_t)B * N);
    ss = ws.take<float>((size_t)B * N);
    a2 = ws.take<float>((size_t)B * N);
    ent = ws.take<float>((size_t)B * N);
    dvec = ws.take<float>((size_t)B * K);
    stats = ws.take<float>((size_t)B * 8);
    losses = ws.take<float>(8);
    argmax = ws.take<int32_t>((size_t)B);
    ok = ws.ok;
  }
};

static size_t dense_saved_bytes(int64_t B, int64_t N, int64_t K) {
  size_t f = sizeof(float);
  return align_up((size_t)B * K * N * f) + align_up((size_t)B * K * K * f) * 2 + align_up((size_t)B * N * f) * 4 +
         align_up((size_t)B * K * f) + align_up((size_t)B * 8 * f) + align_up(8 * f) + align_up((size_t)B * 4) +
         align_up((size_t)B * ((N + 63) / 64) * ((N + 63) / 64) * f) + 4096;
}

template <typename T>
static int dense_fwd(const T* A, const T* S, const T* X, int B, int N, int K, int F, uint32_t flags, int loss_kind,
                     float eps, float link_div, float ent_div, T* Xpool, T* Apool, float* losses_out, Workspace& saved,
                     cudaStream_t st) {
  DensePlan<T> pl(saved, B, N, K);
  float* link_partial = saved.take<float>((size_t)B * ((N + 63) / 64) * ((N + 63) / 64));
  if (!pl.ok || !saved.ok) return TGPB200_ERR_WORKSPACE;
  int rc;
  int64_t NK = (int64_t)N * K, NN = (int64_t)N * N, NF = (int64_t)N * F, KK = (int64_t)K * K, KF = (int64_t)K * F;
  const Mat Smn{S, NK, K, 1};  // S read with the node index as the contraction dim
  bool fused = false;
  if (A && X && Xpool && B > 0) {
    // one pass over A, X, S: Tt, X_pool, M and the row statistics (dense_fused.cu)
    rc = TGPB200_ERR_UNSUPPORTED;
    if constexpr (std::is_same<T, float>::value) {
      const char* e = getenv("TGPB200_FUSED_TS");
      if (!(e && e[0] == '0'))
        rc = tc::dense_fwd_fused_ts(A, S, X, B, N, K, F, eps, pl.Tt, Xpool, pl.M, pl.d, pl.ss, pl.a2, pl.ent, st);
    }
    if (rc == TGPB200_ERR_UNSUPPORTED)
      rc = tc::dense_fwd_fused(A, S, X, B, N, K, F, std::is_same<T, __nv_bfloat16>::value, eps, pl.Tt, Xpool, pl.M,
                               pl.d, pl.ss, pl.a2, pl.ent, st);
    if (rc == TGPB200_OK) fused = true;
    else if (rc != TGPB200_ERR_UNSUPPORTED) return rc;
  }
  if (!fused) {
    if (X && Xpool) {
      // X_pool = S^T X.  The larger of (F, K) becomes the 128-row MMA dimension.
      // (transposed stores are scalar per element; when K already fills the 128-row tile the natural orientation
      //  with vectorised row stores is the faster one)
      if (F >= K && K % 128 != 0)  // D[f, k] = sum_i X[i, f] S[i, k], written transposed into X_pool[k, f]
        rc = mm1<T, T>(B, F, K, N, Mat{X, NF, F, 1}, Smn, Xpool, KF, 1, F, st, "k_tc_gemm:Xpool=StX");
      else
        rc = mm1<T, T>(B, K, F, N, Smn, Mat{X, NF, F, 1}, Xpool, KF, F, 1, st, "k_tc_gemm:Xpool=StX");
      if (rc) return rc;
    }
    if (A) {  // T = S^T A  [K, N]
      if (K % 128 == 0)  // natural orientation: D[k, j] = sum_i S[i, k] A[i, j]
        rc = mm1<T, T>(B, K, N, N, Smn, Mat{A, NN, N, 1}, pl.Tt, NK, N, 1, st, "k_tc_gemm:T=StA");
      else  // D[j, k] = sum_i A[i, j] S[i, k], written transposed (no padding of K up to the 128-row tile)
        rc = mm1<T, T>(B, N, K, N, Mat{A, NN, N, 1}, Smn, pl.Tt, NK, 1, N, st, "k_tc_gemm:T=StA");
      if (rc) return rc;
    }
    if (loss_kind != 0) {  // M = S^T S
      rc = mm1<T, float>(B, K, K, N, Smn, Smn, pl.M, KK, K, 1, st, "k_tc_gemm:M=StS");
      if (rc) return rc;
    }
    int64_t rows = (int64_t)B * N;
    if (rows > 0) {
      constexpr int kVec = sizeof(T) == 2 ? 8 : 4;  // 128-bit loads
      launch("k_row_stats",
             (N % kVec == 0 && K % kVec == 0 && aligned16(A) && aligned16(S)) ? k_row_stats<T, kVec> : k_row_stats<T, 1>,
             (unsigned)ceil_div(rows * 32, 256), 256, 0, st, A, S, rows, N, K, eps, pl.d, pl.ss, pl.a2, pl.ent);
    }
  }
  if (A) {  // A_raw = T S  [K, K]   ((S^T A) S, dense_conn.py:120-121)
    rc = mm1<T, float>(B, K, K, N, Mat{pl.Tt, NK, N, 0}, Smn, pl.Araw, KK, K, 1, st, "k_tc_gemm:Araw=TS");
    if (rc) return rc;
  }
  if (B > 0) {
    // 256 threads up to 8 K elements per CTA: measured 16.6 vs 21.1 us on C2 (the backward kernel prefers 512)
    const StripPlan sp = strip_plan(K, 2, 2, 8193);
    const bool v4 = K % 4 == 0 && aligned16(Apool);
    auto kern = sp.stage ? (v4 ? k_graph_epilogue<T, 4, true> : k_graph_epilogue<T, 1, true>)
                         : (v4 ? k_graph_epilogue<T, 4, false> : k_graph_epilogue<T, 1, false>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
    launch_cluster("k_graph_epilogue", kern, (unsigned)B * sp.R, sp.threads, sp.R, sp.smem, st,
                   A ? pl.Araw : (float*)nullptr, loss_kind != 0 ? pl.M : (float*)nullptr, pl.d, pl.ss, pl.a2, pl.ent,
                   N, K, flags, eps, Apool, pl.dvec, pl.stats, pl.argmax, sp.rows_per, sp.rowp_floats);
    // fp32: the identity holds rtol 1e-5 down to q ~ 0.05 ||A||^2; bf16 (T stored in bf16, rtol 2e-2) down to 0.25
    const float link_tau = (loss_kind == 2 && A) ? (std::is_same<T, float>::value ? 0.05f : 0.25f) : 0.f;
    launch("k_finalize_losses", k_finalize_losses, 1, 256, 0, st, pl.stats, B, eps, link_div, ent_div, link_tau,
           pl.losses);
    if (link_tau > 0.f) {
      const int tn = (N + 63) / 64;
      const int64_t total = (int64_t)tn * tn * B;
      const int64_t cap = (int64_t)tc::device_sm_count() * 8;
      launch("k_link_residual", k_link_residual<T>, (unsigned)(total < cap ? total : cap), 256, 0, st, A, S, N, K, tn, total,
             pl.losses, link_partial);
      launch("k_link_fix", k_link_fix, 1, 1024, 0, st, link_partial, (int64_t)B * tn * tn, link_div, pl.losses);
    }
    if (losses_out) cudaMemcpyAsync(losses_out, pl.losses, 4 * sizeof(float), cudaMemcpyDeviceToDevice, st);
  }
  return launch_status();
}

template <typename T>
static int dense_bwd(const T* A, const T* S, const T* X, const T* gXpool, const T* gApool, const float* gl, int B,
                     int N, int K, int F, uint32_t flags, int loss_kind, float eps, float link_div, float ent_div,
                     T* dS_out, T* dX_out, T* dA_out, Workspace& saved, Workspace& ws, cudaStream_t st) {
  DensePlan<T> pl(saved, B, N, K);
  if (!pl.ok) return TGPB200_ERR_WORKSPACE;
  int64_t NK = (int64_t)N * K, NN = (int64_t)N * N, NF = (int64_t)N * F, KK = (int64_t)K * K, KF = (int64_t)K * F;
  float* Graw = ws.take<float>((size_t)B * KK);
  float* P = ws.take<float>((size_t)B * KK);
  float* coef = ws.take<float>((size_t)B * 4);
  T* W = ws.take<T>((size_t)B * NK);
  T* U = ws.take<T>((size_t)B * NK);
  T* Gt = ws.take<T>((size_t)B * KK);  // Graw / P in the operand dtype (bf16 runs)
  T* Pt = ws.take<T>((size_t)B * KK);
  if (!ws.ok) return TGPB200_ERR_WORKSPACE;
  int rc;
  if (B == 0) return TGPB200_OK;
  const bool have_a = A != nullptr;
  const bool f32 = std::is_same<T, float>::value;
  if (have_a) {
    const StripPlan sp = strip_plan(K, 4, 3);
    const bool v4 = K % 4 == 0 && aligned16(gApool);
    auto kern = sp.stage ? (v4 ? k_graph_bwd<T, 4, true> : k_graph_bwd<T, 1, true>)
                         : (v4 ? k_graph_bwd<T, 4, false> : k_graph_bwd<T, 1, false>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
    launch_cluster("k_graph_bwd", kern, (unsigned)B * sp.R, sp.threads, sp.R, sp.smem, st, pl.Araw,
                   loss_kind != 0 ? pl.M : (float*)nullptr, gApool, pl.dvec, pl.stats, pl.argmax, gl, pl.losses, B, K,
                   flags, loss_kind, eps, link_div, ent_div, Graw, loss_kind != 0 ? P : (float*)nullptr, coef,
                   sp.rows_per, sp.rowp_floats, f32 ? (T*)nullptr : Gt, (f32 || loss_kind == 0) ? (T*)nullptr : Pt);
    if (f32) {
      Gt = reinterpret_cast<T*>(Graw);
      Pt = reinterpret_cast<T*>(P);
    }
  }
  const bool have_x = X && gXpool;
  bool fused_bwd = false;
  if constexpr (std::is_same<T, float>::value) {
    if (have_a && have_x && dX_out && dS_out) {
      // one launch: W = A S in tensor memory -> dS (all four products + element-wise terms) and dX
      const bool ew = loss_kind != 0;
      rc = tc::dense_bwd_fused(A, S, X, pl.Tt, gXpool, Graw, loss_kind != 0 ? P : (float*)nullptr, B, N, K, F,
                               ew ? S : (const float*)nullptr, pl.d, coef, eps, dS_out, dX_out, st);
      if (rc == TGPB200_OK) fused_bwd = true;
      else if (rc != TGPB200_ERR_UNSUPPORTED) return rc;
    }
  }
  if (!fused_bwd) {
  if (have_x && dX_out) {  // dX = S Gx  [N, F]
    rc = mm1<T, T>(B, N, F, K, Mat{S, NK, K, 0}, Mat{gXpool, KF, F, 1}, dX_out, NF, F, 1, st, "k_tc_gemm:dX=SGx");
    if (rc) return rc;
  }
  if (have_a) {  // W = A S  [N, K]
    rc = mm1<T, T>(B, N, K, N, Mat{A, NN, N, 0}, Mat{S, NK, K, 1}, W, NK, K, 1, st, "k_tc_gemm:W=AS");
    if (rc) return rc;
  }
  // dS = X Gx^T + W Graw^T + T^T Graw + S P   (one accumulation chain in TMEM; the element-wise loss terms
  // are added in the same epilogue when the tensor-core engine runs the product)
  bool ew_done = false;
  {
    int kd[4];
    Mat a[4], b[4];
    int n = 0;
    if (have_x) { kd[n] = F; a[n] = Mat{X, NF, F, 0}; b[n] = Mat{gXpool, KF, F, 0}; ++n; }
    if (have_a) {
      kd[n] = K; a[n] = Mat{W, NK, K, 0}; b[n] = Mat{Gt, KK, K, 0}; ++n;
      kd[n] = K; a[n] = Mat{pl.Tt, NK, N, 1}; b[n] = Mat{Gt, KK, K, 1}; ++n;  // T^T Graw
      if (loss_kind != 0) { kd[n] = K; a[n] = Mat{S, NK, K, 0}; b[n] = Mat{Pt, KK, K, 1}; ++n; }
    }
    if (n > 0) {
      EwHook hook;
      if (have_a && loss_kind != 0) {
        hook.S = S, hook.d = pl.d, hook.coef = coef, hook.eps = eps, hook.applied = &ew_done;
      }
      rc = mm<T, T>(B, N, K, n, kd, a, b, dS_out, NK, K, 1, st, hook, "k_tc_gemm:dS");
      if (rc) return rc;
    } else {
      cudaMemsetAsync(dS_out, 0, (size_t)B * NK * sizeof(T), st);
    }
  }
  if (have_a && loss_kind != 0 && !ew_done)
    launch("k_ds_elementwise", k_ds_elementwise<T>, (unsigned)ceil_div((int64_t)B * NK, 256), 256, 0, st, S, pl.d, coef,
           (int64_t)B * NK, N, K, eps, dS_out);
  }  // !fused_bwd
  if (have_a && dA_out) {  // dA = (S Graw) S^T + element-wise terms
    rc = mm1<T, T>(B, N, K, K, Mat{S, NK, K, 0}, Mat{Gt, KK, K, 1}, U, NK, K, 1, st, "k_tc_gemm:U=SG");
    if (rc) return rc;
    rc = mm1<T, T>(B, N, N, K, Mat{U, NK, K, 0}, Mat{S, NK, K, 0}, dA_out, NN, N, 1, st, "k_tc_gemm:dA=USt");
    if (rc) return rc;
    if (loss_kind != 0) {
      constexpr int V = 16 / sizeof(T);
      if (N % V == 0 && aligned16(A) && aligned16(dA_out))
        launch("k_da_elementwise", k_da_elementwise_vec<T>, (unsigned)ceil_div((int64_t)B * NN / V, 256), 256, 0, st, A,
               pl.ss, coef, (int64_t)B * NN / V, N, dA_out);
      else
        launch("k_da_elementwise", k_da_elementwise<T>, (unsigned)ceil_div((int64_t)B * NN, 256), 256, 0, st, A, pl.ss,
               coef, (int64_t)B * NN, N, dA_out);
    }
  }
  return launch_status();
}

}  // namespace tgp

using namespace tgp;

extern "C" {

size_t tgpb200_dense_pool_saved_bytes(int64_t B, int64_t N, int64_t K) { return dense_saved_bytes(B, N, K); }

size_t tgpb200_dense_pool_bwd_workspace_bytes(int64_t B, int64_t N, int64_t K, int need_grad_adj) {
  size_t f = sizeof(float);
  (void)need_grad_adj;
  return 4 * align_up((size_t)B * K * K * f) + align_up((size_t)B * 4 * f) + 2 * align_up((size_t)B * N * K * f) + 8192;
}

// Batched product with the shape-general fallback (tensor-core engine when the operand layout satisfies the TMA
// constraints, FP32-pipe tiles otherwise): out[b] (M x N, row-major) = A[b] B[b].
int tgpb200_bmm(const void* a, const void* b, void* out, int64_t batch, int64_t M, int64_t N, int64_t Kd,
                int64_t a_batch_stride, int64_t a_row_stride, int a_mn_major, int64_t b_batch_stride,
                int64_t b_row_stride, int b_mn_major, int dtype, tgpb200_stream_t stream) {
  if (batch < 0 || M < 0 || N < 0 || Kd < 0 || !out) return TGPB200_ERR_INVALID;
  if (batch == 0 || M == 0 || N == 0) return TGPB200_OK;
  if (batch >= INT32_MAX || M >= INT32_MAX || N >= INT32_MAX || Kd >= INT32_MAX) return TGPB200_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (Kd == 0) {
    cudaMemsetAsync(out, 0, (size_t)batch * M * N * (dtype == TGPB200_F32 ? 4 : 2), st);
    return launch_status();
  }
  if (!a || !b) return TGPB200_ERR_INVALID;
  const Mat A{a, a_batch_stride, a_row_stride, a_mn_major}, Bm{b, b_batch_stride, b_row_stride, b_mn_major};
  if (dtype == TGPB200_F32)
    return mm1<float, float>((int)batch, (int)M, (int)N, (int)Kd, A, Bm, (float*)out, M * N, N, 1, st, "k_tc_gemm:bmm");
  if (dtype == TGPB200_BF16)
    return mm1<__nv_bfloat16, __nv_bfloat16>((int)batch, (int)M, (int)N, (int)Kd, A, Bm, (__nv_bfloat16*)out, M * N, N,
                                             1, st, "k_tc_gemm:bmm");
  return TGPB200_ERR_UNSUPPORTED;
}

int tgpb200_dense_pool_fwd(const void* adj, const void* s, const void* x, int64_t B, int64_t N, int64_t K, int64_t F,
                           int dtype, uint32_t flags, int loss_kind, float eps, float link_div, float ent_div,
                           void* x_pool, void* adj_pool, float* losses, void* saved, size_t saved_bytes,
                           tgpb200_stream_t stream) {
  if (B < 0 || N < 0 || K < 0 || F < 0 || !s || !saved) return TGPB200_ERR_INVALID;
  if (B >= INT32_MAX || N >= 65536 || K >= 32768 || F >= INT32_MAX) return TGPB200_ERR_UNSUPPORTED;
  if (loss_kind < 0 || loss_kind > 2 || (loss_kind != 0 && !adj)) return TGPB200_ERR_INVALID;
  if (adj && !adj_pool) return TGPB200_ERR_INVALID;
  Workspace sv(saved, saved_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TGPB200_F32)
    return dense_fwd<float>((const float*)adj, (const float*)s, (const float*)x, (int)B, (int)N, (int)K, (int)F, flags,
                            loss_kind, eps, link_div, ent_div, (float*)x_pool, (float*)adj_pool, losses, sv, st);
  if (dtype == TGPB200_BF16)
    return dense_fwd<__nv_bfloat16>((const __nv_bfloat16*)adj, (const __nv_bfloat16*)s, (const __nv_bfloat16*)x, (int)B,
                                    (int)N, (int)K, (int)F, flags, loss_kind, eps, link_div, ent_div,
                                    (__nv_bfloat16*)x_pool, (__nv_bfloat16*)adj_pool, losses, sv, st);
  return TGPB200_ERR_UNSUPPORTED;
}

int tgpb200_dense_pool_bwd(const void* adj, const void* s, const void* x, const void* grad_x_pool,
                           const void* grad_adj_pool, const float* grad_losses, int64_t B, int64_t N, int64_t K,
                           int64_t F, int dtype, uint32_t flags, int loss_kind, float eps, float link_div,
                           float ent_div, void* grad_s, void* grad_x, void* grad_adj, void* saved, size_t saved_bytes,
                           void* workspace, size_t workspace_bytes, tgpb200_stream_t stream) {
  if (B < 0 || N < 0 || K < 0 || F < 0 || !s || !saved) return TGPB200_ERR_INVALID;
  if (loss_kind < 0 || loss_kind > 2) return TGPB200_ERR_INVALID;
  Workspace sv(saved, saved_bytes), ws(workspace, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TGPB200_F32)
    return dense_bwd<float>((const float*)adj, (const float*)s, (const float*)x, (const float*)grad_x_pool,
                            (const float*)grad_adj_pool, grad_losses, (int)B, (int)N, (int)K, (int)F, flags, loss_kind,
                            eps, link_div, ent_div, (float*)grad_s, (float*)grad_x, (float*)grad_adj, sv, ws, st);
  if (dtype == TGPB200_BF16)
    return dense_bwd<__nv_bfloat16>((const __nv_bfloat16*)adj, (const __nv_bfloat16*)s, (const __nv_bfloat16*)x,
                                    (const __nv_bfloat16*)grad_x_pool, (const __nv_bfloat16*)grad_adj_pool,
                                    grad_losses, (int)B, (int)N, (int)K, (int)F, flags, loss_kind, eps, link_div,
                                    ent_div, (__nv_bfloat16*)grad_s, (__nv_bfloat16*)grad_x, (__nv_bfloat16*)grad_adj,
                                    sv, ws, st);
  return TGPB200_ERR_UNSUPPORTED;
}

}  // extern "C"
