// fp32 GEMM on CUDA cores: C[M,N] = epilogue(A[M,K] * W[N,K]^T).
//
// Used for the reference-grade precision mode (M3PC_PREC_FP32): every contraction of the MTM and of the TwinQ critic
// (finetune_omtm/model.py:146-171) in fp32, the arithmetic type of the reference.  In bf16 mode the critic's two hidden
// layers run on the tcgen05 GEMM instead (bf16 operands, fp32 accumulate; engine.cu:critic) when its width is a multiple of
// 128, and only fall back to this kernel otherwise; the final 256 -> 1 layer and min(q1, q2) are fp32 in both modes.
// Two register-blocked tilings, arbitrary M, N, K:
//   * 128x128x16 per 256-thread CTA, 8x8 outputs per thread (as 2x2 blocks of 4x4, so every shared-memory read is a
//     conflict-free 16-byte load), the next K slab prefetched into registers while the current one is multiplied and
//     double-buffered in shared memory (one barrier per slab): 4 LDS.128 per 64 FFMA;
//   * 64x64x16, 4x4 outputs per thread, for problems smaller than one big tile in either dimension.
#include "common.cuh"

namespace m3pc {
namespace {

constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                    float* __restrict__ C, int M, int N, int K, const float* __restrict__ bias,
                                                    const float* __restrict__ table, int rows_per_group, int flags) {
  PDL_PROLOGUE();
  __shared__ float As[SB_K][SB_M + 4];
  __shared__ float Ws[SB_K][SB_N + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * SB_M, n0 = blockIdx.x * SB_N;
  float acc[4][4] = {};
  // loader mapping: 256 threads, 64 rows x 16 k: thread -> (row = tid / 4, k4 = (tid % 4) * 4)
  const int lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;
  const bool k_vec = (K & 3) == 0;
  for (int k0 = 0; k0 < K; k0 += SB_K) {
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int r = m0 + lrow, kk = k0 + lk;
      if (r < M) {
        const float* p = A + static_cast<size_t>(r) * K + kk;
        if (k_vec && kk + 3 < K) {
          float4 t = *reinterpret_cast<const float4*>(p);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (kk + i < K) v[i] = p[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[lk + i][lrow] = v[i];
    }
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int r = n0 + lrow, kk = k0 + lk;
      if (r < N) {
        const float* p = W + static_cast<size_t>(r) * K + kk;
        if (k_vec && kk + 3 < K) {
          float4 t = __ldg(reinterpret_cast<const float4*>(p));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (kk + i < K) v[i] = __ldg(p + i);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Ws[lk + i][lrow] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SB_K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
    const float* trow = table ? table + static_cast<size_t>(r / rows_per_group) * N : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + c);
      if (trow) v += __ldg(trow + c);
      if (flags & EPI_GELU) v = gelu_erf(v);
      if (flags & EPI_RELU) v = fmaxf(v, 0.f);
      float* cp = C + static_cast<size_t>(r) * N + c;
      if (flags & EPI_RESIDUAL) v += *cp;
      *cp = v;
    }
  }
}

constexpr int LB_M = 128, LB_N = 128, LB_K = 16, LB_LD = LB_M + 4;

// one thread's share of a 128 x 16 operand slab: rows (tid / 4) and (tid / 4) + 64, k = (tid % 4) * 4 .. + 3
__device__ __forceinline__ void slab_load(const float* __restrict__ P, int rows, int K, int r0, int k0, bool k_vec, float4 (&v)[2]) {
  const int lrow = threadIdx.x >> 2, kk = k0 + (threadIdx.x & 3) * 4;
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    const int r = r0 + lrow + 64 * hlf;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const float* p = P + static_cast<size_t>(r) * K + kk;
      if (k_vec && kk + 3 < K) {
        t = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        if (kk < K) t.x = __ldg(p);
        if (kk + 1 < K) t.y = __ldg(p + 1);
        if (kk + 2 < K) t.z = __ldg(p + 2);
        if (kk + 3 < K) t.w = __ldg(p + 3);
      }
    }
    v[hlf] = t;
  }
}
__device__ __forceinline__ void slab_store(float (*S)[LB_LD], const float4 (&v)[2]) {
  const int lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    S[lk + 0][lrow + 64 * hlf] = v[hlf].x;
    S[lk + 1][lrow + 64 * hlf] = v[hlf].y;
    S[lk + 2][lrow + 64 * hlf] = v[hlf].z;
    S[lk + 3][lrow + 64 * hlf] = v[hlf].w;
  }
}

__global__ void __launch_bounds__(256, 2) sgemm128_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, int M,
                                                          int N, int K, const float* __restrict__ bias, const float* __restrict__ table,
                                                          int rows_per_group, int flags) {
  PDL_PROLOGUE();
  __shared__ __align__(16) float As[2][LB_K][LB_LD];
  __shared__ __align__(16) float Ws[2][LB_K][LB_LD];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * LB_M, n0 = blockIdx.x * LB_N;
  const bool k_vec = (K & 3) == 0;
  float acc[8][8] = {};
  float4 pa[2], pw[2];
  slab_load(A, M, K, m0, 0, k_vec, pa);
  slab_load(W, N, K, n0, 0, k_vec, pw);
  slab_store(As[0], pa);
  slab_store(Ws[0], pw);
  __syncthreads();
  int cur = 0;
  for (int k0 = 0; k0 < K; k0 += LB_K) {
    const bool more = k0 + LB_K < K;
    if (more) {  // the next slab travels while this one is multiplied
      slab_load(A, M, K, m0, k0 + LB_K, k_vec, pa);
      slab_load(W, N, K, n0, k0 + LB_K, k_vec, pw);
    }
#pragma unroll
    for (int k = 0; k < LB_K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]), a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[cur][k][tx * 4]), w1 = *reinterpret_cast<const float4*>(&Ws[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (more) {
      slab_store(As[cur ^ 1], pa);  // last read two iterations ago, behind the barrier below
      slab_store(Ws[cur ^ 1], pw);
      __syncthreads();
      cur ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (r >= M) continue;
    const float* trow = table ? table + static_cast<size_t>(r / rows_per_group) * N : nullptr;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c = n0 + jh * 64 + tx * 4;
      float v[4] = {acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      float* cp = C + static_cast<size_t>(r) * N + c;
      const bool vec = c + 3 < N && (N & 3) == 0;
      float res[4] = {0.f, 0.f, 0.f, 0.f};
      if (flags & EPI_RESIDUAL) {
        if (vec) {
          const float4 t = *reinterpret_cast<const float4*>(cp);
          res[0] = t.x; res[1] = t.y; res[2] = t.z; res[3] = t.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < N) res[j] = cp[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c + j >= N) continue;
        float x = v[j];
        if (bias) x += __ldg(bias + c + j);
        if (trow) x += __ldg(trow + c + j);
        if (flags & EPI_GELU) x = gelu_erf(x);
        if (flags & EPI_RELU) x = fmaxf(x, 0.f);
        v[j] = x + res[j];
      }
      if (vec) {
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < N) cp[j] = v[j];
      }
    }
  }
}

}  // namespace

int gemm_fp32(const float* A, const float* W, float* C, int M, int N, int K, const GemmEpilogue& epi, cudaStream_t st) {
  M3PC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_fp32: empty problem");
  M3PC_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemm_fp32: operands must be 16-byte aligned");
  M3PC_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm_fp32: the output must be 16-byte aligned");
  if (M >= LB_M && N >= LB_N) {
    dim3 big(ceil_div(N, LB_N), ceil_div(M, LB_M));
    M3PC_CHECK_CUDA(launch_k(sgemm128_kernel, big, dim3(256), 0, st, A, W, C, M, N, K, epi.bias, epi.table, epi.rows_per_group > 0 ? epi.rows_per_group : 1, epi.flags));
    M3PC_CHECK_LAUNCH();
    return M3PC_OK;
  }
  dim3 grid(ceil_div(N, SB_N), ceil_div(M, SB_M));
  M3PC_CHECK_CUDA(launch_k(sgemm_kernel, dim3(grid), dim3(256), 0, st, A, W, C, M, N, K, epi.bias, epi.table, epi.rows_per_group > 0 ? epi.rows_per_group : 1, epi.flags));
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}

}  // namespace m3pc
