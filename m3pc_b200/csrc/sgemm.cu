// fp32 GEMM on CUDA cores: C[M,N] = epilogue(A[M,K] * W[N,K]^T).
//
// Used for the reference-grade precision mode (M3PC_PREC_FP32): every contraction of the MTM and of the TwinQ critic
// (finetune_omtm/model.py:146-171) in fp32, the arithmetic type of the reference.  In bf16 mode the critic's two hidden
// layers run on the tcgen05 GEMM instead (bf16 operands, fp32 accumulate; engine.cu:critic) when its width is a multiple of
// 128, and only fall back to this kernel otherwise; the final 256 -> 1 layer and min(q1, q2) are fp32 in both modes.
// Plain register-blocked tiling (64x64x16 per 256-thread CTA, 4x4 outputs per thread); arbitrary M, N, K.
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

}  // namespace

int gemm_fp32(const float* A, const float* W, float* C, int M, int N, int K, const GemmEpilogue& epi, cudaStream_t st) {
  M3PC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_fp32: empty problem");
  M3PC_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemm_fp32: operands must be 16-byte aligned");
  dim3 grid(ceil_div(N, SB_N), ceil_div(M, SB_M));
  M3PC_CHECK_CUDA(launch_k(sgemm_kernel, dim3(grid), dim3(256), 0, st, A, W, C, M, N, K, epi.bias, epi.table, epi.rows_per_group > 0 ? epi.rows_per_group : 1, epi.flags));
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}

}  // namespace m3pc
