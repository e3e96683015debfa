// Skinny bf16 GEMM for M <= 32 rows: C[M,N] = epilogue(A[M,K] * W[N,K]^T), fp32 accumulation on CUDA cores.
//
// The B = 1 passes of the planners (pass 1 of Learner.rtg_guiding / critic_lambda_guiding, learner.py:278-284; the
// zero-shot planners) have 8..17 token rows: a 128-row tcgen05 tile would be > 85 % padding and the kernel is pure
// latency.  Here every warp owns output columns: it streams the column's weight row once (coalesced 16-byte loads)
// and dots it with all M activation rows staged in shared memory, so all SMs pull weights from L2 concurrently.
// Same fused epilogues as the tensor-core kernel (bias, per-token table, GELU, ReLU, residual, bf16 / fp32 out).
#include "common.cuh"

namespace m3pc {
namespace {

constexpr int SK_THREADS = 256;      // 8 warps
constexpr int SK_COLS_PER_WARP = 2;  // columns per warp -> 16 per CTA
constexpr int SK_MAX_M = 32;

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(SK_THREADS) gemm_skinny_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W,
                                                                 void* __restrict__ C, int M, int N, int K, const float* __restrict__ bias,
                                                                 const float* __restrict__ table, int rows_per_group, int flags) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(sk_smem);  // M x K
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // stage A (M*K bf16, 16-byte chunks)
  const int chunks = M * K / 8;
  for (int i = threadIdx.x; i < chunks; i += SK_THREADS)
    reinterpret_cast<uint4*>(As)[i] = __ldg(reinterpret_cast<const uint4*>(A) + i);
  __syncthreads();
  const bool do_gelu = flags & EPI_GELU, do_relu = flags & EPI_RELU, do_res = flags & EPI_RESIDUAL;
  const bool out_f32 = do_res || (flags & EPI_OUT_F32);
#pragma unroll 1
  for (int cw = 0; cw < SK_COLS_PER_WARP; ++cw) {
    const int n = (blockIdx.x * (SK_THREADS / 32) + warp) * SK_COLS_PER_WARP + cw;
    if (n >= N) break;
    float acc[SK_MAX_M];
#pragma unroll
    for (int m = 0; m < SK_MAX_M; ++m) acc[m] = 0.f;
    const uint4* wrow = reinterpret_cast<const uint4*>(W + static_cast<size_t>(n) * K);
    for (int c = lane; c < K / 8; c += 32) {
      float wf[8];
      bf16x8_to_float(__ldg(wrow + c), wf);
#pragma unroll
      for (int m = 0; m < SK_MAX_M; ++m) {
        if (m < M) {
          float af[8];
          bf16x8_to_float(*reinterpret_cast<const uint4*>(As + static_cast<size_t>(m) * K + c * 8), af);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[m] = fmaf(af[i], wf[i], acc[m]);
        }
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < SK_MAX_M; ++m) {
      if (m < M) {
        const float s = warp_sum(acc[m]);
        if (lane == m) mine = s;
      }
    }
    if (lane < M) {
      float v = mine;
      if (bias != nullptr) v += __ldg(bias + n);
      if (table != nullptr) v += __ldg(table + static_cast<size_t>(lane / rows_per_group) * N + n);
      if (do_gelu) v = gelu_erf_fast(v);
      if (do_relu) v = fmaxf(v, 0.f);
      if (out_f32) {
        float* cp = reinterpret_cast<float*>(C) + static_cast<size_t>(lane) * N + n;
        if (do_res) v += *cp;
        *cp = v;
      } else {
        reinterpret_cast<__nv_bfloat16*>(C)[static_cast<size_t>(lane) * N + n] = __float2bfloat16_rn(v);
      }
    }
  }
}

}  // namespace

int gemm_bf16_skinny(const __nv_bfloat16* A, const __nv_bfloat16* W, void* C, int M, int N, int K, const GemmEpilogue& epi, cudaStream_t st) {
  M3PC_REQUIRE(M >= 1 && M <= SK_MAX_M && K % 8 == 0, "gemm_skinny: needs M <= 32 and K % 8 == 0");
  const size_t smem = static_cast<size_t>(M) * K * sizeof(__nv_bfloat16);
  M3PC_REQUIRE(smem <= 160 * 1024, "gemm_skinny: A panel does not fit shared memory");
  static size_t configured = 0;
  if (smem > configured) {
    M3PC_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  const int cols_per_cta = (SK_THREADS / 32) * SK_COLS_PER_WARP;
  gemm_skinny_kernel<<<ceil_div(N, cols_per_cta), SK_THREADS, smem, st>>>(A, W, C, M, N, K, epi.bias, epi.table,
                                                                          epi.rows_per_group > 0 ? epi.rows_per_group : 1, epi.flags);
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}

}  // namespace m3pc
