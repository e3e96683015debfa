// K3: bidirectional multi-head attention for very short sequences (S = kept tokens <= 64, head_dim 128).
//
// Reference call site: nn.MultiheadAttention inside nn.TransformerEncoderLayer (mtm_model.py:379-409):
// softmax(q k^T / sqrt(128)) v, no mask, eval mode.  One CTA per (batch row, head): Q, K, V of that head
// (S x 128 each) live in shared memory as fp32, the S x S score matrix never leaves the SM.
// Activations are token-major: row = token * B + b, so a (b, head) slice is S rows of 128 contiguous values.
#include "kernels.cuh"

namespace m3pc {
namespace {

constexpr int HD = 128;       // head dim
constexpr int QS = HD + 1;    // padded row stride (floats) for conflict-free row-parallel reads
constexpr int ATT_THREADS = 128;

template <typename AT>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AT* __restrict__ q_base, int q_ld, const AT* __restrict__ k_base,
                                                                const AT* __restrict__ v_base, int kv_ld, AT* __restrict__ out, int out_ld,
                                                                int B, int n_q, int S, int n_head) {
  extern __shared__ float sm[];
  float* Vs = sm;                 // S x HD (first: keeps its float4 stores 16-byte aligned)
  float* Qs = Vs + S * HD;        // n_q x QS
  float* Ks = Qs + n_q * QS;      // S x QS
  float* Ps = Ks + S * QS;        // n_q x (S + 1)
  const int b = blockIdx.x / n_head, h = blockIdx.x % n_head;
  const int tid = threadIdx.x;

  // stage Q, K, V (4 elements per access)
  for (int idx = tid; idx < n_q * (HD / 4); idx += ATT_THREADS) {
    const int i = idx / (HD / 4), c = (idx % (HD / 4)) * 4;
    const float4 v = ld4(q_base + (static_cast<size_t>(i) * B + b) * q_ld + h * HD + c);
    float* d = Qs + i * QS + c;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  for (int idx = tid; idx < S * (HD / 4); idx += ATT_THREADS) {
    const int j = idx / (HD / 4), c = (idx % (HD / 4)) * 4;
    const size_t row = (static_cast<size_t>(j) * B + b) * kv_ld + h * HD + c;
    const float4 kv = ld4(k_base + row);
    float* d = Ks + j * QS + c;
    d[0] = kv.x; d[1] = kv.y; d[2] = kv.z; d[3] = kv.w;
    *reinterpret_cast<float4*>(Vs + j * HD + c) = ld4(v_base + row);
  }
  __syncthreads();

  // scores: consecutive threads take consecutive keys j of the same query i (Ks rows conflict-free, Qs broadcast)
  const float scale = 0.08838834764831845f;  // 1/sqrt(128)
  for (int pidx = tid; pidx < n_q * S; pidx += ATT_THREADS) {
    const int i = pidx / S, j = pidx - i * S;
    const float* qi = Qs + i * QS;
    const float* kj = Ks + j * QS;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; d += 4) {
      a0 = fmaf(qi[d], kj[d], a0);
      a1 = fmaf(qi[d + 1], kj[d + 1], a1);
      a2 = fmaf(qi[d + 2], kj[d + 2], a2);
      a3 = fmaf(qi[d + 3], kj[d + 3], a3);
    }
    Ps[i * (S + 1) + j] = ((a0 + a1) + (a2 + a3)) * scale;
  }
  __syncthreads();

  // softmax per query row: one warp per row
  const int lane = tid & 31, warp = tid >> 5;
  for (int i = warp; i < n_q; i += ATT_THREADS / 32) {
    float* pr = Ps + i * (S + 1);
    float m = -INFINITY;
    for (int j = lane; j < S; j += 32) m = fmaxf(m, pr[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float e = expf(pr[j] - m);
      pr[j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < S; j += 32) pr[j] *= inv;
  }
  __syncthreads();

  // out[i, d] = sum_j P[i, j] V[j, d]; thread = d
  for (int i = 0; i < n_q; ++i) {
    const float* pr = Ps + i * (S + 1);
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc = fmaf(pr[j], Vs[j * HD + tid], acc);
    Act<AT>::st(out + (static_cast<size_t>(i) * B + b) * out_ld + h * HD + tid, acc);
  }
}

template <typename AT>
int launch_t(const AT* q, int q_ld, const AT* k, const AT* v, int kv_ld, AT* out, int out_ld, int B, int n_q, int S, int n_head,
             cudaStream_t st) {
  const size_t smem = (static_cast<size_t>(n_q) * QS + static_cast<size_t>(S) * QS + static_cast<size_t>(S) * HD +
                       static_cast<size_t>(n_q) * (S + 1)) * sizeof(float);
  M3PC_REQUIRE(smem <= 200 * 1024, "attention: sequence too long for the shared-memory resident kernel");
  static size_t configured = 0;
  if (smem > configured) {
    M3PC_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  attention_kernel<AT><<<B * n_head, ATT_THREADS, smem, st>>>(q, q_ld, k, v, kv_ld, out, out_ld, B, n_q, S, n_head);
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}

}  // namespace

int launch_attention(const void* qkv, void* out, int B, int S, int n_head, bool bf16, cudaStream_t st) {
  M3PC_REQUIRE(B > 0 && S > 0 && S <= 128 && n_head > 0, "attention: bad shape");
  const int D = n_head * HD;
  if (bf16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(qkv);
    return launch_t<__nv_bfloat16>(p, 3 * D, p + D, p + 2 * D, 3 * D, reinterpret_cast<__nv_bfloat16*>(out), D, B, S, S, n_head, st);
  }
  const float* p = reinterpret_cast<const float*>(qkv);
  return launch_t<float>(p, 3 * D, p + D, p + 2 * D, 3 * D, reinterpret_cast<float*>(out), D, B, S, S, n_head, st);
}

}  // namespace m3pc
