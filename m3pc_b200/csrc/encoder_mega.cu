// Encoder "megakernel": all encoder layers of the batched MTM forward (omtm.forward_encoder, mtm_model.py:619-644; pre-LN
// nn.TransformerEncoderLayer blocks, mtm_model.py:379-391) in ONE persistent launch, for the big-batch passes of the planners
// (pass 2 of Learner.rtg_guiding / critic_lambda_guiding, finetune_omtm/learner.py:292-293).
//
// Observation: every op of a block is local to a batch row's own tokens (GEMMs and LayerNorms are row-wise, attention mixes
// only the <= 16 tokens of one candidate).  With a candidate-major row layout -- a 128-row tile holds cpt = 128 / S whole
// candidates -- a CTA can therefore take its tile through
//     QKV GEMM -> attention -> out-proj GEMM (+residual) -> LN2 -> lin1 GEMM (+GELU) -> lin2 GEMM (+residual) -> next LN
// for every layer without ever synchronising with another tile: no kernel boundaries, no grid barriers, no wave
// quantisation per GEMM, no separate LayerNorm / attention launches.  What the multi-kernel path spent per layer on 7
// launches (profiles/r1c_*: 4 GEMMs at 45 % tensor-pipe activity, 2 LayerNorms, 1 attention) becomes one pipeline.
//
// Structure (same roles as gemm_bf16_2sm_kernel, CTA pairs with cta_group::2 MMAs: the pair shares every weight k-block,
// each CTA loads half of it):
//   warp 0       TMA producer: walks the op list; before the first A load of a GEMM it waits on the `ready` mbarrier that
//                the epilogue warps arrive on when the op producing that operand is complete.
//   warp 1       MMA issuer (leader CTA): tcgen05.mma over the op list's units, two TMEM accumulators (double buffered
//                across units AND across ops).
//   warps 2..17  epilogue + "compute ops": drain accumulators (bias / GELU / bf16 TMA store, or fp32 TMA reduce-add into
//                the residual stream), then run the row-local non-GEMM ops of the block on their own rows: attention
//                (mma.sync, one (candidate, head) pair per warp, softmax in the accumulator registers) and LayerNorm.
// Intermediates (QKV, ATT, HID, Y, X) live in global memory but are private to the CTA and L2-resident; visibility between
// the generic and the async (TMA) proxy is handled with cp.async.bulk.wait_group / fence.proxy.async + mbarriers.
// Numerics are those of the multi-kernel bf16 path (same rounding points), so the two paths agree to fp32 round-off.
#include <algorithm>

#include "kernels.cuh"
#include "tcgen05.cuh"

namespace m3pc {
namespace {

constexpr int MG_EPI_WARPS = 16;
constexpr int MG_THREADS = 32 * (2 + MG_EPI_WARPS);
constexpr int MG_STAGES = 4;
constexpr int MG_MAX_LAYERS = 4;
constexpr int MG_BN = 256;
constexpr float kLnEps = 1e-5f;

struct MegaLayer {
  CUtensorMap w_in, w_out, w_l1, w_l2;  // weights (out, in) row-major = K-major, box 128 rows x 64 k
  const float *in_b, *out_b, *l1_b, *l2_b;
  const float *n2_w, *n2_b;      // norm2 of this layer
  const float *post_w, *post_b;  // LayerNorm applied to x after the layer: the next layer's norm1, or the final encoder norm
};
struct MegaParams {
  CUtensorMap a_y, a_att, a_hid;  // A operands: Y (LN output), ATT, HID -- box 128 rows x 64 k
  CUtensorMap c_qkv, c_hid, c_x;  // outputs: QKV / HID bf16 stores, X fp32 reduce-add -- box 32 rows x 64 bytes
  MegaLayer layer[MG_MAX_LAYERS];
  int n_layers, S, cpt, B, n_pair_tiles, D;
  float* X;              // (R, D)  fp32 residual stream, candidate-major tiles
  __nv_bfloat16* Y;      // (R, D)  LayerNorm output = next GEMM's A operand
  __nv_bfloat16* QKV;    // (R, 3D)
  __nv_bfloat16* ATT;    // (R, D)
  __nv_bfloat16* ENC;    // (S * B, D) token-major output of the final encoder norm (what the decoder kernels consume)
};

struct Smem {
  static constexpr int kABlk = BM * BK * 2;
  static constexpr int kBBlk = 128 * BK * 2;
  static constexpr int kStageBytes = kABlk + kBBlk;
  static constexpr int kScratchPerWarp = 4096;  // staging for TMA stores (2 x 2 KB) / V tile + O tile of the attention op
  static constexpr int kScratchOffset = MG_STAGES * kStageBytes;
  static constexpr int kBiasOffset = kScratchOffset + MG_EPI_WARPS * kScratchPerWarp;
  static constexpr int kBarOffset = kBiasOffset + MG_EPI_WARPS * 64 * 4;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MG_EPI_WARPS * 32) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// byte offset of 16-byte chunk c (0..15) of row r in a [16][256 B] tile, XOR-swizzled (conflict-free ldmatrix / row stores)
__device__ __forceinline__ uint32_t swz(int r, int c) { return static_cast<uint32_t>(r * 256 + ((c ^ (r & 7)) << 4)); }

struct EpiCtx {
  uint8_t* scratch;  // this warp's 4 KB
  float* sbias;      // this warp's 64 floats
  uint32_t nstore;
  uint32_t tmem_base, acc_empty_leader;
  int lane, quarter, cgrp;
};

// Drain this warp's 32 rows x 64 columns of accumulator `acc`: bias (+GELU) -> bf16 TMA store, or fp32 TMA reduce-add.
// `row0`: first row of the warp's 32, `colw`: first of its 64 columns in the output matrix.
template <bool F32, bool GELU>
__device__ __forceinline__ void drain_unit(EpiCtx& c, const CUtensorMap* tmap_c, const float* __restrict__ bias, int row0, int colw, int acc) {
  const int lane = c.lane;
  {
    const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + colw + 2 * lane));
    *reinterpret_cast<float2*>(c.sbias + 2 * lane) = b2;
    __syncwarp();
  }
  constexpr int CW = F32 ? 16 : 32;
  const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
  const uint32_t tacc = c.tmem_base + (static_cast<uint32_t>(c.quarter * 32) << 16) + static_cast<uint32_t>(acc * MG_BN + c.cgrp * 64);
#pragma unroll 1
  for (int cc = 0; cc < 64; cc += CW) {
    uint32_t r[32];
    if (F32) tmem_ld16(tacc + static_cast<uint32_t>(cc), r); else tmem_ld32(tacc + static_cast<uint32_t>(cc), r);
    uint8_t* buf = c.scratch + (c.nstore & 1) * 2048;
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    tmem_ld_wait();
    if (cc + CW == 64) {
      tc_fence_before();
      if (lane == 0) mbar_arrive_remote(c.acc_empty_leader + 8u * static_cast<uint32_t>(acc));
    }
    float v[CW];
#pragma unroll
    for (int j = 0; j < CW / 4; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(c.sbias + cc + 4 * j);
      v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b4.x;
      v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
      v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
      v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
    }
    if (GELU) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = gelu_erf_tanh(v[j]);
    }
    if (F32) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(buf + lane * 64 + ((static_cast<uint32_t>(j) ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack2(v[8 * j + 0], v[8 * j + 1]); o.y = pack2(v[8 * j + 2], v[8 * j + 3]);
        o.z = pack2(v[8 * j + 4], v[8 * j + 5]); o.w = pack2(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(buf + lane * 64 + ((static_cast<uint32_t>(j) ^ sw) << 4)) = o;
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (F32) tma_reduce_add_2d(tmap_c, buf, colw + cc, row0); else tma_store_2d(tmap_c, buf, colw + cc, row0);
      bulk_commit();
    }
    ++c.nstore;
  }
}

// all of this warp's TMA stores / reductions have been performed; make them visible to generic-proxy loads of the CTA
__device__ __forceinline__ void stores_complete(int lane) {
  if (lane == 0) {
    bulk_wait_all();
    fence_proxy_async_all();
  }
  __syncwarp();
}

// LayerNorm of one row held as lane-strided float4s (lane owns columns j*128 + lane*4 + {0..3})
template <int NJ>
__device__ __forceinline__ void warp_ln(const float4 (&v)[NJ], const float* __restrict__ gamma, const float* __restrict__ beta, int lane,
                                        float4 (&o)[NJ]) {
  constexpr float inv_d = 1.0f / (NJ * 128);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * inv_d + kLnEps);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = j * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    o[j].x = (v[j].x - mean) * rstd * g.x + b.x;
    o[j].y = (v[j].y - mean) * rstd * g.y + b.y;
    o[j].z = (v[j].z - mean) * rstd * g.z + b.z;
    o[j].w = (v[j].w - mean) * rstd * g.w + b.w;
  }
}
__device__ __forceinline__ uint2 f4_to_bf4(float4 v) {
  uint2 u;
  u.x = pack2(v.x, v.y);
  u.y = pack2(v.z, v.w);
  return u;
}

// LayerNorm of this warp's 8 rows of the CTA tile.  to_enc: write the token-major ENC matrix (row = token * B + b) instead of Y.
template <int NJ>
__device__ __forceinline__ void ln_rows(const MegaParams& p, int tile, int ew, int lane, const float* gamma, const float* beta, bool to_enc) {
  constexpr int D = NJ * 128;
  const int used = p.cpt * p.S;
#pragma unroll 1
  for (int rb = 0; rb < 8; rb += 2) {  // two rows in flight
    float4 v[2][NJ];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int lr = ew * 8 + rb + i;
      if (lr < used) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          v[i][j] = __ldcg(reinterpret_cast<const float4*>(p.X + (static_cast<size_t>(tile) * BM + lr) * D + j * 128 + lane * 4));
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int lr = ew * 8 + rb + i;
      if (lr < used) {
        float4 o[NJ];
        warp_ln<NJ>(v[i], gamma, beta, lane, o);
        __nv_bfloat16* dst;
        if (to_enc) {
          const int c = lr / p.S, s = lr - c * p.S, b = tile * p.cpt + c;
          if (b >= p.B) continue;
          dst = p.ENC + (static_cast<size_t>(s) * p.B + b) * D;
        } else {
          dst = p.Y + (static_cast<size_t>(tile) * BM + lr) * D;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) *reinterpret_cast<uint2*>(dst + j * 128 + lane * 4) = f4_to_bf4(o[j]);
      }
    }
  }
}

// Bidirectional attention of ONE (candidate, head) pair by one warp, S <= 16 tokens, head_dim 128.
// Q K^T: both operands straight from global memory into mma fragments (16-byte loads, consistent k permutation: lane (g, t)
// holds k = 32 i + 8 t .. + 7 of rows g and g + 8); P V: V staged in the warp's scratch (ldmatrix.trans), P = the score
// accumulators re-used as the A fragment; O staged in scratch and copied out with 16-byte stores.
__device__ __forceinline__ void attend_pair(const __nv_bfloat16* __restrict__ qkv_rows, int ld, int D, int h, int S, uint8_t* scratch,
                                            __nv_bfloat16* __restrict__ out_rows, int ldo, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const uint32_t uV = smem_u32(scratch);
  // V (16 x 128 bf16, pad rows zero) -> scratch, asynchronously
  for (int idx = lane; idx < 16 * 16; idx += 32) {
    const int r = idx >> 4, c = idx & 15;
    if (r < S)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uV + swz(r, c)),
                   "l"(qkv_rows + static_cast<size_t>(r) * ld + 2 * D + h * 128 + c * 8)
                   : "memory");
    else
      *reinterpret_cast<uint4*>(scratch + swz(r, c)) = make_uint4(0, 0, 0, 0);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // scores = Q K^T (16 x 16)
  const __nv_bfloat16* qa = qkv_rows + static_cast<size_t>(g) * ld + h * 128 + t * 8;
  const __nv_bfloat16* qb = qa + static_cast<size_t>(8) * ld;
  const __nv_bfloat16* ka = qa + D;
  const __nv_bfloat16* kb = qb + D;
  uint4 q0[4], q1[4], k0[4], k1[4];
  const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    q0[i] = g < S ? __ldcg(reinterpret_cast<const uint4*>(qa + 32 * i)) : z;
    q1[i] = g + 8 < S ? __ldcg(reinterpret_cast<const uint4*>(qb + 32 * i)) : z;
    k0[i] = g < S ? __ldcg(reinterpret_cast<const uint4*>(ka + 32 * i)) : z;
    k1[i] = g + 8 < S ? __ldcg(reinterpret_cast<const uint4*>(kb + 32 * i)) : z;
  }
  float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mma16816(sc[0], q0[i].x, q1[i].x, q0[i].y, q1[i].y, k0[i].x, k0[i].y);
    mma16816(sc[0], q0[i].z, q1[i].z, q0[i].w, q1[i].w, k0[i].z, k0[i].w);
    mma16816(sc[1], q0[i].x, q1[i].x, q0[i].y, q1[i].y, k1[i].x, k1[i].y);
    mma16816(sc[1], q0[i].z, q1[i].z, q0[i].w, q1[i].w, k1[i].z, k1[i].w);
  }
  // softmax over keys (rows g and g + 8); pad keys masked
  const float sl2 = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int key = 8 * j + 2 * t;
    if (key >= S) sc[j][0] = sc[j][2] = -INFINITY;
    if (key + 1 >= S) sc[j][1] = sc[j][3] = -INFINITY;
    m0 = fmaxf(m0, fmaxf(sc[j][0], sc[j][1]));
    m1 = fmaxf(m1, fmaxf(sc[j][2], sc[j][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    sc[j][0] = exp2f((sc[j][0] - m0) * sl2); sc[j][1] = exp2f((sc[j][1] - m0) * sl2);
    sc[j][2] = exp2f((sc[j][2] - m1) * sl2); sc[j][3] = exp2f((sc[j][3] - m1) * sl2);
    l0 += sc[j][0] + sc[j][1];
    l1 += sc[j][2] + sc[j][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  const uint32_t pa0 = pack2(sc[0][0] * inv0, sc[0][1] * inv0), pa1 = pack2(sc[0][2] * inv1, sc[0][3] * inv1);
  const uint32_t pa2 = pack2(sc[1][0] * inv0, sc[1][1] * inv0), pa3 = pack2(sc[1][2] * inv1, sc[1][3] * inv1);
  // O = P V (16 x 128)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  float o[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int np = 0; np < 8; ++np) {  // two 8-wide dim tiles per ldmatrix.x4.trans
    uint32_t b0, b1, b2, b3;
    const int r = (lane & 7) + 8 * ((lane >> 3) & 1), c = 2 * np + (lane >> 4);
    ldsm_x4_t(uV + swz(r, c), b0, b1, b2, b3);
    mma16816(o[2 * np], pa0, pa1, pa2, pa3, b0, b1);
    mma16816(o[2 * np + 1], pa0, pa1, pa2, pa3, b2, b3);
  }
  __syncwarp();  // every lane has read V: the tile is reused for O
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    *reinterpret_cast<uint32_t*>(scratch + swz(g, n) + 4 * t) = pack2(o[n][0], o[n][1]);
    *reinterpret_cast<uint32_t*>(scratch + swz(g + 8, n) + 4 * t) = pack2(o[n][2], o[n][3]);
  }
  __syncwarp();
  for (int idx = lane; idx < S * 16; idx += 32) {
    const int r = idx >> 4, c = idx & 15;
    *reinterpret_cast<uint4*>(out_rows + static_cast<size_t>(r) * ldo + h * 128 + c * 8) = *reinterpret_cast<const uint4*>(scratch + swz(r, c));
  }
  __syncwarp();  // scratch is free again
}

template <int NJ>
__global__ void __launch_bounds__(MG_THREADS, 1) encoder_mega_kernel(const __grid_constant__ MegaParams p) {
  constexpr int D = NJ * 128, F = 4 * D, H = D / 128;
  constexpr int KB1 = D / BK, KB4 = F / BK;                            // k-blocks of the K = D and K = 4D GEMMs
  constexpr int NT_QKV = 3 * D / MG_BN, NT_D = D / MG_BN, NT_F = F / MG_BN;  // 256-wide n-tiles
  using L = Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + MG_STAGES;
  uint64_t* acc_full = empty_bar + MG_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* ready = acc_empty + 2;  // [0] ATT written, [1] Y = LN2 written, [2] HID stored, [3] Y = post-LN written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int s = 0; s < MG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 2 * MG_EPI_WARPS);
    mbar_init(&acc_empty[1], 2 * MG_EPI_WARPS);
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&ready[i], MG_EPI_WARPS);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * MG_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PDL_PROLOGUE();

  if (warp == 0) {
    if (lane == 0) {
      // =============================================================== TMA producer
      const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[0]), 0);
      int s = 0;
      uint32_t ph = 0, lc = 0;
      auto gemm_loads = [&](const CUtensorMap* ta, const CUtensorMap* tw, int row0, int n_tiles, int num_kb) {
        for (int nt = 0; nt < n_tiles; ++nt) {
          const int n0 = nt * MG_BN + static_cast<int>(crank) * 128;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (leader) mbar_arrive_expect_tx(&full_bar[s], 2u * L::kStageBytes);
            uint8_t* dst = smem + s * L::kStageBytes;
            const uint32_t bar = full_leader + 8u * static_cast<uint32_t>(s);
            tma_load_2d_2sm(dst, ta, bar, kb * BK, row0);
            tma_load_2d_2sm(dst + L::kABlk, tw, bar, kb * BK, n0);
            if (++s == MG_STAGES) { s = 0; ph ^= 1; }
          }
        }
      };
      for (int pt = pair; pt < p.n_pair_tiles; pt += n_pairs) {
        const int row0 = (pt * 2 + static_cast<int>(crank)) * BM;
        for (int l = 0; l < p.n_layers; ++l, ++lc) {
          const MegaLayer& w = p.layer[l];
          gemm_loads(&p.a_y, &w.w_in, row0, NT_QKV, KB1);
          mbar_wait(&ready[0], lc & 1);
          gemm_loads(&p.a_att, &w.w_out, row0, NT_D, KB1);
          mbar_wait(&ready[1], lc & 1);
          gemm_loads(&p.a_y, &w.w_l1, row0, NT_F, KB1);
          mbar_wait(&ready[2], lc & 1);
          gemm_loads(&p.a_hid, &w.w_l2, row0, NT_D, KB4);
          mbar_wait(&ready[3], lc & 1);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // =============================================================== MMA issuer
      constexpr uint32_t idesc = make_idesc_mn(2 * BM, MG_BN);
      int s = 0, it = 0;
      uint32_t ph = 0;
      auto gemm_mma = [&](int n_tiles, int num_kb) {
        for (int nt = 0; nt < n_tiles; ++nt, ++it) {
          const int a = it & 1;
          const uint32_t aph = (it >> 1) & 1;
          mbar_wait(&acc_empty[a], aph ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(a * MG_BN);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
            const uint32_t b_addr = a_addr + L::kABlk;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_2sm(tmem_d, make_smem_desc(a_addr + k * UMMA_K * 2), make_smem_desc(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[s]);
            if (++s == MG_STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit_2sm(&acc_full[a]);
        }
      };
      for (int pt = pair; pt < p.n_pair_tiles; pt += n_pairs)
        for (int l = 0; l < p.n_layers; ++l) {
          gemm_mma(NT_QKV, KB1);
          gemm_mma(NT_D, KB1);
          gemm_mma(NT_F, KB1);
          gemm_mma(NT_D, KB4);
        }
    }
  } else {
    // =============================================================== epilogue + compute warps
    const int ew = warp - 2;
    EpiCtx c;
    c.scratch = smem + L::kScratchOffset + ew * L::kScratchPerWarp;
    c.sbias = reinterpret_cast<float*>(smem + L::kBiasOffset) + ew * 64;
    c.nstore = 0;
    c.tmem_base = tmem_base;
    c.acc_empty_leader = mapa_u32(smem_u32(&acc_empty[0]), 0);
    c.lane = lane;
    c.quarter = warp & 3;
    c.cgrp = ew >> 2;
    int it = 0;
    auto wait_acc = [&]() {
      const int a = it & 1;
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      ++it;
      return a;
    };
    for (int pt = pair; pt < p.n_pair_tiles; pt += n_pairs) {
      const int tile = pt * 2 + static_cast<int>(crank);
      const int row0 = tile * BM + c.quarter * 32;  // first of this warp's 32 accumulator rows
      const int ncand = max(0, min(p.cpt, p.B - tile * p.cpt));
      for (int l = 0; l < p.n_layers; ++l) {
        const MegaLayer& w = p.layer[l];
        const bool last = l + 1 == p.n_layers;
        // ---- G1: QKV = Y W_in^T + b  (bf16 store) ----
        for (int nt = 0; nt < NT_QKV; ++nt) {
          const int a = wait_acc();
          drain_unit<false, false>(c, &p.c_qkv, w.in_b, row0, nt * MG_BN + c.cgrp * 64, a);
        }
        stores_complete(lane);
        epi_bar_sync();
        // ---- C1: attention, one (candidate, head) pair per warp ----
        for (int pr = ew; pr < ncand * H; pr += MG_EPI_WARPS) {
          const int cand = pr / H, h = pr - cand * H;
          const size_t r = static_cast<size_t>(tile) * BM + cand * p.S;
          attend_pair(p.QKV + r * 3 * D, 3 * D, D, h, p.S, c.scratch, p.ATT + r * D, D, lane);
        }
        fence_proxy_async_all();  // generic stores of ATT -> visible to the TMA loads of the next GEMM
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[0]);
        // ---- G2: X += ATT W_out^T + b  (fp32 reduce-add) ----
        for (int nt = 0; nt < NT_D; ++nt) {
          const int a = wait_acc();
          drain_unit<true, false>(c, &p.c_x, w.out_b, row0, nt * MG_BN + c.cgrp * 64, a);
        }
        stores_complete(lane);
        epi_bar_sync();
        // ---- C2: Y = LN2(X) ----
        ln_rows<NJ>(p, tile, ew, lane, w.n2_w, w.n2_b, false);
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[1]);
        // ---- G3: HID = gelu(Y W1^T + b1)  (bf16 store) ----
        for (int nt = 0; nt < NT_F; ++nt) {
          const int a = wait_acc();
          drain_unit<false, true>(c, &p.c_hid, w.l1_b, row0, nt * MG_BN + c.cgrp * 64, a);
        }
        stores_complete(lane);
        if (lane == 0) mbar_arrive(&ready[2]);
        // ---- G4: X += HID W2^T + b2  (fp32 reduce-add) ----
        for (int nt = 0; nt < NT_D; ++nt) {
          const int a = wait_acc();
          drain_unit<true, false>(c, &p.c_x, w.l2_b, row0, nt * MG_BN + c.cgrp * 64, a);
        }
        stores_complete(lane);
        epi_bar_sync();
        // ---- C3: Y = norm1 of the next layer (X), or ENC = final encoder norm (token-major) ----
        ln_rows<NJ>(p, tile, ew, lane, w.post_w, w.post_b, last);
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[3]);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * MG_BN) : "memory");
  }
}

}  // namespace

int encoder_mega_cpt(int S) { return S >= 1 && S <= 16 ? 128 / S : 0; }

int launch_encoder_mega(const EncoderMegaArgs& a, cudaStream_t st) {
  M3PC_REQUIRE(a.D == 512, "encoder_mega: n_embd must be 512");
  M3PC_REQUIRE(a.S >= 1 && a.S <= 16 && a.n_layers >= 1 && a.n_layers <= MG_MAX_LAYERS && a.B >= 1, "encoder_mega: shape out of range");
  M3PC_TRY(gemm_init_driver_api());
  static_assert(Smem::kTotal <= 227 * 1024, "shared memory budget exceeded");
  static bool configured = false;
  if (!configured) {
    M3PC_CHECK_CUDA(cudaFuncSetAttribute(encoder_mega_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kTotal));
    configured = true;
  }
  const int D = a.D, F = 4 * D;
  const int cpt = encoder_mega_cpt(a.S);
  const int n_tiles = ceil_div(a.B, cpt);
  const int n_pair_tiles = ceil_div(n_tiles, 2);
  const uint64_t R = static_cast<uint64_t>(n_pair_tiles) * 2 * BM;  // rows of every tile-layout buffer
  MegaParams p{};
  p.n_layers = a.n_layers; p.S = a.S; p.cpt = cpt; p.B = a.B; p.n_pair_tiles = n_pair_tiles; p.D = D;
  p.X = a.X; p.Y = a.Y; p.QKV = a.QKV; p.ATT = a.ATT; p.ENC = a.ENC;
  M3PC_TRY(make_tmap(&p.a_y, a.Y, R, D, BM));
  M3PC_TRY(make_tmap(&p.a_att, a.ATT, R, D, BM));
  M3PC_TRY(make_tmap(&p.a_hid, a.HID, R, F, BM));
  M3PC_TRY(make_tmap_out(&p.c_qkv, a.QKV, R, 3 * D, false));
  M3PC_TRY(make_tmap_out(&p.c_hid, a.HID, R, F, false));
  M3PC_TRY(make_tmap_out(&p.c_x, a.X, R, D, true));
  for (int l = 0; l < a.n_layers; ++l) {
    const EncoderMegaLayer& s = a.layer[l];
    MegaLayer& w = p.layer[l];
    M3PC_TRY(make_tmap(&w.w_in, s.in_w, 3 * D, D, 128));
    M3PC_TRY(make_tmap(&w.w_out, s.out_w, D, D, 128));
    M3PC_TRY(make_tmap(&w.w_l1, s.l1_w, F, D, 128));
    M3PC_TRY(make_tmap(&w.w_l2, s.l2_w, D, F, 128));
    w.in_b = s.in_b; w.out_b = s.out_b; w.l1_b = s.l1_b; w.l2_b = s.l2_b;
    w.n2_w = s.n2_w; w.n2_b = s.n2_b; w.post_w = s.post_w; w.post_b = s.post_b;
  }
  const int pairs = std::min(n_pair_tiles, gemm_num_sms() / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = Smem::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
  M3PC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, encoder_mega_kernel<4>, p));
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}

}  // namespace m3pc
