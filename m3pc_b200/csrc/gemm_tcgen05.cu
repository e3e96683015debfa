// bf16 GEMM for sm_100a: C[M,N] = epilogue(A[M,K] * W[N,K]^T), fp32 accumulation in TMEM.
//
// This is the dense-contraction kernel of the MTM encoder/decoder blocks (reference call sites:
// nn.TransformerEncoderLayer QKV / out-proj / linear1 / linear2 at omtm/models/mtm_model.py:379-409,
// decoder_embed_dict at :646-661, output_head_dict[*].1 at :428-433).
//
// Persistent, warp-specialised, one CTA per SM (grid = min(tiles, #SM)); each CTA walks output tiles
// (128 x BN, n fastest so concurrently running CTAs share A panels and the whole weight in L2):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of the A (128 x 64) and W (BN x 64) k-blocks into a
//               STAGES-deep shared-memory ring (SWIZZLE_128B), completion on `full` mbarriers.
//   warp 1      allocates TMEM (2 accumulators of BN fp32 columns), then one lane issues tcgen05.mma
//               (kind::f16, bf16 x bf16 -> fp32, M=128, N=BN, K=16 per instruction) from smem descriptors;
//               tcgen05.commit releases ring slots (`empty`) and publishes the accumulator (`acc_full`).
//   warps 2..9  epilogue, overlapped with the next tile's main loop through the second accumulator: tcgen05.ld
//               (each warp its own 32-lane TMEM quarter and half of the columns), transpose through a padded
//               smem tile so that global accesses are full 128-byte lines, fused bias / per-token table /
//               GELU(erf) / ReLU / residual add, 16-byte stores; `acc_empty` hands the accumulator back.
// Clusters: with CL > 1, CL CTAs of a thread-block cluster work on CL consecutive m-tiles of the same n-tile in lockstep;
// each loads 1/CL of the W k-block and TMA-multicasts it to all of them, so a W tile crosses the L2 -> SM fabric once per
// cluster instead of once per CTA (the GEMMs of this model are bound by that traffic, not by the tensor pipe: K is only
// 512..2048).  Ring slots are released cluster-wide: every CTA's tcgen05.commit multicast-arrives on all `empty` barriers.
// Both operands are K-major, which is the layout the activations (row-major (rows, K)) and nn.Linear weights
// ((out, in) row-major) already have -- no transposes anywhere.  Rows beyond M are zero-filled by TMA and masked
// in the epilogue, so M is arbitrary (ragged candidate counts, B = 1).
#include <cuda.h>

#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

namespace m3pc {

namespace {

constexpr int NUM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int XPOSE_LD = 36;  // floats per staged row: 128-bit accesses are bank-conflict free both row- and column-wise

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
int g_force_bn = 0, g_force_cl = 0;  // tuning build: M3PC_GEMM_CONFIG="<bn>x<cl>" pins one single-CTA configuration (tuning / tests)
#ifdef M3PC_TUNING
int g_debug_skip_epi = 0;             // tuning build only, M3PC_GEMM_DEBUG_SKIP_EPI=1: the epilogue warps only hand the accumulator back (results are garbage)
constexpr int EPI_DEBUG_SKIP = 1 << 30;
#else
constexpr int g_debug_skip_epi = 0;
constexpr int EPI_DEBUG_SKIP = 0;     // the release library has no result-corrupting switch
#endif
int g_use_2sm = 1;                   // tuning build: M3PC_GEMM_2SM=0 disables the CTA-pair kernel

struct EpiParams {
  const float* bias;
  const float* table;
  int rows_per_group;
  int flags;
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kXposeOffset = STAGES * kStageBytes;
  static constexpr int kXposeBytesPerWarp = 32 * XPOSE_LD * 4;
  static constexpr int kBarOffset = kXposeOffset + NUM_EPI_WARPS * kXposeBytesPerWarp;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

__device__ __forceinline__ float apply_act(float v, bool do_gelu, bool do_relu) {
  if (do_gelu) v = gelu_erf_fast(v);
  if (do_relu) v = fmaxf(v, 0.0f);
  return v;
}

// One epilogue warp's share of a 128 x BN accumulator tile: TMEM lanes [32*quarter, +32) (= rows), columns
// [half*BN/2, +BN/2), in chunks of 32 columns: tcgen05.ld -> padded smem transpose -> fused epilogue -> 16-byte coalesced stores.
template <int BN>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_acc, float* xp, void* C, int M, int N, int m0, int n0, int quarter, int half,
                                              int lane, const EpiParams& ep) {
  const bool do_gelu = ep.flags & EPI_GELU, do_relu = ep.flags & EPI_RELU, do_res = ep.flags & EPI_RESIDUAL;
  const bool out_f32 = do_res || (ep.flags & EPI_OUT_F32);
  const int row_base = m0 + quarter * 32;
#pragma unroll 1
  for (int c = 0; c < BN / 2; c += 32) {
    const int col_in_tile = half * (BN / 2) + c;
    const int col0 = n0 + col_in_tile;
    // residual rows are independent of the accumulator: issue all 8 coalesced loads first so their DRAM/L2 latency
    // overlaps the TMEM load and the transpose (lane -> 4 consecutive columns, 8 lanes per row, 4 rows per pass)
    float4 res[8];
    if (do_res) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row_base + 4 * i + (lane >> 3);
        res[i] = (row < M) ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(C) + static_cast<size_t>(row) * N + col0 + 4 * (lane & 7))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    uint32_t r[32];
    tmem_ld32(tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(col_in_tile), r);
    tmem_ld_wait();
    // row-per-lane -> staging tile (lane = row)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(xp + lane * XPOSE_LD + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    __syncwarp();
    if (out_f32) {
      // lane -> 4 consecutive columns, 8 lanes per row, 4 rows per pass
      const int cc = 4 * (lane & 7);
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + cc));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        const int row = row_base + rr;
        float4 v = *reinterpret_cast<const float4*>(xp + rr * XPOSE_LD + cc);
        if (row < M) {
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          if (ep.table != nullptr) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(ep.table + static_cast<size_t>(row / ep.rows_per_group) * N + col0 + cc));
            v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w;
          }
          v.x = apply_act(v.x, do_gelu, do_relu); v.y = apply_act(v.y, do_gelu, do_relu);
          v.z = apply_act(v.z, do_gelu, do_relu); v.w = apply_act(v.w, do_gelu, do_relu);
          float* cp = reinterpret_cast<float*>(C) + static_cast<size_t>(row) * N + col0 + cc;
          if (do_res) { v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w; }
          *reinterpret_cast<float4*>(cp) = v;
        }
      }
    } else {
      // lane -> 8 consecutive columns, 4 lanes per row, 8 rows per pass
      const int cc = 8 * (lane & 3);
      float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ep.bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + cc));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + cc + 4));
        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + (lane >> 2);
        const int row = row_base + rr;
        const float4 v0 = *reinterpret_cast<const float4*>(xp + rr * XPOSE_LD + cc);
        const float4 v1 = *reinterpret_cast<const float4*>(xp + rr * XPOSE_LD + cc + 4);
        if (row < M) {
          float v[8] = {v0.x + bb[0], v0.y + bb[1], v0.z + bb[2], v0.w + bb[3], v1.x + bb[4], v1.y + bb[5], v1.z + bb[6], v1.w + bb[7]};
          if (ep.table != nullptr) {
            const float* tr = ep.table + static_cast<size_t>(row / ep.rows_per_group) * N + col0 + cc;
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(tr));
            const float4 t1 = __ldg(reinterpret_cast<const float4*>(tr + 4));
            v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w; v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], do_gelu, do_relu);
          uint4 o;
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]);
          __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2], v[3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]);
          __nv_bfloat162 p3 = __floats2bfloat162_rn(v[6], v[7]);
          o.x = *reinterpret_cast<uint32_t*>(&p0);
          o.y = *reinterpret_cast<uint32_t*>(&p1);
          o.z = *reinterpret_cast<uint32_t*>(&p2);
          o.w = *reinterpret_cast<uint32_t*>(&p3);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(C) + static_cast<size_t>(row) * N + col0 + cc) = o;
        }
      }
    }
    __syncwarp();  // staging tile is reused by the next chunk
  }
}

template <int BN, int STAGES, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                    const __grid_constant__ CUtensorMap tmap_w,
                                                                    void* __restrict__ C, int M, int N, int K, EpiParams ep) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = K / BK;
  const int n_tiles = N / BN;
  // work unit = CL consecutive m-tiles x one n-tile; the CTAs of a cluster walk the same unit sequence in lockstep
  const int total_tiles = n_tiles * (((M + BM - 1) / BM + CL - 1) / CL);
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  const int unit0 = blockIdx.x / CL, unit_step = gridDim.x / CL;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);  // one tcgen05.commit arrival from every CTA of the cluster
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], NUM_EPI_WARPS);
    mbar_init(&acc_empty[1], NUM_EPI_WARPS);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // peers must see initialised barriers before any multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PDL_PROLOGUE();  // everything above is private to this CTA; operands of the previous kernel are touched only below

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        const int m0 = ((tile / n_tiles) * CL + static_cast<int>(crank)) * BM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);  // every CTA of the cluster has consumed this slot
          mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
          uint8_t* sa = smem + s * L::kStageBytes;
          tma_load_2d(sa, &tmap_a, &full_bar[s], kb * BK, m0);
          if (CL == 1) {
            tma_load_2d(sa + L::kABytes, &tmap_w, &full_bar[s], kb * BK, n0);
          } else {  // my 1/CL slice of the W k-block, delivered to every CTA of the cluster
            constexpr int kSliceRows = BN / CL;
            tma_load_2d_mc(sa + L::kABytes + crank * (kSliceRows * BK * 2), &tmap_w, &full_bar[s], kb * BK,
                           n0 + static_cast<int>(crank) * kSliceRows, kMask);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step, ++it) {
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&acc_empty[a], aph ^ 1);  // epilogue has drained this accumulator (first use passes immediately)
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(a * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
          const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance inside the 128-byte swizzle atom: +32 bytes per K=16 step
            umma_bf16(tmem_d, make_smem_desc(a_addr + k * UMMA_K * 2), make_smem_desc(b_addr + k * UMMA_K * 2), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          // slot reusable once these MMAs have read it -- in every CTA of the cluster, since peers multicast into it
          if (CL == 1) umma_commit(&empty_bar[s]); else umma_commit_mc(&empty_bar[s], kMask);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[a]);  // accumulator complete
      }
    }
  } else {
    // ---- epilogue warps: TMEM -> registers -> smem transpose -> coalesced global ----
    const int ew = warp - 2;
    const int quarter = warp & 3;       // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
    const int half = ew >> 2;           // which half of the BN columns
    float* xp = reinterpret_cast<float*>(smem + L::kXposeOffset + ew * L::kXposeBytesPerWarp);
    int it = 0;
    for (int tile = unit0; tile < total_tiles; tile += unit_step, ++it) {
      const int m0 = ((tile / n_tiles) * CL + static_cast<int>(crank)) * BM, n0 = (tile % n_tiles) * BN;
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
      epilogue_tile<BN>(tmem_base + static_cast<uint32_t>(a * BN), xp, C, M, N, m0, n0, quarter, half, lane, ep);
      // all of this warp's TMEM reads are complete (tcgen05.wait::ld above): hand the accumulator back
      tc_fence_before();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // no CTA may exit while a peer can still multicast into / arrive on it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}


// ======================================================================================================================
// CTA-pair kernel (cta_group::2) -- the production path for M > 128, N % 256 == 0.
// A cluster of two CTAs owns a 256 x 256 output tile: one tcgen05.mma of the leader multiplies the pair's 256 x 16 A slab
// with a 256-wide W slab of which each CTA holds (and loads) only half, so a CTA stages 32 KB per k-block instead of 48 KB.
// Epilogue (measured to be what bounds these K = 512 .. 2048 GEMMs, profiles/r1b_*): each of the 8 epilogue warps drains its
// 32-lane x 128-column share of the accumulator in 32-column chunks -- tcgen05.ld (lane = row) -> bias / row table /
// activation in registers -> 16-byte shared-memory stores into a hardware-swizzled staging box (conflict free) -> one TMA
// bulk tensor store per chunk, double buffered.  The residual epilogue never loads C: the staged tile is added to the fp32
// residual stream in L2 by a TMA reduce-add (cp.reduce.async.bulk.tensor .add), one rounding, same result as load-add-store.
// The bias of the tile's columns is staged in shared memory once per tile, before the accumulator is waited for.
constexpr int NUM_EPI_WARPS_2SM = 16;  // 4 per TMEM lane quarter: enough warps per scheduler to hide TMEM / fence / store latencies
constexpr int GEMM_THREADS_2SM = 32 * (2 + NUM_EPI_WARPS_2SM);

template <int STAGES>
struct Smem2 {
  static constexpr int kABlk = BM * BK * 2;   // this CTA's 128 rows of one A k-block
  static constexpr int kBBlk = 128 * BK * 2;  // this CTA's 128 of the 256 W rows of one k-block
  static constexpr int kStageBytes = kABlk + kBBlk;
  static constexpr int kStoreBufBytes = 32 * 64;  // one staged chunk: 32 rows x 64 bytes (16 fp32 or 32 bf16 columns), SWIZZLE_64B
  static constexpr int kStoreOffset = STAGES * kStageBytes;
  static constexpr int kBiasOffset = kStoreOffset + NUM_EPI_WARPS_2SM * 2 * kStoreBufBytes;
  static constexpr int kBarOffset = kBiasOffset + NUM_EPI_WARPS_2SM * 64 * 4;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

// Grouped launches: up to MAX_GROUP independent problems (own operands, output, epilogue, shapes) share one launch; the pair's
// unit range simply runs across problem boundaries, so several small GEMMs of the decoder / heads / critic cost one kernel
// boundary instead of one each.  Split-K (ksplit = 2, residual epilogue only): a unit is one K half of an output tile and
// both halves are reduce-added into the fp32 residual stream by the TMA (bias / row table ride on half 0) -- used where
// whole tiles would leave most pairs idle in the last round (N = D GEMMs with K = 4 D at ~1024 batch rows).
constexpr int MAX_GROUP = 4;
struct alignas(64) GroupProblem {
  CUtensorMap ta, tw, tc;
  const float* bias;
  const float* table;
  int rows_per_group, flags;
  int M, n_tiles;  // rows; N / 256
  int num_kb;      // k-blocks per unit (K / 64 / ksplit)
  int ksplit;      // 1 or 2
  int unit0;       // first unit of this problem in the launch-wide unit order
};
struct GroupParams {
  GroupProblem p[MAX_GROUP];
  int n, total_units;
  int strided;  // 1: pair p works on units p, p + n_pairs, ... (see launch_2sm) instead of one contiguous range
  int tune;     // tuning build only (timing experiments, results are garbage): bit 0 = load A for a pair's first unit only,
                // bit 1 = load W for a pair's first unit only
};
struct UnitCoord { int g, mp, nt, ks; };
__device__ __forceinline__ UnitCoord decode_unit(const GroupParams& gp, int u) {
  UnitCoord c;
  c.g = 0;
#pragma unroll
  for (int i = 1; i < MAX_GROUP; ++i)
    if (i < gp.n && u >= gp.p[i].unit0) c.g = i;
  const GroupProblem& P = gp.p[c.g];
  int lu = u - P.unit0;
  c.ks = 0;
  if (P.ksplit == 2) { c.ks = lu & 1; lu >>= 1; }
  c.mp = lu / P.n_tiles;
  c.nt = lu - c.mp * P.n_tiles;
  return c;
}

template <int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS_2SM, 1) gemm_bf16_2sm_kernel(const __grid_constant__ GroupParams gp) {
  constexpr int BN = 256;
  using L = Smem2<STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);  // waited on by the leader's MMA thread only
  uint64_t* empty_bar = full_bar + STAGES;                                 // ring slot free (signalled in both CTAs)
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;                                      // leader's copy counts both CTAs' epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int total_units = gp.total_units;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // units are ordered m-major and split contiguously: a pair mostly stays on one row block, whose A slab stays hot in L2
  // strided (N = 512 problems with a long K): neighbouring pairs take the two column tiles of the SAME row block at the same
  // time, so its A slab is fetched from HBM once and the second reader hits L2 (contiguous ranges re-read it ~150 MB later)
  const int u_step = gp.strided ? n_pairs : 1;
  const int u_lo = gp.strided ? pair : static_cast<int>(static_cast<long long>(pair) * total_units / n_pairs);
  const int u_hi = gp.strided ? total_units : static_cast<int>(static_cast<long long>(pair + 1) * total_units / n_pairs);

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < gp.n; ++g) {
      prefetch_tmap(&gp.p[g].ta);
      prefetch_tmap(&gp.p[g].tw);
      prefetch_tmap(&gp.p[g].tc);
    }
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 2 * NUM_EPI_WARPS_2SM);
    mbar_init(&acc_empty[1], 2 * NUM_EPI_WARPS_2SM);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {  // the same warp of both CTAs allocates (2 accumulators of 256 fp32 columns = all of TMEM)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PDL_PROLOGUE();  // everything above is private to this CTA pair; operands of the previous kernel are touched only below

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer (both CTAs): own half of every operand slab, completion signalled on the LEADER's barrier ----
      const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[0]), 0);
      int s = 0;
      uint32_t ph = 0;
      for (int u = u_lo; u < u_hi; u += u_step) {
        const UnitCoord uc = decode_unit(gp, u);
        const GroupProblem& P = gp.p[uc.g];
        const int m0 = (uc.mp * 2 + static_cast<int>(crank)) * BM, n0 = uc.nt * BN + static_cast<int>(crank) * 128;
        const int num_kb = P.num_kb, k0 = uc.ks * num_kb;
#ifdef M3PC_TUNING
        const bool load_a = !(gp.tune & 1) || u == u_lo, load_w = !(gp.tune & 2) || u == u_lo;
#else
        constexpr bool load_a = true, load_w = true;
#endif
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2u * ((load_a ? L::kABlk : 0) + (load_w ? L::kBBlk : 0)));
          uint8_t* dst = smem + s * L::kStageBytes;
          const uint32_t bar = full_leader + 8u * static_cast<uint32_t>(s);
          if (load_a) tma_load_2d_2sm(dst, &P.ta, bar, (k0 + kb) * BK, m0);
          if (load_w) tma_load_2d_2sm(dst + L::kABlk, &P.tw, bar, (k0 + kb) * BK, n0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ---- MMA issuer (leader only) ----
      constexpr uint32_t idesc = make_idesc_mn(2 * BM, BN);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int u = u_lo; u < u_hi; u += u_step, ++it) {
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&acc_empty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(a * BN);
        const int num_kb = gp.p[decode_unit(gp, u).g].num_kb;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
          const uint32_t b_addr = a_addr + L::kABlk;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_2sm(tmem_d, make_smem_desc(a_addr + k * UMMA_K * 2), make_smem_desc(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_2sm(&acc_full[a]);
      }
    }
  } else {
    // ---- epilogue warps (both CTAs): 32 rows (TMEM lane quarter) x 64 columns of this CTA's 128 x 256 tile each ----
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp may read
    const int cgrp = ew >> 2;      // which 64 of the tile's 256 columns
    uint8_t* sbuf = smem + L::kStoreOffset + ew * 2 * L::kStoreBufBytes;
    float* sbias = reinterpret_cast<float*>(smem + L::kBiasOffset) + ew * 64;
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(&acc_empty[0]), 0);
    const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);  // SWIZZLE_64B: 16-byte chunk j of row r lives at j ^ ((r >> 1) & 3)
    int it = 0;
    uint32_t nstore = 0;  // chunks staged so far (selects the staging buffer)
    for (int u = u_lo; u < u_hi; u += u_step, ++it) {
      const UnitCoord uc = decode_unit(gp, u);
      const GroupProblem& P = gp.p[uc.g];
      const CUtensorMap* tmap_c = &P.tc;
      const int M = P.M, N = P.n_tiles * BN;
      const bool first_k = uc.ks == 0;  // split-K: bias and row table are added by the first K half only
      EpiParams ep{first_k ? P.bias : nullptr, first_k ? P.table : nullptr, P.rows_per_group, P.flags};
      const bool do_gelu = ep.flags & EPI_GELU, do_relu = ep.flags & EPI_RELU, do_res = ep.flags & EPI_RESIDUAL;
      const bool out_f32 = do_res || (ep.flags & EPI_OUT_F32);
      const bool skip = ep.flags & EPI_DEBUG_SKIP;
      const int cw = out_f32 ? 16 : 32;  // columns per staged chunk (64 bytes of output)
      const int row0 = (uc.mp * 2 + static_cast<int>(crank)) * BM + quarter * 32;
      const int colw = uc.nt * BN + cgrp * 64;  // first column of this warp's share
      // bias of my 64 columns -> shared memory (the previous tile's reads are done: same warp, program order)
      {
        float2 b2 = make_float2(0.f, 0.f);
        if (ep.bias != nullptr) b2 = __ldg(reinterpret_cast<const float2*>(ep.bias + colw + 2 * lane));
        *reinterpret_cast<float2*>(sbias + 2 * lane) = b2;
        __syncwarp();
      }
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
      const bool live = row0 < M && !skip;
      const int row = row0 + lane;
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(a * BN + cgrp * 64);
      if (live) {
#pragma unroll 1
        for (int c = 0; c < 64; c += cw) {
          const int col0 = colw + c;
          uint32_t r[32];
          if (out_f32) tmem_ld16(tacc + static_cast<uint32_t>(c), r); else tmem_ld32(tacc + static_cast<uint32_t>(c), r);
          // the staging buffer about to be overwritten was handed to the TMA two chunks ago: wait until it has been read
          uint8_t* buf = sbuf + (nstore & 1) * L::kStoreBufBytes;
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          tmem_ld_wait();
          if (c + cw == 64) {  // every TMEM read of this tile has landed in registers: hand the accumulator back to the MMA warp
            tc_fence_before();
            if (lane == 0) mbar_arrive_remote(acc_empty_leader + 8u * static_cast<uint32_t>(a));
          }
          if (out_f32) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(sbias + c + 4 * j);
              v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b4.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
            }
            if (ep.table != nullptr && row < M) {
              const float* tr = ep.table + static_cast<size_t>(row / ep.rows_per_group) * N + col0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(tr + 4 * j));
                v[4 * j + 0] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
              }
            }
            if (do_gelu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = gelu_erf_tanh(v[j]);
            }
            if (do_relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(buf + lane * 64 + ((static_cast<uint32_t>(j) ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(sbias + c + 4 * j);
              v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b4.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
            }
            if (ep.table != nullptr && row < M) {
              const float* tr = ep.table + static_cast<size_t>(row / ep.rows_per_group) * N + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(tr + 4 * j));
                v[4 * j + 0] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
              }
            }
            if (do_gelu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf_tanh(v[j]);
            }
            if (do_relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
              __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
              __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
              uint4 o;
              o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
              o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(buf + lane * 64 + ((static_cast<uint32_t>(j) ^ sw) << 4)) = o;
            }
          }
          fence_proxy_async();  // my generic-proxy writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            if (do_res) tma_reduce_add_2d(tmap_c, buf, col0, row0); else tma_store_2d(tmap_c, buf, col0, row0);
            bulk_commit();
          }
          ++nstore;
        }
      } else {  // nothing to store (rows beyond M, or the tuning switch): just release the accumulator
        tc_fence_before();
        if (lane == 0) mbar_arrive_remote(acc_empty_leader + 8u * static_cast<uint32_t>(a));
      }
    }
    if (lane == 0) bulk_wait_read<0>();  // shared memory may be released once the stores have been READ; the writes complete with the grid
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA may exit while its peer can still read its shared memory or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

}  // namespace
int make_tmap(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " (rows=" +
              std::to_string(rows) + ", cols=" + std::to_string(cols) + ")");
    return M3PC_ERR_CUDA;
  }
  return M3PC_OK;
}
namespace {

template <int BN, int STAGES, int CL>
int launch(const __nv_bfloat16* A, const __nv_bfloat16* W, void* C, int M, int N, int K, const GemmEpilogue& epi, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  static_assert(L::kTotal <= 227 * 1024, "shared memory budget exceeded");
  static PerDevice<bool> configured;
  if (!configured.here()) {
    M3PC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured.here() = true;
  }
  CUtensorMap ta, tw;
  M3PC_TRY(make_tmap(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), BM));
  M3PC_TRY(make_tmap(&tw, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), BN / CL));
  EpiParams ep{epi.bias, epi.table, epi.rows_per_group > 0 ? epi.rows_per_group : 1, epi.flags};
  const int units = (N / BN) * ceil_div(ceil_div(M, BM), CL);
  const int clusters = std::min(units, device_num_sms() / CL);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
  M3PC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, STAGES, CL>, ta, tw, C, M, N, K, ep));
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}


// output tensor map: boxes of 32 rows x 64 bytes (16 fp32 / 32 bf16 columns), SWIZZLE_64B
}  // namespace
int make_tmap_out(CUtensorMap* map, void* ptr, uint64_t rows, uint64_t cols, bool f32) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * (f32 ? 4u : 2u)};
  cuuint32_t box[2] = {f32 ? 16u : 32u, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return M3PC_ERR_CUDA;
  }
  return M3PC_OK;
}
namespace {

int g_use_strided = 1;  // tuning build: M3PC_GEMM_STRIDED=0 restores contiguous unit ranges for every problem
int g_use_splitk = 0;  // tuning build: M3PC_GEMM_SPLITK=1 enables split-K (measured +0.4% plans/s at 1024 candidates; off by default: the reduce-add order of the two halves is not reproducible run to run)

// One launch of the CTA-pair kernel over `n` problems (n <= MAX_GROUP; every N a multiple of 256).
template <int STAGES>
int launch_2sm(const GemmProblem* probs, int n, cudaStream_t st) {
  using L = Smem2<STAGES>;
  static_assert(L::kTotal <= 227 * 1024, "shared memory budget exceeded");
  static_assert(sizeof(GroupParams) <= 4000, "kernel parameter space");
  static PerDevice<bool> configured;
  if (!configured.here()) {
    M3PC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_2sm_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured.here() = true;
  }
  const int n_pairs_max = device_num_sms() / 2;
  GroupParams gp{};
  gp.n = n;
  int units = 0;
  for (int g = 0; g < n; ++g) {
    const GemmProblem& q = probs[g];
    GroupProblem& P = gp.p[g];
    const bool res = (q.epi.flags & EPI_RESIDUAL) != 0;
    const bool out_f32 = res || (q.epi.flags & EPI_OUT_F32) != 0;
    M3PC_TRY(make_tmap(&P.ta, q.A, static_cast<uint64_t>(q.M), static_cast<uint64_t>(q.K), BM));
    M3PC_TRY(make_tmap(&P.tw, q.W, static_cast<uint64_t>(q.N), static_cast<uint64_t>(q.K), 128));
    M3PC_TRY(make_tmap_out(&P.tc, q.C, static_cast<uint64_t>(q.M), static_cast<uint64_t>(q.N), out_f32));
    P.bias = q.epi.bias;
    P.table = q.epi.table;
    P.rows_per_group = q.epi.rows_per_group > 0 ? q.epi.rows_per_group : 1;
    P.flags = q.epi.flags | (g_debug_skip_epi ? EPI_DEBUG_SKIP : 0);
    P.M = q.M;
    P.n_tiles = q.N / 256;
    const int tiles = P.n_tiles * ceil_div(q.M, 2 * BM);
    // split K in two where whole tiles quantise badly: single-problem residual GEMMs with a long K whose last round would
    // leave most pairs idle (e.g. 104 tiles on 74 pairs: 2 rounds of whole tiles vs 3 rounds of half tiles)
    P.ksplit = 1;
    if (g_use_splitk && n == 1 && res && q.K >= 1024 && (q.K / BK) % 2 == 0) {
      const double r1 = std::ceil(static_cast<double>(tiles) / n_pairs_max);
      const double r2 = 0.5 * std::ceil(2.0 * tiles / n_pairs_max);
      if (r2 + 0.25 <= r1) P.ksplit = 2;
    }
    P.num_kb = q.K / BK / P.ksplit;
    P.unit0 = units;
    units += tiles * P.ksplit;
  }
  gp.total_units = units;
  if (const char* t = tune_env("M3PC_TUNE_GEMM")) gp.tune = atoi(t);
  const int pairs = std::min(units, n_pairs_max);
  gp.strided = g_use_strided && n == 1 && gp.p[0].n_tiles == 2 && gp.p[0].ksplit == 1 && probs[0].K >= 1024 && pairs % 2 == 0 && units >= 2 * pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(GEMM_THREADS_2SM);
  cfg.dynamicSmemBytes = L::kTotal;
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
  M3PC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_2sm_kernel<STAGES>, gp));
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}

// Modelled time of one configuration: the slower of the tensor pipe (rounds of the persistent grid x MMA cycles per tile)
// and the L2 -> SM operand stream (measured ~10.5 TB/s on B200 for this access pattern, profiles/README.md).
double model_time(int M, int N, int K, int bn, int cl) {
  const int m_tiles = ceil_div(M, BM);
  const int units = (N / bn) * ceil_div(m_tiles, cl);
  const int clusters = std::min(units, device_num_sms() / cl);
  const double rounds = std::ceil(static_cast<double>(units) / clusters);
  const double kb = K / BK;
  const double t_mma = rounds * kb * 4.0 * (bn / 2.0) / 1.7e9 + 2.0e-6;                          // 64 / 128 cycles per K=16 step
  const double bytes = static_cast<double>(units) * cl * kb * (BM * BK * 2.0 + bn * BK * 2.0 / cl);  // per CTA: own A + 1/cl of W
  return std::max(t_mma, bytes / 10.5e12);
}

}  // namespace

int gemm_init_driver_api() {
  if (g_encode_tiled != nullptr) return M3PC_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  M3PC_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return M3PC_ERR_CUDA;
  }
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  if (const char* f = tune_env("M3PC_GEMM_2SM")) g_use_2sm = atoi(f);
  if (const char* f = tune_env("M3PC_GEMM_SPLITK")) g_use_splitk = atoi(f);
  if (const char* f = tune_env("M3PC_GEMM_STRIDED")) g_use_strided = atoi(f);
#ifdef M3PC_TUNING
  if (const char* f = tune_env("M3PC_GEMM_DEBUG_SKIP_EPI")) g_debug_skip_epi = atoi(f);
#endif
  if (const char* f = tune_env("M3PC_GEMM_CONFIG")) {
    if (sscanf(f, "%dx%d", &g_force_bn, &g_force_cl) != 2 || (g_force_bn != 128 && g_force_bn != 256) || (g_force_cl != 1 && g_force_cl != 2))
      g_force_bn = g_force_cl = 0;
  }
  return M3PC_OK;
}

int gemm_bf16_tcgen05(const __nv_bfloat16* A, const __nv_bfloat16* W, void* C, int M, int N, int K, const GemmEpilogue& epi,
                      cudaStream_t st) {
  M3PC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: empty problem");
  M3PC_REQUIRE(K % BK == 0, "gemm_bf16: K must be a multiple of 64");
  M3PC_REQUIRE(N % 128 == 0, "gemm_bf16: N must be a multiple of 128");
  M3PC_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "gemm_bf16: operands must be 16-byte aligned");
  M3PC_TRY(gemm_init_driver_api());
  if (M <= 32 && static_cast<size_t>(M) * K * 2 <= 160 * 1024) return gemm_bf16_skinny(A, W, C, M, N, K, epi, st);
  // CTA-pair kernel wherever a pair has two row tiles to work on and N splits into 256-wide tiles
  if (g_use_2sm && !g_force_bn && N % 256 == 0 && M > BM) {
    const GemmProblem q{A, W, C, M, N, K, epi};
    return launch_2sm<4>(&q, 1, st);
  }
  // pick tile width and cluster size by the modelled time (ties: wider tile, larger cluster = less L2 traffic)
  struct Cand { int bn, cl; } cands[] = {{256, 2}, {256, 1}, {128, 2}, {128, 1}};
  int best = -1;
  double best_t = 1e30;
  for (int i = 0; i < 4; ++i) {
    if (N % cands[i].bn != 0) continue;
    if (g_force_bn && (cands[i].bn != g_force_bn || cands[i].cl != g_force_cl)) continue;
    const double t = model_time(M, N, K, cands[i].bn, cands[i].cl);
    if (t < best_t * 0.98) { best_t = t; best = i; }
  }
  switch (best) {
    case 0: return launch<256, 3, 2>(A, W, C, M, N, K, epi, st);
    case 1: return launch<256, 3, 1>(A, W, C, M, N, K, epi, st);
    case 2: return launch<128, 4, 2>(A, W, C, M, N, K, epi, st);
    default: return launch<128, 4, 1>(A, W, C, M, N, K, epi, st);
  }
}

// Several independent GEMMs in one launch (see GroupParams).  Falls back to one launch per problem where the CTA-pair
// kernel does not apply (a problem with N % 256 != 0, only tiny M, pinned tuning configurations).
int gemm_bf16_grouped(const GemmProblem* probs, int n, cudaStream_t st) {
  M3PC_REQUIRE(n >= 1, "gemm_bf16_grouped: no problems");
  M3PC_TRY(gemm_init_driver_api());
  bool ok = g_use_2sm && !g_force_bn && n <= MAX_GROUP;
  int max_m = 0;
  for (int g = 0; g < n && ok; ++g) {
    const GemmProblem& q = probs[g];
    ok = q.M > 0 && q.N > 0 && q.K > 0 && q.N % 256 == 0 && q.K % BK == 0 && (reinterpret_cast<uintptr_t>(q.A) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(q.W) & 15) == 0 && (reinterpret_cast<uintptr_t>(q.C) & 15) == 0;
    max_m = std::max(max_m, q.M);
  }
  if (ok && max_m > BM && n > 1) return launch_2sm<4>(probs, n, st);
  for (int g = 0; g < n; ++g) M3PC_TRY(gemm_bf16_tcgen05(probs[g].A, probs[g].W, probs[g].C, probs[g].M, probs[g].N, probs[g].K, probs[g].epi, st));
  return M3PC_OK;
}

}  // namespace m3pc
