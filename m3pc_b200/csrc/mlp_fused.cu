// Fused transformer MLP for n_embd = 512:   X[M,512] (fp32, in place)  +=  GELU(Y[M,512] W1[2048,512]^T + b1) W2[512,2048]^T + b2
//
// Reference call site: linear1 -> GELU(erf) -> linear2 of nn.TransformerEncoderLayer and its residual add (mtm_model.py:379-409;
// norm_first: x = x + ff(norm2(x)), Y = norm2(x)).  As two GEMM launches the 2048-wide hidden activation is written to HBM by the
// first (4 KB per row) and read back by the second: 2 x 436 MB per encoder layer at 106 496 rows, a quarter of all the bytes a
// plan step moves.  Here it never leaves the SM pair:
//
//   unit = 128 rows (64 per CTA of a cta_group::2 pair), hidden processed in 8 chunks of 256 columns:
//     MMA1_j   Hacc[b] (128 x 256, fp32 in TMEM)  = Y_unit (128 x 512, resident in shared memory) . W1[256 j .. 256 j + 256, :]^T
//     EPI1_j   Hacc[b] -> + b1 -> GELU -> bf16 -> shared memory, in the K-major SWIZZLE_128B layout tcgen05.mma reads its A operand from
//     MMA2_j   Oacc (128 x 512, fp32 in TMEM)    += H_j (128 x 256) . W2[:, 256 j .. 256 j + 256]^T
//   EPI2       Oacc -> + b2 -> TMA reduce-add into the fp32 residual stream X
//
// Tensor memory: with M = 128 a CTA's 64 x N slice of D takes 128 lanes x N/2 columns (the "2x2" layout, see gemm_ln.cu), so
// Oacc is 256 columns and the two Hacc buffers 128 each: exactly the 512 columns of an SM.  (At 256 rows per pair -- the tile the
// plain GEMMs use -- Oacc alone would fill tensor memory; that is why the hidden cannot stay on chip at that tile size.)
// The MMA warp software-pipelines the two products: MMA1_{j+1} is issued before MMA2_j, so the tensor pipe computes the next
// hidden chunk while the 16 epilogue warps turn the current one into an operand.  Weights (4 MB bf16) stream from L2 through a
// 3-stage ring of 32 KB TMA boxes; per unit a CTA receives 2 MB of them, which is what bounds the kernel (L2 -> SM fabric).
// Arithmetic per element (k order, bias / GELU / bf16 rounding of the hidden, fp32 residual add) is that of the two-launch path:
// results are bit-identical to it (tests/test_gpu_kernels.py::test_fused_mlp_matches_the_two_launch_path).
//
// Measured (profiles/r2l_fused_mlp.txt, 106 496 rows, 1 B200): 438 us against 418 us for the two launches; 428 us when MMA2 does not
// wait for the hidden chunk, 397 us without the GELU arithmetic, 383 us without either -- i.e. the floor of this arrangement is the
// operand stream: every 128-row unit pulls all 4 MB of W1 / W2 through the L2 -> SM fabric (2 MB per CTA per unit, ~60 GB/s per SM
// sustained), where the 256 x 256-tile GEMMs move half as many operand bytes per FLOP.  With the LayerNorm that the two-launch path
// fuses into linear2 (gemm_ln.cu) the plan step is 4 % SLOWER with this kernel, so the engine keeps it off (option "fused_mlp").
#include "common.cuh"
#include "tcgen05.cuh"

namespace m3pc {

int make_tmap(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
int make_tmap_out(CUtensorMap* map, void* ptr, uint64_t rows, uint64_t cols, bool f32);

namespace {

constexpr int MF_D = 512, MF_F = 2048;
constexpr int MF_EPI_WARPS = 16;
constexpr int MF_THREADS = 32 * (2 + MF_EPI_WARPS);
constexpr int MF_STAGES = 3;
constexpr int MF_CHUNK = 256;                // hidden columns per chunk
constexpr int MF_NCHUNK = MF_F / MF_CHUNK;  // 8
constexpr int MF_KB1 = MF_D / BK;           // 8 k-blocks of the first product
constexpr int MF_KB2 = MF_CHUNK / BK;       // 4 k-blocks of the second product per chunk

struct SmemMf {
  static constexpr int kKBlk = 64 * BK * 2;            // 8 KB: 64 rows x 64 k of a K-major SWIZZLE_128B operand
  static constexpr int kWBlk = 128 * BK * 2;           // 16 KB: 128 weight rows x 64 k
  static constexpr int kA1Offset = 0;                  // resident Y tile of the unit: 8 k-blocks
  static constexpr int kHOffset = kA1Offset + MF_KB1 * kKBlk;   // GELU'd hidden chunk: 4 k-blocks
  static constexpr int kRingOffset = kHOffset + MF_KB2 * kKBlk;
  static constexpr int kStageBytes = 2 * kWBlk;        // MMA1: two k-blocks of this CTA's 128 W1 rows; MMA2: one k-block of both column halves of W2
  static constexpr int kStoreOffset = kRingOffset + MF_STAGES * kStageBytes;
  static constexpr int kBoxBytes = 32 * 64;            // 32 rows x 16 fp32, SWIZZLE_64B
  static constexpr int kBarOffset = kStoreOffset + MF_EPI_WARPS * kBoxBytes;
  static constexpr int kTotal = kBarOffset + 512 + 1024;
};
static_assert(SmemMf::kTotal <= 227 * 1024, "shared memory budget exceeded");
static_assert(SmemMf::kHOffset % 1024 == 0 && SmemMf::kRingOffset % 1024 == 0 && SmemMf::kStoreOffset % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");

struct MlpParams {
  CUtensorMap ta, tw1, tw2, tx;
  const float* b1;
  const float* b2;
  int M, n_units;
  int tune;  // tuning build only (timing experiments, results are garbage): bit 0 = MMA2 does not wait for the hidden chunk,
             // bit 1 = EPI1 skips the GELU arithmetic
};

__global__ void __launch_bounds__(MF_THREADS, 1) mlp_fused_2sm_kernel(const __grid_constant__ MlpParams P) {
  using L = SmemMf;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);  // weight ring, leader's copy is waited on
  uint64_t* empty_bar = full_bar + MF_STAGES;
  uint64_t* a_full = empty_bar + MF_STAGES;  // the unit's Y tile has landed (leader's copy)
  uint64_t* a_empty = a_full + 1;            // the unit's last MMA1 has read it
  uint64_t* hacc_full = a_empty + 1;         // [2] MMA1_j complete
  uint64_t* h_full = hacc_full + 2;          // EPI1_j complete in BOTH CTAs (leader's copy counts 2 x 16 warps)
  uint64_t* h_empty = h_full + 1;            // MMA2_j has read the hidden chunk
  uint64_t* o_full = h_empty + 1;            // the unit's last MMA2 complete
  uint64_t* o_empty = o_full + 1;            // EPI2 has drained Oacc in both CTAs (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&P.ta);
    prefetch_tmap(&P.tw1);
    prefetch_tmap(&P.tw2);
    prefetch_tmap(&P.tx);
#pragma unroll
    for (int s = 0; s < MF_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(&hacc_full[0], 1);
    mbar_init(&hacc_full[1], 1);
    mbar_init(h_full, 2 * MF_EPI_WARPS);
    mbar_init(h_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 2 * MF_EPI_WARPS);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {  // all 512 TMEM columns: Oacc [0, 256), Hacc[0] [256, 384), Hacc[1] [384, 512)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PDL_PROLOGUE();

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer (both CTAs): own 64 rows of Y once per unit; own 128 rows of every weight box ----
      const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[0]), 0);
      const uint32_t a_full_leader = mapa_u32(smem_u32(a_full), 0);
      int s = 0;
      uint32_t ph = 0;
      auto stage = [&](const CUtensorMap* map, int k0, int r0, int k1, int r1) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (leader) mbar_arrive_expect_tx(&full_bar[s], 2u * L::kStageBytes);
        uint8_t* dst = smem + L::kRingOffset + s * L::kStageBytes;
        const uint32_t bar = full_leader + 8u * static_cast<uint32_t>(s);
        tma_load_2d_2sm(dst, map, bar, k0, r0);
        tma_load_2d_2sm(dst + L::kWBlk, map, bar, k1, r1);
        if (++s == MF_STAGES) { s = 0; ph ^= 1; }
      };
      auto load_w1 = [&](int j) {  // chunk j of W1: this CTA's 128 of its 256 rows, two k-blocks per stage
        const int r = j * MF_CHUNK + static_cast<int>(crank) * 128;
        for (int st = 0; st < MF_KB1 / 2; ++st) stage(&P.tw1, (2 * st) * BK, r, (2 * st + 1) * BK, r);
      };
      auto load_w2 = [&](int j) {  // k-slice j of W2: this CTA's 128 rows of each 256-row column half, one k-block per stage
        for (int kb = 0; kb < MF_KB2; ++kb) {
          const int k = j * MF_CHUNK + kb * BK;
          stage(&P.tw2, k, static_cast<int>(crank) * 128, k, 256 + static_cast<int>(crank) * 128);
        }
      };
      int it = 0;
      for (int u = pair; u < P.n_units; u += n_pairs, ++it) {
        const int m0 = (u * 2 + static_cast<int>(crank)) * 64;
        mbar_wait(a_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u);
        if (leader) mbar_arrive_expect_tx(a_full, 2u * MF_KB1 * L::kKBlk);
        for (int kb = 0; kb < MF_KB1; ++kb) tma_load_2d_2sm(smem + L::kA1Offset + kb * L::kKBlk, &P.ta, a_full_leader, kb * BK, m0);
        load_w1(0);
        for (int j = 0; j < MF_NCHUNK; ++j) {  // the order the MMA warp consumes the ring in
          if (j + 1 < MF_NCHUNK) load_w1(j + 1);
          load_w2(j);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ---- MMA issuer (leader only) ----
      constexpr uint32_t idesc = make_idesc_mn(128, 256);
      const uint32_t a1_addr = smem_u32(smem + L::kA1Offset), h_addr = smem_u32(smem + L::kHOffset), ring_addr = smem_u32(smem + L::kRingOffset);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int u = pair; u < P.n_units; u += n_pairs, ++it) {
        auto mma1 = [&](int j) {
          const uint32_t d = tmem_base + 256u + static_cast<uint32_t>(j & 1) * 128u;
          for (int st = 0; st < MF_KB1 / 2; ++st) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t w = ring_addr + s * L::kStageBytes;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int kb = 2 * st + kk;
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_bf16_2sm(d, make_smem_desc(a1_addr + kb * L::kKBlk + k * UMMA_K * 2), make_smem_desc(w + kk * L::kWBlk + k * UMMA_K * 2), idesc,
                              (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[s]);
            if (++s == MF_STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit_2sm(&hacc_full[j & 1]);
          if (j == MF_NCHUNK - 1) umma_commit_2sm(a_empty);  // the Y tile may be overwritten by the next unit's
        };
        auto mma2 = [&](int j) {
#ifdef M3PC_TUNING
          if (!(P.tune & 1))
#endif
          mbar_wait_cluster(h_full, static_cast<uint32_t>(it * MF_NCHUNK + j) & 1u);  // both CTAs' epilogue warps have written H_j
          tc_fence_after();
          if (j == 0) {
            mbar_wait(o_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u);  // the previous unit's Oacc has been drained
            tc_fence_after();
          }
          for (int kb = 0; kb < MF_KB2; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t w = ring_addr + s * L::kStageBytes;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint32_t acc = (j | kb | k) != 0 ? 1u : 0u;
              const uint64_t da = make_smem_desc(h_addr + kb * L::kKBlk + k * UMMA_K * 2);
              umma_bf16_2sm(tmem_base, da, make_smem_desc(w + k * UMMA_K * 2), idesc, acc);
              umma_bf16_2sm(tmem_base + 128u, da, make_smem_desc(w + L::kWBlk + k * UMMA_K * 2), idesc, acc);
            }
            umma_commit_2sm(&empty_bar[s]);
            if (++s == MF_STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit_2sm(h_empty);
          if (j == MF_NCHUNK - 1) umma_commit_2sm(o_full);
        };
        mbar_wait(a_full, static_cast<uint32_t>(it) & 1u);
        tc_fence_after();
        mma1(0);
        for (int j = 0; j < MF_NCHUNK; ++j) {
          if (j + 1 < MF_NCHUNK) mma1(j + 1);  // the next hidden chunk is computed while the epilogue warps convert this one
          mma2(j);
        }
      }
    }
  } else {
    // ---- epilogue warps (both CTAs) ----
    const int ew = warp - 2;
    const int q = warp & 3;           // TMEM lanes [32 q, 32 q + 32)
    const int grp = ew >> 2;          // 0..3
    const int rowq = q & 1, lanehalf = q >> 1;
    const int r = rowq * 32 + lane;   // row of this CTA's 64
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t h_full_leader = mapa_u32(smem_u32(h_full), 0), o_empty_leader = mapa_u32(smem_u32(o_empty), 0);
    // EPI1: this warp converts TMEM columns [32 grp, +32) of its lanes = hidden columns lanehalf * 128 + 32 grp .. + 32 of the chunk
    const int hcol = lanehalf * 128 + grp * 32;
    uint8_t* hdst = smem + L::kHOffset + (hcol >> 6) * L::kKBlk + r * 128;  // k-block of the chunk, row r
    const uint32_t hchunk0 = static_cast<uint32_t>((hcol & 63) >> 3);        // first 16-byte chunk of the row segment
    const uint32_t hsw = static_cast<uint32_t>(r & 7);                        // SWIZZLE_128B: chunk c of row r lives at c ^ (r & 7)
    // EPI2: output columns of this warp (see gemm_ln.cu, ROWS = 128)
    const int colw = (grp >> 1) * 256 + lanehalf * 128 + (grp & 1) * 64;
    const uint32_t tcol = static_cast<uint32_t>((grp >> 1) * 128 + (grp & 1) * 64);
    uint8_t* sbuf = smem + L::kStoreOffset + ew * L::kBoxBytes;
    const uint32_t sw64 = static_cast<uint32_t>((lane >> 1) & 3);  // SWIZZLE_64B staging box
    int it = 0;
    for (int u = pair; u < P.n_units; u += n_pairs, ++it) {
#pragma unroll 1
      for (int j = 0; j < MF_NCHUNK; ++j) {
        const uint32_t b = static_cast<uint32_t>(j & 1);
        mbar_wait(&hacc_full[b], static_cast<uint32_t>(it * (MF_NCHUNK / 2) + (j >> 1)) & 1u);
        tc_fence_after();
        uint32_t acc[32];
        tmem_ld32(tmem_base + lane_addr + 256u + b * 128u + static_cast<uint32_t>(grp * 32), acc);
        tmem_ld_wait();
        const float* bp = P.b1 + j * MF_CHUNK + hcol;
        uint4 o[4];
#ifdef M3PC_TUNING
        if (P.tune & 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = make_uint4(acc[8 * i], acc[8 * i + 1], acc[8 * i + 2], acc[8 * i + 3]);
        } else
#endif
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(bp + 8 * i)), bb = __ldg(reinterpret_cast<const float4*>(bp + 8 * i + 4));
          const float v0 = gelu_erf_tanh(__uint_as_float(acc[8 * i + 0]) + ba.x), v1 = gelu_erf_tanh(__uint_as_float(acc[8 * i + 1]) + ba.y);
          const float v2 = gelu_erf_tanh(__uint_as_float(acc[8 * i + 2]) + ba.z), v3 = gelu_erf_tanh(__uint_as_float(acc[8 * i + 3]) + ba.w);
          const float v4 = gelu_erf_tanh(__uint_as_float(acc[8 * i + 4]) + bb.x), v5 = gelu_erf_tanh(__uint_as_float(acc[8 * i + 5]) + bb.y);
          const float v6 = gelu_erf_tanh(__uint_as_float(acc[8 * i + 6]) + bb.z), v7 = gelu_erf_tanh(__uint_as_float(acc[8 * i + 7]) + bb.w);
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v0, v1), p1 = __floats2bfloat162_rn(v2, v3);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v4, v5), p3 = __floats2bfloat162_rn(v6, v7);
          o[i].x = *reinterpret_cast<uint32_t*>(&p0); o[i].y = *reinterpret_cast<uint32_t*>(&p1);
          o[i].z = *reinterpret_cast<uint32_t*>(&p2); o[i].w = *reinterpret_cast<uint32_t*>(&p3);
        }
        // the hidden buffer is free once MMA2 of the previous chunk has read it (first chunk ever: passes immediately)
        mbar_wait(h_empty, (static_cast<uint32_t>(it * MF_NCHUNK + j) & 1u) ^ 1u);
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(hdst + (((hchunk0 + i) ^ hsw) << 4)) = o[i];
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(h_full_leader);  // release at cluster scope: the peer's H rows are read by the leader's MMAs
      }
      // ---- EPI2: Oacc + b2 -> X (TMA reduce-add into the fp32 residual stream) ----
      const int row0 = (u * 2 + static_cast<int>(crank)) * 64 + rowq * 32;
      mbar_wait(o_full, static_cast<uint32_t>(it) & 1u);
      tc_fence_after();
      if (row0 < P.M) {
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int col0 = colw + ci * 16;
          uint32_t rr[32];
          tmem_ld16(tmem_base + lane_addr + tcol + static_cast<uint32_t>(ci * 16), rr);
          if (lane == 0) bulk_wait_read<0>();  // the staging box's previous reduce has been read
          __syncwarp();
          tmem_ld_wait();
          if (ci == 3) {  // Oacc drained by this warp: the next unit's MMA2_0 may overwrite it
            tc_fence_before();
            if (lane == 0) mbar_arrive_remote(o_empty_leader);
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.b2 + col0 + 4 * jj));
            *reinterpret_cast<float4*>(sbuf + lane * 64 + ((static_cast<uint32_t>(jj) ^ sw64) << 4)) =
                make_float4(__uint_as_float(rr[4 * jj + 0]) + b4.x, __uint_as_float(rr[4 * jj + 1]) + b4.y, __uint_as_float(rr[4 * jj + 2]) + b4.z,
                            __uint_as_float(rr[4 * jj + 3]) + b4.w);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&P.tx, sbuf, col0, row0);
            bulk_commit();
          }
        }
      } else {
        tc_fence_before();
        if (lane == 0) mbar_arrive_remote(o_empty_leader);
      }
    }
    if (lane == 0) bulk_wait_read<0>();  // shared memory may be released once the stores have been read; the writes complete with the grid
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

// X (M, 512) fp32 += GELU(Y W1^T + b1) W2^T + b2 with Y (M, 512) bf16, W1 (2048, 512) bf16, W2 (512, 2048) bf16, b1 (2048), b2 (512) fp32
int mlp_fused_bf16(const __nv_bfloat16* Y, const __nv_bfloat16* W1, const float* b1, const __nv_bfloat16* W2, const float* b2, float* X, int M,
                   cudaStream_t st) {
  M3PC_REQUIRE(M > 0 && Y && W1 && b1 && W2 && b2 && X, "mlp_fused: bad argument");
  M3PC_REQUIRE(((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(W1) | reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(X) |
                 reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(b2)) & 15) == 0,
               "mlp_fused: operands must be 16-byte aligned");
  M3PC_TRY(gemm_init_driver_api());
  static PerDevice<bool> configured;
  if (!configured.here()) {
    M3PC_CHECK_CUDA(cudaFuncSetAttribute(mlp_fused_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemMf::kTotal));
    configured.here() = true;
  }
  MlpParams P{};
  M3PC_TRY(make_tmap(&P.ta, Y, static_cast<uint64_t>(M), MF_D, 64));
  M3PC_TRY(make_tmap(&P.tw1, W1, MF_F, MF_D, 128));
  M3PC_TRY(make_tmap(&P.tw2, W2, MF_D, MF_F, 128));
  M3PC_TRY(make_tmap_out(&P.tx, X, static_cast<uint64_t>(M), MF_D, true));
  P.b1 = b1;
  P.b2 = b2;
  P.M = M;
  P.n_units = ceil_div(M, 128);
  if (const char* t = tune_env("M3PC_TUNE_MLP")) P.tune = atoi(t);
  const int pairs = std::min(P.n_units, device_num_sms() / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(MF_THREADS);
  cfg.dynamicSmemBytes = SmemMf::kTotal;
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
  M3PC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mlp_fused_2sm_kernel, P));
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}

}  // namespace m3pc
