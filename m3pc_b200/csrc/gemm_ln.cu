// Tensor-core GEMM with a fused residual + LayerNorm epilogue (N = n_embd = 512):
//
//     X[M,512] (fp32)  <-  R + A[M,K] W[512,K]^T + bias          R = X (in place) or table[row / rows_per_group] (first block)
//     Y[M,512] (bf16)  <-  LayerNorm(X; gamma, beta)             eps 1e-5, statistics in fp32
//
// Reference call sites: the out-projection / linear2 of nn.TransformerEncoderLayer followed by the next pre-LN norm
// (mtm_model.py:379-409; norm_first: x = x + sa(norm1(x)); x = x + ff(norm2(x))).  The unfused path writes X through a TMA
// reduce-add (read-modify-write of 4 KB per row) and a separate LayerNorm kernel reads it again; here a CTA pair owns 256
// rows x ALL 512 columns -- exactly the pair's tensor memory (2 x 256 fp32 columns per CTA) -- so the epilogue sees whole
// rows: it loads the residual tile through the TMA, adds, stores X, keeps the updated rows in TMEM (tcgen05.st), exchanges
// the row statistics between the four column-slice warps of a row, and writes the normalised bf16 operand of the next GEMM.
// 6 KB of HBM traffic per row and one launch instead of 8 KB and two.  The accumulator is single-buffered (TMEM is full),
// so the epilogue of a row block is not overlapped with the MMAs of the next one; the operand ring keeps prefetching.
//
// Structure (per CTA of the pair, 18 warps): warp 0 lane 0 = TMA producer (A: own 128 rows; W: own 128 rows of each 256-row
// half), warp 1 lane 0 of the leader = MMA issuer (two tcgen05.mma cta_group::2 M256 x N256 x K16 per k-step, one per
// column half), warps 2..17 = epilogue (TMEM lane quarter = warp % 4, column slice of 128 = (warp - 2) / 4).
#include "common.cuh"
#include "tcgen05.cuh"

namespace m3pc {

int make_tmap(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
int make_tmap_out(CUtensorMap* map, void* ptr, uint64_t rows, uint64_t cols, bool f32);

namespace {

constexpr int LN_N = 512;
constexpr int LN_EPI_WARPS = 16;
constexpr int LN_THREADS = 32 * (2 + LN_EPI_WARPS);
constexpr int LN_STAGES = 3;

struct SmemLn {
  static constexpr int kABlk = BM * BK * 2;        // 16 KB: this CTA's 128 rows of one A k-block
  static constexpr int kBBlk = 128 * BK * 2;       // 16 KB: this CTA's 128 rows of one 256-row half of W
  static constexpr int kStageBytes = kABlk + 2 * kBBlk;
  static constexpr int kBoxBytes = 32 * 64;        // one staged chunk: 32 rows x 64 bytes (16 fp32 / 32 bf16 columns), SWIZZLE_64B
  static constexpr int kStoreOffset = LN_STAGES * kStageBytes;
  static constexpr int kConstOffset = kStoreOffset + LN_EPI_WARPS * 2 * kBoxBytes;  // bias | gamma | beta, 512 floats each
  static constexpr int kStatsOffset = kConstOffset + 3 * LN_N * 4;                  // float2 [128 rows][4 slices]
  static constexpr int kBarOffset = kStatsOffset + 128 * 4 * 8;
  static constexpr int kTotal = kBarOffset + 512 + 1024;
};

struct LnGemmParams {
  CUtensorMap ta, tw, tx, ty;
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* table;  // null: residual = X
  int rows_per_group;
  int M, num_kb, n_units;
  int tune;  // tuning build only (timing experiments, results are garbage): bit 0 = no residual loads, bit 1 = no X store,
             // bit 2 = no Y store (pass 2 only drains TMEM), bit 3 = epilogue only hands the accumulator back
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__global__ void __launch_bounds__(LN_THREADS, 1) gemm_ln_2sm_kernel(const __grid_constant__ LnGemmParams P) {
  using L = SmemLn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);  // leader's copy is the one waited on
  uint64_t* empty_bar = full_bar + LN_STAGES;
  uint64_t* acc_full = empty_bar + LN_STAGES;
  uint64_t* acc_empty = acc_full + 1;                                      // leader's copy counts both CTAs' epilogue warps
  uint64_t* res_bar = acc_empty + 1;                                       // [LN_EPI_WARPS][2]: residual box landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * LN_EPI_WARPS);
  float* sconst = reinterpret_cast<float*>(smem + L::kConstOffset);
  float2* sstats = reinterpret_cast<float2*>(smem + L::kStatsOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&P.ta);
    prefetch_tmap(&P.tw);
    prefetch_tmap(&P.tx);
    prefetch_tmap(&P.ty);
#pragma unroll
    for (int s = 0; s < LN_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 2 * LN_EPI_WARPS);
    for (int i = 0; i < 2 * LN_EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {  // all 512 TMEM columns: one 128-lane x 512-column fp32 accumulator per CTA
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(LN_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // epilogue constants are weights (never written by a preceding kernel): staged before the programmatic dependency wait
  for (int i = threadIdx.x; i < LN_N; i += LN_THREADS) {
    sconst[i] = P.bias != nullptr ? __ldg(P.bias + i) : 0.f;
    sconst[LN_N + i] = __ldg(P.gamma + i);
    sconst[2 * LN_N + i] = __ldg(P.beta + i);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PDL_PROLOGUE();

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer (both CTAs) ----
      const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[0]), 0);
      int s = 0;
      uint32_t ph = 0;
      for (int u = pair; u < P.n_units; u += n_pairs) {
        const int m0 = (u * 2 + static_cast<int>(crank)) * BM;
        for (int kb = 0; kb < P.num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2u * L::kStageBytes);
          uint8_t* dst = smem + s * L::kStageBytes;
          const uint32_t bar = full_leader + 8u * static_cast<uint32_t>(s);
          tma_load_2d_2sm(dst, &P.ta, bar, kb * BK, m0);
          tma_load_2d_2sm(dst + L::kABlk, &P.tw, bar, kb * BK, static_cast<int>(crank) * 128);
          tma_load_2d_2sm(dst + L::kABlk + L::kBBlk, &P.tw, bar, kb * BK, 256 + static_cast<int>(crank) * 128);
          if (++s == LN_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ---- MMA issuer (leader only): both column halves per k-step ----
      constexpr uint32_t idesc = make_idesc_mn(2 * BM, 256);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int u = pair; u < P.n_units; u += n_pairs, ++it) {
        mbar_wait(acc_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < P.num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
          const uint32_t b0_addr = a_addr + L::kABlk, b1_addr = b0_addr + L::kBBlk;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
            const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 2);
            umma_bf16_2sm(tmem_base, da, make_smem_desc(b0_addr + k * UMMA_K * 2), idesc, acc);
            umma_bf16_2sm(tmem_base + 256u, da, make_smem_desc(b1_addr + k * UMMA_K * 2), idesc, acc);
          }
          umma_commit_2sm(&empty_bar[s]);
          if (++s == LN_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_2sm(acc_full);
      }
    }
  } else {
    // ---- epilogue warps (both CTAs): 32 rows (TMEM lane quarter) x 128 columns each ----
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int cgrp = ew >> 2;
    uint8_t* sbuf = smem + L::kStoreOffset + ew * 2 * L::kBoxBytes;
    uint64_t* rbar = res_bar + 2 * ew;
    uint32_t rph[2] = {0u, 0u};
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(acc_empty), 0);
    const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);  // SWIZZLE_64B: 16-byte chunk j of row r lives at j ^ ((r >> 1) & 3)
#ifdef M3PC_TUNING
    const int tune = P.tune;
#else
    constexpr int tune = 0;
#endif
    const bool from_x = P.table == nullptr && !(tune & 1);
    const float* sbias = sconst;
    const float* sgamma = sconst + LN_N;
    const float* sbeta = sconst + 2 * LN_N;
    int it = 0;
    uint32_t nbox = 0;  // boxes handed to the TMA store so far (selects the staging buffer)
    for (int u = pair; u < P.n_units; u += n_pairs, ++it) {
      const int row0 = (u * 2 + static_cast<int>(crank)) * BM + quarter * 32;
      const int row = row0 + lane;
      const bool live = row0 < P.M && !(tune & 8);
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(cgrp * 128);
      // short K: the residual tile is pulled into L2 while the MMAs of this row block are still running (each box load then waits
      // for an L2 hit instead of an HBM round trip).  Not for long K: the A stream evicts the prefetched lines before they are
      // used and the tile is read from HBM twice (measured: 838 MB instead of 654 MB per launch at K = 2048).
      if (from_x && live && lane < 8 && P.num_kb <= 8) tma_prefetch_l2_2d(&P.tx, cgrp * 128 + lane * 16, row0);
      // the box of the first chunk travels all the way to shared memory before the accumulator is ready
      if (from_x && live && lane == 0) {
        bulk_wait_read<0>();  // the buffer's last store has been read
        mbar_arrive_expect_tx(&rbar[nbox & 1], L::kBoxBytes);
        tma_load_2d(sbuf + (nbox & 1) * L::kBoxBytes, &P.tx, &rbar[nbox & 1], cgrp * 128, row0);
      }
      mbar_wait(acc_full, static_cast<uint32_t>(it) & 1u);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
      if (live) {
        // ---- pass 1: x = acc + bias + residual -> X (TMA store) and back into TMEM; row sums ----
#pragma unroll 1
        for (int ci = 0; ci < 8; ++ci) {
          const int col0 = cgrp * 128 + ci * 16;
          const uint32_t b = nbox & 1;
          uint8_t* buf = sbuf + b * L::kBoxBytes;
          if (from_x && ci + 1 < 8 && lane == 0) {  // prefetch the next residual box into the other buffer
            bulk_wait_read<0>();
            mbar_arrive_expect_tx(&rbar[b ^ 1], L::kBoxBytes);
            tma_load_2d(sbuf + (b ^ 1) * L::kBoxBytes, &P.tx, &rbar[b ^ 1], col0 + 16, row0);
          }
          __syncwarp();
          uint32_t r[32];
          tmem_ld16(tacc + static_cast<uint32_t>(ci * 16), r);
          if (from_x) {
            mbar_wait(&rbar[b], rph[b]);
            rph[b] ^= 1u;
          } else {
            if (lane == 0) bulk_wait_read<1>();  // the store issued from this buffer two boxes ago has been read
            __syncwarp();
          }
          tmem_ld_wait();
          float v[16];
          const float* trow = (from_x || P.table == nullptr) ? nullptr : P.table + static_cast<size_t>(min(row, P.M - 1) / P.rows_per_group) * LN_N + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4* slot = reinterpret_cast<float4*>(buf + lane * 64 + ((static_cast<uint32_t>(j) ^ sw) << 4));
            const float4 res = from_x ? *slot : (trow != nullptr ? __ldg(reinterpret_cast<const float4*>(trow + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f));
            const float4 b4 = *reinterpret_cast<const float4*>(sbias + col0 + 4 * j);
            v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b4.x + res.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y + res.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z + res.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w + res.w;
            *slot = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s1 += v[j];
            s2 = fmaf(v[j], v[j], s2);
          }
          tmem_st16(tacc + static_cast<uint32_t>(ci * 16), v);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (!(tune & 2)) tma_store_2d(&P.tx, buf, col0, row0);
            bulk_commit();
          }
          ++nbox;
        }
        tmem_st_wait();
      }
      // ---- row statistics across the four column-slice warps of this lane quarter ----
      sstats[(quarter * 32 + lane) * 4 + cgrp] = make_float2(s1, s2);
      named_bar_sync(1 + quarter, 128);
      float mean, rstd;
      {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float2 p = sstats[(quarter * 32 + lane) * 4 + g];
          t1 += p.x;
          t2 += p.y;
        }
        mean = t1 * (1.0f / LN_N);
        const float var = fmaxf(t2 * (1.0f / LN_N) - mean * mean, 0.f);
        rstd = rsqrtf(var + 1e-5f);
      }
      named_bar_sync(1 + quarter, 128);  // everybody has read the statistics: the next row block may overwrite them
      if (live) {
        // ---- pass 2: y = (x - mean) * rstd * gamma + beta -> bf16 -> Y ----
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int col0 = cgrp * 128 + ci * 32;
          uint8_t* buf = sbuf + (nbox & 1) * L::kBoxBytes;
          uint32_t r[32];
          tmem_ld32(tacc + static_cast<uint32_t>(ci * 32), r);
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          tmem_ld_wait();
          if (ci == 3) {  // the accumulator has been drained: the MMA warp may start the next row block
            tc_fence_before();
            if (lane == 0) mbar_arrive_remote(acc_empty_leader);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float y[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = col0 + 8 * j + q;
              y[q] = (__uint_as_float(r[8 * j + q]) - mean) * rstd * sgamma[c] + sbeta[c];
            }
            __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0], y[1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(y[2], y[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(y[4], y[5]);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(y[6], y[7]);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
            o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(buf + lane * 64 + ((static_cast<uint32_t>(j) ^ sw) << 4)) = o;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (!(tune & 4)) tma_store_2d(&P.ty, buf, col0, row0);
            bulk_commit();
          }
          ++nbox;
        }
      } else {
        tc_fence_before();
        if (lane == 0) mbar_arrive_remote(acc_empty_leader);
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(LN_N) : "memory");
  }
}

}  // namespace

// X (M, 512) fp32 in place (or table rows as the residual), Y (M, 512) bf16 = LayerNorm(X).  K % 64 == 0.
int gemm_ln_bf16(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, float* X, __nv_bfloat16* Y, const float* gamma,
                 const float* beta, const float* table, int rows_per_group, int M, int K, cudaStream_t st) {
  M3PC_REQUIRE(M > 0 && K > 0 && K % BK == 0, "gemm_ln: K must be a positive multiple of 64");
  M3PC_REQUIRE(A && W && X && Y && gamma && beta, "gemm_ln: null operand");
  M3PC_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0,
               "gemm_ln: operands must be 16-byte aligned");
  M3PC_TRY(gemm_init_driver_api());
  static PerDevice<bool> configured;
  if (!configured.here()) {
    static_assert(SmemLn::kTotal <= 227 * 1024, "shared memory budget exceeded");
    M3PC_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLn::kTotal));
    configured.here() = true;
  }
  LnGemmParams P{};
  M3PC_TRY(make_tmap(&P.ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), BM));
  M3PC_TRY(make_tmap(&P.tw, W, static_cast<uint64_t>(LN_N), static_cast<uint64_t>(K), 128));
  M3PC_TRY(make_tmap_out(&P.tx, X, static_cast<uint64_t>(M), LN_N, true));
  M3PC_TRY(make_tmap_out(&P.ty, Y, static_cast<uint64_t>(M), LN_N, false));
  P.bias = bias; P.gamma = gamma; P.beta = beta; P.table = table;
  P.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  P.M = M;
  P.num_kb = K / BK;
  P.n_units = ceil_div(M, 2 * BM);
  if (const char* t = tune_env("M3PC_TUNE_LN")) P.tune = atoi(t);
  const int pairs = std::min(P.n_units, device_num_sms() / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(LN_THREADS);
  cfg.dynamicSmemBytes = SmemLn::kTotal;
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
  M3PC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln_2sm_kernel, P));
  M3PC_CHECK_LAUNCH();
  return M3PC_OK;
}

}  // namespace m3pc
