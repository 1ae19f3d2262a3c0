// Tensor-mode GEMM for sm_100a: bf16 operands staged by TMA (128B swizzle) into a
// multi-stage shared-memory ring, tcgen05.mma (cta_group::1, 128 x BN x 16) issued
// by one thread with fp32 accumulators in TMEM, and a fused epilogue
// (bias / ReLU / ReLU-mask / dropout / residual) whose results leave the SM through
// swizzled shared-memory staging tiles and TMA stores -- or TMA fp32 reduce-adds for
// split-K / accumulate -- so the output is written in full 128-byte lines, clipped at the
// matrix edge by the tensor map, with no per-thread global stores.  Persistent: one CTA
// per SM walks the tile list; TMEM holds two accumulators so the epilogue of tile i
// overlaps the mainloop of tile i+1.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer, warps 2..9 = epilogue (TMEM lane quadrant = warp_id % 4, two warps per quadrant).
//
// Operand layouts: K-major ([rows, K], K contiguous: activations, nn.Linear
// weights) and MN-major ([K, rows], rows contiguous: the transposed views that
// dgrad / wgrad need) are both fed straight from the row-major tensors -- no
// transposed copies are ever written to HBM.
#include <cuda.h>
#include <stdlib.h>

#include "epilogue.cuh"
#include "ptx.cuh"

namespace dgpt {

using namespace ptx;

static constexpr int TBM = 128;       // tile M (UMMA M)
static constexpr int TBK = 64;        // k-block: 64 bf16 = one 128-byte swizzle row
static constexpr int UMMA_K = 16;
static constexpr int kThreads = 64 + 256;  // TMA warp, MMA warp, 8 epilogue warps
static constexpr int kStageBudget = 160 * 1024;
static constexpr int kStagingBytes = 8 * 2 * 4096;  // 8 epilogue warps x 2 buffers x (32 rows x 128 B)

enum StoreMode { kStoreDirect = 0, kStoreTma = 1, kStoreTmaAdd = 2 };

struct TcParams {
  int M, N, K;
  int m_tiles, n_tiles, split_k, kb_total, kb_per_split;
  int vec_ok;      // epilogue operands (bias / residual / relu_aux) allow 16-byte loads
  int store_mode;  // StoreMode for D
  int cta_group;   // 1, or 2 = CTA pairs (cluster of 2) sharing each MMA
#ifdef DGPT_GEMM_TS
  long long* ts;   // cycle stamps of (block 0, first epilogue warp, lane 0) for the first tiles (debug builds)
#endif
  int debug;       // DGPT_GEMM_DEBUG: 1 = epilogue skipped, 2 = no TMA loads / MMAs (timing experiments only)
  Epilogue ep;
};

// CG = CTAs sharing one MMA (tcgen05 cta_group).  With CG = 2 a pair of CTAs computes a 256 x BN tile: each CTA
// stages its own 128 rows of A but only HALF of B (the MMA reads the other half from the peer's shared
// memory), so every SM ingests and re-reads a third (BN = 256) or a quarter (BN = 128) fewer operand
// bytes per MMA cycle -- operand delivery (L2 -> SM ~43 B/cycle/SM measured) and shared-memory bandwidth
// are what bound the single-CTA mainloop.
template <int BN, int CG = 1>
struct TcCfg {
  static constexpr int kBBytes = (BN / CG) * TBK * 2;
  static constexpr int kStageBytes = TBM * TBK * 2 + kBBytes;
  static constexpr int kStages = kStageBudget / kStageBytes > 8 ? 8 : kStageBudget / kStageBytes;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // 128 / 256 / 512 (powers of two)
  static constexpr int kBiasBytes = 2 * BN * 4;  // bias slice of the tile, double-buffered with the accumulator
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kStagingBytes + kBiasBytes + 256 /*barriers*/;
};

// --------------------------------------------------------------------------
// epilogue math on 32 consecutive columns [n, n+32) of accumulator row m (all in registers).
// Fast path only: n + 32 <= N, 16-byte aligned operands.  Keep this small -- the kernel's
// instruction footprint must stay inside the instruction cache (an earlier version that
// unrolled the ragged path 32x was 180 KB of SASS and stalled on instruction fetch).
// --------------------------------------------------------------------------
// `pre` holds the chunk's residual (8 x float4) or ReLU-mask operand (4 x uint4 of bf16 / 8 x float4 of
// fp32), fetched one chunk ahead; `bias_s` is the tile's bias slice in shared memory.  (With ~226 KB
// of shared memory in use the L1 data cache is nearly gone: an un-prefetched global load here costs a
// full L2 round trip per chunk on the only warp of the scheduler, which is what bounded the epilogue.)
// EPI: compile-time feature mask of the fused epilogue (kEpiBias | kEpiRelu | ...), or -1 to test the
// run-time flags.  The specialisations matter: with run-time flags the compiler if-converts the whole
// body (~600 predicated instructions per 32-column step, measured 700 cycles) instead of ~100.
enum { kEpiBias = 1, kEpiRelu = 2, kEpiAux = 4, kEpiDrop = 8, kEpiRes = 16 };
template <int EPI> __device__ __forceinline__ bool epi_has(int bit, bool runtime) { return EPI < 0 ? runtime : (EPI & bit) != 0; }

template <int EPI>
__device__ __forceinline__ void epilogue_prefetch(const Epilogue& e, int m, int n, uint4 (&pre)[8]) {
  if (e.first_split && epi_has<EPI>(kEpiRes, e.residual != nullptr)) {
    const uint4* rp = reinterpret_cast<const uint4*>(e.residual + (int64_t)m * e.ldr + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) pre[j] = __ldg(rp + j);
  } else if (epi_has<EPI>(kEpiAux, e.relu_aux != nullptr)) {
    if (e.aux_dtype == DGPT_BF16) {
      const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(e.relu_aux) + (int64_t)m * e.ld_aux + n);
#pragma unroll
      for (int j = 0; j < 4; ++j) pre[j] = __ldg(ap + j);
    } else {
      const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(e.relu_aux) + (int64_t)m * e.ld_aux + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) pre[j] = __ldg(ap + j);
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_math32(const Epilogue& e, int m, int n, float (&v)[32], const float* bias_s,
                                                const uint4 (&pre)[8]) {
  if (e.first_split && epi_has<EPI>(kEpiBias, e.bias != nullptr)) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias_s + j);  // shared-memory broadcast
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (epi_has<EPI>(kEpiRelu, e.relu != 0)) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (epi_has<EPI>(kEpiAux, e.relu_aux != nullptr)) {
    if (e.aux_dtype == DGPT_BF16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {pre[j].x, pre[j].y, pre[j].z, pre[j].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          // bf16 > 0  <=>  sign bit clear and magnitude non-zero
          const uint32_t lo = w[t] & 0xFFFFu, hi = w[t] >> 16;
          if (!(lo != 0 && lo < 0x8000u)) v[j * 8 + t * 2] = 0.f;
          if (!(hi != 0 && hi < 0x8000u)) v[j * 8 + t * 2 + 1] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!(__uint_as_float(pre[j].x) > 0.f)) v[4 * j] = 0.f;
        if (!(__uint_as_float(pre[j].y) > 0.f)) v[4 * j + 1] = 0.f;
        if (!(__uint_as_float(pre[j].z) > 0.f)) v[4 * j + 2] = 0.f;
        if (!(__uint_as_float(pre[j].w) > 0.f)) v[4 * j + 3] = 0.f;
      }
    }
  }
  if (epi_has<EPI>(kEpiDrop, e.thr != 0)) {
    const uint64_t q0 = ((uint64_t)m * (uint64_t)e.N + (uint64_t)n) >> 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const u32x4 b = dropout_bits4(e.seed, e.site, q0 + j);
      v[4 * j] = b.x >= e.thr ? v[4 * j] * e.inv_keep : 0.f;
      v[4 * j + 1] = b.y >= e.thr ? v[4 * j + 1] * e.inv_keep : 0.f;
      v[4 * j + 2] = b.z >= e.thr ? v[4 * j + 2] * e.inv_keep : 0.f;
      v[4 * j + 3] = b.w >= e.thr ? v[4 * j + 3] * e.inv_keep : 0.f;
    }
  }
  if (e.first_split && epi_has<EPI>(kEpiRes, e.residual != nullptr)) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] += __uint_as_float(pre[j].x); v[4 * j + 1] += __uint_as_float(pre[j].y);
      v[4 * j + 2] += __uint_as_float(pre[j].z); v[4 * j + 3] += __uint_as_float(pre[j].w);
    }
  }
}

// Ragged edge (n + 32 > N), unaligned operands, outputs the TMA cannot address, and the optional
// second output: one compact element-wise loop over a local copy of the 32 values.
// (Epilogue BY VALUE: a reference would pin the caller's copy in local memory, turning every flag test of
// the fast path into a local-memory load.)
__device__ __noinline__ void epilogue_slow32(const Epilogue e, float* v, int m, int n, int store_main) {
  if (m >= e.M) return;
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    if (n + j >= e.N) break;
    const float x = epilogue_value(e, m, n + j, v[j]);
    v[j] = x;
    if (store_main) {
      const int64_t i = (int64_t)m * e.ldd + n + j;
      if (e.d_dtype == DGPT_F32) {
        float* d = reinterpret_cast<float*>(e.D);
        if (e.atomic) atomicAdd(d + i, x);
        else d[i] = e.accumulate ? d[i] + x : x;
      } else {
        reinterpret_cast<__nv_bfloat16*>(e.D)[i] = __float2bfloat16_rn(x);
      }
    }
    if (e.D2) {
      const int64_t i = (int64_t)m * e.ldd2 + n + j;
      if (e.d2_dtype == DGPT_F32) reinterpret_cast<float*>(e.D2)[i] = x;
      else reinterpret_cast<__nv_bfloat16*>(e.D2)[i] = __float2bfloat16_rn(x);
    }
  }
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// one row (128 bytes) of a 32-row SWIZZLE_128B staging tile
__device__ __forceinline__ void stage_row_f32(uint8_t* tile, int row, const float (&v)[32]) {
  uint8_t* rp = tile + row * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(rp + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void stage_half_row_bf16(uint8_t* tile, int row, int half, const float (&v)[32]) {
  uint8_t* rp = tile + row * 128;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 w;
    w.x = pack2(v[8 * j], v[8 * j + 1]); w.y = pack2(v[8 * j + 2], v[8 * j + 3]);
    w.z = pack2(v[8 * j + 4], v[8 * j + 5]); w.w = pack2(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(rp + (((half * 4 + j) ^ (row & 7)) << 4)) = w;
  }
}

// --------------------------------------------------------------------------
// the kernel
// --------------------------------------------------------------------------
template <int BN, int A_MN, int B_MN, int EPI, int CG>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_d, TcParams p) {
  using Cfg = TcCfg<BN, CG>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kABytes = TBM * TBK * 2;
  constexpr int kBBytes = Cfg::kBBytes;
  constexpr uint32_t kIdesc = make_idesc_bf16(TBM * CG, BN, A_MN, B_MN);

  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + (size_t)kStages * Cfg::kStageBytes;  // 1024-aligned: stage sizes are multiples of 8 KB
  float* bias_s = reinterpret_cast<float*>(staging + kStagingBytes);  // [2][BN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStagingBytes + Cfg::kBiasBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef DGPT_GEMM_TS
  const long long t_block0 = clock64();
#endif

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    if (p.store_mode != kStoreDirect) prefetch_tensormap(&map_d);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8 * CG);  // one arrive per epilogue warp of every CTA sharing the accumulator
    }
    fence_barrier_init();
  }
  if (CG == 2) {  // both CTAs of the pair have initialised their barriers before the paired TMEM allocation
    __syncthreads();
    cluster_sync_all();
  }
  if (warp == 1) {
    if (CG == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer's TMEM and barriers exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile walk: with CG = 2 the pair handles pair-tiles (256 rows) and CTA rank r takes rows [128 r, 128 r + 128)
  const int crank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int tiles_mn = (p.m_tiles / CG) * p.n_tiles;
  const int total_tiles = tiles_mn * p.split_k;
  const int t_first = blockIdx.x / CG, t_step = gridDim.x / CG;

  if (warp == 0) {
    // ------------------------------ TMA producer ---------------------------
    // the whole warp walks the loop (uniform control flow); one elected lane issues
    {
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        const int ks = t / tiles_mn, mn = t - ks * tiles_mn;
        const int m0 = ((mn / p.n_tiles) * CG + crank) * TBM, n0 = (mn % p.n_tiles) * BN;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1 && !(p.debug & 2); ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = stage_base + (size_t)s * Cfg::kStageBytes;
          uint8_t* sb = sa + kABytes;
          const int k0 = kb * TBK;
          if (!elect_one()) {
            if (++s == kStages) { s = 0; ph ^= 1; }
            continue;
          }
          if (CG == 2) {
            // both CTAs load into their own shared memory; all bytes are credited to the LEADER's barrier,
            // which the leader arms for the pair's total
            if (crank == 0) mbar_expect_tx(&full_bar[s], 2 * (kABytes + kBBytes));
            const uint32_t lbar = mapa_u32(&full_bar[s], 0);
            if (A_MN) {
#pragma unroll
              for (int c = 0; c < TBM / 64; ++c) tma_load_2d_2sm(sa + c * 8192, &map_a, lbar, m0 + c * 64, k0);
            } else {
              tma_load_2d_2sm(sa, &map_a, lbar, k0, m0);
            }
            const int nh = n0 + crank * (BN / 2);  // this CTA's half of the B tile
            if (B_MN) {
#pragma unroll
              for (int c = 0; c < BN / 128; ++c) tma_load_2d_2sm(sb + c * 8192, &map_b, lbar, nh + c * 64, k0);
            } else {
              tma_load_2d_2sm(sb, &map_b, lbar, k0, nh);
            }
            if (++s == kStages) { s = 0; ph ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[s], kABytes + kBBytes);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < TBM / 64; ++c) tma_load_2d(sa + c * 8192, &map_a, &full_bar[s], m0 + c * 64, k0);
          } else {
            tma_load_2d(sa, &map_a, &full_bar[s], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &map_b, &full_bar[s], n0 + c * 64, k0);
          } else {
            tma_load_2d(sb, &map_b, &full_bar[s], k0, n0);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer -----------------------------
    if (crank == 0) {  // with CG = 2 only the pair's leader issues MMAs; whole warp loops, one elected lane issues
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        const int ks = t / tiles_mn;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1 && !(p.debug & 2); ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)s * Cfg::kStageBytes);
          const uint32_t sb = sa + kABytes;
          // K-major: 32 bytes per UMMA_K inside the 128B swizzle row, 8-row groups 1024 B apart.
          // MN-major: 16 k-rows (2 KB) per UMMA_K, 64-element MN chunks 8 KB apart.
          // (the start-address field counts 16-byte units: + 2 / + 128 per UMMA_K step)
          const uint64_t da0 = A_MN ? make_smem_desc_sw128(sa, 8192, 1024) : make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db0 = B_MN ? make_smem_desc_sw128(sb, 8192, 1024) : make_smem_desc_sw128(sb, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k) {
              const uint64_t da = da0 + (uint64_t)(k * (A_MN ? 128 : 2));
              const uint64_t db = db0 + (uint64_t)(k * (B_MN ? 128 : 2));
              if (CG == 1) tc_mma_bf16(d_tmem, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else tc_mma_bf16_2sm(d_tmem, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if (CG == 1) tc_commit(&empty_bar[s]);  // smem stage is free once these MMAs retire
            else tc_commit_2sm(&empty_bar[s]);      // ... in both CTAs of the pair
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) {
          if (CG == 1) tc_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
          else tc_commit_2sm(&tmem_full[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue -------------------------------
    // Eight warps: TMEM lane quadrant = warp % 4 (hardware rule), and the two warps of a quadrant
    // split the tile's columns.  Two warps per scheduler hide each other's fixed latencies
    // (measured per 32-column step of one warp: tcgen05.ld+wait ~590 cycles, bias/ReLU ~400,
    // pack+STS ~190, fence+TMA store ~260); the TMEM load of step i+1 is issued before step i's math.
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lanes [32*quad, 32*quad+32)
    const int half = ew >> 2;    // columns [half*BN/2, (half+1)*BN/2)
    constexpr int kCols = BN / 2;
    uint8_t* my_stage = staging + ew * 8192;
    int sbuf = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    Epilogue ep = p.ep;
    epilogue_resolve_seed(ep);
    const bool bf16_out = ep.d_dtype == DGPT_BF16;
    const int mode = p.store_mode;
    for (int t = t_first; t < total_tiles; t += t_step) {
      const int ks = t / tiles_mn, mn = t - ks * tiles_mn;
      const int m0 = ((mn / p.n_tiles) * CG + crank) * TBM, n0 = (mn % p.n_tiles) * BN;
      ep.first_split = (ks == 0);
      const int mrow0 = (p.debug & 1) ? p.M : m0 + quad * 32;  // debug bit 0: skip all epilogue work
      const int m = mrow0 + lane;
      const bool row_ok = m < p.M;
      // stage this tile's bias slice (each warp an eighth) while the accumulator is still being produced
      float* bias_t = bias_s + acc * BN;
      if (p.vec_ok && ep.first_split && epi_has<EPI>(kEpiBias, ep.bias != nullptr)) {
        const int col = ew * (BN / 8) + lane * 4;
        if (lane * 4 < BN / 8 && n0 + col < p.N)
          *reinterpret_cast<float4*>(bias_t + col) = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + col));
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps
      const int nbeg = n0 + half * kCols;
      const bool active = mrow0 < p.M && nbeg < p.N;
      uint4 pre[8];
      if (p.vec_ok && row_ok && nbeg + 32 <= p.N) epilogue_prefetch<EPI>(ep, m, nbeg, pre);
#ifdef DGPT_GEMM_TS
      const bool stamp = p.ts && blockIdx.x == 0 && warp == 2 && lane == 0 && t < 3 * t_step;
      long long* tsp = p.ts + (t / t_step) * 64;
      int tsi = 0;
#define TS_MARK() do { if (stamp) tsp[tsi++] = clock64(); } while (0)
#else
#define TS_MARK() do { } while (0)
#endif
      TS_MARK();
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      TS_MARK();
      const uint32_t row_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * kCols);
      // 32 accumulator columns per step; a staging tile (32 rows x 128 B) holds 32 fp32 or 64 bf16
      // columns, so bf16 output issues one TMA store every second step
      uint8_t* tile = my_stage + sbuf * 4096;
      uint32_t r[32];
#ifdef DGPT_GEMM_TS
      {  // micro-benchmark: 1 load vs 4 back-to-back loads (latency- or throughput-bound?)
        uint32_t q0[32], q1[32], q2[32], q3[32];
        TS_MARK();
        tmem_ld32(row_addr, q0);
        tmem_ld_wait();
        TS_MARK();
        tmem_ld32(row_addr, q0); tmem_ld32(row_addr + 32, q1); tmem_ld32(row_addr + 64, q2); tmem_ld32(row_addr + 96, q3);
        tmem_ld_wait();
        TS_MARK();
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) x ^= q0[j] ^ q1[j] ^ q2[j] ^ q3[j];
        if (x == 0x12345678u) tile[0] = 1;
        TS_MARK();
      }
#endif
      if (active) tmem_ld32(row_addr, r);
#pragma unroll 1
      for (int c = 0; c < kCols && active; c += 32) {
        const int n = nbeg + c;
        if (n >= p.N) break;
        const bool opens_tile = !bf16_out || (c & 32) == 0;
        if (mode != kStoreDirect && opens_tile) {
          tile = my_stage + sbuf * 4096;
          if (lane == 0) bulk_wait_read<1>();  // the store issued two tiles ago has drained this buffer
          __syncwarp();
        }
        TS_MARK();
        tmem_ld_wait();
        TS_MARK();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const bool has_next = c + 32 < kCols && n + 32 < p.N;
        if (has_next) tmem_ld32(row_addr + c + 32, r);  // in flight during this step's math
        if (p.vec_ok && n + 32 <= p.N && row_ok) {
          epilogue_math32<EPI>(ep, m, n, v, bias_t + half * kCols + c, pre);
        } else {
          float tmp[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) tmp[j] = v[j];
          epilogue_slow32(ep, tmp, m, n, mode == kStoreDirect);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tmp[j];
        }
        // operands of the next step: in flight during the staging / store below and the next TMEM wait
        if (has_next && p.vec_ok && row_ok && n + 64 <= p.N) epilogue_prefetch<EPI>(ep, m, n + 32, pre);
        TS_MARK();
        if (mode == kStoreDirect) continue;
        const bool closes_tile = !bf16_out || (c & 32) != 0 || !has_next;
        if (bf16_out) stage_half_row_bf16(tile, lane, (c >> 5) & 1, v);
        else stage_row_f32(tile, lane, v);
        TS_MARK();
        if (closes_tile) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const int ncol = bf16_out ? nbeg + (c & ~63) : n;
            if (mode == kStoreTmaAdd) tma_reduce_add_2d(&map_d, tile, ncol, mrow0);
            else tma_store_2d(&map_d, tile, ncol, mrow0);
            bulk_commit();
          }
          sbuf ^= 1;
        }
        TS_MARK();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));  // the leader's MMA thread waits for both CTAs
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();  // all output tiles have landed before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
#ifdef DGPT_GEMM_TS
  if (p.ts && threadIdx.x == 0) p.ts[1500 + blockIdx.x] = clock64() - t_block0;
#endif
  if (CG == 2) cluster_sync_all();  // neither CTA may free TMEM / retire while the pair's MMAs or arrives are in flight
  if (warp == 1) {
    tc_fence_after();
    if (CG == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
  }
}

// --------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
    else
      cudaGetLastError();
  }
  return fn;
}

// 2-D tensor map with 128-byte swizzle and zero fill out of bounds: inner (contiguous) extent `inner`
// elements, `outer` rows of pitch ld elements, box = box_inner x box_outer elements (box_inner * esize == 128).
int make_tmap_2d(CUtensorMap* map, const void* base, int dtype, int64_t inner, int64_t outer, int64_t ld,
                 int box_inner, int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return DGPT_E_DEVICE;
  }
  const int esize = dtype == DGPT_F32 ? 4 : 2;
  DGPT_REQUIRE(((uintptr_t)base & 15) == 0 && (ld * esize) % 16 == 0,
               "TMA needs a 16-byte aligned base and row pitch (ld=%lld elements of %d bytes)", (long long)ld, esize);
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, dtype == DGPT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld outer=%lld ld=%lld)", (int)r,
              (long long)inner, (long long)outer, (long long)ld);
    return DGPT_E_ARG;
  }
  return DGPT_OK;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
  return make_tmap_2d(map, base, DGPT_BF16, inner, outer, ld, 64, box_outer);
}

template <int BN, int A_MN, int B_MN, int EPI, int CG>
static int launch_one(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const TcParams& p, int grid,
                      cudaStream_t st) {
  using Cfg = TcCfg<BN, CG>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI, CG>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute(%zu B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return DGPT_E_LAUNCH;
    }
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (CG == 2) {
    // a persistent grid must not exceed the number of CTA pairs that can be resident at once (GPCs with an
    // odd number of usable SMs leave single SMs that cannot host a pair)
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3(2 * 74);
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q) != cudaSuccess || max_clusters <= 0) {
        cudaGetLastError();
        max_clusters = 64;
      }
      if (getenv("DGPT_GEMM_VERBOSE")) fprintf(stderr, "gemm_tc: max active 2-CTA clusters = %d\n", max_clusters);
    }
    if (grid > 2 * max_clusters) cfg.gridDim = dim3((unsigned)(2 * max_clusters));
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, md, p);
  if (e != cudaSuccess) {
    set_error("gemm_tc: launch: %s", cudaGetErrorString(e));
    return DGPT_E_LAUNCH;
  }
  return check_launch("gemm_tc");
}

// Pair tiles (cta_group::2) are opt-in (dgpt_gemm_set_cta_group): on the model's shapes they measured no faster
// than single-CTA tiles -- the 20-30 us GEMMs are bound by fixed per-launch / per-tile latencies and the
// epilogue, not by operand delivery -- so the default stays 1.
static int g_cta_group = -1;  // -1: not set yet (DGPT_GEMM_CTA_GROUP, default 1)
void set_gemm_cta_group(int g) { g_cta_group = (g == 2) ? 2 : 1; }
static int gemm_cta_group() {
  if (g_cta_group < 0) {
    const char* e = getenv("DGPT_GEMM_CTA_GROUP");
    g_cta_group = (e && atoi(e) == 2) ? 2 : 1;
  }
  return g_cta_group;
}

// pair tiles whenever requested and the row-tile count is even, single-CTA tiles otherwise
template <int BN, int A_MN, int B_MN, int EPI>
static int launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb_half, const CUtensorMap& md,
                      const TcParams& p, int sms, cudaStream_t st) {
  if (p.cta_group == 2) {
    const int total = (p.m_tiles / 2) * p.n_tiles * p.split_k;
    return launch_one<BN, A_MN, B_MN, EPI, 2>(ma, mb_half, md, p, 2 * min(total, sms / 2), st);
  }
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  return launch_one<BN, A_MN, B_MN, EPI, 1>(ma, mb, md, p, min(total, sms), st);
}

int launch_gemm_tc(const dgpt_gemm_args* a, cudaStream_t st) {
  DGPT_REQUIRE(a->K > 0, "gemm(bf16): K must be positive");
  const int a_mn = a->a_major == DGPT_MAJOR_MN, b_mn = a->b_major == DGPT_MAJOR_MN;
  DGPT_REQUIRE(!(a_mn && !b_mn), "gemm(bf16): A MN-major with B K-major is not instantiated");
  // tile N: widest tile that still yields >= ~1 wave of CTAs
  int sms = dgpt_sm_count();
  if (sms <= 0) sms = 148;
  const int m_tiles = ceil_div(a->M, TBM);
  // 128 x 256 tiles ingest 25 % fewer operand bytes per MMA cycle than 128 x 128; a ragged last column
  // tile (N = 1152 -> 4.5 tiles) costs less than that as soon as N >= 1024 (TMA zero-fills, the store clips)
  int BN = 256;
  if (a->N <= 128 || (a->N % 256 != 0 && a->N < 1024)) BN = 128;
  if (BN == 256 && m_tiles * ceil_div(a->N, 256) * (a->split_k > 1 ? a->split_k : 1) < sms) BN = 128;
  const int n_tiles = ceil_div(a->N, BN);
  TcParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.m_tiles = m_tiles; p.n_tiles = n_tiles;
  p.kb_total = ceil_div(a->K, TBK);
  p.split_k = a->split_k > 1 ? min(a->split_k, p.kb_total) : 1;
  p.kb_per_split = ceil_div(p.kb_total, p.split_k);
  p.split_k = ceil_div(p.kb_total, p.kb_per_split);
  p.ep = make_epilogue(a);
  p.ep.atomic = p.split_k > 1;
  if (p.split_k > 1 && !a->accumulate) {
    cudaError_t e = cudaMemset2DAsync(a->D, (size_t)a->ldd * 4, 0, (size_t)a->N * 4, (size_t)a->M, st);
    if (e != cudaSuccess) { set_error("gemm_tc: memset: %s", cudaGetErrorString(e)); return DGPT_E_LAUNCH; }
  }
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  p.vec_ok = (a->N % 4 == 0) && (!a->bias || al16(a->bias)) && (!a->residual || (al16(a->residual) && a->ldr % 4 == 0)) &&
             (!a->relu_aux || (al16(a->relu_aux) && a->ld_aux % 8 == 0));

  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("DGPT_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
#ifdef DGPT_GEMM_TS
    static long long* ts = nullptr;
    if (!ts) { cudaMalloc(&ts, 2048 * sizeof(long long)); }
    cudaMemset(ts, 0, 2048 * sizeof(long long));
    p.ts = ts;
#endif
  }
  CUtensorMap ma, mb, md;
  int rc;
  if (a_mn) rc = make_tmap_bf16_2d(&ma, a->A, a->M, a->K, a->lda, 64);
  else rc = make_tmap_bf16_2d(&ma, a->A, a->K, a->M, a->lda, TBM);
  if (rc) return rc;
  CUtensorMap mb_half;  // B as loaded by one CTA of a pair: half of the tile's rows
  if (b_mn) rc = make_tmap_bf16_2d(&mb, a->B, a->N, a->K, a->ldb, 64);
  else rc = make_tmap_bf16_2d(&mb, a->B, a->K, a->N, a->ldb, BN);
  if (rc) return rc;
  if (b_mn) mb_half = mb;
  else if ((rc = make_tmap_bf16_2d(&mb_half, a->B, a->K, a->N, a->ldb, BN / 2))) return rc;
  p.cta_group = (gemm_cta_group() == 2 && m_tiles % 2 == 0) ? 2 : 1;
  // output through TMA when its pitch allows it (always true for the model's buffers)
  const int desz = a->d_dtype == DGPT_F32 ? 4 : 2;
  p.store_mode = kStoreDirect;
  if (al16(a->D) && ((int64_t)a->ldd * desz) % 16 == 0) {
    rc = make_tmap_2d(&md, a->D, a->d_dtype, a->N, a->M, a->ldd, 128 / desz, 32);
    if (rc) return rc;
    p.store_mode = (p.ep.atomic || a->accumulate) ? kStoreTmaAdd : kStoreTma;
  } else {
    md = ma;
  }
  if (p.store_mode == kStoreDirect || a->D2 || (a->residual && a->relu_aux)) p.vec_ok = 0;  // everything through the compact element-wise path

  // epilogue specialisation: the combinations the model uses get their own instantiation, the rest run
  // the generic (run-time flag) kernel
  const int epi = (a->bias ? kEpiBias : 0) | (a->relu ? kEpiRelu : 0) | (a->relu_aux ? kEpiAux : 0) |
                  (a->dropout_p > 0.f ? kEpiDrop : 0) | (a->residual ? kEpiRes : 0);
#ifdef DGPT_GEMM_TS
  {  // debug build: run synchronously and print the stamp deltas of the first three tiles of block 0
    int rcx = (BN == 256) ? launch_cfg<256, 0, 0, kEpiBias | kEpiRelu>(ma, mb, mb_half, md, p, sms, st)
                          : launch_cfg<128, 0, 0, kEpiBias | kEpiRelu>(ma, mb, mb_half, md, p, sms, st);
    cudaStreamSynchronize(st);
    long long h[2048];
    cudaMemcpy(h, p.ts, sizeof(h), cudaMemcpyDeviceToHost);
    {
      printf("block durations (cycles):");
      for (int i = 0; i < 148; ++i) printf(" %lld", h[1500 + i]);
      printf("\n");
    }
    for (int t = 0; t < 3; ++t) {
      printf("tile %d:", t);
      for (int i = 1; i < 48 && h[t * 64 + i]; ++i) printf(" %lld", h[t * 64 + i] - h[t * 64 + i - 1]);
      printf("\nblock durations (cycles):");
      for (int i = 0; i < 148; ++i) printf(" %lld", h[1500 + i]);
      printf("\n");
    }
    fflush(stdout);
    return rcx;
  }
#endif
#define TC_EPI(BN_, A_, B_, E_) \
  if (epi == (E_)) return launch_cfg<BN_, A_, B_, (E_)>(ma, mb, mb_half, md, p, sms, st);
#define TC_DISPATCH(BN_)                                                          \
  if (BN == BN_) {                                                                \
    if (!a_mn && !b_mn) {                                                         \
      TC_EPI(BN_, 0, 0, 0)                                                        \
      TC_EPI(BN_, 0, 0, kEpiBias)                                                 \
      TC_EPI(BN_, 0, 0, kEpiBias | kEpiRelu)                                      \
      TC_EPI(BN_, 0, 0, kEpiBias | kEpiRes)                                       \
      TC_EPI(BN_, 0, 0, kEpiBias | kEpiDrop | kEpiRes)                            \
      return launch_cfg<BN_, 0, 0, -1>(ma, mb, mb_half, md, p, sms, st);                  \
    }                                                                             \
    if (!a_mn && b_mn) {                                                          \
      TC_EPI(BN_, 0, 1, 0)                                                        \
      TC_EPI(BN_, 0, 1, kEpiAux)                                                  \
      return launch_cfg<BN_, 0, 1, -1>(ma, mb, mb_half, md, p, sms, st);                  \
    }                                                                             \
    TC_EPI(BN_, 1, 1, 0)                                                          \
    return launch_cfg<BN_, 1, 1, -1>(ma, mb, mb_half, md, p, sms, st);                    \
  }
  TC_DISPATCH(128)
  TC_DISPATCH(256)
#undef TC_DISPATCH
#undef TC_EPI
  set_error("gemm_tc: no tile configuration for N=%d", a->N);
  return DGPT_E_ARG;
}

}  // namespace dgpt
