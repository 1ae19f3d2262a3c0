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
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer, warps 2..5 = epilogue (TMEM lane quadrant = warp_id % 4).
//
// Operand layouts: K-major ([rows, K], K contiguous: activations, nn.Linear
// weights) and MN-major ([K, rows], rows contiguous: the transposed views that
// dgrad / wgrad need) are both fed straight from the row-major tensors -- no
// transposed copies are ever written to HBM.
#include <cuda.h>

#include "epilogue.cuh"
#include "ptx.cuh"

namespace dgpt {

using namespace ptx;

static constexpr int TBM = 128;       // tile M (UMMA M)
static constexpr int TBK = 64;        // k-block: 64 bf16 = one 128-byte swizzle row
static constexpr int UMMA_K = 16;
static constexpr int kThreads = 192;
static constexpr int kStageBudget = 192 * 1024;
static constexpr int kStagingBytes = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 B)

enum StoreMode { kStoreDirect = 0, kStoreTma = 1, kStoreTmaAdd = 2 };

struct TcParams {
  int M, N, K;
  int m_tiles, n_tiles, split_k, kb_total, kb_per_split;
  int vec_ok;      // epilogue operands (bias / residual / relu_aux) allow 16-byte loads
  int store_mode;  // StoreMode for D
  Epilogue ep;
};

template <int BN>
struct TcCfg {
  static constexpr int kStageBytes = TBM * TBK * 2 + BN * TBK * 2;
  static constexpr int kStages = kStageBudget / kStageBytes > 8 ? 8 : kStageBudget / kStageBytes;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // 128 / 256 / 512 (powers of two)
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kStagingBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// --------------------------------------------------------------------------
// epilogue math on 32 consecutive columns [n, n+32) of accumulator row m (all in registers)
// --------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_math32(const Epilogue& e, int vec_ok, int m, int n, float (&v)[32]) {
  if (m >= e.M) return;
  if (!(vec_ok && n + 32 <= e.N)) {  // ragged edge / unaligned operands: guarded scalar path (fully unrolled)
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < e.N) v[j] = epilogue_value(e, m, n + j, v[j]);
    return;
  }
  if (e.first_split && e.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (e.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (e.relu_aux) {
    if (e.aux_dtype == DGPT_BF16) {
      const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(e.relu_aux) + (int64_t)m * e.ld_aux + n);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 a = __ldg(ap + j);
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          // bf16 > 0  <=>  sign bit clear and magnitude non-zero
          const uint32_t lo = w[t] & 0xFFFFu, hi = w[t] >> 16;
          if (!(lo != 0 && lo < 0x8000u)) v[j * 8 + t * 2] = 0.f;
          if (!(hi != 0 && hi < 0x8000u)) v[j * 8 + t * 2 + 1] = 0.f;
        }
      }
    } else {
      const float4* ap = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.relu_aux) + (int64_t)m * e.ld_aux + n);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = __ldg(ap + j);
        if (!(a.x > 0.f)) v[4 * j] = 0.f;
        if (!(a.y > 0.f)) v[4 * j + 1] = 0.f;
        if (!(a.z > 0.f)) v[4 * j + 2] = 0.f;
        if (!(a.w > 0.f)) v[4 * j + 3] = 0.f;
      }
    }
  }
  if (e.thr) {
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
  if (e.first_split && e.residual) {
    const float4* rp = reinterpret_cast<const float4*>(e.residual + (int64_t)m * e.ldr + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = __ldg(rp + j);
      v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
    }
  }
}

// element-wise stores for outputs the TMA cannot address (unaligned pitch) and for the optional D2
__device__ __forceinline__ void direct_store32(const Epilogue& e, bool main_out, int m, int n, const float (&v)[32]) {
  if (m >= e.M) return;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (n + j >= e.N) continue;
    if (main_out) {
      const int64_t i = (int64_t)m * e.ldd + n + j;
      if (e.d_dtype == DGPT_F32) {
        float* d = reinterpret_cast<float*>(e.D);
        if (e.atomic) atomicAdd(d + i, v[j]);
        else d[i] = e.accumulate ? d[i] + v[j] : v[j];
      } else {
        reinterpret_cast<__nv_bfloat16*>(e.D)[i] = __float2bfloat16_rn(v[j]);
      }
    } else {
      const int64_t i = (int64_t)m * e.ldd2 + n + j;
      if (e.d2_dtype == DGPT_F32) reinterpret_cast<float*>(e.D2)[i] = v[j];
      else reinterpret_cast<__nv_bfloat16*>(e.D2)[i] = __float2bfloat16_rn(v[j]);
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
template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_d, TcParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kABytes = TBM * TBK * 2;
  constexpr int kBBytes = BN * TBK * 2;
  constexpr uint32_t kIdesc = make_idesc_bf16(TBM, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + (size_t)kStages * Cfg::kStageBytes;  // 1024-aligned: stage sizes are multiples of 8 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

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
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_mn * p.split_k;

  if (warp == 0) {
    // ------------------------------ TMA producer ---------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ks = t / tiles_mn, mn = t - ks * tiles_mn;
        const int m0 = (mn / p.n_tiles) * TBM, n0 = (mn % p.n_tiles) * BN;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = stage_base + (size_t)s * Cfg::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[s], kABytes + kBBytes);
          const int k0 = kb * TBK;
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
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ks = t / tiles_mn;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)s * Cfg::kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < TBK / UMMA_K; ++k) {
            // K-major: 32 bytes per UMMA_K inside the 128B swizzle row, 8-row groups 1024 B apart.
            // MN-major: 16 k-rows (2 KB) per UMMA_K, 64-element MN chunks 8 KB apart.
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sb + k * 32, 16, 1024);
            tc_mma_bf16(d_tmem, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);  // smem stage is free once these MMAs retire
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        tc_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue -------------------------------
    const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32)
    uint8_t* my_stage = staging + quad * 8192;
    int sbuf = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    Epilogue ep = p.ep;
    epilogue_resolve_seed(ep);
    const bool bf16_out = ep.d_dtype == DGPT_BF16;
    const int mode = p.store_mode;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int ks = t / tiles_mn, mn = t - ks * tiles_mn;
      const int m0 = (mn / p.n_tiles) * TBM, n0 = (mn % p.n_tiles) * BN;
      ep.first_split = (ks == 0);
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const int mrow0 = m0 + quad * 32;
      const int m = mrow0 + lane;
      const uint32_t row_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      if (mode != kStoreDirect && bf16_out) {
        // 64 columns (= 128 bytes of bf16) per staging tile
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          if (n0 + c >= p.N || mrow0 >= p.M) break;
          uint32_t r0[32], r1[32];
          tmem_ld32(row_addr + c, r0);
          tmem_ld32(row_addr + c + 32, r1);
          uint8_t* tile = my_stage + sbuf * 4096;
          if (lane == 0) bulk_wait_read<1>();  // the store issued two tiles ago has drained this buffer
          __syncwarp();
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
          epilogue_math32(ep, p.vec_ok, m, n0 + c, v);
          if (ep.D2) direct_store32(ep, false, m, n0 + c, v);
          stage_half_row_bf16(tile, lane, 0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r1[j]);
          epilogue_math32(ep, p.vec_ok, m, n0 + c + 32, v);
          if (ep.D2) direct_store32(ep, false, m, n0 + c + 32, v);
          stage_half_row_bf16(tile, lane, 1, v);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_d, tile, n0 + c, mrow0);
            bulk_commit();
          }
          sbuf ^= 1;
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          if (n0 + c >= p.N || mrow0 >= p.M) break;
          uint32_t r[32];
          tmem_ld32(row_addr + c, r);
          uint8_t* tile = my_stage + sbuf * 4096;
          if (mode != kStoreDirect) {
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_math32(ep, p.vec_ok, m, n0 + c, v);
          if (ep.D2) direct_store32(ep, false, m, n0 + c, v);
          if (mode == kStoreDirect) {
            direct_store32(ep, true, m, n0 + c, v);
          } else {
            stage_row_f32(tile, lane, v);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (mode == kStoreTmaAdd) tma_reduce_add_2d(&map_d, tile, n0 + c, mrow0);
              else tma_store_2d(&map_d, tile, n0 + c, mrow0);
              bulk_commit();
            }
            sbuf ^= 1;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();  // all output tiles have landed before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
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

template <int BN, int A_MN, int B_MN>
static int launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const TcParams& p, int grid,
                      cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute(%zu B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return DGPT_E_LAUNCH;
    }
    attr_done = true;
  }
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, kThreads, Cfg::kSmemBytes, st>>>(ma, mb, md, p);
  return check_launch("gemm_tc");
}

int launch_gemm_tc(const dgpt_gemm_args* a, cudaStream_t st) {
  DGPT_REQUIRE(a->K > 0, "gemm(bf16): K must be positive");
  const int a_mn = a->a_major == DGPT_MAJOR_MN, b_mn = a->b_major == DGPT_MAJOR_MN;
  DGPT_REQUIRE(!(a_mn && !b_mn), "gemm(bf16): A MN-major with B K-major is not instantiated");
  // tile N: widest tile that still yields >= ~1 wave of CTAs
  int sms = dgpt_sm_count();
  if (sms <= 0) sms = 148;
  const int m_tiles = ceil_div(a->M, TBM);
  int BN = 256;
  if (a->N <= 64) BN = 64;
  else if (a->N <= 128 || a->N % 256 != 0) BN = 128;
  if (BN == 256 && m_tiles * (a->N / 256) * (a->split_k > 1 ? a->split_k : 1) < sms) BN = 128;
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

  CUtensorMap ma, mb, md;
  int rc;
  if (a_mn) rc = make_tmap_bf16_2d(&ma, a->A, a->M, a->K, a->lda, 64);
  else rc = make_tmap_bf16_2d(&ma, a->A, a->K, a->M, a->lda, TBM);
  if (rc) return rc;
  if (b_mn) rc = make_tmap_bf16_2d(&mb, a->B, a->N, a->K, a->ldb, 64);
  else rc = make_tmap_bf16_2d(&mb, a->B, a->K, a->N, a->ldb, BN);
  if (rc) return rc;
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

  const int total = m_tiles * n_tiles * p.split_k;
  const int grid = min(total, sms);
#define TC_DISPATCH(BN_)                                                                   \
  if (BN == BN_) {                                                                         \
    if (!a_mn && !b_mn) return launch_cfg<BN_, 0, 0>(ma, mb, md, p, grid, st);             \
    if (!a_mn && b_mn) return launch_cfg<BN_, 0, 1>(ma, mb, md, p, grid, st);              \
    return launch_cfg<BN_, 1, 1>(ma, mb, md, p, grid, st);                                 \
  }
  TC_DISPATCH(64)
  TC_DISPATCH(128)
  TC_DISPATCH(256)
#undef TC_DISPATCH
  set_error("gemm_tc: no tile configuration for N=%d", a->N);
  return DGPT_E_ARG;
}

}  // namespace dgpt
