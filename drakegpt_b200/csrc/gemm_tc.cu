// Tensor-mode GEMM for sm_100a: bf16 operands staged by TMA (128B swizzle) into a
// multi-stage shared-memory ring, tcgen05.mma (cta_group::1, 128 x BN x 16) issued
// by one thread with fp32 accumulators in TMEM, and a fused epilogue
// (bias / ReLU / ReLU-mask / dropout / residual / fp32+bf16 stores / split-K
// red.add) read back with tcgen05.ld.  Persistent: one CTA per SM walks the tile
// list; TMEM holds two accumulators so the epilogue of tile i overlaps the
// mainloop of tile i+1.
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
static constexpr size_t kSmemBudget = 200 * 1024;

struct TcParams {
  int M, N, K;
  int m_tiles, n_tiles, split_k, kb_total, kb_per_split;
  int vec_ok;  // epilogue may use 16-byte accesses
  Epilogue ep;
};

template <int BN>
struct TcCfg {
  static constexpr int kStageBytes = TBM * TBK * 2 + BN * TBK * 2;
  static constexpr int kStages = (int)(kSmemBudget / kStageBytes) > 8 ? 8 : (int)(kSmemBudget / kStageBytes);
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // 128 / 256 / 512 (powers of two)
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// --------------------------------------------------------------------------
// epilogue for one 32-column chunk of one accumulator row
// --------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_chunk32(const Epilogue& e, int vec_ok, int m, int n, uint32_t (&r)[32]) {
  if (m >= e.M) return;
  if (vec_ok && n + 32 <= e.N) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
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
    if (e.d_dtype == DGPT_F32) {
      float* d = reinterpret_cast<float*>(e.D) + (int64_t)m * e.ldd + n;
      if (e.atomic) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(d + j, v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (e.accumulate) {
            const float4 c = reinterpret_cast<float4*>(d)[j];
            o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
            v[4 * j] = o.x; v[4 * j + 1] = o.y; v[4 * j + 2] = o.z; v[4 * j + 3] = o.w;
          }
          reinterpret_cast<float4*>(d)[j] = o;
        }
      }
    } else {
      uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.D) + (int64_t)m * e.ldd + n);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        d[j] = o;
      }
    }
    if (e.D2) {
      if (e.d2_dtype == DGPT_F32) {
        float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.D2) + (int64_t)m * e.ldd2 + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.D2) + (int64_t)m * e.ldd2 + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
          uint4 o;
          o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
          o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
          d[j] = o;
        }
      }
    }
    return;
  }
  // ragged / unaligned tail: element-wise
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    if (n + j < e.N) epilogue_store(e, m, n + j, epilogue_value(e, m, n + j, __uint_as_float(r[j])));
  }
}

// --------------------------------------------------------------------------
// the kernel
// --------------------------------------------------------------------------
template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kABytes = TBM * TBK * 2;
  constexpr int kBBytes = BN * TBK * 2;
  constexpr uint32_t kIdesc = make_idesc_bf16(TBM, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
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
    int acc = 0;
    uint32_t acc_ph = 0;
    Epilogue ep = p.ep;
    epilogue_resolve_seed(ep);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int ks = t / tiles_mn, mn = t - ks * tiles_mn;
      const int m0 = (mn / p.n_tiles) * TBM, n0 = (mn % p.n_tiles) * BN;
      ep.first_split = (ks == 0);
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const int m = m0 + quad * 32 + lane;
      const uint32_t row_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= p.N) break;
        uint32_t r[32];
        tmem_ld32(row_addr + c, r);
        tmem_ld_wait();
        epilogue_chunk32(ep, p.vec_ok, m, n0 + c, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
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

// 2-D bf16 tensor map: inner (contiguous) extent `inner`, `outer` rows of pitch ld elements;
// box = 64 x box_outer elements, 128-byte swizzle, zero fill out of bounds.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return DGPT_E_DEVICE;
  }
  DGPT_REQUIRE(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0,
               "tensor-core operand needs a 16-byte aligned base and row pitch (ld=%lld)", (long long)ld);
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld outer=%lld ld=%lld)", (int)r,
              (long long)inner, (long long)outer, (long long)ld);
    return DGPT_E_ARG;
  }
  return DGPT_OK;
}

template <int BN, int A_MN, int B_MN>
static int launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int grid, cudaStream_t st) {
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
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, kThreads, Cfg::kSmemBytes, st>>>(ma, mb, p);
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
  p.vec_ok = (a->N % 4 == 0) && al16(a->D) && (a->ldd % 8 == 0) && (!a->D2 || (al16(a->D2) && a->ldd2 % 8 == 0)) &&
             (!a->bias || al16(a->bias)) && (!a->residual || (al16(a->residual) && a->ldr % 4 == 0)) &&
             (!a->relu_aux || (al16(a->relu_aux) && a->ld_aux % 8 == 0));

  CUtensorMap ma, mb;
  int rc;
  if (a_mn) rc = make_tmap_bf16_2d(&ma, a->A, a->M, a->K, a->lda, 64);
  else rc = make_tmap_bf16_2d(&ma, a->A, a->K, a->M, a->lda, TBM);
  if (rc) return rc;
  if (b_mn) rc = make_tmap_bf16_2d(&mb, a->B, a->N, a->K, a->ldb, 64);
  else rc = make_tmap_bf16_2d(&mb, a->B, a->K, a->N, a->ldb, BN);
  if (rc) return rc;

  const int total = m_tiles * n_tiles * p.split_k;
  const int grid = min(total, sms);
#define TC_DISPATCH(BN_)                                                                   \
  if (BN == BN_) {                                                                         \
    if (!a_mn && !b_mn) return launch_cfg<BN_, 0, 0>(ma, mb, p, grid, st);                 \
    if (!a_mn && b_mn) return launch_cfg<BN_, 0, 1>(ma, mb, p, grid, st);                  \
    return launch_cfg<BN_, 1, 1>(ma, mb, p, grid, st);                                     \
  }
  TC_DISPATCH(64)
  TC_DISPATCH(128)
  TC_DISPATCH(256)
#undef TC_DISPATCH
  set_error("gemm_tc: no tile configuration for N=%d", a->N);
  return DGPT_E_ARG;
}

}  // namespace dgpt
