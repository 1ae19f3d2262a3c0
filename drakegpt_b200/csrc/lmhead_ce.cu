// Fused LM head + cross-entropy for sm_100a (tensor mode).
//
// Replaces, in one kernel, `logits = lm_head(x)` (src/model.py:599) and `F.cross_entropy(logits, targets)`
// (src/model.py:604-607) plus the first step of their backward, d(loss)/d(logits):
//
//   one CTA per 128 rows of x:  TMA loads the [128 x K] bf16 activation tile and the whole [V x K] bf16 weight
//   (V = 80 characters: 60 KB, it fits next to the activation tile) -> tcgen05.mma 128 x V x 16 into V fp32 TMEM
//   columns -> 128 epilogue threads, one per row, read their logits row straight from TMEM: + bias, row max,
//   sum of exponentials, loss_sum += (lse - logit[target]) / M, and dlogits = (softmax - onehot) * dloss / M
//   written as bf16 for the backward GEMMs.  The fp32 logits never touch HBM unless the caller asks for them
//   (the reference API returns logits; the training step does not need them).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dgpt {

using namespace ptx;

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_outer);

static constexpr int kLmThreads = 64 + 128;  // TMA warp, MMA warp, 4 epilogue warps
static constexpr int kLmMaxKb = 8;           // K <= 512

struct LmCeP {
  const float* bias;
  const int64_t* targets;
  float* loss_sum;
  __nv_bfloat16* dlogits;
  float* logits;
  const float* dloss;
  int ld_dl, ld_lg;
  int M, V, kb;  // kb = K / 64
};

__global__ void __launch_bounds__(kLmThreads, 1)
lmhead_ce_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, LmCeP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int b_bytes = p.V * 128;  // one k-block of W: V rows x 128 B (a multiple of 1024: V % 8 == 0)
  uint8_t* sA = smem;                                  // [kb][128 x 64] bf16
  uint8_t* sB = smem + (size_t)p.kb * 16384;           // [kb][V x 64] bf16
  float* bias_s = reinterpret_cast<float*>(sB + (size_t)p.kb * b_bytes);  // [V]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + 256);        // [kb]
  uint64_t* acc_bar = full_bar + kLmMaxKb;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_w);
    for (int i = 0; i < p.kb; ++i) mbar_init(&full_bar[i], 1);
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_grid_sync();  // prologue (barriers, TMEM) overlapped the previous kernel; global memory from here on

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < p.kb; ++kb) {
        mbar_expect_tx(&full_bar[kb], 16384 + b_bytes);
        tma_load_2d(sA + kb * 16384, &map_x, &full_bar[kb], kb * 64, m0);
        tma_load_2d(sB + (size_t)kb * b_bytes, &map_w, &full_bar[kb], kb * 64, 0);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, p.V, 0, 0);
    for (int kb = 0; kb < p.kb; ++kb) {
      mbar_wait(&full_bar[kb], 0);
      tc_fence_after();
      const uint64_t da = make_smem_desc_sw128(smem_u32(sA + kb * 16384), 16, 1024);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sB + (size_t)kb * b_bytes), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(acc_bar);
    __syncwarp();
  } else {
    // ------------------------------ epilogue: one thread per row --------------------------
    const int et = threadIdx.x - 64;  // 0..127
    for (int i = et; i < p.V; i += 128) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int quad = warp & 3;  // TMEM lane quadrant of this warp
    const int m = m0 + quad * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
    const int tgt = (m < p.M && p.targets) ? (int)p.targets[m] : -1;
    const int nch = p.V >> 4;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    // pass 1: row max; pass 2: sum of exponentials and the target logit; pass 3: outputs.  TMEM re-reads are cheap
    // (the V columns are read three times instead of holding up to 256 logits in registers).
    float mx = -INFINITY;
    for (int c = 0; c < nch; ++c) {
      uint32_t r[16];
      tmem_ld16(taddr + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(r[j]) + bias_s[c * 16 + j]);
    }
    float se = 0.f, vt = 0.f;
    for (int c = 0; c < nch; ++c) {
      uint32_t r[16];
      tmem_ld16(taddr + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = __uint_as_float(r[j]) + bias_s[c * 16 + j];
        se += __expf(v - mx);
        vt = (c * 16 + j == tgt) ? v : vt;
      }
    }
    float row_loss = 0.f;
    if (tgt >= 0) row_loss = (logf(se) + mx - vt) / (float)p.M;
    // (tcgen05.ld is warp-collective: the loop runs for every lane of the warp, only the stores are predicated on
    // the row being inside the matrix -- the last row tile may be ragged)
    const bool row_ok = m < p.M;
    const bool want_dl = p.dlogits != nullptr && tgt >= 0 && row_ok;
    if (p.dlogits != nullptr || p.logits != nullptr) {
      const float scale = (p.dloss ? p.dloss[0] : 1.f) / (float)p.M;
      const float inv = 1.f / se;
      __nv_bfloat16* dr = want_dl ? p.dlogits + (int64_t)m * p.ld_dl : nullptr;
      float* lr = (p.logits && row_ok) ? p.logits + (int64_t)m * p.ld_lg : nullptr;
      for (int c = 0; c < nch; ++c) {
        uint32_t r[16];
        tmem_ld16(taddr + c * 16, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + bias_s[c * 16 + j];
        if (lr) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(lr + c * 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (want_dl) {
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float d0 = (__expf(v[j] - mx) * inv - ((c * 16 + j) == tgt ? 1.f : 0.f)) * scale;
            const float d1 = (__expf(v[j + 1] - mx) * inv - ((c * 16 + j + 1) == tgt ? 1.f : 0.f)) * scale;
            __nv_bfloat162 t = __floats2bfloat162_rn(d0, d1);
            w[j >> 1] = *reinterpret_cast<uint32_t*>(&t);
          }
          *reinterpret_cast<uint4*>(dr + c * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(dr + c * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
    }
    if (p.loss_sum && p.targets) {
      row_loss = warp_sum(row_loss);
      if (lane == 0) atomicAdd(p.loss_sum, row_loss);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

static size_t lm_smem_bytes(int V, int kb) { return (size_t)kb * (16384 + V * 128) + 256 * 4 + (kLmMaxKb + 1) * 8 + 16; }

bool lmhead_ce_supported(int V, int K) {
  if (V % 16 != 0 || V < 16 || V > 256 || K % 64 != 0 || K < 64 || K / 64 > kLmMaxKb) return false;
  return lm_smem_bytes(V, K / 64) <= 227 * 1024;
}

}  // namespace dgpt

using namespace dgpt;

extern "C" {

int dgpt_lmhead_ce_supported(int V, int K) { return lmhead_ce_supported(V, K) ? 1 : 0; }

int dgpt_lmhead_ce(const void* x, int ldx, const void* w, int ldw, const float* bias, const int64_t* targets,
                   float* loss_sum, void* dlogits, int ld_dl, float* logits, int ld_lg, const float* dloss, int M,
                   int V, int K, void* stream) {
  DGPT_DEVICE_OR_RETURN();
  DGPT_REQUIRE(M >= 0 && x && w, "lmhead_ce: bad arguments");
  if (M == 0) return DGPT_OK;
  DGPT_REQUIRE(lmhead_ce_supported(V, K), "lmhead_ce: needs V %% 16 == 0, 16 <= V <= 256, K %% 64 == 0, K <= 512 and "
               "K * (256 + 2 V) bytes of shared memory (V=%d K=%d)", V, K);
  DGPT_REQUIRE(!dlogits || (targets && ld_dl >= V && ld_dl % 8 == 0 && ((uintptr_t)dlogits & 15) == 0),
               "lmhead_ce: dlogits needs targets, a 16-byte aligned base and ld_dl %% 8 == 0 (ld_dl=%d)", ld_dl);
  DGPT_REQUIRE(!logits || (ld_lg >= V && ld_lg % 4 == 0 && ((uintptr_t)logits & 15) == 0),
               "lmhead_ce: logits needs a 16-byte aligned base and ld_lg %% 4 == 0 (ld_lg=%d)", ld_lg);
  DGPT_REQUIRE(!targets || loss_sum, "lmhead_ce: targets need loss_sum");
  CUtensorMap mx, mw;
  int rc;
  if ((rc = make_tmap_bf16_2d(&mx, x, K, M, ldx, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&mw, w, K, V, ldw, V))) return rc;
  const int kb = K / 64;
  const size_t smem = lm_smem_bytes(V, kb);
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(lmhead_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) { set_error("lmhead_ce: smem attribute: %s", cudaGetErrorString(e)); return DGPT_E_LAUNCH; }
    attr_bytes = 227 * 1024;
  }
  LmCeP p;
  p.bias = bias; p.targets = targets; p.loss_sum = loss_sum; p.dlogits = (__nv_bfloat16*)dlogits; p.logits = logits;
  p.dloss = dloss; p.ld_dl = ld_dl; p.ld_lg = ld_lg; p.M = M; p.V = V; p.kb = kb;
  launch_pdl(lmhead_ce_kernel, dim3(ceil_div(M, 128)), dim3(kLmThreads), smem, (cudaStream_t)stream, mx, mw, p);
  return check_launch("lmhead_ce");
}

}  // extern "C"
