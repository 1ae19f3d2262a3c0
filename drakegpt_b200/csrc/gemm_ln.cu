// Residual GEMM with the following LayerNorm folded into its epilogue (tensor mode, sm_100a).
//
// Replaces, in one kernel, the tail of a transformer sub-block and the head of the next one
// (src/model_component.py:505-506, `x = x + self.sa(self.ln1(x))` / `x = x + self.ffwd(self.ln2(x))`):
//
//   x_out[M, N] (fp32) = dropout(A[M, K] @ W[N, K]^T + bias) + residual[M, N]
//   y[M, N]     (bf16) = (x_out - mean) * rstd * gamma + beta,   mean[M], rstd[M] kept for the backward pass
//
// One CTA per 128 rows and the FULL row of N <= 384 output columns: the fp32 accumulator row of every token sits in
// TMEM (row = lane), so the row statistics need no second kernel and no second trip of x_out through HBM:
//   warp 0     TMA producer: [128 x 64] A and [N x 64] W k-blocks through a shared-memory ring
//   warp 1     tcgen05.mma 128 x N x 16 (two N / 2 halves when N > 256) into N TMEM columns
//   warps 2-9  epilogue, two threads per row (column halves).  Once the accumulator is complete the ring is dead
//              and becomes staging space: the TMA drops the CTA's whole fp32 residual tile into it (every block in
//              flight at once), then
//                pass 1: TMEM -> + bias -> dropout -> + residual -> back to TMEM, x_out block leaves by TMA store
//                pass 2: TMEM -> sum of squared deviations (two-pass variance, like the stand-alone LayerNorm)
//                pass 3: TMEM -> normalise, scale, shift -> bf16 block -> TMA store
// Same arithmetic order as gemm_tc's epilogue followed by ln_fwd (bias, dropout, residual; two-pass statistics).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dgpt {

using namespace ptx;

int make_tmap_2d(CUtensorMap* map, const void* base, int dtype, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                 int box_outer);

static constexpr int kGlThreads = 64 + 256;
static constexpr int kGlMaxStages = 6;
static constexpr int kGlMaxBlk = 6;  // fp32 blocks (32 columns) per epilogue warp: N / 64

struct GemmLnP {
  const float *bias, *gamma, *beta;
  float *mean, *rstd;
  int M, N, kb, stages;
  float eps;
  uint32_t thr;  // dropout threshold (0 = off)
  float inv_keep;
  uint32_t site;
  uint64_t seed;
  const uint64_t* seed_dev;
};

__device__ __forceinline__ uint32_t gl_pack2(uint32_t a, uint32_t b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(a), __uint_as_float(b));
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int TMEM_COLS>
__global__ void __launch_bounds__(kGlThreads, 1)
gemm_res_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_x,
                   const __grid_constant__ CUtensorMap map_y, GemmLnP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int N = p.N;
  const int stage_bytes = 16384 + N * 128;
  uint8_t* ring = smem;
  float* bias_s = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes);  // [N]
  float* gamma_s = bias_s + N;
  float* beta_s = gamma_s + N;
  float* red_s = beta_s + N;  // [2 passes][2 halves][128 rows]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red_s + 512);
  uint64_t* empty_bar = full_bar + kGlMaxStages;
  uint64_t* acc_bar = empty_bar + kGlMaxStages;
  uint64_t* res_bar = acc_bar + 1;  // [8 warps][kGlMaxBlk]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8 * kGlMaxBlk);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int nmma = N > 256 ? N / 2 : N;  // columns per MMA instruction
  const int nh = N / nmma;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_r);
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_y);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    for (int i = 0; i < 8 * kGlMaxBlk; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_grid_sync();  // barriers and TMEM were set up while the previous kernel drained; global memory from here on

  if (warp == 0) {
    // ------------------------------ TMA producer ---------------------------
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < p.kb; ++kb) {
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* sa = ring + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], stage_bytes);
        tma_load_2d(sa, &map_a, &full_bar[s], kb * 64, m0);
        for (int h = 0; h < nh; ++h) tma_load_2d(sa + 16384 + h * nmma * 128, &map_b, &full_bar[s], kb * 64, h * nmma);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer -----------------------------
    const uint32_t idesc = make_idesc_bf16(128, nmma, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < p.kb; ++kb) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
      const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
      const uint64_t db = make_smem_desc_sw128(sa + 16384, 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          for (int h = 0; h < nh; ++h)  // (descriptor start addresses count 16-byte units: a half is nmma * 8 of them)
            tc_mma_bf16(tmem + (uint32_t)(h * nmma), da + (uint64_t)(k * 2), db + (uint64_t)(h * nmma * 8 + k * 2), idesc,
                        (kb | k) ? 1u : 0u);
        tc_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    if (elect_one()) tc_commit(acc_bar);
    __syncwarp();
  } else {
    // ------------------------------ epilogue -------------------------------
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    const int et = threadIdx.x - 64;
    for (int i = et; i < N; i += 256) {
      bias_s[i] = p.bias ? p.bias[i] : 0.f;
      gamma_s[i] = p.gamma ? p.gamma[i] : 1.f;  // gamma == NULL: y is the plain bf16 copy of x_out (no statistics)
      beta_s[i] = p.gamma ? p.beta[i] : 0.f;
    }
    uint64_t seed = p.seed;
    if (p.thr && p.seed_dev) seed += *p.seed_dev;
    const int nblk = N >> 6;  // 32-column fp32 blocks of this warp's half row
    const int col_beg = half * (N >> 1);
    const int row = quad * 32 + lane, m = m0 + row, row7 = lane & 7;
    const uint32_t row_addr = tmem + ((uint32_t)(quad * 32) << 16);
    uint8_t* my_stage = ring + (size_t)(ew * nblk) * 4096;
    uint64_t* my_res = res_bar + ew * kGlMaxBlk;
    mbar_wait(acc_bar, 0);  // every MMA has retired: the accumulator is complete and the ring is free
    tc_fence_after();
    if (lane == 0) {
      for (int b = 0; b < nblk; ++b) {
        mbar_expect_tx(&my_res[b], 4096);
        tma_load_2d(my_stage + b * 4096, &map_r, &my_res[b], col_beg + b * 32, m0 + quad * 32);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // bias / gamma / beta staged
    // ---- pass 1: x_out = dropout(acc + bias) + residual ----
    float sum = 0.f;
    for (int b = 0; b < nblk; ++b) {
      const int n = col_beg + b * 32;
      uint8_t* tile = my_stage + b * 4096;
      uint32_t r[32];
      tmem_ld32(row_addr + n, r);
      mbar_wait(&my_res[b], 0);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(bias_s + n + j);
        r[j] = __float_as_uint(__uint_as_float(r[j]) + bb.x);
        r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + bb.y);
        r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + bb.z);
        r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + bb.w);
      }
      if (p.thr) {  // the 32 columns are one mask group (N % 32 == 0): one hash, then a multiply-add per element
        const DropGroup g = dropout_group(seed, p.site, ((uint64_t)m * (uint64_t)N + (uint64_t)n) >> 5);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = dropout_word(g, j) >= p.thr ? __float_as_uint(__uint_as_float(r[j]) * p.inv_keep) : 0u;
      }
      const uint8_t* rr = tile + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(rr + ((j ^ row7) << 4));
        r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + x.x);
        r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + x.y);
        r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + x.z);
        r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + x.w);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        sum += (__uint_as_float(r[j]) + __uint_as_float(r[j + 1])) + (__uint_as_float(r[j + 2]) + __uint_as_float(r[j + 3]));
      tmem_st32(row_addr + n, r);
      uint8_t* wr = tile + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(wr + ((j ^ row7) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&map_x, tile, n, m0 + quad * 32);
        bulk_commit();
      }
    }
    tmem_st_wait();
    red_s[half * 128 + row] = sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float mean = p.gamma ? (red_s[row] + red_s[128 + row]) / (float)N : 0.f;
    // ---- pass 2: variance around the mean ----
    float ss = 0.f;
    for (int b = 0; b < nblk && p.gamma; ++b) {
      uint32_t r[32];
      tmem_ld32(row_addr + col_beg + b * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float d0 = __uint_as_float(r[j]) - mean, d1 = __uint_as_float(r[j + 1]) - mean;
        const float d2 = __uint_as_float(r[j + 2]) - mean, d3 = __uint_as_float(r[j + 3]) - mean;
        ss += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    }
    red_s[256 + half * 128 + row] = ss;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float rstd = p.gamma ? rsqrtf((red_s[256 + row] + red_s[384 + row]) / (float)N + p.eps) : 1.f;
    if (half == 0 && m < p.M && p.gamma) {
      p.mean[m] = mean;
      p.rstd[m] = rstd;
    }
    // ---- pass 3: y = (x_out - mean) * rstd * gamma + beta, bf16, 64 columns per staging block ----
    if (lane == 0) bulk_wait_read<0>();  // this warp's x_out stores have read their staging blocks
    __syncwarp();
    for (int b = 0; b < (nblk >> 1); ++b) {
      const int n = col_beg + b * 64;
      uint8_t* tile = my_stage + b * 4096;
      uint32_t r0[32], r1[32];
      tmem_ld32(row_addr + n, r0);
      tmem_ld32(row_addr + n + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {  // (float4 broadcast reads: a quarter of the shared-memory wavefronts of scalar ones)
        const float4 g0 = *reinterpret_cast<const float4*>(gamma_s + n + j), b0 = *reinterpret_cast<const float4*>(beta_s + n + j);
        const float4 g1 = *reinterpret_cast<const float4*>(gamma_s + n + 32 + j), b1 = *reinterpret_cast<const float4*>(beta_s + n + 32 + j);
        r0[j] = __float_as_uint((__uint_as_float(r0[j]) - mean) * rstd * g0.x + b0.x);
        r0[j + 1] = __float_as_uint((__uint_as_float(r0[j + 1]) - mean) * rstd * g0.y + b0.y);
        r0[j + 2] = __float_as_uint((__uint_as_float(r0[j + 2]) - mean) * rstd * g0.z + b0.z);
        r0[j + 3] = __float_as_uint((__uint_as_float(r0[j + 3]) - mean) * rstd * g0.w + b0.w);
        r1[j] = __float_as_uint((__uint_as_float(r1[j]) - mean) * rstd * g1.x + b1.x);
        r1[j + 1] = __float_as_uint((__uint_as_float(r1[j + 1]) - mean) * rstd * g1.y + b1.y);
        r1[j + 2] = __float_as_uint((__uint_as_float(r1[j + 2]) - mean) * rstd * g1.z + b1.z);
        r1[j + 3] = __float_as_uint((__uint_as_float(r1[j + 3]) - mean) * rstd * g1.w + b1.w);
      }
      uint8_t* wr = tile + lane * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 w;
        w.x = gl_pack2(r0[8 * j], r0[8 * j + 1]); w.y = gl_pack2(r0[8 * j + 2], r0[8 * j + 3]);
        w.z = gl_pack2(r0[8 * j + 4], r0[8 * j + 5]); w.w = gl_pack2(r0[8 * j + 6], r0[8 * j + 7]);
        *reinterpret_cast<uint4*>(wr + ((j ^ row7) << 4)) = w;
        w.x = gl_pack2(r1[8 * j], r1[8 * j + 1]); w.y = gl_pack2(r1[8 * j + 2], r1[8 * j + 3]);
        w.z = gl_pack2(r1[8 * j + 4], r1[8 * j + 5]); w.w = gl_pack2(r1[8 * j + 6], r1[8 * j + 7]);
        *reinterpret_cast<uint4*>(wr + (((4 + j) ^ row7) << 4)) = w;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&map_y, tile, n, m0 + quad * 32);
        bulk_commit();
      }
    }
    if (lane == 0) bulk_wait<0>();  // every output block has landed before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

static size_t gl_fixed_bytes(int N) { return (size_t)3 * N * 4 + 512 * 4 + (2 * kGlMaxStages + 1 + 8 * kGlMaxBlk) * 8 + 16; }

// ring stages: as many as fit, and at least the CTA's fp32 residual tile (128 x N x 4 bytes) of staging space
static int gl_stages(int N) {
  const size_t stage = 16384 + (size_t)N * 128, room = 227 * 1024 - gl_fixed_bytes(N);
  int s = (int)(room / stage);
  return s > kGlMaxStages ? kGlMaxStages : s;
}

bool gemm_res_ln_supported(int N, int K) {
  if (!(N == 128 || N == 256 || N == 384) || K < 64 || K % 64 != 0) return false;
  const int s = gl_stages(N);
  return s >= 2 && (size_t)s * (16384 + (size_t)N * 128) >= (size_t)512 * N;
}

}  // namespace dgpt

using namespace dgpt;

extern "C" {

int dgpt_gemm_res_ln_supported(int N, int K) { return gemm_res_ln_supported(N, K) ? 1 : 0; }

int dgpt_gemm_res_ln(const void* a, int lda, const void* w, int ldw, const float* bias, const float* residual, int ldr,
                     float* x_out, int ldx, const float* gamma, const float* beta, void* y, int ldy, float* mean,
                     float* rstd, int M, int N, int K, float eps, float dropout_p, uint64_t seed,
                     const uint64_t* seed_dev, uint32_t site, void* stream) {
  DGPT_DEVICE_OR_RETURN();
  DGPT_REQUIRE(M >= 0 && a && w && residual && x_out && y && (!gamma || (beta && mean && rstd)), "gemm_res_ln: bad arguments");
  if (M == 0) return DGPT_OK;
  DGPT_REQUIRE(gemm_res_ln_supported(N, K), "gemm_res_ln: needs N in {128, 256, 384} and K %% 64 == 0 (N=%d K=%d)", N, K);
  DGPT_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gemm_res_ln: dropout_p %g outside [0, 1)", (double)dropout_p);
  CUtensorMap ma, mb, mr, mx, my;
  int rc;
  const int nmma = N > 256 ? N / 2 : N;
  if ((rc = make_tmap_2d(&ma, a, DGPT_BF16, K, M, lda, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&mb, w, DGPT_BF16, K, N, ldw, 64, nmma))) return rc;
  if ((rc = make_tmap_2d(&mr, residual, DGPT_F32, N, M, ldr, 32, 32))) return rc;
  if ((rc = make_tmap_2d(&mx, x_out, DGPT_F32, N, M, ldx, 32, 32))) return rc;
  if ((rc = make_tmap_2d(&my, y, DGPT_BF16, N, M, ldy, 64, 32))) return rc;
  GemmLnP p;
  p.bias = bias; p.gamma = gamma; p.beta = beta; p.mean = mean; p.rstd = rstd;
  p.M = M; p.N = N; p.kb = K / 64; p.stages = gl_stages(N); p.eps = eps;
  p.thr = dropout_threshold(dropout_p);
  p.inv_keep = 1.f / (1.f - dropout_p);
  p.site = site; p.seed = seed; p.seed_dev = seed_dev;
  const size_t smem = (size_t)p.stages * (16384 + (size_t)N * 128) + gl_fixed_bytes(N);
  // (the three instantiations share one function-pointer type: the attribute flag is per kernel, not per lambda)
  auto launch = [&](void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, GemmLnP), bool& attr_done) -> int {
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
      if (e != cudaSuccess) { set_error("gemm_res_ln: smem attribute: %s", cudaGetErrorString(e)); return DGPT_E_LAUNCH; }
      attr_done = true;
    }
    launch_pdl(kern, dim3(ceil_div(M, 128)), dim3(kGlThreads), smem, (cudaStream_t)stream, ma, mb, mr, mx, my, p);
    return check_launch("gemm_res_ln");
  };
  static bool done128 = false, done256 = false, done512 = false;
  if (N <= 128) return launch(gemm_res_ln_kernel<128>, done128);
  if (N <= 256) return launch(gemm_res_ln_kernel<256>, done256);
  return launch(gemm_res_ln_kernel<512>, done512);
}

}  // extern "C"
