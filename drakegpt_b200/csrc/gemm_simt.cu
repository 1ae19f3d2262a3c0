// Exact-mode GEMM: fp32 operands, fp32 FMA accumulation on the CUDA cores, any
// shape, either operand K-major or MN-major.  Used for the small reference
// models (C=32, T=8, whose tiles do not fill a tcgen05 instruction) and as the
// on-device fp32 cross-check of the tensor-core path.  64x64x16 tiles, 256
// threads, 4x4 outputs per thread, shared-memory staged.
#include "epilogue.cuh"

namespace dgpt {

static constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A,
                                                       const float* __restrict__ B, int M, int N, int K,
                                                       int64_t a_sm, int64_t a_sk, int64_t b_sn,
                                                       int64_t b_sk, Epilogue ep) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  epilogue_resolve_seed(ep);
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // stage A tile: pick the thread mapping that walks the contiguous dimension
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm, kk;
      if (a_sk == 1) { kk = tid & 15; mm = (tid >> 4) + 16 * i; }
      else           { mm = tid & 63; kk = (tid >> 6) + 4 * i; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[gm * a_sm + gk * a_sk] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int nn, kk;
      if (b_sk == 1) { kk = tid & 15; nn = (tid >> 4) + 16 * i; }
      else           { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? B[gn * b_sn + gk * b_sk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) epilogue_store(ep, m, n, epilogue_value(ep, m, n, acc[i][j]));
    }
  }
}

int launch_gemm_f32(const dgpt_gemm_args* a, cudaStream_t st) {
  const int64_t a_sm = a->a_major == DGPT_MAJOR_K ? a->lda : 1, a_sk = a->a_major == DGPT_MAJOR_K ? 1 : a->lda;
  const int64_t b_sn = a->b_major == DGPT_MAJOR_K ? a->ldb : 1, b_sk = a->b_major == DGPT_MAJOR_K ? 1 : a->ldb;
  Epilogue ep = make_epilogue(a);
  dim3 grid(ceil_div(a->N, BN), ceil_div(a->M, BM));
  gemm_f32_kernel<<<grid, 256, 0, st>>>((const float*)a->A, (const float*)a->B, a->M, a->N, a->K, a_sm,
                                        a_sk, b_sn, b_sk, ep);
  return check_launch("gemm_f32");
}

}  // namespace dgpt
