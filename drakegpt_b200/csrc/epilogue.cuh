// GEMM epilogue shared by the fp32 CUDA-core kernel and the tcgen05 kernel:
// bias, ReLU, ReLU-mask from a saved activation, counter-based dropout,
// residual add, accumulate, and up to two typed outputs.
#pragma once
#include "common.cuh"

namespace dgpt {

struct Epilogue {
  void* D;
  void* D2;
  const float* bias;
  const float* residual;
  const void* relu_aux;
  int M, N;
  int ldd, ldd2, ldr, ld_aux;
  int d_dtype, d2_dtype, aux_dtype;
  int relu, accumulate, atomic;  // atomic: split-K partial -> red.add into fp32 D
  int first_split;               // bias/residual are applied by split 0 only
  uint32_t thr;                  // dropout threshold (0 = off)
  float inv_keep;
  uint32_t site;
  uint64_t seed;
  const uint64_t* seed_dev;
};

inline Epilogue make_epilogue(const dgpt_gemm_args* a) {
  Epilogue e;
  e.D = a->D; e.D2 = a->D2; e.bias = a->bias; e.residual = a->residual; e.relu_aux = a->relu_aux;
  e.M = a->M; e.N = a->N;
  e.ldd = a->ldd; e.ldd2 = a->ldd2; e.ldr = a->ldr; e.ld_aux = a->ld_aux;
  e.d_dtype = a->d_dtype; e.d2_dtype = a->d2_dtype; e.aux_dtype = a->aux_dtype;
  e.relu = a->relu; e.accumulate = a->accumulate; e.atomic = 0; e.first_split = 1;
  e.thr = dropout_threshold(a->dropout_p);
  e.inv_keep = 1.f / (1.f - a->dropout_p);
  e.site = a->site; e.seed = a->seed; e.seed_dev = a->seed_dev;
  return e;
}

// fold the device-side seed offset in once per thread (graph replays bump it)
__device__ __forceinline__ void epilogue_resolve_seed(Epilogue& e) {
  if (e.thr && e.seed_dev) e.seed += *e.seed_dev;
  e.seed_dev = nullptr;
}

__device__ __forceinline__ float epilogue_value(const Epilogue& e, int m, int n, float v) {
  if (e.first_split && e.bias) v += e.bias[n];
  if (e.relu) v = fmaxf(v, 0.f);
  if (e.relu_aux) {
    const int64_t i = (int64_t)m * e.ld_aux + n;
    const float a = e.aux_dtype == DGPT_F32 ? reinterpret_cast<const float*>(e.relu_aux)[i]
                                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.relu_aux)[i]);
    v = a > 0.f ? v : 0.f;
  }
  if (e.thr) v = dropout_keep(e.seed, e.site, (uint64_t)m * (uint64_t)e.N + (uint64_t)n, e.thr) ? v * e.inv_keep : 0.f;
  if (e.first_split && e.residual) v += e.residual[(int64_t)m * e.ldr + n];
  return v;
}

__device__ __forceinline__ void epilogue_store(const Epilogue& e, int m, int n, float v) {
  const int64_t i = (int64_t)m * e.ldd + n;
  if (e.d_dtype == DGPT_F32) {
    float* d = reinterpret_cast<float*>(e.D);
    if (e.atomic) atomicAdd(d + i, v);
    else { if (e.accumulate) v += d[i]; d[i] = v; }
  } else {
    reinterpret_cast<__nv_bfloat16*>(e.D)[i] = __float2bfloat16_rn(v);
  }
  if (e.D2) {
    const int64_t j = (int64_t)m * e.ldd2 + n;
    if (e.d2_dtype == DGPT_F32) reinterpret_cast<float*>(e.D2)[j] = v;
    else reinterpret_cast<__nv_bfloat16*>(e.D2)[j] = __float2bfloat16_rn(v);
  }
}

}  // namespace dgpt
