// Exact-mode fused causal attention (CUDA cores, fp32 math), forward and backward.
// One CTA per (batch, head) slice (forward additionally splits the query rows);
// K/V live in shared memory, one warp owns one query row at a time, the T x T
// probabilities never reach HBM in the forward pass.  Handles Tq != Tk (KV-cached
// decode: the Tq queries are the LAST Tq positions of the Tk keys).
#include "common.cuh"

namespace dgpt {

struct AttnP {
  const void *q, *k, *v, *d_o;
  void *o, *dq, *dk, *dv;
  float* lse;
  float* scratch;
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
  int64_t dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs, do_bs, do_rs;
  int B, NH, H, Tq, Tk;
  float scale, inv_keep;
  uint32_t thr, site;
  uint64_t seed;
  const uint64_t* seed_dev;
};

static constexpr int kWarps = 8;
static constexpr int kRowsPerCta = 64;  // forward query rows per CTA

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) attn_fwd_simt_kernel(AttnP p) {
  extern __shared__ float smem[];
  if (p.thr && p.seed_dev) p.seed += *p.seed_dev;
  const int H = p.H, Tk = p.Tk, Tq = p.Tq, HP = H + 1;
  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int off = Tk - Tq;
  const int r0 = blockIdx.y * kRowsPerCta;
  const int r1 = min(Tq, r0 + kRowsPerCta);
  const int nk = min(Tk, r1 + off);  // keys this CTA can see
  float* Ks = smem;                   // [Tk][H+1]
  float* Vs = Ks + (size_t)Tk * HP;   // [Tk][H]
  float* qs = Vs + (size_t)Tk * H;    // [kWarps][H]
  float* ps = qs + kWarps * H;        // [kWarps][Tk]
  const T* kb = (const T*)p.k + b * p.k_bs + h * H;
  const T* vb = (const T*)p.v + b * p.v_bs + h * H;
  for (int i = threadIdx.x; i < nk * H; i += blockDim.x) {
    const int j = i / H, d = i - j * H;
    Ks[j * HP + d] = to_f32(kb[j * p.k_rs + d]);
    Vs[j * H + d] = to_f32(vb[j * p.v_rs + d]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* q = qs + w * H;
  float* pr = ps + (size_t)w * Tk;
  for (int i = r0 + w; i < r1; i += kWarps) {
    const T* qr = (const T*)p.q + b * p.q_bs + i * p.q_rs + h * H;
    for (int d = lane; d < H; d += 32) q[d] = to_f32(qr[d]);
    __syncwarp();
    const int n = i + off + 1;
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) {
      float s = 0.f;
      for (int d = 0; d < H; ++d) s = fmaf(q[d], Ks[j * HP + d], s);
      s *= p.scale;
      pr[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float e = expf(pr[j] - mx);
      pr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    if (lane == 0 && p.lse) p.lse[((int64_t)b * p.NH + h) * Tq + i] = mx + logf(sum);
    const uint64_t base = (((uint64_t)b * p.NH + h) * Tq + i) * (uint64_t)Tk;
    for (int j = lane; j < n; j += 32) {
      float pv = pr[j] * inv;
      if (p.thr) pv = dropout_keep(p.seed, p.site, base + j, p.thr) ? pv * p.inv_keep : 0.f;
      pr[j] = pv;
    }
    __syncwarp();
    T* orow = (T*)p.o + b * p.o_bs + i * p.o_rs + h * H;
    for (int d = lane; d < H; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(pr[j], Vs[j * H + d], acc);
      orow[d] = from_f32<T>(acc);
    }
    __syncwarp();
  }
}

// Backward.  Phase 1 (row-parallel): recompute P from q,k and the saved lse,
// dPd = dO V^T, D = sum_j Pd*dPd, dS = P*(dP-D)*scale, dQ = dS K.  dS and Pd rows
// go to a global scratch so that phase 2 (column-parallel) can form
// dK = dS^T Q and dV = Pd^T dO without atomics.
template <typename T>
__global__ void __launch_bounds__(kWarps * 32) attn_bwd_simt_kernel(AttnP p) {
  extern __shared__ float smem[];
  if (p.thr && p.seed_dev) p.seed += *p.seed_dev;
  const int H = p.H, Tk = p.Tk, Tq = p.Tq, HP = H + 1;
  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int off = Tk - Tq;
  float* Ks = smem;                      // [Tk][H+1]   (phase 2: Q  [Tq][H])
  float* Vs = Ks + (size_t)Tk * HP;      // [Tk][H+1]   (phase 2: dO [Tq][H])
  float* qs = Vs + (size_t)Tk * HP;      // [kWarps][H]
  float* gs = qs + kWarps * H;           // [kWarps][H]  dO row
  float* ps = gs + kWarps * H;           // [kWarps][Tk] p, then dS
  float* ds = ps + (size_t)kWarps * Tk;  // [kWarps][Tk] dP
  float* dS_g = p.scratch + (int64_t)bh * Tq * Tk * 2;
  float* Pd_g = dS_g + (int64_t)Tq * Tk;
  const T* kb = (const T*)p.k + b * p.k_bs + h * H;
  const T* vb = (const T*)p.v + b * p.v_bs + h * H;
  for (int i = threadIdx.x; i < Tk * H; i += blockDim.x) {
    const int j = i / H, d = i - j * H;
    Ks[j * HP + d] = to_f32(kb[j * p.k_rs + d]);
    Vs[j * HP + d] = to_f32(vb[j * p.v_rs + d]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* q = qs + w * H;
  float* g = gs + w * H;
  float* pr = ps + (size_t)w * Tk;
  float* dr = ds + (size_t)w * Tk;
  for (int i = w; i < Tq; i += kWarps) {
    const T* qr = (const T*)p.q + b * p.q_bs + i * p.q_rs + h * H;
    const T* gr = (const T*)p.d_o + b * p.do_bs + i * p.do_rs + h * H;
    for (int d = lane; d < H; d += 32) {
      q[d] = to_f32(qr[d]);
      g[d] = to_f32(gr[d]);
    }
    __syncwarp();
    const int n = i + off + 1;
    const float lse = p.lse[((int64_t)b * p.NH + h) * Tq + i];
    const uint64_t base = (((uint64_t)b * p.NH + h) * Tq + i) * (uint64_t)Tk;
    float D = 0.f;
    for (int j = lane; j < n; j += 32) {
      float s = 0.f, dpd = 0.f;
      for (int d = 0; d < H; ++d) {
        s = fmaf(q[d], Ks[j * HP + d], s);
        dpd = fmaf(g[d], Vs[j * HP + d], dpd);
      }
      const float pj = expf(s * p.scale - lse);
      const bool keep = p.thr == 0 || dropout_keep(p.seed, p.site, base + j, p.thr);
      const float pd = keep ? pj * p.inv_keep : 0.f;
      const float dp = keep ? dpd * p.inv_keep : 0.f;
      D = fmaf(pj, dp, D);
      pr[j] = pj;
      dr[j] = dp;
      Pd_g[(int64_t)i * Tk + j] = pd;
    }
    D = warp_sum(D);
    for (int j = lane; j < n; j += 32) {
      const float v = pr[j] * (dr[j] - D) * p.scale;
      pr[j] = v;
      dS_g[(int64_t)i * Tk + j] = v;
    }
    __syncwarp();
    T* dqr = (T*)p.dq + b * p.dq_bs + i * p.dq_rs + h * H;
    for (int d = lane; d < H; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(pr[j], Ks[j * HP + d], acc);
      dqr[d] = from_f32<T>(acc);
    }
    __syncwarp();
  }
  __syncthreads();
  // phase 2: Q and dO into shared memory (unpadded), one warp per key column
  float* Qs = Ks;
  float* Gs = Vs;
  for (int i = threadIdx.x; i < Tq * H; i += blockDim.x) {
    const int r = i / H, d = i - r * H;
    Qs[r * H + d] = to_f32(((const T*)p.q + b * p.q_bs + r * p.q_rs + h * H)[d]);
    Gs[r * H + d] = to_f32(((const T*)p.d_o + b * p.do_bs + r * p.do_rs + h * H)[d]);
  }
  __syncthreads();
  for (int j = w; j < Tk; j += kWarps) {
    const int i0 = max(j - off, 0);
    T* dkr = (T*)p.dk + b * p.dk_bs + j * p.dk_rs + h * H;
    T* dvr = (T*)p.dv + b * p.dv_bs + j * p.dv_rs + h * H;
    for (int d = lane; d < H; d += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = i0; i < Tq; ++i) {
        ak = fmaf(dS_g[(int64_t)i * Tk + j], Qs[i * H + d], ak);
        av = fmaf(Pd_g[(int64_t)i * Tk + j], Gs[i * H + d], av);
      }
      dkr[d] = from_f32<T>(ak);
      dvr[d] = from_f32<T>(av);
    }
  }
}

// KV-cached decode step (Tq == 1, no dropout): one warp per (batch, head) streams that head's K and V
// rows straight from the cache (each row is one contiguous 2*H- or 4*H-byte segment), scores and
// probabilities stay in shared memory, nothing is staged through a per-CTA copy of the whole cache.
// HBM-bound: 2 * Tk * H elements read per (b, h).
static constexpr int kDecWarps = 4;

template <typename T>
__global__ void __launch_bounds__(kDecWarps * 32) attn_decode_kernel(AttnP p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int bh = blockIdx.x * kDecWarps + w;
  if (bh >= p.B * p.NH) return;
  const int b = bh / p.NH, h = bh % p.NH, H = p.H, Tk = p.Tk;
  float* qs = smem + w * (H + Tk);  // [H] query, then [Tk] scores / probabilities
  float* ps = qs + H;
  const T* qr = (const T*)p.q + b * p.q_bs + h * H;  // the single query row (t = 0 of the q view)
  for (int d = lane; d < H; d += 32) qs[d] = to_f32(qr[d]) * p.scale;
  __syncwarp();
  const T* kb = (const T*)p.k + b * p.k_bs + h * H;
  const T* vb = (const T*)p.v + b * p.v_bs + h * H;
  float mx = -INFINITY;
  for (int j = lane; j < Tk; j += 32) {
    const T* kr = kb + j * p.k_rs;
    float sacc = 0.f;
    if constexpr (sizeof(T) == 2) {
      if ((H & 7) == 0) {
        for (int d = 0; d < H; d += 8) {
          const uint4 u = *reinterpret_cast<const uint4*>(kr + d);
          const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uw[t]));
            sacc = fmaf(qs[d + 2 * t], f.x, sacc);
            sacc = fmaf(qs[d + 2 * t + 1], f.y, sacc);
          }
        }
      } else {
        for (int d = 0; d < H; ++d) sacc = fmaf(qs[d], to_f32(kr[d]), sacc);
      }
    } else {
      for (int d = 0; d < H; ++d) sacc = fmaf(qs[d], to_f32(kr[d]), sacc);
    }
    ps[j] = sacc;
    mx = fmaxf(mx, sacc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < Tk; j += 32) {
    const float e = expf(ps[j] - mx);
    ps[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  T* orow = (T*)p.o + b * p.o_bs + h * H;
  for (int d = lane; d < H; d += 32) {  // lanes walk the head dimension: each V row is read coalesced
    float acc = 0.f;
    for (int j = 0; j < Tk; ++j) acc = fmaf(ps[j], to_f32(vb[j * p.v_rs + d]), acc);
    orow[d] = from_f32<T>(acc * inv);
  }
}

static AttnP make_params(const dgpt_attn_args* a) {
  AttnP p;
  p.q = a->q; p.k = a->k; p.v = a->v; p.d_o = a->d_o; p.o = a->o;
  p.dq = a->dq; p.dk = a->dk; p.dv = a->dv; p.lse = a->lse; p.scratch = (float*)a->scratch;
  p.q_bs = a->q_bs; p.q_rs = a->q_rs; p.k_bs = a->k_bs; p.k_rs = a->k_rs;
  p.v_bs = a->v_bs; p.v_rs = a->v_rs; p.o_bs = a->o_bs; p.o_rs = a->o_rs;
  p.dq_bs = a->dq_bs; p.dq_rs = a->dq_rs; p.dk_bs = a->dk_bs; p.dk_rs = a->dk_rs;
  p.dv_bs = a->dv_bs; p.dv_rs = a->dv_rs; p.do_bs = a->do_bs; p.do_rs = a->do_rs;
  p.B = a->B; p.NH = a->NH; p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk;
  p.scale = a->scale; p.inv_keep = 1.f / (1.f - a->dropout_p);
  p.thr = dropout_threshold(a->dropout_p); p.site = a->site; p.seed = a->seed; p.seed_dev = a->seed_dev;
  return p;
}

static constexpr size_t kMaxSmem = 227 * 1024;

int launch_attn_fwd_simt(const dgpt_attn_args* a, cudaStream_t st) {
  if (a->Tq == 1 && a->dropout_p == 0.f && a->lse == nullptr) {  // KV-cached decode step
    AttnP p = make_params(a);
    const size_t smem = (size_t)kDecWarps * (a->H + a->Tk) * sizeof(float);
    DGPT_REQUIRE(smem <= 48 * 1024, "attn_fwd(decode): Tk=%d H=%d needs %zu B of shared memory", a->Tk, a->H, smem);
    const int grid = ceil_div((int64_t)a->B * a->NH, kDecWarps);
    if (a->dtype == DGPT_F32) attn_decode_kernel<float><<<grid, kDecWarps * 32, smem, st>>>(p);
    else attn_decode_kernel<__nv_bfloat16><<<grid, kDecWarps * 32, smem, st>>>(p);
    return check_launch("attn_decode");
  }
  const size_t smem = ((size_t)a->Tk * (a->H + 1) + (size_t)a->Tk * a->H + kWarps * a->H +
                       (size_t)kWarps * a->Tk) * sizeof(float);
  DGPT_REQUIRE(smem <= kMaxSmem, "attn_fwd(exact): Tk=%d H=%d needs %zu B of shared memory (max %zu)",
               a->Tk, a->H, smem, kMaxSmem);
  AttnP p = make_params(a);
  dim3 grid(a->B * a->NH, ceil_div(a->Tq, kRowsPerCta));
  if (a->dtype == DGPT_F32) {
    cudaFuncSetAttribute(attn_fwd_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    attn_fwd_simt_kernel<float><<<grid, kWarps * 32, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(attn_fwd_simt_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    attn_fwd_simt_kernel<__nv_bfloat16><<<grid, kWarps * 32, smem, st>>>(p);
  }
  return check_launch("attn_fwd_simt");
}

int launch_attn_bwd_simt(const dgpt_attn_args* a, cudaStream_t st) {
  const size_t smem = (2 * (size_t)a->Tk * (a->H + 1) + 2 * kWarps * a->H + 2 * (size_t)kWarps * a->Tk) * sizeof(float);
  DGPT_REQUIRE(smem <= kMaxSmem && a->Tq <= a->Tk,
               "attn_bwd(exact): Tq=%d Tk=%d H=%d needs %zu B of shared memory (max %zu)", a->Tq, a->Tk,
               a->H, smem, kMaxSmem);
  DGPT_REQUIRE(a->scratch != nullptr, "attn_bwd(exact): scratch is NULL");
  AttnP p = make_params(a);
  dim3 grid(a->B * a->NH);
  if (a->dtype == DGPT_F32) {
    cudaFuncSetAttribute(attn_bwd_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    attn_bwd_simt_kernel<float><<<grid, kWarps * 32, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(attn_bwd_simt_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    attn_bwd_simt_kernel<__nv_bfloat16><<<grid, kWarps * 32, smem, st>>>(p);
  }
  return check_launch("attn_bwd_simt");
}

}  // namespace dgpt
