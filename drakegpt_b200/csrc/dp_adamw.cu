// Data-parallel optimizer step as ONE kernel over NVLink peer memory (sm_100a, one process per GPU):
//
//     reduce-scatter of the gradient arenas  ->  AdamW on this rank's 1/N shard  ->  all-gather of the updated
//     fp32 parameters and their bf16 shadows
//
// Every rank maps the gradient / parameter / shadow arenas of all its peers (CUDA IPC, exchanged once on the host) and
// owns the contiguous shard [lo, hi) of the flat arena.  Per step:
//   barrier A   every rank has finished its backward pass (remote flag stores + local spin; one flag per peer)
//   phase 1     g_sum[i] = sum over ranks r of g_r[i] for i in my shard: the local slice plus N - 1 peer slices read
//               straight over NVLink with 16-byte loads, all N loads of an element group in flight at once;
//               AdamW (decoupled decay, bias correction from the device-side step counter, 1 / world folded into
//               grad_scale) on the shard only -- the Adam moments exist only for the shard (1 / N of the memory and
//               1 / N of the optimizer's HBM traffic per rank instead of a full replica on every rank)
//   phase 2     the updated fp32 parameters and bf16 shadows of the shard are stored into every rank's arenas
//               (remote 16-byte stores), i.e. the all-gather is the optimizer's own write-back
//   barrier B   all my stores have been fenced (system scope) and every peer has finished reading my gradients and
//               writing into my arenas: the kernel may retire; the caller clears the gradient arena afterwards.
// Against NCCL all-reduce (2 (N-1)/N x 43 MB each way) + a replicated AdamW this moves (N-1)/N x 43 MB in and
// (N-1)/N x 65 MB out per rank over NVLink with no intermediate buffers, and the optimizer math rides on the transfer.
//
// Replaces: the gradient mean that torch DDP would add around src/train.py:149-151 (the reference is single-process;
// SURVEY 8e) + optimizer.zero_grad() / AdamW.step().
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace dgpt {

static constexpr int kMaxRanks = 8;

struct DpPeers {
  const float* g[kMaxRanks];      // gradient arena of every rank (index = rank; [me] is local)
  float* p[kMaxRanks];            // parameter arena of every rank
  __nv_bfloat16* sh[kMaxRanks];   // bf16 shadow arena of every rank (may be all NULL)
  uint32_t* flags[kMaxRanks];     // flag array of every rank: [2][kMaxRanks] uint32 (A and B barriers)
  const uint8_t* need32;          // per 64-element block of the arena: peers need the fp32 value (NULL: all blocks)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// non-temporal 16-byte peer loads: every byte is read exactly once
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// spin until every peer's flag for this barrier reaches `epoch`; returns false on timeout (a peer never arrived)
__device__ __forceinline__ bool wait_peers(const uint32_t* my_flags, int world, int me, uint32_t epoch, long long timeout) {
  const long long t0 = clock64();
  for (int r = 0; r < world; ++r) {
    if (r == me) continue;
    while ((int32_t)(ld_acquire_sys(my_flags + r) - epoch) < 0) {
      __nanosleep(64);
      if (clock64() - t0 > timeout) return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(512, 1)
dp_adamw_kernel(DpPeers pe, float* __restrict__ m, float* __restrict__ v, int64_t lo, int64_t hi, int world, int me,
                const float* __restrict__ hyper, const int64_t* __restrict__ step, const uint32_t* __restrict__ epoch_ptr,
                uint32_t* __restrict__ done_blocks, int* __restrict__ status) {
  __shared__ int s_ok;
  const uint32_t epoch = *epoch_ptr + 1u;
  const long long timeout = 4000000000ll;  // ~2 s of SM clocks: a peer that never launches must not hang the GPU
  // ---- barrier A: my backward pass is complete (stream order) -> tell every peer; wait for theirs ----
  if (blockIdx.x == 0 && threadIdx.x < world && (int)threadIdx.x != me)
    st_release_sys(pe.flags[threadIdx.x] + me, epoch);
  if (threadIdx.x == 0) s_ok = wait_peers(pe.flags[me], world, me, epoch, timeout) ? 1 : 0;
  __syncthreads();
  if (!s_ok) {
    if (threadIdx.x == 0) *status = 1;
    return;
  }
  // ---- phase 1 + 2: reduce my shard over all ranks, AdamW, write back to all ranks ----
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], gs = hyper[5];
  const double t = (double)(*step + 1);
  const float bc1 = (float)(1.0 - pow((double)b1, t)), bc2 = (float)(1.0 - pow((double)b2, t));
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  const int64_t nq = (hi - lo) >> 2;  // shard boundaries are multiples of 64 elements
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = lo + (q << 2);
    float4 acc[kMaxRanks];
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < world) acc[r] = ld_peer_f4(pe.g[r] + i);
    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < world) { gg.x += acc[r].x; gg.y += acc[r].y; gg.z += acc[r].z; gg.w += acc[r].w; }
    float4 pp = *reinterpret_cast<const float4*>(pe.p[me] + i);
    float4 mm = reinterpret_cast<float4*>(m)[q];
    float4 vv = reinterpret_cast<float4*>(v)[q];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = ga[j] * gs;
      pa[j] *= decay;
      ma[j] = ma[j] + (gj - ma[j]) * (1.f - b1);
      va[j] = va[j] * b2 + gj * gj * (1.f - b2);
      pa[j] -= step_size * ma[j] / (sqrtf(va[j]) * inv_sqrt_bc2 + eps);
    }
    reinterpret_cast<float4*>(m)[q] = mm;
    reinterpret_cast<float4*>(v)[q] = vv;
    __nv_bfloat162 l2 = __floats2bfloat162_rn(pp.x, pp.y), h2 = __floats2bfloat162_rn(pp.z, pp.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&l2);
    pk.y = *reinterpret_cast<uint32_t*>(&h2);
    // GEMM weights are only ever read through their bf16 shadows during training: their fp32 masters stay with the
    // owner (need32 = 0 for those blocks; PeerAdamW.sync_master() fetches them one-sidedly for checkpoints)
    const bool all32 = pe.need32 == nullptr || pe.need32[i >> 6] != 0;
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < world) {
        if (all32 || r == me) *reinterpret_cast<float4*>(pe.p[r] + i) = pp;
        if (pe.sh[r]) *reinterpret_cast<uint2*>(pe.sh[r] + i) = pk;
      }
  }
  // ---- barrier B: my remote stores are fenced; the LAST block of this rank tells every peer and waits for theirs ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(done_blocks, 1u);
    s_ok = (prev == gridDim.x - 1) ? 2 : 1;
  }
  __syncthreads();
  if (s_ok != 2) return;
  if (threadIdx.x == 0) *done_blocks = 0;  // ready for the next launch
  __threadfence_system();                   // (cumulativity: the other blocks' fenced stores before my flag stores)
  if (threadIdx.x < world && (int)threadIdx.x != me) st_release_sys(pe.flags[threadIdx.x] + kMaxRanks + me, epoch);
  if (threadIdx.x == 0) {
    if (!wait_peers(pe.flags[me] + kMaxRanks, world, me, epoch, timeout)) *status = 2;
  }
}

// epoch and AdamW step counters advance on the stream, after the kernel (graph-replay safe, no host state)
__global__ void dp_counters_kernel(uint32_t* epoch, int64_t* step) {
  *epoch += 1u;
  *step += 1;
}

}  // namespace dgpt

using namespace dgpt;

extern "C" {

// ---- CUDA IPC plumbing: export a device pointer of this process / map one exported by a peer process ----
// handle_out: 64 bytes (cudaIpcMemHandle_t) of the ALLOCATION containing ptr; *offset_out: ptr - allocation base.
int dgpt_ipc_export(const void* ptr, void* handle_out, int64_t* offset_out) {
  DGPT_DEVICE_OR_RETURN();
  DGPT_REQUIRE(ptr && handle_out && offset_out, "ipc_export: NULL argument");
  CUdeviceptr base = 0;
  size_t size = 0;
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      set_error("ipc_export: cuMemGetAddressRange entry point not available");
      return DGPT_E_DEVICE;
    }
    fn = (RangeFn)sym;
  }
  if (fn(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) {
    set_error("ipc_export: cuMemGetAddressRange failed");
    return DGPT_E_ARG;
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (void*)base);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("ipc_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return DGPT_E_DEVICE;
  }
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (int64_t)((CUdeviceptr)ptr - base);
  return DGPT_OK;
}

int dgpt_ipc_open(const void* handle, int64_t offset, void** ptr_out) {
  DGPT_DEVICE_OR_RETURN();
  DGPT_REQUIRE(handle && ptr_out, "ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("ipc_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    return DGPT_E_DEVICE;
  }
  *ptr_out = (char*)base + offset;
  return DGPT_OK;
}

/* bytes from a peer-mapped (or local) device buffer into a local one, on the stream (one-sided fetch over NVLink) */
int dgpt_peer_copy(void* dst, const void* src, int64_t bytes, void* stream) {
  DGPT_DEVICE_OR_RETURN();
  DGPT_REQUIRE(dst && src && bytes >= 0, "peer_copy: bad arguments");
  if (bytes == 0) return DGPT_OK;
  cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("peer_copy: %s", cudaGetErrorString(e));
    return DGPT_E_LAUNCH;
  }
  return DGPT_OK;
}

int dgpt_ipc_close(void* ptr, int64_t offset) {
  if (!ptr) return DGPT_OK;
  cudaError_t e = cudaIpcCloseMemHandle((char*)ptr - offset);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("ipc_close: %s", cudaGetErrorString(e));
    return DGPT_E_DEVICE;
  }
  return DGPT_OK;
}

// peers: host array of 4 * world + 1 device pointers: g[world], p[world], shadow[world] (entries may be NULL),
// flags[world], then need32 (uint8 per 64-element block of the arena, or NULL = broadcast every fp32 value);
// entry [me] of each group is this rank's own buffer.  m / v: this rank's moment shards ([hi - lo] floats).
// epoch (device uint32): barrier generation, bumped by this call (on the stream) after the kernel.
// scratch (device, 2 x uint32, zero-initialised once): [0] block counter, [1] status (0 ok, 1 / 2 = a peer never
// reached barrier A / B within the timeout).
int dgpt_dp_adamw(const void* const* peers, int world, int me, float* m, float* v, int64_t lo, int64_t hi,
                  const float* hyper, int64_t* step, uint32_t* epoch, uint32_t* scratch, int sms, void* stream) {
  DGPT_DEVICE_OR_RETURN();
  DGPT_REQUIRE(peers && world >= 1 && world <= kMaxRanks && me >= 0 && me < world, "dp_adamw: world=%d me=%d", world, me);
  DGPT_REQUIRE(lo >= 0 && hi >= lo && lo % 64 == 0 && hi % 64 == 0, "dp_adamw: shard [%lld, %lld) must be 64-element aligned",
               (long long)lo, (long long)hi);
  DGPT_REQUIRE(m && v && hyper && step && epoch && scratch, "dp_adamw: NULL argument");
  DpPeers pe;
  memset(&pe, 0, sizeof(pe));
  for (int r = 0; r < world; ++r) {
    pe.g[r] = (const float*)peers[r];
    pe.p[r] = (float*)peers[world + r];
    pe.sh[r] = (__nv_bfloat16*)peers[2 * world + r];
    pe.flags[r] = (uint32_t*)peers[3 * world + r];
    DGPT_REQUIRE(pe.g[r] && pe.p[r] && pe.flags[r], "dp_adamw: missing peer pointer for rank %d", r);
  }
  pe.need32 = (const uint8_t*)peers[4 * world];
  DGPT_REQUIRE(!pe.need32 || pe.sh[me], "dp_adamw: need32 (bf16-only broadcast) requires the shadow arenas");
  if (sms <= 0) sms = dgpt_sm_count();
  if (sms <= 0) sms = 148;
  cudaStream_t st = (cudaStream_t)stream;
  // every block must be resident at once (the last one waits for the peers): one 512-thread block per SM at most
  dp_adamw_kernel<<<sms, 512, 0, st>>>(pe, m, v, lo, hi, world, me, hyper, step, epoch, scratch, (int*)(scratch + 1));
  int rc = check_launch("dp_adamw");
  if (rc) return rc;
  dp_counters_kernel<<<1, 1, 0, st>>>(epoch, step);
  return check_launch("dp_adamw");
}

}  // extern "C"
