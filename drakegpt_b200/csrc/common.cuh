// Shared device/host helpers for the drakegpt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/drakegpt_b200.h"

namespace dgpt {

// ---------------------------------------------------------------------------
// error plumbing (thread-local message, negative return codes)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);
int require_device();

#define DGPT_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      dgpt::set_error(__VA_ARGS__);    \
      return DGPT_E_ARG;               \
    }                                  \
  } while (0)

#define DGPT_DEVICE_OR_RETURN()            \
  do {                                     \
    int _rc = dgpt::require_device();      \
    if (_rc != 0) return _rc;              \
  } while (0)

// ---------------------------------------------------------------------------
// Philox4x32-10, the one dropout / sampling RNG of the library.
// ---------------------------------------------------------------------------
struct u32x4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                        uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return u32x4{c0, c1, c2, c3};
}

// ---------------------------------------------------------------------------
// Dropout mask: counter-based, two levels.  The mask is regenerated inside the attention and
// GEMM-epilogue inner loops (forward AND backward), so its cost is on the critical path:
//   * one SplitMix64 hash per GROUP of 32 consecutive elements:
//       (s0, s1) = mix64(seed + group * golden + (site + 1) * K)
//   * one 32-bit multiply-add per element e of the group:
//       word(e) = (e odd ? s1 : s0) * MUL[e] + ADD[e]      (MUL odd: a bijection of the group hash)
//     keep iff word(e) >= threshold, P(drop) = threshold / 2^32.
// ~3 integer instructions per element when a thread walks whole groups (MUL / ADD fold into
// immediates), against ~8 with one 64-bit hash per four elements and ~25 with Philox4x32-10.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct DropGroup {
  uint32_t s0, s1;
};
__host__ __device__ __forceinline__ DropGroup dropout_group(uint64_t seed, uint32_t site, uint64_t group) {
  const uint64_t h = mix64(seed + group * 0x9E3779B97F4A7C15ull + (uint64_t)(site + 1u) * 0xD1B54A32D192ED03ull);
  return DropGroup{(uint32_t)h, (uint32_t)(h >> 32)};
}
__host__ __device__ constexpr uint32_t drop_mul(uint32_t e) {
  uint32_t x = (e + 1u) * 0x9E3779B1u;
  x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13;
  return x | 1u;
}
__host__ __device__ constexpr uint32_t drop_add(uint32_t e) {
  uint32_t x = (e + 33u) * 0xC2B2AE3Du;
  x ^= x >> 16; x *= 0x27D4EB2Fu; x ^= x >> 15;
  return x;
}
// uniform 32-bit word of element e (0..31) of the group
__host__ __device__ __forceinline__ uint32_t dropout_word(const DropGroup& g, uint32_t e) {
  return ((e & 1u) ? g.s1 : g.s0) * drop_mul(e) + drop_add(e);
}

// the four words of elements [4 quad, 4 quad + 3] (any quad; one group hash per call)
__host__ __device__ __forceinline__ u32x4 dropout_bits4(uint64_t seed, uint32_t site, uint64_t quad) {
  const DropGroup g = dropout_group(seed, site, quad >> 3);
  const uint32_t e0 = ((uint32_t)quad & 7u) * 4u;
  return u32x4{dropout_word(g, e0), dropout_word(g, e0 + 1u), dropout_word(g, e0 + 2u), dropout_word(g, e0 + 3u)};
}

__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  // keep iff word >= threshold; P(drop) = threshold / 2^32 (0 = dropout off)
  double t = (double)p * 4294967296.0 + 0.5;
  if (t < 0.0) t = 0.0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

__host__ __device__ __forceinline__ uint32_t pick4(const u32x4& r, int lane) {
  return lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
}

__host__ __device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t site, uint64_t index,
                                                      uint32_t threshold) {
  return dropout_word(dropout_group(seed, site, index >> 5), (uint32_t)index & 31u) >= threshold;
}

// ---------------------------------------------------------------------------
// typed loads / stores for the two activation types
// ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Effective SM clock of a kernel (DGPT_CLOCK_PROBE=1): one thread stamps clock64 and %globaltimer at kernel entry and
// exit into a 4-word device buffer (clock_probe_buffer(), NULL when the probe is off) that the kernel receives as an
// argument; dgpt_debug_clock_probe() returns (cycles, nanoseconds) of the last stamped kernel.
unsigned long long* clock_probe_buffer();
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void clock_probe_begin(unsigned long long* buf) {
  if (buf && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
    buf[0] = (unsigned long long)clock64();
    buf[1] = global_timer_ns();
  }
}
__device__ __forceinline__ void clock_probe_end(unsigned long long* buf) {
  if (buf && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
    buf[2] = (unsigned long long)clock64();
    buf[3] = global_timer_ns();
  }
}

// ---------------------------------------------------------------------------
// Programmatic dependent launch.  Kernels that begin with pdl_grid_sync() are launched through launch_pdl():
// their CTAs may be scheduled (and run their on-chip prologue) while the previous kernel of the stream is still
// draining; pdl_grid_sync() then waits until that kernel has completed and its writes are visible.
// DGPT_PDL=0 turns the launch attribute off (the device-side calls are then no-ops).
// ---------------------------------------------------------------------------
bool pdl_enabled();
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface in check_launch()
}

}  // namespace dgpt
