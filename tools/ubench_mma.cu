// Raw tcgen05.mma issue / execution rate: one CTA per SM, operands resident in shared memory (no TMA, no
// pipeline), M = 128, K = 16 per instruction, accumulating into one TMEM tile.  Modes:
//   0: NMMA back-to-back MMAs, one commit at the end
//   1: a commit (to a barrier nobody waits on) after every 4 MMAs
//   2: a commit after every 4 MMAs AND the issuing warp waits for it before the next 4 (fully serialised)
//   3: like 1, plus an mbarrier try_wait on an already-completed barrier + tcgen05.fence between groups
//      (the per-k-block bookkeeping of the GEMM mainloop without any data dependence)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma tools/ubench_mma.cu && tools/ubench_mma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../drakegpt_b200/csrc/ptx.cuh"

using namespace dgpt::ptx;

// MODE is a template parameter on purpose: with run-time mode tests the compiler emits several predicated UTCHMMA
// variants per MMA, and a predicated-off UTCHMMA still costs issue time (the numbers of ALL modes then change).
template <int BN, int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate(int groups, long long* out) {
  constexpr int mode = MODE;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (4 * 49152) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
  if (warp == 0) {
    // a completed barrier for mode 3
    if (lane == 0) mbar_arrive(&bars[2]);
    __syncwarp();
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int g = 0; g < groups; ++g) {
      const uint32_t sa = smem_u32(smem + (g & 3) * 49152), sb = sa + 16384;
      if (mode == 3) {
        mbar_wait(&bars[2], 0);
        tc_fence_after();
      }
      const uint64_t da0 = make_smem_desc_sw128(sa, 16, 1024), db0 = make_smem_desc_sw128(sb, 16, 1024);
      if (mode == 5 || mode == 6) {
        // bookkeeping sliced between the MMA issues of the group (the tensor pipe buffers about one MMA):
        //   5: MMA0 | ready-barrier wait + fence | MMA1 MMA2 MMA3 | commit
        //   6: MMA0 | commit (of the previous group) | MMA1 | ready-barrier wait + fence | MMA2 MMA3
        if (elect_one()) tc_mma_bf16(tmem, da0, db0, idesc, g ? 1u : 0u);
        __syncwarp();
        if (mode == 5) {
          mbar_wait(&bars[2], 0);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 1; k < 4; ++k) tc_mma_bf16(tmem, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, 1u);
            tc_commit(&bars[1]);
          }
          __syncwarp();
        } else {
          if (elect_one()) {
            if (g) tc_commit(&bars[1]);
            tc_mma_bf16(tmem, da0 + 2, db0 + 2, idesc, 1u);
          }
          __syncwarp();
          mbar_wait(&bars[2], 0);
          tc_fence_after();
          if (elect_one()) {
            tc_mma_bf16(tmem, da0 + 4, db0 + 4, idesc, 1u);
            tc_mma_bf16(tmem, da0 + 6, db0 + 6, idesc, 1u);
          }
          __syncwarp();
        }
        continue;
      }
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode == 4) tc_mma_bf16_ts(tmem, tmem + 256 + (g & 3) * 32 + k * 8, db0 + (uint64_t)(2 * k), idesc, (g | k) ? 1u : 0u);
          else tc_mma_bf16(tmem, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, (g | k) ? 1u : 0u);
        }
        if (mode == 1 || mode == 3) tc_commit(&bars[1]);
        if (mode == 2) tc_commit(&bars[0]);
      }
      __syncwarp();
      if (mode == 2) {
        mbar_wait(&bars[0], ph);
        ph ^= 1;
        tc_fence_after();
      }
    }
    if (elect_one()) tc_commit(&bars[3]);
    __syncwarp();
    mbar_wait(&bars[3], 0);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  const int groups = 512;
  const int smem = 4 * 49152;

  const char* names[] = {"back-to-back", "commit every 4", "commit+wait every 4", "commit + ready-barrier wait + fence every 4",
                         "back-to-back, A from TMEM (TS mode)", "MMA0 | wait+fence | MMA1-3 | commit",
                         "MMA0 | commit(prev) | MMA1 | wait+fence | MMA2-3"};
  auto run = [&](auto kern, int bn, int mode) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) kern<<<148, 128, smem>>>(groups, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    long long h;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("N=%d %-52s %7.1f cycles per 4-MMA k-block (floor %d)\n", bn, names[mode], (double)h / groups, 2 * bn);
  };
  run(mma_rate<128, 0>, 128, 0); run(mma_rate<128, 1>, 128, 1); run(mma_rate<128, 2>, 128, 2); run(mma_rate<128, 3>, 128, 3);
  run(mma_rate<128, 4>, 128, 4); run(mma_rate<128, 5>, 128, 5); run(mma_rate<128, 6>, 128, 6);
  run(mma_rate<256, 0>, 256, 0); run(mma_rate<256, 1>, 256, 1); run(mma_rate<256, 2>, 256, 2); run(mma_rate<256, 3>, 256, 3);
  run(mma_rate<256, 5>, 256, 5); run(mma_rate<256, 6>, 256, 6);
  return 0;
}
