// Micro-benchmarks behind the GEMM epilogue design: TMEM read throughput per warp /
// per SM, the cost of fence.proxy.async, and of small shared->global bulk stores.  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_epi tools/ubench_epi.cu && tools/ubench_epi
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../drakegpt_b200/csrc/ptx.cuh"

using namespace dgpt::ptx;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// mode 0: per iteration 2 x LDTM.x32 (different register sets) + wait      (latency-ish, 8 KB / iter / warp)
// mode 1: per iteration 1 x LDTM.x32 + wait                                 (pure round trip)
// mode 2: STS 4 KB + fence.proxy.async + syncwarp                            (fence cost)
// mode 3: mode 2 + bulk store 4 KB + commit + wait_group.read 1              (staging pipeline, 2 buffers)
// mode 4: like 3 without the fence (invalid for real use; isolates the fence)
// mode 5: LDTM.x16 x 2 + wait
// mode 6: like 3 but one 16 KB store per four iterations (larger stores)
__global__ void __launch_bounds__(512, 1) ubench(int mode, int nwarps, int iters, long long* out, uint8_t* gbuf) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const int quad = warp & 3;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    uint8_t* stage = smem + (warp & 7) * 16384;
    uint8_t* gdst = gbuf + ((size_t)blockIdx.x * 16 + warp) * 16384;
    asm volatile("bar.sync 1, %0;" ::"r"(nwarps * 32) : "memory");
    t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < iters; ++i) {
        uint32_t a[32], b[32];
        tmem_ld32(taddr + ((i * 64) & 127), a);
        tmem_ld32(taddr + ((i * 64 + 32) & 127), b);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= a[j] ^ b[j];
      }
    } else if (mode == 1) {
      for (int i = 0; i < iters; ++i) {
        uint32_t a[32];
        tmem_ld32(taddr + ((i * 32) & 127), a);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= a[j];
      }
    } else if (mode == 5) {
      for (int i = 0; i < iters; ++i) {
        uint32_t a[16], b[16];
        tmem_ld16(taddr + ((i * 32) & 127), a);
        tmem_ld16(taddr + ((i * 32 + 16) & 127), b);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= a[j] ^ b[j];
      }
    } else if (mode == 7) {
      // transpose through shared memory with the generic proxy only: STS (swizzled rows) -> LDS (row-contiguous)
      // -> coalesced STG.128 (each instruction writes 4 full 128-byte lines)
      for (int i = 0; i < iters; ++i) {
        uint8_t* tile = stage + (i & 1) * 4096;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(i, j, lane, acc);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int row = j * 4 + (lane >> 3), ch = lane & 7;
          const uint4 v = *reinterpret_cast<const uint4*>(tile + row * 128 + ((ch ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(gdst + (i & 3) * 4096 + row * 128 + ch * 16) = v;
        }
        __syncwarp();
      }
    } else if (mode == 8) {
      // direct row-strided STG.128 from registers (each lane owns a row: 32 different lines per instruction)
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(gdst + (i & 3) * 128 + lane * 512 + j * 16) = make_uint4(i, j, lane, acc);
      }
    } else {
      int sbuf = 0;
      for (int i = 0; i < iters; ++i) {
        uint8_t* tile = stage + sbuf * 4096;
        if (mode == 3 || mode == 4) {
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
        }
        if (mode == 6 && (i & 3) == 0) {
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
        }
        if (mode == 6) tile = stage + (i & 3) * 4096;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(i, j, lane, acc);
        if (mode != 4) fence_proxy_async();
        __syncwarp();
        if (mode == 3 || mode == 4) {
          if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst + sbuf * 4096),
                         "r"(smem_u32(tile)), "r"(4096)
                         : "memory");
            bulk_commit();
          }
          sbuf ^= 1;
        }
        if (mode == 6 && (i & 3) == 3) {
          if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(stage)),
                         "r"(16384)
                         : "memory");
            bulk_commit();
          }
        }
      }
      if (lane == 0) bulk_wait<0>();
    }
    t1 = clock64();
  }
  if (acc == 0x12345678u) out[1000] = acc;
  if (blockIdx.x == 0 && lane == 0 && warp < nwarps) out[warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  long long* out;
  uint8_t* gbuf;
  cudaMalloc(&out, 2048 * sizeof(long long));
  cudaMalloc(&gbuf, (size_t)148 * 16 * 16384);
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384);
  const char* names[] = {"2xLDTM.x32+wait", "1xLDTM.x32+wait", "STS4K+fence", "STS4K+fence+bulk4K", "STS4K+bulk4K(nofence)",
                         "2xLDTM.x16+wait", "STS4K+fence, 16K store/4 it", "STS4K+LDS+STG coalesced", "STG.128 row-strided 4K"};
  const int iters = 256;
  printf("grid = %d CTAs\n", grid);
  for (int mode = 2; mode < 9; ++mode) {
    if (mode == 5) continue;
    for (int nw : {1, 4, 8, 16}) {
      if (nw == 16 && mode >= 2 && mode != 5) continue;
      cudaMemset(out, 0, 2048 * sizeof(long long));
      for (int rep = 0; rep < 2; ++rep) ubench<<<grid, 512, 8 * 16384>>>(mode, nw, iters, out, gbuf);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d nw %d: %s\n", mode, nw, cudaGetErrorString(e)); return 1; }
      long long h[16];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("%-28s warps=%2d  %8.1f cycles/iter/warp  (SM-wide: %.1f cycles per warp-iter)\n", names[mode], nw,
             (double)mx / iters, (double)mx / iters / nw);
    }
  }
  return 0;
}
