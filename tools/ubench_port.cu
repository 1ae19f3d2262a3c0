// Do an SM's TMA loads and its output stores share one memory port?  One CTA per SM: warp 0 streams bulk
// loads global -> shared (ring of 4 x 16 KB, mbarrier-tracked), warps 1..8 stream stores shared -> global,
// either as bulk (TMA-path) stores or as coalesced STG.128 (LSU path).  Prints cycles for each alone and combined.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_port tools/ubench_port.cu && tools/ubench_port
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../drakegpt_b200/csrc/ptx.cuh"

using namespace dgpt::ptx;

// mode bit 0: loads on; bit 1: bulk stores on; bit 2: STG stores on
__global__ void __launch_bounds__(288, 1) port(int mode, int load_iters, int store_iters, const uint8_t* src, uint8_t* dst,
                                                long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * 16384 + 8 * 8192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    if ((mode & 1) && lane == 0) {
      const uint8_t* s = src + (size_t)blockIdx.x * 16 * 16384;  // 256 KB per CTA, re-read from L2
      for (int i = 0; i < load_iters + 4; ++i) {
        if (i >= 4) mbar_wait(&bars[i & 3], ((i - 4) >> 2) & 1);  // the load that used this slot has landed
        if (i < load_iters) {
          mbar_expect_tx(&bars[i & 3], 16384);
          bulk_load_1d(smem + (i & 3) * 16384, s + (size_t)(i & 15) * 16384, 16384, &bars[i & 3]);
        }
      }
    }
  } else {
    uint8_t* stage = smem + 4 * 16384 + (warp - 1) * 8192;
    uint8_t* d = dst + ((size_t)blockIdx.x * 8 + (warp - 1)) * 8 * 4096;  // 32 KB per warp, rewritten in L2
    if (mode & 2) {
      for (int i = 0; i < store_iters; ++i) {
        uint8_t* tile = stage + (i & 1) * 4096;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(i, j, lane, warp);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + (size_t)(i & 7) * 4096),
                       "r"(smem_u32(tile)), "r"(4096)
                       : "memory");
          bulk_commit();
        }
      }
      if (lane == 0) bulk_wait<0>();
    } else if (mode & 4) {
      for (int i = 0; i < store_iters; ++i) {
        uint8_t* tile = stage + (i & 1) * 4096;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(i, j, lane, warp);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int row = j * 4 + (lane >> 3), ch = lane & 7;
          const uint4 v = *reinterpret_cast<const uint4*>(tile + row * 128 + ((ch ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(d + (size_t)(i & 7) * 4096 + row * 128 + ch * 16) = v;
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const int load_iters = 384, store_iters = 96;  // per CTA: 6 MB loaded, 8 warps x 384 KB = 3 MB stored
  uint8_t *src, *dst;
  long long* out;
  cudaMalloc(&src, (size_t)148 * 16 * 16384);
  cudaMalloc(&dst, (size_t)148 * 8 * 8 * 4096);
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaMemset(src, 1, (size_t)148 * 16 * 16384);
  const int smem = 4 * 16384 + 8 * 8192 + 64;
  cudaFuncSetAttribute(port, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"", "loads only", "bulk stores only", "loads + bulk stores", "STG stores only", "loads + STG stores"};
  for (int mode : {1, 2, 3, 4, 5}) {
    for (int rep = 0; rep < 3; ++rep) port<<<148, 288, smem>>>(mode, load_iters, store_iters, src, dst, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    double avg = 0;
    for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i] / 148.0; }
    printf("%-22s avg %8.0f max %8lld cycles   (loads %.1f B/cyc/SM, stores %.1f B/cyc/SM)\n", names[mode], avg, mx,
           (mode & 1) ? load_iters * 16384.0 / avg : 0.0, (mode & 6) ? 8.0 * store_iters * 4096 / avg : 0.0);
  }
  return 0;
}
