// Microbenchmark: do the divergent halves of one warp hide each other's latency on sm_100a?
// One warp per block.  Variant 0: all 32 lanes run path A.  Variant 1: lanes 0-15 run path A, lanes 16-31 run path B (distinct code).
// Variant 2: four groups of 8 lanes on four distinct paths.  Each path is a dependent chain (ALU / shared-memory / global pointer chase).
// If the hardware interleaves divergent paths while one waits, the diverged variants take about as long as variant 0; if it
// serialises them, 2x / 4x.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 4096

template <int SALT>
__device__ __noinline__ uint32_t alu_chain(uint32_t x) {
#pragma unroll 1
  for (int i = 0; i < N; ++i) x = x * (2654435761u + SALT) + (uint32_t)i + SALT;
  return x;
}
template <int SALT>
__device__ __noinline__ uint32_t smem_chain(const uint32_t* s, uint32_t x) {
#pragma unroll 1
  for (int i = 0; i < N; ++i) x = s[(x + SALT) & 1023];
  return x;
}
template <int SALT>
__device__ __noinline__ uint32_t gmem_chain(const uint32_t* g, uint32_t x) {
#pragma unroll 1
  for (int i = 0; i < N / 8; ++i) x = g[(x + SALT) & ((1u << 20) - 1)];
  return x;
}

__global__ void k(int variant, int kind, const uint32_t* g, uint32_t* out, long long* cycles) {
  __shared__ uint32_t s[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (i * 7919u + 13u) & 1023u;
  __syncwarp();
  const int lane = threadIdx.x;
  int path = 0;
  if (variant == 1) path = lane >> 4;
  if (variant == 2) path = lane >> 3;
  uint32_t x = lane >> 3;   // same start inside a group
  long long t0 = clock64();
  if (kind == 0) {
    if (path == 0) x = alu_chain<0>(x); else if (path == 1) x = alu_chain<1>(x); else if (path == 2) x = alu_chain<2>(x); else x = alu_chain<3>(x);
  } else if (kind == 1) {
    if (path == 0) x = smem_chain<0>(s, x); else if (path == 1) x = smem_chain<1>(s, x); else if (path == 2) x = smem_chain<2>(s, x); else x = smem_chain<3>(s, x);
  } else {
    if (path == 0) x = gmem_chain<0>(g, x); else if (path == 1) x = gmem_chain<1>(g, x); else if (path == 2) x = gmem_chain<2>(g, x); else x = gmem_chain<3>(g, x);
  }
  __syncwarp();
  long long t1 = clock64();
  out[blockIdx.x * 32 + lane] = x;
  if (lane == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  uint32_t* g; uint32_t* out; long long* cyc;
  cudaMalloc(&g, 4u << 20); cudaMalloc(&out, 4096 * 32 * 4); cudaMalloc(&cyc, 4096 * 8);
  uint32_t* h = (uint32_t*)malloc(4u << 20);
  for (uint32_t i = 0; i < (1u << 20); ++i) h[i] = (i * 2654435761u + 12345u) & ((1u << 20) - 1);
  cudaMemcpy(g, h, 4u << 20, cudaMemcpyHostToDevice);
  const char* kinds[3] = {"alu", "smem", "gmem"};
  for (int blocks : {1, 148 * 6}) {
    for (int kind = 0; kind < 3; ++kind) {
      for (int variant = 0; variant < 3; ++variant) {
        k<<<blocks, 32>>>(variant, kind, g, out, cyc);   // warm
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<blocks, 32>>>(variant, kind, g, out, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("blocks %4d  %-4s  variant %d (%d paths): %lld cycles in block 0, %.3f ms\n", blocks, kinds[kind], variant, variant == 0 ? 1 : variant * 2, c, ms);
      }
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
