// Global-store pattern microbenchmark for the GEMM epilogue question (profiles/README.md): 148 CTAs x 256 threads write a
// [6512 x 2048] bf16 matrix tile by tile (128 x 256 tiles, the epilogue's warp -> (row quarter, column half) mapping).
//   mode 0: one row per thread, 32-byte stores (the GEMM epilogue since r02)      mode 2: same with 16-byte stores
//   mode 1: warp-coalesced (16 lanes x 16 B = one 256-byte row segment per half warp)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/store_pattern profiles/store_pattern.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void stg256(void* p, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256) k(uint16_t* C, int M, int N, int mode) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, half = warp >> 2;
  const int tiles_m = (M + 127) / 128, tiles_n = N / 256;
  for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x) {
    const int mt = t % tiles_m, nt = t / tiles_m;
    if (mode == 1) {
      // warp = 32 rows x 128 columns (256 B per row): 2 rows per instruction
      for (int rr = 0; rr < 32; rr += 2) {
        const int m = mt * 128 + q * 32 + rr + (lane >> 4);
        if (m < M) *reinterpret_cast<uint4*>(C + (size_t)m * N + nt * 256 + half * 128 + (lane & 15) * 8) = make_uint4(t, t, t, t);
      }
    } else {
      const int m = mt * 128 + q * 32 + lane;
      if (m >= M) continue;
      for (int c = 0; c < 4; ++c) {
        uint16_t* p = C + (size_t)m * N + nt * 256 + half * 128 + c * 32;
        if (mode == 0) {
          stg256(p, t);
          stg256(p + 16, t);
        } else {
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(p + 8 * j) = make_uint4(t, t, t, t);
        }
      }
    }
  }
}

int main() {
  const int M = 6512, N = 2048;
  uint16_t* C;
  cudaMalloc(&C, (size_t)M * N * 2 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    for (int i = 0; i < 3; ++i) k<<<148, 256>>>(C, M, N, mode);
    cudaEventRecord(e0);
    for (int i = 0; i < 40; ++i) k<<<148, 256>>>(C + (size_t)(i % 8) * M * N, M, N, mode);  // 8 x 26.7 MB > L2
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d: %.2f us per 26.7 MB matrix = %.0f GB/s   (%s)\n", mode, ms * 1e3 / 40, (double)M * N * 2 / (ms * 1e-3 / 40) / 1e9,
           mode == 0 ? "row per thread, 32-byte stores" : mode == 1 ? "warp-coalesced 16-byte stores" : "row per thread, 16-byte stores");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
