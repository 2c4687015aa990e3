// Throughput probe: scalar FFMA vs packed fma.rn.f32x2 (FFMA2) on sm_100a, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%0,%1};\n"
               "fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0,%1}, rc;}" : "+f"(dx), "+f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
template <int PACKED>
__global__ void k(float* out, float a, float b, int iters) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) ffma2(acc[i], acc[i + 1], a, a, b, b);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int packed = 0; packed < 2; ++packed) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (packed) k<1><<<148 * 8, 256>>>(out, 0.999f, 1e-3f, iters); else k<0><<<148 * 8, 256>>>(out, 0.999f, 1e-3f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 8 * 256 * 16 * iters;
      if (rep == 2) printf("%s: %.3f ms, %.2f T lane-FMA/s (%.1f TFLOP/s)\n", packed ? "FFMA2 " : "FFMA  ", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
