// Accuracy of cos.approx.ftz.f32(pi z) for z in [-2, 2] with and without an exact reduction of z to [-1, 1] half-turns
// (rollout_tc2.cu needs to know whether the FADD/FMUL/FRND/FFMA reduction before MUFU.COS can go).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(int n, float* e_plain, float* e_red) {
  float mp = 0.f, mr = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float z = -2.0f + 4.0f * (float)i / (float)n;
    const double ref = cos(3.14159265358979323846 * (double)z);
    float a, b;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(3.14159265358979f * z));
    float zr = fmaf(-2.0f, rintf(0.5f * z), z);
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(3.14159265358979f * zr));
    mp = fmaxf(mp, (float)fabs((double)a - ref));
    mr = fmaxf(mr, (float)fabs((double)b - ref));
  }
  atomicMax((int*)e_plain, __float_as_int(mp));
  atomicMax((int*)e_red, __float_as_int(mr));
}
int main() {
  float *d, h[2];
  cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
  k<<<148, 256>>>(1 << 24, d, d + 1);
  cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("max abs error of cos.approx(pi z), z in [-2,2]: plain %.3e   with exact reduction to [-1,1] %.3e\n", h[0], h[1]);
  return 0;
}
