// Issue-rate probe for the softmax instruction mix on sm_100a: cycles per warp-instruction of FFMA, FFMA2, FADD2,
// MUFU.EX2, FMNMX, F2FP, IMAD with 1 / 2 / 4 warps per scheduler (8 independent chains per thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/pipe_rates.bin scripts/pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm volatile("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

template <int OP>
__global__ void probe(float* out, long long* cyc, int iters, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (OP == 0) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(v[i]) : "f"(seed)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(v[i + 1]) : "f"(seed)); }
      if (OP == 1) { ffma2(v[i], v[i + 1], v[i], v[i + 1], seed, seed, seed, seed); }
      if (OP == 2) { fadd2(v[i], v[i + 1], v[i], v[i + 1], seed, seed); }
      if (OP == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i + 1])); }
      if (OP == 4) { asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(seed)); asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i + 1]) : "f"(seed)); }
      if (OP == 5) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(v[i + 1])); v[i] = __uint_as_float(r); }
      if (OP == 6) { int a = __float_as_int(v[i]), b = __float_as_int(v[i + 1]); asm volatile("mad.lo.s32 %0, %1, 8388608, %0;" : "+r"(a) : "r"(b)); v[i] = __int_as_float(a); }
      if (OP == 7) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(v[i + 1]), "f"(seed)); }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1024 * 148 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps = 4; warps <= 16; warps *= 2) {
    probe<OP><<<148, warps * 32>>>(out, cyc, iters, 0.5f);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s warps/scheduler %d: %.2f cycles per warp-instruction per scheduler\n", name, warps / 4,
           double(h) / (double(iters) * per_iter * (warps / 4)));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA", 16); run<1>("FFMA2", 8); run<2>("FADD2", 8); run<3>("MUFU.EX2", 16); run<4>("FMNMX", 16);
  run<5>("F2FP", 8); run<6>("IMAD", 8); run<7>("FMNMX3", 8);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
