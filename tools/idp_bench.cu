// developer tool: issue throughput of integer / DPX / half2 min-max / dot-product instructions on this GPU, alone and mixed
// (8 independent chains per thread): which ones share a pipe?
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b) {
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<unsigned*>(&r);
}
template <int OP> __global__ void k(unsigned* out, unsigned a, unsigned b, int iters) {
  unsigned x[8];
  for (int i = 0; i < 8; ++i) x[i] = 0x64006400u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = x[i] * a + b;
      else if (OP == 1) x[i] = __dp4a(x[i], a, b);
      else if (OP == 2) x[i] = __vmaxu2(x[i], a) + 1;            // VIMNMX.U16x2 (+ add to keep the chain alive)
      else if (OP == 3) x[i] = hmax2u(x[i], a) + 1;              // HMNMX2
      else if (OP == 4) x[i] = (i & 1) ? hmax2u(x[i], a) + 1 : __vmaxu2(x[i], a) + 1;      // half the chains each
      else if (OP == 5) x[i] = (i & 1) ? x[i] * a + b : __vmaxu2(x[i], a) + 1;              // IMAD + VIMNMX
      else x[i] = __vimax3_u16x2(x[i], a, b) + 1;
    }
  }
  unsigned s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> float run(unsigned* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<OP><<<148 * 8, 256>>>(d, 0x64056405u, 5, iters);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
  }
  return ms;
}
int main() {
  unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 20000;
  const char* names[] = {"IMAD", "IDP.4A", "VIMNMX.U16x2 + IADD", "HMNMX2 + IADD", "VIMNMX/HMNMX2 mix + IADD", "IMAD / VIMNMX+IADD mix", "VIMNMX3.U16x2 + IADD"};
  float ms[7] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters), run<5>(d, iters), run<6>(d, iters)};
  for (int op = 0; op < 7; ++op) {
    const double chains = 148.0 * 8 * 256 * 8.0 * iters;
    printf("%-28s %.2f ms  -> %.1f chain-steps/clk/SM\n", names[op], ms[op], chains / (ms[op] * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
