// developer tool: H2D bandwidth from default-pinned vs write-combined host memory, alone and with a concurrent D2H stream
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const size_t B = 478ull << 20, O = 131ull << 20;
  void *hd, *hw, *ho, *d, *d2;
  CK(cudaHostAlloc(&hd, B, cudaHostAllocPortable));
  CK(cudaHostAlloc(&hw, B, cudaHostAllocPortable | cudaHostAllocWriteCombined));
  CK(cudaHostAlloc(&ho, O, cudaHostAllocPortable));
  CK(cudaMalloc(&d, B)); CK(cudaMalloc(&d2, O));
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int mode = 0; mode < 4; ++mode) {
    void* src = (mode & 1) ? hw : hd;
    const bool dup = mode & 2;
    for (int it = 0; it < 2; ++it) {
      CK(cudaEventRecord(a, s1));
      for (int k = 0; k < 4; ++k) { CK(cudaMemcpyAsync(d, src, B, cudaMemcpyHostToDevice, s1)); if (dup) CK(cudaMemcpyAsync(ho, d2, O, cudaMemcpyDeviceToHost, s2)); }
      CK(cudaEventRecord(b, s1)); CK(cudaDeviceSynchronize());
    }
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    printf("%s%s: H2D %.1f GB/s\n", (mode & 1) ? "write-combined" : "default pinned", dup ? " + concurrent D2H" : "", 4.0 * B / ms / 1e6);
  }
  return 0;
}
