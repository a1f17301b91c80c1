// developer tool: the floor of a small host->device->host round trip on this box, the fixed cost under the one-frame
// latency path.  Wall clock (median of 500) of: an empty kernel + sync; a pinned H2D copy of one KITTI frame + sync; the
// same copy on a copy stream handed to a compute stream through an event, followed by a small kernel; a 120 KB D2H + sync;
// cudaStreamSynchronize vs cudaEventSynchronize vs spinning on cudaEventQuery.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/h2d_probe tools/h2d_probe.cu
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <vector>
#include <cuda_runtime.h>

__global__ void k_touch(const unsigned* in, unsigned* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] + 1;
}
__global__ void k_empty() {}

static double med(const std::function<void()>& f, int iters = 500) {
  for (int i = 0; i < 30; ++i) f();
  std::vector<double> t;
  for (int i = 0; i < iters; ++i) {
    const auto a = std::chrono::steady_clock::now();
    f();
    t.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count());
  }
  std::sort(t.begin(), t.end());
  return t[t.size() / 2];
}

int main() {
  const size_t B = 1241 * 376, O = 120 * 1024;
  void *hp, *ho, *pg; unsigned *d, *d2;
  cudaHostAlloc(&hp, B, cudaHostAllocPortable); cudaHostAlloc(&ho, O, cudaHostAllocPortable);
  pg = malloc(B); memset(pg, 1, B); memset(hp, 1, B);
  cudaMalloc(&d, B + 64); cudaMalloc(&d2, B + 64);
  cudaStream_t s, c; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking);
  cudaEvent_t e, e2, eb; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); cudaEventCreateWithFlags(&e2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&eb, cudaEventDisableTiming | cudaEventBlockingSync);
  const int n = (int)(B / 4);
  printf("empty kernel + stream sync             %.1f us\n", med([&] { k_empty<<<1, 32, 0, s>>>(); cudaStreamSynchronize(s); }));
  printf("empty kernel + event sync              %.1f us\n", med([&] { k_empty<<<1, 32, 0, s>>>(); cudaEventRecord(e, s); cudaEventSynchronize(e); }));
  printf("empty kernel + blocking-event sync     %.1f us\n", med([&] { k_empty<<<1, 32, 0, s>>>(); cudaEventRecord(eb, s); cudaEventSynchronize(eb); }));
  printf("empty kernel + event query spin        %.1f us\n", med([&] { k_empty<<<1, 32, 0, s>>>(); cudaEventRecord(e, s); while (cudaEventQuery(e) == cudaErrorNotReady) {} }));
  for (size_t bytes : {(size_t)4096, (size_t)65536, B}) {
    printf("H2D pinned %7zu B + sync             %.1f us\n", bytes, med([&] { cudaMemcpyAsync(d, hp, bytes, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s); }));
    printf("H2D pageable %7zu B + sync           %.1f us\n", bytes, med([&] { cudaMemcpyAsync(d, pg, bytes, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s); }));
  }
  printf("memcpy pageable->pinned %zu B         %.1f us\n", B, med([&] { memcpy(hp, pg, B); }));
  printf("H2D pinned frame, same stream, kernel, sync          %.1f us\n",
         med([&] { cudaMemcpyAsync(d, hp, B, cudaMemcpyHostToDevice, s); k_touch<<<(n + 255) / 256, 256, 0, s>>>(d, d2, n); cudaStreamSynchronize(s); }));
  printf("H2D pinned frame on copy stream, event, kernel, sync %.1f us\n",
         med([&] { cudaStreamWaitEvent(c, e2, 0); cudaMemcpyAsync(d, hp, B, cudaMemcpyHostToDevice, c); cudaEventRecord(e, c); cudaStreamWaitEvent(s, e, 0);
                   k_touch<<<(n + 255) / 256, 256, 0, s>>>(d, d2, n); cudaEventRecord(e2, s); cudaStreamSynchronize(s); }));
  printf("kernel reads the pinned frame in place (zero copy), sync %.1f us\n",
         med([&] { k_touch<<<(n + 255) / 256, 256, 0, s>>>((const unsigned*)hp, d2, n); cudaStreamSynchronize(s); }));
  printf("D2H %zu B pinned + sync                %.1f us\n", O, med([&] { cudaMemcpyAsync(ho, d, O, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }));
  printf("D2H 3 copies (56K + 64K + 4) + sync    %.1f us\n",
         med([&] { cudaMemcpyAsync(ho, d, 56000, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync((char*)ho + 56000, d + 20000, 64000, cudaMemcpyDeviceToHost, s);
                   cudaMemcpyAsync((char*)ho + 120000, d + 40000, 4, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }));
  printf("kernel writes 120 KB into pinned host memory, sync   %.1f us\n",
         med([&] { k_touch<<<(30720 + 255) / 256, 256, 0, s>>>(d, (unsigned*)ho, 30720); cudaStreamSynchronize(s); }));
  return 0;
}
