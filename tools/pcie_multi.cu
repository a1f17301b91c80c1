// developer tool: the host<->device copy ceiling of a multi-GPU box, the number the frame-parallel e2e arm of bench.py is
// bounded by.  For several subsets of the visible GPUs, every GPU of the subset copies concurrently from its own host
// thread and its own pinned buffers: the bench's per-step traffic (478 MB H2D per eye-batch, 131 MB D2H), H2D alone and
// H2D with a concurrent D2H stream.  Prints per-GPU and aggregate GB/s so that a shared limit (host memory, IOMMU, root
// complex, PCIe switch) shows up as per-GPU rates falling while the aggregate stays flat.
//   nvcc -O2 -std=c++17 -o tools/bin/pcie_multi tools/pcie_multi.cu -lpthread     (run on the GPU box)
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static std::atomic<int> g_ready{0};
static std::atomic<bool> g_go{false};

struct Result { double h2d = 0, d2h = 0, seconds = 0; };

static void worker(int dev, bool withD2H, bool wc, int iters, Result* out) {
  const size_t B = 478ull << 20, O = 131ull << 20;
  cudaSetDevice(dev);
  void *hi = nullptr, *ho = nullptr, *di = nullptr, *dout = nullptr;
  cudaHostAlloc(&hi, B, cudaHostAllocPortable | (wc ? cudaHostAllocWriteCombined : 0));
  cudaHostAlloc(&ho, O, cudaHostAllocPortable);
  cudaMalloc(&di, B); cudaMalloc(&dout, O);
  std::memset(ho, 0, O);
  if (!wc) std::memset(hi, 1, B);
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaMemcpyAsync(di, hi, B, cudaMemcpyHostToDevice, s1); cudaDeviceSynchronize();      // warm
  g_ready.fetch_add(1);
  while (!g_go.load()) std::this_thread::yield();
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < iters; ++k) {
    cudaMemcpyAsync(di, hi, B, cudaMemcpyHostToDevice, s1);
    if (withD2H) cudaMemcpyAsync(ho, dout, O, cudaMemcpyDeviceToHost, s2);
  }
  cudaDeviceSynchronize();
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  out->h2d = (double)iters * B / out->seconds / 1e9;
  out->d2h = withD2H ? (double)iters * O / out->seconds / 1e9 : 0;
  cudaFreeHost(hi); cudaFreeHost(ho); cudaFree(di); cudaFree(dout);
}

static void run(const std::vector<int>& devs, bool withD2H, bool wc) {
  g_ready = 0; g_go = false;
  std::vector<Result> res(devs.size());
  std::vector<std::thread> th;
  for (size_t i = 0; i < devs.size(); ++i) th.emplace_back(worker, devs[i], withD2H, wc, 8, &res[i]);
  while (g_ready.load() < (int)devs.size()) std::this_thread::yield();
  g_go = true;
  for (auto& t : th) t.join();
  double sum = 0, sumo = 0;
  std::string per;
  for (size_t i = 0; i < devs.size(); ++i) { sum += res[i].h2d; sumo += res[i].d2h; char b[64]; snprintf(b, sizeof b, " gpu%d %.1f", devs[i], res[i].h2d); per += b; }
  printf("%-8s %-10s %zu GPUs: H2D total %6.1f GB/s (D2H %5.1f) |%s\n", wc ? "wc" : "pinned", withD2H ? "H2D+D2H" : "H2D", devs.size(), sum, sumo, per.c_str());
  fflush(stdout);
}

int main() {
  int n = 0;
  cudaGetDeviceCount(&n);
  printf("%d visible GPUs\n", n);
  std::vector<std::vector<int>> sets = {{0}};
  if (n >= 2) { sets.push_back({0, 1}); }
  if (n >= 4) { sets.push_back({0, 2}); sets.push_back({0, 1, 2, 3}); }
  if (n >= 8) { sets.push_back({0, 4}); sets.push_back({0, 2, 4, 6}); sets.push_back({4, 5, 6, 7}); sets.push_back({0, 1, 2, 3, 4, 5, 6, 7}); }
  for (const auto& s : sets) { run(s, false, false); run(s, true, false); }
  if (n >= 2) { std::vector<int> all; for (int i = 0; i < n; ++i) all.push_back(i); run(all, true, true); }
  return 0;
}
