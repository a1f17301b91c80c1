// developer tool: where one eye's one-frame latency goes, in C++ against the C ABI (no interpreter in the way).
// Wall clock, median of 500: each phase followed by ivg_sync, then the whole ivg_extract.
//   g++ -O2 -std=c++17 -Iinclude -o tools/bin/latency_steps tools/latency_steps.cpp -Liv_slam_b200/lib -livslam_gpu -Wl,-rpath,$PWD/iv_slam_b200/lib
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>
#include "ivslam_gpu.h"

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

int main(int argc, char** argv) {
  const int W = 1241, H = 376;
  std::vector<uint8_t> img((size_t)W * H);
  // argv[1]: a raw 1241x376 8-bit frame (python -c "from iv_slam_b200 import synthetic as S; S.make_stereo_pair(1241,376,0)[0].tofile('/tmp/left.raw')")
  FILE* f = argc > 1 ? std::fopen(argv[1], "rb") : nullptr;
  if (!f || std::fread(img.data(), 1, img.size(), f) != img.size()) { std::fprintf(stderr, "usage: latency_steps frame.raw\n"); return 2; }
  std::fclose(f);
  ivg_extractor* h = nullptr;
  if (ivg_extractor_create(&h, 0, 2000, 1.2f, 8, 20, 7, 0)) return 1;
  ivg_set_graph_mode(h, 1);
  const int cap = ivg_max_keypoints(h);
  void *pimg, *pk, *pd, *pn;
  ivg_host_alloc(&pimg, img.size()); ivg_host_alloc(&pk, (size_t)cap * 28); ivg_host_alloc(&pd, (size_t)cap * 32); ivg_host_alloc(&pn, 64);
  std::memcpy(pimg, img.data(), img.size());
  std::vector<uint8_t> kq((size_t)cap * 28), dq((size_t)cap * 32);
  int n = 0;
  for (int pinned = 0; pinned < 2 && !(argc > 2); ++pinned) {
    const uint8_t* src = pinned ? (const uint8_t*)pimg : img.data();
    const double up = med([&] { ivg_upload_batch(h, 1, src, W, H, W, (size_t)W * H, nullptr, 0, 0); ivg_sync(h); });
    const double run = med([&] { ivg_run_batch(h); ivg_sync(h); });
    const double dl = med([&] { ivg_download_batch(h, (ivg_keypoint*)pk, (uint8_t*)pd, cap, (int*)pn); ivg_sync(h); });
    const double sy = med([&] { ivg_sync(h); });
    const double all_pin = med([&] { ivg_extract(h, src, W, H, W, nullptr, 0, (ivg_keypoint*)pk, (uint8_t*)pd, cap, &n); });
    const double all_pg = med([&] { ivg_extract(h, src, W, H, W, nullptr, 0, (ivg_keypoint*)kq.data(), dq.data(), cap, &n); });
    std::printf("%s image: upload+sync %.1f  run+sync %.1f  download+sync %.1f  bare sync %.1f | ivg_extract into pinned results %.1f, into pageable results %.1f us (%d keypoints)\n",
                pinned ? "pinned  " : "pageable", up, run, dl, sy, all_pin, all_pg, n);
  }
  // threading patterns around the same call (pageable image, pageable results): the reference spawns two std::threads per
  // frame (src/Frame.cc:115-125)
  ivg_extractor* h2 = nullptr;
  if (ivg_extractor_create(&h2, 0, 2000, 1.2f, 8, 20, 7, 0)) return 1;
  ivg_set_graph_mode(h2, 1);
  std::vector<uint8_t> kq2((size_t)cap * 28), dq2((size_t)cap * 32), img2(img);
  int n2 = 0;
  auto exL = [&] { ivg_extract(h, img.data(), W, H, W, nullptr, 0, (ivg_keypoint*)kq.data(), dq.data(), cap, &n); };
  auto exR = [&] { ivg_extract(h2, img2.data(), W, H, W, nullptr, 0, (ivg_keypoint*)kq2.data(), dq2.data(), cap, &n2); };
  std::printf("one eye, calling thread %.1f | one eye, fresh std::thread %.1f | two eyes, calling thread, one after the other %.1f | "
              "two eyes, two fresh std::threads %.1f us\n",
              med(exL), med([&] { std::thread t(exL); t.join(); }), med([&] { exL(); exR(); }),
              med([&] { std::thread a(exL), b(exR); a.join(); b.join(); }));
  std::printf("empty fresh std::thread spawn+join %.1f | two %.1f us\n", med([] { std::thread t([] {}); t.join(); }),
              med([] { std::thread a([] {}), b([] {}); a.join(); b.join(); }));
  {
    // two persistent threads, woken by a spin flag: what the two-eye frame costs without thread creation
    std::atomic<int> go{0}, done{0}; std::atomic<bool> quit{false};
    auto worker = [&](int bit, const std::function<void()>& f) {
      int seen = 0;
      while (!quit.load(std::memory_order_acquire)) {
        if (go.load(std::memory_order_acquire) != seen) { ++seen; f(); done.fetch_add(1, std::memory_order_acq_rel); }
      }
      (void)bit;
    };
    std::function<void()> fl = exL, fr = exR;
    std::thread a(worker, 0, std::cref(fl)), b(worker, 1, std::cref(fr));
    const double t = med([&] { done.store(0); go.fetch_add(1); while (done.load(std::memory_order_acquire) < 2) {} });
    quit = true; a.join(); b.join();
    std::printf("two eyes, two persistent spinning threads %.1f us\n", t);
  }
  ivg_extractor_destroy(h2);
  ivg_extractor_destroy(h);
  return 0;
}
