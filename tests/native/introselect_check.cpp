// Host-side check of iv_slam_b200/csrc/introselect.h against the real libstdc++ std::nth_element
// (same comparator as cv::KeyPointsFilter::retainBest). Built and driven by tests/test_introselect.py.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include "../../iv_slam_b200/csrc/introselect.h"

struct KP { float x, y, size, angle, response; int octave, class_id; };

extern "C" long introselect_mismatches(int trials, int max_n, int max_val, unsigned seed, int adversarial) {
  std::mt19937 rng(seed);
  long bad = 0;
  for (int t = 0; t < trials; ++t) {
    int n = 1 + rng() % max_n;
    int nth = rng() % n;
    std::vector<KP> ref(n);
    std::vector<ivg::SelItem> mine(n);
    int vals = 1 + rng() % max_val;
    for (int i = 0; i < n; ++i) {
      float r;
      if (adversarial == 1) r = (float)(i % vals);                    // organ-pipe / sawtooth
      else if (adversarial == 2) r = (float)((i * 7919) % vals) * 0.37f;
      else r = (float)(rng() % vals);
      ref[i] = KP{0, 0, 0, 0, r, 0, i};
      uint32_t bits; std::memcpy(&bits, &r, 4);
      mine[i] = ivg::SelItem{bits, (uint32_t)i};
    }
    std::nth_element(ref.begin(), ref.begin() + nth, ref.end(), [](const KP& a, const KP& b) { return a.response > b.response; });
    ivg::sel_nth_element(mine.data(), nth, n);
    for (int i = 0; i < n; ++i)
      if ((uint32_t)ref[i].class_id != mine[i].val) { ++bad; break; }
  }
  return bad;
}

// Forces the heap_select branch by calling it the way __introselect does at depth 0.
extern "C" long heapselect_mismatches(int trials, int max_n, int max_val, unsigned seed) {
  std::mt19937 rng(seed);
  long bad = 0;
  for (int t = 0; t < trials; ++t) {
    int n = 4 + rng() % max_n;
    int nth = rng() % n;
    std::vector<KP> ref(n);
    std::vector<ivg::SelItem> mine(n);
    int vals = 1 + rng() % max_val;
    for (int i = 0; i < n; ++i) {
      float r = (float)(rng() % vals);
      ref[i] = KP{0, 0, 0, 0, r, 0, i};
      uint32_t bits; std::memcpy(&bits, &r, 4);
      mine[i] = ivg::SelItem{bits, (uint32_t)i};
    }
    auto cmp = [](const KP& a, const KP& b) { return a.response > b.response; };
    std::__heap_select(ref.begin(), ref.begin() + nth + 1, ref.end(), __gnu_cxx::__ops::__iter_comp_iter(cmp));
    std::iter_swap(ref.begin(), ref.begin() + nth);
    ivg::sel_heap_select(mine.data(), nth + 1, n);
    ivg::sel_swap(mine.data(), 0, nth);
    for (int i = 0; i < n; ++i)
      if ((uint32_t)ref[i].class_id != mine[i].val) { ++bad; break; }
  }
  return bad;
}
