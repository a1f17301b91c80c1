// introselect.h — std::nth_element as libstdc++ 13 executes it, restated on a plain array.
//
// Why: the reference trims each FAST cell and each pyramid level with cv::KeyPointsFilter::retainBest followed by
// vector::resize(n) (introspective_ORB_SLAM/src/ORBextractor.cc:1146-1148, :1162-1166).  retainBest is
// std::nth_element(begin, begin+n-1, end, response-greater); FAST responses are small integers, so ties at the cut
// are common and WHICH of the tied keypoints survive — and the order of all survivors — is whatever permutation
// libstdc++'s introselect leaves behind.  That permutation depends only on comparator
// outcomes, so one GPU thread can replay it verbatim.  Algorithm structure follows the published libstdc++
// (bits/stl_algo.h __introselect / __unguarded_partition_pivot / __move_median_to_first / __insertion_sort /
// __heap_select; bits/stl_heap.h __make_heap / __adjust_heap / __push_heap / __pop_heap), GCC 13.
//
// Compiles as plain C++ too, so tests/ can check it against the real std::nth_element on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define IVG_HD __host__ __device__ __forceinline__
#else
#define IVG_HD inline
#endif

namespace ivg {

struct SelItem { uint32_t key; uint32_t val; };   // key: response (non-negative float bits compare like unsigned), val: payload

IVG_HD bool sel_before(const SelItem& a, const SelItem& b) { return a.key > b.key; }   // KeypointResponseGreater
IVG_HD void sel_swap(SelItem* a, int i, int j) { SelItem t = a[i]; a[i] = a[j]; a[j] = t; }

IVG_HD void sel_adjust_heap(SelItem* a, int hole, int len, SelItem value) {
  const int top = hole;
  int second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (sel_before(a[second], a[second - 1])) second--;
    a[hole] = a[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    a[hole] = a[second - 1];
    hole = second - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && sel_before(a[parent], value)) {
    a[hole] = a[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a[hole] = value;
}

IVG_HD void sel_heap_select(SelItem* a, int middle, int last) {   // range [0,last), heap on [0,middle)
  const int len = middle;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    while (true) {
      SelItem v = a[parent];
      sel_adjust_heap(a, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  for (int i = middle; i < last; ++i)
    if (sel_before(a[i], a[0])) {
      SelItem v = a[i];
      a[i] = a[0];
      sel_adjust_heap(a, 0, len, v);
    }
}

// std::nth_element(a, a+nth, a+n, greater-by-key)
IVG_HD void sel_nth_element(SelItem* a, int nth, int n) {
  if (n == 0 || nth == n) return;
  int first = 0, last = n;
  int lg = 0;
  for (int t = n; t > 1; t >>= 1) ++lg;
  int depth = 2 * lg;
  while (last - first > 3) {
    if (depth == 0) {
      sel_heap_select(a + first, nth + 1 - first, last - first);
      sel_swap(a, first, nth);
      return;
    }
    --depth;
    const int mid = first + (last - first) / 2;
    const int A = first + 1, B = mid, C = last - 1;
    if (sel_before(a[A], a[B])) {
      if (sel_before(a[B], a[C])) sel_swap(a, first, B);
      else if (sel_before(a[A], a[C])) sel_swap(a, first, C);
      else sel_swap(a, first, A);
    } else if (sel_before(a[A], a[C])) sel_swap(a, first, A);
    else if (sel_before(a[B], a[C])) sel_swap(a, first, C);
    else sel_swap(a, first, B);
    const SelItem pivot = a[first];
    int f = first + 1, l = last;
    while (true) {
      while (sel_before(a[f], pivot)) ++f;
      --l;
      while (sel_before(pivot, a[l])) --l;
      if (!(f < l)) break;
      sel_swap(a, f, l);
      ++f;
    }
    if (f <= nth) first = f; else last = f;
  }
  // __insertion_sort(first, last)
  for (int i = first + 1; i < last; ++i) {
    const SelItem v = a[i];
    if (sel_before(v, a[first])) {
      for (int j = i; j > first; --j) a[j] = a[j - 1];
      a[first] = v;
    } else {
      int j = i;
      while (sel_before(v, a[j - 1])) { a[j] = a[j - 1]; --j; }
      a[j] = v;
    }
  }
}

}  // namespace ivg
