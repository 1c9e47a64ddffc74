// Binary max-heap primitives whose element ORDER under score ties is identical to
// libstdc++'s std::make_heap / std::pop_heap / std::push_heap -- the routines the
// reference decoder's list merge is built on
// (/root/reference/viterbi/viterbi_convolutional_code.cpp:762-795, comparator :54).
//
// The algorithms are restated from the published libstdc++ behaviour
// (bits/stl_heap.h: __push_heap, __adjust_heap, __pop_heap, __make_heap):
//   * sift-up moves a parent down only while parent.score <  value.score (strict);
//   * the pop/make "adjust" walks the hole all the way to a leaf, at each level taking
//     the right child unless right.score < left.score (strict), handles the lone left
//     child of an even-length heap, then sifts the saved value back up from the leaf.
// The heap compares scores only; the tag rides along.  Usable from host and device so
// that tests/test_heap_order.py can pin it against the real library on the CPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define LVA_HD __host__ __device__ __forceinline__
#else
#define LVA_HD inline
#endif

namespace lva {

struct HeapRef {
  float *score;
  uint32_t *tag;
};

LVA_HD void heap_sift_up(HeapRef h, int hole, int top, float vs, uint32_t vt) {
  int parent = (hole - 1) / 2;
  while (hole > top && h.score[parent] < vs) {
    h.score[hole] = h.score[parent];
    h.tag[hole] = h.tag[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  h.score[hole] = vs;
  h.tag[hole] = vt;
}

LVA_HD void heap_adjust(HeapRef h, int hole, int len, float vs, uint32_t vt) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (h.score[child] < h.score[child - 1]) child--;
    h.score[hole] = h.score[child];
    h.tag[hole] = h.tag[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h.score[hole] = h.score[child - 1];
    h.tag[hole] = h.tag[child - 1];
    hole = child - 1;
  }
  heap_sift_up(h, hole, top, vs, vt);
}

// std::make_heap over n elements
LVA_HD void heap_make(HeapRef h, int n) {
  if (n < 2) return;
  for (int parent = (n - 2) / 2;; parent--) {
    heap_adjust(h, parent, n, h.score[parent], h.tag[parent]);
    if (parent == 0) return;
  }
}

// std::pop_heap over n elements: afterwards the former maximum sits at index n-1 and
// [0, n-1) is a heap again.
LVA_HD void heap_pop(HeapRef h, int n) {
  if (n < 2) return;
  float vs = h.score[n - 1];
  uint32_t vt = h.tag[n - 1];
  h.score[n - 1] = h.score[0];
  h.tag[n - 1] = h.tag[0];
  heap_adjust(h, 0, n - 1, vs, vt);
}

// std::push_heap over n elements: element n-1 is the newcomer.
LVA_HD void heap_push(HeapRef h, int n) { heap_sift_up(h, n - 1, 0, h.score[n - 1], h.tag[n - 1]); }

}  // namespace lva
