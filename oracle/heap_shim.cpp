// TEST INFRASTRUCTURE ONLY: host build of the product's device heap primitives
// (nanopore_dna_storage_b200/csrc/heap_order.cuh) so tests/test_heap_order.py can pin them against
// the real libstdc++ routines (lvo_std_heap_op in lva_oracle.cpp) on the CPU.
#include "../nanopore_dna_storage_b200/csrc/heap_order.cuh"

extern "C" void lva_heap_op(int op, float *score, uint32_t *tag, uint32_t n) {
  lva::HeapRef h{score, tag};
  if (op == 0) lva::heap_make(h, (int)n);
  else if (op == 1) lva::heap_pop(h, (int)n);
  else lva::heap_push(h, (int)n);
}
