// tcgen05 / TMEM fused dense chain (fp16 hi/lo split, fp32 accumulate) - see DESIGN.md.
#include "bb_common.cuh"

int bb_tc_prepare(bb_ctx*, Chain* c) {
  c->tc_ok = false;
  return BB_ERR_UNSUPPORTED;
}

int bb_tc_launch(bb_ctx*, const Chain*, const void*, int, int64_t, const float*, const float*, const float*,
                 const float*, void*, int, int, int*, cudaStream_t) {
  return BB_ERR_UNSUPPORTED;
}
