// tcgen05 attention entry (placeholder until the TMEM flash kernel lands in this file):
// reports "unsupported" so the dispatcher in lnx_attn_simt.cu uses the CUDA-core kernel.
#include "lnx_common.cuh"

int lnx_attn_fwd_tc(const void*, const void*, const void*, void*, float*, int, int, int, int, cudaStream_t) {
  return LNX_ERR_UNSUPPORTED;
}
