// tcgen05 tile kernel for the fused local-attention forward (placeholder until
// the tile kernel lands: reports "not supported" so the gather kernel serves
// every call).
#include "common.cuh"
namespace gfla {
bool local_attn_fwd_tc_supported(int, int, int, int, int, int, int, int, int, const void*, const void*) { return false; }
int local_attn_fwd_tc(const void*, const void*, const void*, void*, void*, int, int, int, int, int, int, int, int, int,
                      cudaStream_t) { return GFLA_E_NOTSUP; }
}  // namespace gfla
