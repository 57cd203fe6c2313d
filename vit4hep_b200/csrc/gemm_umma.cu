// placeholder until the tcgen05 GEMM lands: reports "unsupported" so callers use the SIMT GEMM
#include "kernels.cuh"
namespace v4h {
struct UmmaContext { int unused; };
UmmaContext* umma_context_create() { return new UmmaContext(); }
void umma_context_destroy(UmmaContext* c) { delete c; }
bool gemm_umma_supported(const GemmDesc&) { return false; }
int gemm_umma(UmmaContext*, const GemmDesc&, cudaStream_t) {
  return fail(V4H_ERR_UNSUPPORTED, "gemm_umma: not built");
}
}  // namespace v4h
