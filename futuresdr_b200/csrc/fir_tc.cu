// fir_tc.cu -- tcgen05 block-Toeplitz FIR (placeholder until the tensor-core path lands).
#include "fir.cuh"

bool fir_tc_supported(const b2s_fir *) { return false; }
int32_t fir_tc_prepare(b2s_fir *f) { return b2s_fail(f->ctx, B2S_EUNSUPPORTED, "tensor path not built"); }
int32_t fir_tc_launch(b2s_fir *f, const void *, size_t, void *, size_t, cudaStream_t) {
    return b2s_fail(f->ctx, B2S_EUNSUPPORTED, "tensor path not built");
}
void fir_tc_release(b2s_fir *) {}
