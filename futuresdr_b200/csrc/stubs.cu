// stubs.cu -- entry points not implemented yet (temporary; each returns B2S_EUNSUPPORTED).
#include "common.cuh"
#define NOTYET(ctx) b2s_fail((ctx), B2S_EUNSUPPORTED, "%s: not implemented yet", __func__)
extern "C" {
int32_t b2s_resamp_plan(b2s_ctx *c, b2s_kind, const float *, size_t, size_t, size_t, b2s_resamp **) { return NOTYET(c); }
void b2s_resamp_destroy(b2s_resamp *) {}
size_t b2s_resamp_length(const b2s_resamp *) { return 0; }
int32_t b2s_resamp_exec(b2s_resamp *, const void *, size_t, void *, size_t, size_t *, size_t *, int32_t *) { return NOTYET(nullptr); }
int32_t b2s_pfbarb_plan_c32(b2s_ctx *c, const float *, size_t, size_t, float, b2s_pfbarb **) { return NOTYET(c); }
void b2s_pfbarb_destroy(b2s_pfbarb *) {}
int32_t b2s_pfbarb_reset(b2s_pfbarb *) { return NOTYET(nullptr); }
int32_t b2s_pfbarb_exec(b2s_pfbarb *, const void *, size_t, void *, size_t, size_t *, size_t *, int32_t *) { return NOTYET(nullptr); }
int32_t b2s_fft_plan_c32(b2s_ctx *c, size_t, int32_t, int32_t, int32_t, float, b2s_fft **) { return NOTYET(c); }
void b2s_fft_destroy(b2s_fft *) {}
size_t b2s_fft_length(const b2s_fft *) { return 0; }
int32_t b2s_fft_exec(b2s_fft *, const void *, size_t, void *, size_t, size_t *, size_t *) { return NOTYET(nullptr); }
int32_t b2s_apply_create(b2s_ctx *c, b2s_op, float, b2s_apply **) { return NOTYET(c); }
void b2s_apply_destroy(b2s_apply *) {}
int32_t b2s_apply_reset(b2s_apply *) { return NOTYET(nullptr); }
int32_t b2s_apply_exec(b2s_apply *, const void *, size_t, void *, size_t, size_t *, size_t *) { return NOTYET(nullptr); }
int32_t b2s_ring_create(b2s_ctx *c, size_t, size_t, size_t, int32_t, int32_t, b2s_ring **) { return NOTYET(c); }
void b2s_ring_destroy(b2s_ring *) {}
int32_t b2s_ring_acquire_empty(b2s_ring *, b2s_slot **) { return NOTYET(nullptr); }
int32_t b2s_ring_submit_full(b2s_ring *, b2s_slot *, size_t, int32_t) { return NOTYET(nullptr); }
int32_t b2s_ring_acquire_full(b2s_ring *, b2s_slot **, size_t *) { return NOTYET(nullptr); }
int32_t b2s_ring_release(b2s_ring *, b2s_slot *) { return NOTYET(nullptr); }
int32_t b2s_ring_carry_halo(b2s_ring *, const b2s_slot *, size_t, size_t, b2s_slot *) { return NOTYET(nullptr); }
void *b2s_slot_device_ptr(const b2s_slot *) { return nullptr; }
void *b2s_slot_host_ptr(const b2s_slot *) { return nullptr; }
size_t b2s_slot_halo_valid(const b2s_slot *) { return 0; }
int32_t b2s_slot_fetch_to_host(b2s_slot *, size_t) { return NOTYET(nullptr); }
int32_t b2s_slot_wait(b2s_slot *) { return NOTYET(nullptr); }
size_t b2s_ring_free_slots(const b2s_ring *) { return 0; }
size_t b2s_ring_full_slots(const b2s_ring *) { return 0; }
}
