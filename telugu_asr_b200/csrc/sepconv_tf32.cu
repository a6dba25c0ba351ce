// TF32 tensor-core SeparableConv1D (tcgen05) — placeholder until the kernel lands.
#include "common.cuh"
using namespace tasr;
extern "C" int tasr_sepconv_plan_create(const TasrSepConvLayer*, TasrSepConvPlan** out, tasr_stream_t) {
  if (out) *out = nullptr;
  return fail(TASR_ERR_UNSUPPORTED, "tasr_sepconv_plan_create: TF32 path not built in this revision");
}
extern "C" int tasr_sepconv_plan_destroy(TasrSepConvPlan*) { return TASR_OK; }
extern "C" int tasr_sepconv1d_tf32(const TasrSepConvPlan*, const float*, int32_t, int32_t, float*, int32_t, tasr_stream_t) {
  return fail(TASR_ERR_UNSUPPORTED, "tasr_sepconv1d_tf32: TF32 path not built in this revision");
}
