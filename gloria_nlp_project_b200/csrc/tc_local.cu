// bf16 tensor-core mode (tcgen05 / TMEM / TMA) -- placeholder entry points until the kernels land.
#include "common.cuh"

using namespace gloria;

extern "C" int gloria_b200_tc_spad(int S) { return (S + 127) / 128 * 128; }
extern "C" int gloria_b200_tc_lpad(int Lcap) { return (Lcap + 15) / 16 * 16; }
extern "C" int gloria_b200_tc_supported(int D, int S, int Lcap) {
  (void)D; (void)S; (void)Lcap;
  return GLORIA_ERR_UNSUPPORTED;
}
extern "C" int gloria_b200_tc_prepack(const float*, const float*, const int32_t*, int, int, int, int, int, int, int,
                                      void*, void*, void*, float*, void*) {
  return fail(GLORIA_ERR_UNSUPPORTED, "tensor-core path not built yet");
}
extern "C" size_t gloria_b200_tc_workspace(int, int, int, int, int) { return 0; }
extern "C" int gloria_b200_tc_local_sim_fwd(const void*, const void*, const void*, const float*, const int32_t*, int,
                                            int, int, int, int, float, float, int, float, float*, float*, float*,
                                            void*, size_t, void*) {
  return fail(GLORIA_ERR_UNSUPPORTED, "tensor-core path not built yet");
}
extern "C" int gloria_b200_tc_local_sim_bwd(const void*, const void*, const void*, const float*, const int32_t*, int,
                                            int, int, int, int, int, int, float, float, int, float, const float*,
                                            const float*, const float*, float*, float*, void*, size_t, void*) {
  return fail(GLORIA_ERR_UNSUPPORTED, "tensor-core path not built yet");
}
