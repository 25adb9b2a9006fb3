// placeholder: CLAHE / crop-resize kernels land in a later commit
#include "common.cuh"
extern "C" int b2u_clahe_u8(const uint8_t*, uint8_t*, int, int, int, float, int, void*, size_t, void*) {
  b2u_set_error("clahe: not built yet"); return B2U_ERR_UNSUPPORTED; }
extern "C" int b2u_crop_resize(const uint8_t*, int, int, int, const int*, int, int, int, uint8_t*, float*, void*) {
  b2u_set_error("crop_resize: not built yet"); return B2U_ERR_UNSUPPORTED; }
