// placeholder until the tcgen05 kernels land
#include "common.cuh"
#include "internal.h"
int b2u_tc_compiled(void) { return 0; }
int b2u_tc_conv3x3_ok(int, int, int, int) { return 0; }
int b2u_tc_wgrad_ok(int, int, int, int) { return 0; }
int b2u_tc_convt_ok(int, int, int, int) { return 0; }
#define NOTC b2u_set_error("tensor path not compiled"); return B2U_ERR_UNSUPPORTED
int b2u_tc_conv3x3(const void*, int, int, const float*, int, const float*, int, void*, int, int, double*, const void*, int, int, int, int, int, int, void*, size_t, void*) { NOTC; }
int b2u_tc_conv3x3_wgrad(const void*, int, int, const void*, int, int, float*, float*, int, int, int, void*, size_t, void*) { NOTC; }
int b2u_tc_convt_fwd(const void*, int, int, const float*, const float*, void*, int, int, int, int, int, void*, size_t, void*) { NOTC; }
int b2u_tc_convt_dgrad(const void*, int, int, const float*, void*, int, int, const void*, int, int, int, int, int, int, void*, size_t, void*) { NOTC; }
int b2u_tc_convt_wgrad(const void*, int, int, const void*, int, int, float*, float*, int, int, int, void*, size_t, void*) { NOTC; }
