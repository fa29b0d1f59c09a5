// TEMPORARY stubs (replaced by lsqr.cu / conic.cu)
#include "common.cuh"
extern "C" {
#define NI(ctx) do { if (ctx) (ctx)->err = "not implemented yet"; return -5; } while (0)
int32_t diffopt_b200_lsqr_csc(diffopt_b200_ctx* ctx, int64_t, int64_t, const int64_t*, const int64_t*, const double*, int32_t, const double*, double, double, double, int64_t, double*, double*, int32_t) { NI(ctx); }
int32_t diffopt_b200_conic_setup(diffopt_b200_ctx* ctx, int64_t, int64_t, const int64_t*, const int64_t*, const double*, const double*, const double*, const double*, const double*, const double*, int64_t, const int32_t*, const int64_t*, int32_t) { NI(ctx); }
int32_t diffopt_b200_conic_get_vp(diffopt_b200_ctx* ctx, double*, int32_t) { NI(ctx); }
int32_t diffopt_b200_conic_dpi_apply(diffopt_b200_ctx* ctx, const double*, int32_t, double*, int32_t) { NI(ctx); }
int32_t diffopt_b200_conic_M_apply(diffopt_b200_ctx* ctx, const double*, int32_t, double*, int32_t) { NI(ctx); }
int32_t diffopt_b200_conic_forward(diffopt_b200_ctx* ctx, int64_t, const int64_t*, const int64_t*, const double*, const double*, const double*, double, double, double, int64_t, double*, double*, double*, int32_t) { NI(ctx); }
int32_t diffopt_b200_conic_reverse(diffopt_b200_ctx* ctx, const double*, double, double, double, int64_t, double*, double*, double*, double*, int32_t) { NI(ctx); }
}
