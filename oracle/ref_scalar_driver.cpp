// ref_scalar_driver.cpp -- TEST INFRASTRUCTURE.  Thin array loops (ours) around the REFERENCE's
// own per-element templates, included from where they lie (-I/root/reference/torchlsq/csrc/ops):
//   kernels/lsq_kernel.h  (which includes ../global_scope.h)
// Built by oracle/Makefile into oracle/_ref/libref_scalar.so.  The host-side constants follow
// cpu/lsq_cpu.cpp:40-47 (per-tensor s, inv_s) and :103 / :250 (grad scaler).
#include <cmath>
#include <cstdint>
#include <limits>
#include <tuple>

#include "kernels/lsq_kernel.h"

extern "C" {

void ref_fwd_tensor_f32(const float* x, float* y, int64_t n, float scale, float shift, int64_t qmin, int64_t qmax,
                        int64_t tmin, int64_t tmax, int init_mode) {
    const float b = shift;
    const float s = std::max(std::abs(scale), std::numeric_limits<float>::epsilon());
    const float inv_s = 1.0f / s;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++)
        y[i] = lsq_forward_kernel_per_tensor<float>(x[i], s, inv_s, b, (float)qmin, (float)qmax, (float)tmin,
                                                    (float)tmax, init_mode != 0);
}

// per-element outputs (dx, ds_i * gs, db_i * gs), exactly what the reference stores before at::sum
void ref_bwd_tensor_f32(const float* g, const float* x, float* dx, float* ds, float* db, int64_t n, float scale,
                        float shift, int64_t qmin, int64_t qmax, int64_t tmin, int64_t tmax, float grad_scaler,
                        int sym, int init_mode) {
    const float b = shift;
    const float s = std::max(std::abs(scale), std::numeric_limits<float>::epsilon());
    const float inv_s = 1.0f / s;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        auto r = lsq_backward_kernel_per_tensor<float>(g[i], x[i], s, inv_s, b, (float)qmin, (float)qmax, (float)tmin,
                                                       (float)tmax, grad_scaler, sym != 0, init_mode != 0);
        dx[i] = std::get<0>(r); ds[i] = std::get<1>(r); db[i] = std::get<2>(r);
    }
}

void ref_fwd_channel_f32(const float* x, float* y, int64_t outer, int64_t C, int64_t inner, const float* scale,
                         const float* shift, int64_t qmin, int64_t qmax, int64_t tmin, int64_t tmax, int init_mode) {
    const float eps = std::numeric_limits<float>::epsilon();
#pragma omp parallel for collapse(2)
    for (int64_t o = 0; o < outer; o++)
        for (int64_t c = 0; c < C; c++)
            for (int64_t i = 0; i < inner; i++) {
                const int64_t k = (o * C + c) * inner + i;
                y[k] = lsq_forward_kernel_per_channel<float>(x[k], scale[c], shift[c], (float)qmin, (float)qmax,
                                                             (float)tmin, (float)tmax, init_mode != 0, eps);
            }
}

void ref_bwd_channel_f32(const float* g, const float* x, float* dx, float* ds, float* db, int64_t outer, int64_t C,
                         int64_t inner, const float* scale, const float* shift, int64_t qmin, int64_t qmax,
                         int64_t tmin, int64_t tmax, float grad_scaler, int sym, int init_mode) {
    const float eps = std::numeric_limits<float>::epsilon();
#pragma omp parallel for collapse(2)
    for (int64_t o = 0; o < outer; o++)
        for (int64_t c = 0; c < C; c++)
            for (int64_t i = 0; i < inner; i++) {
                const int64_t k = (o * C + c) * inner + i;
                auto r = lsq_backward_kernel_per_channel<float>(g[k], x[k], scale[c], shift[c], (float)qmin, (float)qmax,
                                                                (float)tmin, (float)tmax, grad_scaler, sym != 0,
                                                                init_mode != 0, eps);
                dx[k] = std::get<0>(r); ds[k] = std::get<1>(r); db[k] = std::get<2>(r);
            }
}

}  // extern "C"
