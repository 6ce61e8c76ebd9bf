// jds_kernels.cuh - host-callable launchers of the staged kernels (jds_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "jds_internal.cuh"

namespace jds {

// strides are in elements of the buffer's type, per unit (0 = shared by all units)
void launch_forward(bool exact, const Geom& g, int prefilter, const uint8_t* rgb,
                    size_t rgb_stride, void* fwd, size_t fwd_stride, int units, cudaStream_t s);
void launch_codec(bool exact, const Geom& g, const void* fwd, size_t fwd_stride, void* rec,
                  size_t rec_stride, const QTables* tables, int table_stride, int16_t* coeffs,
                  size_t coeff_stride, bool hist, DevMetrics* metrics, int units, cudaStream_t s);
void launch_inverse(bool exact, const Geom& g, const uint8_t* rgb, size_t rgb_stride,
                    const void* fwd, size_t fwd_stride, const void* rec, size_t rec_stride,
                    uint8_t* recon, size_t recon_stride, double* err_y, double* err_rgb,
                    DevMetrics* metrics, int units, cudaStream_t s);
void launch_ssim(bool exact, int H, int W, const uint8_t* a, size_t a_stride, const uint8_t* b,
                 size_t b_stride, DevMetrics* metrics, int units, cudaStream_t s);
void launch_selected_block(const Geom& g, const uint8_t* rgb, int bx, int by,
                           const QTables* tables, void* out, cudaStream_t s);
size_t selected_out_bytes();

}  // namespace jds
