// jds_kernels.cuh - host-callable launchers of the staged kernels (jds_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "jds_internal.cuh"

namespace jds {

// Per-device one-time setup, called from jds_ctx_create for the context's device: dynamic
// shared-memory opt-ins / carve-outs and constant tables.  Idempotent and free of shared
// mutable state, so contexts on different threads / devices never race.
cudaError_t fused_configure_device();      // jds_fused.cu
cudaError_t ssim_configure_device();       // jds_ssim.cu
cudaError_t entropy_configure_device();    // jds_entropy.cu
cudaError_t exact_fused_configure_device();   // jds_fused_exact.cu

// strides are in elements of the buffer's type, per unit (0 = shared by all units)
void launch_forward(bool exact, const Geom& g, int prefilter, const uint8_t* rgb,
                    size_t rgb_stride, void* fwd, size_t fwd_stride, int units, cudaStream_t s,
                    bool chroma_only = false);
void launch_codec(bool exact, const Geom& g, const void* fwd, size_t fwd_stride, void* rec,
                  size_t rec_stride, const QTables* tables, int table_stride, int16_t* coeffs,
                  size_t coeff_stride, bool hist, DevMetrics* metrics, int units, cudaStream_t s,
                  bool chroma_only = false);
void launch_inverse(bool exact, const Geom& g, const uint8_t* rgb, size_t rgb_stride,
                    const void* fwd, size_t fwd_stride, const void* rec, size_t rec_stride,
                    uint8_t* recon, size_t recon_stride, double* err_y, double* err_rgb,
                    DevMetrics* metrics, int units, cudaStream_t s);
void launch_ssim(bool exact, int H, int W, const uint8_t* a, size_t a_stride, const uint8_t* b,
                 size_t b_stride, DevMetrics* metrics, int units, cudaStream_t s);
// strip-streaming SSIM / SSE kernel (jds_ssim.cu); needs W % 16 == 0 and 16-byte aligned
// image bases / unit strides (TMA bulk copies)
bool ssim_strip_supported(int H, int W, const void* a, size_t a_stride, const void* b,
                          size_t b_stride);
cudaError_t launch_ssim_strip(int H, int W, const uint8_t* a, size_t a_stride, const uint8_t* b,
                              size_t b_stride, DevMetrics* metrics, int units, bool want_ssim,
                              bool want_sse, int sm_count, cudaStream_t s);
// fused fast-mode codec kernels (jds_fused.cu)
bool fused_supported(const Geom& g, int prefilter, const void* rgb, size_t rgb_stride,
                     const void* recon, size_t recon_stride);
size_t fused_chroma_plane_floats(const Geom& g);
cudaError_t launch_fused_chroma(const Geom& g, int prefilter, const uint8_t* rgb, size_t rgb_stride,
                                float* cplanes, size_t cplane_stride, const QTables* tables,
                                int table_stride, int16_t* coeffs, size_t coeff_stride,
                                DevMetrics* metrics, int units, cudaStream_t s, int stage = 0,
                                float* fcoef = nullptr);
// hoisted sweeps: `stage` 1 writes the quality-independent forward DCT coefficients of one
// frame into `fcoef` (fused_fcoef_floats(g) floats; 0 = this geometry is not hoisted), `stage` 2
// runs quantisation .. reconstruction of `units` quality points from it; 0 = the full kernels
size_t fused_fcoef_floats(const Geom& g);
cudaError_t launch_fused_luma(const Geom& g, const uint8_t* rgb, size_t rgb_stride,
                              const float* cplanes, size_t cplane_stride, const QTables* tables,
                              int table_stride, int16_t* coeffs, size_t coeff_stride,
                              uint8_t* recon, size_t recon_stride, DevMetrics* metrics, int units,
                              cudaStream_t s, int stage = 0, float* fcoef = nullptr);
// fused exact-mode luma + compose kernel (jds_fused_exact.cu): Y never touches HBM; `rec` holds
// the reconstructed chroma planes the chroma-only k_codec wrote (plane layout of the staged path)
cudaError_t launch_exact_luma(const Geom& g, const uint8_t* rgb, size_t rgb_stride, const double* rec,
                              size_t rec_stride, const QTables* tables, int table_stride,
                              int16_t* coeffs, size_t coeff_stride, uint8_t* recon,
                              size_t recon_stride, DevMetrics* metrics, int units, cudaStream_t s);
bool exact_chroma_supported(const Geom& g, int prefilter);
cudaError_t launch_exact_chroma(const Geom& g, int prefilter, const uint8_t* rgb, size_t rgb_stride, double* rec,
                                size_t rec_stride, const QTables* tables, int table_stride,
                                int16_t* coeffs, size_t coeff_stride, DevMetrics* metrics, int units,
                                cudaStream_t s);
void launch_hist50(const int16_t* coeffs, size_t coeff_stride, size_t n_coeffs, DevMetrics* metrics,
                   int units, int sm_count, cudaStream_t s);
constexpr int VALUE_HIST_BINS = 2048;          // int16 coefficient value v -> bin v + 1024
void launch_value_hist(const int16_t* coeffs, size_t n_coeffs, unsigned long long* hist,
                       int sm_count, cudaStream_t s);
void launch_heat_u8(const double* err, uint8_t* out, size_t n, cudaStream_t s);
int launch_resize_area_u8(int H, int W, int dh, int dw, const uint8_t* src, uint8_t* dst,
                          cudaStream_t s);
void launch_color_f64(int direction, long long n, const double* in, double* out, cudaStream_t s);
void launch_subsample_plane(const double* p, int H, int W, int hc, int wc, int sub, int prefilter,
                            double* out, cudaStream_t s);
void launch_upsample_plane(const double* p, int h, int w, int H, int W, double* out, cudaStream_t s);
void launch_sse_u8(const uint8_t* a, const uint8_t* b, long long n_px, DevMetrics* m, int sm_count,
                   cudaStream_t s);
void launch_bitcount(const int16_t* c, unsigned long long n, DevMetrics* m, int sm_count, cudaStream_t s);
enum { JDS_BLOCKOP_DCT2 = 0, JDS_BLOCKOP_IDCT2 = 1, JDS_BLOCKOP_ENCODE = 2, JDS_BLOCKOP_DECODE = 3,
       JDS_BLOCKOP_QUANTIZE = 4, JDS_BLOCKOP_DEQUANTIZE = 5 };
void launch_block_ops(int op, long long n_blocks, const double* in, const int16_t* in_q,
                      const double* qtable, double* out, int16_t* out_q, cudaStream_t s);
void launch_selected_block(const Geom& g, const uint8_t* rgb, int bx, int by,
                           const QTables* tables, void* out, cudaStream_t s);
size_t selected_out_bytes();
// baseline-JPEG entropy-coded size of a coefficient array (jds_entropy.cu): bits of the three
// non-interleaved scans; blocks [0,ny) Y, then nc Cb, then nc Cr
cudaError_t launch_entropy_bits(const int16_t* coeffs, long long ny, long long nc,
                                unsigned long long* scan_bits, cudaStream_t s);
// the bitstream of those scans (jds_entropy.cu).  Every scan owns a run of CTAs of
// ENTROPY_CTA_BLOCKS blocks each.
constexpr int ENTROPY_CTA_BLOCKS = 64;
struct EntropyGrid {
    long long n[3];       // blocks of the scan
    long long first[3];   // index of its first block in the coefficient array
    int cta0[3];          // its first CTA
    int ctas;
};
struct EntropyLayout {                    // device-resident result of the size passes
    unsigned long long bits[3];           // entropy-coded bits of each scan
    unsigned long long ubytes[3];         // ceil(bits / 8)
    unsigned long long ubase[3];          // byte offset of the scan in the unstuffed buffer
    unsigned long long ff[3];             // 0xFF bytes of the scan (each gets a stuffed 0x00)
    unsigned long long total_ubytes;      // end of the last scan, rounded up to 4
    unsigned long long stuffed_bytes;     // sum of ubytes + ff
    unsigned int invalid;                 // a value outside the baseline code tables was seen
    unsigned int pad;
};
EntropyGrid make_entropy_grid(long long ny, long long nc);
cudaError_t launch_entropy_sizes(const int16_t* coeffs, const EntropyGrid& g, uint32_t* blk_bits,
                                 uint32_t* part, unsigned long long* cta_off, EntropyLayout* lay,
                                 cudaStream_t s);
cudaError_t launch_entropy_pack(const int16_t* coeffs, const EntropyGrid& g, const uint32_t* blk_bits,
                                const unsigned long long* cta_off, const EntropyLayout* lay,
                                uint32_t* ubuf, size_t ubuf_bytes, cudaStream_t s);
cudaError_t launch_stuff_sizes(const uint32_t* ubuf, size_t ubuf_bytes, EntropyLayout* lay, uint32_t* cnt,
                               unsigned long long* chunk_off, cudaStream_t s);
cudaError_t launch_stuff_scatter(const uint32_t* ubuf, size_t ubuf_bytes, const EntropyLayout* lay,
                                 const unsigned long long* chunk_off, uint8_t* out, cudaStream_t s);
size_t jfif_write_headers(uint8_t* out, int height, int width, int sub, const uint8_t q_raster[64]);
size_t jfif_write_sos(uint8_t* out, int comp);
// chroma-aliasing demo front end (jds_alias.cu)
size_t alias_scratch_floats(int H, int W);
int launch_alias_subsample(int H, int W, int prefilter, const uint8_t* rgb, float* scratch,
                           uint8_t* out, cudaStream_t s);     // returns the kernels launched
void launch_alias_luma3(size_t n_px, const uint8_t* rgb, uint8_t* out, cudaStream_t s);
void launch_alias_diff(size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out, cudaStream_t s);

}  // namespace jds
