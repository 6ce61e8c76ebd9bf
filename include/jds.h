/*
 * jds.h - C ABI of libjds.so, the B200 (sm_100a) implementation of JPEG-DSP Studio's
 * compression round trip.
 *
 * The reference has no FFI of its own: its operator boundary for this path is the
 * Python function
 *     engines/pipeline.py:17-21   compress_reconstruct(image_rgb, params, selected_block_idx)
 * called from gui/worker.py:29 (single run), gui/worker.py:68 (quality sweep) and
 * main.py:81 (CLI).  This header is what a ctypes/cffi binding for that call binds
 * (see INTEGRATION.md for the stub): plain pointers and sizes, no Python.h, no torch
 * types, status codes instead of exceptions.  jpeg_dsp_studio_b200/_native.py is
 * that binding.
 *
 * Conventions
 *   - every entry point returns 0 (JDS_OK) or a negative jds_status; the message for
 *     the calling thread's last failure is jds_last_error().
 *   - image buffers are packed RGB uint8, row-major H x W x 3, no row padding.
 *   - a buffer argument may be a host pointer or a device pointer; the matching
 *     `*_loc` argument says which (JDS_HOST / JDS_DEVICE).  Host buffers are copied
 *     with cudaMemcpyAsync on the context's stream inside the call.
 *   - all work of a context is issued on one CUDA stream (jds_ctx_set_stream to use
 *     the caller's); calls return after that stream has been synchronised, so the
 *     outputs and the metrics struct are valid on return.
 *   - no entry point falls back to the CPU: without a CUDA device every compute
 *     entry point fails with JDS_ERR_CUDA.
 */
#ifndef JDS_H_
#define JDS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JDS_ABI_VERSION 1

typedef enum jds_status {
    JDS_OK = 0,
    JDS_ERR_INVALID = -1,      /* bad argument (NULL, size, quality, mode ...)        */
    JDS_ERR_UNSUPPORTED = -2,  /* valid for the reference but not implemented (odd
                                  plane sizes with chroma subsampling)               */
    JDS_ERR_CUDA = -3,         /* CUDA runtime error; text in jds_last_error()       */
    JDS_ERR_NOMEM = -4,
    JDS_ERR_CAPACITY = -5      /* caller's output buffer too small; the needed size is
                                  in the size outputs of the call                    */
} jds_status;

/* subsampling_mode of models/compression_params.py:13 */
enum { JDS_SUB_444 = 0, JDS_SUB_422 = 1, JDS_SUB_420 = 2 };

/* arithmetic mode (BASELINE.json north_star) */
enum {
    JDS_EXACT = 0, /* fp64, operation order of the reference's NumPy/SciPy/OpenCV calls:
                      int16 coefficients and uint8 pixels bit-identical             */
    JDS_FAST = 1   /* fp32, fused arithmetic; PSNR within 1e-3 dB, SSIM within 1e-5  */
};

enum { JDS_HOST = 0, JDS_DEVICE = 1 };

/* what to produce besides the reconstructed image (IntermediateData fields,
 * models/intermediate_data.py:8-23, filled at engines/pipeline.py:119-124) */
enum {
    JDS_OUT_RECON   = 1u << 0, /* reconstructed_image, uint8 H x W x 3                */
    JDS_OUT_COEFFS  = 1u << 1, /* all_quantized_coeffs, int16, Y|Cb|Cr, block raster  */
    JDS_OUT_ERR_Y   = 1u << 2, /* error_map_y, fp64 H x W                             */
    JDS_OUT_ERR_RGB = 1u << 3, /* error_map_rgb, fp64 H x W                           */
    JDS_OUT_HIST    = 1u << 4, /* quantized_histogram (50 bins over [-100,100])       */
    JDS_OUT_SSIM    = 1u << 5, /* SSIM partial sums (utils/metrics.py:12-14,21)       */
    JDS_OUT_PSNR    = 1u << 6  /* SSE partial sums (utils/metrics.py:11,20); always on */
};

/* CompressionParams (models/compression_params.py:7-20) plus the frame geometry. */
typedef struct jds_params {
    int32_t height;
    int32_t width;
    int32_t quality;      /* 1..100                                                  */
    int32_t subsampling;  /* JDS_SUB_*                                               */
    int32_t prefilter;    /* use_prefilter: 3x3 sigma=0.75 Gaussian before decimation */
    int32_t precision;    /* JDS_EXACT / JDS_FAST                                    */
    uint32_t outputs;     /* JDS_OUT_* flags                                         */
    int32_t reserved;
} jds_params;

/*
 * Metric partials of one round trip.  The host turns them into the floats of
 * CompressionResult exactly as utils/metrics.py does:
 *   psnr_rgb = 10 log10(255^2 / (sse_rgb / (3 H W)))           (metrics.py:11)
 *   psnr_y   = 10 log10(255^2 / (sse_y / (H W)))               (metrics.py:17-20)
 *   ssim_rgb = mean_c(ssim_sum[c] / ssim_count), c = R,G,B     (metrics.py:12-14)
 *   ssim_y   = ssim_sum[3] / ssim_count                        (metrics.py:21)
 *   bits     = 2 * luma_blocks + coeff_bits                    (metrics.py:63-85)
 */
typedef struct jds_metrics {
    uint64_t sse_rgb;       /* sum (a-b)^2 over the 3*H*W uint8 samples, exact          */
    double   sse_y;         /* sum (Ya-Yb)^2, Y = .299r+.587g+.114b of the uint8 images */
    double   ssim_sum[4];   /* sum of the SSIM map over the (H-6)(W-6) windows: R,G,B,Y */
    uint64_t ssim_count;    /* (H-6)*(W-6)                                              */
    uint64_t coeff_bits;    /* 6*nnz + sum_{v!=0} (bit_length(|v|) + 1), exact          */
    uint64_t nnz;           /* non-zero quantised coefficients                          */
    uint64_t total_coeffs;  /* 64 * (blocks of Y + Cb + Cr)                             */
    uint64_t luma_blocks;   /* ceil(H/8)*ceil(W/8)                                      */
    int64_t  hist50[50];    /* np.histogram(coeffs, 50, (-100,100)) (pipeline.py:124)   */
    double   gpu_ms;        /* device time of this unit's kernels (CUDA events)         */
    uint64_t reserved[3];
} jds_metrics;

typedef struct jds_ctx jds_ctx;

/* ---- library / context ------------------------------------------------------- */
int jds_abi_version(void);
const char* jds_last_error(void);
int jds_device_count(int* count);
int jds_ctx_create(int device, jds_ctx** ctx);
int jds_ctx_destroy(jds_ctx* ctx);
/* use the caller's cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); 0 = default */
int jds_ctx_set_stream(jds_ctx* ctx, void* cuda_stream);
int jds_ctx_synchronize(jds_ctx* ctx);
/* stream ordering against the caller's own streams (cudaEvent_t handles, same device):
 * wait_event - everything this context enqueues afterwards waits for the event (frames the
 * caller produced on another stream); record_event - records the event behind the context's
 * work so far (results of the non-synchronising *_records entry points) */
int jds_ctx_wait_event(jds_ctx* ctx, void* cuda_event);
int jds_ctx_record_event(jds_ctx* ctx, void* cuda_event);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int jds_ctx_launch_count(jds_ctx* ctx, uint64_t* launches);

/* accumulated device time (ms, CUDA events on the context's stream) and launch counts of
 * the four stage kernels - forward colour, block codec, inverse colour, SSIM - since
 * the last reset; bench.py's roofline line is computed from these */
int jds_ctx_stage_times(jds_ctx* ctx, double ms[4], uint64_t launches[4], int reset);
/* per-kernel events are recorded only while enabled (off by default: ~20 us per call) */
int jds_ctx_stage_timing(jds_ctx* ctx, int enable);

/* ---- host-only helpers (no GPU needed) ---------------------------------------- */
/* engines/quantizer.py:7-19 scale_quant_matrix(JPEG_LUMA_Q50, quality) -> 64 doubles */
int jds_quant_table(int quality, double table[64]);
/* element counts of the outputs for a frame geometry (padding per block_processor.py:7-16) */
int jds_coeff_count(int height, int width, int subsampling, uint64_t* count);
int jds_plane_dims(int height, int width, int subsampling, int* chroma_h, int* chroma_w);

/* ---- the round trip ------------------------------------------------------------- */
/*
 * One frame: replaces engines/pipeline.py:17-167 (everything except the six 8x8
 * selected-block arrays, which jds_selected_block produces).
 * rgb    [in]  H*W*3 uint8.
 * recon  [out] H*W*3 uint8                        (JDS_OUT_RECON, else may be NULL)
 * coeffs [out] jds_coeff_count() int16            (JDS_OUT_COEFFS, else NULL)
 * err_y, err_rgb [out] H*W fp64 each              (JDS_OUT_ERR_*, else NULL)
 * metrics [out] host struct, always filled (fields not requested stay 0).
 */
int jds_roundtrip(jds_ctx* ctx, const jds_params* params,
                  const uint8_t* rgb, int rgb_loc,
                  uint8_t* recon, int16_t* coeffs, double* err_y, double* err_rgb,
                  int out_loc, jds_metrics* metrics);

/*
 * A batch of n_frames frames of identical geometry and parameters, frame k at
 * rgb + k*H*W*3 (BASELINE.json config 5: frames sharded across GPUs by the caller).
 * recon/coeffs are batch-strided the same way or NULL; metrics is an array of
 * n_frames host structs.  Error maps are not produced in batch mode.
 */
int jds_roundtrip_batch(jds_ctx* ctx, const jds_params* params, int n_frames,
                        const uint8_t* rgb, int rgb_loc,
                        uint8_t* recon, int16_t* coeffs, int out_loc,
                        jds_metrics* metrics);

/*
 * jds_roundtrip_batch split in two, for callers that stream batch after batch: begin enqueues the
 * batch's copies and kernels and returns; jds_ctx_finish waits for them and fills `metrics`.
 * With two contexts used alternately the PCIe transfers of consecutive batches overlap (the
 * drain of one hides the fill of the next).  rgb / recon / coeffs / metrics must stay valid until
 * jds_ctx_finish; no other call on the context in between (JDS_ERR_INVALID).  jds_ctx_finish
 * without a pending batch is a no-op.
 */
int jds_roundtrip_batch_begin(jds_ctx* ctx, const jds_params* params, int n_frames,
                              const uint8_t* rgb, int rgb_loc,
                              uint8_t* recon, int16_t* coeffs, int out_loc,
                              jds_metrics* metrics);
int jds_ctx_finish(jds_ctx* ctx);

/*
 * Tile-band sharding of ONE frame (second row): rows [row0, row1) of the frame as
 * one rank's share of engines/pipeline.py:17-167, for single-image latency on several GPUs.
 * params->height / width describe the WHOLE frame and rgb points at its first row (host or
 * device; only the band and its halo are read / transferred).  row0 and row1 are multiples of
 * 16 (row1 may also be the frame height).  The library runs the band with the halo of MCU rows
 * that makes every owned pixel, coefficient and SSIM window identical to the whole-frame run
 * (16 rows; 32 with the prefilter), then reduces over the band's OWN rows only:
 *   recon_rows  [out] (row1-row0)*W*3 uint8: rows row0..row1 of reconstructed_image
 *   coeffs_rows [out] the band's blocks, Y | Cb | Cr, each in block raster order: block rows
 *               [row0/8, row1/8) of Y and [row0/8v, row1/8v) of Cb, Cr (v = 2 at 4:2:0)
 *   metrics     [out] partial sums: sse over the band's rows, SSIM over the window centres in
 *               the band (rows 3..H-4 of the frame), bits / nnz / blocks of the band's blocks.
 * Adding the partials of all bands field by field gives the jds_metrics of jds_roundtrip on the
 * whole frame (integers exactly; the fp64 sums up to summation order).  Error maps and hist50
 * are whole-frame outputs (JDS_ERR_UNSUPPORTED), as is 4:2:0 with an odd height unless the band
 * is the whole frame (cv2's area taps then depend on the frame height).
 */
int jds_roundtrip_band(jds_ctx* ctx, const jds_params* params, const uint8_t* rgb, int rgb_loc,
                       int row0, int row1, uint8_t* recon_rows, int16_t* coeffs_rows, int out_loc,
                       jds_metrics* metrics);

/*
 * Quality sweep on one frame (gui/worker.py:55-74 BatchSweepWorker.run): the same
 * frame at each quality in qualities[0..n_q); params->quality is ignored.
 * recon is n_q frames or NULL (the sweep's consumers read only the metrics,
 * gui/compression_tab.py:752-754); metrics is an array of n_q host structs.
 */
int jds_sweep(jds_ctx* ctx, const jds_params* params, const int32_t* qualities, int n_q,
              const uint8_t* rgb, int rgb_loc,
              uint8_t* recon, int out_loc, jds_metrics* metrics);

/*
 * jds_sweep whose results stay on the device, for sweeps sharded over several GPUs
 * (BASELINE.json config 4; the loop of gui/worker.py:55-74 with its points dealt round-robin
 * to the ranks).  After the kernels one record of JDS_RECORD_FIELDS doubles per quality is
 * written to `records` (DEVICE memory, `capacity` rows):
 *   unit (= unit0 + i*unit_step), quality, sse_rgb, sse_y, ssim_sum[R,G,B,Y], ssim_count,
 *   coeff_bits, nnz, total_coeffs, luma_blocks          (the fields of jds_metrics)
 * rows n_q .. capacity-1 are marked empty (unit = -1), so every rank contributes `capacity`
 * rows to one all-gather.  Unlike every other entry point this one returns WITHOUT
 * synchronising: the records are ordered on the context's stream (jds_ctx_set_stream), the
 * caller enqueues its collective / copy there and synchronises once.  capacity <=
 * JDS_SWEEP_RECORDS_MAX; params->outputs may hold JDS_OUT_SSIM | JDS_OUT_PSNR only.
 */
#define JDS_RECORD_FIELDS 13
#define JDS_SWEEP_RECORDS_MAX 128
int jds_sweep_records(jds_ctx* ctx, const jds_params* params, const int32_t* qualities, int n_q,
                      const uint8_t* rgb, int rgb_loc, int unit0, int unit_step,
                      double* records, int capacity);

/* jds_roundtrip_batch with device-resident records (rows as for jds_sweep_records, quality =
 * params->quality, no synchronisation): frames and the optional reconstruction are DEVICE
 * buffers.  For callers that stream batch after batch and read the metrics at the end. */
int jds_roundtrip_batch_records(jds_ctx* ctx, const jds_params* params, int n_frames,
                                const uint8_t* rgb, uint8_t* recon, int unit0, int unit_step,
                                double* records, int capacity);

/*
 * The six 8x8 arrays of IntermediateData.selected_block_* for luma block
 * (block_row, block_col) (engines/pipeline.py:126-151): original, shifted, dct,
 * dequantized, reconstructed as fp64[64] and quantized as int16[64], all host
 * pointers.  *present = 0 (arrays untouched) when row*blocks_per_row+col is outside
 * [0, n_blocks) - the reference then leaves the fields None.
 */
int jds_selected_block(jds_ctx* ctx, const jds_params* params,
                       const uint8_t* rgb, int rgb_loc, int block_row, int block_col,
                       double original[64], double shifted[64], double dct[64],
                       int16_t quantized[64], double dequantized[64],
                       double reconstructed[64], int* present);

/*
 * GUI plot payload.  One round trip that returns what the reference's plots
 * draw instead of the bulky intermediates they are drawn from:
 *   value_hist[JDS_VALUE_HIST_BINS]  count of every int16 coefficient value, value v at
 *       index v + 1024 (replaces all_quantized_coeffs for the "Quantized Coefficient
 *       Distribution" plot: gui/compression_tab.py:662-667 -> gui/widgets/mpl_canvas.py:81-100;
 *       the host bins it exactly as matplotlib's hist / np.histogram(bins=50) would)
 *   heat_y, heat_rgb  H*W uint8 = trunc(clip(error_map * 10, 0, 255)), the amplified error
 *       map of gui/widgets/mpl_canvas.py:116-118 (replaces the 8 B/px fp64 error maps)
 *   recon  H*W*3 uint8 reconstruction.
 * recon, heat_y, heat_rgb and value_hist may each be NULL.  params->outputs selects the
 * metrics (JDS_OUT_SSIM | JDS_OUT_PSNR | JDS_OUT_HIST); in JDS_EXACT precision all outputs are
 * bit-identical to the reference's arrays reduced the same way.
 */
#define JDS_VALUE_HIST_BINS 2048
int jds_plot_payload(jds_ctx* ctx, const jds_params* params, const uint8_t* rgb, int rgb_loc,
                     uint8_t* recon, uint8_t* heat_y, uint8_t* heat_rgb, int64_t* value_hist,
                     int out_loc, jds_metrics* metrics);

/*
 * Preview downscale: the step in front of the round trip when the GUI's
 * preview mode is on (gui/compression_tab.py:532-552).
 *   jds_preview_size: the reference's size rule - frames that fit the target are kept,
 *       otherwise scale = min(target_w/w, target_h/h), new = int(dim * scale)  (:541-547).
 *   jds_resize_area:  cv2.resize(uint8 RGB, (out_w, out_h), interpolation=cv2.INTER_AREA)
 *       (:549-552), bit-identical to OpenCV 4.13 for shrinking (out <= in on both axes);
 *       enlarging returns JDS_ERR_UNSUPPORTED.  rgb / out: host or device, like jds_roundtrip.
 */
int jds_preview_size(int height, int width, int target_w, int target_h, int* out_h, int* out_w);
int jds_resize_area(jds_ctx* ctx, const uint8_t* rgb, int rgb_loc, int height, int width,
                    uint8_t* out, int out_h, int out_w, int out_loc);

/*
 * Chroma-aliasing demo: one arm of AliasingDemoWorker.run
 * (gui/dialogs/aliasing_demo_dialog.py:98-166).
 *   1. _process_with_explicit_subsample(prefilter) (:125-150): OpenCV FLOAT32 arithmetic -
 *      cvtColor(RGB2YCrCb) of the float image, optional GaussianBlur(5x5, 0.8) of Cb / Cr,
 *      [::2, ::2] decimation, resize(INTER_LINEAR) back to H x W, cvtColor(YCrCb2RGB),
 *      clip + truncate to uint8 - bit-identical to OpenCV 4.13 -> `subsampled`
 *   2. the hot path at 4:4:4, prefilter off, quality `quality` (:152-158) on that frame
 *      -> `recon` (bit-identical to the reference in JDS_EXACT precision)
 *   3. compute_metrics(original, recon) (:69-83): m_rgb = squared error / SSIM sums of the RGB
 *      frames (psnr_rgb, ssim_rgb as for jds_metrics); m_luma = the same sums of OpenCV's
 *      integer luma Y = (4899 R + 9617 G + 1868 B + 8192) >> 14 replicated to three channels:
 *      psnr_y from sse_rgb / 3 over H*W samples, ssim_y = ssim_sum[0] / ssim_count
 *   4. _compute_difference (:162-166): clip(|original - recon| * 10, 0, 255) uint8 -> `diff`
 * subsampled, recon, diff: H*W*3 uint8 each, may be NULL; frames of at least 8 x 8.
 */
int jds_aliasing_demo(jds_ctx* ctx, const uint8_t* rgb, int rgb_loc, int height, int width,
                      int quality, int prefilter, int precision, uint8_t* subsampled,
                      uint8_t* recon, uint8_t* diff, int out_loc, jds_metrics* m_rgb,
                      jds_metrics* m_luma);
/* step 3 alone: compute_metrics(a, b) (gui/dialogs/aliasing_demo_dialog.py:69-83) for two
 * uint8 RGB frames (host or device): m_rgb / m_luma as for jds_aliasing_demo */
int jds_aliasing_metrics(jds_ctx* ctx, const uint8_t* a, const uint8_t* b, int loc, int height,
                         int width, jds_metrics* m_rgb, jds_metrics* m_luma);

/*
 * Entropy-coded size: the reference reports only "Estimated (no entropy coding)"
 * (utils/metrics.py:51-92) and leaves ZIGZAG_ORDER (utils/constants.py:18-27) unused.  This is
 * the exact number of bits a baseline JPEG (ITU-T T.81: zig-zag scan, DC differences, run/size
 * Huffman codes with the Annex K tables - luminance tables for Y, chrominance for Cb/Cr)
 * spends on `coeffs` (int16, the order of all_quantized_coeffs, jds_coeff_count() values,
 * host or device) when the three components are coded as non-interleaved scans:
 * scan_bits[0..2] = Y, Cb, Cr, before byte stuffing and padding.
 */
int jds_entropy_bits(jds_ctx* ctx, const int16_t* coeffs, int loc, int height, int width,
                     int subsampling, uint64_t scan_bits[3]);

/*
 * The entropy-coded BYTES of those three scans, coded on the device: zig-zag order
 * (utils/constants.py:18-27 ZIGZAG_ORDER, which the reference defines and never uses), DC
 * differences, run/size Huffman codes of the Annex K tables, the final byte of each scan padded
 * with 1-bits, 0x00 stuffed after every 0xFF (T.81 F.1.2, B.1.1.5).
 *   out [out] the three scans back to back (scan k starts at scan_bytes[0] + .. + scan_bytes[k-1]),
 *             host or device (out_loc); NULL = sizes only.  out_capacity < total returns
 *             JDS_ERR_CAPACITY with scan_bytes / scan_bits filled.
 * Values without a baseline code (DC difference beyond 11 bits, AC beyond 10 bits - the round
 * trip never produces them: |q| <= 1016, engines/quantizer.py:22-24) return JDS_ERR_INVALID.
 */
int jds_entropy_encode(jds_ctx* ctx, const int16_t* coeffs, int loc, int height, int width,
                       int subsampling, uint8_t* out, int out_loc, uint64_t out_capacity,
                       uint64_t scan_bytes[3], uint64_t scan_bits[3]);
/*
 * A complete baseline JFIF file (SOI, APP0, DQT, SOF0, 4 x DHT, three non-interleaved
 * SOS + scan, EOI) of a round trip's coefficients - the file the reference's "compression"
 * stands for but never writes (utils/metrics.py:57-61).  qtable: the 64 raster-order values of
 * jds_quant_table (one table for all components, engines/pipeline.py:43).  out: HOST buffer or
 * NULL (size only); *out_bytes = file size.  Subsampled modes need even sizes (JPEG rounds
 * component sizes up, engines/color_space.py:44-49 rounds down).
 */
int jds_jfif_encode(jds_ctx* ctx, const int16_t* coeffs, int loc, int height, int width,
                    int subsampling, const double qtable[64], uint8_t* out, uint64_t out_capacity,
                    uint64_t* out_bytes, uint64_t scan_bits[3]);

/*
 * Stand-alone 8x8 block operators, exact (reference) arithmetic, host buffers:
 *   op 0 dct2, 1 idct2 (engines/dct_engine.py:7-14), 2 encode_block (-128 then DCT, :17-20),
 *   3 decode_block (IDCT, +128, clip, :23-27): in/out = n_blocks*64 fp64;
 *   4 quantize (engines/quantizer.py:22-24): in fp64, qtable[64], out_q int16;
 *   5 dequantize (:27-29): in_q int16, qtable[64], out fp64.
 * Unused pointers are NULL.
 */
int jds_block_op(jds_ctx* ctx, int op, int64_t n_blocks, const double* in, const int16_t* in_q,
                 const double* qtable, double* out, int16_t* out_q);

/*
 * The reference's stage functions as stand-alone operators, exact arithmetic, so that every
 * row of the path can be called (and checked against the reference) on its own.
 *   jds_color_convert   direction 0: rgb_to_ycbcr (engines/color_space.py:8-14);
 *                       direction 1: ycbcr_to_rgb incl. np.clip(0,255) (:17-24).
 *                       in/out: n_pixels x 3 fp64, host.
 *   jds_subsample_plane subsample_chroma of ONE chroma plane (:27-53): optional
 *                       cv2.GaussianBlur(3x3, 0.75), then cv2.resize(INTER_AREA) to
 *                       (W//2, H) / (W//2, H//2); out = jds_plane_dims() fp64 samples, host.
 *   jds_upsample_plane  upsample_chroma(method='bilinear') of ONE plane (:56-66):
 *                       cv2.resize(INTER_LINEAR) to out_height x out_width (enlarging only).
 *   jds_compare_images  compute_psnr_ssim (utils/metrics.py:9-28) partial sums of two uint8
 *                       RGB frames (host or device): sse_rgb, sse_y, ssim_sum[4], ssim_count.
 *   jds_bitrate_partials estimate_bitrate_no_entropy (utils/metrics.py:63-83) counts of an
 *                       int16 coefficient array: non-zeros and 6 + ceil(log2(|v|+1)) + 1 bits
 *                       per non-zero (block overhead is added by the caller).
 */
int jds_color_convert(jds_ctx* ctx, int direction, int64_t n_pixels, const double* in, double* out);
int jds_subsample_plane(jds_ctx* ctx, const double* plane, int height, int width, int subsampling,
                        int prefilter, double* out);
int jds_upsample_plane(jds_ctx* ctx, const double* plane, int height, int width, double* out,
                       int out_height, int out_width);
int jds_compare_images(jds_ctx* ctx, const uint8_t* a, const uint8_t* b, int loc, int height,
                       int width, jds_metrics* metrics);
int jds_bitrate_partials(jds_ctx* ctx, const int16_t* coeffs, int loc, uint64_t n, uint64_t* nnz,
                         uint64_t* coeff_bits);

#ifdef __cplusplus
}
#endif
#endif /* JDS_H_ */
