// jds_api.cu - the C ABI of libjds.so (include/jds.h): context, buffers, staging of
// host data, chunking of batches / sweeps, and the launch sequence of the kernels.
// No CPU compute path exists here: every entry point that produces pixels or metrics
// launches the kernels of jds_kernels.cu and fails when no CUDA device is usable.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/jds.h"
#include "jds_host.h"
#include "jds_kernels.cuh"

using namespace jds;

// ------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define JDS_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return fail(JDS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                    \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

// ------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct jds_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double stage_ms[4] = {0, 0, 0, 0};   // forward, codec, inverse, ssim (accumulated)
    uint64_t stage_launches[4] = {0, 0, 0, 0};
    DevBuf planes, in, recon, coeffs, errs, metrics, tables, selected;
    void* h_metrics = nullptr;   // pinned
    size_t h_metrics_bytes = 0;
    void* h_tables = nullptr;    // pinned
    size_t h_tables_bytes = 0;
    void* h_selected = nullptr;  // pinned
    uint64_t launches = 0;
    int sm_count = 148;
    bool legacy_ssim = false;     // JDS_LEGACY_SSIM=1: use the tile kernel (debug / A-B runs)
    bool no_fused = false;        // JDS_NO_FUSED=1: fast mode through the staged kernels
    size_t scratch_budget = (size_t)1 << 30;
};

static int ensure(jds_ctx* c, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return JDS_OK;
    if (b.p) {
        JDS_CUDA(cudaStreamSynchronize(c->stream));
        JDS_CUDA(cudaFree(b.p));
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess)
        return fail(JDS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    b.bytes = want;
    return JDS_OK;
}

static int ensure_pinned(void** p, size_t* have, size_t bytes) {
    if (*have >= bytes) return JDS_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    JDS_CUDA(cudaMallocHost(p, bytes));
    *have = bytes;
    return JDS_OK;
}

extern "C" int jds_abi_version(void) { return JDS_ABI_VERSION; }

extern "C" const char* jds_last_error(void) { return g_last_error.c_str(); }

extern "C" int jds_device_count(int* count) {
    if (!count) return fail(JDS_ERR_INVALID, "count is NULL");
    *count = 0;
    JDS_CUDA(cudaGetDeviceCount(count));
    return JDS_OK;
}

extern "C" int jds_ctx_create(int device, jds_ctx** out) {
    if (!out) return fail(JDS_ERR_INVALID, "ctx out pointer is NULL");
    *out = nullptr;
    int n = 0;
    JDS_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n)
        return fail(JDS_ERR_INVALID, "device %d out of range (%d CUDA devices)", device, n);
    JDS_CUDA(cudaSetDevice(device));
    jds_ctx* c = new jds_ctx();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->evs[i]);
    if (e != cudaSuccess) {
        delete c;
        return fail(JDS_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    }
    c->own_stream = true;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    const char* ls = getenv("JDS_LEGACY_SSIM");
    c->legacy_ssim = ls && atoi(ls) != 0;
    const char* nf = getenv("JDS_NO_FUSED");
    c->no_fused = nf && atoi(nf) != 0;
    const char* mb = getenv("JDS_SCRATCH_MB");
    if (mb && atol(mb) > 0) c->scratch_budget = (size_t)atol(mb) << 20;
    *out = c;
    return JDS_OK;
}

extern "C" int jds_ctx_destroy(jds_ctx* c) {
    if (!c) return JDS_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->planes, &c->in, &c->recon, &c->coeffs, &c->errs,
                      &c->metrics, &c->tables, &c->selected};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->h_metrics) cudaFreeHost(c->h_metrics);
    if (c->h_tables) cudaFreeHost(c->h_tables);
    if (c->h_selected) cudaFreeHost(c->h_selected);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 5; ++i)
        if (c->evs[i]) cudaEventDestroy(c->evs[i]);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return JDS_OK;
}

extern "C" int jds_ctx_set_stream(jds_ctx* c, void* cuda_stream) {
    if (!c) return fail(JDS_ERR_INVALID, "ctx is NULL");
    JDS_CUDA(cudaSetDevice(c->device));
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return JDS_OK;
}

extern "C" int jds_ctx_synchronize(jds_ctx* c) {
    if (!c) return fail(JDS_ERR_INVALID, "ctx is NULL");
    JDS_CUDA(cudaSetDevice(c->device));
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    return JDS_OK;
}

extern "C" int jds_ctx_stage_times(jds_ctx* c, double ms[4], uint64_t launches[4], int reset) {
    if (!c || !ms || !launches) return fail(JDS_ERR_INVALID, "NULL argument");
    for (int i = 0; i < 4; ++i) {
        ms[i] = c->stage_ms[i];
        launches[i] = c->stage_launches[i];
        if (reset) {
            c->stage_ms[i] = 0;
            c->stage_launches[i] = 0;
        }
    }
    return JDS_OK;
}

extern "C" int jds_ctx_launch_count(jds_ctx* c, uint64_t* launches) {
    if (!c || !launches) return fail(JDS_ERR_INVALID, "NULL argument");
    *launches = c->launches;
    return JDS_OK;
}

// ------------------------------------------------------------------------------
// host-only helpers
// ------------------------------------------------------------------------------
extern "C" int jds_quant_table(int quality, double table[64]) {
    if (!table) return fail(JDS_ERR_INVALID, "table is NULL");
    if (quality < 1 || quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", quality);
    quant_table_host(quality, table);
    return JDS_OK;
}

static int make_geom(int H, int W, int sub, Geom* g) {
    switch (geom_init(H, W, sub, g)) {
        case 0: return JDS_OK;
        case 1: return fail(JDS_ERR_INVALID, "bad frame size %dx%d", H, W);
        case 2: return fail(JDS_ERR_INVALID, "Unknown subsampling mode: %d", sub);
        case 3:
            return fail(JDS_ERR_UNSUPPORTED,
                        "odd width %d with chroma subsampling: OpenCV's fractional INTER_AREA / "
                        "non-2x INTER_LINEAR path is not implemented", W);
        default:
            return fail(JDS_ERR_UNSUPPORTED,
                        "odd height %d with 4:2:0: OpenCV's fractional INTER_AREA / non-2x "
                        "INTER_LINEAR path is not implemented", H);
    }
}

extern "C" int jds_coeff_count(int height, int width, int subsampling, uint64_t* count) {
    if (!count) return fail(JDS_ERR_INVALID, "count is NULL");
    Geom g;
    int rc = make_geom(height, width, subsampling, &g);
    if (rc) return rc;
    *count = 64ull * (uint64_t)(g.nblk_y + 2 * g.nblk_c);
    return JDS_OK;
}

extern "C" int jds_plane_dims(int height, int width, int subsampling, int* ch, int* cw) {
    if (!ch || !cw) return fail(JDS_ERR_INVALID, "NULL argument");
    Geom g;
    int rc = make_geom(height, width, subsampling, &g);
    if (rc) return rc;
    *ch = g.hc;
    *cw = g.wc;
    return JDS_OK;
}

static int check_params(const jds_params* p) {
    if (!p) return fail(JDS_ERR_INVALID, "params is NULL");
    if (p->precision != JDS_EXACT && p->precision != JDS_FAST)
        return fail(JDS_ERR_INVALID, "bad precision %d", p->precision);
    return JDS_OK;
}

// ------------------------------------------------------------------------------
// the launch sequence over `units` units (frames of a batch or sweep points)
// ------------------------------------------------------------------------------
struct UnitJob {
    const jds_params* p;
    Geom g;
    int units;
    const int32_t* qualities;   // per unit (sweep) or NULL (params->quality for all)
    bool shared_input;          // sweep: one frame, forward stage run once
    const uint8_t* rgb;
    int rgb_loc;
    uint8_t* recon;
    int16_t* coeffs;
    double* err_y;
    double* err_rgb;
    int out_loc;
    jds_metrics* metrics;
};

static int run_job(jds_ctx* c, const UnitJob& J) {
    const Geom& g = J.g;
    const jds_params* p = J.p;
    const bool exact = p->precision == JDS_EXACT;
    const size_t esz = exact ? sizeof(double) : sizeof(float);
    const size_t frame_bytes = (size_t)g.H * g.W * 3;
    const size_t planes_elems = (size_t)(g.plane_y + 2 * g.plane_c);
    const size_t ncoef = 64ull * (size_t)(g.nblk_y + 2 * g.nblk_c);
    const bool want_coeffs = (p->outputs & JDS_OUT_COEFFS) && J.coeffs;
    const bool want_hist = (p->outputs & JDS_OUT_HIST) != 0;
    const bool want_ssim = (p->outputs & JDS_OUT_SSIM) != 0;
    const bool want_recon = (p->outputs & JDS_OUT_RECON) && J.recon;
    const bool want_ey = (p->outputs & JDS_OUT_ERR_Y) && J.err_y;
    const bool want_ergb = (p->outputs & JDS_OUT_ERR_RGB) && J.err_rgb;
    if ((want_ey || want_ergb) && J.units != 1)
        return fail(JDS_ERR_INVALID, "error maps are single-frame outputs");

    JDS_CUDA(cudaSetDevice(c->device));

    // units per chunk from the scratch budget
    const size_t per_unit = planes_elems * esz * (J.shared_input ? 1 : 2) + frame_bytes +
                            (want_coeffs ? ncoef * 2 : 0);
    int chunk = (int)(c->scratch_budget / (per_unit ? per_unit : 1));
    if (chunk < 1) chunk = 1;
    if (chunk > J.units) chunk = J.units;
    if (chunk > 65535) chunk = 65535;

    int rc;
    const int fwd_units = J.shared_input ? 1 : chunk;
    if ((rc = ensure(c, c->planes, planes_elems * esz * (size_t)(fwd_units + chunk)))) return rc;
    if ((rc = ensure(c, c->metrics, sizeof(DevMetrics) * (size_t)chunk))) return rc;
    if ((rc = ensure(c, c->tables, sizeof(QTables) * (size_t)chunk))) return rc;
    if ((rc = ensure_pinned(&c->h_metrics, &c->h_metrics_bytes, sizeof(DevMetrics) * (size_t)chunk)))
        return rc;
    if ((rc = ensure_pinned(&c->h_tables, &c->h_tables_bytes, sizeof(QTables) * (size_t)chunk)))
        return rc;
    const bool in_host = J.rgb_loc == JDS_HOST;
    const bool out_host = J.out_loc == JDS_HOST;
    if (in_host && (rc = ensure(c, c->in, frame_bytes * (size_t)fwd_units))) return rc;
    if ((!want_recon || out_host) && (rc = ensure(c, c->recon, frame_bytes * (size_t)chunk)))
        return rc;
    if (want_coeffs && out_host && (rc = ensure(c, c->coeffs, ncoef * 2 * (size_t)chunk))) return rc;
    if ((want_ey || want_ergb) && out_host &&
        (rc = ensure(c, c->errs, (size_t)g.H * g.W * 8 * 2)))
        return rc;

    char* planes = (char*)c->planes.p;
    void* fwd = planes;
    void* rec = planes + planes_elems * esz * (size_t)fwd_units;
    const size_t fwd_stride = J.shared_input ? 0 : planes_elems;
    const size_t rec_stride = planes_elems;
    DevMetrics* d_metrics = (DevMetrics*)c->metrics.p;
    QTables* d_tables = (QTables*)c->tables.p;
    QTables* h_tables = (QTables*)c->h_tables;
    DevMetrics* h_metrics = (DevMetrics*)c->h_metrics;
    cudaStream_t s = c->stream;

    for (int u0 = 0; u0 < J.units; u0 += chunk) {
        const int n = (J.units - u0 < chunk) ? (J.units - u0) : chunk;
        // --- quantiser tables ---
        const int n_tables = J.qualities ? n : 1;
        for (int i = 0; i < n_tables; ++i)
            fill_tables(J.qualities ? J.qualities[u0 + i] : p->quality, &h_tables[i]);
        JDS_CUDA(cudaMemcpyAsync(d_tables, h_tables, sizeof(QTables) * n_tables,
                                 cudaMemcpyHostToDevice, s));
        // --- inputs ---
        const uint8_t* d_rgb;
        size_t rgb_stride = J.shared_input ? 0 : frame_bytes;
        if (in_host) {
            if (!J.shared_input || u0 == 0) {
                const size_t nin = J.shared_input ? 1 : (size_t)n;
                const uint8_t* src = J.rgb + (J.shared_input ? 0 : (size_t)u0 * frame_bytes);
                JDS_CUDA(cudaMemcpyAsync(c->in.p, src, frame_bytes * nin,
                                         cudaMemcpyHostToDevice, s));
            }
            d_rgb = (const uint8_t*)c->in.p;
        } else {
            d_rgb = J.rgb + (J.shared_input ? 0 : (size_t)u0 * frame_bytes);
        }
        // --- outputs ---
        uint8_t* d_recon = (want_recon && !out_host) ? J.recon + (size_t)u0 * frame_bytes
                                                     : (uint8_t*)c->recon.p;
        int16_t* d_coeffs = nullptr;
        if (want_coeffs)
            d_coeffs = out_host ? (int16_t*)c->coeffs.p : J.coeffs + (size_t)u0 * ncoef;
        double* d_ey = nullptr;
        double* d_ergb = nullptr;
        if (want_ey) d_ey = out_host ? (double*)c->errs.p : J.err_y;
        if (want_ergb) d_ergb = out_host ? (double*)c->errs.p + (size_t)g.H * g.W : J.err_rgb;

        JDS_CUDA(cudaMemsetAsync(d_metrics, 0, sizeof(DevMetrics) * n, s));
        JDS_CUDA(cudaEventRecord(c->ev0, s));
        JDS_CUDA(cudaEventRecord(c->evs[0], s));
        // fast mode on block-aligned frames runs the fused kernels (jds_fused.cu); exact
        // mode, prefiltered / ragged frames and the GUI-only outputs (histogram, error
        // maps) run the staged kernels (jds_kernels.cu)
        const bool fused = !exact && !c->no_fused && !want_hist && !want_ey && !want_ergb &&
                           fused_supported(g, p->prefilter, d_rgb, rgb_stride, d_recon, frame_bytes) &&
                           ssim_strip_supported(g.H, g.W, d_rgb, rgb_stride, d_recon, frame_bytes);
        bool do_fwd, do_inv, do_ssim;
        if (fused) {
            float* cpl = (float*)c->planes.p;
            const size_t cpl_stride = fused_chroma_plane_floats(g);
            do_fwd = g.sub != 0;
            if (do_fwd) {
                JDS_CUDA(launch_fused_chroma(g, d_rgb, rgb_stride, cpl, cpl_stride, d_tables,
                                             J.qualities ? 1 : 0, d_coeffs, ncoef, d_metrics, n, s));
                c->launches++;
            }
            JDS_CUDA(cudaEventRecord(c->evs[1], s));
            JDS_CUDA(launch_fused_luma(g, d_rgb, rgb_stride, cpl, cpl_stride, d_tables,
                                       J.qualities ? 1 : 0, d_coeffs, ncoef, d_recon, frame_bytes,
                                       d_metrics, n, s));
            c->launches++;
            JDS_CUDA(cudaEventRecord(c->evs[2], s));
            JDS_CUDA(cudaEventRecord(c->evs[3], s));
            do_inv = false;
            do_ssim = true;     // squared errors always come from the strip kernel here
            JDS_CUDA(launch_ssim_strip(g.H, g.W, d_rgb, rgb_stride, d_recon, frame_bytes, d_metrics,
                                       n, want_ssim, true, c->sm_count, s));
            c->launches++;
        } else {
            do_fwd = !J.shared_input || u0 == 0;
            do_inv = true;
            if (do_fwd) {
                launch_forward(exact, g, p->prefilter, d_rgb, rgb_stride, fwd, fwd_stride,
                               J.shared_input ? 1 : n, s);
                c->launches++;
            }
            JDS_CUDA(cudaEventRecord(c->evs[1], s));
            launch_codec(exact, g, fwd, fwd_stride, rec, rec_stride, d_tables,
                         J.qualities ? 1 : 0, d_coeffs, ncoef, want_hist, d_metrics, n, s);
            JDS_CUDA(cudaEventRecord(c->evs[2], s));
            launch_inverse(exact, g, d_rgb, rgb_stride, fwd, fwd_stride, rec, rec_stride, d_recon,
                           frame_bytes, d_ey, d_ergb, d_metrics, n, s);
            JDS_CUDA(cudaEventRecord(c->evs[3], s));
            c->launches += 2;
            do_ssim = want_ssim && g.H >= 7 && g.W >= 7;
            if (do_ssim) {
                if (!c->legacy_ssim &&
                    ssim_strip_supported(g.H, g.W, d_rgb, rgb_stride, d_recon, frame_bytes)) {
                    JDS_CUDA(launch_ssim_strip(g.H, g.W, d_rgb, rgb_stride, d_recon, frame_bytes,
                                               d_metrics, n, true, false, c->sm_count, s));
                } else {
                    launch_ssim(exact, g.H, g.W, d_rgb, rgb_stride, d_recon, frame_bytes,
                                d_metrics, n, s);
                }
                c->launches++;
            }
        }
        JDS_CUDA(cudaEventRecord(c->evs[4], s));
        JDS_CUDA(cudaEventRecord(c->ev1, s));
        JDS_CUDA(cudaGetLastError());
        // --- results back ---
        JDS_CUDA(cudaMemcpyAsync(h_metrics, d_metrics, sizeof(DevMetrics) * n,
                                 cudaMemcpyDeviceToHost, s));
        if (out_host) {
            if (want_recon)
                JDS_CUDA(cudaMemcpyAsync(J.recon + (size_t)u0 * frame_bytes, d_recon,
                                         frame_bytes * n, cudaMemcpyDeviceToHost, s));
            if (want_coeffs)
                JDS_CUDA(cudaMemcpyAsync(J.coeffs + (size_t)u0 * ncoef, d_coeffs,
                                         ncoef * 2 * (size_t)n, cudaMemcpyDeviceToHost, s));
            if (want_ey)
                JDS_CUDA(cudaMemcpyAsync(J.err_y, d_ey, (size_t)g.H * g.W * 8,
                                         cudaMemcpyDeviceToHost, s));
            if (want_ergb)
                JDS_CUDA(cudaMemcpyAsync(J.err_rgb, d_ergb, (size_t)g.H * g.W * 8,
                                         cudaMemcpyDeviceToHost, s));
        }
        JDS_CUDA(cudaStreamSynchronize(s));
        float ms = 0.f;
        JDS_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        {
            const bool ran[4] = {do_fwd, true, do_inv, do_ssim};
            for (int k = 0; k < 4; ++k) {
                float t = 0.f;
                JDS_CUDA(cudaEventElapsedTime(&t, c->evs[k], c->evs[k + 1]));
                if (ran[k]) {
                    c->stage_ms[k] += t;
                    c->stage_launches[k] += 1;
                }
            }
        }
        for (int i = 0; i < n; ++i) {
            jds_metrics* m = &J.metrics[u0 + i];
            const DevMetrics& d = h_metrics[i];
            memset(m, 0, sizeof *m);
            m->sse_rgb = d.sse_rgb;
            m->sse_y = d.sse_y;
            for (int k = 0; k < 4; ++k) m->ssim_sum[k] = d.ssim_sum[k];
            m->ssim_count = (want_ssim && g.H >= 7 && g.W >= 7)
                                ? (uint64_t)(g.H - 6) * (uint64_t)(g.W - 6) : 0;
            m->coeff_bits = d.coeff_bits;
            m->nnz = d.nnz;
            m->total_coeffs = ncoef;
            m->luma_blocks = (uint64_t)g.nblk_y;
            for (int k = 0; k < 50; ++k) m->hist50[k] = (int64_t)d.hist[k];
            m->gpu_ms = (double)ms / n;
        }
    }
    return JDS_OK;
}

// ------------------------------------------------------------------------------
// entry points
// ------------------------------------------------------------------------------
extern "C" int jds_roundtrip(jds_ctx* c, const jds_params* p, const uint8_t* rgb, int rgb_loc,
                             uint8_t* recon, int16_t* coeffs, double* err_y, double* err_rgb,
                             int out_loc, jds_metrics* metrics) {
    if (!c || !rgb || !metrics) return fail(JDS_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    J.p = p;
    J.units = 1;
    J.rgb = rgb;
    J.rgb_loc = rgb_loc;
    J.recon = recon;
    J.coeffs = coeffs;
    J.err_y = err_y;
    J.err_rgb = err_rgb;
    J.out_loc = out_loc;
    J.metrics = metrics;
    return run_job(c, J);
}

extern "C" int jds_roundtrip_batch(jds_ctx* c, const jds_params* p, int n_frames,
                                   const uint8_t* rgb, int rgb_loc, uint8_t* recon,
                                   int16_t* coeffs, int out_loc, jds_metrics* metrics) {
    if (!c || !rgb || !metrics) return fail(JDS_ERR_INVALID, "NULL argument");
    if (n_frames < 1) return fail(JDS_ERR_INVALID, "n_frames must be >= 1, got %d", n_frames);
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    if (p->outputs & (JDS_OUT_ERR_Y | JDS_OUT_ERR_RGB))
        return fail(JDS_ERR_INVALID, "error maps are not produced in batch mode");
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    J.p = p;
    J.units = n_frames;
    J.rgb = rgb;
    J.rgb_loc = rgb_loc;
    J.recon = recon;
    J.coeffs = coeffs;
    J.out_loc = out_loc;
    J.metrics = metrics;
    return run_job(c, J);
}

extern "C" int jds_sweep(jds_ctx* c, const jds_params* p, const int32_t* qualities, int n_q,
                         const uint8_t* rgb, int rgb_loc, uint8_t* recon, int out_loc,
                         jds_metrics* metrics) {
    if (!c || !rgb || !metrics || !qualities) return fail(JDS_ERR_INVALID, "NULL argument");
    if (n_q < 1) return fail(JDS_ERR_INVALID, "n_q must be >= 1, got %d", n_q);
    int rc = check_params(p);
    if (rc) return rc;
    for (int i = 0; i < n_q; ++i)
        if (qualities[i] < 1 || qualities[i] > 100)
            return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", qualities[i]);
    if (p->outputs & (JDS_OUT_ERR_Y | JDS_OUT_ERR_RGB | JDS_OUT_COEFFS))
        return fail(JDS_ERR_INVALID, "sweeps produce metrics and (optionally) recon only");
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    J.p = p;
    J.units = n_q;
    J.qualities = qualities;
    J.shared_input = true;
    J.rgb = rgb;
    J.rgb_loc = rgb_loc;
    J.recon = recon;
    J.out_loc = out_loc;
    J.metrics = metrics;
    return run_job(c, J);
}

extern "C" int jds_selected_block(jds_ctx* c, const jds_params* p, const uint8_t* rgb,
                                  int rgb_loc, int block_row, int block_col,
                                  double original[64], double shifted[64], double dct[64],
                                  int16_t quantized[64], double dequantized[64],
                                  double reconstructed[64], int* present) {
    if (!c || !rgb || !present || !original || !shifted || !dct || !quantized || !dequantized ||
        !reconstructed)
        return fail(JDS_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    Geom g;
    if ((rc = make_geom(p->height, p->width, JDS_SUB_444, &g))) return rc;
    // engines/pipeline.py:134-137: target = row * blocks_per_row + col, valid iff in range
    const long long target = (long long)block_row * g.nbx_y + block_col;
    if (target < 0 || target >= g.nblk_y) {
        *present = 0;
        return JDS_OK;
    }
    const int by = (int)(target / g.nbx_y), bx = (int)(target % g.nbx_y);
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t frame_bytes = (size_t)g.H * g.W * 3;
    const size_t ob = selected_out_bytes();
    if ((rc = ensure(c, c->selected, ob))) return rc;
    if ((rc = ensure(c, c->tables, sizeof(QTables)))) return rc;
    if ((rc = ensure_pinned(&c->h_tables, &c->h_tables_bytes, sizeof(QTables)))) return rc;
    if (!c->h_selected) JDS_CUDA(cudaMallocHost(&c->h_selected, ob));
    const uint8_t* d_rgb = rgb;
    if (rgb_loc == JDS_HOST) {
        // only the 8 (reflect-mapped) rows of the block are needed, but the kernel
        // indexes the frame: stage the whole frame (a few MB, once per GUI click)
        if ((rc = ensure(c, c->in, frame_bytes))) return rc;
        JDS_CUDA(cudaMemcpyAsync(c->in.p, rgb, frame_bytes, cudaMemcpyHostToDevice, c->stream));
        d_rgb = (const uint8_t*)c->in.p;
    }
    fill_tables(p->quality, (QTables*)c->h_tables);
    JDS_CUDA(cudaMemcpyAsync(c->tables.p, c->h_tables, sizeof(QTables), cudaMemcpyHostToDevice,
                             c->stream));
    launch_selected_block(g, d_rgb, bx, by, (const QTables*)c->tables.p, c->selected.p, c->stream);
    c->launches++;
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaMemcpyAsync(c->h_selected, c->selected.p, ob, cudaMemcpyDeviceToHost, c->stream));
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    const double* h = (const double*)c->h_selected;
    memcpy(original, h, 64 * 8);
    memcpy(shifted, h + 64, 64 * 8);
    memcpy(dct, h + 128, 64 * 8);
    memcpy(dequantized, h + 192, 64 * 8);
    memcpy(reconstructed, h + 256, 64 * 8);
    memcpy(quantized, h + 320, 64 * 2);
    *present = 1;
    return JDS_OK;
}
