// jds_host.h - host-side helpers shared by the C ABI (jds_api.cu) and the CPU
// emulation harness used by the tests (tests/emul): frame geometry and the
// quantiser tables in the forms the two arithmetic policies consume.
#pragma once
#include <math.h>
#include <string.h>
#include "jds_internal.cuh"

namespace jds {

// returns 0 ok, 1 bad size, 2 bad mode
inline int geom_init(int H, int W, int sub, Geom* g) {
    if (H < 1 || W < 1) return 1;
    if (sub < 0 || sub > 2) return 2;
    if (sub != 0 && W < 2) return 1;              // cv2.resize to width 0 raises
    if (sub == 2 && H < 2) return 1;
    memset(g, 0, sizeof *g);
    // an odd width (odd height under 4:2:0) sends cv2.resize(INTER_AREA) down its general
    // (fractional) path for both axes
    g->general = (sub != 0 && (W % 2)) || (sub == 2 && (H % 2));
    g->H = H;
    g->W = W;
    g->sub = sub;
    g->wc = sub == 0 ? W : W / 2;                 // engines/color_space.py:44,48
    g->hc = sub == 2 ? H / 2 : H;
    g->Hp = (H + 7) / 8 * 8;                      // engines/block_processor.py:10-11
    g->Wp = (W + 7) / 8 * 8;
    g->hcp = (g->hc + 7) / 8 * 8;
    g->wcp = (g->wc + 7) / 8 * 8;
    g->nbx_y = g->Wp / 8;
    g->nby_y = g->Hp / 8;
    g->nbx_c = g->wcp / 8;
    g->nby_c = g->hcp / 8;
    g->W4 = 4 * (W / 4);
    g->sx = (double)g->wc / (double)W;
    g->sy = (double)g->hc / (double)H;
    g->nblk_y = (long long)g->nbx_y * g->nby_y;
    g->nblk_c = (long long)g->nbx_c * g->nby_c;
    g->plane_y = (long long)g->Hp * g->Wp;
    g->plane_c = (long long)g->hcp * g->wcp;
    return 0;
}

inline void fill_tables(int quality, QTables* t) {
    quant_table_host(quality, t->q);
    for (int i = 0; i < 64; ++i) {
        volatile double r = 1.0 / t->q[i];        // correctly rounded reciprocal
        t->rq[i] = r;
        t->dqx[i] = ldexp(t->q[i], -exact_coeff_shift(i) - 4);   // Q * 2^-shift / 16, exact
    }
    struct Aan {                                        // AAN scale factors
        double s[8];
        Aan() {
            s[0] = 1.0;
            for (int k = 1; k < 8; ++k) s[k] = sqrt(2.0) * cos(k * M_PI / 16.0);
        }
    };
    static const Aan aan;                               // C++11: initialised once, thread safe
    const double* s = aan.s;
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            const double fwd = (2.0 * sqrt(2.0) * s[u]) * (2.0 * sqrt(2.0) * s[v]);
            const double inv = (s[u] / (2.0 * sqrt(2.0))) * (s[v] / (2.0 * sqrt(2.0)));
            t->fq[u * 8 + v] = (float)(1.0 / (t->q[u * 8 + v] * fwd));
            t->dq[u * 8 + v] = (float)(t->q[u * 8 + v] * inv);
        }
}

}  // namespace jds
