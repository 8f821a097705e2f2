/* step_c_impl.h -- body of the C restatement, included once per precision by step_c.c.
 * REAL, QCONST (9 or 13), SUFFIX(name) are defined by the includer (compile-time Q lets gcc
 * unroll/vectorise the population loops). */

static void SUFFIX(eq)(const oracle_desc* d, REAL rho, REAL ux, REAL uy, const REAL* w, REAL c2, REAL tc4, REAL tc2,
                       REAL tc6, REAL* out) {
    /* src/dynamics.py:70-74 (D2Q9) and :101-102 (D2Q13) */
    const REAL uu = ux * ux + uy * uy;
    for (int q = 0; q < QCONST; ++q) {
        const REAL ku = (REAL)KSI[q][0] * ux + (REAL)KSI[q][1] * uy;
        REAL poly = (REAL)1 + ku / c2 + ku * ku / tc4 - uu / tc2;
        if (QCONST == 13) poly = poly + ku * ku * ku / tc6 - (REAL)3 * ku * uu / tc4;
        out[q] = w[q] * rho * poly;
    }
}

int SUFFIX(fvdbm_oracle_step)(const oracle_desc* d, REAL* pdf, REAL* rho, REAL* vel, REAL* pdf_eq, REAL* flux,
                              REAL* npdf, REAL* nrho, REAL* nvel, int nsteps) {
    const int Q = QCONST, K = d->K, M = d->M;
    const int64_t N = d->N, F = d->F, P = d->P;
    const REAL* fdist = (const REAL*)d->face_dists;
    const REAL* fn = (const REAL*)d->face_n;
    const REAL* fL = (const REAL*)d->face_L;
    const REAL* ncd = (const REAL*)d->node_cell_dist;
    REAL w[16];
    for (int q = 0; q < Q; ++q) w[q] = (REAL)d->lat_w[q];
    const REAL c2 = (REAL)d->cs2, tc4 = (REAL)d->two_cs4, tc2 = (REAL)d->two_cs2, tc6 = (REAL)d->two_cs6;
    const REAL dt = (REAL)d->delta_t, inv_tau = (REAL)(1.0 / d->tau);
    REAL* pdf_new = (REAL*)malloc((size_t)N * Q * sizeof(REAL));
    if (!pdf_new) return -2;
    for (int s = 0; s < nsteps; ++s) {
        /* S1 + S2: src/containers.py:93-105, src/dynamics.py:35-47 */
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < N; ++c) {
            const REAL* f = pdf + c * Q;
            REAL r = 0, jx = 0, jy = 0;
            for (int q = 0; q < Q; ++q) { r += f[q]; jx += (REAL)KSI[q][0] * f[q]; jy += (REAL)KSI[q][1] * f[q]; }
            rho[c] = r; vel[2 * c] = jx / r; vel[2 * c + 1] = jy / r;
            SUFFIX(eq)(d, r, vel[2 * c], vel[2 * c + 1], w, c2, tc4, tc2, tc6, pdf_eq + c * Q);
        }
        /* S3: src/containers.py:339-404, utils/utils.py:34-60 (only nodes with a BC type change) */
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t p = 0; p < P; ++p) {
            const int type = d->node_type[p];
            if (type == 0) continue;
            REAL sw = 0, srho = 0, sux = 0, suy = 0, sneq[16];
            for (int q = 0; q < Q; ++q) sneq[q] = 0;
            for (int m = 0; m < M; ++m) {
                REAL wt = (REAL)1 / ncd[p * M + m];
                if (wt < 0) wt = 0;                                  /* utils/utils.py:59 */
                int64_t c = d->node_cell_idx[p * M + m];
                const int pad = c == -1;
                if (c < 0) c += N;                                   /* python negative index */
                sw += wt; srho += rho[c] * wt; sux += vel[2 * c] * wt; suy += vel[2 * c + 1] * wt;
                if (!pad)                                            /* src/containers.py:388-390 */
                    for (int q = 0; q < Q; ++q) sneq[q] += (pdf[c * Q + q] - pdf_eq[c * Q + q]) * wt;
            }
            if (type == 1) nrho[p] = srho / sw;
            if (type == 2) { nvel[2 * p] = sux / sw; nvel[2 * p + 1] = suy / sw; }
            REAL e[16];
            SUFFIX(eq)(d, nrho[p], nvel[2 * p], nvel[2 * p + 1], w, c2, tc4, tc2, tc6, e);
            for (int q = 0; q < Q; ++q) npdf[p * Q + q] = e[q] + sneq[q] / sw;
        }
        /* S4: src/containers.py:191-287, utils/utils.py:153-154 */
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < F; ++j) {
            int64_t s0 = d->face_cell_idx[2 * j], s1 = d->face_cell_idx[2 * j + 1];
            const REAL d0 = fdist[2 * j], d1 = fdist[2 * j + 1];
            const int g0 = s0 == -1, g1 = s1 == -1;
            if (s0 < 0) s0 += N;
            if (s1 < 0) s1 += N;
            REAL f0[16], f1[16];
            for (int q = 0; q < Q; ++q) { f0[q] = pdf[s0 * Q + q]; f1[q] = pdf[s1 * Q + q]; }
            if (g0 || g1) {
                const int64_t na = d->face_node_idx[2 * j], nb = d->face_node_idx[2 * j + 1];
                for (int q = 0; q < Q; ++q) {
                    const REAL g = (npdf[na * Q + q] + npdf[nb * Q + q]) / (REAL)2;
                    const REAL k0 = pdf[s0 * Q + q], k1 = pdf[s1 * Q + q];
                    if (g0) f0[q] = g + (g - k1) * (d0 / d1);
                    if (g1) f1[q] = g + (g - k0) * (d1 / d0);
                }
            }
            for (int q = 0; q < Q; ++q) {
                const REAL varpi = (REAL)KSI[q][0] * fn[2 * j] + (REAL)KSI[q][1] * fn[2 * j + 1];
                REAL fs;
                if (d->scheme == 0) fs = varpi >= 0 ? f0[q] : f1[q];
                else {
                    const REAL dd = d0 + d1;
                    fs = f0[q] + (f1[q] - f0[q]) * (d0 / dd - (varpi * dt) / ((REAL)2 * dd));
                }
                flux[j * Q + q] = fs * varpi * fL[j];
            }
        }
        /* S5: src/containers.py:107-121 */
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < N; ++c) {
            REAL tot[16];
            for (int q = 0; q < Q; ++q) tot[q] = 0;
            for (int k = 0; k < K; ++k) {
                int64_t j = d->cell_face_idx[c * K + k];
                if (j < 0) j += F;
                const REAL sg = (REAL)d->cell_face_sign[c * K + k];
                for (int q = 0; q < Q; ++q) tot[q] += flux[j * Q + q] * sg;
            }
            for (int q = 0; q < Q; ++q) {
                const REAL f = pdf[c * Q + q];
                pdf_new[c * Q + q] = f + dt * (inv_tau * (pdf_eq[c * Q + q] - f) - tot[q]);
            }
        }
        memcpy(pdf, pdf_new, (size_t)N * Q * sizeof(REAL));
    }
    free(pdf_new);
    return 0;
}
