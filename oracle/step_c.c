/* step_c.c -- plain C (+OpenMP) restatement of the reference hot path Environment.step()
 * (/root/reference/src/environment.py:55-65), in the reference's own data model: AoS arrays,
 * original numbering, five separate sweeps S1..S5 with materialised pdf_eq and flux arrays.
 *
 * TEST INFRASTRUCTURE ONLY: it is the timed CPU baseline of bench.py ("cpu_baseline", kind
 * "port", and the --impl reference arm) and a second checker; nothing in fvdbm_jax_b200 links or
 * calls it.  Pinned against tests/golden/*.npz (minted from the reference's own sources under
 * oracle/jaxshim) by tests/test_oracle_golden.py.  JAX itself cannot run here (absent, no
 * network), so this port -- not jax.jit -- is what gets timed; unlike XLA it skips the work the
 * reference computes and then masks away (ghosts of interior faces, untyped nodes), so it is a
 * conservative (fast) stand-in.
 *
 * Built by __graft_entry__.build():  gcc -O3 -march=native -fopenmp -fPIC -shared
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct oracle_desc {
    int32_t scheme, Q, K, M;
    int64_t N, F, P;
    double tau, delta_t;
    double lat_w[16];
    double cs2, two_cs4, two_cs2, two_cs6;
    const int32_t* cell_face_idx;   /* [N*K] Cells.face_indices            src/containers.py:60  */
    const int32_t* cell_face_sign;  /* [N*K] Cells.face_normals            src/containers.py:62  */
    const int32_t* face_cell_idx;   /* [F*2] Faces.stencil_cells_index     src/containers.py:151 */
    const void* face_dists;         /* [F*2] Faces.stencil_dists           src/containers.py:152 */
    const int32_t* face_node_idx;   /* [F*2] Faces.nodes_index             src/containers.py:150 */
    const void* face_n;             /* [F*2] Faces.n                       src/containers.py:153 */
    const void* face_L;             /* [F]   Faces.L                       src/containers.py:154 */
    const int32_t* node_type;       /* [P]   Nodes.type                    src/containers.py:305 */
    const int32_t* node_cell_idx;   /* [P*M] Nodes.cells_index             src/containers.py:306 */
    const void* node_cell_dist;     /* [P*M] Nodes.cell_dists              src/containers.py:307 */
} oracle_desc;

/* src/dynamics.py:54-62 and :81-93 */
static const int KSI[13][2] = {{0, 0}, {1, 0}, {0, 1}, {-1, 0}, {0, -1}, {1, 1}, {-1, 1}, {-1, -1}, {1, -1},
                               {2, 0}, {0, 2}, {-2, 0}, {0, -2}};

#define REAL float
#define QCONST 9
#define SUFFIX(name) name##_f32_q9
#include "step_c_impl.h"
#undef QCONST
#undef SUFFIX
#define QCONST 13
#define SUFFIX(name) name##_f32_q13
#include "step_c_impl.h"
#undef QCONST
#undef SUFFIX
#undef REAL

#define REAL double
#define QCONST 9
#define SUFFIX(name) name##_f64_q9
#include "step_c_impl.h"
#undef QCONST
#undef SUFFIX
#define QCONST 13
#define SUFFIX(name) name##_f64_q13
#include "step_c_impl.h"
#undef QCONST
#undef SUFFIX
#undef REAL

int fvdbm_oracle_step_f32(const oracle_desc* d, float* pdf, float* rho, float* vel, float* pdf_eq, float* flux,
                          float* npdf, float* nrho, float* nvel, int nsteps) {
    if (d->Q == 9) return fvdbm_oracle_step_f32_q9(d, pdf, rho, vel, pdf_eq, flux, npdf, nrho, nvel, nsteps);
    if (d->Q == 13) return fvdbm_oracle_step_f32_q13(d, pdf, rho, vel, pdf_eq, flux, npdf, nrho, nvel, nsteps);
    return -1;
}
int fvdbm_oracle_step_f64(const oracle_desc* d, double* pdf, double* rho, double* vel, double* pdf_eq, double* flux,
                          double* npdf, double* nrho, double* nvel, int nsteps) {
    if (d->Q == 9) return fvdbm_oracle_step_f64_q9(d, pdf, rho, vel, pdf_eq, flux, npdf, nrho, nvel, nsteps);
    if (d->Q == 13) return fvdbm_oracle_step_f64_q13(d, pdf, rho, vel, pdf_eq, flux, npdf, nrho, nvel, nsteps);
    return -1;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host thread it can get */
void fvdbm_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int fvdbm_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
