// core.cuh -- per-element arithmetic of the FVDBM step, shared by every kernel.
//
// Everything here is `FVDBM_HD` (host+device) on purpose: tests/hostsim.cpp compiles the very same
// functions with g++ and drives them on the CPU against the oracle, so the arithmetic and the
// index decoding are checked in the GPU-less build container (the GPU tests then only have to
// catch staging / synchronisation mistakes).  The product library never runs them on the host.
//
// Reference formulas (paths relative to /root/reference):
//   moments      src/dynamics.py:35-47      rho = sum f ; u = KSI^T f / rho
//   equilibrium  src/dynamics.py:70-74      feq = W rho (1 + ku/C^2 + ku^2/(2C^4) - uu/(2C^2))
//                src/dynamics.py:101-102    D2Q13 adds  ku^3/(2C^6) - 3 ku uu/(2C^4)
//   upwind flux  src/containers.py:215-240  f* = (KSI.n >= 0) ? f_slot0 : f_slot1
//   LW flux      src/containers.py:242-278  f* = f0 + (f1-f0) (d0/(d0+d1) - varpi dt/(2(d0+d1)))
//   ghost        src/containers.py:280-287, utils/utils.py:153-154
//   cell update  src/containers.py:115-121  f += dt ( (1/tau)(feq - f) - sum_k s_k flux_k )
#pragma once
#include <cstdint>
#include <cstddef>

#if defined(__CUDACC__)
#define FVDBM_HD __host__ __device__ __forceinline__
#else
#define FVDBM_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define FVDBM_LDG(p) __ldg(p)
#else
#define FVDBM_LDG(p) (*(p))
#endif

namespace fvdbm {

constexpr int kTW = 32;
constexpr int32_t kHole = INT32_MIN;

template <typename real>
struct Params {
    real w[16];
    real inv_cs2, inv_2cs4, inv_2cs2, inv_2cs6, three_inv_2cs4;
    real inv_tau, dt;
};

// lattice velocities (src/dynamics.py:54-62, :81-93); constant-folded after unrolling
FVDBM_HD constexpr int kx(int q) {
    return (q == 1 || q == 5 || q == 8) ? 1 : (q == 3 || q == 6 || q == 7) ? -1 : (q == 9) ? 2 : (q == 11) ? -2 : 0;
}
FVDBM_HD constexpr int ky(int q) {
    return (q == 2 || q == 5 || q == 6) ? 1 : (q == 4 || q == 7 || q == 8) ? -1 : (q == 10) ? 2 : (q == 12) ? -2 : 0;
}

template <int Q>
FVDBM_HD size_t pdf_index(int64_t cell) {       // index of population 0; population q at + q*32
    return (size_t)(cell >> 5) * (size_t)(Q * kTW) + (size_t)(cell & 31);
}

template <typename real>
FVDBM_HD real ksi_dot(int q, real x, real y) {
    // sum of +-x, +-y, +-2x, +-2y in the same order as dot(KSI[q], (x,y))
    real a = real(kx(q)) * x, b = real(ky(q)) * y;
    return (kx(q) == 0) ? b : (ky(q) == 0) ? a : a + b;
}

template <typename real, int Q>
FVDBM_HD void moments(const real* f, real& rho, real& ux, real& uy) {
    real r = f[0];
#pragma unroll
    for (int q = 1; q < Q; ++q) r += f[q];
    real jx = real(0), jy = real(0);
#pragma unroll
    for (int q = 1; q < Q; ++q) {
        if (kx(q) != 0) jx += real(kx(q)) * f[q];
        if (ky(q) != 0) jy += real(ky(q)) * f[q];
    }
    rho = r;
    ux = jx / r;
    uy = jy / r;
}

template <typename real, int Q>
FVDBM_HD real feq(int q, real rho, real ux, real uy, real uu, const Params<real>& P) {
    const real ku = ksi_dot<real>(q, ux, uy);
    real poly = real(1) + ku * P.inv_cs2 + ku * ku * P.inv_2cs4 - uu * P.inv_2cs2;
    if (Q == 13) poly = poly + ku * ku * ku * P.inv_2cs6 - ku * uu * P.three_inv_2cs4;
    return P.w[q] * rho * poly;
}

// signed flux of one face accumulated into fl[] (fl[q] += s * Phi_q), evaluated in FACE orientation
// so that both cells of an interior face obtain bit-identical Phi (exact conservation):
//   slot = stencil slot of this cell, fn = populations of the other slot (neighbour or ghost)
//   mx,my = n*L ; alpha = d0/(d0+d1) ; gdt = dt/(2(d0+d1)L)
template <typename real, int Q, int SCHEME>
FVDBM_HD void side_flux(real* fl, const real* f, const real* fn, int slot, real sgn,
                        real mx, real my, real alpha, real gdt) {
#pragma unroll
    for (int q = 1; q < Q; ++q) {               // q = 0: KSI = 0 -> zero flux
        const real w = ksi_dot<real>(q, mx, my);
        const real f0 = slot ? fn[q] : f[q];
        const real f1 = slot ? f[q] : fn[q];
        real fs;
        if (SCHEME == 0) fs = (w >= real(0)) ? f0 : f1;
        else fs = f0 + (f1 - f0) * (alpha - w * gdt);
        fl[q] += sgn * (fs * w);
    }
}

// BGK relaxation + flux divergence (src/containers.py:121)
template <typename real, int Q>
FVDBM_HD void relax_update(real* out, const real* f, const real* fl, const Params<real>& P) {
    real rho, ux, uy;
    moments<real, Q>(f, rho, ux, uy);
    const real uu = ux * ux + uy * uy;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const real e = feq<real, Q>(q, rho, ux, uy, uu, P);
        out[q] = f[q] + P.dt * (P.inv_tau * (e - f[q]) - fl[q]);
    }
}

// boundary-side tables + tracked node populations (plan.hpp)
template <typename real>
struct GhostTables {
    const int32_t* bf_na;
    const int32_t* bf_nb;
    const real* bf_ratio;
    const real* npdf;      // SoA [Q][NTpad]
    int64_t NTpad;
    // border kernel only: node values staged per tile in shared memory, [slot][Q]; bf_la/bf_lb give
    // the slots of a boundary side's two nodes (null/unused elsewhere)
    const real* snode = nullptr;
    const int32_t* bf_la = nullptr;
    const int32_t* bf_lb = nullptr;
};

// One cell of the fused step: K sides (neighbour populations through `load_nbr(pos, fn)`, ghosts
// from the boundary tables), then relaxation + update.  `code` / `coef` hold this cell's K side
// codes and K*NC side coefficients (plan.hpp).
template <typename real, int Q, int K, int SCHEME, typename NbrLoader>
FVDBM_HD void advance_cell(const Params<real>& P, const GhostTables<real>& G, const real* f, const int32_t* code,
                           const real* coef, NbrLoader&& load_nbr, real* out) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    real fl[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) fl[q] = real(0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int32_t cd = code[k];
        const real mx = coef[k * NC + 0], my = coef[k * NC + 1];
        real alpha = real(0), gdt = real(0);
        if (SCHEME != 0) { alpha = coef[k * NC + 2]; gdt = P.dt * coef[k * NC + 3]; }
        real fn[Q];
        int32_t v;
        if (cd >= 0) {
            v = cd;
            load_nbr((int64_t)(cd >> 2), fn);
        } else {
            v = -(cd + 1);
            const int32_t b = v >> 2;
            const real ratio = FVDBM_LDG(G.bf_ratio + b);
            if (G.snode) {                      // node values staged by this tile (k_border)
                const real* ga = G.snode + (size_t)FVDBM_LDG(G.bf_la + b) * Q;
                const real* gb = G.snode + (size_t)FVDBM_LDG(G.bf_lb + b) * Q;
#pragma unroll
                for (int q = 1; q < Q; ++q) {
                    const real g = (ga[q] + gb[q]) / real(2);
                    fn[q] = g + (g - f[q]) * ratio;
                }
            } else {
                const int32_t na = FVDBM_LDG(G.bf_na + b), nb = FVDBM_LDG(G.bf_nb + b);
#pragma unroll
                for (int q = 1; q < Q; ++q) {   // containers.py:285-287: mean of the two node PDFs, extrapolated
                    const real g = (FVDBM_LDG(G.npdf + q * G.NTpad + na) + FVDBM_LDG(G.npdf + q * G.NTpad + nb)) / real(2);
                    fn[q] = g + (g - f[q]) * ratio;
                }
            }
        }
        fn[0] = real(0);
        side_flux<real, Q, SCHEME>(fl, f, fn, v & 1, (v & 2) ? real(-1) : real(1), mx, my, alpha, gdt);
    }
    relax_update<real, Q>(out, f, fl, P);
}

// One active boundary node (src/containers.py:339-404): ring sums are supplied by the caller
// (warp-reduced on the GPU); returns the node's rho, vel and populations.
template <typename real, int Q>
FVDBM_HD void node_finish(const Params<real>& P, int type, real sw, real srho, real sux, real suy, const real* sneq,
                          real& rho_n, real& ux_n, real& uy_n, real* pdf_n) {
    if (type == 1) rho_n = srho / sw;                                  // containers.py:348-351
    if (type == 2) { ux_n = sux / sw; uy_n = suy / sw; }               // containers.py:343-346
    const real uu = ux_n * ux_n + uy_n * uy_n;
#pragma unroll
    for (int q = 0; q < Q; ++q)                                         // containers.py:353-361
        pdf_n[q] = feq<real, Q>(q, rho_n, ux_n, uy_n, uu, P) + sneq[q] / sw;
}

// contribution of one ring cell to a node's sums
template <typename real, int Q>
FVDBM_HD void node_accumulate(const Params<real>& P, const real* f, real w, real& sw, real& srho, real& sux, real& suy,
                              real* sneq) {
    real rho, ux, uy;
    moments<real, Q>(f, rho, ux, uy);
    const real uu = ux * ux + uy * uy;
    sw += w; srho += rho * w; sux += ux * w; suy += uy * w;
#pragma unroll
    for (int q = 0; q < Q; ++q) sneq[q] += (f[q] - feq<real, Q>(q, rho, ux, uy, uu, P)) * w;
}

// decode helpers for side codes (plan.hpp)
FVDBM_HD int code_slot(int32_t v) { return v & 1; }
FVDBM_HD int code_neg(int32_t v) { return (v >> 1) & 1; }
FVDBM_HD int32_t code_index(int32_t v) { return v >> 2; }

}  // namespace fvdbm
