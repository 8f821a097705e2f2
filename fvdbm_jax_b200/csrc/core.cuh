// core.cuh -- per-element arithmetic of the FVDBM step, shared by every kernel.
//
// ONE canonical operation sequence.  Every function here is written over a value type V that is
// either a scalar (float / double: one cell) or, on the device, a packed float2 (two cells, Blackwell
// FFMA2 / FADD2 / FMUL2) and spells out every add / mul / fma explicitly (v_add, v_mul, v_fma ...),
// so no compiler is free to contract or reassociate: the thread-per-cell kernels, the two-cells-per-
// thread packed kernel, the TMA kernel and the g++ build in tests/hostsim all produce the SAME bits.
// (ptxas fuses `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 even under --fmad=false, so the sequence is
// arranged such that a product only ever feeds an fma multiplicand / addend or another product.)
//
// Everything is `FVDBM_HD` (host+device): tests/hostsim.cpp drives the very same functions on the CPU
// against the oracle.  The product library never runs them on the host.
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
#include <cmath>

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
constexpr int kNodeLanes = 8;      // lanes cooperating on one boundary node (ring slots j, j+8, ... per lane; xor butterfly over 8)
constexpr int32_t kHole = INT32_MIN;

// ---- value types ----------------------------------------------------------------------------------
template <typename V> struct VT;
template <> struct VT<float>  { using S = float;  using M = bool; static constexpr int L = 1; };
template <> struct VT<double> { using S = double; using M = bool; static constexpr int L = 1; };

// host build: plain operators are safe because hostsim is compiled with -ffp-contract=off
FVDBM_HD float v_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
FVDBM_HD float v_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
FVDBM_HD float v_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
FVDBM_HD double v_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
FVDBM_HD double v_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
FVDBM_HD double v_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
FVDBM_HD float v_neg(float a) { return -a; }
FVDBM_HD double v_neg(double a) { return -a; }
FVDBM_HD float v_div(float a, float b) { return a / b; }       // IEEE (nvcc default -prec-div=true)
FVDBM_HD double v_div(double a, double b) { return a / b; }
FVDBM_HD float v_sel(bool m, float a, float b) { return m ? a : b; }
FVDBM_HD double v_sel(bool m, double a, double b) { return m ? a : b; }
FVDBM_HD bool v_ge0(float a) { return a >= 0.0f; }
FVDBM_HD bool v_ge0(double a) { return a >= 0.0; }
FVDBM_HD bool v_le0(float a) { return a <= 0.0f; }
FVDBM_HD bool v_le0(double a) { return a <= 0.0; }
FVDBM_HD bool m_eq(bool a, bool b) { return a == b; }
FVDBM_HD bool m_sel(bool c, bool a, bool b) { return c ? a : b; }
template <typename V> FVDBM_HD V v_bcast(typename VT<V>::S s);
template <> FVDBM_HD float v_bcast<float>(float s) { return s; }
template <> FVDBM_HD double v_bcast<double>(double s) { return s; }
FVDBM_HD float v_lane(float a, int) { return a; }
FVDBM_HD double v_lane(double a, int) { return a; }

#if defined(__CUDACC__)
// packed pair: lane x = first cell, lane y = second cell of the thread
struct bool2 { bool x, y; };
template <> struct VT<float2> { using S = float; using M = bool2; static constexpr int L = 2; };
__device__ __forceinline__ float2 v_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 v_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 v_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 v_neg(float2 a) { return make_float2(-a.x, -a.y); }     // folds into operand modifiers
__device__ __forceinline__ float2 v_div(float2 a, float2 b) { return make_float2(a.x / b.x, a.y / b.y); }
__device__ __forceinline__ float2 v_sel(bool2 m, float2 a, float2 b) { return make_float2(m.x ? a.x : b.x, m.y ? a.y : b.y); }
__device__ __forceinline__ bool2 v_ge0(float2 a) { return bool2{a.x >= 0.0f, a.y >= 0.0f}; }
__device__ __forceinline__ bool2 v_le0(float2 a) { return bool2{a.x <= 0.0f, a.y <= 0.0f}; }
__device__ __forceinline__ bool2 m_eq(bool2 a, bool2 b) { return bool2{a.x == b.x, a.y == b.y}; }
__device__ __forceinline__ bool2 m_sel(bool2 c, bool2 a, bool2 b) { return bool2{c.x ? a.x : b.x, c.y ? a.y : b.y}; }
template <> __device__ __forceinline__ float2 v_bcast<float2>(float s) { return make_float2(s, s); }
__device__ __forceinline__ float v_lane(float2 a, int i) { return i ? a.y : a.x; }
#endif

template <typename V> FVDBM_HD V v_sub(V a, V b) { return v_add(a, v_neg(b)); }

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
// weight class of population q: W is constant on {0}, {1..4}, {5..8}, {9..12} (checked at create time)
FVDBM_HD constexpr int wclass(int q) { return q == 0 ? 0 : q <= 4 ? 1 : q <= 8 ? 5 : 9; }

template <int Q>
FVDBM_HD size_t pdf_index(int64_t cell) {       // index of population 0; population q at + q*32
    return (size_t)(cell >> 5) * (size_t)(Q * kTW) + (size_t)(cell & 31);
}

// Two population layouts (chosen per handle, api.cu):
//   lay 0  tiled AoSoA (above): every population streams coalesced; a neighbour's Q-1 moving populations lie in Q-1
//          different 32-byte sectors.
//   lay 1  records: the Q-1 MOVING populations of a cell are contiguous ((Q-1)*sizeof(real) bytes: exactly one sector
//          for D2Q9 fp32), the rest population q = 0 -- never gathered, KSI_0 = 0 -- lives in a separate array behind
//          the records.  A neighbour gather is two 128-bit loads from one sector.
template <int Q>
FVDBM_HD size_t pdf_off(int lay, int64_t Npad, int64_t cell, int q) {
    return lay == 0 ? pdf_index<Q>(cell) + (size_t)q * kTW
                    : (q == 0 ? (size_t)Npad * (Q - 1) + (size_t)cell : (size_t)cell * (Q - 1) + (size_t)(q - 1));
}

// KSI_q . (x, y) given s = x + y and d = x - y (each rounded once); negations and doubling are exact
template <typename V>
FVDBM_HD V ksi_dot(int q, V x, V y, V s, V d) {
    const int a = kx(q), b = ky(q);
    if (a == 0 && b == 0) return v_bcast<V>(typename VT<V>::S(0));
    if (b == 0) return a == 1 ? x : a == -1 ? v_neg(x) : a == 2 ? v_add(x, x) : v_neg(v_add(x, x));
    if (a == 0) return b == 1 ? y : b == -1 ? v_neg(y) : b == 2 ? v_add(y, y) : v_neg(v_add(y, y));
    if (a == b) return a == 1 ? s : v_neg(s);
    return a == 1 ? d : v_neg(d);
}

template <typename V, int Q>
FVDBM_HD void moments(const V* f, V& rho, V& ux, V& uy) {
    using S = typename VT<V>::S;
    V r = f[0];
#pragma unroll
    for (int q = 1; q < Q; ++q) r = v_add(r, f[q]);
    V jx = f[1], jy = f[2];                          // kx(1) = ky(2) = +1 open the two sums
#pragma unroll
    for (int q = 2; q < Q; ++q) {
        if (kx(q) == 1) jx = v_add(jx, f[q]);
        else if (kx(q) == -1) jx = v_sub(jx, f[q]);
        else if (kx(q) != 0) jx = v_fma(v_bcast<V>(S(kx(q))), f[q], jx);
    }
#pragma unroll
    for (int q = 3; q < Q; ++q) {
        if (ky(q) == 1) jy = v_add(jy, f[q]);
        else if (ky(q) == -1) jy = v_sub(jy, f[q]);
        else if (ky(q) != 0) jy = v_fma(v_bcast<V>(S(ky(q))), f[q], jy);
    }
    rho = r;
    ux = v_div(jx, r);
    uy = v_div(jy, r);
}

// equilibrium split as  feq_q = (W_q rho) * poly_q  so that consumers can fuse the last product
template <typename V, int Q>
struct Equilibrium {
    using S = typename VT<V>::S;
    V rho, ux, uy, us, ud, base, ut;
    FVDBM_HD Equilibrium(V rho_, V ux_, V uy_, const Params<S>& P) : rho(rho_), ux(ux_), uy(uy_) {
        const V uu = v_fma(ux, ux, v_mul(uy, uy));
        us = v_add(ux, uy); ud = v_sub(ux, uy);
        base = v_fma(v_neg(uu), v_bcast<V>(P.inv_2cs2), v_bcast<V>(S(1)));          // 1 - uu/(2C^2)
        ut = (Q == 13) ? v_mul(uu, v_bcast<V>(P.three_inv_2cs4)) : uu;
    }
    FVDBM_HD V poly(int q, const Params<S>& P) const {
        const V ku = ksi_dot<V>(q, ux, uy, us, ud);
        V p = v_fma(ku, v_fma(ku, v_bcast<V>(P.inv_2cs4), v_bcast<V>(P.inv_cs2)), base);
        if (Q == 13) p = v_fma(ku, v_fma(v_mul(ku, ku), v_bcast<V>(P.inv_2cs6), v_neg(ut)), p);
        return p;
    }
    FVDBM_HD V wrho(int q, const Params<S>& P) const { return v_mul(v_bcast<V>(P.w[wclass(q)]), rho); }
    FVDBM_HD V value(int q, const Params<S>& P) const { return v_mul(wrho(q, P), poly(q, P)); }
};

// Signed flux of one side accumulated into fl[] in CELL orientation.  The planner folds the two
// orientation signs of a side (sigma = Cells.face_normals entry, varsigma = +1 in stencil slot 0 / -1
// in slot 1; all folds are exact sign flips) into its coefficients:
//     Mx,My = sigma n L        A = varsigma d0/(d0+d1)        Gd = dt * sigma varsigma /(2(d0+d1)L)
// With W_q = KSI_q.(Mx,My) = sigma varpi_q L:
//     c_q  = A - W_q Gd            = varsigma (alpha - varpi_q dt/(2(d0+d1)))
//     f*   = f_slot0 + (fn - f) c_q      (fn - f = varsigma (f_slot1 - f_slot0))
//     fl_q += f* W_q               (one rounding: the product is never materialised)
// Both cells of a face see the same f* bits and exactly opposite products -> conservation as exact as
// the reference's shared flux array.  Upwind picks f_slot0 / f_slot1 by the sign of varpi_q.
template <typename V, int Q, int SCHEME>
FVDBM_HD void side_flux(V* fl, const V* f, const V* fn, typename VT<V>::M slot1, typename VT<V>::M neg,
                        V Mx, V My, V A, V Gd) {
    const V Ms = v_add(Mx, My), Md = v_sub(Mx, My);
    if (SCHEME == 0) {
        // varpi_q >= 0  <=>  sigma W_q >= 0 ; the own cell is upstream iff (slot == 0) == (varpi_q >= 0)
        const typename VT<V>::M gx = m_sel(neg, v_le0(Mx), v_ge0(Mx)), lx = m_sel(neg, v_ge0(Mx), v_le0(Mx));
        const typename VT<V>::M gy = m_sel(neg, v_le0(My), v_ge0(My)), ly = m_sel(neg, v_ge0(My), v_le0(My));
        const typename VT<V>::M gs = m_sel(neg, v_le0(Ms), v_ge0(Ms)), ls = m_sel(neg, v_ge0(Ms), v_le0(Ms));
        const typename VT<V>::M gd = m_sel(neg, v_le0(Md), v_ge0(Md)), ld = m_sel(neg, v_ge0(Md), v_le0(Md));
#pragma unroll
        for (int q = 1; q < Q; ++q) {
            const int a = kx(q), b = ky(q);
            const typename VT<V>::M ge = b == 0 ? (a > 0 ? gx : lx) : a == 0 ? (b > 0 ? gy : ly)
                                       : a == b ? (a > 0 ? gs : ls) : (a > 0 ? gd : ld);
            const V fs = v_sel(m_eq(ge, slot1), fn[q], f[q]);      // slot0 upstream & own in slot 0 -> own, ...
            fl[q] = v_fma(fs, ksi_dot<V>(q, Mx, My, Ms, Md), fl[q]);
        }
    } else {
        const V nGd = v_neg(Gd);
#pragma unroll
        for (int q = 1; q < Q; ++q) {
            const V W = ksi_dot<V>(q, Mx, My, Ms, Md);
            const V c = v_fma(W, nGd, A);
            const V fs = v_fma(v_sub(fn[q], f[q]), c, v_sel(slot1, fn[q], f[q]));
            fl[q] = v_fma(fs, W, fl[q]);
        }
    }
}

// BGK relaxation + flux divergence (src/containers.py:121):  out = f + dt ( (feq - f)/tau - fl )
template <typename V, int Q>
FVDBM_HD void relax_update(V* out, const V* f, const V* fl, const Params<typename VT<V>::S>& P) {
    V rho, ux, uy;
    moments<V, Q>(f, rho, ux, uy);
    const Equilibrium<V, Q> E(rho, ux, uy, P);
    V wr[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) wr[q] = (q == wclass(q)) ? E.wrho(q, P) : wr[wclass(q)];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const V g = v_fma(wr[q], E.poly(q, P), v_neg(f[q]));                      // feq - f
        out[q] = v_fma(v_bcast<V>(P.dt), v_fma(v_bcast<V>(P.inv_tau), g, v_neg(fl[q])), f[q]);
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
};

// populations on the far side of one side of ONE cell: the neighbour's (gathered by `gather(pos, fn)`)
// or the ghost cell's (src/containers.py:285-287: mean of the two node PDFs, extrapolated through it)
template <typename S, int Q, typename Gather>
FVDBM_HD void far_populations(const GhostTables<S>& G, int32_t cd, const S* f, Gather&& gather, S* fn) {
    if (cd >= 0) {
        gather((int64_t)(cd >> 2), fn);
    } else {
        const int32_t b = (-(cd + 1)) >> 2;
        const S ratio = FVDBM_LDG(G.bf_ratio + b);
        const int32_t na = FVDBM_LDG(G.bf_na + b), nb = FVDBM_LDG(G.bf_nb + b);
#pragma unroll
        for (int q = 1; q < Q; ++q) {
            const S g = v_mul(v_add(FVDBM_LDG(G.npdf + q * G.NTpad + na), FVDBM_LDG(G.npdf + q * G.NTpad + nb)), S(0.5));
            fn[q] = v_fma(v_sub(g, f[q]), ratio, g);
        }
    }
    fn[0] = S(0);
}
FVDBM_HD int code_slot(int32_t cd) { return (cd >= 0 ? cd : -(cd + 1)) & 1; }
FVDBM_HD int code_neg(int32_t cd) { return ((cd >= 0 ? cd : -(cd + 1)) >> 1) & 1; }
FVDBM_HD int32_t code_index(int32_t cd) { return (cd >= 0 ? cd : -(cd + 1)) >> 2; }

// One cell (scalar V) of the fused step: K sides, then relaxation + update.  `code` / `coef` hold this
// cell's K side codes and K*NC side coefficients (plan.hpp).  The packed kernel has its own driver
// (kernels.cuh: k_fused_pair) around the same side_flux / relax_update.
template <typename real, int Q, int K, int SCHEME, typename Gather>
FVDBM_HD void advance_cell(const Params<real>& P, const GhostTables<real>& G, const real* f, const int32_t* code,
                           const real* coef, Gather&& gather, real* out) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    real fl[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) fl[q] = real(0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        real fn[Q];
        far_populations<real, Q>(G, code[k], f, gather, fn);
        real A = real(0), Gd = real(0);
        if (SCHEME != 0) { A = coef[k * NC + 2]; Gd = v_mul(P.dt, coef[k * NC + 3]); }
        side_flux<real, Q, SCHEME>(fl, f, fn, code_slot(code[k]) != 0, code_neg(code[k]) != 0, coef[k * NC + 0],
                                   coef[k * NC + 1], A, Gd);
    }
    relax_update<real, Q>(out, f, fl, P);
}

// One active boundary node (src/containers.py:339-404): ring sums are supplied by the caller
// (warp-reduced on the GPU); returns the node's rho, vel and populations.
template <typename real, int Q>
FVDBM_HD void node_finish(const Params<real>& P, int type, real sw, real srho, real sux, real suy, const real* sneq,
                          real& rho_n, real& ux_n, real& uy_n, real* pdf_n) {
    const real inv = v_div(real(1), sw);        // weighted_avg (utils/utils.py:34-60): one IEEE reciprocal, then products
    if (type == 1) rho_n = v_mul(srho, inv);                                 // containers.py:348-351
    if (type == 2) { ux_n = v_mul(sux, inv); uy_n = v_mul(suy, inv); }       // containers.py:343-346
    const Equilibrium<real, Q> E(rho_n, ux_n, uy_n, P);
#pragma unroll
    for (int q = 0; q < Q; ++q)                                               // containers.py:353-361
        pdf_n[q] = v_fma(sneq[q], inv, E.value(q, P));
}

// contribution of one ring cell to a node's sums
template <typename real, int Q>
FVDBM_HD void node_accumulate(const Params<real>& P, const real* f, real w, real& sw, real& srho, real& sux, real& suy,
                              real* sneq) {
    real rho, ux, uy;
    moments<real, Q>(f, rho, ux, uy);
    const Equilibrium<real, Q> E(rho, ux, uy, P);
    sw = v_add(sw, w); srho = v_fma(rho, w, srho); sux = v_fma(ux, w, sux); suy = v_fma(uy, w, suy);
#pragma unroll
    for (int q = 0; q < Q; ++q) sneq[q] = v_fma(v_sub(f[q], E.value(q, P)), w, sneq[q]);
}

}  // namespace fvdbm
