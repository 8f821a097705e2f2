// hostsim.cpp -- TEST-ONLY CPU driver for the device code's per-cell functions.
//
// Compiles csrc/core.cuh (the FVDBM_HD arithmetic + side decoding used by the CUDA kernels) and
// csrc/plan.hpp (the host planner) with g++, and walks the planned AoSoA layout on the CPU exactly
// the way k_nodes + the fused cell kernels do on the GPU -- same canonical operation
// sequence, same reduction order, so the GPU results are required to be BIT-IDENTICAL to this walk.  tests/test_hostsim.py compares it with the
// oracle, so layout/encoding/arithmetic mistakes are caught in the GPU-less build container.
// This file is NOT part of libfvdbm_b200.so and no product path can reach it.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../fvdbm_jax_b200/csrc/core.cuh"
#include "../../fvdbm_jax_b200/csrc/plan.hpp"

using namespace fvdbm;

static std::string g_err;

template <typename real, int Q, int K, int SCHEME>
static int run(const fvdbm_desc& d, int nsteps, void* o_pdf, void* o_npdf, void* o_nrho, void* o_nvel,
               void* o_prev_rho, void* o_prev_vel) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    Plan<real> pl;
    if (!pl.build(d)) { g_err = pl.error; return -1; }
    if (!pl.fused_ok) { g_err = "mesh not fused-capable: " + pl.why_not; return -4; }
    Params<real> P{};
    for (int q = 0; q < 16; ++q) P.w[q] = (real)d.lat_w[q];
    P.inv_cs2 = real(1) / (real)d.cs2; P.inv_2cs4 = real(1) / (real)d.two_cs4; P.inv_2cs2 = real(1) / (real)d.two_cs2;
    P.inv_2cs6 = Q == 13 ? real(1) / (real)d.two_cs6 : real(0); P.three_inv_2cs4 = real(3) / (real)d.two_cs4;
    P.inv_tau = (real)(1.0 / d.tau); P.dt = (real)d.delta_t;
    const size_t ne = (size_t)(pl.Npad / TW) * Q * TW;
    std::vector<real> buf[2] = {std::vector<real>(ne, real(0)), std::vector<real>(ne, real(0))};
    const real* in = static_cast<const real*>(d.cell_pdf);
    for (int64_t i = 0; i < pl.N; ++i)
        for (int q = 0; q < Q; ++q) buf[0][pdf_index<Q>(pl.pos[i]) + q * kTW] = in[i * Q + q];
    buf[1] = buf[0];
    std::vector<real> npdf = pl.tn_pdf, nrho = pl.tn_rho, nvel = pl.tn_vel;
    int cur = 0;
    for (int s = 0; s < nsteps; ++s) {
        const real* pin = buf[cur].data();
        // k_nodes: lane l of the node's group accumulates ring slots l, l+8, ...; xor butterfly over the 8 lanes
        // (same order as the GPU)
        constexpr int NL = kNodeLanes;
        for (int64_t t = 0; t < pl.NA; ++t) {
            real sw[NL], srho[NL], sux[NL], suy[NL], sneq[NL][Q];
            for (int l = 0; l < NL; ++l) {
                sw[l] = srho[l] = sux[l] = suy[l] = 0;
                for (int q = 0; q < Q; ++q) sneq[l][q] = 0;
                for (int64_t j = l; j < pl.MR; j += NL) {
                    const real w = pl.ring_fw[(size_t)t * pl.MR + j];
                    if (!(w != real(0))) continue;
                    real f[Q];
                    for (int q = 0; q < Q; ++q) f[q] = pin[pdf_index<Q>(pl.ring_fcell[(size_t)t * pl.MR + j]) + q * kTW];
                    node_accumulate<real, Q>(P, f, w, sw[l], srho[l], sux[l], suy[l], sneq[l]);
                }
            }
            auto butterfly = [](real* v) {
                for (int o = NL / 2; o > 0; o >>= 1) {
                    real n[NL];
                    for (int l = 0; l < NL; ++l) n[l] = v[l] + v[l ^ o];
                    for (int l = 0; l < NL; ++l) v[l] = n[l];
                }
            };
            butterfly(sw); butterfly(srho); butterfly(sux); butterfly(suy);
            real sq[Q];
            for (int q = 0; q < Q; ++q) {
                real v[NL];
                for (int l = 0; l < NL; ++l) v[l] = sneq[l][q];
                butterfly(v);
                sq[q] = v[0];
            }
            real r = nrho[t], x = nvel[t], y = nvel[pl.NTpad + t], g[Q];
            node_finish<real, Q>(P, pl.tn_type[t], sw[0], srho[0], sux[0], suy[0], sq, r, x, y, g);
            if (pl.tn_type[t] == 1) nrho[t] = r;
            if (pl.tn_type[t] == 2) { nvel[t] = x; nvel[pl.NTpad + t] = y; }
            for (int q = 0; q < Q; ++q) npdf[(size_t)q * pl.NTpad + t] = g[q];
        }
        // k_fused_direct
        GhostTables<real> G{pl.bf_na.data(), pl.bf_nb.data(), pl.bf_ratio.data(), npdf.data(), pl.NTpad};
        real* pout = buf[cur ^ 1].data();
        for (int64_t c = 0; c < round_up(pl.Oend, PAD_TO); ++c) {
            const size_t tile = (size_t)(c >> 5); const int lane = (int)(c & 31);
            int32_t code[K]; real coef[K * NC], f[Q], out[Q];
            code[0] = pl.ccode[tile * (K * kTW) + lane];
            if (code[0] == kHole) continue;
            for (int k = 1; k < K; ++k) code[k] = pl.ccode[tile * (K * kTW) + k * kTW + lane];
            for (int i = 0; i < K * NC; ++i) coef[i] = pl.ccoef[tile * (K * NC * kTW) + i * kTW + lane];
            for (int q = 0; q < Q; ++q) f[q] = pin[tile * (Q * kTW) + q * kTW + lane];
            auto load_nbr = [pin](int64_t nb, real* fn) {
                for (int q = 1; q < Q; ++q) fn[q] = pin[pdf_index<Q>(nb) + q * kTW];
            };
            advance_cell<real, Q, K, SCHEME>(P, G, f, code, coef, load_nbr, out);
            for (int q = 0; q < Q; ++q) pout[tile * (Q * kTW) + q * kTW + lane] = out[q];
        }
        cur ^= 1;
    }
    real* op = static_cast<real*>(o_pdf);
    for (int64_t i = 0; i < pl.N; ++i)
        for (int q = 0; q < Q; ++q) op[i * Q + q] = buf[cur][pdf_index<Q>(pl.pos[i]) + q * kTW];
    real* on = static_cast<real*>(o_npdf); real* orh = static_cast<real*>(o_nrho); real* ov = static_cast<real*>(o_nvel);
    for (int64_t t = 0; t < pl.NT; ++t) {
        const int64_t n = pl.tn_orig[t];
        for (int q = 0; q < Q; ++q) on[n * Q + q] = npdf[(size_t)q * pl.NTpad + t];
        orh[n] = nrho[t]; ov[2 * n] = nvel[t]; ov[2 * n + 1] = nvel[pl.NTpad + t];
    }
    if (nsteps > 0 && o_prev_rho && o_prev_vel) {
        real* pr = static_cast<real*>(o_prev_rho); real* pv = static_cast<real*>(o_prev_vel);
        for (int64_t i = 0; i < pl.N; ++i) {
            real f[Q], r, x, y;
            for (int q = 0; q < Q; ++q) f[q] = buf[cur ^ 1][pdf_index<Q>(pl.pos[i]) + q * kTW];
            moments<real, Q>(f, r, x, y);
            pr[i] = r; pv[2 * i] = x; pv[2 * i + 1] = y;
        }
    }
    return 0;
}

template <typename real, int Q, int K>
static int by_scheme(const fvdbm_desc& d, int n, void* a, void* b, void* c, void* e, void* f, void* g) {
    return d.scheme == 0 ? run<real, Q, K, 0>(d, n, a, b, c, e, f, g) : run<real, Q, K, 1>(d, n, a, b, c, e, f, g);
}
template <typename real>
static int by_shape(const fvdbm_desc& d, int n, void* a, void* b, void* c, void* e, void* f, void* g) {
    if (d.Q == 9 && d.K == 3) return by_scheme<real, 9, 3>(d, n, a, b, c, e, f, g);
    if (d.Q == 9 && d.K == 4) return by_scheme<real, 9, 4>(d, n, a, b, c, e, f, g);
    if (d.Q == 13 && d.K == 3) return by_scheme<real, 13, 3>(d, n, a, b, c, e, f, g);
    if (d.Q == 13 && d.K == 4) return by_scheme<real, 13, 4>(d, n, a, b, c, e, f, g);
    g_err = "unsupported (Q,K)";
    return -4;
}

extern "C" int hostsim_run(const fvdbm_desc* d, int nsteps, void* o_pdf, void* o_npdf, void* o_nrho, void* o_nvel,
                           void* o_prev_rho, void* o_prev_vel) {
    if (d->dtype == 32) return by_shape<float>(*d, nsteps, o_pdf, o_npdf, o_nrho, o_nvel, o_prev_rho, o_prev_vel);
    return by_shape<double>(*d, nsteps, o_pdf, o_npdf, o_nrho, o_nvel, o_prev_rho, o_prev_vel);
}
extern "C" const char* hostsim_error() { return g_err.c_str(); }
