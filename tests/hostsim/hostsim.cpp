// hostsim.cpp -- TEST-ONLY CPU driver for the device code's per-cell functions.
//
// Compiles csrc/core.cuh (the FVDBM_HD arithmetic + side decoding used by the CUDA kernels) and
// csrc/plan.hpp (the host planner) with g++, and walks the planned AoSoA layout on the CPU exactly
// the way k_nodes + k_fused_direct do on the GPU.  tests/test_hostsim.py compares it with the
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
        // k_nodes
        for (int64_t t = 0; t < pl.NA; ++t) {
            real sw = 0, srho = 0, sux = 0, suy = 0, sneq[Q];
            for (int q = 0; q < Q; ++q) sneq[q] = 0;
            for (int i = pl.ring_off[t]; i < pl.ring_off[t + 1]; ++i) {
                real f[Q];
                for (int q = 0; q < Q; ++q) f[q] = pin[pdf_index<Q>(pl.ring_cell[i]) + q * kTW];
                node_accumulate<real, Q>(P, f, pl.ring_w[i], sw, srho, sux, suy, sneq);
            }
            real r = nrho[t], x = nvel[t], y = nvel[pl.NTpad + t], g[Q];
            node_finish<real, Q>(P, pl.tn_type[t], sw, srho, sux, suy, sneq, r, x, y, g);
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
            if (!(getenv("HOSTSIM_COEF_LAYOUT") && atoi(getenv("HOSTSIM_COEF_LAYOUT")) == 1)) {
                for (int i = 0; i < K * NC; ++i) coef[i] = pl.ccoef[tile * (K * NC * kTW) + i * kTW + lane];
            } else {          // face layout: shared record per face
                for (int k = 0; k < K; ++k) {
                    const int32_t rec = pl.cface[tile * (K * kTW) + k * kTW + lane];
                    for (int i = 0; i < NC; ++i) coef[k * NC + i] = pl.fcoef[(size_t)rec * NC + i];
                }
            }
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

// ---- temporal blocking mirror: same schedule and per-tile algorithm as api.cu superstep() / k_fused2 ----
template <typename real, int Q, int K, int SCHEME>
static int run_temporal(const fvdbm_desc& d, int npairs, void* o_pdf) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    constexpr int T2 = Plan<real>::T2;
    Plan<real> pl;
    if (!pl.build(d)) { g_err = pl.error; return -1; }
    if (!pl.fused_ok || !pl.t2_ok) { g_err = "temporal tiles unavailable"; return -4; }
    Params<real> P{};
    for (int q = 0; q < 16; ++q) P.w[q] = (real)d.lat_w[q];
    P.inv_cs2 = real(1) / (real)d.cs2; P.inv_2cs4 = real(1) / (real)d.two_cs4; P.inv_2cs2 = real(1) / (real)d.two_cs2;
    P.inv_2cs6 = Q == 13 ? real(1) / (real)d.two_cs6 : real(0); P.three_inv_2cs4 = real(3) / (real)d.two_cs4;
    P.inv_tau = (real)(1.0 / d.tau); P.dt = (real)d.delta_t;
    const size_t ne = (size_t)(pl.Npad / TW) * Q * TW;
    std::vector<real> buf[3] = {std::vector<real>(ne, real(0)), std::vector<real>(ne, real(0)), std::vector<real>(ne, real(0))};
    const real* in = static_cast<const real*>(d.cell_pdf);
    for (int64_t i = 0; i < pl.N; ++i)
        for (int q = 0; q < Q; ++q) buf[0][pdf_index<Q>(pl.pos[i]) + q * kTW] = in[i * Q + q];
    std::vector<real> npdf = pl.tn_pdf, nrho = pl.tn_rho, nvel = pl.tn_vel;
    const int64_t end = round_up(pl.Oend, PAD_TO);

    auto nodes = [&](const real* pin) {                       // k_nodes over every active node
        for (int64_t t = 0; t < pl.NA; ++t) {
            real sw = 0, srho = 0, sux = 0, suy = 0, sneq[Q];
            for (int q = 0; q < Q; ++q) sneq[q] = 0;
            for (int i = pl.ring_off[t]; i < pl.ring_off[t + 1]; ++i) {
                real f[Q];
                for (int q = 0; q < Q; ++q) f[q] = pin[pdf_index<Q>(pl.ring_cell[i]) + q * kTW];
                node_accumulate<real, Q>(P, f, pl.ring_w[i], sw, srho, sux, suy, sneq);
            }
            real r = nrho[t], x = nvel[t], y = nvel[pl.NTpad + t], g[Q];
            node_finish<real, Q>(P, pl.tn_type[t], sw, srho, sux, suy, sneq, r, x, y, g);
            if (pl.tn_type[t] == 1) nrho[t] = r;
            if (pl.tn_type[t] == 2) { nvel[t] = x; nvel[pl.NTpad + t] = y; }
            for (int q = 0; q < Q; ++q) npdf[(size_t)q * pl.NTpad + t] = g[q];
        }
    };
    auto single = [&](const real* pin, real* pout, int64_t c) {   // k_fused_direct for one position
        GhostTables<real> G{pl.bf_na.data(), pl.bf_nb.data(), pl.bf_ratio.data(), npdf.data(), pl.NTpad};
        const size_t tile = (size_t)(c >> 5); const int lane = (int)(c & 31);
        int32_t code[K]; real coef[K * NC], f[Q], out[Q];
        code[0] = pl.ccode[tile * (K * kTW) + lane];
        if (code[0] == kHole) return;
        for (int k = 1; k < K; ++k) code[k] = pl.ccode[tile * (K * kTW) + k * kTW + lane];
        for (int i = 0; i < K * NC; ++i) coef[i] = pl.ccoef[tile * (K * NC * kTW) + i * kTW + lane];
        for (int q = 0; q < Q; ++q) f[q] = pin[tile * (Q * kTW) + q * kTW + lane];
        auto load_nbr = [pin](int64_t nb, real* fn) { for (int q = 1; q < Q; ++q) fn[q] = pin[pdf_index<Q>(nb) + q * kTW]; };
        advance_cell<real, Q, K, SCHEME>(P, G, f, code, coef, load_nbr, out);
        for (int q = 0; q < Q; ++q) pout[tile * (Q * kTW) + q * kTW + lane] = out[q];
    };
    int cur = 0;
    for (int s = 0; s < npairs; ++s) {
        const real* A = buf[cur].data(); real* B = buf[(cur + 1) % 3].data(); real* C = buf[(cur + 2) % 3].data();
        // thin passes (main stream in the engine)
        nodes(A);
        for (int32_t c : pl.l2_list) single(A, B, c);
        for (int64_t c = pl.D1start; c < end; ++c) single(A, B, c);
        nodes(B);
        for (int64_t c = pl.D1start; c < end; ++c) single(B, C, c);
        // tiles (side stream in the engine): k_fused2
        GhostTables<real> G0{nullptr, nullptr, nullptr, nullptr, 0};
        for (int64_t t = 0; t < pl.t2_tiles; ++t) {
            const int64_t t0 = t * T2, off = pl.t2_off[t], n12 = pl.t2_off[t + 1] - off, n1 = pl.t2_n1[t];
            const int64_t n01 = T2 + n1, nent = T2 + n12;
            std::vector<real> s0((size_t)Q * nent), s1((size_t)Q * n01, real(0));
            auto entry_pos = [&](int64_t e) -> int64_t { return e < T2 ? t0 + e : pl.t2_pos[off + e - T2]; };
            for (int64_t e = 0; e < nent; ++e)
                for (int q = 0; q < Q; ++q) s0[(size_t)q * nent + e] = A[pdf_index<Q>(entry_pos(e)) + q * kTW];
            auto advance_entry = [&](int64_t e, const std::vector<real>& src, int64_t stride, real* out) -> bool {
                const uint16_t* l = &pl.t2_lnbr[(size_t)(pl.t2_loff[t] + e) * K];
                if (l[0] == 0xFFFF) return false;
                const int64_t pc = entry_pos(e);
                const size_t tile = (size_t)(pc >> 5); const int lane = (int)(pc & 31);
                int32_t code[K]; real coef[K * NC], f[Q];
                for (int k = 0; k < K; ++k) code[k] = (int32_t)l[k];
                for (int i = 0; i < K * NC; ++i) coef[i] = pl.ccoef[tile * (K * NC * kTW) + i * kTW + lane];
                for (int q = 0; q < Q; ++q) f[q] = src[(size_t)q * stride + e];
                auto load_nbr = [&](int64_t nb, real* fn) { for (int q = 1; q < Q; ++q) fn[q] = src[(size_t)q * stride + nb]; };
                advance_cell<real, Q, K, SCHEME>(P, G0, f, code, coef, load_nbr, out);
                return true;
            };
            for (int64_t e = 0; e < n01; ++e) {
                real out[Q];
                if (advance_entry(e, s0, nent, out)) for (int q = 0; q < Q; ++q) s1[(size_t)q * n01 + e] = out[q];
            }
            for (int64_t e = 0; e < T2; ++e) {
                real out[Q];
                if (advance_entry(e, s1, n01, out)) for (int q = 0; q < Q; ++q) C[pdf_index<Q>(t0 + e) + q * kTW] = out[q];
            }
        }
        cur = (cur + 2) % 3;
    }
    real* op = static_cast<real*>(o_pdf);
    for (int64_t i = 0; i < pl.N; ++i)
        for (int q = 0; q < Q; ++q) op[i * Q + q] = buf[cur][pdf_index<Q>(pl.pos[i]) + q * kTW];
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
extern "C" int hostsim_run_temporal(const fvdbm_desc* d, int npairs, void* o_pdf) {
#define T_DISPATCH(real)                                                                                   \
    if (d->Q == 9 && d->K == 3) return d->scheme == 0 ? run_temporal<real, 9, 3, 0>(*d, npairs, o_pdf) : run_temporal<real, 9, 3, 1>(*d, npairs, o_pdf); \
    if (d->Q == 9 && d->K == 4) return d->scheme == 0 ? run_temporal<real, 9, 4, 0>(*d, npairs, o_pdf) : run_temporal<real, 9, 4, 1>(*d, npairs, o_pdf); \
    if (d->Q == 13 && d->K == 3) return d->scheme == 0 ? run_temporal<real, 13, 3, 0>(*d, npairs, o_pdf) : run_temporal<real, 13, 3, 1>(*d, npairs, o_pdf); \
    if (d->Q == 13 && d->K == 4) return d->scheme == 0 ? run_temporal<real, 13, 4, 0>(*d, npairs, o_pdf) : run_temporal<real, 13, 4, 1>(*d, npairs, o_pdf);
    if (d->dtype == 32) { T_DISPATCH(float) } else { T_DISPATCH(double) }
#undef T_DISPATCH
    g_err = "unsupported (Q,K)";
    return -4;
}
extern "C" const char* hostsim_error() { return g_err.c_str(); }
