// kernels.cuh -- sm_100a kernels of the FVDBM step (see DESIGN.md for the roofline of each).
//
//   k_nodes         S3  boundary nodes: 8 lanes per active node, shuffle butterfly over its ring
//   k_fused_rec     S1+S2+S4+S5 cell-centric over the RECORD layout, one 256-bit access per record, FFMA2 over population
//                   pairs; fp32 D2Q9 default.   k_fused_pair: TWO cells per thread over AoSoA (fp32 D2Q13 >= 4M cells)
//   k_fused_direct  same arithmetic, one cell per thread, either layout (fp64 default; fp32 A/B partner)
//   k_fused_tma     same arithmetic; persistent CTAs, cp.async.bulk (TMA) + mbarrier ring of tiles
//   k_s_*           staged (reference-shaped) kernels S1/S2, S4, S5 -- general meshes + observables
//   k_export_* / k_import_* / k_pack / k_unpack   layout conversion at the API boundary
// All fused kernels run the one canonical operation sequence of core.cuh -> bit-identical results.
#pragma once
#include <cuda_runtime.h>
#include "core.cuh"

namespace fvdbm {

template <typename real>
struct FusedArgs {
    Params<real> P;
    const real* __restrict__ pdf_in;
    real* __restrict__ pdf_out;
    const int32_t* __restrict__ ccode;
    const real* __restrict__ ccoef;    // per-side coefficients, tiled like the codes
    GhostTables<real> G;               // boundary sides + tracked node populations
    int64_t cell_begin, cell_end;      // position range, multiples of PAD_TO (512)
    int reverse;                       // 1: sweep tiles from the top (L2 reuse of last step's writes)
    int64_t Npad;                      // padded position count (record layout: offset of the rest populations)
    int prefetch_dist;                 // >0: every CTA asks L2 to prefetch the streaming operands of the CTA
                                       // `prefetch_dist` blocks ahead (cp.async.bulk.prefetch.L2)
};

#ifndef FVDBM_PAIR_THREADS
#define FVDBM_PAIR_THREADS 128
#endif
#ifndef FVDBM_PAIR_MINCTAS
#define FVDBM_PAIR_MINCTAS 5          // A/B on B200 (profiles/r2_ab_pair_kernel.jsonl): 128x5 (96 regs) beats 128x4, 128x6 (spills),
#endif                                // 256x2 and 64x8 in both the burst and the power-capped sustained regime

// Programmatic dependent launch (PDL): a kernel launched with programmaticStreamSerialization may start while its
// predecessor in the stream is still running; it must pass pdl_wait() before touching anything the predecessor
// writes.  Every kernel triggers its own dependents only AFTER its wait, so at any time at most two consecutive
// kernels overlap and everything older is complete.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// The AoSoA layout makes the streaming operands of any run of `cells` consecutive positions three
// contiguous blocks (populations, side coefficients, side codes): one bulk L2 prefetch each, issued by a
// single thread, moves the DRAM latency of a later CTA's first loads off its warps.
template <typename real, int Q, int K, int NC>
__device__ __forceinline__ void prefetch_cells_l2(const FusedArgs<real>& a, int64_t first_cell, int cells) {
    if (first_cell < a.cell_begin || first_cell + cells > a.cell_end) return;
    const size_t mt = (size_t)(first_cell >> 5);
    prefetch_l2_bulk(a.pdf_in + mt * (Q * kTW), (uint32_t)(cells * Q * sizeof(real)));
    prefetch_l2_bulk(a.ccoef + mt * (K * NC * kTW), (uint32_t)(cells * K * NC * sizeof(real)));
    prefetch_l2_bulk(a.ccode + mt * (K * kTW), (uint32_t)(cells * K * sizeof(int32_t)));
}
// same for the record layout: records and rest populations of the run are two contiguous blocks
template <typename real, int Q, int K, int NC>
__device__ __forceinline__ void prefetch_cells_rec_l2(const FusedArgs<real>& a, int64_t first_cell, int cells) {
    if (first_cell < a.cell_begin || first_cell + cells > a.cell_end) return;
    const size_t mt = (size_t)(first_cell >> 5);
    prefetch_l2_bulk(a.pdf_in + (size_t)first_cell * (Q - 1), (uint32_t)(cells * (Q - 1) * sizeof(real)));
    prefetch_l2_bulk(a.pdf_in + (size_t)a.Npad * (Q - 1) + first_cell, (uint32_t)(cells * sizeof(real)));
    prefetch_l2_bulk(a.ccoef + mt * (K * NC * kTW), (uint32_t)(cells * K * NC * sizeof(real)));
    prefetch_l2_bulk(a.ccode + mt * (K * kTW), (uint32_t)(cells * K * sizeof(int32_t)));
}

// measured on B200 (profiles/r1_experiment_occupancy_cachehints.txt): fp32 is best left to ptxas
// (5 CTAs/SM; forcing 6 or 8 CTAs spills and loses 2-18 %), fp64 gains 15 % from 3 CTAs/SM.
#define FVDBM_DIRECT_BOUNDS __launch_bounds__(256, (sizeof(real) == 8 ? 3 : (Q == 9 ? 5 : 4)))

// record layout (core.cuh: lay 1): the Q-1 moving populations of `cell` with 128-bit accesses
template <int Q>
__device__ __forceinline__ void load_record(const float* __restrict__ base, int64_t cell, float* out) {      // out[1..Q-1]
    const float4* p = reinterpret_cast<const float4*>(base) + cell * ((Q - 1) / 4);
#pragma unroll
    for (int i = 0; i < (Q - 1) / 4; ++i) {
        const float4 v = __ldg(p + i);
        out[1 + 4 * i] = v.x; out[2 + 4 * i] = v.y; out[3 + 4 * i] = v.z; out[4 + 4 * i] = v.w;
    }
}
template <int Q>
__device__ __forceinline__ void load_record(const double* __restrict__ base, int64_t cell, double* out) {
    const double2* p = reinterpret_cast<const double2*>(base) + cell * ((Q - 1) / 2);
#pragma unroll
    for (int i = 0; i < (Q - 1) / 2; ++i) {
        const double2 v = __ldg(p + i);
        out[1 + 2 * i] = v.x; out[2 + 2 * i] = v.y;
    }
}
template <int Q>
__device__ __forceinline__ void store_record(float* __restrict__ base, int64_t cell, const float* in) {
    float4* p = reinterpret_cast<float4*>(base) + cell * ((Q - 1) / 4);
#pragma unroll
    for (int i = 0; i < (Q - 1) / 4; ++i) p[i] = make_float4(in[1 + 4 * i], in[2 + 4 * i], in[3 + 4 * i], in[4 + 4 * i]);
}
template <int Q>
__device__ __forceinline__ void store_record(double* __restrict__ base, int64_t cell, const double* in) {
    double2* p = reinterpret_cast<double2*>(base) + cell * ((Q - 1) / 2);
#pragma unroll
    for (int i = 0; i < (Q - 1) / 2; ++i) p[i] = make_double2(in[1 + 2 * i], in[2 + 2 * i]);
}

// ------------------------------------------------------------------------------------------------
// V1: thread per cell, everything through L1/L2.  LAY = 0: tiled AoSoA populations; LAY = 1: record layout
// (neighbour gathers are (Q-1)*sizeof(real)/16 128-bit loads from one or two sectors; fp64 and D2Q13 use this
// kernel for FVDBM_VARIANT_REC, fp32 D2Q9 has the packed k_fused_rec below).
// ------------------------------------------------------------------------------------------------
template <typename real, int Q, int K, int SCHEME, int LAY>
__global__ void FVDBM_DIRECT_BOUNDS k_fused_direct(const FusedArgs<real> a) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    const int64_t nblk = gridDim.x;
    const int64_t blk = a.reverse ? (nblk - 1 - blockIdx.x) : blockIdx.x;
    const int64_t c = a.cell_begin + blk * blockDim.x + threadIdx.x;
    if (a.prefetch_dist > 0 && threadIdx.x == 0) {
        const int64_t first = a.cell_begin + (blk + (a.reverse ? -a.prefetch_dist : a.prefetch_dist)) * (int64_t)blockDim.x;
        if (LAY == 0) prefetch_cells_l2<real, Q, K, NC>(a, first, (int)blockDim.x);
        else prefetch_cells_rec_l2<real, Q, K, NC>(a, first, (int)blockDim.x);
    }
    if (c >= a.cell_end) return;
    const size_t tile = (size_t)(c >> 5);
    const int lane = (int)(c & 31);
    // Issue every independent streaming load (side codes, side coefficients, own populations) BEFORE the
    // first use of any of them (a full memory round trip was otherwise spent on the hole test of code[0]).
    // An early `return` would let ptxas sink the loads below it again, so padding positions are not
    // skipped: they run the (in-bounds, harmless) arithmetic on a neutral code and only their stores are
    // suppressed.
    const int32_t* gc = a.ccode + tile * (K * kTW) + lane;
    int32_t code[K];
#pragma unroll
    for (int k = 0; k < K; ++k) code[k] = __ldg(gc + k * kTW);
    real coef[K * NC];
    const real* gco = a.ccoef + tile * (K * NC * kTW) + lane;
#pragma unroll
    for (int i = 0; i < K * NC; ++i) coef[i] = __ldg(gco + i * kTW);
    real f[Q], out[Q];
    if (LAY == 0) {
        const real* gp = a.pdf_in + tile * (Q * kTW) + lane;
#pragma unroll
        for (int q = 0; q < Q; ++q) f[q] = __ldg(gp + q * kTW);
    } else {
        load_record<Q>(a.pdf_in, c, f);
        f[0] = __ldg(a.pdf_in + (size_t)a.Npad * (Q - 1) + c);
    }
    const bool live = code[0] != kHole;
    if (!live) code[0] = 0;                        // neutral: interior side towards position 0
    // PDL: everything above reads data older than the predecessor (the node kernel); only the ghost sides below read
    // what it writes.  The streaming loads stay in flight across the wait.
    pdl_wait();
    pdl_trigger();
    const real* pin = a.pdf_in;
    auto gather = [pin](int64_t nb, real* fn) {
        if (LAY == 0) {
            const real* pn = pin + pdf_index<Q>(nb);
#pragma unroll
            for (int q = 1; q < Q; ++q) fn[q] = __ldg(pn + q * kTW);
        } else load_record<Q>(pin, nb, fn);
    };
    advance_cell<real, Q, K, SCHEME>(a.P, a.G, f, code, coef, gather, out);
    if (live) {
        if (LAY == 0) {
            real* go = a.pdf_out + tile * (Q * kTW) + lane;
#pragma unroll
            for (int q = 0; q < Q; ++q) go[q * kTW] = out[q];
        } else {
            store_record<Q>(a.pdf_out, c, out);
            a.pdf_out[(size_t)a.Npad * (Q - 1) + c] = out[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// V3 (fp32): TWO cells per thread.  Thread t of a warp owns positions (2j, 2j+1) of one 32-wide AoSoA
// tile (16 threads per tile, a warp covers two tiles), so every streaming operand -- own populations,
// side codes, side coefficients, results -- is ONE 64-bit access per thread (two full 128-byte lines
// per warp request), and the whole per-cell arithmetic runs on packed pairs: FFMA2 / FADD2 / FMUL2 do
// the work of two scalar instructions in one issue slot with identical per-lane rounding.  Only the
// neighbour gathers (8 populations per side and cell, through L1/L2) and the f_slot0 select stay
// 32-bit.  Per cell this issues ~45 % fewer instructions than the thread-per-cell kernel, which is
// what was limiting it once the power cap pulls the SM clock down.
// ------------------------------------------------------------------------------------------------
template <int Q, int K, int SCHEME>
__global__ void __launch_bounds__(FVDBM_PAIR_THREADS, FVDBM_PAIR_MINCTAS) k_fused_pair(const FusedArgs<float> a) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    const int64_t nblk = gridDim.x;
    const int64_t blk = a.reverse ? (nblk - 1 - blockIdx.x) : blockIdx.x;
    const int64_t c = a.cell_begin + 2 * (blk * blockDim.x + threadIdx.x);       // even position; pair (c, c+1)
    if (a.prefetch_dist > 0 && threadIdx.x == 0)
        prefetch_cells_l2<float, Q, K, NC>(a, a.cell_begin + 2 * (blk + (a.reverse ? -a.prefetch_dist : a.prefetch_dist)) * (int64_t)blockDim.x,
                                           2 * (int)blockDim.x);
    if (c >= a.cell_end) return;
    const size_t tile = (size_t)(c >> 5);
    const int lane = (int)(c & 31);
    const int2* gc = reinterpret_cast<const int2*>(a.ccode + tile * (K * kTW) + lane);
    int2 code[K];
#pragma unroll
    for (int k = 0; k < K; ++k) code[k] = __ldg(gc + k * (kTW / 2));
    const float2* gco = reinterpret_cast<const float2*>(a.ccoef + tile * (K * NC * kTW) + lane);
    float2 coef[K * NC];
#pragma unroll
    for (int i = 0; i < K * NC; ++i) coef[i] = __ldg(gco + i * (kTW / 2));
    const float2* gp = reinterpret_cast<const float2*>(a.pdf_in + tile * (Q * kTW) + lane);
    float2 f[Q], out[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q] = __ldg(gp + q * (kTW / 2));
    const bool live0 = code[0].x != kHole, live1 = code[0].y != kHole;
    if (!live0) code[0].x = 0;                     // neutral: interior side towards position 0
    if (!live1) code[0].y = 0;
    pdl_wait();                                    // see k_fused_direct
    pdl_trigger();
    const float* pin = a.pdf_in;
    auto gather = [pin](int64_t nb, float* fn) {
        const float* pn = pin + pdf_index<Q>(nb);
#pragma unroll
        for (int q = 1; q < Q; ++q) fn[q] = __ldg(pn + q * kTW);
    };
    float2 fl[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) fl[q] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float fx[Q], fy[Q];
        float2 fn[Q];
        if (code[k].x >= 0 && code[k].y >= 0) {     // both sides interior (all but the O(sqrt N) border cells)
            gather((int64_t)(code[k].x >> 2), fx);
            gather((int64_t)(code[k].y >> 2), fy);
        } else {
            float f0[Q], f1[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) { f0[q] = f[q].x; f1[q] = f[q].y; }
            far_populations<float, Q>(a.G, code[k].x, f0, gather, fx);
            far_populations<float, Q>(a.G, code[k].y, f1, gather, fy);
        }
#pragma unroll
        for (int q = 1; q < Q; ++q) fn[q] = make_float2(fx[q], fy[q]);
        fn[0] = make_float2(0.0f, 0.0f);
        float2 A = make_float2(0.0f, 0.0f), Gd = A;
        if (SCHEME != 0) { A = coef[k * NC + 2]; Gd = v_mul(v_bcast<float2>(a.P.dt), coef[k * NC + 3]); }
        const bool2 slot1{code_slot(code[k].x) != 0, code_slot(code[k].y) != 0};
        const bool2 neg{code_neg(code[k].x) != 0, code_neg(code[k].y) != 0};
        side_flux<float2, Q, SCHEME>(fl, f, fn, slot1, neg, coef[k * NC + 0], coef[k * NC + 1], A, Gd);
    }
    relax_update<float2, Q>(out, f, fl, a.P);
    float2* go = reinterpret_cast<float2*>(a.pdf_out + tile * (Q * kTW) + lane);
    if (live0 && live1) {
#pragma unroll
        for (int q = 0; q < Q; ++q) go[q * (kTW / 2)] = out[q];
    } else {
        float* g1 = a.pdf_out + tile * (Q * kTW) + lane;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (live0) g1[q * kTW] = out[q].x;
            if (live1) g1[q * kTW + 1] = out[q].y;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// V4 (fp32, D2Q9): thread per cell over the RECORD layout (core.cuh: lay 1).  The eight moving populations of a
// cell are one 32-byte sector, so a neighbour gather is ONE 256-bit load (LDG.E.ENL2.256) of ONE sector (the
// AoSoA kernels touch eight sectors with eight LDG.E.32), the own populations are one 256-bit + one 32-bit load,
// the result one 256-bit + one 32-bit store; and the arithmetic is packed over POPULATION pairs (q1,q2) (q3,q4) (q5,q6) (q7,q8):
// per side the four distinct KSI.M values form two pairs W0 = (Mx, My), W2 = (Mx+My, My-Mx) and their negations,
// so c = A - W Gd, f* = f_slot0 + (fn - f) c and fl += f* W are four FFMA2 each, with scalar side coefficients as
// broadcast operands.  Same canonical operation sequence per (cell, population) as every other kernel -> same bits.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }
// one D2Q9 fp32 record (8 moving populations, 32 bytes, 32-byte aligned) = ONE 256-bit access (sm_100: LDG.E.ENL2.256)
__device__ __forceinline__ void ld_record256(const float* p, float2* o) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(o[0].x), "=f"(o[0].y), "=f"(o[1].x), "=f"(o[1].y), "=f"(o[2].x), "=f"(o[2].y), "=f"(o[3].x), "=f"(o[3].y)
                 : "l"(p));
}
__device__ __forceinline__ void st_record256(float* p, const float2* o) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(o[0].x), "f"(o[0].y), "f"(o[1].x), "f"(o[1].y), "f"(o[2].x), "f"(o[2].y), "f"(o[3].x), "f"(o[3].y)
                 : "memory");
}

#ifndef FVDBM_REC_THREADS
#define FVDBM_REC_THREADS 256
#endif
#ifndef FVDBM_REC_MINCTAS
#define FVDBM_REC_MINCTAS 5           // A/B on B200: 256x5 (48 regs, no spills) 0.2058 ms sustained vs 256x3/4 (56 regs) 0.2093, 128x8 0.2072
#endif
template <int K, int SCHEME>
__global__ void __launch_bounds__(FVDBM_REC_THREADS, FVDBM_REC_MINCTAS) k_fused_rec(const FusedArgs<float> a) {
    constexpr int Q = 9, NC = SCHEME == 0 ? 2 : 4;
    const int64_t Npad = a.Npad;
    const int64_t nblk = gridDim.x;
    const int64_t blk = a.reverse ? (nblk - 1 - blockIdx.x) : blockIdx.x;
    const int64_t c = a.cell_begin + blk * blockDim.x + threadIdx.x;
    const float* rest_in = a.pdf_in + (size_t)Npad * (Q - 1);
    if (a.prefetch_dist > 0 && threadIdx.x == 0)
        prefetch_cells_rec_l2<float, Q, K, NC>(a, a.cell_begin + (blk + (a.reverse ? -a.prefetch_dist : a.prefetch_dist)) * (int64_t)blockDim.x,
                                               (int)blockDim.x);
    if (c >= a.cell_end) return;
    const size_t tile = (size_t)(c >> 5);
    const int lane = (int)(c & 31);
    const int32_t* gc = a.ccode + tile * (K * kTW) + lane;
    int32_t code[K];
#pragma unroll
    for (int k = 0; k < K; ++k) code[k] = __ldg(gc + k * kTW);
    float coef[K * NC];
    const float* gco = a.ccoef + tile * (K * NC * kTW) + lane;
#pragma unroll
    for (int i = 0; i < K * NC; ++i) coef[i] = __ldg(gco + i * kTW);
    float2 f[4];                                                                          // (q1,q2) (q3,q4) (q5,q6) (q7,q8)
    ld_record256(a.pdf_in + (size_t)c * (Q - 1), f);
    const float f0 = __ldg(rest_in + c);
    const bool live = code[0] != kHole;
    if (!live) code[0] = 0;                        // neutral: interior side towards position 0
    pdl_wait();
    pdl_trigger();
    float2 fl[4] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int32_t cd = code[k];
        float2 fn[4];
        if (cd >= 0) {
            ld_record256(a.pdf_in + (size_t)(cd >> 2) * (Q - 1), fn);
        } else {                                    // ghost side (border cells only): scalar path of core.cuh
            const float fo[Q] = {f0, f[0].x, f[0].y, f[1].x, f[1].y, f[2].x, f[2].y, f[3].x, f[3].y};
            float g[Q];
            far_populations<float, Q>(a.G, cd, fo, [](int64_t, float*) {}, g);
            fn[0] = f2(g[1], g[2]); fn[1] = f2(g[3], g[4]); fn[2] = f2(g[5], g[6]); fn[3] = f2(g[7], g[8]);
        }
        const bool slot1 = code_slot(cd) != 0, neg = code_neg(cd) != 0;
        const float Mx = coef[k * NC + 0], My = coef[k * NC + 1];
        const float Ms = v_add(Mx, My), Md = v_sub(Mx, My), nMd = v_sub(My, Mx);     // nMd == -Md exactly
        const float2 W0 = f2(Mx, My), W2 = f2(Ms, nMd);
        const float2 W[4] = {W0, v_neg(W0), W2, v_neg(W2)};                          // KSI_q . M for q = 1..8
        if (SCHEME == 0) {
            // varpi_q >= 0 <=> sigma W_q >= 0; the own cell is upstream iff (slot == 0) == (varpi_q >= 0)   (core.cuh)
            const bool gx = neg ? Mx <= 0.f : Mx >= 0.f, lx = neg ? Mx >= 0.f : Mx <= 0.f;
            const bool gy = neg ? My <= 0.f : My >= 0.f, ly = neg ? My >= 0.f : My <= 0.f;
            const bool gs = neg ? Ms <= 0.f : Ms >= 0.f, ls = neg ? Ms >= 0.f : Ms <= 0.f;
            const bool gd = neg ? Md <= 0.f : Md >= 0.f, ld = neg ? Md >= 0.f : Md <= 0.f;
            const bool2 ge[4] = {bool2{gx, gy}, bool2{lx, ly}, bool2{gs, ld}, bool2{ls, gd}};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 fs = v_sel(bool2{ge[j].x == slot1, ge[j].y == slot1}, fn[j], f[j]);
                fl[j] = v_fma(fs, W[j], fl[j]);
            }
        } else {
            const float A = coef[k * NC + 2], Gd = v_mul(a.P.dt, coef[k * NC + 3]);
            const float2 nGd = v_bcast<float2>(-Gd), A2 = v_bcast<float2>(A);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 cj = v_fma(W[j], nGd, A2);
                const float2 fs = v_fma(v_sub(fn[j], f[j]), cj, slot1 ? fn[j] : f[j]);
                fl[j] = v_fma(fs, W[j], fl[j]);
            }
        }
    }
    // moments, equilibrium, relaxation: the scalar canonical order for the sums, packed pairs for the rest
    float r = f0;
    r = v_add(r, f[0].x); r = v_add(r, f[0].y); r = v_add(r, f[1].x); r = v_add(r, f[1].y);
    r = v_add(r, f[2].x); r = v_add(r, f[2].y); r = v_add(r, f[3].x); r = v_add(r, f[3].y);
    float jx = f[0].x, jy = f[0].y;                                   // q1, q2
    jx = v_sub(jx, f[1].x); jx = v_add(jx, f[2].x); jx = v_sub(jx, f[2].y); jx = v_sub(jx, f[3].x); jx = v_add(jx, f[3].y);
    jy = v_sub(jy, f[1].y); jy = v_add(jy, f[2].x); jy = v_add(jy, f[2].y); jy = v_sub(jy, f[3].x); jy = v_sub(jy, f[3].y);
    const float ux = v_div(jx, r), uy = v_div(jy, r);
    const float uu = v_fma(ux, ux, v_mul(uy, uy));
    const float us = v_add(ux, uy), nud = v_sub(uy, ux);               // nud == -(ux - uy) exactly
    const float base = v_fma(-uu, a.P.inv_2cs2, 1.0f);
    const float2 ku0 = f2(ux, uy), ku2 = f2(us, nud);
    const float2 ku[4] = {ku0, v_neg(ku0), ku2, v_neg(ku2)};
    const float2 b2 = v_bcast<float2>(a.P.inv_2cs4), a2 = v_bcast<float2>(a.P.inv_cs2), base2 = v_bcast<float2>(base);
    const float wr0 = v_mul(a.P.w[0], r), wr1 = v_mul(a.P.w[1], r), wr5 = v_mul(a.P.w[5], r);
    const float2 dt2 = v_bcast<float2>(a.P.dt), it2 = v_bcast<float2>(a.P.inv_tau);
    float2 out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 poly = v_fma(ku[j], v_fma(ku[j], b2, a2), base2);
        const float2 g = v_fma(v_bcast<float2>(j < 2 ? wr1 : wr5), poly, v_neg(f[j]));
        out[j] = v_fma(dt2, v_fma(it2, g, v_neg(fl[j])), f[j]);
    }
    const float g0 = v_fma(wr0, base, -f0);                            // poly_0 == base exactly (KSI_0 = 0)
    const float out0 = v_fma(a.P.dt, v_fma(a.P.inv_tau, g0, -0.0f), f0);
    if (live) {
        st_record256(a.pdf_out + (size_t)c * (Q - 1), out);
        a.pdf_out[(size_t)Npad * (Q - 1) + c] = out0;
    }
}

// ------------------------------------------------------------------------------------------------
// V2: persistent CTAs; each CTA walks tiles of blockDim.x cells.  Thread 0 keeps `stages-1` tiles
// in flight with cp.async.bulk (TMA bulk copies: populations, side codes, side coefficients are
// each one contiguous block thanks to the AoSoA layout) completing on per-stage mbarriers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

constexpr int kTmaHeader = 128;   // bytes reserved for the mbarriers in front of the stage ring

template <typename real, int Q, int K, int SCHEME>
__host__ __device__ constexpr size_t tma_stage_bytes(int tile_cells) {
    return (size_t)tile_cells * ((Q + K * (SCHEME == 0 ? 2 : 4)) * sizeof(real) + K * sizeof(int32_t));
}

template <typename real, int Q, int K, int SCHEME>
__global__ void __launch_bounds__(512) k_fused_tma(const FusedArgs<real> a, const int stages) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    extern __shared__ __align__(128) unsigned char smem[];
    const int TC = blockDim.x;
    const int tid = threadIdx.x;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    const size_t pdf_bytes = (size_t)TC * Q * sizeof(real);
    const size_t coef_bytes = (size_t)TC * K * NC * sizeof(real);
    const size_t code_bytes = (size_t)TC * K * sizeof(int32_t);
    const size_t stage_bytes = pdf_bytes + coef_bytes + code_bytes;
    unsigned char* ring = smem + kTmaHeader;

    const int64_t ntiles = (a.cell_end - a.cell_begin) / TC;
    const int64_t first = blockIdx.x, stride = gridDim.x;

    auto tile_base = [&](int64_t i) -> int64_t {        // i-th tile of this CTA -> first position
        const int64_t t = first + i * stride;
        return a.cell_begin + (a.reverse ? (ntiles - 1 - t) : t) * TC;
    };
    auto issue = [&](int s, int64_t base) {
        unsigned char* st = ring + (size_t)s * stage_bytes;
        const size_t mt = (size_t)(base >> 5);
        mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
        bulk_g2s(st, a.pdf_in + mt * (Q * kTW), (uint32_t)pdf_bytes, &full[s]);
        bulk_g2s(st + pdf_bytes, a.ccoef + mt * (K * NC * kTW), (uint32_t)coef_bytes, &full[s]);
        bulk_g2s(st + pdf_bytes + coef_bytes, a.ccode + mt * (K * kTW), (uint32_t)code_bytes, &full[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    pdl_wait();
    pdl_trigger();
    __syncthreads();
    const int64_t my_tiles = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    if (tid == 0)
        for (int s = 0; s < stages - 1 && s < my_tiles; ++s) issue(s, tile_base(s));

    const int mt_local = tid >> 5, lane = tid & 31;
    for (int64_t it = 0; it < my_tiles; ++it) {
        const int s = (int)(it % stages);
        const uint32_t parity = (uint32_t)((it / stages) & 1);
        if (tid == 0) {
            const int64_t nx = it + stages - 1;
            if (nx < my_tiles) issue((int)(nx % stages), tile_base(nx));
        }
        while (!mbar_try_wait(&full[s], parity)) {}
        const int64_t base = tile_base(it);
        unsigned char* st = ring + (size_t)s * stage_bytes;
        const real* s_pdf = reinterpret_cast<const real*>(st);
        const real* s_coef = reinterpret_cast<const real*>(st + pdf_bytes) + (size_t)mt_local * (K * NC * kTW) + lane;
        const int32_t* s_code = reinterpret_cast<const int32_t*>(st + pdf_bytes + coef_bytes) + (size_t)mt_local * (K * kTW) + lane;
        int32_t code[K];
        code[0] = s_code[0];
        if (code[0] != kHole) {
#pragma unroll
            for (int k = 1; k < K; ++k) code[k] = s_code[k * kTW];
            real coef[K * NC];
#pragma unroll
            for (int i = 0; i < K * NC; ++i) coef[i] = s_coef[i * kTW];
            const real* sp = s_pdf + (size_t)mt_local * (Q * kTW) + lane;
            real f[Q], out[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) f[q] = sp[q * kTW];
            const real* pin = a.pdf_in;
            auto gather = [pin, s_pdf, base, TC](int64_t nb, real* fn) {
                const int64_t loc = nb - base;
                if (loc >= 0 && loc < TC) {            // neighbour staged in this tile: shared memory
                    const real* pn = s_pdf + pdf_index<Q>(loc);
#pragma unroll
                    for (int q = 1; q < Q; ++q) fn[q] = pn[q * kTW];
                } else {                                // halo of the tile: L2 / L1
                    const real* pn = pin + pdf_index<Q>(nb);
#pragma unroll
                    for (int q = 1; q < Q; ++q) fn[q] = __ldg(pn + q * kTW);
                }
            };
            advance_cell<real, Q, K, SCHEME>(a.P, a.G, f, code, coef, gather, out);
            real* go = a.pdf_out + ((size_t)(base >> 5) + mt_local) * (Q * kTW) + lane;
#pragma unroll
            for (int q = 0; q < Q; ++q) go[q * kTW] = out[q];
        }
        __syncthreads();      // stage s may be refilled by the next iteration's issue
    }
}

// ------------------------------------------------------------------------------------------------
// S3 boundary nodes (src/containers.py:339-404): warp per active node.
// ------------------------------------------------------------------------------------------------
template <typename real>
struct NodeArgs {
    Params<real> P;
    const real* __restrict__ pdf;          // current populations
    int lay; int64_t Npad;                 // their layout (core.cuh: pdf_off)
    const int32_t* __restrict__ ring_cell;  // fixed-width ring table [NA][MR] (plan.hpp: ring_fcell)
    const real* __restrict__ ring_w;        //                               (ring_fw; 0 = unused slot)
    int MR;
    const int32_t* __restrict__ tn_type;
    real* __restrict__ npdf;               // [Q][NTpad]
    real* __restrict__ nrho;               // [NTpad]
    real* __restrict__ nvel;               // [2][NTpad]
    int64_t NTpad;
    int NA;
};

template <typename real>
__device__ __forceinline__ real group_sum(real v) {          // xor butterfly inside an aligned group of kNodeLanes lanes
#pragma unroll
    for (int o = kNodeLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// kNodeLanes (8) lanes per active node, four nodes per warp: lane s of a group walks ring slots s, s+8, ... (one
// ring cell per lane covers the usual vertex valence of 6-8), the 13 partial sums are combined by a 3-level xor
// butterfly (every lane of the group ends with the same totals), and lane 0 stores rho / vel / populations.
// (A full warp per node left 24+ lanes idle and spent ~1500 warp-instructions per node: 26 us for the 19.5k
// boundary nodes of the porous config; profiles/r2_launches_porous.csv.)
template <typename real, int Q>
__global__ void __launch_bounds__(256) k_nodes(const NodeArgs<real> a) {
    const int gtid = (int)(blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
    const int sub = threadIdx.x & (kNodeLanes - 1);
    const bool valid = gtid / kNodeLanes < a.NA;
    const int node = valid ? gtid / kNodeLanes : 0;          // idle groups shadow node 0 so that shuffles stay convergent
    // static data (type, ring table) is independent of the predecessor kernel: issue those loads before the PDL wait
    const int type = a.tn_type[node];
    const size_t i0 = (size_t)node * a.MR + sub;
    real w_first = real(0);
    int32_t c_first = 0;
    if (sub < a.MR) { w_first = a.ring_w[i0]; c_first = a.ring_cell[i0]; }
    pdl_wait();                                    // the ring gathers read the populations the previous cell kernel wrote
    pdl_trigger();
    real rho_n = a.nrho[node], ux_n = a.nvel[node], uy_n = a.nvel[a.NTpad + node];
    real sw = real(0), srho = real(0), sux = real(0), suy = real(0);
    real sneq[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) sneq[q] = real(0);
    for (int j = sub; j < a.MR; j += kNodeLanes) {          // same slot order per lane as the CSR ring
        const real w = j == sub ? w_first : a.ring_w[(size_t)node * a.MR + j];
        if (w != real(0)) {
            const int32_t c = j == sub ? c_first : a.ring_cell[(size_t)node * a.MR + j];
            real f[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) f[q] = a.pdf[pdf_off<Q>(a.lay, a.Npad, (int64_t)c, q)];
            node_accumulate<real, Q>(a.P, f, w, sw, srho, sux, suy, sneq);
        }
    }
    sw = group_sum(sw); srho = group_sum(srho); sux = group_sum(sux); suy = group_sum(suy);
#pragma unroll
    for (int q = 0; q < Q; ++q) sneq[q] = group_sum(sneq[q]);
    real pdf_n[Q];
    node_finish<real, Q>(a.P, type, sw, srho, sux, suy, sneq, rho_n, ux_n, uy_n, pdf_n);
    if (valid && sub == 0) {
        if (type == 1) a.nrho[node] = rho_n;
        if (type == 2) { a.nvel[node] = ux_n; a.nvel[a.NTpad + node] = uy_n; }
#pragma unroll
        for (int q = 0; q < Q; ++q) a.npdf[q * a.NTpad + node] = pdf_n[q];
    }
}

// ------------------------------------------------------------------------------------------------
// staged path: S1+S2, S4, S5 as separate kernels over the reference's data model.
// ------------------------------------------------------------------------------------------------
template <typename real, int Q>
__global__ void __launch_bounds__(256) k_s_moments(const Params<real> P, const real* __restrict__ pdf, int lay,
                                                  const int32_t* __restrict__ ipos, int64_t Npad,
                                                  real* __restrict__ rho, real* __restrict__ ux,
                                                  real* __restrict__ uy, real* __restrict__ pdf_eq) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= Npad || ipos[c] < 0) return;
    real f[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q] = pdf[pdf_off<Q>(lay, Npad, c, q)];
    real r, x, y;
    moments<real, Q>(f, r, x, y);
    rho[c] = r; ux[c] = x; uy[c] = y;
    const Equilibrium<real, Q> E(r, x, y, P);
#pragma unroll
    for (int q = 0; q < Q; ++q) pdf_eq[pdf_off<Q>(lay, Npad, c, q)] = E.value(q, P);
}

template <typename real>
struct FaceArgs {
    Params<real> P;
    const real* __restrict__ pdf;
    int lay; int64_t Npad;
    const int32_t* __restrict__ fcell;     // [F*2] positions, -1 ghost
    const int32_t* __restrict__ fnode;     // [F*2] tracked node ids (ghost faces only)
    const real* __restrict__ fdist;        // [F*2]
    const real* __restrict__ fn;           // [F*2]
    const real* __restrict__ fL;           // [F]
    const real* __restrict__ npdf;
    int64_t NTpad, F;
    int64_t last_pos;                      // position of original cell N-1 (python -1 indexing)
    real* __restrict__ flux;               // [F*Q] reference layout
};

template <typename real, int Q, int SCHEME>
__global__ void __launch_bounds__(256) k_s_faces(const FaceArgs<real> a) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= a.F) return;
    const int32_t s0 = a.fcell[2 * j], s1 = a.fcell[2 * j + 1];
    const real d0 = a.fdist[2 * j], d1 = a.fdist[2 * j + 1];
    const real nx = a.fn[2 * j], ny = a.fn[2 * j + 1], L = a.fL[j];
    const int64_t c0 = s0 < 0 ? a.last_pos : (int64_t)s0, c1 = s1 < 0 ? a.last_pos : (int64_t)s1;
    real f0[Q], f1[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { f0[q] = a.pdf[pdf_off<Q>(a.lay, a.Npad, c0, q)]; f1[q] = a.pdf[pdf_off<Q>(a.lay, a.Npad, c1, q)]; }
    if (s0 < 0 || s1 < 0) {
        const int32_t na = a.fnode[2 * j], nb = a.fnode[2 * j + 1];
        real g0[Q], g1[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const real g = (na >= 0 && nb >= 0) ? (a.npdf[q * a.NTpad + na] + a.npdf[q * a.NTpad + nb]) / real(2) : real(0);
            g0[q] = g + (g - f1[q]) * (d0 / d1);     // ghost in slot 0, known = slot 1
            g1[q] = g + (g - f0[q]) * (d1 / d0);     // ghost in slot 1, known = slot 0
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (s0 < 0) f0[q] = g0[q];
            if (s1 < 0) f1[q] = g1[q];
        }
    }
    real* out = a.flux + j * Q;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const real varpi = ksi_dot<real>(q, nx, ny, nx + ny, nx - ny);
        real fs;
        if (SCHEME == 0) fs = (varpi >= real(0)) ? f0[q] : f1[q];
        else {
            const real dd = d0 + d1;
            fs = f0[q] + (f1[q] - f0[q]) * (d0 / dd - (varpi * a.P.dt) / (real(2) * dd));
        }
        out[q] = fs * varpi * L;
    }
}

template <typename real, int Q, int K>
__global__ void __launch_bounds__(256) k_s_cells(const Params<real> P, int lay, const real* __restrict__ pdf,
                                                const real* __restrict__ pdf_eq, const real* __restrict__ flux,
                                                const int32_t* __restrict__ cface, const int32_t* __restrict__ csign,
                                                const int32_t* __restrict__ ipos, int64_t Npad, int64_t No,
                                                const real* __restrict__ inv_area, real* __restrict__ pdf_out) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= Npad) return;
    const int32_t o = ipos[c];
    if (o < 0 || o >= No) return;
    real fl[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) fl[q] = real(0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int64_t j = cface[(size_t)k * Npad + c];
        const real s = real(csign[(size_t)k * Npad + c]);
#pragma unroll
        for (int q = 0; q < Q; ++q) fl[q] += flux[j * Q + q] * s;
    }
    const real ia = inv_area ? inv_area[c] : real(1);        // optional physically consistent mode; NULL = reference
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const size_t ix = pdf_off<Q>(lay, Npad, c, q);
        const real f = pdf[ix];
        pdf_out[ix] = f + P.dt * (P.inv_tau * (pdf_eq[ix] - f) - fl[q] * ia);
    }
}

// ------------------------------------------------------------------------------------------------
// API-boundary layout conversion
// ------------------------------------------------------------------------------------------------
template <typename real, int Q>
__global__ void k_export_cells(const real* __restrict__ pdf, int lay, int64_t Npad, const int32_t* __restrict__ pos, int64_t N,
                               real* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t c = pos[i];
#pragma unroll
    for (int q = 0; q < Q; ++q) out[i * Q + q] = pdf[pdf_off<Q>(lay, Npad, c, q)];
}

// switch the population layout of a whole buffer (set_option(VARIANT) across layouts)
template <typename real, int Q>
__global__ void k_relayout(const real* __restrict__ src, int lay_src, real* __restrict__ dst, int lay_dst, int64_t Npad) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= Npad) return;
#pragma unroll
    for (int q = 0; q < Q; ++q) dst[pdf_off<Q>(lay_dst, Npad, c, q)] = src[pdf_off<Q>(lay_src, Npad, c, q)];
}

template <typename real, int Q>
__global__ void k_import_cells(real* __restrict__ pdf, int lay, int64_t Npad, const int32_t* __restrict__ pos, int64_t N,
                               const real* __restrict__ in) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t c = pos[i];
#pragma unroll
    for (int q = 0; q < Q; ++q) pdf[pdf_off<Q>(lay, Npad, c, q)] = in[i * Q + q];
}

// rho / vel / pdf_eq of the given populations in reference layout (any output may be null)
template <typename real, int Q>
__global__ void k_export_moments(const Params<real> P, const real* __restrict__ pdf, int lay, int64_t Npad,
                                 const int32_t* __restrict__ pos, int64_t N, real* __restrict__ rho, real* __restrict__ vel,
                                 real* __restrict__ pdf_eq) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t c = pos[i];
    real f[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q] = pdf[pdf_off<Q>(lay, Npad, c, q)];
    real r, x, y;
    moments<real, Q>(f, r, x, y);
    if (rho) rho[i] = r;
    if (vel) { vel[2 * i] = x; vel[2 * i + 1] = y; }
    if (pdf_eq) {
        const Equilibrium<real, Q> E(r, x, y, P);
#pragma unroll
        for (int q = 0; q < Q; ++q) pdf_eq[i * Q + q] = E.value(q, P);
    }
}

// failure detection: count owned cells with a non-finite population (one atomic per warp)
template <typename real, int Q>
__global__ void k_count_nonfinite(const real* __restrict__ pdf, int lay, const int32_t* __restrict__ ipos, int64_t Npad, int64_t No,
                                  unsigned long long* __restrict__ count) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool bad = false;
    if (c < Npad) {
        const int32_t o = ipos[c];
        if (o >= 0 && o < No) {
#pragma unroll
            for (int q = 0; q < Q; ++q) bad |= !isfinite(pdf[pdf_off<Q>(lay, Npad, c, q)]);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

// halo exchange helpers: list[i] = position ; buf layout [count][Q]
template <typename real, int Q>
__global__ void k_pack(const real* __restrict__ pdf, int lay, int64_t Npad, const int32_t* __restrict__ list, int64_t n,
                       real* __restrict__ buf) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * Q) return;
    const int64_t i = t / Q; const int q = (int)(t % Q);
    buf[t] = pdf[pdf_off<Q>(lay, Npad, list[i], q)];
}

template <typename real, int Q>
__global__ void k_unpack(real* __restrict__ pdf, int lay, int64_t Npad, const int32_t* __restrict__ list, int64_t n,
                         const real* __restrict__ buf) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * Q) return;
    const int64_t i = t / Q; const int q = (int)(t % Q);
    pdf[pdf_off<Q>(lay, Npad, list[i], q)] = buf[t];
}

}  // namespace fvdbm
