// kernels.cuh -- sm_100a kernels of the FVDBM step (see DESIGN.md for the roofline of each).
//
//   k_nodes         S3  boundary nodes: one warp per active node, shuffle reduction over its ring
//   k_fused_direct  S1+S2+S4+S5 cell-centric, thread per cell, operands through L1/L2
//   k_fused_tma     same arithmetic; persistent CTAs, cp.async.bulk (TMA) + mbarrier ring of tiles
//   k_s_*           staged (reference-shaped) kernels S1/S2, S4, S5 -- general meshes + observables
//   k_export_* / k_import_* / k_pack / k_unpack   layout conversion at the API boundary
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
    const real* __restrict__ ccoef;    // cell layout (LAYOUT 0): per-side coefficients, tiled like the codes
    const int32_t* __restrict__ cface; // face layout (LAYOUT 1): side -> record index
    const real* __restrict__ fcoef;    //                          records [NF][NC], one per face
    GhostTables<real> G;               // boundary sides + tracked node populations
    int64_t cell_begin, cell_end;      // position range, multiples of the CTA tile
    int reverse;                       // 1: sweep tiles from the top (L2 reuse of last step's writes)
    const int32_t* __restrict__ list;  // optional explicit position list (direct kernel): cell = list[i]
    int64_t list_n;
};

// ---- build-time tuning knobs (A/B'd on B200, see profiles/) -----------------------------------
#ifndef FVDBM_DIRECT_MINCTAS
#define FVDBM_DIRECT_MINCTAS 0        // >0: __launch_bounds__(256, N) for the direct kernel
#endif
#ifndef FVDBM_PREFETCH_NBR
#define FVDBM_PREFETCH_NBR 0          // 1: prefetch.global.L1 the neighbour populations of sides 1..K-1 up front
#endif
#ifndef FVDBM_STREAM_HINTS
#define FVDBM_STREAM_HINTS 0          // 1: ld.global.cs for the never-reused side records, st.global.cs for stores
#endif
// measured on B200 (profiles/r1_experiment_occupancy_cachehints.txt): fp32 is best left to ptxas
// (48 regs, 5 CTAs/SM; forcing 6 or 8 CTAs spills and loses 2-18 %), fp64 gains 15 % from 3 CTAs/SM
// (96 -> 80 regs: 0.492 -> 0.426 ms at 10M cells).
#if FVDBM_DIRECT_MINCTAS > 0
#define FVDBM_DIRECT_BOUNDS __launch_bounds__(256, FVDBM_DIRECT_MINCTAS)
#else
#define FVDBM_DIRECT_BOUNDS __launch_bounds__(256, (sizeof(real) == 8 ? 3 : (Q == 9 ? 5 : 4)))
#endif

template <typename T>
__device__ __forceinline__ T ld_static(const T* p) {     // side codes / coefficients: streamed once per step
#if FVDBM_STREAM_HINTS
    return __ldcs(p);
#else
    return __ldg(p);
#endif
}
template <typename T>
__device__ __forceinline__ void st_result(T* p, T v) {   // new populations: not re-read during this step
#if FVDBM_STREAM_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}

// one face record (NC coefficients, NC*sizeof(real) bytes, naturally aligned) with the widest loads
template <typename real, int NC>
__device__ __forceinline__ void load_face_record(const real* __restrict__ base, int32_t rec, real* out);
template <> __device__ __forceinline__ void load_face_record<float, 4>(const float* __restrict__ base, int32_t rec, float* out) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(base) + rec);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <> __device__ __forceinline__ void load_face_record<float, 2>(const float* __restrict__ base, int32_t rec, float* out) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(base) + rec);
    out[0] = v.x; out[1] = v.y;
}
template <> __device__ __forceinline__ void load_face_record<double, 4>(const double* __restrict__ base, int32_t rec, double* out) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(base) + 2 * (size_t)rec);
    const double2 b = __ldg(reinterpret_cast<const double2*>(base) + 2 * (size_t)rec + 1);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}
template <> __device__ __forceinline__ void load_face_record<double, 2>(const double* __restrict__ base, int32_t rec, double* out) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(base) + rec);
    out[0] = a.x; out[1] = a.y;
}

// ------------------------------------------------------------------------------------------------
// V1: thread per cell, everything through L1/L2.
// ------------------------------------------------------------------------------------------------
template <typename real, int Q, int K, int SCHEME, int LAYOUT>
__global__ void FVDBM_DIRECT_BOUNDS k_fused_direct(const FusedArgs<real> a) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    const int64_t nblk = gridDim.x;
    const int64_t blk = a.reverse ? (nblk - 1 - blockIdx.x) : blockIdx.x;
    int64_t c = a.cell_begin + blk * blockDim.x + threadIdx.x;
    if (a.list != nullptr) {                       // thin, list-driven pass (temporal schedule: level-2 cells)
        const int64_t i = blk * blockDim.x + threadIdx.x;
        if (i >= a.list_n) return;
        c = a.list[i];
    } else if (c >= a.cell_end) return;
    const size_t tile = (size_t)(c >> 5);
    const int lane = (int)(c & 31);
    // Issue every independent streaming load (side codes, side coefficients, own populations) BEFORE the
    // first use of any of them: the ncu source view showed ~30 % of the stall samples on the hole test of
    // code[0], i.e. a full memory round trip spent before the other loads were even in flight.  An early
    // `return` would let ptxas sink the loads below it again, so padding positions are not skipped: they
    // run the (in-bounds, harmless) arithmetic on a neutral code and only their stores are suppressed.
    const int32_t* gc = a.ccode + tile * (K * kTW) + lane;
    int32_t code[K];
#pragma unroll
    for (int k = 0; k < K; ++k) code[k] = ld_static(gc + k * kTW);
    real coef[K * NC];
    if (LAYOUT == 0) {
        const real* gco = a.ccoef + tile * (K * NC * kTW) + lane;
#pragma unroll
        for (int i = 0; i < K * NC; ++i) coef[i] = ld_static(gco + i * kTW);
    }
    const real* gp = a.pdf_in + tile * (Q * kTW) + lane;
    real f[Q], out[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q] = __ldg(gp + q * kTW);
    const bool live = code[0] != kHole;
    if (!live) code[0] = 0;                        // neutral: interior side towards position 0
    if (LAYOUT != 0) {
        const int32_t* gf = a.cface + tile * (K * kTW) + lane;
#pragma unroll
        for (int k = 0; k < K; ++k) load_face_record<real, NC>(a.fcoef, ld_static(gf + k * kTW), coef + k * NC);
    }
#if FVDBM_PREFETCH_NBR
    // pull the neighbours' lines towards L1 while side 0 is being computed (no registers held)
#pragma unroll
    for (int k = 1; k < K; ++k)
        if (code[k] >= 0) {
            const real* pn = a.pdf_in + pdf_index<Q>((int64_t)(code[k] >> 2));
#pragma unroll
            for (int q = 1; q < Q; ++q) asm volatile("prefetch.global.L1 [%0];" ::"l"(pn + q * kTW));
        }
#endif
    const real* pin = a.pdf_in;
    auto load_nbr = [pin](int64_t nb, real* fn) {
        const real* pn = pin + pdf_index<Q>(nb);
#pragma unroll
        for (int q = 1; q < Q; ++q) fn[q] = __ldg(pn + q * kTW);
    };
    advance_cell<real, Q, K, SCHEME>(a.P, a.G, f, code, coef, load_nbr, out);
    real* go = a.pdf_out + tile * (Q * kTW) + lane;
    if (live) {
#pragma unroll
        for (int q = 0; q < Q; ++q) st_result(go + q * kTW, out[q]);
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

template <typename real, int Q, int K, int SCHEME, int LAYOUT>
__host__ __device__ constexpr size_t tma_stage_bytes(int tile_cells) {
    return LAYOUT == 0 ? (size_t)tile_cells * ((Q + K * (SCHEME == 0 ? 2 : 4)) * sizeof(real) + K * sizeof(int32_t))
                       : (size_t)tile_cells * (Q * sizeof(real) + 2 * K * sizeof(int32_t));
}

template <typename real, int Q, int K, int SCHEME, int LAYOUT>
__global__ void __launch_bounds__(512) k_fused_tma(const FusedArgs<real> a, const int stages) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    extern __shared__ __align__(128) unsigned char smem[];
    const int TC = blockDim.x;
    const int tid = threadIdx.x;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    const size_t pdf_bytes = (size_t)TC * Q * sizeof(real);
    // second region of a stage: per-side coefficients (cell layout) or per-side record ids (face layout)
    const size_t coef_bytes = LAYOUT == 0 ? (size_t)TC * K * NC * sizeof(real) : (size_t)TC * K * sizeof(int32_t);
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
        if (LAYOUT == 0) bulk_g2s(st + pdf_bytes, a.ccoef + mt * (K * NC * kTW), (uint32_t)coef_bytes, &full[s]);
        else bulk_g2s(st + pdf_bytes, a.cface + mt * (K * kTW), (uint32_t)coef_bytes, &full[s]);
        bulk_g2s(st + pdf_bytes + coef_bytes, a.ccode + mt * (K * kTW), (uint32_t)code_bytes, &full[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
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
        const int32_t* s_face = reinterpret_cast<const int32_t*>(st + pdf_bytes) + (size_t)mt_local * (K * kTW) + lane;
        const int32_t* s_code = reinterpret_cast<const int32_t*>(st + pdf_bytes + coef_bytes) + (size_t)mt_local * (K * kTW) + lane;
        int32_t code[K];
        code[0] = s_code[0];
        if (code[0] != kHole) {
#pragma unroll
            for (int k = 1; k < K; ++k) code[k] = s_code[k * kTW];
            real coef[K * NC];
            if (LAYOUT == 0) {
#pragma unroll
                for (int i = 0; i < K * NC; ++i) coef[i] = s_coef[i * kTW];
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) load_face_record<real, NC>(a.fcoef, s_face[k * kTW], coef + k * NC);
            }
            const real* sp = s_pdf + (size_t)mt_local * (Q * kTW) + lane;
            real f[Q], out[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) f[q] = sp[q * kTW];
            const real* pin = a.pdf_in;
            auto load_nbr = [pin, s_pdf, base, TC](int64_t nb, real* fn) {
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
            advance_cell<real, Q, K, SCHEME>(a.P, a.G, f, code, coef, load_nbr, out);
            real* go = a.pdf_out + ((size_t)(base >> 5) + mt_local) * (Q * kTW) + lane;
#pragma unroll
            for (int q = 0; q < Q; ++q) go[q * kTW] = out[q];
        }
        __syncthreads();      // stage s may be refilled by the next iteration's issue
    }
}

// ------------------------------------------------------------------------------------------------
// Temporal blocking: TWO iterations per pass.  One CTA per tile of 256 cells at level >= 2 (no ghost
// sides within two rings).  Phase 0 stages the time-t populations of the tile, its face-neighbour
// ring and that ring's ring in shared memory; phase 1 advances tile + ring 1 to t+1 in shared memory
// (ring 1 redundantly -- the neighbouring tiles compute the same bits); phase 2 advances the tile to
// t+2 and writes it out.  Per cell and iteration the arithmetic is exactly advance_cell(), so the
// result is bit-identical to two single steps, while DRAM sees one read + one write of the
// populations and one read of the side coefficients per TWO iterations.
// ------------------------------------------------------------------------------------------------
template <typename real>
struct Fused2Args {
    Params<real> P;
    const real* __restrict__ pdf_in;
    real* __restrict__ pdf_out;
    const real* __restrict__ ccoef;
    const int32_t* __restrict__ t2_off;
    const int32_t* __restrict__ t2_n1;
    const int32_t* __restrict__ t2_pos;
    const int64_t* __restrict__ t2_loff;
    const uint16_t* __restrict__ t2_lnbr;
    int s0_stride, s1_stride;              // shared-memory strides (entries) of the two staging arrays
};

template <typename real, int Q, int K, int SCHEME>
__global__ void __launch_bounds__(256, (sizeof(real) == 4 ? 4 : 2)) k_fused2(const Fused2Args<real> a) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    constexpr int T2 = 256;
    extern __shared__ __align__(16) unsigned char smem_raw2[];
    real* s0 = reinterpret_cast<real*>(smem_raw2);
    real* s1 = s0 + (size_t)Q * a.s0_stride;
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int64_t t0 = (int64_t)tile * T2;
    const int off = a.t2_off[tile];
    const int n12 = a.t2_off[tile + 1] - off, n1 = a.t2_n1[tile];
    const int n01 = T2 + n1, nent = T2 + n12;
    const int S0 = a.s0_stride, S1 = a.s1_stride;
    const GhostTables<real> G0{nullptr, nullptr, nullptr, nullptr, 0};

    // phase 0: stage populations at time t (own: coalesced; rings: gathers sorted by position)
    for (int e = tid; e < nent; e += T2) {
        const int64_t pc = e < T2 ? t0 + e : (int64_t)a.t2_pos[off + e - T2];
        const real* src = a.pdf_in + pdf_index<Q>(pc);
#pragma unroll
        for (int q = 0; q < Q; ++q) s0[q * S0 + e] = __ldg(src + q * kTW);
    }
    __syncthreads();

    const uint16_t* lbase = a.t2_lnbr + (size_t)a.t2_loff[tile] * K;
    int32_t code_own[K];
    real coef_own[K * NC];
    bool own_live = false;
    // phase 1: tile + ring 1 -> t+1 (shared memory to shared memory)
    for (int e = tid; e < n01; e += T2) {
        int32_t code[K];
#pragma unroll
        for (int k = 0; k < K; ++k) code[k] = (int32_t)lbase[(size_t)e * K + k];
        if (code[0] == 0xFFFF) continue;                       // padding position
        const int64_t pc = e < T2 ? t0 + e : (int64_t)a.t2_pos[off + e - T2];
        const real* gco = a.ccoef + (size_t)(pc >> 5) * (K * NC * kTW) + (pc & 31);
        real coef[K * NC], f[Q], out[Q];
#pragma unroll
        for (int i = 0; i < K * NC; ++i) coef[i] = __ldg(gco + i * kTW);
#pragma unroll
        for (int q = 0; q < Q; ++q) f[q] = s0[q * S0 + e];
        auto load_nbr = [s0, S0](int64_t nb, real* fn) {
#pragma unroll
            for (int q = 1; q < Q; ++q) fn[q] = s0[q * S0 + (int)nb];
        };
        advance_cell<real, Q, K, SCHEME>(a.P, G0, f, code, coef, load_nbr, out);
#pragma unroll
        for (int q = 0; q < Q; ++q) s1[q * S1 + e] = out[q];
        if (e < T2) {                                          // keep the own cell's side records for phase 2
            own_live = true;
#pragma unroll
            for (int k = 0; k < K; ++k) code_own[k] = code[k];
#pragma unroll
            for (int i = 0; i < K * NC; ++i) coef_own[i] = coef[i];
        }
    }
    __syncthreads();

    // phase 2: tile -> t+2
    if (own_live) {
        real f[Q], out[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) f[q] = s1[q * S1 + tid];
        auto load_nbr = [s1, S1](int64_t nb, real* fn) {
#pragma unroll
            for (int q = 1; q < Q; ++q) fn[q] = s1[q * S1 + (int)nb];
        };
        advance_cell<real, Q, K, SCHEME>(a.P, G0, f, code_own, coef_own, load_nbr, out);
        real* go = a.pdf_out + pdf_index<Q>(t0 + tid);
#pragma unroll
        for (int q = 0; q < Q; ++q) go[q * kTW] = out[q];
    }
}

// ------------------------------------------------------------------------------------------------
// S3 boundary nodes (src/containers.py:339-404): warp per active node.
// ------------------------------------------------------------------------------------------------
template <typename real>
struct NodeArgs {
    Params<real> P;
    const real* __restrict__ pdf;          // current populations (AoSoA)
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
__device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One active node evaluated by one warp: lanes walk the ring, shfl_xor tree reduction (every lane ends
// with the same sums), then rho / vel / populations of the node (identical on all lanes).
template <typename real, int Q>
__device__ __forceinline__ void warp_eval_node(const NodeArgs<real>& a, int node, int lane, real& rho_n, real& ux_n, real& uy_n,
                                               real* pdf_n) {
    // prescribed values / type are independent of the ring: issue their loads up front so they overlap
    // the ring gathers instead of adding a dependent memory round trip after the reduction
    const int type = a.tn_type[node];
    rho_n = a.nrho[node]; ux_n = a.nvel[node]; uy_n = a.nvel[a.NTpad + node];
    real sw = real(0), srho = real(0), sux = real(0), suy = real(0);
    real sneq[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) sneq[q] = real(0);
    for (int j = lane; j < a.MR; j += 32) {                 // lane j <-> ring slot j (same order as the CSR)
        const size_t i = (size_t)node * a.MR + j;
        const real w = a.ring_w[i];
        if (w != real(0)) {
            const real* p = a.pdf + pdf_index<Q>((int64_t)a.ring_cell[i]);
            real f[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) f[q] = p[q * kTW];
            node_accumulate<real, Q>(a.P, f, w, sw, srho, sux, suy, sneq);
        }
    }
    sw = warp_sum(sw); srho = warp_sum(srho); sux = warp_sum(sux); suy = warp_sum(suy);
#pragma unroll
    for (int q = 0; q < Q; ++q) sneq[q] = warp_sum(sneq[q]);
    node_finish<real, Q>(a.P, type, sw, srho, sux, suy, sneq, rho_n, ux_n, uy_n, pdf_n);
}

template <typename real, int Q>
__device__ __forceinline__ void store_node(const NodeArgs<real>& a, int node, real rho_n, real ux_n, real uy_n, const real* pdf_n) {
    const int type = a.tn_type[node];
    if (type == 1) a.nrho[node] = rho_n;
    if (type == 2) { a.nvel[node] = ux_n; a.nvel[a.NTpad + node] = uy_n; }
#pragma unroll
    for (int q = 0; q < Q; ++q) a.npdf[q * a.NTpad + node] = pdf_n[q];
}

template <typename real, int Q>
__global__ void __launch_bounds__(256) k_nodes(const NodeArgs<real> a) {
    const int node = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (node >= a.NA) return;
    real rho_n, ux_n, uy_n, pdf_n[Q];
    warp_eval_node<real, Q>(a, node, lane, rho_n, ux_n, uy_n, pdf_n);
    __syncwarp();
    if (lane == 0) store_node<real, Q>(a, node, rho_n, ux_n, uy_n, pdf_n);
}

// ------------------------------------------------------------------------------------------------
// Border kernel: one CTA per tile of BORDER_TILE (=256) border cells.  Phase A: the tile's warps
// evaluate the boundary nodes its ghost sides reference (S3) into shared memory -- a node shared by
// two tiles is evaluated by both, identically -- and publish them for observation; phase B: the
// tile's cells run the ordinary fused update taking ghost populations from shared memory.  This
// keeps ONE kernel on the per-step critical path [nodes -> border cells] while the interior cells
// run on the side stream.
// ------------------------------------------------------------------------------------------------
template <typename real>
struct BorderArgs {
    FusedArgs<real> F;
    NodeArgs<real> N;
    const int32_t* __restrict__ bt_off;
    const int32_t* __restrict__ bt_nodes;
    const int32_t* __restrict__ bf_la;
    const int32_t* __restrict__ bf_lb;
};

template <typename real, int Q, int K, int SCHEME>
__global__ void __launch_bounds__(256) k_border(const BorderArgs<real> b) {
    constexpr int NC = SCHEME == 0 ? 2 : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* s_node = reinterpret_cast<real*>(smem_raw);
    const FusedArgs<real>& a = b.F;
    const int tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = b.bt_off[tile], nn = b.bt_off[tile + 1] - n0;
    for (int i = warp; i < nn; i += 8) {
        const int t = b.bt_nodes[n0 + i];
        if (t < b.N.NA) {                                   // active: evaluate from its ring
            real rho_n, ux_n, uy_n, pdf_n[Q];
            warp_eval_node<real, Q>(b.N, t, lane, rho_n, ux_n, uy_n, pdf_n);
            if (lane == 0) {
                store_node<real, Q>(b.N, t, rho_n, ux_n, uy_n, pdf_n);
#pragma unroll
                for (int q = 0; q < Q; ++q) s_node[i * Q + q] = pdf_n[q];
            }
        } else if (lane < Q) {                              // type-0 boundary node: keeps its stored PDFs
            s_node[i * Q + lane] = b.N.npdf[lane * b.N.NTpad + t];
        }
    }
    __syncthreads();
    const int64_t c = a.cell_begin + (int64_t)tile * blockDim.x + threadIdx.x;
    if (c >= a.cell_end) return;
    const size_t mt = (size_t)(c >> 5);
    const int32_t* gc = a.ccode + mt * (K * kTW) + lane;
    int32_t code[K];
    code[0] = gc[0];
    if (code[0] == kHole) return;
#pragma unroll
    for (int k = 1; k < K; ++k) code[k] = gc[k * kTW];
    real coef[K * NC];
    if (a.cface == nullptr) {
        const real* gco = a.ccoef + mt * (K * NC * kTW) + lane;
#pragma unroll
        for (int i = 0; i < K * NC; ++i) coef[i] = gco[i * kTW];
    } else {
        const int32_t* gf = a.cface + mt * (K * kTW) + lane;
#pragma unroll
        for (int k = 0; k < K; ++k) load_face_record<real, NC>(a.fcoef, gf[k * kTW], coef + k * NC);
    }
    const real* pin = a.pdf_in;
    const real* gp = pin + mt * (Q * kTW) + lane;
    real f[Q], out[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q] = gp[q * kTW];
    auto load_nbr = [pin](int64_t nb, real* fn) {
        const real* pn = pin + pdf_index<Q>(nb);
#pragma unroll
        for (int q = 1; q < Q; ++q) fn[q] = __ldg(pn + q * kTW);
    };
    GhostTables<real> G = a.G;
    G.snode = s_node; G.bf_la = b.bf_la; G.bf_lb = b.bf_lb;
    advance_cell<real, Q, K, SCHEME>(a.P, G, f, code, coef, load_nbr, out);
    real* go = a.pdf_out + mt * (Q * kTW) + lane;
#pragma unroll
    for (int q = 0; q < Q; ++q) go[q * kTW] = out[q];
}

// ------------------------------------------------------------------------------------------------
// staged path: S1+S2, S4, S5 as separate kernels over the reference's data model.
// ------------------------------------------------------------------------------------------------
template <typename real, int Q>
__global__ void __launch_bounds__(256) k_s_moments(const Params<real> P, const real* __restrict__ pdf,
                                                  const int32_t* __restrict__ ipos, int64_t Npad,
                                                  real* __restrict__ rho, real* __restrict__ ux,
                                                  real* __restrict__ uy, real* __restrict__ pdf_eq) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= Npad || ipos[c] < 0) return;
    const real* p = pdf + pdf_index<Q>(c);
    real f[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q] = p[q * kTW];
    real r, x, y;
    moments<real, Q>(f, r, x, y);
    rho[c] = r; ux[c] = x; uy[c] = y;
    const real uu = x * x + y * y;
    real* e = pdf_eq + pdf_index<Q>(c);
#pragma unroll
    for (int q = 0; q < Q; ++q) e[q * kTW] = feq<real, Q>(q, r, x, y, uu, P);
}

template <typename real>
struct FaceArgs {
    Params<real> P;
    const real* __restrict__ pdf;
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
    const real* p0 = a.pdf + pdf_index<Q>(s0 < 0 ? a.last_pos : (int64_t)s0);
    const real* p1 = a.pdf + pdf_index<Q>(s1 < 0 ? a.last_pos : (int64_t)s1);
    real f0[Q], f1[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { f0[q] = p0[q * kTW]; f1[q] = p1[q * kTW]; }
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
        const real varpi = ksi_dot<real>(q, nx, ny);
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
__global__ void __launch_bounds__(256) k_s_cells(const Params<real> P, const real* __restrict__ pdf,
                                                const real* __restrict__ pdf_eq, const real* __restrict__ flux,
                                                const int32_t* __restrict__ cface, const int32_t* __restrict__ csign,
                                                const int32_t* __restrict__ ipos, int64_t Npad, int64_t No,
                                                real* __restrict__ pdf_out) {
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
    const size_t ix = pdf_index<Q>(c);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const real f = pdf[ix + q * kTW];
        pdf_out[ix + q * kTW] = f + P.dt * (P.inv_tau * (pdf_eq[ix + q * kTW] - f) - fl[q]);
    }
}

// ------------------------------------------------------------------------------------------------
// API-boundary layout conversion
// ------------------------------------------------------------------------------------------------
template <typename real, int Q>
__global__ void k_export_cells(const real* __restrict__ pdf, const int32_t* __restrict__ pos, int64_t N,
                               real* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const real* p = pdf + pdf_index<Q>(pos[i]);
#pragma unroll
    for (int q = 0; q < Q; ++q) out[i * Q + q] = p[q * kTW];
}

template <typename real, int Q>
__global__ void k_import_cells(real* __restrict__ pdf, const int32_t* __restrict__ pos, int64_t N,
                               const real* __restrict__ in) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    real* p = pdf + pdf_index<Q>(pos[i]);
#pragma unroll
    for (int q = 0; q < Q; ++q) p[q * kTW] = in[i * Q + q];
}

// rho / vel / pdf_eq of the given populations in reference layout (any output may be null)
template <typename real, int Q>
__global__ void k_export_moments(const Params<real> P, const real* __restrict__ pdf, const int32_t* __restrict__ pos,
                                 int64_t N, real* __restrict__ rho, real* __restrict__ vel, real* __restrict__ pdf_eq) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const real* p = pdf + pdf_index<Q>(pos[i]);
    real f[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q] = p[q * kTW];
    real r, x, y;
    moments<real, Q>(f, r, x, y);
    if (rho) rho[i] = r;
    if (vel) { vel[2 * i] = x; vel[2 * i + 1] = y; }
    if (pdf_eq) {
        const real uu = x * x + y * y;
#pragma unroll
        for (int q = 0; q < Q; ++q) pdf_eq[i * Q + q] = feq<real, Q>(q, r, x, y, uu, P);
    }
}

// failure detection: count owned cells with a non-finite population (one atomic per warp)
template <typename real, int Q>
__global__ void k_count_nonfinite(const real* __restrict__ pdf, const int32_t* __restrict__ ipos, int64_t Npad, int64_t No,
                                  unsigned long long* __restrict__ count) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool bad = false;
    if (c < Npad) {
        const int32_t o = ipos[c];
        if (o >= 0 && o < No) {
            const real* p = pdf + pdf_index<Q>(c);
#pragma unroll
            for (int q = 0; q < Q; ++q) bad |= !isfinite(p[q * kTW]);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

// halo exchange helpers: list[i] = position ; buf layout [count][Q]
template <typename real, int Q>
__global__ void k_pack(const real* __restrict__ pdf, const int32_t* __restrict__ list, int64_t n, real* __restrict__ buf) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * Q) return;
    const int64_t i = t / Q; const int q = (int)(t % Q);
    buf[t] = pdf[pdf_index<Q>(list[i]) + q * kTW];
}

template <typename real, int Q>
__global__ void k_unpack(real* __restrict__ pdf, const int32_t* __restrict__ list, int64_t n, const real* __restrict__ buf) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * Q) return;
    const int64_t i = t / Q; const int q = (int)(t % Q);
    pdf[pdf_index<Q>(list[i]) + q * kTW] = buf[t];
}

}  // namespace fvdbm
