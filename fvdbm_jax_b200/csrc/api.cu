// api.cu -- C ABI of libfvdbm_b200.so (include/fvdbm_b200.h): handle, dispatch, transfers.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include <map>
#include <initializer_list>

#include <dlfcn.h>
#if __has_include(<nccl.h>)
#include <nccl.h>
#else   // minimal declarations (ABI-stable subset) when the header is absent at build time
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
#endif

#include "../../include/fvdbm_b200.h"
#include "plan.hpp"
#include "mesh.hpp"
#include "kernels.cuh"

using namespace fvdbm;

namespace {

thread_local std::string g_create_error;

// NCCL entry points resolved at run time (torch's bundled libnccl.so.2 if already loaded, else the
// system one), so libfvdbm_b200.so itself has no link-time NCCL dependency.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        const char* names[] = {getenv("FVDBM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { error = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define NCCL_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); if (!field) { error = "missing symbol " name; lib = nullptr; return false; }
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId") NCCL_SYM(CommInitRank, "ncclCommInitRank") NCCL_SYM(CommDestroy, "ncclCommDestroy")
        NCCL_SYM(Send, "ncclSend") NCCL_SYM(Recv, "ncclRecv") NCCL_SYM(GroupStart, "ncclGroupStart") NCCL_SYM(GroupEnd, "ncclGroupEnd")
        NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;

struct Engine {
    virtual ~Engine() {}
    std::string err;
    virtual int step(int n) = 0;
    virtual int step_timed(int n, float* ms) = 0;
    virtual int step_phase(int phase) = 0;
    virtual int sync() = 0;
    virtual int get(int field, void* dst, size_t bytes) = 0;
    virtual int set(int field, const void* src, size_t bytes) = 0;
    virtual int set_async(int field, const void* src, size_t bytes) = 0;
    virtual int get_async(int field, void* dst, size_t bytes, int64_t* ticket) = 0;
    virtual int wait_ticket(int64_t ticket) = 0;
    virtual int set_params(double tau, double dt) = 0;
    virtual int set_option(int opt, int64_t v) = 0;
    virtual int info(int key, int64_t* v) const = 0;
    virtual int check_finite(int64_t* bad) = 0;
    virtual int halo_set_lists(const int32_t* s, int64_t ns, const int32_t* r, int64_t nr) = 0;
    virtual int halo_pack(void* buf) = 0;
    virtual int halo_unpack(const void* buf) = 0;
    virtual void* stream_handle() = 0;
    virtual int comm_init(int nranks, int rank, const void* id) = 0;
    virtual int halo_set_peers(const int32_t* sp, const int64_t* sc, int ns, const int32_t* rp, const int64_t* rc, int nr) = 0;
};

#define CU_TRY(expr)                                                                          \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            err = std::string(#expr) + ": " + cudaGetErrorString(e_);                         \
            return FVDBM_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc(&p, count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    cudaError_t upload(const T* v, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, v, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
    size_t bytes() const { return n * sizeof(T); }
};

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// launch with (pdl = true) or without the programmatic-stream-serialization attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl,
                            Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename real, int Q, int K, int SCHEME>
struct EngineT final : Engine {
    Plan<real> plan;
    Params<real> P{};
    int device = 0;
    cudaStream_t stream = nullptr;      // main stream (highest priority): nodes, border cells, transfers
    cudaStream_t stream2 = nullptr;     // side stream (lowest priority): interior cells, forked/joined per step
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int overlap = 1;
    int mode = FVDBM_MODE_FUSED;
    int variant = FVDBM_VARIANT_DIRECT;   // replaced by default_variant() at init: fp32 D2Q9 REC, fp64 DIRECT; TMA is opt-in
    int tile_cells = 256, stages = 3, graph_steps = 0, ctas_per_sm = 0, reverse_sweep = 0;
    int pdl = 0;                         // programmatic dependent launch chain (default: on below 1.5M cells)
    int lay = 0;                         // population layout of pdf[0..1] (core.cuh: 0 tiled AoSoA, 1 records); follows the variant
    int prefetch_dist = 296;             // CTAs of L2 look-ahead (0.4 of a resident wave); measured on B200: burst 0.231 -> 0.194 ms per
                                         // 10M-cell iteration, sustained +2-3 % (profiles/r2_ab_pair_kernel.jsonl)
    int num_sms = 148;
    int cur = 0;
    int64_t steps = 0, launches = 0;
    bool phase0_done = false;

    DevBuf<real> pdf[2];                 // ping-pong: pdf[prev] *is* the lagged state of the reference's observables
    int prev = 1;                        // buffer holding the populations before the last iteration
    int nxt() const { return cur ^ 1; }
    DevBuf<int32_t> ccode, bf_na, bf_nb, pos, ipos, ring_off, ring_cell, tn_type;
    DevBuf<real> ccoef, bf_ratio, ring_w, npdf, nrho, nvel, s_inv_area;
    DevBuf<int32_t> s_cface, s_csign, s_fcell, s_fnode;
    DevBuf<real> s_fdist, s_fn, s_fL;
    DevBuf<real> s_rho, s_ux, s_uy, s_feq, s_flux;     // staged dynamics (lazy)
    // host <-> device staging in reference layout.  Uploads land in `inbox` on their own copy stream, so the
    // PCIe transfer overlaps whatever the main stream is still computing; the import kernel is ordered behind
    // it on the main stream.  Downloads: an export kernel on the main stream fills one of two `outbox` slots,
    // the D2H copy runs on a second copy stream (full-duplex PCIe, overlaps the next iterations).
    DevBuf<real> inbox;
    cudaStream_t xin = nullptr, xout = nullptr;
    cudaEvent_t inbox_filled = nullptr, inbox_consumed = nullptr;
    bool inbox_used = false;
    struct Outbox {
        DevBuf<real> buf;
        cudaEvent_t ready = nullptr, drained = nullptr;
        bool used = false;
        int64_t moments_steps = -1, moments_rows = 0;   // slot holds [rho | vel] of the populations before step `moments_steps`
    } outbox[2];
    static constexpr int kTickets = 32;
    cudaEvent_t ticket_ev[kTickets] = {};
    int64_t tickets = 0;
    DevBuf<int32_t> halo_send, halo_recv;
    DevBuf<unsigned long long> counter;
    // native exchange (optional)
    ncclComm_t comm = nullptr;
    std::vector<int> send_peers, recv_peers;
    std::vector<int64_t> send_counts, recv_counts;
    DevBuf<real> sendbuf, recvbuf;
    bool native_exchange() const { return comm != nullptr && (!send_peers.empty() || !recv_peers.empty()); }
    std::map<int, cudaGraphExec_t> graphs;             // key: starting `cur`

    ~EngineT() override {
        cudaSetDevice(device);
        drop_graphs();
        if (comm && g_nccl.CommDestroy) { cudaStreamSynchronize(stream); g_nccl.CommDestroy(comm); }
        if (xin) { cudaStreamSynchronize(xin); cudaStreamDestroy(xin); }
        if (xout) { cudaStreamSynchronize(xout); cudaStreamDestroy(xout); }
        cudaEvent_t evs[] = {inbox_filled, inbox_consumed, outbox[0].ready, outbox[0].drained, outbox[1].ready, outbox[1].drained};
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : ticket_ev) if (e) cudaEventDestroy(e);
        if (stream2) { cudaStreamSynchronize(stream2); cudaStreamDestroy(stream2); }
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
    }
    void drop_graphs() {
        for (auto& g : graphs) cudaGraphExecDestroy(g.second);
        graphs.clear();
    }

    void fill_params(const fvdbm_desc& d) {
        for (int q = 0; q < 16; ++q) P.w[q] = (real)d.lat_w[q];
        const real cs2 = (real)d.cs2, t4 = (real)d.two_cs4, t2 = (real)d.two_cs2, t6 = (real)d.two_cs6;
        P.inv_cs2 = real(1) / cs2;
        P.inv_2cs4 = real(1) / t4;
        P.inv_2cs2 = real(1) / t2;
        P.inv_2cs6 = (Q == 13) ? real(1) / t6 : real(0);
        P.three_inv_2cs4 = real(3) / t4;
        P.inv_tau = (real)(1.0 / d.tau);
        P.dt = (real)d.delta_t;
    }

    int init(const fvdbm_desc& d) {
        if (!plan.build(d)) { err = plan.error; return FVDBM_ERR_ARG; }
        device = d.device_id;
        int ndev = 0;
        CU_TRY(cudaGetDeviceCount(&ndev));
        if (device < 0 || device >= ndev) { err = "device_id out of range"; return FVDBM_ERR_ARG; }
        CU_TRY(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        num_sms = prop.multiProcessorCount;
        int prio_lo = 0, prio_hi = 0;
        CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CU_TRY(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_hi));
        CU_TRY(cudaStreamCreateWithPriority(&stream2, cudaStreamNonBlocking, prio_lo));
        CU_TRY(cudaStreamCreateWithFlags(&xin, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&xout, cudaStreamNonBlocking));
        cudaEvent_t* evs[] = {&inbox_filled, &inbox_consumed, &outbox[0].ready, &outbox[0].drained, &outbox[1].ready, &outbox[1].drained};
        for (cudaEvent_t* e : evs) CU_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (cudaEvent_t& e : ticket_ev) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        if (const char* e = getenv("FVDBM_OVERLAP")) overlap = atoi(e);
        fill_params(d);
        mode = d.mode == FVDBM_MODE_STAGED ? FVDBM_MODE_STAGED : FVDBM_MODE_FUSED;
        if (mode == FVDBM_MODE_FUSED && !plan.fused_ok) {
            if (d.mode == FVDBM_MODE_FUSED) { err = "fused mode unavailable: " + plan.why_not; return FVDBM_ERR_ARG; }
            mode = FVDBM_MODE_STAGED;
        }
        if (mode == FVDBM_MODE_STAGED && plan.No < plan.N) { err = "staged mode does not support halo cells"; return FVDBM_ERR_ARG; }
        const size_t npdf_elems = (size_t)(plan.Npad / TW) * Q * TW;
        CU_TRY(pdf[0].alloc(npdf_elems));
        CU_TRY(pdf[1].alloc(npdf_elems));
        CU_TRY(cudaMemsetAsync(pdf[0].p, 0, pdf[0].bytes(), stream));
        CU_TRY(cudaMemsetAsync(pdf[1].p, 0, pdf[1].bytes(), stream));
        CU_TRY(pos.upload(plan.pos, stream));
        CU_TRY(ipos.upload(plan.ipos, stream));
        if (plan.fused_ok) {
            CU_TRY(ccode.upload(plan.ccode, stream));
            CU_TRY(ccoef.upload(plan.ccoef, stream));
            CU_TRY(bf_na.upload(plan.bf_na, stream));
            CU_TRY(bf_nb.upload(plan.bf_nb, stream));
            CU_TRY(bf_ratio.upload(plan.bf_ratio, stream));
        }
        CU_TRY(s_inv_area.upload(plan.s_inv_area, stream));
        CU_TRY(ring_cell.upload(plan.ring_fcell, stream));     // fixed-width ring table [NA][MR]
        CU_TRY(ring_w.upload(plan.ring_fw, stream));
        CU_TRY(tn_type.upload(plan.tn_type, stream));
        CU_TRY(npdf.upload(plan.tn_pdf, stream));
        CU_TRY(nrho.upload(plan.tn_rho, stream));
        CU_TRY(nvel.upload(plan.tn_vel, stream));
        CU_TRY(s_cface.upload(plan.s_cface, stream));
        CU_TRY(s_csign.upload(plan.s_csign, stream));
        CU_TRY(s_fcell.upload(plan.s_fcell, stream));
        CU_TRY(s_fnode.upload(plan.s_fnode, stream));
        CU_TRY(s_fdist.upload(static_cast<const real*>(d.face_dists), (size_t)plan.F * 2, stream));
        CU_TRY(s_fn.upload(static_cast<const real*>(d.face_n), (size_t)plan.F * 2, stream));
        CU_TRY(s_fL.upload(static_cast<const real*>(d.face_L), (size_t)plan.F, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        // host staging vectors are no longer needed
        plan.ccode = {}; plan.ccoef = {}; plan.s_cface = {}; plan.s_csign = {}; plan.s_fcell = {}; plan.s_fnode = {};
        plan.ring_cell = {}; plan.ring_w = {}; plan.ring_fcell = {}; plan.ring_fw = {}; plan.tn_pdf = {}; plan.s_inv_area = {};
        int rc = set(FVDBM_CELL_PDF, d.cell_pdf, (size_t)plan.N * Q * sizeof(real));
        if (rc) return rc;
        // opt-in shared memory for the TMA kernel
        CU_TRY(cudaFuncSetAttribute(k_fused_tma<real, Q, K, SCHEME>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)prop.sharedMemPerBlockOptin));
        max_smem = prop.sharedMemPerBlockOptin;
        variant = default_variant();
        // debugging / A-B overrides of the defaults (documented in include/fvdbm_b200.h; the same knobs are
        // reachable through fvdbm_set_option): FVDBM_VARIANT, FVDBM_TILE_CELLS, FVDBM_STAGES,
        // FVDBM_GRAPH_STEPS, FVDBM_CTAS_PER_SM, FVDBM_REVERSE_SWEEP, FVDBM_OVERLAP
        if (const char* e = getenv("FVDBM_VARIANT")) variant = atoi(e);
        if (const char* e = getenv("FVDBM_TILE_CELLS")) tile_cells = atoi(e);
        if (const char* e = getenv("FVDBM_STAGES")) stages = atoi(e);
        // latency-bound meshes: batch iterations in CUDA graphs by default (20k cells: 13.2 -> 7.2 us/step)
        // PDL chain + graphs win up to ~1M cells (1.0M: 25.1 vs 26.6 us per iteration), the two-stream schedule from 2M
        // (43.0 vs 44.9 us); profiles/r2_schedule_sweep_rec256.jsonl
        if (plan.No < 1500000) { graph_steps = 50; pdl = 1; }
        if (const char* e = getenv("FVDBM_PDL")) pdl = atoi(e) ? 1 : 0;
        if (const char* e = getenv("FVDBM_GRAPH_STEPS")) graph_steps = atoi(e);
        if (const char* e = getenv("FVDBM_CTAS_PER_SM")) ctas_per_sm = atoi(e);
        if (const char* e = getenv("FVDBM_REVERSE_SWEEP")) reverse_sweep = atoi(e);
        if (const char* e = getenv("FVDBM_PREFETCH_DIST")) prefetch_dist = atoi(e);
        if (variant == FVDBM_VARIANT_AUTO) variant = default_variant();
        return sanitize_options();
    }
    size_t max_smem = 0;
    int occ_cache = 0;
    // fp32 D2Q9: the record-layout kernel (one 256-bit access per record, FFMA2 over population pairs) at every size: with
    // 256-bit accesses it is as fast as or faster than the thread-per-cell AoSoA kernel from 10k to 10M cells (20k: 4.88 vs
    // 5.05 us, cylinder 194k: 8.50 vs 9.11, 500k: 12.3 vs 13.2, 1M: 24.6 vs 26.7, porous 2M: 49.2 vs 50.9, 4M: 81.3 vs 81.0,
    // 10M sustained: 0.204 vs 0.228 ms; only 50k-100k squares are 1-2 % behind) -- profiles/r2_variant_crossover.jsonl,
    // r2_small_configs_rec256.jsonl, r2_ab_record_kernel.jsonl.  fp32 D2Q13: thread-per-cell over records up to 64k cells
    // (5.57 vs 5.82 us), the packed two-cells-per-thread kernel from 4M.  fp64: thread-per-cell AoSoA (the fp64 record is
    // 64 B and its strided 128-bit accesses lose: 0.518 vs 0.359 ms).
    int default_variant() const {
        if (sizeof(real) == 4 && mode == FVDBM_MODE_FUSED) {
            const bool big = plan.No >= (int64_t(1) << 22), small = plan.No <= (int64_t(1) << 16);
            if (Q == 9 || small) return FVDBM_VARIANT_REC;
            if (big) return FVDBM_VARIANT_PAIR;
        }
        return FVDBM_VARIANT_DIRECT;
    }

    size_t stage_bytes(int tc) const { return tma_stage_bytes<real, Q, K, SCHEME>(tc); }

    int sanitize_options() {
        if (variant != FVDBM_VARIANT_DIRECT && variant != FVDBM_VARIANT_TMA && variant != FVDBM_VARIANT_PAIR && variant != FVDBM_VARIANT_REC) { err = "unknown variant"; return FVDBM_ERR_ARG; }
        if (variant == FVDBM_VARIANT_REC && mode != FVDBM_MODE_FUSED) { err = "the record-layout kernel needs the fused mode"; return FVDBM_ERR_ARG; }
        if (variant == FVDBM_VARIANT_PAIR && sizeof(real) != 4) { err = "the packed two-cells-per-thread kernel exists for fp32 only"; return FVDBM_ERR_ARG; }
        if (tile_cells != 128 && tile_cells != 256 && tile_cells != 512) { err = "tile_cells must be 128, 256 or 512"; return FVDBM_ERR_ARG; }
        if (stages < 2 || stages > 8) { err = "stages must be in 2..8"; return FVDBM_ERR_ARG; }
        while (stages > 2 && kTmaHeader + stages * stage_bytes(tile_cells) > max_smem) --stages;
        while (tile_cells > 128 && kTmaHeader + stages * stage_bytes(tile_cells) > max_smem) tile_cells /= 2;
        if (graph_steps < 0) graph_steps = 0;
        if (graph_steps & 1) ++graph_steps;
        return set_layout(variant == FVDBM_VARIANT_REC ? 1 : 0);
    }

    // ---------------------------------------------------------------- launches
    FusedArgs<real> fused_args(int64_t begin, int64_t end) const {
        FusedArgs<real> a;
        a.P = P;
        a.pdf_in = pdf[cur].p; a.pdf_out = pdf[nxt()].p;
        a.ccode = ccode.p; a.ccoef = ccoef.p;
        a.G.bf_na = bf_na.p; a.G.bf_nb = bf_nb.p; a.G.bf_ratio = bf_ratio.p;
        a.G.npdf = npdf.p; a.G.NTpad = plan.NTpad;
        a.cell_begin = begin; a.cell_end = end;
        a.reverse = (reverse_sweep && cur == 1) ? 1 : 0;
        a.Npad = plan.Npad;
        a.prefetch_dist = prefetch_dist;
        return a;
    }

    NodeArgs<real> node_args(int64_t count) const {
        NodeArgs<real> a;
        a.P = P; a.pdf = pdf[cur].p; a.lay = lay; a.Npad = plan.Npad;
        a.ring_cell = ring_cell.p; a.ring_w = ring_w.p; a.MR = (int)plan.MR; a.tn_type = tn_type.p;
        a.npdf = npdf.p; a.nrho = nrho.p; a.nvel = nvel.p; a.NTpad = plan.NTpad; a.NA = (int)count;
        return a;
    }
    int launch_nodes() {
        const int64_t count = plan.NA;
        if (count == 0) return FVDBM_OK;
        NodeArgs<real> a = node_args(count);
        CU_TRY(launch_k(k_nodes<real, Q>, blocks_for(count * kNodeLanes, 256), 256, 0, stream, pdl_chain(), a));
        ++launches;
        CU_TRY(cudaGetLastError());
        return FVDBM_OK;
    }

    int launch_fused(int64_t begin, int64_t end, cudaStream_t st) {
        if (end <= begin) return FVDBM_OK;
        FusedArgs<real> a = fused_args(begin, end);
        if (variant == FVDBM_VARIANT_REC) {
            launch_rec(a, end - begin, st);
        } else if (variant == FVDBM_VARIANT_PAIR) {
            launch_pair(a, end - begin, st);
        } else if (variant == FVDBM_VARIANT_DIRECT) {
            CU_TRY(launch_k(k_fused_direct<real, Q, K, SCHEME, 0>, blocks_for(end - begin, 256), 256, 0, st, pdl_chain(), a));
        } else {
            const size_t smem = kTmaHeader + (size_t)stages * stage_bytes(tile_cells);
            int per_sm = ctas_per_sm;
            if (per_sm <= 0) {
                if (occ_cache <= 0) {
                    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, k_fused_tma<real, Q, K, SCHEME>, tile_cells, smem));
                    if (occ_cache < 1) occ_cache = 1;
                }
                per_sm = occ_cache;
            }
            const int64_t ntiles = (end - begin) / tile_cells;
            int64_t grid = (int64_t)num_sms * per_sm;
            if (grid > ntiles) grid = ntiles;
            CU_TRY(launch_k(k_fused_tma<real, Q, K, SCHEME>, (unsigned)grid, (unsigned)tile_cells, smem, st, pdl_chain(), a, stages));
        }
        ++launches;
        CU_TRY(cudaGetLastError());
        return FVDBM_OK;
    }

    // fp32 only: two cells per thread, 128 threads (= 256 cells) per CTA
    void launch_pair(const FusedArgs<float>& a, int64_t cells, cudaStream_t st) {
        launch_k(k_fused_pair<Q, K, SCHEME>, blocks_for(cells / 2, FVDBM_PAIR_THREADS), FVDBM_PAIR_THREADS, 0, st, pdl_chain(), a);
    }
    void launch_pair(const FusedArgs<double>&, int64_t, cudaStream_t) {}
    // record layout: packed k_fused_rec for fp32 D2Q9, the thread-per-cell kernel over records otherwise
    void launch_rec(const FusedArgs<float>& a, int64_t cells, cudaStream_t st) {
        if constexpr (Q == 9) launch_k(k_fused_rec<K, SCHEME>, blocks_for(cells, FVDBM_REC_THREADS), FVDBM_REC_THREADS, 0, st, pdl_chain(), a);
        else launch_k(k_fused_direct<float, Q, K, SCHEME, 1>, blocks_for(cells, 256), 256, 0, st, pdl_chain(), a);
    }
    void launch_rec(const FusedArgs<double>& a, int64_t cells, cudaStream_t st) {
        launch_k(k_fused_direct<double, Q, K, SCHEME, 1>, blocks_for(cells, 256), 256, 0, st, pdl_chain(), a);
    }
    static constexpr bool rec_available() { return true; }

    // the population buffers follow the variant's layout; switching re-lays both out (rare: set_option only)
    int set_layout(int want) {
        if (want == lay || !pdf[0].p) { lay = want; return FVDBM_OK; }
        DevBuf<real> tmp;
        CU_TRY(tmp.alloc(pdf[0].n));
        for (int b = 0; b < 2; ++b) {
            k_relayout<real, Q><<<blocks_for(plan.Npad, 256), 256, 0, stream>>>(pdf[b].p, lay, tmp.p, want, plan.Npad);
            ++launches;
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaMemcpyAsync(pdf[b].p, tmp.p, pdf[b].bytes(), cudaMemcpyDeviceToDevice, stream));
        }
        CU_TRY(cudaStreamSynchronize(stream));
        lay = want;
        return FVDBM_OK;
    }

    int ensure_staged_buffers() {
        if (s_flux.p) return FVDBM_OK;
        CU_TRY(s_flux.alloc((size_t)plan.F * Q));
        if (mode == FVDBM_MODE_STAGED) {
            CU_TRY(s_rho.alloc(plan.Npad)); CU_TRY(s_ux.alloc(plan.Npad)); CU_TRY(s_uy.alloc(plan.Npad));
            CU_TRY(s_feq.alloc(pdf[0].n));
        }
        return FVDBM_OK;
    }

    int launch_faces(const real* src_pdf) {
        FaceArgs<real> a;
        a.P = P; a.pdf = src_pdf; a.lay = lay; a.Npad = plan.Npad; a.fcell = s_fcell.p; a.fnode = s_fnode.p; a.fdist = s_fdist.p; a.fn = s_fn.p;
        a.fL = s_fL.p; a.npdf = npdf.p; a.NTpad = plan.NTpad; a.F = plan.F;
        a.last_pos = plan.pos[plan.N - 1]; a.flux = s_flux.p;
        k_s_faces<real, Q, SCHEME><<<blocks_for(plan.F, 256), 256, 0, stream>>>(a);
        ++launches;
        CU_TRY(cudaGetLastError());
        return FVDBM_OK;
    }

    int step_staged_once() {
        int rc = ensure_staged_buffers();
        if (rc) return rc;
        k_s_moments<real, Q><<<blocks_for(plan.Npad, 256), 256, 0, stream>>>(P, pdf[cur].p, lay, ipos.p, plan.Npad, s_rho.p,
                                                                             s_ux.p, s_uy.p, s_feq.p);
        ++launches;
        if ((rc = launch_nodes())) return rc;
        if ((rc = launch_faces(pdf[cur].p))) return rc;
        k_s_cells<real, Q, K><<<blocks_for(plan.Npad, 256), 256, 0, stream>>>(P, lay, pdf[cur].p, s_feq.p, s_flux.p, s_cface.p,
                                                                            s_csign.p, ipos.p, plan.Npad, plan.No, s_inv_area.p, pdf[nxt()].p);
        ++launches;
        CU_TRY(cudaGetLastError());
        prev = cur; cur = nxt(); ++steps;
        return FVDBM_OK;
    }

    int64_t owned_end() const { return round_up(plan.Oend, PAD_TO); }

    // interior cells on the side stream (forked from / joined into the main stream)
    int fork_interior() {
        if (plan.Bstart == 0) { phase0_done = true; return FVDBM_OK; }
        if (!overlap) { int rc = launch_fused(0, plan.Bstart, stream); phase0_done = true; return rc; }
        CU_TRY(cudaEventRecord(ev_fork, stream));
        CU_TRY(cudaStreamWaitEvent(stream2, ev_fork, 0));
        int rc = launch_fused(0, plan.Bstart, stream2);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(ev_join, stream2));
        phase0_done = true;
        forked = true;
        return FVDBM_OK;
    }
    bool forked = false;

    // One iteration: interior cells run concurrently with [node kernel -> border cells]; the caller may
    // have issued the interior part earlier (step_phase(0)) to overlap it with a halo exchange.
    // Latency-bound meshes: ONE stream, [node kernel -> all cells] per iteration, every launch programmatically
    // dependent on the previous one (PDL) so launch latency and the kernels' prologues (index math, streaming loads)
    // overlap the predecessor's tail.  Bandwidth-bound meshes keep the two-stream overlap schedule below.
    bool pdl_chain() const { return pdl && mode == FVDBM_MODE_FUSED && !native_exchange() && plan.No == plan.N && !phase0_done; }

    int step_fused_once() {
        int rc;
        if (pdl_chain()) {
            if ((rc = launch_nodes())) return rc;
            if ((rc = launch_fused(0, owned_end(), stream))) return rc;
            prev = cur; cur = nxt(); ++steps;
            return FVDBM_OK;
        }
        const bool xchg = native_exchange();
        if (xchg && phase0_done) { err = "step_phase(0) cannot be combined with the native exchange"; return FVDBM_ERR_STATE; }
        if (!phase0_done && (rc = fork_interior())) return rc;       // fork first: the exchange must not delay it
        if (xchg && (rc = exchange())) return rc;
        if ((rc = launch_nodes())) return rc;
        if ((rc = launch_fused(plan.Bstart, owned_end(), stream))) return rc;
        if (forked) CU_TRY(cudaStreamWaitEvent(stream, ev_join, 0));
        forked = false;
        phase0_done = false;
        prev = cur; cur = nxt(); ++steps;
        return FVDBM_OK;
    }

#define NCCL_TRY(expr)                                                                        \
    do {                                                                                      \
        ncclResult_t r_ = (expr);                                                             \
        if (r_ != ncclSuccess) {                                                              \
            err = std::string(#expr) + ": " + g_nccl.GetErrorString(r_);                      \
            return FVDBM_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

    // pack -> grouped ncclSend/ncclRecv with every neighbour -> unpack, all on the main stream
    // (the interior update is already running on the side stream)
    int exchange() {
        int rc;
        if ((rc = halo_pack(sendbuf.p))) return rc;
        const ncclDataType_t dt = sizeof(real) == 4 ? ncclFloat32 : ncclFloat64;
        NCCL_TRY(g_nccl.GroupStart());
        size_t off = 0;
        for (size_t i = 0; i < send_peers.size(); ++i) {
            NCCL_TRY(g_nccl.Send(sendbuf.p + off * Q, (size_t)send_counts[i] * Q, dt, send_peers[i], comm, stream));
            off += (size_t)send_counts[i];
        }
        off = 0;
        for (size_t i = 0; i < recv_peers.size(); ++i) {
            NCCL_TRY(g_nccl.Recv(recvbuf.p + off * Q, (size_t)recv_counts[i] * Q, dt, recv_peers[i], comm, stream));
            off += (size_t)recv_counts[i];
        }
        NCCL_TRY(g_nccl.GroupEnd());
        return halo_unpack(recvbuf.p);
    }

    int comm_init(int nranks, int rank, const void* id) override {
        CU_TRY(cudaSetDevice(device));
        if (!id || nranks < 1 || rank < 0 || rank >= nranks) { err = "bad communicator arguments"; return FVDBM_ERR_ARG; }
        if (!g_nccl.load()) { err = g_nccl.error; return FVDBM_ERR_CUDA; }
        if (comm) { g_nccl.CommDestroy(comm); comm = nullptr; }
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        NCCL_TRY(g_nccl.CommInitRank(&comm, nranks, uid, rank));
        drop_graphs();
        return FVDBM_OK;
    }

    int halo_set_peers(const int32_t* sp, const int64_t* sc, int ns, const int32_t* rp, const int64_t* rc, int nr) override {
        CU_TRY(cudaSetDevice(device));
        int64_t tot_s = 0, tot_r = 0;
        for (int i = 0; i < ns; ++i) tot_s += sc[i];
        for (int i = 0; i < nr; ++i) tot_r += rc[i];
        if (tot_s != (int64_t)halo_send.n || tot_r != (int64_t)halo_recv.n) {
            err = "peer counts do not add up to the halo lists"; return FVDBM_ERR_ARG;
        }
        send_peers.assign(sp, sp + ns); send_counts.assign(sc, sc + ns);
        recv_peers.assign(rp, rp + nr); recv_counts.assign(rc, rc + nr);
        CU_TRY(sendbuf.alloc((size_t)std::max<int64_t>(tot_s, 1) * Q));
        CU_TRY(recvbuf.alloc((size_t)std::max<int64_t>(tot_r, 1) * Q));
        drop_graphs();
        return FVDBM_OK;
    }

    int step_once() { return mode == FVDBM_MODE_STAGED ? step_staged_once() : step_fused_once(); }

    int step_phase(int phase) override {
        CU_TRY(cudaSetDevice(device));
        if (mode != FVDBM_MODE_FUSED) { err = "step_phase needs the fused mode"; return FVDBM_ERR_STATE; }
        if (phase == 0) {
            if (phase0_done) { err = "phase 0 already issued"; return FVDBM_ERR_STATE; }
            return fork_interior();
        }
        if (phase == 1) return step_fused_once();
        err = "phase must be 0 or 1";
        return FVDBM_ERR_ARG;
    }

    int build_graph(int start_cur, cudaGraphExec_t* out) {
        cudaGraph_t g = nullptr;
        const int64_t s0 = steps, l0 = launches;
        CU_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        int rc = FVDBM_OK;
        for (int i = 0; i < graph_steps && rc == FVDBM_OK; ++i) rc = step_once();
        cudaError_t e = cudaStreamEndCapture(stream, &g);
        steps = s0; launches = l0; cur = start_cur;
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) { err = std::string("graph capture: ") + cudaGetErrorString(e); return FVDBM_ERR_CUDA; }
        e = cudaGraphInstantiate(out, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { err = std::string("graph instantiate: ") + cudaGetErrorString(e); return FVDBM_ERR_CUDA; }
        return FVDBM_OK;
    }

    int step(int n) override {
        if (n < 0) { err = "nsteps must be >= 0"; return FVDBM_ERR_ARG; }
        CU_TRY(cudaSetDevice(device));
        int rc;
        if (mode == FVDBM_MODE_STAGED && (rc = ensure_staged_buffers())) return rc;
        while (n > 0) {
            if (graph_steps > 0 && n >= graph_steps && !phase0_done && !native_exchange()) {   // NCCL ops stay out of graphs
                auto it = graphs.find(cur);
                if (it == graphs.end()) {
                    cudaGraphExec_t ge;
                    const int64_t l0 = launches;
                    // count launches of one captured batch by replaying the bookkeeping
                    if ((rc = build_graph(cur, &ge))) return rc;
                    (void)l0;
                    it = graphs.emplace(cur, ge).first;
                }
                CU_TRY(cudaGraphLaunch(it->second, stream));
                steps += graph_steps;
                launches += (int64_t)graph_steps * launches_per_step();
                n -= graph_steps;           // graph_steps is even: `cur` is unchanged
            } else {
                if ((rc = step_once())) return rc;
                --n;
            }
        }
        return FVDBM_OK;
    }

    int64_t launches_per_step() const {
        if (mode == FVDBM_MODE_STAGED) return 3 + (plan.NA > 0 ? 1 : 0);
        // own kernels only: the grouped ncclSend/Recv of a native exchange is NCCL's launch, not counted
        if (pdl_chain()) return (plan.NA > 0 ? 1 : 0) + 1;
        return (plan.NA > 0 ? 1 : 0) + (plan.Bstart > 0 ? 1 : 0) + (owned_end() > plan.Bstart ? 1 : 0) +
               (native_exchange() ? (halo_send.n ? 1 : 0) + (halo_recv.n ? 1 : 0) : 0);
    }

    int step_timed(int n, float* ms) override {
        CU_TRY(cudaSetDevice(device));
        cudaEvent_t e0, e1;
        CU_TRY(cudaEventCreate(&e0));
        CU_TRY(cudaEventCreate(&e1));
        CU_TRY(cudaEventRecord(e0, stream));
        int rc = step(n);
        if (rc == FVDBM_OK) {
            cudaEventRecord(e1, stream);
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { err = std::string("step_timed: ") + cudaGetErrorString(e); rc = FVDBM_ERR_CUDA; }
            else if (ms) cudaEventElapsedTime(ms, e0, e1);
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return rc;
    }

    int sync() override {
        CU_TRY(cudaSetDevice(device));
        CU_TRY(cudaStreamSynchronize(xin));
        CU_TRY(cudaStreamSynchronize(stream));
        CU_TRY(cudaStreamSynchronize(xout));
        CU_TRY(cudaGetLastError());
        return FVDBM_OK;
    }

    // ---------------------------------------------------------------- transfers
    int wait_ticket(int64_t t) override {
        CU_TRY(cudaSetDevice(device));
        if (t < 0 || t >= tickets) { err = "unknown transfer ticket"; return FVDBM_ERR_ARG; }
        if (t + kTickets <= tickets) return FVDBM_OK;            // recycled: that transfer completed long ago
        CU_TRY(cudaEventSynchronize(ticket_ev[t % kTickets]));
        return FVDBM_OK;
    }
    // D2H of `bytes` from an outbox slot on the download stream; the ticket completes when dst is filled
    int ship(Outbox& o, const real* src, void* dst, size_t bytes, int64_t* ticket) {
        CU_TRY(cudaEventRecord(o.ready, stream));
        CU_TRY(cudaStreamWaitEvent(xout, o.ready, 0));
        CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, xout));
        CU_TRY(cudaEventRecord(o.drained, xout));
        o.used = true;
        const int64_t t = tickets++;
        CU_TRY(cudaEventRecord(ticket_ev[t % kTickets], xout));
        if (ticket) *ticket = t;
        return FVDBM_OK;
    }
    int claim_outbox(Outbox& o, size_t elems) {                 // next export may overwrite the slot once its last D2H drained
        if (o.buf.n < elems) {
            if (o.used) CU_TRY(cudaEventSynchronize(o.drained));
            CU_TRY(o.buf.alloc(elems));
            o.moments_steps = -1;
        }
        if (o.used) CU_TRY(cudaStreamWaitEvent(stream, o.drained, 0));
        return FVDBM_OK;
    }

    int get_async(int field, void* dst, size_t bytes, int64_t* ticket) override {
        CU_TRY(cudaSetDevice(device));
        if (!dst) { err = "dst is null"; return FVDBM_ERR_ARG; }
        const int64_t N = plan.N, F = plan.F, Pn = plan.P;
        auto expect = [&](size_t elems) {
            if (bytes != elems * sizeof(real)) { err = "size mismatch for field"; return false; }
            return true;
        };
        // cell fields may also be fetched for the owned prefix only ([0,N_owned) rows)
        int64_t rows = N;
        auto expect_cells = [&](size_t per) {
            if (bytes == (size_t)N * per * sizeof(real)) { rows = N; return true; }
            if (bytes == (size_t)plan.No * per * sizeof(real)) { rows = plan.No; return true; }
            err = "size mismatch for field";
            return false;
        };
        int rc;
        switch (field) {
        case FVDBM_CELL_PDF: case FVDBM_CELL_PDF_PREV: case FVDBM_CELL_PDF_EQ: {
            if (!expect_cells(Q)) return FVDBM_ERR_ARG;
            if (field != FVDBM_CELL_PDF && steps == 0) { err = "no step taken yet"; return FVDBM_ERR_STATE; }
            Outbox& o = outbox[tickets & 1];
            if ((rc = claim_outbox(o, (size_t)N * Q))) return rc;
            o.moments_steps = -1;
            if (field == FVDBM_CELL_PDF_EQ)
                k_export_moments<real, Q><<<blocks_for(rows, 256), 256, 0, stream>>>(P, pdf[prev].p, lay, plan.Npad, pos.p, rows, nullptr, nullptr, o.buf.p);
            else
                k_export_cells<real, Q><<<blocks_for(rows, 256), 256, 0, stream>>>(pdf[field == FVDBM_CELL_PDF ? cur : prev].p, lay, plan.Npad, pos.p, rows, o.buf.p);
            ++launches;
            CU_TRY(cudaGetLastError());
            return ship(o, o.buf.p, dst, bytes, ticket);
        }
        case FVDBM_CELL_RHO: case FVDBM_CELL_VEL: {
            // rho and vel share ONE export pass: the slot keeps [rho (rows) | vel (2 rows)] of the lagged
            // populations, so fetching the second of the two costs no kernel
            if (!expect_cells(field == FVDBM_CELL_RHO ? 1 : 2)) return FVDBM_ERR_ARG;
            if (steps == 0) { err = "no step taken yet"; return FVDBM_ERR_STATE; }
            Outbox* o = nullptr;
            for (Outbox& c : outbox) if (c.moments_steps == steps && c.moments_rows == rows) o = &c;
            if (!o) {
                o = &outbox[tickets & 1];
                if ((rc = claim_outbox(*o, (size_t)N * 3))) return rc;
                k_export_moments<real, Q><<<blocks_for(rows, 256), 256, 0, stream>>>(P, pdf[prev].p, lay, plan.Npad, pos.p, rows, o->buf.p, o->buf.p + rows, nullptr);
                ++launches;
                CU_TRY(cudaGetLastError());
                o->moments_steps = steps; o->moments_rows = rows;
            }
            return ship(*o, field == FVDBM_CELL_RHO ? o->buf.p : o->buf.p + rows, dst, bytes, ticket);
        }
        case FVDBM_FACE_FLUX: {
            if (!expect((size_t)F * Q)) return FVDBM_ERR_ARG;
            if (steps == 0) { err = "no step taken yet"; return FVDBM_ERR_STATE; }
            if ((rc = ensure_staged_buffers())) return rc;
            if (mode != FVDBM_MODE_STAGED && (rc = launch_faces(pdf[prev].p))) return rc;
            CU_TRY(cudaMemcpyAsync(dst, s_flux.p, bytes, cudaMemcpyDeviceToHost, stream));
            CU_TRY(cudaStreamSynchronize(stream));
            break;
        }
        case FVDBM_NODE_PDF: case FVDBM_NODE_RHO: case FVDBM_NODE_VEL: {
            const size_t per = field == FVDBM_NODE_RHO ? 1 : field == FVDBM_NODE_VEL ? 2 : Q;
            if (!expect((size_t)Pn * per)) return FVDBM_ERR_ARG;
            const DevBuf<real>& src = field == FVDBM_NODE_RHO ? nrho : field == FVDBM_NODE_VEL ? nvel : npdf;
            std::vector<real> host(src.n);
            if (src.n) CU_TRY(cudaMemcpyAsync(host.data(), src.p, src.bytes(), cudaMemcpyDeviceToHost, stream));
            CU_TRY(cudaStreamSynchronize(stream));
            real* out = static_cast<real*>(dst);
            for (int64_t t = 0; t < plan.NT; ++t)
                for (size_t j = 0; j < per; ++j) out[(size_t)plan.tn_orig[t] * per + j] = host[j * plan.NTpad + t];
            break;
        }
        default: err = "unknown field"; return FVDBM_ERR_ARG;
        }
        // fields served synchronously (O(sqrt N) node data, the staged flux observable): already complete
        const int64_t t = tickets++;
        CU_TRY(cudaEventRecord(ticket_ev[t % kTickets], stream));
        if (ticket) *ticket = t;
        return FVDBM_OK;
    }

    int get(int field, void* dst, size_t bytes) override {
        int64_t t = -1;
        int rc = get_async(field, dst, bytes, &t);
        return rc ? rc : wait_ticket(t);
    }

    // upload; returns once `src` may be reused when wait_host is set, immediately otherwise (src must then stay
    // valid until fvdbm_sync / a later blocking call).  The state change itself is ordered on the main stream.
    int set_impl(int field, const void* src, size_t bytes, bool wait_host) {
        CU_TRY(cudaSetDevice(device));
        if (!src) { err = "src is null"; return FVDBM_ERR_ARG; }
        const int64_t N = plan.N, Pn = plan.P;
        switch (field) {
        case FVDBM_CELL_PDF: {
            int64_t rows = N;                 // all local cells, or only the owned prefix
            if (bytes == (size_t)plan.No * Q * sizeof(real)) rows = plan.No;
            else if (bytes != (size_t)N * Q * sizeof(real)) { err = "size mismatch for field"; return FVDBM_ERR_ARG; }
            if (inbox.n < (size_t)N * Q) CU_TRY(inbox.alloc((size_t)N * Q));
            if (inbox_used) CU_TRY(cudaStreamWaitEvent(xin, inbox_consumed, 0));     // previous import has read the inbox
            CU_TRY(cudaMemcpyAsync(inbox.p, src, bytes, cudaMemcpyHostToDevice, xin));
            CU_TRY(cudaEventRecord(inbox_filled, xin));
            CU_TRY(cudaStreamWaitEvent(stream, inbox_filled, 0));
            k_import_cells<real, Q><<<blocks_for(rows, 256), 256, 0, stream>>>(pdf[cur].p, lay, plan.Npad, pos.p, rows, inbox.p);
            ++launches;
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaEventRecord(inbox_consumed, stream));
            inbox_used = true;
            for (Outbox& o : outbox) o.moments_steps = -1;
            if (wait_host) CU_TRY(cudaEventSynchronize(inbox_filled));
            return FVDBM_OK;
        }
        case FVDBM_NODE_PDF: case FVDBM_NODE_RHO: case FVDBM_NODE_VEL: {
            const size_t per = field == FVDBM_NODE_RHO ? 1 : field == FVDBM_NODE_VEL ? 2 : Q;
            if (bytes != (size_t)Pn * per * sizeof(real)) { err = "size mismatch for field"; return FVDBM_ERR_ARG; }
            DevBuf<real>& dstb = field == FVDBM_NODE_RHO ? nrho : field == FVDBM_NODE_VEL ? nvel : npdf;
            std::vector<real> host(dstb.n, real(0));
            const real* in = static_cast<const real*>(src);
            for (int64_t t = 0; t < plan.NT; ++t)
                for (size_t j = 0; j < per; ++j) host[j * plan.NTpad + t] = in[(size_t)plan.tn_orig[t] * per + j];
            if (dstb.n) CU_TRY(cudaMemcpyAsync(dstb.p, host.data(), dstb.bytes(), cudaMemcpyHostToDevice, stream));
            CU_TRY(cudaStreamSynchronize(stream));
            return FVDBM_OK;
        }
        default: err = "field is not settable"; return FVDBM_ERR_ARG;
        }
    }
    int set(int field, const void* src, size_t bytes) override { return set_impl(field, src, bytes, true); }
    int set_async(int field, const void* src, size_t bytes) override { return set_impl(field, src, bytes, false); }

    int check_finite(int64_t* bad) override {
        CU_TRY(cudaSetDevice(device));
        if (!bad) { err = "null argument"; return FVDBM_ERR_ARG; }
        if (!counter.p) CU_TRY(counter.alloc(1));
        CU_TRY(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), stream));
        k_count_nonfinite<real, Q><<<blocks_for(plan.Npad, 256), 256, 0, stream>>>(pdf[cur].p, lay, ipos.p, plan.Npad, plan.No, counter.p);
        ++launches;
        CU_TRY(cudaGetLastError());
        unsigned long long host = 0;
        CU_TRY(cudaMemcpyAsync(&host, counter.p, sizeof(host), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        *bad = (int64_t)host;
        return FVDBM_OK;
    }

    int set_params(double tau, double dt) override {
        if (!(tau > 0)) { err = "tau must be positive"; return FVDBM_ERR_ARG; }
        P.inv_tau = (real)(1.0 / tau);
        P.dt = (real)dt;
        drop_graphs();
        return FVDBM_OK;
    }

    int set_option(int opt, int64_t v) override {
        const int old_variant = variant, old_tile = tile_cells, old_stages = stages, old_graph = graph_steps;
        switch (opt) {
        case FVDBM_OPT_VARIANT: variant = v == FVDBM_VARIANT_AUTO ? default_variant() : (int)v; break;
        case FVDBM_OPT_TILE_CELLS: tile_cells = (int)v; break;
        case FVDBM_OPT_STAGES: stages = (int)v; break;
        case FVDBM_OPT_GRAPH_STEPS: graph_steps = (int)v; break;
        case FVDBM_OPT_CTAS_PER_SM: ctas_per_sm = (int)v; break;
        case FVDBM_OPT_REVERSE_SWEEP: reverse_sweep = v ? 1 : 0; break;
        case FVDBM_OPT_PDL: pdl = v ? 1 : 0; break;
        case FVDBM_OPT_PREFETCH_DIST: if (v < 0 || v > (1 << 20)) { err = "prefetch distance out of range"; return FVDBM_ERR_ARG; } prefetch_dist = (int)v; break;
        case FVDBM_OPT_TEMPORAL:
            if (v) { err = "temporal blocking was removed in ABI 2 (measured slower than the single-step kernel; DESIGN.md)"; return FVDBM_ERR_UNSUPPORTED; }
            break;
        default: err = "unknown option"; return FVDBM_ERR_ARG;
        }
        occ_cache = 0;
        int rc = sanitize_options();
        if (rc) { variant = old_variant; tile_cells = old_tile; stages = old_stages; graph_steps = old_graph; return rc; }
        drop_graphs();
        return FVDBM_OK;
    }

    int info(int key, int64_t* v) const override {
        if (!v) return FVDBM_ERR_ARG;
        switch (key) {
        case FVDBM_INFO_MODE: *v = mode; break;
        case FVDBM_INFO_STEPS: *v = steps; break;
        case FVDBM_INFO_LAUNCHES: *v = launches; break;
        case FVDBM_INFO_TRACKED_NODES: *v = plan.NT; break;
        case FVDBM_INFO_BOUNDARY_SIDES: *v = plan.NB; break;
        case FVDBM_INFO_DEVICE_BYTES:
            *v = (int64_t)(pdf[0].bytes() + pdf[1].bytes() + ccode.bytes() + ccoef.bytes() + bf_na.bytes() + bf_nb.bytes() +
                           bf_ratio.bytes() + pos.bytes() + ipos.bytes() + ring_cell.bytes() + ring_w.bytes() + tn_type.bytes() +
                           npdf.bytes() + nrho.bytes() + nvel.bytes() + s_cface.bytes() + s_csign.bytes() + s_fcell.bytes() +
                           s_fnode.bytes() + s_fdist.bytes() + s_fn.bytes() + s_fL.bytes() + s_inv_area.bytes() + s_rho.bytes() +
                           s_ux.bytes() + s_uy.bytes() + s_feq.bytes() + s_flux.bytes() + inbox.bytes() + outbox[0].buf.bytes() + outbox[1].buf.bytes() + halo_send.bytes() +
                           halo_recv.bytes() + sendbuf.bytes() + recvbuf.bytes() + counter.bytes());
            break;
        case FVDBM_INFO_VARIANT: *v = variant; break;
        case FVDBM_INFO_NPAD: *v = plan.Npad; break;
        case FVDBM_INFO_FUSED_OK: *v = plan.fused_ok ? 1 : 0; break;
        case FVDBM_INFO_HALO_CELLS: *v = plan.N - plan.No; break;
        case FVDBM_INFO_OWNED_CELLS: *v = plan.No; break;
        case FVDBM_INFO_GRAPH_STEPS: *v = graph_steps; break;
        default: return FVDBM_ERR_ARG;
        }
        return FVDBM_OK;
    }

    // ---------------------------------------------------------------- halo
    int halo_set_lists(const int32_t* s, int64_t ns, const int32_t* r, int64_t nr) override {
        CU_TRY(cudaSetDevice(device));
        std::vector<int32_t> hs((size_t)ns), hr((size_t)nr);
        for (int64_t i = 0; i < ns; ++i) {
            if (s[i] < 0 || s[i] >= plan.N) { err = "send cell out of range"; return FVDBM_ERR_ARG; }
            hs[i] = plan.pos[s[i]];
        }
        for (int64_t i = 0; i < nr; ++i) {
            if (r[i] < 0 || r[i] >= plan.N) { err = "recv cell out of range"; return FVDBM_ERR_ARG; }
            hr[i] = plan.pos[r[i]];
        }
        CU_TRY(halo_send.upload(hs, stream));
        CU_TRY(halo_recv.upload(hr, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        return FVDBM_OK;
    }
    int halo_pack(void* buf) override {
        CU_TRY(cudaSetDevice(device));
        if (halo_send.n == 0) return FVDBM_OK;
        k_pack<real, Q><<<blocks_for((int64_t)halo_send.n * Q, 256), 256, 0, stream>>>(pdf[cur].p, lay, plan.Npad, halo_send.p, (int64_t)halo_send.n,
                                                                                   static_cast<real*>(buf));
        ++launches;
        CU_TRY(cudaGetLastError());
        return FVDBM_OK;
    }
    int halo_unpack(const void* buf) override {
        CU_TRY(cudaSetDevice(device));
        if (halo_recv.n == 0) return FVDBM_OK;
        k_unpack<real, Q><<<blocks_for((int64_t)halo_recv.n * Q, 256), 256, 0, stream>>>(pdf[cur].p, lay, plan.Npad, halo_recv.p, (int64_t)halo_recv.n,
                                                                                     static_cast<const real*>(buf));
        ++launches;
        CU_TRY(cudaGetLastError());
        return FVDBM_OK;
    }
    void* stream_handle() override { return (void*)stream; }
};

template <typename real, int Q, int K>
Engine* make_scheme(const fvdbm_desc& d, int* rc, std::string* msg) {
    Engine* e = nullptr;
    if (d.scheme == FVDBM_SCHEME_UPWIND) {
        auto* t = new EngineT<real, Q, K, 0>(); *rc = t->init(d); e = t;
    } else if (d.scheme == FVDBM_SCHEME_LAX_WENDROFF) {
        auto* t = new EngineT<real, Q, K, 1>(); *rc = t->init(d); e = t;
    } else { *rc = FVDBM_ERR_ARG; *msg = "Unknown flux scheme"; return nullptr; }
    if (*rc) { *msg = e->err; delete e; return nullptr; }
    return e;
}

template <typename real>
Engine* make_engine(const fvdbm_desc& d, int* rc, std::string* msg) {
    if (d.Q == 9 && d.K == 3) return make_scheme<real, 9, 3>(d, rc, msg);
    if (d.Q == 9 && d.K == 4) return make_scheme<real, 9, 4>(d, rc, msg);
    if (d.Q == 13 && d.K == 3) return make_scheme<real, 13, 3>(d, rc, msg);
    if (d.Q == 13 && d.K == 4) return make_scheme<real, 13, 4>(d, rc, msg);
    *rc = FVDBM_ERR_UNSUPPORTED; *msg = "unsupported (Q,K): Q in {9,13}, K in {3,4}";
    return nullptr;
}

struct PlanBox {
    int dtype = 32;
    Plan<float> f;
    Plan<double> d;
};

}  // namespace

namespace {
template <typename real>
int64_t plan_array(const Plan<real>& p, const std::string& k, const void** ptr, int32_t* eb) {
#define I32(name) if (k == #name) { *ptr = p.name.data(); *eb = 4; return (int64_t)p.name.size(); }
#define REAL(name) if (k == #name) { *ptr = p.name.data(); *eb = (int32_t)sizeof(real); return (int64_t)p.name.size(); }
    I32(pos) I32(ipos) I32(ccode) I32(bf_na) I32(bf_nb) I32(tn_orig) I32(tn_type) I32(node_track)
    I32(ring_off) I32(ring_cell) I32(ring_fcell) I32(s_cface) I32(s_csign) I32(s_fcell) I32(s_fnode)
    REAL(ccoef) REAL(bf_ratio) REAL(ring_w) REAL(ring_fw) REAL(tn_pdf) REAL(tn_rho) REAL(tn_vel) REAL(s_inv_area)
#undef I32
#undef REAL
    return -1;
}
template <typename real>
int64_t plan_scalar(const Plan<real>& p, const std::string& k) {
    if (k == "N") return p.N; if (k == "F") return p.F; if (k == "P") return p.P; if (k == "No") return p.No;
    if (k == "Npad") return p.Npad; if (k == "Bstart") return p.Bstart; if (k == "Oend") return p.Oend;
    if (k == "Hstart") return p.Hstart; if (k == "NB") return p.NB; if (k == "NT") return p.NT;
    if (k == "NTpad") return p.NTpad; if (k == "NA") return p.NA; if (k == "MR") return p.MR; if (k == "NC") return p.NC;
    if (k == "fused_ok") return p.fused_ok ? 1 : 0;
    return -1;
}
}  // namespace

struct fvdbm_handle { Engine* e; };
struct fvdbm_plan { PlanBox b; };

extern "C" {

int fvdbm_abi_version(void) { return FVDBM_ABI_VERSION; }

const char* fvdbm_last_error(const fvdbm_handle* h) { return h ? h->e->err.c_str() : g_create_error.c_str(); }

int fvdbm_create(const fvdbm_desc* desc, fvdbm_handle** out) {
    if (!desc || !out) { g_create_error = "null argument"; return FVDBM_ERR_ARG; }
    *out = nullptr;
    if (desc->abi_version != FVDBM_ABI_VERSION) { g_create_error = "abi_version mismatch"; return FVDBM_ERR_ARG; }
    int rc = FVDBM_OK;
    std::string msg;
    Engine* e = nullptr;
    if (desc->dtype == 32) e = make_engine<float>(*desc, &rc, &msg);
    else if (desc->dtype == 64) e = make_engine<double>(*desc, &rc, &msg);
    else { rc = FVDBM_ERR_ARG; msg = "dtype must be 32 or 64"; }
    if (!e) { g_create_error = msg; return rc ? rc : FVDBM_ERR_ARG; }
    *out = new fvdbm_handle{e};
    return FVDBM_OK;
}

void fvdbm_destroy(fvdbm_handle* h) { if (h) { delete h->e; delete h; } }
int fvdbm_step(fvdbm_handle* h, int n) { return h ? h->e->step(n) : FVDBM_ERR_ARG; }
int fvdbm_step_timed(fvdbm_handle* h, int n, float* ms) { return h ? h->e->step_timed(n, ms) : FVDBM_ERR_ARG; }
int fvdbm_step_phase(fvdbm_handle* h, int phase) { return h ? h->e->step_phase(phase) : FVDBM_ERR_ARG; }
int fvdbm_sync(fvdbm_handle* h) { return h ? h->e->sync() : FVDBM_ERR_ARG; }
int fvdbm_get(fvdbm_handle* h, int field, void* dst, size_t bytes) { return h ? h->e->get(field, dst, bytes) : FVDBM_ERR_ARG; }
int fvdbm_set(fvdbm_handle* h, int field, const void* src, size_t bytes) { return h ? h->e->set(field, src, bytes) : FVDBM_ERR_ARG; }
int fvdbm_set_async(fvdbm_handle* h, int field, const void* src, size_t bytes) { return h ? h->e->set_async(field, src, bytes) : FVDBM_ERR_ARG; }
int fvdbm_get_async(fvdbm_handle* h, int field, void* dst, size_t bytes, int64_t* ticket) { return h ? h->e->get_async(field, dst, bytes, ticket) : FVDBM_ERR_ARG; }
int fvdbm_wait(fvdbm_handle* h, int64_t ticket) { return h ? h->e->wait_ticket(ticket) : FVDBM_ERR_ARG; }
int fvdbm_set_params(fvdbm_handle* h, double tau, double dt) { return h ? h->e->set_params(tau, dt) : FVDBM_ERR_ARG; }
int fvdbm_set_option(fvdbm_handle* h, int opt, int64_t v) { return h ? h->e->set_option(opt, v) : FVDBM_ERR_ARG; }
int fvdbm_info(const fvdbm_handle* h, int key, int64_t* v) { return h ? h->e->info(key, v) : FVDBM_ERR_ARG; }
int fvdbm_check_finite(fvdbm_handle* h, int64_t* bad) { return h ? h->e->check_finite(bad) : FVDBM_ERR_ARG; }
int fvdbm_halo_set_lists(fvdbm_handle* h, const int32_t* s, int64_t ns, const int32_t* r, int64_t nr) {
    return h ? h->e->halo_set_lists(s, ns, r, nr) : FVDBM_ERR_ARG;
}
int fvdbm_halo_pack(fvdbm_handle* h, void* buf) { return h ? h->e->halo_pack(buf) : FVDBM_ERR_ARG; }
int fvdbm_halo_unpack(fvdbm_handle* h, const void* buf) { return h ? h->e->halo_unpack(buf) : FVDBM_ERR_ARG; }
void* fvdbm_stream(fvdbm_handle* h) { return h ? h->e->stream_handle() : nullptr; }
int fvdbm_comm_unique_id(void* id_out) {
    if (!id_out) { g_create_error = "null argument"; return FVDBM_ERR_ARG; }
    if (!g_nccl.load()) { g_create_error = g_nccl.error; return FVDBM_ERR_CUDA; }
    ncclUniqueId uid;
    if (g_nccl.GetUniqueId(&uid) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return FVDBM_ERR_CUDA; }
    memcpy(id_out, &uid, sizeof(uid));
    return FVDBM_OK;
}
int fvdbm_comm_init(fvdbm_handle* h, int nranks, int rank, const void* id) { return h ? h->e->comm_init(nranks, rank, id) : FVDBM_ERR_ARG; }
int fvdbm_halo_set_peers(fvdbm_handle* h, const int32_t* sp, const int64_t* sc, int ns, const int32_t* rp, const int64_t* rc, int nr) {
    return h ? h->e->halo_set_peers(sp, sc, ns, rp, rc, nr) : FVDBM_ERR_ARG;
}

// ---- host-only helpers ---------------------------------------------------------------------------
// Hilbert-curve key of every cell centroid (same curve as reorder.hilbert_index), OpenMP over cells.
int fvdbm_sfc_keys(const double* points, const int32_t* elements, int64_t ncells, int K, int bits, double lo_x, double lo_y,
                   double scale, int64_t* keys) {
    if (!points || !keys || ncells < 0 || (elements && (K < 3 || K > 4)) || bits < 1 || bits > 30) { g_create_error = "bad argument"; return FVDBM_ERR_ARG; }
    const int64_t n1 = (int64_t(1) << bits) - 1;
    const int nthreads = fvdbm::plan_threads();
    (void)nthreads;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t c = 0; c < ncells; ++c) {
        double cx = 0, cy = 0;
        if (elements) {
            for (int k = 0; k < K; ++k) { cx += points[2 * (int64_t)elements[c * K + k]]; cy += points[2 * (int64_t)elements[c * K + k] + 1]; }
            cx /= K; cy /= K;
        } else { cx = points[2 * c]; cy = points[2 * c + 1]; }      // points ARE the centroids
        int64_t x = (int64_t)((cx - lo_x) * scale), y = (int64_t)((cy - lo_y) * scale);
        x = x < 0 ? 0 : (x > n1 ? n1 : x); y = y < 0 ? 0 : (y > n1 ? n1 : y);
        int64_t d = 0;
        for (int64_t s = int64_t(1) << (bits - 1); s > 0; s >>= 1) {
            const int64_t rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
            d += s * s * ((3 * rx) ^ ry);
            if (ry == 0) {
                if (rx == 1) { x = n1 - x; y = n1 - y; }
                const int64_t t = x; x = y; y = t;
            }
        }
        keys[c] = d;
    }
    return FVDBM_OK;
}

// ---- host-only Mesher-equivalent (csrc/mesh.hpp) ---------------------------------------------------
int64_t fvdbm_mesh_ring_width(const int32_t* cells, const int32_t* point_alias, int64_t N, int64_t P) {
    if ((!cells && N > 0) || N < 0 || P < 0) { g_create_error = "bad argument"; return FVDBM_ERR_ARG; }
    const int64_t m = fvdbm::mesh_ring_width(cells, point_alias, N, P);
    if (m < 0) g_create_error = "point id out of range";
    return m;
}

int fvdbm_mesh_properties(const fvdbm_mesh_desc* d) {
    if (!d) { g_create_error = "null argument"; return FVDBM_ERR_ARG; }
    const bool has_c = d->N > 0, has_f = d->F > 0, has_r = d->P > 0 && d->M > 0;
    if (d->N < 0 || d->F < 0 || d->P < 0 || d->M < 0 || (d->P > 0 && !d->points) || (has_c && !d->cells) || (has_f && !d->faces) ||
        (has_c && (!d->cell_centers || !d->cell_face_indices || !d->cell_face_normals || !d->cell_face_normal_signs)) ||
        (has_f && (!d->faces_out || !d->face_centers || !d->face_normals || !d->face_lengths || !d->face_cell_indices ||
                   !d->face_cell_center_distances || !d->stencil_norms || !d->cc_stencil_dist || !d->face_stencil_angles)) ||
        (has_r && (!d->point_cell_indices || !d->point_cell_center_distances))) {
        g_create_error = "fvdbm_mesh_properties: null array";
        return FVDBM_ERR_ARG;
    }
    fvdbm::MeshIn in;
    in.N = d->N; in.F = d->F; in.P = d->P;
    in.points = d->points; in.cells = d->cells; in.faces = d->faces; in.alias = d->point_alias;
    fvdbm::MeshOut o;
    o.M = d->M;
    o.cell_centers = d->cell_centers; o.cell_face_indices = d->cell_face_indices; o.cell_face_normals = d->cell_face_normals;
    o.cell_face_normal_signs = d->cell_face_normal_signs; o.faces = d->faces_out; o.face_centers = d->face_centers;
    o.face_normals = d->face_normals; o.face_lengths = d->face_lengths; o.face_cell_indices = d->face_cell_indices;
    o.face_cell_center_distances = d->face_cell_center_distances; o.stencil_norms = d->stencil_norms;
    o.cc_stencil_dist = d->cc_stencil_dist; o.face_stencil_angles = d->face_stencil_angles;
    o.point_cell_indices = d->point_cell_indices; o.point_cell_center_distances = d->point_cell_center_distances;
    std::string err;
    const int rc = fvdbm::mesh_properties(in, o, err);
    if (rc == 0) return FVDBM_OK;
    g_create_error = err;
    return rc == -2 ? FVDBM_ERR_STATE : FVDBM_ERR_ARG;
}

int64_t fvdbm_mesh_unique_edges(const int32_t* cells, int64_t N, int K, int64_t P, const int32_t* point_alias, int32_t* faces_out) {
    if (N > 0 && (!cells || !faces_out)) { g_create_error = "null argument"; return FVDBM_ERR_ARG; }
    const int64_t f = fvdbm::mesh_unique_edges(cells, N, K, P, point_alias, faces_out);
    if (f < 0) { g_create_error = "fvdbm_mesh_unique_edges: point id out of range, K not 3 or 4, or too many cell edges"; return FVDBM_ERR_ARG; }
    return f;
}

// ---- host-only planning -------------------------------------------------------------------------
int fvdbm_plan_create(const fvdbm_desc* desc, fvdbm_plan** out) {
    if (!desc || !out) { g_create_error = "null argument"; return FVDBM_ERR_ARG; }
    *out = nullptr;
    auto* p = new fvdbm_plan();
    p->b.dtype = desc->dtype;
    bool ok = false;
    if (desc->dtype == 32) { ok = p->b.f.build(*desc); if (!ok) g_create_error = p->b.f.error; }
    else if (desc->dtype == 64) { ok = p->b.d.build(*desc); if (!ok) g_create_error = p->b.d.error; }
    else g_create_error = "dtype must be 32 or 64";
    if (!ok) { delete p; return FVDBM_ERR_ARG; }
    *out = p;
    return FVDBM_OK;
}
void fvdbm_plan_destroy(fvdbm_plan* p) { delete p; }

int64_t fvdbm_plan_array(const fvdbm_plan* p, const char* key, const void** ptr, int32_t* eb) {
    if (!p || !key || !ptr || !eb) return -1;
    return p->b.dtype == 32 ? plan_array(p->b.f, key, ptr, eb) : plan_array(p->b.d, key, ptr, eb);
}
int64_t fvdbm_plan_scalar(const fvdbm_plan* p, const char* key) {
    if (!p || !key) return -1;
    return p->b.dtype == 32 ? plan_scalar(p->b.f, key) : plan_scalar(p->b.d, key);
}

}  // extern "C"
