// plan.hpp -- host-side planning: reference layout (AoS, original numbering) -> device layout.
//
// Pure C++ (no CUDA) so it can be exercised on a CPU-only box through fvdbm_plan_create().  The O(N)
// passes are OpenMP-parallel (FVDBM_PLAN_THREADS, default omp_get_max_threads()).
//
// Device layout produced here (DESIGN.md "Data layout in HBM"):
//   * cells live at "positions" pos[i] in [0,Npad); positions are grouped
//       [ interior | pad | border | pad | halo | pad ],  every group starting on a PAD_TO boundary;
//     border = owned cells with a boundary side or a halo neighbour (O(sqrt N)); interior = the rest.
//   * populations are tiled AoSoA: value (cell p, population q) at (p>>5)*(Q*32) + q*32 + (p&31),
//     so a warp reads 128 contiguous bytes per population and any CTA tile (multiple of 32 cells)
//     is one contiguous block for cp.async.bulk.
//   * per (cell,k) "side" record replaces Cells.face_indices/face_normals + Faces.stencil_*:
//       ccode: interior  (nbr_pos<<2) | (sign<0)<<1 | slot        slot = stencil slot of THIS cell
//              boundary  -(((bside<<2) | (sign<0)<<1 | slot) + 1)
//              hole      INT32_MIN in k=0 (padding position, skipped)
//       ccoef: NC reals per side, in CELL orientation (sigma = sign entry, varsigma = +1 in slot 0, -1 in
//              slot 1; sign folds are exact): M = sigma n L [inv_area] (2), and for Lax-Wendroff
//              A = varsigma d0/(d0+d1), G = sigma varsigma /(2 (d0+d1) L [inv_area])
//              (src/containers.py:266-277; core.cuh: side_flux)
//   * boundary sides: the two tracked-node ids of the face and ratio d_ghost/d_known
//     (src/containers.py:280-287, utils/utils.py:153-154)
//   * tracked nodes (type 1/2, or on a face with a ghost slot): compact ids, ring CSR with weights
//     w = 1/d, negative -> 0 (utils/utils.py:58-59), zero-weight entries dropped.
#pragma once
#include <cstdint>
#include <climits>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#if defined(_OPENMP)
#include <omp.h>
#endif
#include "../../include/fvdbm_b200.h"

namespace fvdbm {

constexpr int TW = 32;            // lanes of one AoSoA mini-tile
constexpr int PAD_TO = 512;       // group alignment = largest CTA tile
constexpr int32_t HOLE = INT32_MIN;

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

inline int plan_threads() {
    int n = 1;
#if defined(_OPENMP)
    n = omp_get_max_threads();
#endif
    if (const char* e = getenv("FVDBM_PLAN_THREADS")) n = atoi(e);
    return n < 1 ? 1 : n;
}

template <typename real>
struct Plan {
    int64_t N = 0, F = 0, P = 0, No = 0;
    int Q = 9, K = 3, M = 0, scheme = 0, NC = 2;
    int64_t Npad = 0, Bstart = 0, Oend = 0, Hstart = 0;
    std::vector<int32_t> pos, ipos;
    bool fused_ok = true;
    std::string why_not;
    std::vector<int32_t> ccode;                 // side codes
    std::vector<real> ccoef;                    // per-side coefficients, tiled like the codes
    int64_t NB = 0;
    std::vector<int32_t> bf_na, bf_nb;
    std::vector<real> bf_ratio;
    int64_t NT = 0, NTpad = 0, NA = 0;          // tracked nodes; [0,NA) active (type 1 or 2)
    std::vector<int32_t> tn_orig, tn_type, node_track, ring_off, ring_cell;
    std::vector<real> ring_w, tn_pdf, tn_rho, tn_vel;
    // the same rings as a fixed-width table [NA][MR] (zero weight = unused slot): lets the node kernel index
    // its ring entries directly instead of first loading CSR offsets (one dependent memory level less)
    std::vector<int32_t> ring_fcell;
    std::vector<real> ring_fw;
    int64_t MR = 0;
    std::vector<int32_t> s_cface, s_csign, s_fcell, s_fnode;
    std::vector<real> s_inv_area;               // staged path: optional 1/area per position (empty = 1)
    std::string error;

    bool fail(const std::string& msg) { error = msg; return false; }

    bool build(const fvdbm_desc& d) {
        N = d.N; F = d.F; P = d.P; Q = d.Q; K = d.K; M = d.M; scheme = d.scheme;
        No = (d.N_owned <= 0 || d.N_owned > d.N) ? d.N : d.N_owned;
        NC = scheme == FVDBM_SCHEME_LAX_WENDROFF ? 4 : 2;
        if (N <= 0 || F <= 0 || P < 0) return fail("N and F must be positive");
        if (N >= (int64_t(1) << 29)) return fail("at most 2^29-1 cells per handle");
        if (!(Q == 9 || Q == 13)) return fail("Q must be 9 or 13");
        if (!(K == 3 || K == 4)) return fail("K must be 3 or 4");
        if (scheme != FVDBM_SCHEME_UPWIND && scheme != FVDBM_SCHEME_LAX_WENDROFF)
            return fail("Unknown flux scheme");
        if (!d.cell_face_idx || !d.cell_face_sign || !d.face_cell_idx || !d.face_dists || !d.face_node_idx ||
            !d.face_n || !d.face_L || !d.cell_pdf)
            return fail("missing static array");
        if (P > 0 && (!d.node_type || !d.node_pdf || !d.node_rho || !d.node_vel)) return fail("missing node array");
        if (P > 0 && M > 0 && (!d.node_cell_idx || !d.node_cell_dist)) return fail("missing node ring arrays");
        for (int q = 0; q < Q; ++q) {      // core.cuh evaluates W per weight class
            const int cls = q == 0 ? 0 : q <= 4 ? 1 : q <= 8 ? 5 : 9;
            if (d.lat_w[q] != d.lat_w[cls]) return fail("lattice weights must be constant on {0},{1..4},{5..8},{9..12}");
        }
        const real* fdist = static_cast<const real*>(d.face_dists);
        const real* fn = static_cast<const real*>(d.face_n);
        const real* fL = static_cast<const real*>(d.face_L);
        const real* inv_area = static_cast<const real*>(d.cell_inv_area);     // optional, NULL = reference (no area)
        const int nthreads = plan_threads();
        (void)nthreads;

        // ---- range checks -------------------------------------------------------------------
        {
            int bad_cell = 0, bad_face = 0, bad_type = 0;
#pragma omp parallel for num_threads(nthreads) reduction(| : bad_cell)
            for (int64_t i = 0; i < No * K; ++i)      // halo cells carry no sides (never updated)
                if (d.cell_face_idx[i] < 0 || d.cell_face_idx[i] >= F) bad_cell |= 1;
#pragma omp parallel for num_threads(nthreads) reduction(| : bad_face)
            for (int64_t i = 0; i < F * 2; ++i)
                if (d.face_cell_idx[i] < -1 || d.face_cell_idx[i] >= N) bad_face |= 1;
            for (int64_t n = 0; n < P; ++n)
                if (d.node_type[n] < 0 || d.node_type[n] > 2) bad_type |= 1;
            if (bad_cell) return fail("cell_face_idx out of range (ragged / -1 padded cells are not supported)");
            if (bad_face) return fail("face_cell_idx out of range");
            if (bad_type) return fail("node_type must be 0 (none), 1 (velocity) or 2 (density)");   // containers.py:339-351
        }

        // ---- side analysis in original numbering -----------------------------------------------
        // other[i*K+k] = neighbour cell (>=0), -1 boundary, -2 inconsistent
        std::vector<int32_t> other((size_t)N * K, -2);
        std::vector<uint8_t> slot((size_t)N * K, 0);
        fused_ok = true;
        {
            int why = 0;      // 3: sign not +-1, 2: stencil does not contain the cell once, 1: boundary face without nodes
#pragma omp parallel for num_threads(nthreads) reduction(max : why)
            for (int64_t c = 0; c < No; ++c)
                for (int k = 0; k < K; ++k) {
                    const int64_t j = d.cell_face_idx[c * K + k];
                    const int32_t a = d.face_cell_idx[2 * j], b = d.face_cell_idx[2 * j + 1];
                    const int32_t s = d.cell_face_sign[c * K + k];
                    if (s != 1 && s != -1) { why = std::max(why, 3); continue; }
                    if (a == c && b != c) { slot[c * K + k] = 0; other[c * K + k] = b; }
                    else if (b == c && a != c) { slot[c * K + k] = 1; other[c * K + k] = a; }
                    else { why = std::max(why, 2); continue; }
                    if (other[c * K + k] == -1) {
                        const int32_t na = d.face_node_idx[2 * j], nb = d.face_node_idx[2 * j + 1];
                        if (na < 0 || na >= P || nb < 0 || nb >= P) why = std::max(why, 1);
                    }
                }
            if (why) {
                fused_ok = false;
                why_not = why == 3 ? "cell_face_sign not +-1"
                        : why == 2 ? "a cell lists a face whose stencil does not contain it exactly once"
                                   : "boundary face without valid nodes";
            }
        }

        // ---- positions ------------------------------------------------------------------------
        std::vector<int32_t> order(N);               // rank -> original cell
        if (d.cell_perm) {
            std::vector<uint8_t> seen(N, 0);
            for (int64_t i = 0; i < N; ++i) {
                int32_t r = d.cell_perm[i];
                if (r < 0 || r >= N || seen[r]) return fail("cell_perm is not a bijection onto [0,N)");
                if ((i < No) != (r < No)) return fail("cell_perm must keep owned cells in [0,N_owned)");
                seen[r] = 1; order[r] = (int32_t)i;
            }
        } else for (int64_t i = 0; i < N; ++i) order[i] = (int32_t)i;

        pos.assign(N, -1);
        const bool has_halo = No < N;
        if (has_halo && !fused_ok) return fail("halo handles need a consistent mesh: " + why_not);
        int64_t p = 0;
        if (fused_ok) {
            // interior = owned cells whose K sides are all interior faces to owned cells: they need neither
            // node values nor halo copies, so the engine updates them concurrently with the exchange /
            // node kernel / border update (api.cu: step_fused_once).
            std::vector<uint8_t> border(N, 0);
#pragma omp parallel for num_threads(nthreads)
            for (int64_t c = 0; c < No; ++c)
                for (int k = 0; k < K; ++k) {
                    const int32_t o = other[c * K + k];
                    if (o == -1 || o >= No) border[c] = 1;
                }
            for (int64_t r = 0; r < No; ++r) if (!border[order[r]]) pos[order[r]] = (int32_t)p++;
            Bstart = round_up(p, PAD_TO); p = Bstart;
            for (int64_t r = 0; r < No; ++r) if (border[order[r]]) pos[order[r]] = (int32_t)p++;
            Oend = p; Hstart = round_up(p, PAD_TO); p = Hstart;
            for (int64_t r = No; r < N; ++r) pos[order[r]] = (int32_t)p++;
        } else {
            Bstart = 0;
            for (int64_t r = 0; r < N; ++r) pos[order[r]] = (int32_t)p++;
            Oend = p; Hstart = round_up(p, PAD_TO);
        }
        Npad = round_up(std::max<int64_t>(p, 1), PAD_TO);
        ipos.assign(Npad, -1);
#pragma omp parallel for num_threads(nthreads)
        for (int64_t i = 0; i < N; ++i) ipos[pos[i]] = (int32_t)i;

        // ---- tracked nodes --------------------------------------------------------------------
        node_track.assign(P, -1);
        std::vector<uint8_t> want(P, 0);
        for (int64_t n = 0; n < P; ++n) if (d.node_type[n] != 0) want[n] = 1;
        for (int64_t j = 0; j < F; ++j)
            if (d.face_cell_idx[2 * j] == -1 || d.face_cell_idx[2 * j + 1] == -1)
                for (int e = 0; e < 2; ++e) {
                    int32_t n = d.face_node_idx[2 * j + e];
                    if (n >= 0 && n < P) want[n] = 1;
                }
        // order: active nodes (type 1 / 2, re-evaluated every step) | inactive tracked nodes (keep their PDFs)
        tn_orig.clear(); tn_type.clear(); NA = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (int64_t n = 0; n < P; ++n) {
                if (!want[n]) continue;
                const bool active = d.node_type[n] != 0;
                if ((active ? 0 : 1) != pass) continue;
                node_track[n] = (int32_t)tn_orig.size();
                tn_orig.push_back((int32_t)n);
                tn_type.push_back(d.node_type[n]);
                if (active) ++NA;
            }
        NT = (int64_t)tn_orig.size();
        NTpad = round_up(std::max<int64_t>(NT, 1), TW);
        const real* npdf = static_cast<const real*>(d.node_pdf);
        const real* nrho = static_cast<const real*>(d.node_rho);
        const real* nvel = static_cast<const real*>(d.node_vel);
        tn_pdf.assign((size_t)Q * NTpad, real(0));
        tn_rho.assign(NTpad, real(0));
        tn_vel.assign(2 * NTpad, real(0));
        for (int64_t t = 0; t < NT; ++t) {
            int64_t n = tn_orig[t];
            for (int q = 0; q < Q; ++q) tn_pdf[(size_t)q * NTpad + t] = npdf[n * Q + q];
            tn_rho[t] = nrho[n];
            tn_vel[t] = nvel[2 * n]; tn_vel[NTpad + t] = nvel[2 * n + 1];
        }
        // ring CSR of the active nodes
        ring_off.assign(NA + 1, 0);
        ring_cell.clear(); ring_w.clear();
        const real* ncd = static_cast<const real*>(d.node_cell_dist);
        for (int64_t t = 0; t < NA; ++t) {
            int64_t n = tn_orig[t];
            for (int m = 0; m < M; ++m) {
                int32_t c = d.node_cell_idx[n * M + m];
                real w = real(1) / ncd[n * M + m];
                if (w < 0) w = 0;                                   // utils/utils.py:59
                if (!(w != 0)) continue;                            // zero weight contributes nothing
                if (c < 0 || c >= N) return fail("node ring entry with positive weight but invalid cell index");
                ring_cell.push_back(pos[c]);
                ring_w.push_back(w);
            }
            ring_off[t + 1] = (int32_t)ring_cell.size();
        }

        MR = 1;
        for (int64_t t = 0; t < NA; ++t) MR = std::max<int64_t>(MR, ring_off[t + 1] - ring_off[t]);
        ring_fcell.assign((size_t)std::max<int64_t>(NA, 1) * MR, 0);
        ring_fw.assign((size_t)std::max<int64_t>(NA, 1) * MR, real(0));
        for (int64_t t = 0; t < NA; ++t)
            for (int i = ring_off[t]; i < ring_off[t + 1]; ++i) {
                ring_fcell[(size_t)t * MR + (i - ring_off[t])] = ring_cell[i];
                ring_fw[(size_t)t * MR + (i - ring_off[t])] = ring_w[i];
            }

        // ---- staged statics (general path + observables) ---------------------------------------
        s_cface.assign((size_t)K * Npad, 0);
        s_csign.assign((size_t)K * Npad, 0);
#pragma omp parallel for num_threads(nthreads)
        for (int64_t c = 0; c < No; ++c)
            for (int k = 0; k < K; ++k) {
                s_cface[(size_t)k * Npad + pos[c]] = d.cell_face_idx[c * K + k];
                s_csign[(size_t)k * Npad + pos[c]] = d.cell_face_sign[c * K + k];
            }
        s_inv_area.clear();
        if (inv_area) {
            s_inv_area.assign(Npad, real(1));
            for (int64_t c = 0; c < No; ++c) s_inv_area[pos[c]] = inv_area[c];
        }
        s_fcell.resize(2 * F); s_fnode.resize(2 * F);
#pragma omp parallel for num_threads(nthreads)
        for (int64_t j = 0; j < F; ++j) {
            bool ghost = d.face_cell_idx[2 * j] == -1 || d.face_cell_idx[2 * j + 1] == -1;
            for (int e = 0; e < 2; ++e) {
                int32_t c = d.face_cell_idx[2 * j + e];
                s_fcell[2 * j + e] = c < 0 ? -1 : pos[c];
                int32_t n = d.face_node_idx[2 * j + e];
                s_fnode[2 * j + e] = (ghost && n >= 0 && n < P) ? node_track[n] : -1;
            }
        }

        // ---- fused side records -----------------------------------------------------------------
        NB = 0; bf_na.clear(); bf_nb.clear(); bf_ratio.clear();
        if (fused_ok) {
            const int64_t ntile = Npad / TW;
            ccode.assign((size_t)ntile * K * TW, 0);
            ccoef.assign((size_t)ntile * K * NC * TW, real(0));
            // boundary sides are numbered in position order; they only occur in the border group
            std::vector<int32_t> bside_base(std::max<int64_t>(Oend - Bstart, 0) + 1, 0);
            for (int64_t pc = Bstart; pc < Oend; ++pc) {
                const int64_t c = ipos[pc];
                int nb = 0;
                for (int k = 0; k < K; ++k) nb += other[c * K + k] == -1;
                bside_base[pc - Bstart + 1] = bside_base[pc - Bstart] + nb;
            }
            NB = bside_base[std::max<int64_t>(Oend - Bstart, 0)];
            if (NB >= (int64_t(1) << 28)) return fail("too many boundary sides");
            bf_na.assign(NB, 0); bf_nb.assign(NB, 0); bf_ratio.assign(NB, real(0));
#pragma omp parallel for num_threads(nthreads) schedule(static)
            for (int64_t pc = 0; pc < Npad; ++pc) {
                const int64_t c = ipos[pc];
                const int64_t tile = pc >> 5, lane = pc & 31;
                if (c < 0 || c >= No) { ccode[(size_t)(tile * K) * TW + lane] = HOLE; continue; }
                int64_t nb = pc >= Bstart ? bside_base[pc - Bstart] : 0;
                const real ia = inv_area ? inv_area[c] : real(1);
                for (int k = 0; k < K; ++k) {
                    const int64_t j = d.cell_face_idx[c * K + k];
                    const int sl = slot[c * K + k];
                    const int neg = d.cell_face_sign[c * K + k] < 0 ? 1 : 0;
                    const int32_t o = other[c * K + k];
                    int32_t code;
                    if (o >= 0) code = (pos[o] << 2) | (neg << 1) | sl;
                    else {
                        const real dg = fdist[2 * j + (1 - sl)], dk = fdist[2 * j + sl];
                        bf_na[nb] = node_track[d.face_node_idx[2 * j]];
                        bf_nb[nb] = node_track[d.face_node_idx[2 * j + 1]];
                        bf_ratio[nb] = dg / dk;
                        code = -(int32_t)(((nb << 2) | (neg << 1) | sl) + 1);
                        ++nb;
                    }
                    ccode[(size_t)(tile * K + k) * TW + lane] = code;
                    const real sg = neg ? real(-1) : real(1), vs = sl ? real(-1) : real(1);
                    real co[4] = {0, 0, 0, 0};
                    real L = fL[j];
                    if (inv_area) L = L * ia;
                    co[0] = sg * (fn[2 * j] * L);
                    co[1] = sg * (fn[2 * j + 1] * L);
                    if (NC == 4) {
                        const real d0 = fdist[2 * j], d1 = fdist[2 * j + 1], dd = d0 + d1;
                        co[2] = vs * (d0 / dd);
                        co[3] = (sg * vs) * (real(1) / (real(2) * dd * L));
                    }
                    for (int i = 0; i < NC; ++i) ccoef[(size_t)((tile * K + k) * NC + i) * TW + lane] = co[i];
                }
            }
        }
        return true;
    }
};

}  // namespace fvdbm
