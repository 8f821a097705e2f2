// plan.hpp -- host-side planning: reference layout (AoS, original numbering) -> device layout.
//
// Pure C++ (no CUDA) so it can be exercised on a CPU-only box through fvdbm_plan_create().
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
//       ccoef: NC reals per side: m = n*L (2), and for Lax-Wendroff alpha = d0/(d0+d1),
//              gamma = 1/(2 (d0+d1) L)   (src/containers.py:266-277 with varpi*L folded into m)
//   * boundary sides: the two tracked-node ids of the face and ratio d_ghost/d_known
//     (src/containers.py:280-287, utils/utils.py:153-154)
//   * tracked nodes (type != 0, or on a face with a ghost slot): compact ids, ring CSR with weights
//     w = 1/d, negative -> 0 (utils/utils.py:58-59), zero-weight entries dropped.
#pragma once
#include <cstdint>
#include <climits>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/fvdbm_b200.h"

namespace fvdbm {

constexpr int TW = 32;            // lanes of one AoSoA mini-tile
constexpr int PAD_TO = 512;       // group alignment = largest CTA tile
constexpr int32_t HOLE = INT32_MIN;
constexpr int BORDER_TILE = 256;  // cells per CTA of the border kernel

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

template <typename real>
struct Plan {
    int64_t N = 0, F = 0, P = 0, No = 0;
    int Q = 9, K = 3, M = 0, scheme = 0, NC = 2;
    int64_t Npad = 0, Bstart = 0, Oend = 0, Hstart = 0, D1start = 0;
    std::vector<int32_t> l2_list;               // positions of the level-2 cells (inside the tiled group)
    std::vector<int32_t> pos, ipos;
    bool fused_ok = true;
    std::string why_not;
    std::vector<int32_t> ccode, cface;          // side codes; side -> shared face record (face layout)
    std::vector<real> ccoef, fcoef;             // per-side coefficients (cell layout) / per-face records [NF][NC]
    int64_t NF = 0;                             // face records referenced by owned cells
    int64_t NB = 0;
    std::vector<int32_t> bf_na, bf_nb;
    std::vector<real> bf_ratio;
    int64_t NT = 0, NTpad = 0, NA = 0, NO = 0;  // tracked nodes; [0,NA) active (type != 0); [0,NO) active "orphans"
                                                // not on any owned boundary side (only those need k_nodes when fused)
    // border tiles (BORDER_TILE cells each, from Bstart): the tracked nodes each tile's boundary sides use
    std::vector<int32_t> bt_off, bt_nodes, bf_la, bf_lb;
    int64_t max_tile_nodes = 0;
    std::vector<int32_t> tn_orig, tn_type, node_track, tn_active, ring_off, ring_cell;
    std::vector<real> ring_w, tn_pdf, tn_rho, tn_vel;
    // the same rings as a fixed-width table [NA][MR] (zero weight = unused slot): lets the node kernel index
    // its ring entries directly instead of first loading CSR offsets (one dependent memory level less)
    std::vector<int32_t> ring_fcell;
    std::vector<real> ring_fw;
    int64_t MR = 0;
    std::vector<int32_t> s_cface, s_csign, s_fcell, s_fnode;
    std::string error;

    bool fail(const std::string& msg) { error = msg; return false; }

    // ---- temporal blocking (two iterations per pass) -------------------------------------------------
    // Tiles of T2 consecutive positions over [0, D1start) (cells at level >= 2).  Entry list of a tile:
    //   [ own (T2, implicit positions) | ring1 = face neighbours of own | ring2 = face neighbours of ring1 ]
    // The kernel stages the time-t populations of all entries in shared memory, advances own+ring1 to
    // t+1 there (ring1 redundantly: neighbouring tiles do the same, identically), then own to t+2.
    // t2_lnbr[entry][k] = (local id of the side's other cell << 2) | (sign<0)<<1 | slot, 16 bit.
    static constexpr int T2 = 256;
    std::vector<int32_t> t2_off, t2_n1, t2_pos;     // per tile: offset into t2_pos, #ring1; ring positions
    std::vector<int64_t> t2_loff;                   // per tile: offset (in entries) into t2_lnbr
    std::vector<uint16_t> t2_lnbr;
    int64_t t2_tiles = 0, t2_max_entries = 0, t2_max_n01 = 0;
    bool t2_ok = false;

    void build_temporal_tiles(const fvdbm_desc& d, const std::vector<int32_t>& other, const std::vector<uint8_t>& slot) {
        t2_ok = false;
        t2_tiles = D1start / T2;
        if (t2_tiles == 0) return;
        t2_off.assign(t2_tiles + 1, 0); t2_n1.assign(t2_tiles, 0); t2_loff.assign(t2_tiles + 1, 0);
        t2_pos.clear(); t2_lnbr.clear();
        std::vector<int32_t> stamp(Npad, -1), lid(Npad, 0);
        std::vector<int32_t> ring;
        for (int64_t t = 0; t < t2_tiles; ++t) {
            const int64_t t0 = t * T2;
            ring.clear();
            for (int e = 0; e < T2; ++e) { stamp[t0 + e] = (int32_t)t; lid[t0 + e] = e; }
            // collect a ring (unstamped face neighbours of the given entries), sort it by position so that
            // consecutive threads gather consecutive positions (shared 32 B sectors), then assign local ids
            auto grow = [&](int64_t from_begin, int64_t from_end, bool from_own) {
                const size_t first = ring.size();
                for (int64_t e = from_begin; e < from_end; ++e) {
                    const int64_t pc = from_own ? t0 + e : ring[e];
                    const int64_t c = ipos[pc];
                    if (c < 0 || c >= No) continue;
                    for (int k = 0; k < K; ++k) {
                        const int32_t o = other[c * K + k];
                        if (o < 0) continue;                    // cannot happen for level >= 1 cells
                        const int64_t np = pos[o];
                        if (stamp[np] != (int32_t)t) { stamp[np] = (int32_t)t; ring.push_back((int32_t)np); }
                    }
                }
                std::sort(ring.begin() + first, ring.end());
                for (size_t i = first; i < ring.size(); ++i) lid[ring[i]] = (int32_t)(T2 + i);
            };
            grow(0, T2, true);
            const int64_t n1 = (int64_t)ring.size();
            grow(0, n1, false);
            const int64_t n2 = (int64_t)ring.size() - n1;
            t2_n1[t] = (int32_t)n1;
            t2_off[t + 1] = t2_off[t] + (int32_t)ring.size();
            t2_pos.insert(t2_pos.end(), ring.begin(), ring.end());
            t2_loff[t + 1] = t2_loff[t] + (T2 + n1);
            for (int64_t e = 0; e < T2 + n1; ++e) {
                const int64_t pc = e < T2 ? t0 + e : ring[e - T2];
                const int64_t c = ipos[pc];
                for (int k = 0; k < K; ++k) {
                    uint16_t v = 0xFFFF;                        // hole (padding position): skipped by the kernel
                    if (c >= 0 && c < No) {
                        const int32_t o = other[c * K + k];
                        const int neg = d.cell_face_sign[c * K + k] < 0 ? 1 : 0;
                        v = (uint16_t)((lid[pos[o]] << 2) | (neg << 1) | slot[c * K + k]);
                    }
                    t2_lnbr.push_back(v);
                }
            }
            t2_max_entries = std::max<int64_t>(t2_max_entries, T2 + n1 + n2);
            t2_max_n01 = std::max<int64_t>(t2_max_n01, T2 + n1);
        }
        t2_ok = t2_max_entries < (1 << 14);
    }

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
        const real* fdist = static_cast<const real*>(d.face_dists);
        const real* fn = static_cast<const real*>(d.face_n);
        const real* fL = static_cast<const real*>(d.face_L);

        // ---- range checks -------------------------------------------------------------------
        for (int64_t i = 0; i < No * K; ++i)      // halo cells carry no sides (never updated)
            if (d.cell_face_idx[i] < 0 || d.cell_face_idx[i] >= F)
                return fail("cell_face_idx out of range (ragged / -1 padded cells are not supported)");
        for (int64_t i = 0; i < F * 2; ++i)
            if (d.face_cell_idx[i] < -1 || d.face_cell_idx[i] >= N) return fail("face_cell_idx out of range");

        // ---- side analysis in original numbering -----------------------------------------------
        // other[i*K+k] = neighbour cell (>=0), -1 boundary, -2 inconsistent
        std::vector<int32_t> other(N * K, -2);
        std::vector<uint8_t> slot(N * K, 0);
        fused_ok = true;
        for (int64_t c = 0; c < No && fused_ok; ++c)
            for (int k = 0; k < K; ++k) {
                int64_t j = d.cell_face_idx[c * K + k];
                int32_t a = d.face_cell_idx[2 * j], b = d.face_cell_idx[2 * j + 1];
                int32_t s = d.cell_face_sign[c * K + k];
                if (s != 1 && s != -1) { fused_ok = false; why_not = "cell_face_sign not +-1"; break; }
                if (a == c && b != c) { slot[c * K + k] = 0; other[c * K + k] = b; }
                else if (b == c && a != c) { slot[c * K + k] = 1; other[c * K + k] = a; }
                else { fused_ok = false; why_not = "a cell lists a face whose stencil does not contain it exactly once"; break; }
                if (other[c * K + k] == -1) {
                    int32_t na = d.face_node_idx[2 * j], nb = d.face_node_idx[2 * j + 1];
                    if (na < 0 || na >= P || nb < 0 || nb >= P) { fused_ok = false; why_not = "boundary face without valid nodes"; break; }
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
            // level = face-graph distance to the nearest border cell, capped at 3.  Positions are grouped
            // [level>=2 | level 1 | level 0 = border]: the single-step schedule uses interior = [0,Bstart);
            // the two-step temporal schedule tiles [0,D1start) and runs thin single-step passes over
            // [D1start,end) plus the explicit list of level-2 cells (l2_list), which stay in locality order
            // inside the tiled group so that no tile is a thin strip with a huge ring.
            std::vector<uint8_t> lvl(N, 3);
            for (int64_t c = 0; c < No; ++c)
                for (int k = 0; k < K; ++k) {
                    int32_t o = other[c * K + k];
                    if (o == -1 || o >= No) lvl[c] = 0;
                }
            for (int L = 0; L < 2; ++L)
                for (int64_t c = 0; c < No; ++c)
                    if (lvl[c] == L)
                        for (int k = 0; k < K; ++k) {
                            int32_t o = other[c * K + k];
                            if (o >= 0 && o < No && lvl[o] > L + 1) lvl[o] = (uint8_t)(L + 1);
                        }
            l2_list.clear();
            for (int64_t r = 0; r < No; ++r)
                if (lvl[order[r]] >= 2) {
                    if (lvl[order[r]] == 2) l2_list.push_back((int32_t)p);
                    pos[order[r]] = (int32_t)p++;
                }
            D1start = round_up(p, PAD_TO); p = D1start;
            for (int64_t r = 0; r < No; ++r) if (lvl[order[r]] == 1) pos[order[r]] = (int32_t)p++;
            Bstart = round_up(p, PAD_TO); p = Bstart;
            for (int64_t r = 0; r < No; ++r) if (lvl[order[r]] == 0) pos[order[r]] = (int32_t)p++;
            Oend = p; Hstart = round_up(p, PAD_TO); p = Hstart;
            for (int64_t r = No; r < N; ++r) pos[order[r]] = (int32_t)p++;
        } else {
            Bstart = 0;
            for (int64_t r = 0; r < N; ++r) pos[order[r]] = (int32_t)p++;
            Oend = p; Hstart = round_up(p, PAD_TO);
        }
        Npad = round_up(std::max<int64_t>(p, 1), PAD_TO);
        ipos.assign(Npad, -1);
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
        // order: active orphans | active nodes on an owned boundary side (evaluated by the border kernel,
        // tile by tile) | inactive tracked nodes (keep their stored PDFs)
        std::vector<uint8_t> on_side(P, 0);
        if (fused_ok)
            for (int64_t c = 0; c < No; ++c)
                for (int k = 0; k < K; ++k)
                    if (other[c * K + k] == -1) {
                        const int64_t j = d.cell_face_idx[c * K + k];
                        on_side[d.face_node_idx[2 * j]] = 1; on_side[d.face_node_idx[2 * j + 1]] = 1;
                    }
        tn_orig.clear(); tn_type.clear(); NA = 0; NO = 0;
        for (int pass = 0; pass < 3; ++pass)
            for (int64_t n = 0; n < P; ++n) {
                if (!want[n]) continue;
                const bool active = d.node_type[n] != 0;
                const int cls = !active ? 2 : (on_side[n] ? 1 : 0);
                if (cls != pass) continue;
                node_track[n] = (int32_t)tn_orig.size();
                tn_orig.push_back((int32_t)n);
                tn_type.push_back(d.node_type[n]);
                if (active) ++NA;
                if (cls == 0) ++NO;
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
        for (int64_t c = 0; c < No; ++c)
            for (int k = 0; k < K; ++k) {
                s_cface[(size_t)k * Npad + pos[c]] = d.cell_face_idx[c * K + k];
                s_csign[(size_t)k * Npad + pos[c]] = d.cell_face_sign[c * K + k];
            }
        s_fcell.resize(2 * F); s_fnode.resize(2 * F);
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
            for (int64_t q = 0; q < Npad; ++q)
                if (ipos[q] < 0 || ipos[q] >= No) ccode[(size_t)(q >> 5) * K * TW + (q & 31)] = HOLE;
            // two interchangeable coefficient layouts:
            //   cell layout: ccoef[tile][k*NC+i][lane]   (perfectly coalesced, interior faces stored twice)
            //   face layout: fcoef[rec][NC] + cface[tile][k][lane] -> rec (one 16 B record per face, shared by
            //                both cells; records numbered by first touch in position order for locality)
            cface.assign((size_t)ntile * K * TW, 0);
            const int64_t nbt = (round_up(Oend, PAD_TO) - Bstart) / BORDER_TILE;
            bt_off.assign(nbt + 1, 0); bt_nodes.clear(); bf_la.clear(); bf_lb.clear(); max_tile_nodes = 0;
            std::vector<int32_t> slot_of(std::max<int64_t>(NT, 1), -1), stamp(std::max<int64_t>(NT, 1), -1);
            int64_t cur_bt = -1;
            auto tile_slot = [&](int64_t bt, int32_t t) -> int32_t {   // slot of tracked node t in border tile bt
                while (cur_bt < bt) { ++cur_bt; bt_off[cur_bt] = (int32_t)bt_nodes.size(); }
                if (stamp[t] != (int32_t)bt) { stamp[t] = (int32_t)bt; slot_of[t] = (int32_t)(bt_nodes.size() - bt_off[bt]); bt_nodes.push_back(t); }
                return slot_of[t];
            };
            std::vector<int32_t> rec_of(F, -1);
            fcoef.clear();
            NF = 0;
            for (int64_t pc = 0; pc < Npad; ++pc) {
                const int64_t c = ipos[pc];
                if (c < 0 || c >= No) continue;
                const int64_t tile = pc >> 5, lane = pc & 31;
                for (int k = 0; k < K; ++k) {
                    const int64_t j = d.cell_face_idx[c * K + k];
                    const int sl = slot[c * K + k];
                    const int neg = d.cell_face_sign[c * K + k] < 0 ? 1 : 0;
                    const int32_t o = other[c * K + k];
                    int32_t code;
                    if (o >= 0) code = (pos[o] << 2) | (neg << 1) | sl;
                    else {
                        const real dg = fdist[2 * j + (1 - sl)], dk = fdist[2 * j + sl];
                        const int32_t ta = node_track[d.face_node_idx[2 * j]], tb = node_track[d.face_node_idx[2 * j + 1]];
                        bf_na.push_back(ta);
                        bf_nb.push_back(tb);
                        bf_ratio.push_back(dg / dk);
                        const int64_t bt = (pc - Bstart) / BORDER_TILE;       // boundary cells live in the border group
                        bf_la.push_back(tile_slot(bt, ta));
                        bf_lb.push_back(tile_slot(bt, tb));
                        code = -(int32_t)(((NB << 2) | (neg << 1) | sl) + 1);
                        ++NB;
                    }
                    ccode[(size_t)(tile * K + k) * TW + lane] = code;
                    real co[4] = {0, 0, 0, 0};
                    const real L = fL[j];
                    co[0] = fn[2 * j] * L;
                    co[1] = fn[2 * j + 1] * L;
                    if (NC == 4) {
                        const real d0 = fdist[2 * j], d1 = fdist[2 * j + 1], dd = d0 + d1;
                        co[2] = d0 / dd;
                        co[3] = real(1) / (real(2) * dd * L);
                    }
                    for (int i = 0; i < NC; ++i) ccoef[(size_t)((tile * K + k) * NC + i) * TW + lane] = co[i];
                    if (rec_of[j] < 0) {
                        rec_of[j] = (int32_t)NF++;
                        for (int i = 0; i < NC; ++i) fcoef.push_back(co[i]);
                    }
                    cface[(size_t)(tile * K + k) * TW + lane] = rec_of[j];
                }
            }
            while (cur_bt < nbt) { ++cur_bt; bt_off[cur_bt] = (int32_t)bt_nodes.size(); }
            for (int64_t t = 0; t < nbt; ++t) max_tile_nodes = std::max<int64_t>(max_tile_nodes, bt_off[t + 1] - bt_off[t]);
            if (NB >= (int64_t(1) << 28)) return fail("too many boundary sides");
            if (!has_halo) build_temporal_tiles(d, other, slot);
        }
        return true;
    }
};

}  // namespace fvdbm
