// Native Mesher-equivalent (host only, OpenMP): the producer of the hot path's static arrays, SURVEY.md 8(f)-1.
//
// Computes what the reference's Mesher.calc_mesh_properties does with per-element Python loops and dict lookups
// (/root/reference/src/mesher.py:63-316, ~130 us per cell) for triangle meshes of 10^7-10^8 cells in about a second:
//   cell centres            mesher.py:113-120        face -> first two cells, slot order   mesher.py:197-266
//   cell -> face ids        mesher.py:122-138        ghost distances                       mesher.py:268-283
//   outward normals / signs mesher.py:140-169        node -> ring cells and distances      mesher.py:286-316
//   boundary-face flip      mesher.py:80-110         cell-centre stencil (cc_* fluxes)     mesher.py:506-558
//   face centres/normals/L  mesher.py:172-195
//
// No sort and no hash table: two CSR tables keyed by the canonical vertex id -- vertex -> corner slots (3c+k), which IS the
// node ring the path needs anyway, and min-vertex -> faces -- are filled with atomic cursors and then ordered segment by
// segment (segments hold ~6 / ~3 entries), so the result does not depend on the thread schedule.  A cell edge finds its
// face by scanning the ~3 faces of its smaller vertex; a face finds its cells by scanning the ring of its smaller vertex.
//
// Contract (tests/test_host_logic.py): every integer output equals the NumPy implementation in mesher.py (which is
// itself bit-identical to the reference Mesher on the integer arrays) and every float output equals it BIT FOR BIT: each
// expression below spells the same IEEE operations in the same order as the NumPy code (no contraction: the host
// compiler has no FMA target and is passed -ffp-contract=off).  `alias` (periodic identification, an extension) maps a
// point id to its canonical id; connectivity uses canonical ids, geometry a cell's own vertices.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace fvdbm {

struct MeshIn {
    int64_t N = 0, F = 0, P = 0;
    const double* points = nullptr;      // (P,2)
    const int32_t* cells = nullptr;      // (N,3) counter-clockwise
    const int32_t* faces = nullptr;      // (F,2)
    const int32_t* alias = nullptr;      // (P) or null
};

struct MeshOut {                         // all preallocated by the caller
    int32_t M = 0;                       // ring width (mesh_ring_width)
    double* cell_centers = nullptr;              // (N,2)
    int64_t* cell_face_indices = nullptr;        // (N,3)
    double* cell_face_normals = nullptr;         // (N,3,2)
    int32_t* cell_face_normal_signs = nullptr;   // (N,3)
    int32_t* faces = nullptr;                    // (F,2)  node order after the periodic re-expression and the boundary flip
    double* face_centers = nullptr;              // (F,2)
    double* face_normals = nullptr;              // (F,2)
    double* face_lengths = nullptr;              // (F)
    int64_t* face_cell_indices = nullptr;        // (F,2)
    double* face_cell_center_distances = nullptr;// (F,2)
    double* stencil_norms = nullptr;             // (F,2)
    double* cc_stencil_dist = nullptr;           // (F,2)
    double* face_stencil_angles = nullptr;       // (F)
    int64_t* point_cell_indices = nullptr;       // (P,M)  -1 padded
    double* point_cell_center_distances = nullptr;// (P,M) -1 padded
};

namespace meshdetail {

inline int threads() {
    int n = 1;
#if defined(_OPENMP)
    n = omp_get_max_threads();
#endif
    if (const char* e = getenv("FVDBM_PLAN_THREADS")) n = atoi(e);
    return n < 1 ? 1 : n;
}

// CSR keyed by vertex: start[v] .. start[v+1] lists the slots s (ascending) with key(s) == v; key(s) < 0 skips the slot.
template <typename KeyFn>
bool build_csr(int64_t nslots, int64_t P, KeyFn key, std::vector<int64_t>& start, std::vector<int32_t>& items, int nt) {
    std::vector<int32_t> cnt((size_t)P, 0);
    int bad = 0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : bad)
    for (int64_t s = 0; s < nslots; ++s) {
        const int64_t v = key(s);
        if (v >= P) { bad |= 1; continue; }
        if (v < 0) { bad |= (v < -1); continue; }
#pragma omp atomic
        cnt[(size_t)v]++;
    }
    if (bad) return false;
    start.assign((size_t)P + 1, 0);
    for (int64_t v = 0; v < P; ++v) start[(size_t)v + 1] = start[(size_t)v] + cnt[(size_t)v];
    items.resize((size_t)start[(size_t)P]);
    std::fill(cnt.begin(), cnt.end(), 0);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t s = 0; s < nslots; ++s) {
        const int64_t v = key(s);
        if (v < 0) continue;
        int32_t at;
#pragma omp atomic capture
        at = cnt[(size_t)v]++;
        items[(size_t)(start[(size_t)v] + at)] = (int32_t)s;
    }
#pragma omp parallel for num_threads(nt) schedule(static, 4096)
    for (int64_t v = 0; v < P; ++v) {                 // tiny segments: insertion sort
        int32_t* a = items.data() + start[(size_t)v];
        const int64_t n = start[(size_t)v + 1] - start[(size_t)v];
        for (int64_t i = 1; i < n; ++i) {
            const int32_t x = a[i];
            int64_t j = i - 1;
            while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
            a[j + 1] = x;
        }
    }
    return true;
}

struct V2 { double x, y; };
inline V2 pt(const double* p, int64_t i) { return {p[2 * i], p[2 * i + 1]}; }
inline double dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline V2 sub(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 mid(V2 a, V2 b) { return {(a.x + b.x) / 2.0, (a.y + b.y) / 2.0}; }
inline double len(V2 a) { return std::sqrt(a.x * a.x + a.y * a.y); }
inline V2 normal(V2 p0, V2 p1) {                      // unit left normal of p0 -> p1 (utils/utils.py:162-173)
    const V2 t = sub(p1, p0);
    const V2 n = {-t.y, t.x};
    const double l = len(n);
    return {n.x / l, n.y / l};
}

}  // namespace meshdetail

// Widest node ring (cells around a canonical vertex) or -1 for an out-of-range vertex id.
inline int64_t mesh_ring_width(const int32_t* cells, const int32_t* alias, int64_t N, int64_t P) {
    const int nt = meshdetail::threads();
    (void)nt;
    std::vector<int32_t> cnt((size_t)P, 0);
    int bad = 0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : bad)
    for (int64_t s = 0; s < 3 * N; ++s) {
        int64_t v = cells[s];
        if (v < 0 || v >= P) { bad |= 1; continue; }
        if (alias) v = alias[v];
        if (v < 0 || v >= P) { bad |= 1; continue; }
#pragma omp atomic
        cnt[(size_t)v]++;
    }
    if (bad) return -1;
    int32_t m = 0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(max : m)
    for (int64_t v = 0; v < P; ++v) m = cnt[(size_t)v] > m ? cnt[(size_t)v] : m;
    return m;
}

// Unique undirected edges of a K-gon mesh (K = 3 or 4), one row per face, ordered by (min, max) canonical vertex id; a row
// keeps the point ids of the FIRST half-edge (cell-major order) that produced the key -- meshgen.unique_edges' contract,
// which np.unique(key, return_index=True) gives after an O(n log n) sort.  `faces_out` holds up to N*K rows; returns the
// number of faces or -1 for an id out of range.
inline int64_t mesh_unique_edges(const int32_t* cells, int64_t N, int K, int64_t P, const int32_t* alias, int32_t* faces_out) {
    using namespace meshdetail;
    if (N < 0 || P < 0 || K < 3 || K > 4 || N * K > INT32_MAX) return -1;
    const int nt = threads();
    auto canon = [alias](int64_t v) -> int64_t { return alias ? alias[v] : v; };
    int bad = 0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : bad)
    for (int64_t s = 0; s < N * K; ++s) bad |= (cells[s] < 0 || cells[s] >= P || (alias && (alias[cells[s]] < 0 || alias[cells[s]] >= P)));
    if (bad) return -1;
    auto other = [&](int64_t h) -> int64_t { const int64_t c = h / K, k = h - c * K; return cells[c * K + (k + 1) % K]; };
    std::vector<int64_t> start;
    std::vector<int32_t> items;
    if (!build_csr(N * K, P, [&](int64_t h) { return std::min(canon(cells[h]), canon(other(h))); }, start, items, nt)) return -1;
    // per smaller vertex: order its half-edges by (larger vertex, half-edge id), keep the first of every run
    std::vector<int64_t> nuniq((size_t)P + 1, 0);
#pragma omp parallel for num_threads(nt) schedule(static, 4096)
    for (int64_t v = 0; v < P; ++v) {
        int32_t* a = items.data() + start[(size_t)v];
        const int64_t n = start[(size_t)v + 1] - start[(size_t)v];
        auto hi = [&](int32_t h) -> int64_t { return std::max(canon(cells[h]), canon(other(h))); };
        for (int64_t i = 1; i < n; ++i) {                   // stable insertion sort by hi (segments arrive ascending in h)
            const int32_t x = a[i];
            const int64_t hx = hi(x);
            int64_t j = i - 1;
            while (j >= 0 && hi(a[j]) > hx) { a[j + 1] = a[j]; --j; }
            a[j + 1] = x;
        }
        int64_t u = 0;
        for (int64_t i = 0; i < n; ++i) if (i == 0 || hi(a[i]) != hi(a[i - 1])) ++u;
        nuniq[(size_t)v + 1] = u;
    }
    for (int64_t v = 0; v < P; ++v) nuniq[(size_t)v + 1] += nuniq[(size_t)v];
#pragma omp parallel for num_threads(nt) schedule(static, 4096)
    for (int64_t v = 0; v < P; ++v) {
        const int32_t* a = items.data() + start[(size_t)v];
        const int64_t n = start[(size_t)v + 1] - start[(size_t)v];
        int64_t w = nuniq[(size_t)v];
        int64_t last = -1;
        for (int64_t i = 0; i < n; ++i) {
            const int64_t h = a[i], hv = std::max(canon(cells[h]), canon(other(h)));
            if (i == 0 || hv != last) { faces_out[2 * w] = cells[h]; faces_out[2 * w + 1] = (int32_t)other(h); ++w; }
            last = hv;
        }
    }
    return nuniq[(size_t)P];
}

// 0 ok; -1 bad argument / id out of range; -2 a cell edge is missing from `faces` (the reference raises KeyError).
inline int mesh_properties(const MeshIn& in, MeshOut& o, std::string& err) {
    using namespace meshdetail;
    const int64_t N = in.N, F = in.F, P = in.P;
    const int32_t* alias = in.alias;
    const double* pts = in.points;
    const int32_t* cells = in.cells;
    if (N < 0 || F < 0 || P < 0 || 3 * N > INT32_MAX || F > INT32_MAX) { err = "mesh too large for 32-bit slot ids"; return -1; }
    const int nt = threads();
    auto canon = [alias](int64_t v) -> int64_t { return alias ? alias[v] : v; };

    {   // ids in range (everything below indexes without checks)
        int bad = 0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : bad)
        for (int64_t s = 0; s < 3 * N; ++s) bad |= (cells[s] < 0 || cells[s] >= P);
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : bad)
        for (int64_t s = 0; s < 2 * F; ++s) bad |= (in.faces[s] < 0 || in.faces[s] >= P);
        if (alias)
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : bad)
            for (int64_t v = 0; v < P; ++v) bad |= (alias[v] < 0 || alias[v] >= P);
        if (bad) { err = "point id out of range"; return -1; }
    }

    // ---- cell centres (mesher.py:113-120) -------------------------------------------------------
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t c = 0; c < N; ++c) {
        const V2 a = pt(pts, cells[3 * c]), b = pt(pts, cells[3 * c + 1]), d = pt(pts, cells[3 * c + 2]);
        o.cell_centers[2 * c] = ((a.x + b.x) + d.x) / 3.0;
        o.cell_centers[2 * c + 1] = ((a.y + b.y) + d.y) / 3.0;
    }
    auto cc = [&o](int64_t c) -> V2 { return {o.cell_centers[2 * c], o.cell_centers[2 * c + 1]}; };

    // ---- the two tables -------------------------------------------------------------------------
    std::vector<int64_t> rstart, bstart;
    std::vector<int32_t> ring, bucket;
    if (!build_csr(3 * N, P, [&](int64_t s) { return canon(cells[s]); }, rstart, ring, nt) ||
        !build_csr(F, P, [&](int64_t f) { return std::min(canon(in.faces[2 * f]), canon(in.faces[2 * f + 1])); }, bstart, bucket, nt)) {
        err = "point id out of range";
        return -1;
    }
    if (o.M < 0) { err = "bad ring width"; return -1; }

    // ---- cell edge -> face: the LAST face carrying the key wins (dict comprehension, mesher.py:129) ----
    int missing = 0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : missing)
    for (int64_t h = 0; h < 3 * N; ++h) {
        const int64_t c = h / 3, k = h - 3 * c;
        const int64_t a = canon(cells[h]), b = canon(cells[3 * c + (k + 1) % 3]);
        const int64_t lo = std::min(a, b), hi = std::max(a, b);
        int64_t found = -1;
        for (int64_t j = bstart[(size_t)lo]; j < bstart[(size_t)lo + 1]; ++j) {
            const int64_t f = bucket[(size_t)j];
            if (std::max(canon(in.faces[2 * f]), canon(in.faces[2 * f + 1])) == hi) found = f;
        }
        if (found < 0) { missing |= 1; found = 0; }
        o.cell_face_indices[h] = found;
    }
    if (missing) { err = "a cell edge is missing from `faces`"; return -2; }

    // ---- per face: adjacent cells, orientation, geometry, stencils ------------------------------------
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t f = 0; f < F; ++f) {
        int64_t f0 = in.faces[2 * f], f1 = in.faces[2 * f + 1];
        const int64_t a = canon(f0), b = canon(f1);
        const int64_t lo = std::min(a, b), hi = std::max(a, b);
        // first two half-edges (ascending cell, then corner) with this key (mesher.py:206-220)
        int cnt = 0;
        int64_t he[2] = {0, 0}, prev = -1;
        for (int64_t j = rstart[(size_t)lo]; j < rstart[(size_t)lo + 1] && cnt < 2; ++j) {
            const int64_t c = ring[(size_t)j] / 3;
            if (c == prev) continue;                       // degenerate cell listing the vertex twice
            prev = c;
            for (int k = 0; k < 3 && cnt < 2; ++k) {
                const int64_t x = canon(cells[3 * c + k]), y = canon(cells[3 * c + (k + 1) % 3]);
                if (std::min(x, y) == lo && std::max(x, y) == hi) he[cnt++] = 3 * c + k;
            }
        }
        const int64_t c0 = cnt > 0 ? he[0] / 3 : -1, c1 = cnt > 1 ? he[1] / 3 : -1;
        auto ha = [&](int64_t h) -> int64_t { return cells[h]; };
        auto hb = [&](int64_t h) -> int64_t { return cells[3 * (h / 3) + (h % 3 + 1) % 3]; };
        if (alias && cnt > 0) {        // express the face with the point ids of its first half-edge (unwrapped coordinates)
            const int64_t own0 = ha(he[0]), own1 = hb(he[0]);
            const bool same = canon(own0) == a;
            f0 = same ? own0 : own1;
            f1 = same ? own1 : own0;
        }
        V2 p0 = pt(pts, f0), p1 = pt(pts, f1);
        if (cnt == 1) {                // boundary face: stored node order must give an outward normal (mesher.py:80-110)
            const V2 nb = normal(p0, p1), mb = mid(p0, p1);
            if (dot(nb, sub(cc(c0), mb)) >= 0) { std::swap(f0, f1); std::swap(p0, p1); }
        }
        o.faces[2 * f] = (int32_t)f0;
        o.faces[2 * f + 1] = (int32_t)f1;
        // centres / normals / lengths (mesher.py:172-195)
        const V2 fc = mid(p0, p1), fn = normal(p0, p1), t = sub(p1, p0);
        o.face_centers[2 * f] = fc.x; o.face_centers[2 * f + 1] = fc.y;
        o.face_normals[2 * f] = fn.x; o.face_normals[2 * f + 1] = fn.y;
        o.face_lengths[f] = len(t);
        // stencil slots along the face normal with projected distances (mesher.py:222-266); every cell measures from
        // the midpoint of its OWN copy of the edge (periodic-safe)
        const V2 cc0 = cc(std::max<int64_t>(c0, 0)), cc1 = cc(std::max<int64_t>(c1, 0));
        V2 hm0 = {0, 0}, hm1 = {0, 0};
        if (cnt > 0) hm0 = mid(pt(pts, ha(he[0])), pt(pts, hb(he[0])));
        if (cnt > 1) hm1 = mid(pt(pts, ha(he[1])), pt(pts, hb(he[1])));
        const double d0 = cnt > 0 ? dot(fn, sub(cc0, hm0)) : -1.0;
        const double d1 = cnt > 1 ? dot(fn, sub(cc1, hm1)) : -1.0;
        const bool interior = cnt > 1;
        const bool swp = interior && !(d0 < d1);
        const int64_t s0 = swp ? c1 : c0, s1 = swp ? c0 : c1;
        o.face_cell_indices[2 * f] = s0;
        o.face_cell_indices[2 * f + 1] = s1;
        double e0 = std::fabs(swp ? d1 : d0), e1 = std::fabs(swp ? d0 : d1);
        if (s0 == -1) e0 = e1;         // ghost distance = distance of the real cell (mesher.py:268-283)
        if (s1 == -1) e1 = e0;
        o.face_cell_center_distances[2 * f] = e0;
        o.face_cell_center_distances[2 * f + 1] = e1;
        // cell-centre stencil (mesher.py:506-558)
        V2 v = interior ? sub(cc1, cc0) : sub(fc, cc0);
        if (cnt == 0) v = {0.0, 0.0};
        const double vn = len(v);
        V2 sn = vn > 0 ? V2{v.x / vn, v.y / vn} : v;
        const double fl = len(fn);
        const V2 fnu = {fn.x / fl, fn.y / fl};
        const double sl = len(sn) + 1e-14;
        const V2 snu = {sn.x / sl, sn.y / sl};
        if (dot(fnu, snu) < 0) sn = {-sn.x, -sn.y};
        o.stencil_norms[2 * f] = sn.x; o.stencil_norms[2 * f + 1] = sn.y;
        const double q0 = cnt > 0 ? dot(sn, sub(cc0, hm0)) : -1.0;
        const double q1 = cnt > 1 ? dot(sn, sub(cc1, hm1)) : -1.0;
        double g0 = std::fabs(swp ? q1 : q0), g1 = std::fabs(swp ? q0 : q1);
        if (s0 == -1) g0 = g1;
        if (s1 == -1) g1 = g0;
        o.cc_stencil_dist[2 * f] = g0;
        o.cc_stencil_dist[2 * f + 1] = g1;
        const double sl2 = len(sn);
        const V2 snn = {sn.x / sl2, sn.y / sl2};
        double ca = dot(fnu, snn);
        ca = ca < -1.0 ? -1.0 : (ca > 1.0 ? 1.0 : ca);    // NaN (face without a cell) passes through, like np.clip
        o.face_stencil_angles[f] = std::acos(ca);
    }

    // ---- per cell corner: outward normal and its sign against the face normal (mesher.py:140-169) ----------
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t h = 0; h < 3 * N; ++h) {
        const int64_t c = h / 3, k = h - 3 * c;
        const V2 p0 = pt(pts, cells[h]), p1 = pt(pts, cells[3 * c + (k + 1) % 3]);
        V2 hn = normal(p0, p1);
        const bool out = dot(hn, sub(cc(c), mid(p0, p1))) < 0;
        if (!out) hn = {-hn.x, -hn.y};
        o.cell_face_normals[2 * h] = hn.x;
        o.cell_face_normals[2 * h + 1] = hn.y;
        const int64_t f = o.cell_face_indices[h];
        const double s = dot(hn, V2{o.face_normals[2 * f], o.face_normals[2 * f + 1]});
        o.cell_face_normal_signs[h] = s > 0 ? 1 : (s < 0 ? -1 : 0);
    }

    // ---- node -> ring cells (ascending), padded with -1; distance node - centroid (mesher.py:286-316) --------
    const int64_t M = o.M;
    int wide = 0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(| : wide)
    for (int64_t v = 0; v < P; ++v) {
        const int64_t b = rstart[(size_t)v], n = rstart[(size_t)v + 1] - b;
        if (n > M) { wide |= 1; continue; }
        for (int64_t j = 0; j < M; ++j) {
            if (j < n) {
                const int64_t s = ring[(size_t)(b + j)], c = s / 3;
                const V2 own = sub(pt(pts, cells[s]), cc(c));
                o.point_cell_indices[v * M + j] = c;
                o.point_cell_center_distances[v * M + j] = len(own);
            } else {
                o.point_cell_indices[v * M + j] = -1;
                o.point_cell_center_distances[v * M + j] = -1.0;
            }
        }
    }
    if (wide) { err = "ring wider than M (call fvdbm_mesh_ring_width first)"; return -1; }
    return 0;
}

}  // namespace fvdbm
