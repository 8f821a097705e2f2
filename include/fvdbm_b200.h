/* fvdbm_b200.h -- C ABI of the B200-native FVDBM time-step engine (libfvdbm_b200.so).
 *
 * The reference (MADALERAU/FVDBM-JAX) has no FFI layer: its boundary for this path is the Python
 * class `Environment` (src/environment.py:10-68) whose jitted `step()` (:55-65) runs five vmapped
 * stages over the `Cells` / `Faces` / `Nodes` pytrees (src/containers.py).  This header is what a
 * JAX-FFI custom call / ctypes binding for that method binds to: plain
 * pointers and sizes, no torch / jax / C++ types.
 *
 * Array arguments use the REFERENCE layout (row-major AoS, original element numbering):
 *   cell_face_idx  [N*K] i32   Cells.face_indices           src/containers.py:60, mesher.py:634
 *   cell_face_sign [N*K] i32   Cells.face_normals (+1/-1)   src/containers.py:62, mesher.py:635
 *   face_cell_idx  [F*2] i32   Faces.stencil_cells_index    src/containers.py:151 (-1 = ghost)
 *   face_dists     [F*2] real  Faces.stencil_dists          src/containers.py:152
 *   face_node_idx  [F*2] i32   Faces.nodes_index            src/containers.py:150
 *   face_n         [F*2] real  Faces.n                      src/containers.py:153
 *   face_L         [F]   real  Faces.L                      src/containers.py:154
 *   node_type      [P]   i32   Nodes.type (0 none,1 vel,2 rho) src/containers.py:305, mesher.py:716,739
 *   node_cell_idx  [P*M] i32   Nodes.cells_index (-1 pad)   src/containers.py:306
 *   node_cell_dist [P*M] real  Nodes.cell_dists  (-1 pad)   src/containers.py:307
 *   cell_pdf [N*Q], node_pdf [P*Q], node_rho [P], node_vel [P*2]  initial dynamic state
 * node_type values other than 0/1/2 are rejected (the reference updates only types 1 and 2).
 * `real` is float when dtype==32 and double when dtype==64.
 *
 * Threading: one handle = one device + one CUDA stream; calls on a handle are not thread-safe,
 * distinct handles are independent.  fvdbm_step only enqueues work; fvdbm_get / fvdbm_sync block;
 * fvdbm_set returns when the source buffer may be reused (the import is stream-ordered).
 * Errors: every int-returning call gives 0 on success, <0 on failure; fvdbm_last_error(h) (or
 * fvdbm_last_error(NULL) for create-time failures) returns the message.
 */
#ifndef FVDBM_B200_H
#define FVDBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVDBM_ABI_VERSION 2   /* 2: fvdbm_desc.cell_inv_area, FVDBM_VARIANT_PAIR; temporal blocking removed */

/* status codes */
#define FVDBM_OK              0
#define FVDBM_ERR_ARG        -1   /* bad argument / inconsistent mesh description        */
#define FVDBM_ERR_CUDA       -2   /* CUDA runtime failure (no device, OOM, launch error) */
#define FVDBM_ERR_STATE      -3   /* call not valid in the handle's current state        */
#define FVDBM_ERR_UNSUPPORTED -4  /* (Q,K,dtype,scheme) combination not compiled in      */

/* flux schemes: Faces.flux_scheme, src/containers.py:197-203 */
#define FVDBM_SCHEME_UPWIND       0
#define FVDBM_SCHEME_LAX_WENDROFF 1

/* execution modes */
#define FVDBM_MODE_AUTO   0   /* fused when the mesh is cell/face consistent, else staged        */
#define FVDBM_MODE_STAGED 1   /* five reference-shaped kernels S1..S5 (general, materialises all) */
#define FVDBM_MODE_FUSED  2   /* node kernel + one cell-centric kernel per step                   */

/* fused-kernel variants (fvdbm_set_option(h, FVDBM_OPT_VARIANT, v)) */
#define FVDBM_VARIANT_AUTO   0   /* fp32: REC for D2Q9 (every size) and for D2Q13 with <= 64k owned cells; PAIR
                                    for D2Q13 >= 4M; otherwise (and fp64) DIRECT -- measured choices, api.cu:
                                    default_variant                                                           */
#define FVDBM_VARIANT_DIRECT 1   /* thread per cell, all operands through L1/L2               */
#define FVDBM_VARIANT_TMA    2   /* persistent CTAs, cp.async.bulk + mbarrier tile pipeline   */
#define FVDBM_VARIANT_PAIR   3   /* fp32 only: two cells per thread, packed FFMA2 math, 64-bit
                                    coalesced streaming accesses                              */
#define FVDBM_VARIANT_REC    4   /* thread per cell over the RECORD layout (a cell's Q-1 moving populations are
                                    contiguous -- one 32-byte sector for fp32 D2Q9 -- so a neighbour gather is
                                    one 256-bit load of one sector, a few 128-bit loads otherwise); fp32 D2Q9
                                    additionally packs the arithmetic over population pairs (FFMA2)                              */
/* All variants execute the same canonical operation sequence: results are bit-identical.
 * Debug / A-B environment overrides read once at fvdbm_create (each mirrors an fvdbm_option):
 *   FVDBM_VARIANT, FVDBM_TILE_CELLS, FVDBM_STAGES, FVDBM_GRAPH_STEPS, FVDBM_CTAS_PER_SM,
 *   FVDBM_REVERSE_SWEEP, FVDBM_PDL, FVDBM_PREFETCH_DIST, FVDBM_OVERLAP (0: no side stream for the interior cells),
 *   FVDBM_PLAN_THREADS (host planner threads), FVDBM_NCCL_LIB (path of libnccl to dlopen). */

/* fields for fvdbm_get / fvdbm_set (reference attribute in brackets) */
enum fvdbm_field {
    FVDBM_CELL_PDF    = 0,  /* [N*Q] cells.pdf      current populations                         */
    FVDBM_CELL_RHO    = 1,  /* [N]   cells.rho      moments of the PDFs *before* the last step   */
    FVDBM_CELL_VEL    = 2,  /* [N*2] cells.vel      (one-step lag, src/environment.py:60-64)     */
    FVDBM_CELL_PDF_EQ = 3,  /* [N*Q] cells.pdf_eq                                                */
    FVDBM_FACE_FLUX   = 4,  /* [F*Q] faces.pdf      flux of the last step                        */
    FVDBM_NODE_PDF    = 5,  /* [P*Q] nodes.pdf      only rows of tracked nodes are written       */
    FVDBM_NODE_RHO    = 6,  /* [P]   nodes.rho      (tracked = type!=0 or on a boundary face);   */
    FVDBM_NODE_VEL    = 7,  /* [P*2] nodes.vel      other rows of dst are left untouched         */
    FVDBM_CELL_PDF_PREV = 8 /* [N*Q] populations before the last step (diagnostic)               */
};

/* integer queries for fvdbm_info */
enum fvdbm_info_key {
    FVDBM_INFO_MODE = 0,          /* FVDBM_MODE_STAGED or FVDBM_MODE_FUSED actually in use */
    FVDBM_INFO_STEPS = 1,         /* steps taken since create                             */
    FVDBM_INFO_LAUNCHES = 2,      /* kernels launched by this handle since create          */
    FVDBM_INFO_TRACKED_NODES = 3,
    FVDBM_INFO_BOUNDARY_SIDES = 4,/* (cell,k) pairs whose face has a ghost cell            */
    FVDBM_INFO_DEVICE_BYTES = 5,
    FVDBM_INFO_VARIANT = 6,
    FVDBM_INFO_NPAD = 7,
    FVDBM_INFO_FUSED_OK = 8,      /* 1 if the mesh admits the fused path                   */
    FVDBM_INFO_HALO_CELLS = 9,
    FVDBM_INFO_OWNED_CELLS = 10,
    FVDBM_INFO_GRAPH_STEPS = 11   /* iterations per CUDA-graph launch (0 = graphs off): the natural batch size */
};

enum fvdbm_option {
    FVDBM_OPT_VARIANT = 0,        /* FVDBM_VARIANT_*                                         */
    FVDBM_OPT_TILE_CELLS = 1,     /* TMA variant: cells per CTA tile (128, 256, 512)         */
    FVDBM_OPT_STAGES = 2,         /* TMA variant: pipeline depth (2..4)                      */
    FVDBM_OPT_GRAPH_STEPS = 3,    /* steps captured per CUDA graph (0 = no graph)            */
    FVDBM_OPT_CTAS_PER_SM = 4,    /* persistent grid = 148 * this (0 = occupancy query)      */
    FVDBM_OPT_REVERSE_SWEEP = 5,  /* 1: odd steps sweep tiles backwards (L2 reuse of writes) */
    FVDBM_OPT_PREFETCH_DIST = 7,  /* >0: each CTA bulk-prefetches into L2 the streaming operands of the CTA this
                                     many blocks ahead (direct / pair kernels); 0 = off                      */
    FVDBM_OPT_PDL = 8,            /* 1: single-stream [nodes -> cells] chain with programmatic dependent launches
                                     (default below 1.5M cells, where the step is launch-latency bound); 0: two-stream
                                     overlap schedule (default above)                                          */
    FVDBM_OPT_TEMPORAL = 6        /* removed in ABI 2 (two-iterations-per-pass temporal blocking halved the DRAM
                                     traffic but measured slower); 0 accepted, 1 -> ERR_UNSUPPORTED */
};

typedef struct fvdbm_handle fvdbm_handle;

typedef struct fvdbm_desc {
    int32_t abi_version;     /* FVDBM_ABI_VERSION */
    int32_t device_id;       /* CUDA device ordinal */
    int32_t dtype;           /* 32 or 64 */
    int32_t scheme;          /* FVDBM_SCHEME_* */
    int32_t Q;               /* 9 (D2Q9) or 13 (D2Q13): src/dynamics.py:50-102 */
    int32_t K;               /* faces per cell: 3 (triangles) or 4 (quads) */
    int32_t M;               /* ring width of node_cell_idx / node_cell_dist */
    int32_t mode;            /* FVDBM_MODE_* */
    int64_t N, F, P;         /* cells, faces, nodes */
    int64_t N_owned;         /* cells [0,N_owned) are updated, [N_owned,N) are halo copies
                                refreshed by the caller (multi-GPU); 0 or N = all owned      */
    double tau, delta_t;     /* D2Q9(tau, delta_t): src/dynamics.py:65-68 */
    /* lattice constants evaluated by the host in the handle's precision, passed as doubles
       (exactly representable): W[q], C^2, 2C^4, 2C^2, 2C^6 (D2Q13 only) */
    double lat_w[16];
    double cs2, two_cs4, two_cs2, two_cs6;
    const int32_t* cell_face_idx;
    const int32_t* cell_face_sign;
    const int32_t* face_cell_idx;
    const void*    face_dists;
    const int32_t* face_node_idx;
    const void*    face_n;
    const void*    face_L;
    const int32_t* node_type;
    const int32_t* node_cell_idx;
    const void*    node_cell_dist;
    const void*    cell_pdf;
    const void*    node_pdf;
    const void*    node_rho;
    const void*    node_vel;
    const int32_t* cell_perm;  /* optional [N]: storage position of original cell i (a bijection
                                  onto [0,N), must map owned cells onto [0,N_owned)); NULL = identity */
    const void*    cell_inv_area; /* optional [N] real: 1/area of every cell, multiplying the flux divergence
                                  (physically consistent update on non-unit cells).  NULL = the reference's
                                  behaviour: no area anywhere (src/containers.py:115-121), i.e. parity mode */
} fvdbm_desc;

int  fvdbm_abi_version(void);
int  fvdbm_create(const fvdbm_desc* desc, fvdbm_handle** out);
void fvdbm_destroy(fvdbm_handle* h);
const char* fvdbm_last_error(const fvdbm_handle* h);

/* advance nsteps iterations of Environment.step() (src/environment.py:55-65); asynchronous */
int  fvdbm_step(fvdbm_handle* h, int nsteps);
/* same, bracketed by CUDA events on the handle's stream; blocks; *ms = device time */
int  fvdbm_step_timed(fvdbm_handle* h, int nsteps, float* ms);
int  fvdbm_sync(fvdbm_handle* h);

/* host <-> device transfer of one field in reference layout; bytes must match exactly */
int  fvdbm_get(fvdbm_handle* h, int field, void* dst, size_t bytes);
int  fvdbm_set(fvdbm_handle* h, int field, const void* src, size_t bytes);
/* Pipelined variants for callers that stream data every few iterations (e.g. coupling, in-situ output).
 * fvdbm_set_async: the H2D copy runs on its own stream and overlaps the iterations already enqueued; the state
 *   change is ordered behind them.  `src` must stay valid until fvdbm_sync (or any later fvdbm_wait / fvdbm_get).
 * fvdbm_get_async: an export pass ordered behind the enqueued iterations fills a device staging slot (two
 *   slots; cells.rho and cells.vel share one pass), the D2H copy then overlaps later iterations.  `dst` is
 *   complete once fvdbm_wait(h, *ticket) returns.  Use pinned host memory for real overlap. */
int  fvdbm_set_async(fvdbm_handle* h, int field, const void* src, size_t bytes);
int  fvdbm_get_async(fvdbm_handle* h, int field, void* dst, size_t bytes, int64_t* ticket);
int  fvdbm_wait(fvdbm_handle* h, int64_t ticket);
/* change tau / delta_t without rebuilding (they are Python floats baked into the jit in the
   reference, src/dynamics.py:65-68) */
int  fvdbm_set_params(fvdbm_handle* h, double tau, double delta_t);
int  fvdbm_set_option(fvdbm_handle* h, int option, int64_t value);
int  fvdbm_info(const fvdbm_handle* h, int key, int64_t* value);
/* failure detection (the reference only shows blow-ups in plots): number of owned cells whose current
   populations contain a NaN / Inf; one small reduction kernel, blocks until the count is known */
int  fvdbm_check_finite(fvdbm_handle* h, int64_t* nonfinite_cells);

/* ---- multi-GPU halo plumbing (one handle per rank; cells [N_owned,N) are halo copies) -------
 * Send lists are given in the caller's ORIGINAL local numbering; the library translates.
 * fvdbm_halo_pack gathers the current PDFs of `count` cells into a device buffer laid out
 * [count][Q] (real) ; fvdbm_halo_unpack scatters such a buffer into local cells.  Both run on
 * the handle's stream.  Device pointers may come from any allocator (torch, cudaMalloc). */
int  fvdbm_halo_set_lists(fvdbm_handle* h, const int32_t* send_cells, int64_t n_send,
                          const int32_t* recv_cells, int64_t n_recv);
int  fvdbm_halo_pack(fvdbm_handle* h, void* dev_send_buf);
int  fvdbm_halo_unpack(fvdbm_handle* h, const void* dev_recv_buf);
/* split stepping for overlap: phase 0 = interior cells (no halo / boundary dependence),
   phase 1 = node kernel + remaining cells + buffer swap.  fvdbm_step == phase 0 then 1. */
int  fvdbm_step_phase(fvdbm_handle* h, int phase);
/* raw stream handle (cudaStream_t) so the host can order its collectives against the engine */
void* fvdbm_stream(fvdbm_handle* h);

/* ---- native exchange: the engine owns an NCCL communicator and runs the whole distributed
 * iteration inside fvdbm_step (pack -> ncclSend/ncclRecv with the neighbour ranks, overlapped with
 * the interior update -> unpack -> node kernel -> border cells), no host code per iteration.
 * The 128-byte id comes from fvdbm_comm_unique_id on one rank and is distributed by the host
 * (torch.distributed / MPI / a file).  libnccl.so.2 is loaded with dlopen on first use. */
#define FVDBM_COMM_ID_BYTES 128
int  fvdbm_comm_unique_id(void* id_out);
int  fvdbm_comm_init(fvdbm_handle* h, int nranks, int rank, const void* id);
/* peers and per-peer cell counts of the lists given to fvdbm_halo_set_lists (same order) */
int  fvdbm_halo_set_peers(fvdbm_handle* h, const int32_t* send_peers, const int64_t* send_counts, int n_send_peers,
                          const int32_t* recv_peers, const int64_t* recv_counts, int n_recv_peers);

/* ---- host-only decomposition helper (no GPU needed): Hilbert-curve key of every cell centroid of a raw mesh
 * (points [P*2] f64, elements [ncells*K] i32), centroid -> integer grid by (c - lo) * scale, `bits` per axis;
 * OpenMP over cells (FVDBM_PLAN_THREADS).  elements == NULL: `points` already holds the ncells centroids.
 * Used by partition.sfc_owner_from_raw (10^8-cell meshes) and by the Hilbert renumbering. */
int  fvdbm_sfc_keys(const double* points, const int32_t* elements, int64_t ncells, int K, int bits,
                    double lo_x, double lo_y, double scale, int64_t* keys_out);

/* ---- host-only Mesher-equivalent (no GPU needed; OpenMP, FVDBM_PLAN_THREADS): everything the reference's
 * Mesher.calc_mesh_properties (/root/reference/src/mesher.py:319-383, built from :63-316 and :506-558) derives from a raw
 * triangle mesh -- the caller-side step right before the path.  Inputs: points [P*2] f64, cells [N*3]
 * i32 counter-clockwise (mesher.py:63-78), faces [F*2] i32, optional point_alias [P] i32 (periodic identification;
 * NULL = none).  Every output is caller-allocated and named after the Mesher attribute it replaces; integer results
 * are bit-identical to the reference Mesher's, floats to a few ulp of it (and bit-identical to fvdbm_jax_b200.mesher's
 * NumPy path).  M = fvdbm_mesh_ring_width(...) sizes the two ring arrays.  Returns FVDBM_OK, FVDBM_ERR_ARG (id out of
 * range, null pointer) or FVDBM_ERR_STATE (a cell edge is missing from `faces`: the reference raises KeyError). */
typedef struct fvdbm_mesh_desc {
    int64_t N, F, P;
    int32_t M, reserved;
    const double*  points;
    const int32_t* cells;
    const int32_t* faces;
    const int32_t* point_alias;
    double*  cell_centers;                 /* [N*2]   mesher.py:113-120 */
    int64_t* cell_face_indices;            /* [N*3]   mesher.py:122-138 */
    double*  cell_face_normals;            /* [N*3*2] mesher.py:140-158 (outward) */
    int32_t* cell_face_normal_signs;       /* [N*3]   mesher.py:160-169 */
    int32_t* faces_out;                    /* [F*2]   mesher.py:80-110 (boundary faces flipped to an outward normal) */
    double*  face_centers;                 /* [F*2]   mesher.py:172-178 */
    double*  face_normals;                 /* [F*2]   mesher.py:180-187 */
    double*  face_lengths;                 /* [F]     mesher.py:189-195 */
    int64_t* face_cell_indices;            /* [F*2]   mesher.py:197-266 (slot 0 -> slot 1 along the normal, -1 = ghost) */
    double*  face_cell_center_distances;   /* [F*2]   mesher.py:222-283 */
    double*  stencil_norms;                /* [F*2]   mesher.py:506-544 */
    double*  cc_stencil_dist;              /* [F*2]   mesher.py:231-256 on the stencil direction */
    double*  face_stencil_angles;          /* [F]     mesher.py:546-558 */
    int64_t* point_cell_indices;           /* [P*M]   mesher.py:286-300, -1 padded */
    double*  point_cell_center_distances;  /* [P*M]   mesher.py:302-316, -1 padded */
} fvdbm_mesh_desc;
int64_t fvdbm_mesh_ring_width(const int32_t* cells, const int32_t* point_alias, int64_t N, int64_t P);   /* <0: bad id */
int  fvdbm_mesh_properties(const fvdbm_mesh_desc* mesh);
/* `faces` of a raw mesh that only has points + elements (K = 3 or 4 vertices per cell): the unique undirected edges, one
 * row per face ordered by (min, max) canonical point id, each row carrying the point ids of the first cell edge that
 * produced it (what the reference gets from meshpy's `mesh.faces`, /root/reference/src/mesher.py:57).  faces_out holds up
 * to N*K rows; returns the number of faces, or FVDBM_ERR_ARG. */
int64_t fvdbm_mesh_unique_edges(const int32_t* cells, int64_t N, int K, int64_t P, const int32_t* point_alias,
                                int32_t* faces_out);

/* ---- host-only planning (no GPU needed): builds the device layout from a desc so that the
 * layout logic is unit-testable on CPU.  key = name of a plan array, see csrc/plan.hpp. */
typedef struct fvdbm_plan fvdbm_plan;
int  fvdbm_plan_create(const fvdbm_desc* desc, fvdbm_plan** out);
void fvdbm_plan_destroy(fvdbm_plan* p);
/* returns element count (and fills *ptr / *elem_bytes) or <0 if the key is unknown */
int64_t fvdbm_plan_array(const fvdbm_plan* p, const char* key, const void** ptr, int32_t* elem_bytes);
int64_t fvdbm_plan_scalar(const fvdbm_plan* p, const char* key);

#ifdef __cplusplus
}
#endif
#endif /* FVDBM_B200_H */
