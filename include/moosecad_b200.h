/*
 * moosecad_b200 — C ABI of the B200-native meshing hot path.
 *
 * Drop-in boundary for ONE path of dvc94ch/moosecad: evaluating the luascad CSG
 * signed-distance tree over tessellate3d's 5-tet cubic lattice and running isosurface
 * stuffing on it.  The reference has no FFI of its own; its seams are Rust
 * traits.  Each entry point below names the reference interface it replaces (paths are
 * relative to the reference repository).  INTEGRATION.md shows the Rust `extern "C"`
 * block and the patched call sites.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in any signature.  A CUDA stream is
 *     passed as `void*` (cudaStream_t), device memory as `void*`.
 *   - every function returning `int` returns MC_OK (0) or a negative mc_status; the
 *     message is available from mc_last_error() (thread-local).  Nothing throws or
 *     aborts across the ABI (the reference panics on e.g. a non-finite bbox,
 *     tessellate3d/src/lib.rs:44-46; here that is MC_ERR_BBOX).
 *   - the caller owns tapes and contexts; the library owns every buffer reachable from a
 *     mc_mesh until mc_mesh_free().
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device and fails
 *     with MC_ERR_CUDA when none is usable.
 *   - a context is bound to one device and one stream; calls on one context are
 *     synchronous with respect to the host unless stated otherwise and must not be
 *     issued concurrently from several threads (the reference is single-threaded).
 */
#ifndef MOOSECAD_B200_H
#define MOOSECAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_ABI_VERSION 1

typedef enum mc_status {
    MC_OK = 0,
    MC_ERR_ARG = -1,         /* bad handle / node id / argument */
    MC_ERR_BBOX = -2,        /* non-finite root bounding box (reference: unwrap() panic) */
    MC_ERR_SINGULAR = -3,    /* affine matrix not invertible (reference: panic) */
    MC_ERR_UNSUPPORTED = -4, /* node kind outside this path (bend / twist) */
    MC_ERR_CUDA = -5,        /* CUDA runtime error or no device */
    MC_ERR_LIMIT = -6,       /* tape deeper / larger than the device evaluator supports */
    MC_ERR_DIVERGED = -7,    /* edge bisection did not reach the error threshold */
    MC_ERR_STATE = -8        /* call made in the wrong phase */
} mc_status;

typedef struct mc_tape mc_tape; /* a recorded scene: the flat node list luascad lowers to */
typedef struct mc_ctx mc_ctx;   /* device + stream + reusable workspace */
typedef struct mc_mesh mc_mesh; /* result of one tessellation (device resident, host views on demand) */
typedef struct mc_surface mc_surface; /* Mesh::to_boundary(): the surface as a triangle mesh */

const char* mc_last_error(void);
int mc_abi_version(void);

/* ------------------------------------------------------------------------------------
 * Scene lowering.  One call per node created in luascad::geom (luascad/src/lib.rs:89-281).
 * Nodes are immutable once pushed; functions return the new node id (>= 0) or a negative
 * mc_status.  Children are referenced by id and may be shared (the reference clones them).
 * Bounding boxes are propagated with the rules of implicit3d 0.14.2 / bbox 0.11.2.
 * ------------------------------------------------------------------------------------ */
mc_tape* mc_tape_new(void);
void mc_tape_free(mc_tape*);
int32_t mc_tape_num_nodes(const mc_tape*);

/* geom.plane_x/y/z (negative=0) and geom.plane_neg_x/y/z (negative=1); luascad/src/lib.rs:89-123 */
int32_t mc_tape_plane(mc_tape*, int axis, int negative, double d);
/* geom.plane_hesian; luascad/src/lib.rs:125-132 */
int32_t mc_tape_plane_hesian(mc_tape*, double nx, double ny, double nz, double p);
/* geom.plane_3_points; luascad/src/lib.rs:134-152 */
int32_t mc_tape_plane_3_points(mc_tape*, const double a[3], const double b[3], const double c[3]);
/* geom.sphere; luascad/src/lib.rs:154-158 */
int32_t mc_tape_sphere(mc_tape*, double radius);
/* geom.i_cylinder; luascad/src/lib.rs:160-164 */
int32_t mc_tape_i_cylinder(mc_tape*, double radius);
/* geom.i_cone (offset 0) and the cone inside geom.cylinder; luascad/src/lib.rs:166-170,200 */
int32_t mc_tape_i_cone(mc_tape*, double slope, double offset);
/* Cone::set_bbox; luascad/src/lib.rs:202-205.  bbox = {minx,miny,minz,maxx,maxy,maxz} */
int mc_tape_set_bbox(mc_tape*, int32_t node, const double bbox[6]);
/* test fixture of tessellate3d's own tests (`UnitSphere`, tessellate3d/src/lib.rs:460-484):
 * value = |p|^2 - 1, bbox [-1,1]^3 */
int32_t mc_tape_unit_sphere_sq(mc_tape*);
/* Union::from_vec / Intersection::from_vec / Intersection::difference_from_vec
 * (luascad/src/lib.rs:173-187,209-217,252-273).  n == 1 returns a copy of the child. */
int32_t mc_tape_union(mc_tape*, const int32_t* children, int n, double r);
int32_t mc_tape_intersection(mc_tape*, const int32_t* children, int n, double r);
int32_t mc_tape_difference(mc_tape*, const int32_t* children, int n, double r);
int32_t mc_tape_negation(mc_tape*, int32_t child);
/* Object::translate / rotate / scale (luascad/src/lib.rs:221-237): a fresh affine node, or
 * the composed matrix when `node` already is an affine node */
int32_t mc_tape_translate(mc_tape*, int32_t node, double x, double y, double z);
int32_t mc_tape_rotate(mc_tape*, int32_t node, double x, double y, double z);
int32_t mc_tape_scale(mc_tape*, int32_t node, double x, double y, double z);
/* geom.bend / geom.twist (luascad/src/lib.rs:239-249): not on this path -> MC_ERR_UNSUPPORTED */
int32_t mc_tape_bend(mc_tape*, int32_t node, double width);
int32_t mc_tape_twist(mc_tape*, int32_t node, double height);
/* Object::set_parameters(&PrimitiveParameters{fade_range, r_multiplier}); src/main.rs:53-58.
 * Applies to every node of the tape (main applies it to every output). */
int mc_tape_set_parameters(mc_tape*, double fade_range, double r_multiplier);
/* Object::bbox(); src/main.rs:29-31 */
int mc_tape_bbox(const mc_tape*, int32_t node, double out_bbox[6]);
/* size of the device program for (root, slack): 8-byte words, value-stack depth, affine depth */
int mc_tape_program_info(const mc_tape*, int32_t root, double slack, uint64_t* words,
                         int32_t* value_depth, int32_t* point_depth, uint64_t* flops_per_point);

/* content hash of that program and the root bounding box (everything the device sees of the
 * scene): equal scenes lowered twice hash alike.  Key of the slab-partition cache (slabs.py). */
int mc_tape_program_hash(const mc_tape*, int32_t root, double slack, uint64_t* hash);
/* the flat device program itself (csrc/mc_program.h), host only: returns its length in 8-byte words and, when
 * `out` is given (capacity words), copies it.  tests/program_interp.py evaluates it on the CPU against the oracle. */
int64_t mc_tape_program_words(const mc_tape*, int32_t root, double slack, uint64_t* out, uint64_t capacity);

/* ------------------------------------------------------------------------------------
 * Device context.
 * ------------------------------------------------------------------------------------ */
/* device: CUDA ordinal; stream: cudaStream_t to launch on (NULL = the legacy default stream) */
mc_ctx* mc_ctx_new(int device, void* stream);
void mc_ctx_free(mc_ctx*);
int mc_ctx_device(const mc_ctx*);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t mc_ctx_launch_count(const mc_ctx*);
/* bytes currently held by the reusable workspace */
uint64_t mc_ctx_workspace_bytes(const mc_ctx*);
/* release the cached workspace */
int mc_ctx_trim(mc_ctx*);

/* ------------------------------------------------------------------------------------
 * ObjectAdaptor::value — Object::approx_value(p, slack) for a batch of points
 * (src/main.rs:33-35,95).  points: n * 3 doubles (xyz AoS), out: n doubles; HOST buffers.
 * ------------------------------------------------------------------------------------ */
int mc_eval_points(mc_ctx*, const mc_tape*, int32_t root, double slack, const double* points,
                   uint64_t n, double* out);
/* same, DEVICE buffers, asynchronous on the context's stream */
int mc_eval_points_device(mc_ctx*, const mc_tape*, int32_t root, double slack,
                          const void* d_points, uint64_t n, void* d_out);

/* ------------------------------------------------------------------------------------
 * The mesher.  Replaces
 *   IsosurfaceStuffing::new(&ObjectAdaptor{implicit, res}, res, err).tessellate()
 * (src/main.rs:74-82 -> tessellate3d/src/lib.rs:36-52,420-431).
 * ------------------------------------------------------------------------------------ */
typedef struct mc_slab {
    /* z-slab sharding: this call meshes cell layers [z_begin, z_end) of the
     * global lattice.  A NULL mc_slab pointer means the whole lattice (single GPU); an empty
     * range (z_begin == z_end) is a legitimate slab that owns nothing. */
    uint64_t z_begin, z_end;
} mc_slab;

typedef struct mc_options {
    uint32_t struct_size;       /* sizeof(mc_options) */
    uint32_t flags;             /* MC_OPT_* */
    uint32_t max_bisect_iters;  /* 0 = default (200); reference loops until convergence */
    uint32_t reserved;
} mc_options;
#define MC_OPT_KEEP_PHI 1u     /* keep the lattice SDF field readable after the call */
#define MC_OPT_SKIP_QUALITY 2u /* skip check_for_inverted_tets (stdout-only in the reference) */
#define MC_OPT_NO_TILE_PRUNING 4u /* evaluate the full SDF program at every lattice point */
#define MC_OPT_FORCE_TILE_PRUNING 8u /* use the tile-specialised evaluator even for tiny programs */
#define MC_OPT_COUNT_FLOPS 16u /* instrumented SDF kernel: count the FP64 operations it executes (slower) */
#define MC_OPT_PER_TILE_QUALITY 32u /* plain lattice tets: evaluate (step, parity) classes per tile, not globally */
#define MC_OPT_RECLASSIFY_EMIT 64u /* A/B check: the tet emission reads the field and sorts every trimmed tet again instead of using the shape bytes the classification recorded */

/* Load estimate for z-slab sharding (no reference counterpart).  An interval evaluation
 * of the SDF program over the 16^3-point tiles sorts every tile into a cost class:
 *   0 sign not proven (full SDF evaluation, possibly surface work)
 *   1 proven inside, neighbourhood too (no SDF evaluation; arithmetic output)
 *   2 proven inside next to other tiles (shell evaluation + output)
 *   3 proven outside, neighbourhood too (nothing to do)
 *   4 proven outside next to other tiles (shell evaluation)
 * mc_tile_layer_classes fills counts[layer * MC_TILE_CLASS_STRIDE + q] for the tile layers
 * [layer_begin, layer_end) of the n = ceil((dim[2] + 1) / 16) tile layers (the other rows are zeroed, so
 * ranks can each take a range and sum the tables): q = 0..4 the number of tiles per class, q = 5 the summed
 * per-point work (approximate FP64 operations of the tile-specialised program) of the class-0 tiles,
 * q = 6 the same for the shell-evaluated tiles (classes 2 and 4).  mc_estimate_layer_costs turns them
 * into a relative cost per CELL layer z (n = dim[2] doubles) with the weights below (least-squares
 * fit to per-slab device times, scripts/slab_balance.py).  Deterministic, so every rank derives the
 * same partition. */
#define MC_TILE_CLASS_STRIDE 7
#define MC_SLAB_COST_UNPROVEN 0.05
#define MC_SLAB_COST_NEG_SKIPPED 0.07
#define MC_SLAB_COST_NEG_SHELL 0.07
#define MC_SLAB_COST_POS_SKIPPED 0.002
#define MC_SLAB_COST_POS_SHELL 0.002
#define MC_SLAB_COST_UNPROVEN_WORK 1.9e-4
#define MC_SLAB_COST_SHELL_WORK 0.0
int mc_tile_layer_classes(mc_ctx*, const mc_tape* volume, int32_t root, double resolution, uint64_t layer_begin,
                          uint64_t layer_end, uint64_t* counts, uint64_t n);
/* the MC_TILE_CLASS_STRIDE weights mc_estimate_layer_costs applies to those counts */
void mc_slab_class_weights(double* out);
int mc_estimate_layer_costs(mc_ctx*, const mc_tape* volume, int32_t root, double resolution, double* cost, uint64_t n);

/* Phase 1: SDF field, warp, classification, counting.  After it returns, the per-slab counts
 * (mc_mesh_counts) are known; nothing has been emitted yet. */
int mc_tessellate_begin(mc_ctx*, const mc_tape* volume, int32_t root, double resolution,
                        double relative_error, const mc_slab* slab, const mc_options* opts,
                        mc_mesh** out);
/* Phase 2: number and emit this slab's vertices (lattice vertices, warped where applicable,
 * and the bisected cut vertices).  vertex_base: global index of this slab's first vertex
 * (exclusive scan of mc_mesh_counts()[0] over the ranks; 0 on one GPU).  Afterwards
 * mc_mesh_upper_plane_table() is valid. */
int mc_tessellate_emit_vertices(mc_mesh*, uint64_t vertex_base);
/* (slab plumbing: no reference counterpart)
 * Optional, between phases 1 and 2: launch the part of phase 2 that does not need vertex_base
 * (coordinates + bisection) and return without waiting.  A multi-rank caller issues it before
 * the blocking exchange of the counts, so the device works while the ranks wait for each other. */
int mc_tessellate_emit_vertices_local(mc_mesh*);
/* Phase 3: emit this slab's tets with global vertex ids.  tet_base: global index of this
 * slab's first tet.  d_lower_plane: device pointer to the id table of the lattice plane
 * z = z_begin produced by the rank below (its mc_mesh_upper_plane_table), or NULL for the
 * first slab. */
int mc_tessellate_emit_tets(mc_mesh*, uint64_t tet_base, const void* d_lower_plane);
/* (slab plumbing: no reference counterpart; the statistics themselves are
 * check_for_inverted_tets, tessellate3d/src/lib.rs:400-418)
 * Optional, before phase 3 on a slab with z_begin > 0: the coordinates of the rank below's top
 * tile layer (its mc_mesh_upper_layer_coords; `count` vertices with global ids starting at
 * first_global_id = its vertex_base + its first_local), so that check_for_inverted_tets can also
 * evaluate the trimmed tets that touch the shared plane.  Without it those tets are left out of
 * the aspect-ratio / inverted statistics of mc_mesh_stats (the mesh itself is unaffected). */
int mc_tessellate_set_lower_coords(mc_mesh*, const void* d_coords, uint64_t first_global_id, uint64_t count);
/* what the rank ABOVE needs for that: the vertices of the tile layer holding this slab's top plane
 * are the tail of the slab's vertex range: local index of the first one + how many (valid after
 * phase 1; count = 0 on the top slab), a device pointer to their xyz (valid after
 * mc_tessellate_emit_vertices(_local)), and an asynchronous copy into a caller's DEVICE buffer
 * (count * 24 bytes) on the context's stream: the send buffer of the NVLink exchange. */
int mc_mesh_upper_layer_vertices(const mc_mesh*, uint64_t* first_local, uint64_t* count);
int mc_mesh_upper_layer_coords(mc_mesh*, void** d_coords);
int mc_mesh_copy_upper_layer_coords(mc_mesh*, void* d_dst);
/* all phases for the whole lattice on one device */
int mc_tessellate(mc_ctx*, const mc_tape* volume, int32_t root, double resolution,
                  double relative_error, const mc_options* opts, mc_mesh** out);
/* The same job sharded into z-slabs over n devices driven by THIS process (the
 * `devices, ndev` form of the mesher call; what a single-process host like src/main.rs:74-82 uses).
 * ctxs[i]: the context (device + stream + workspace) that meshes slab i, bottom to top; the contexts
 * must be distinct (they may share a device).  One host thread per context runs inside the call;
 * slab boundaries come from the load estimate; the per-slab counts are exchanged in host memory, the
 * id table of every shared plane and the coordinates next to it travel device to device
 * (cudaMemcpyPeerAsync).  out[i] receives slab i's mesh: GLOBAL vertex ids, tets / boundary sides
 * numbered from the slab's global tet offset; the caller frees each with mc_mesh_free. */
int mc_tessellate_multi(mc_ctx* const* ctxs, int n, const mc_tape* volume, int32_t root, double resolution,
                        double relative_error, const mc_options* opts, mc_mesh** out);
void mc_mesh_free(mc_mesh*);

/* counts: [0] vertices owned by this slab, [1] tets, [2] lattice points evaluated (incl.
 * halo), [3] lattice points owned, [4] cut edges, [5] warped vertices */
int mc_mesh_counts(const mc_mesh*, uint64_t counts[6]);
/* tile specialisation of the SDF program (0 tiles when the plain evaluator was used): number of
 * tiles, sum over tiles of the compacted program size in 8-byte words, size of the full program */
int mc_mesh_prune_stats(const mc_mesh*, uint64_t* n_tiles, uint64_t* pruned_words_total, uint64_t* full_words);
/* tile bookkeeping of the last run: [0] tiles (16^3 lattice points each), [1] tiles whose SDF
 * evaluation was skipped (proven interior / exterior), [2] point tiles and [3] cell tiles that
 * needed the general stuffing path (the rest took the uniform-sign fast paths), [4] FP64
 * operations executed by the SDF kernel (MC_OPT_COUNT_FLOPS runs only, else 0), [5] sub-boxes
 * (8 x 4 x 4 points, 32 per tile) of fully evaluated tiles whose sign was proven, i.e. skipped */
int mc_mesh_tile_stats(const mc_mesh*, uint64_t out[6]);
/* id table of the plane z = z_end for the rank above: one uint64 per lattice point of the
 * plane, (global base id << 16) | mask.  Device pointer + length in entries. */
int mc_mesh_upper_plane_table(mc_mesh*, void** d_table, uint64_t* n_entries);
/* copy that table into a caller-provided DEVICE buffer (n_entries * 8 bytes), asynchronously on
 * the context's stream: the send buffer of the NVLink exchange */
int mc_mesh_copy_upper_plane_table(mc_mesh*, void* d_dst);
/* mirrors the reference's stdout (tessellate3d/src/lib.rs:247,355,366,408-417,421,424):
 * dim, tet count after cut_lattice / before trim / after trim, the 7 case counters,
 * aspect-ratio min/max and the number of inverted tets */
int mc_mesh_stats(const mc_mesh*, uint64_t dim[3], uint64_t ntets_stage[3], uint64_t stats[7],
                  double ar_minmax[2], uint64_t* inverted);
/* CUDA-event stage times of the last run in milliseconds:
 * [0] sdf field [1] warp [2] classify+count [3] vertex scan+emit [4] bisection
 * [5] tet emit [6] quality [7] boundary [8] side-sets */
int mc_mesh_stage_ms(const mc_mesh*, double ms[9]);

/* Result views.  Layout = the mesh crate's (mesh/src/lib.rs:58-65): coords xyz AoS f64,
 * tets 4 x node id, 0-based.  Host copies are made on first request and stay valid until
 * mc_mesh_free. */
uint64_t mc_mesh_num_vertices(const mc_mesh*);
uint64_t mc_mesh_num_tets(const mc_mesh*);
const double* mc_mesh_coords(mc_mesh*);
const uint64_t* mc_mesh_tets(mc_mesh*);
/* device-resident views (global ids as uint32 when index_bytes == 4, else uint64) */
int mc_mesh_device_views(mc_mesh*, void** d_coords, void** d_tets, int* index_bytes);
/* copy into caller-provided HOST buffers (pinned or pageable): coords = 3*nv doubles,
 * tets = 4*nt entries of index_bytes (4 or 8) each */
int mc_mesh_copy_to_host(mc_mesh*, double* coords, void* tets, int index_bytes);
/* the lattice SDF field: (dim+1)^3 doubles, x fastest, z slowest.  Needs MC_OPT_KEEP_PHI (or a run
 * that evaluated every point: MC_OPT_NO_TILE_PRUNING / a tiny program); otherwise most entries
 * are +-1 placeholders of the proven sign and the call fails with MC_ERR_STATE. */
int mc_mesh_copy_phi(mc_mesh*, double* out, uint64_t n);

/* (test / measurement plumbing: no reference counterpart; SURVEY.md 8c "cheap order-independent
 * parity probes")  Order- and numbering-independent 64-bit checksums of the emitted mesh, computed on
 * the device: out[0] = sum over the vertices of H(x, y, z bit patterns), out[1] = sum over the tets of
 * H(the four vertex hashes, rotated by the even permutation that sorts them first), both mod 2^64 and
 * therefore additive over the slabs of one lattice; out[2] = tets left out because a node's
 * coordinates are not on this device (0 unless a slab was emitted without mc_tessellate_set_lower_coords);
 * out[3] = 0.  tests/canon.py holds the numpy mirror that is applied to the oracle's mesh. */
int mc_mesh_checksum(mc_mesh*, uint64_t out[4]);

/* ------------------------------------------------------------------------------------
 * Mesh::boundary() (mesh/src/lib.rs:341-357; src/main.rs:86) and the side-set loop
 * (src/main.rs:87-106).
 * ------------------------------------------------------------------------------------ */
/* number of surface sides; computes them on first call */
int mc_mesh_boundary(mc_mesh*, uint64_t* n_sides);
/* (elem, side) pairs, side in 0..3 per Tet4::face (mesh/src/lib.rs:661-668), sorted by (elem, side)
 * (the reference's HashSet order is arbitrary); host view, copied on first request */
const uint64_t* mc_mesh_boundary_sides(mc_mesh*);
/* the same pairs on the device (uint64 x 2 per side) */
int mc_mesh_boundary_device(mc_mesh*, void** d_sides, uint64_t* n_sides);
/* adds one side-set: surface sides whose three nodes all satisfy
 * |approx_value(p, EPSILON)| <= EPSILON for the given boundary object.  Returns its index. */
int mc_mesh_add_sideset(mc_mesh*, const mc_tape* boundary, int32_t root);
uint64_t mc_mesh_sideset_len(const mc_mesh*, int index);
const uint64_t* mc_mesh_sideset(mc_mesh*, int index);
int mc_mesh_sideset_device(mc_mesh*, int index, void** d_pairs, uint64_t* n);

/* Mesh::to_boundary() (mesh/src/lib.rs:359-386; what meshview renders): one Tri3 per surface side (the side's
 * nodes in Tet4::face order), vertices compacted to the referenced ones (Mesh::compact).  Host views: coords
 * xyz AoS f64, tris 3 x node id, valid until mc_surface_free.  Single device (not on a z-slab). */
int mc_mesh_to_boundary(mc_mesh*, mc_surface** out);
void mc_surface_free(mc_surface*);
uint64_t mc_surface_num_vertices(const mc_surface*);
uint64_t mc_surface_num_tris(const mc_surface*);
const double* mc_surface_coords(const mc_surface*);
const uint64_t* mc_surface_tris(const mc_surface*);

/* ------------------------------------------------------------------------------------
 * The consumer right after the path: exodus::write(&mesh, path) (exodus/src/lib.rs:14-155) builds
 * three big arrays from the mesh on one thread.  These calls produce them on the device, in the
 * writer's layout, and stream them into caller HOST buffers (pinned or pageable; chunked, the
 * layout kernel of one chunk overlaps the copy of the previous one), so that the Rust side goes
 * straight to `put_values`:
 *   coord        [num_dim = 3][num_nodes] f64 (exodus/src/lib.rs:55-63): this mesh / slab fills columns
 *                [vertex_base, vertex_base + n_vertices) of the three rows of a buffer with
 *                num_nodes_total columns (one GPU: vertex_base = 0, num_nodes_total = n_vertices)
 *   connect1     [num_el_in_blk1][4] i64, 1-based (exodus/src/lib.rs:84-96): this slab's n_tets rows
 *   elem_ss<i> / side_ss<i>  i32, 1-based (exodus/src/lib.rs:118-132) of side-set `index`
 *                (mc_mesh_sideset_len entries each); MC_ERR_LIMIT when an element number does not
 *                fit i32 (the reference's `*elem as i32 + 1` would wrap silently)
 * ------------------------------------------------------------------------------------ */
int mc_mesh_exodus_coord(mc_mesh*, double* coord, uint64_t num_nodes_total);
int mc_mesh_exodus_connect(mc_mesh*, int64_t* connect1);
int mc_mesh_exodus_sideset(mc_mesh*, int index, int32_t* elem_ss, int32_t* side_ss);

/* ------------------------------------------------------------------------------------
 * mesh crate predicates on arbitrary tet meshes (mesh/src/lib.rs:292-316), evaluated on the
 * device; used by the parity tests of check_for_inverted_tets.  HOST buffers.
 * ------------------------------------------------------------------------------------ */
int mc_tet_quality(mc_ctx*, const double* coords, uint64_t nv, const uint64_t* tets, uint64_t nt,
                   double* volume, double* aspect_ratio);

/* ------------------------------------------------------------------------------------
 * Device capability probes used by bench.py's roofline (measured, not assumed).
 * ------------------------------------------------------------------------------------ */
/* sustained DADD+DMUL (no FMA) rate in flop/s over `ms` milliseconds of work */
int mc_probe_fp64_nofma(mc_ctx*, double* flops_per_s);

/* ------------------------------------------------------------------------------------
 * Host-only self description (no device needed): the shape tables behind the tet emission and
 * Mesh::boundary (csrc/mc_shapes.hpp), copied into `out` (`bytes` must be sizeof(mc::ShapeTables),
 * returned by a call with out == NULL as a positive number).  Used by the CPU tests.
 * ------------------------------------------------------------------------------------ */
int64_t mc_shape_tables(void* out, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* MOOSECAD_B200_H */
