// Flat SDF program ("op-tape") format shared by the host lowering (mc_tape.cpp) and the
// device evaluators (mc_vm.cuh point evaluator, mc_prune.cuh interval evaluator).
//
// The program is a pre-order walk of the CSG tree (luascad/src/lib.rs:89-281 builds the tree;
// implicit3d 0.14.2 `Object::approx_value` defines the arithmetic).  It is
// an array of 8-byte words.  Every instruction starts with a header word
//     bits  0..7   opcode
//     bits  8..15  small argument (axis | inverted<<2, or child count k | union<<7)
//     bits 16..23  nesting level of the guard (warp-uniform all-pass bit; 255 = none)
//     bits 24..39  decision id (tile pruning, mc_prune.cuh): guards and boolean children
//     bits 40..63  a word index (jump target / back pointer), opcode dependent
// followed by its f64 parameters.  The `slack` argument of approx_value is a compile-time
// constant of each node (root slack, +r per boolean, /scale_min per affine), computed on the
// host with the same IEEE operations the reference performs at run time.
#pragma once
#include <stdint.h>

enum mc_opcode : uint32_t {
    MC_OP_END = 0,      // [hdr]                               return top of stack
    MC_OP_PLANE = 1,    // [hdr(axis,inv) d]                   push plane value
    MC_OP_NPLANE = 2,   // [hdr nx ny nz p0]                   push dot(p,n) - p0
    MC_OP_USQ = 3,      // [hdr]                               push |p|^2 - 1 (test fixture)
    MC_OP_CONE = 4,     // [hdr slope offset mult]             push cone value (no guard)
    MC_OP_GSPHERE = 5,  // [hdr(dec) bbox[6] slack r]          push guarded sphere
    MC_OP_GCYL = 6,     // [hdr(dec) bbox[6] slack r]          push guarded cylinder
    MC_OP_GUARD = 7,    // [hdr(level, dec, jump) bbox[6] slack]  a = bbox.distance(p); if no lane
                        //                                     has a <= slack: push a, jump past ENDG
    MC_OP_ENDG = 8,     // [hdr(level, dec, guard index)]      top = (a <= slack) ? top : a
    MC_OP_NEG = 9,      // [hdr]                               top = -top
    MC_OP_BOOL = 10,    // [hdr(k | union<<7, dec base) r exact_range]   pop k, push rvmin / rvmax
    MC_OP_APUSH = 11,   // [hdr m[12] (3x4 row major)]         save p; p = M*p
    MC_OP_APOP = 12,    // [hdr scale_min]                     restore p; top *= scale_min
    MC_OP_CHILD = 13,   // [hdr(union, dec, end of child)]     marks a boolean child (no-op)
    // only produced by the per-tile program compaction:
    MC_OP_BOXDIST = 14, // [hdr bbox[6]]                       push bbox.distance(p)  (guard fails on the whole tile)
    MC_OP_SPHERE = 15,  // [hdr r]                             push |p| - r            (guard passes on the whole tile)
    MC_OP_CYL = 16,     // [hdr r]                             push |p.xy| - r
    // fused forms emitted by the lowering for the shapes geom.* builds all the time:
    MC_OP_GBOX = 17,    // [hdr(dec) bbox[6] slack d[6] r exact_range]  guarded geom.cube: rvmax of the planes
                        //                                     X, Y, Z, NegX, NegY, NegZ (luascad/src/lib.rs:173-187)
    MC_OP_BOX = 18,     // [hdr d[6] r exact_range]            the same without the guard (compaction only)
    MC_OP_GAPUSH = 19,  // [hdr(level, dec, jump) bbox[6] slack m[12]]  GUARD immediately followed by APUSH
};

#define MC_NO_LEVEL 255u
#define MC_HDR(op, small, level, dec, idx)                                                        \
    ((uint64_t)(op) | ((uint64_t)((small) & 0xff) << 8) | ((uint64_t)((level) & 0xff) << 16) |    \
     ((uint64_t)((dec) & 0xffff) << 24) | ((uint64_t)((uint32_t)(idx) & 0xffffffu) << 40))
#define MC_HDR_OP(h) ((uint32_t)((h) & 0xff))
#define MC_HDR_SMALL(h) ((uint32_t)(((h) >> 8) & 0xff))
#define MC_HDR_LEVEL(h) ((uint32_t)(((h) >> 16) & 0xff))
#define MC_HDR_DEC(h) ((uint32_t)(((h) >> 24) & 0xffff))
#define MC_HDR_IDX(h) ((uint32_t)((h) >> 40))

#ifdef __CUDACC__
#define MC_HD __host__ __device__
#else
#define MC_HD
#endif
// number of words of an instruction given its opcode
static MC_HD inline int mc_op_words(uint32_t op) {
    switch (op) {
    case MC_OP_PLANE: return 2;
    case MC_OP_NPLANE: return 5;
    case MC_OP_CONE: return 4;
    case MC_OP_GSPHERE: case MC_OP_GCYL: return 9;
    case MC_OP_GUARD: return 8;
    case MC_OP_BOOL: return 3;
    case MC_OP_APUSH: return 13;
    case MC_OP_APOP: return 2;
    case MC_OP_BOXDIST: return 7;
    case MC_OP_SPHERE: case MC_OP_CYL: return 2;
    case MC_OP_GBOX: return 16;
    case MC_OP_BOX: return 9;
    case MC_OP_GAPUSH: return 20;
    default: return 1;  // END, USQ, ENDG, NEG, CHILD
    }
}

// tile pruning decisions (one byte per decision id per tile)
#define MC_DEC_KEEP 0   // guard: mixed; child: needed
#define MC_DEC_PASS 1   // guard passes at every point of the tile
#define MC_DEC_FAIL 2   // guard fails at every point of the tile / boolean child can be dropped

// evaluator limits (device local-memory stacks); the host lowering rejects deeper programs
#define MC_TILE_PROGRAM_CAP_WORDS 4096u /* shared-memory budget of a tile-specialised program (32 KB) */
#define MC_MAX_VALUE_DEPTH 48
#define MC_MAX_POINT_DEPTH 6
#define MC_MAX_PROGRAM_WORDS (1u << 22)
// lattice points evaluated per lane by the field kernels (shares instruction decode / dispatch)
#define MC_MAX_DECISIONS 65535
