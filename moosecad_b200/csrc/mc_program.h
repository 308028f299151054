// Flat SDF program ("op-tape") format shared by the host lowering (mc_tape.cpp) and the
// device evaluator (mc_vm.cuh).
//
// The program is a pre-order walk of the CSG tree (luascad/src/lib.rs:89-281 builds the tree;
// implicit3d 0.14.2 `Object::approx_value` defines the arithmetic).  It is
// an array of 8-byte words.  Every instruction starts with a header word
//     bits  0..7   opcode
//     bits  8..15  small argument (axis | inverted<<2, or child count k)
//     bits 16..31  nesting level of the guard (for the warp-uniform all-pass bit)
//     bits 32..63  a word index (jump target / back pointer), opcode dependent
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
    MC_OP_GSPHERE = 5,  // [hdr bbox[6] slack r]               push guarded sphere
    MC_OP_GCYL = 6,     // [hdr bbox[6] slack r]               push guarded cylinder
    MC_OP_GUARD = 7,    // [hdr(level, jump) bbox[6] slack]    a = bbox.distance(p); if no lane has
                        //                                     a <= slack: push a, jump past ENDG
    MC_OP_ENDG = 8,     // [hdr(level, guard index)]           top = (a <= slack) ? top : a
    MC_OP_NEG = 9,      // [hdr]                               top = -top
    MC_OP_BOOL = 10,    // [hdr(k | union<<7) r exact_range]   pop k, push rvmin / rvmax
    MC_OP_APUSH = 11,   // [hdr m[12] (3x4 row major)]         save p; p = M*p
    MC_OP_APOP = 12,    // [hdr scale_min]                     restore p; top *= scale_min
};

#define MC_HDR(op, small, level, idx)                                                     \
    ((uint64_t)(op) | ((uint64_t)((small) & 0xff) << 8) | ((uint64_t)((level) & 0xffff) << 16) | \
     ((uint64_t)(uint32_t)(idx) << 32))
#define MC_HDR_OP(h) ((uint32_t)((h) & 0xff))
#define MC_HDR_SMALL(h) ((uint32_t)(((h) >> 8) & 0xff))
#define MC_HDR_LEVEL(h) ((uint32_t)(((h) >> 16) & 0xffff))
#define MC_HDR_IDX(h) ((uint32_t)((h) >> 32))

// evaluator limits (device local-memory stacks); the host lowering rejects deeper programs
#define MC_MAX_VALUE_DEPTH 48
#define MC_MAX_POINT_DEPTH 6
