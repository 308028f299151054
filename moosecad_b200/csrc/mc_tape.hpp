// Host side of the scene lowering: the node list recorded by luascad::geom
// (luascad/src/lib.rs:89-281), bounding-box propagation (bbox 0.11.2 / implicit3d 0.14.2
// rules) and the flattening into the device program of mc_program.h.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "mc_program.h"

namespace mc {

struct Box {
    double lo[3], hi[3];
};

struct Mat4 {  // row-major 4x4, last row stays (0,0,0,1) for every constructor below
    double a[4][4];
};

enum NodeKind : int {
    NK_PLANE, NK_NPLANE, NK_SPHERE, NK_CYL, NK_CONE, NK_USQ, NK_NEG, NK_UNION, NK_INTER, NK_AFFINE
};

struct Node {
    NodeKind kind;
    int axis = 0;            // plane
    bool inverted = false;   // plane
    double p[4] = {0, 0, 0, 0};  // plane d | nplane n,p0 | sphere/cyl r | cone slope,offset,mult
    std::vector<int32_t> children;
    double r = 0.0;          // boolean smoothing radius
    Mat4 m;                  // affine
    double scale_min = 1.0;  // affine
    Box bbox;
};

struct Program {
    std::vector<uint64_t> words;
    int value_depth = 0;  // value-stack slots needed
    int point_depth = 0;  // nested affine depth
    uint64_t flops = 0;   // evaluate-all flop model per point
    uint64_t n_exp = 0, n_ln = 0;
    // tile pruning side tables (mc_prune.cuh)
    std::vector<uint32_t> inst_start;   // word index of every instruction, in program order
    std::vector<uint16_t> word_inst;    // instruction index of every word that starts an instruction (else 0xffff)
    uint32_t n_decisions = 0;
    int max_bool_children = 0;
    // register-stack evaluator applicable (mc_vm.cuh, SMALL)
    bool small() const { return value_depth <= 4 && point_depth <= 1 && max_bool_children <= 4; }
};

}  // namespace mc

struct mc_tape {
    std::vector<mc::Node> nodes;
    double fade_range = 0.1;     // src/main.rs:45
    double r_multiplier = 1.0;   // src/main.rs:46

    int32_t push(mc::Node&& n);
    bool valid(int32_t id) const { return id >= 0 && (size_t)id < nodes.size(); }
    // returns MC_OK or a negative status; on error `err` is filled
    int compile(int32_t root, double slack, mc::Program& out, std::string& err) const;
};

namespace mc {
Box box_infinity();
Box box_neg_infinity();
Mat4 mat_identity();
Mat4 mat_mul(const Mat4& a, const Mat4& b);
Mat4 mat_append_translation(const Mat4& m, const double s[3]);
Mat4 mat_append_scaling(const Mat4& m, const double s[3]);
Mat4 mat_euler(double roll, double pitch, double yaw);
bool mat_inverse(const Mat4& m, Mat4& out);
void mat_transform_point(const Mat4& m, const double p[3], double out[3]);
Box box_transform(const Box& b, const Mat4& m);
}  // namespace mc
