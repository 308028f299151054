// Scene lowering: node list -> bounding boxes -> flat device program.  Host only.
// See mc_tape.hpp.  Arithmetic that the reference performs on the host (bbox propagation,
// matrix composition / inversion, per-node slack) is done here with the same IEEE operation
// order, so the constants the device sees are the ones the reference would compute.
#include "mc_tape.hpp"

#include <cmath>
#include <limits>

#include "../../include/moosecad_b200.h"

namespace mc {

static const double INF = std::numeric_limits<double>::infinity();

Box box_infinity() { return {{-INF, -INF, -INF}, {INF, INF, INF}}; }
Box box_neg_infinity() { return {{INF, INF, INF}, {-INF, -INF, -INF}}; }

Mat4 mat_identity() {
    Mat4 m;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) m.a[i][j] = (i == j) ? 1.0 : 0.0;
    return m;
}

// nalgebra gemv accumulation: start with column 0, then add column k products in order
Mat4 mat_mul(const Mat4& x, const Mat4& y) {
    Mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = x.a[i][0] * y.a[0][j];
            for (int k = 1; k < 4; ++k) acc = x.a[i][k] * y.a[k][j] + acc;
            r.a[i][j] = acc;
        }
    return r;
}

// Matrix4::append_translation: M[i][j] += M[3][j] * shift[i]
Mat4 mat_append_translation(const Mat4& m, const double s[3]) {
    Mat4 r = m;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) r.a[i][j] += r.a[3][j] * s[i];
    return r;
}

// Matrix4::append_nonuniform_scaling: row i *= s[i]
Mat4 mat_append_scaling(const Mat4& m, const double s[3]) {
    Mat4 r = m;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) r.a[i][j] *= s[i];
    return r;
}

// Rotation3::from_euler_angles(roll, pitch, yaw).to_homogeneous()
Mat4 mat_euler(double roll, double pitch, double yaw) {
    const double sr = std::sin(roll), cr = std::cos(roll);
    const double sp = std::sin(pitch), cp = std::cos(pitch);
    const double sy = std::sin(yaw), cy = std::cos(yaw);
    Mat4 m = mat_identity();
    m.a[0][0] = cy * cp;
    m.a[0][1] = cy * sp * sr - sy * cr;
    m.a[0][2] = cy * sp * cr + sy * sr;
    m.a[1][0] = sy * cp;
    m.a[1][1] = sy * sp * sr + cy * cr;
    m.a[1][2] = sy * sp * cr - cy * sr;
    m.a[2][0] = -sp;
    m.a[2][1] = cp * sr;
    m.a[2][2] = cp * cr;
    return m;
}

// 4x4 inverse by cofactor expansion over the column-major element list e[col*4+row]
// (nalgebra `try_inverse` for 4x4): every cofactor is a signed sum of six triple products,
// accumulated left to right; the result is cofactor * (1/det).
static const signed char COF[16][6][4] = {
    {{1, 5, 10, 15}, {-1, 5, 11, 14}, {-1, 9, 6, 15}, {1, 9, 7, 14}, {1, 13, 6, 11}, {-1, 13, 7, 10}},
    {{-1, 1, 10, 15}, {1, 1, 11, 14}, {1, 9, 2, 15}, {-1, 9, 3, 14}, {-1, 13, 2, 11}, {1, 13, 3, 10}},
    {{1, 1, 6, 15}, {-1, 1, 7, 14}, {-1, 5, 2, 15}, {1, 5, 3, 14}, {1, 13, 2, 7}, {-1, 13, 3, 6}},
    {{-1, 1, 6, 11}, {1, 1, 7, 10}, {1, 5, 2, 11}, {-1, 5, 3, 10}, {-1, 9, 2, 7}, {1, 9, 3, 6}},
    {{-1, 4, 10, 15}, {1, 4, 11, 14}, {1, 8, 6, 15}, {-1, 8, 7, 14}, {-1, 12, 6, 11}, {1, 12, 7, 10}},
    {{1, 0, 10, 15}, {-1, 0, 11, 14}, {-1, 8, 2, 15}, {1, 8, 3, 14}, {1, 12, 2, 11}, {-1, 12, 3, 10}},
    {{-1, 0, 6, 15}, {1, 0, 7, 14}, {1, 4, 2, 15}, {-1, 4, 3, 14}, {-1, 12, 2, 7}, {1, 12, 3, 6}},
    {{1, 0, 6, 11}, {-1, 0, 7, 10}, {-1, 4, 2, 11}, {1, 4, 3, 10}, {1, 8, 2, 7}, {-1, 8, 3, 6}},
    {{1, 4, 9, 15}, {-1, 4, 11, 13}, {-1, 8, 5, 15}, {1, 8, 7, 13}, {1, 12, 5, 11}, {-1, 12, 7, 9}},
    {{-1, 0, 9, 15}, {1, 0, 11, 13}, {1, 8, 1, 15}, {-1, 8, 3, 13}, {-1, 12, 1, 11}, {1, 12, 3, 9}},
    {{1, 0, 5, 15}, {-1, 0, 7, 13}, {-1, 4, 1, 15}, {1, 4, 3, 13}, {1, 12, 1, 7}, {-1, 12, 3, 5}},
    {{-1, 0, 5, 11}, {1, 0, 7, 9}, {1, 4, 1, 11}, {-1, 4, 3, 9}, {-1, 8, 1, 7}, {1, 8, 3, 5}},
    {{-1, 4, 9, 14}, {1, 4, 10, 13}, {1, 8, 5, 14}, {-1, 8, 6, 13}, {-1, 12, 5, 10}, {1, 12, 6, 9}},
    {{1, 0, 9, 14}, {-1, 0, 10, 13}, {-1, 8, 1, 14}, {1, 8, 2, 13}, {1, 12, 1, 10}, {-1, 12, 2, 9}},
    {{-1, 0, 5, 14}, {1, 0, 6, 13}, {1, 4, 1, 14}, {-1, 4, 2, 13}, {-1, 12, 1, 6}, {1, 12, 2, 5}},
    {{1, 0, 5, 10}, {-1, 0, 6, 9}, {-1, 4, 1, 10}, {1, 4, 2, 9}, {1, 8, 1, 6}, {-1, 8, 2, 5}},
};

bool mat_inverse(const Mat4& m, Mat4& out) {
    double e[16], o[16];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) e[c * 4 + r] = m.a[r][c];
    for (int k = 0; k < 16; ++k) {
        double acc = 0.0;
        for (int t = 0; t < 6; ++t) {
            const signed char* q = COF[k][t];
            double first = (q[0] < 0) ? -e[q[1]] : e[q[1]];
            double term = first * e[q[2]] * e[q[3]];
            // the leading sign is folded into the first factor (exact), later terms are added
            // or subtracted
            if (t == 0) acc = term;
            else acc = acc + term;
        }
        o[k] = acc;
    }
    // NOTE: x - t == x + (-t) exactly in IEEE arithmetic, so folding the sign is value-preserving.
    double det = e[0] * o[0] + e[1] * o[4] + e[2] * o[8] + e[3] * o[12];
    if (det == 0.0) return false;
    double inv_det = 1.0 / det;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out.a[r][c] = o[c * 4 + r] * inv_det;
    return true;
}

// Matrix4::transform_point: (M3*p + t) / n with n = M[3,0..3].p + M[3,3] (skipped when n == 0)
void mat_transform_point(const Mat4& m, const double p[3], double out[3]) {
    double n = ((m.a[3][0] * p[0] + m.a[3][1] * p[1]) + m.a[3][2] * p[2]) + m.a[3][3];
    for (int i = 0; i < 3; ++i) {
        double acc = m.a[i][0] * p[0];
        acc = m.a[i][1] * p[1] + acc;
        acc = m.a[i][2] * p[2] + acc;
        out[i] = acc + m.a[i][3];
    }
    if (n != 0.0)
        for (int i = 0; i < 3; ++i) out[i] = out[i] / n;
}

// BoundingBox::transform: box of the eight transformed corners
Box box_transform(const Box& b, const Mat4& m) {
    Box r = box_neg_infinity();
    for (int c = 0; c < 8; ++c) {
        double p[3] = {(c & 1) ? b.hi[0] : b.lo[0], (c & 2) ? b.hi[1] : b.lo[1],
                       (c & 4) ? b.hi[2] : b.lo[2]};
        double q[3];
        mat_transform_point(m, p, q);
        for (int k = 0; k < 3; ++k) {
            r.lo[k] = std::fmin(r.lo[k], q[k]);
            r.hi[k] = std::fmax(r.hi[k], q[k]);
        }
    }
    return r;
}

static bool box_is_infinite(const Box& b) {
    for (int k = 0; k < 3; ++k)
        if (b.lo[k] != -INF || b.hi[k] != INF) return false;
    return true;
}

// ---- lowering -------------------------------------------------------------------------
namespace {

struct Lowering {
    const mc_tape& tape;
    Program& prog;
    std::string& err;
    int depth = 0, max_depth = 0;      // value stack
    int pdepth = 0, max_pdepth = 0;    // point stack
    int level = 0;                     // guard nesting
    uint32_t next_dec = 0;             // tile-pruning decision ids
    int lvl(int l) const { return l < 255 ? l : 255; }

    void word(uint64_t w) { prog.words.push_back(w); }
    void dword(double d) {
        uint64_t u;
        static_assert(sizeof(u) == sizeof(d), "");
        __builtin_memcpy(&u, &d, 8);
        prog.words.push_back(u);
    }
    void pushed() {
        ++depth;
        if (depth > max_depth) max_depth = depth;
    }
    void box_words(const Box& b, double slack) {
        for (int k = 0; k < 3; ++k) dword(b.lo[k]);
        for (int k = 0; k < 3; ++k) dword(b.hi[k]);
        dword(slack);
    }

    int emit(int32_t id, double slack) {
        const Node& n = tape.nodes[(size_t)id];
        switch (n.kind) {
        case NK_PLANE:
            word(MC_HDR(MC_OP_PLANE, n.axis | (n.inverted ? 4 : 0), 0, 0, 0));
            dword(n.p[0]);
            prog.flops += 3;
            pushed();
            return MC_OK;
        case NK_NPLANE:
            word(MC_HDR(MC_OP_NPLANE, 0, 0, 0, 0));
            for (int k = 0; k < 4; ++k) dword(n.p[k]);
            prog.flops += 6;
            pushed();
            return MC_OK;
        case NK_USQ:
            word(MC_HDR(MC_OP_USQ, 0, 0, 0, 0));
            prog.flops += 6;
            pushed();
            return MC_OK;
        case NK_CONE:
            word(MC_HDR(MC_OP_CONE, 0, 0, 0, 0));
            dword(n.p[0]); dword(n.p[1]); dword(n.p[2]);
            prog.flops += 8;
            pushed();
            return MC_OK;
        case NK_SPHERE:
        case NK_CYL:
            word(MC_HDR(n.kind == NK_SPHERE ? MC_OP_GSPHERE : MC_OP_GCYL, 0, 0, next_dec++, 0));
            box_words(n.bbox, slack);
            dword(n.p[0]);
            prog.flops += 12 + (n.kind == NK_SPHERE ? 7 : 5);
            pushed();
            return MC_OK;
        case NK_NEG: {
            int rc = emit(n.children[0], slack);
            if (rc != MC_OK) return rc;
            word(MC_HDR(MC_OP_NEG, 0, 0, 0, 0));
            prog.flops += 1;
            return MC_OK;
        }
        case NK_UNION:
        case NK_INTER: {
            const size_t k = n.children.size();
            // geom.cube: Intersection of PlaneX, PlaneY, PlaneZ, PlaneNegX, PlaneNegY, PlaneNegZ
            if (n.kind == NK_INTER && k == 6 && !box_is_infinite(n.bbox)) {
                bool is_cube = true;
                for (size_t i = 0; i < 6; ++i) {
                    const Node& c = tape.nodes[(size_t)n.children[i]];
                    is_cube = is_cube && c.kind == NK_PLANE && c.axis == (int)(i % 3) && c.inverted == (i >= 3);
                }
                if (is_cube) {
                    word(MC_HDR(MC_OP_GBOX, 0x3f, 0, next_dec, 0));  // small = mask of live planes
                    next_dec += 7;  // the guard + one per plane (tile pruning)
                    box_words(n.bbox, slack);
                    for (size_t i = 0; i < 6; ++i) dword(tape.nodes[(size_t)n.children[i]].p[0]);
                    dword(n.r);
                    dword(n.r * tape.r_multiplier);
                    prog.flops += 12 + 6 * 2 + 3 * 6 + (2 * 6 + 4 * 6 + 2);
                    prog.n_exp += 6;
                    prog.n_ln += 1;
                    pushed();
                    return MC_OK;
                }
            }
            if (k > 127) { err = "boolean node with more than 127 children"; return MC_ERR_LIMIT; }
            const bool guarded = !box_is_infinite(n.bbox);
            size_t guard_at = 0;
            uint32_t guard_dec = 0;
            int my_level = 0;
            if (guarded) {
                guard_at = prog.words.size();
                guard_dec = next_dec++;
                word(0);  // patched below
                box_words(n.bbox, slack);
                my_level = level++;
                prog.flops += 12;
            }
            const double child_slack = slack + n.r;  // approx_value(p, slack + self.r)
            const uint32_t dec_base = next_dec;
            next_dec += (uint32_t)k;
            for (size_t ci = 0; ci < k; ++ci) {
                const size_t child_at = prog.words.size();
                word(0);  // CHILD marker, patched with the end of the child
                int rc = emit(n.children[ci], child_slack);
                if (rc != MC_OK) return rc;
                prog.words[child_at] = MC_HDR(MC_OP_CHILD, n.kind == NK_UNION ? 1 : 0, 0, dec_base + ci, prog.words.size());
            }
            if ((int)k > prog.max_bool_children) prog.max_bool_children = (int)k;
            word(MC_HDR(MC_OP_BOOL, (int)k | (n.kind == NK_UNION ? 0x80 : 0), 0, dec_base, 0));
            dword(n.r);
            dword(n.r * tape.r_multiplier);  // exact_range, set_parameters (src/main.rs:53-58)
            depth -= (int)k - 1;
            prog.flops += 3 * k + (2 * k + 4 * k + 2);
            prog.n_exp += k;
            prog.n_ln += 1;
            if (guarded) {
                --level;
                word(MC_HDR(MC_OP_ENDG, 0, lvl(my_level), guard_dec, guard_at));
                prog.words[guard_at] = MC_HDR(MC_OP_GUARD, 0, lvl(my_level), guard_dec, prog.words.size());
            }
            return MC_OK;
        }
        case NK_AFFINE: {
            const size_t guard_at = prog.words.size();
            const uint32_t guard_dec = next_dec++;
            word(0);  // GAPUSH header, patched below
            box_words(n.bbox, slack);
            const int my_level = level++;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 4; ++j) dword(n.m.a[i][j]);
            ++pdepth;
            if (pdepth > max_pdepth) max_pdepth = pdepth;
            int rc = emit(n.children[0], slack / n.scale_min);
            if (rc != MC_OK) return rc;
            --pdepth;
            word(MC_HDR(MC_OP_APOP, 0, 0, 0, 0));
            dword(n.scale_min);
            --level;
            word(MC_HDR(MC_OP_ENDG, 0, lvl(my_level), guard_dec, guard_at));
            prog.words[guard_at] = MC_HDR(MC_OP_GAPUSH, 0, lvl(my_level), guard_dec, prog.words.size());
            prog.flops += 12 + 21 + 1;
            return MC_OK;
        }
        }
        err = "unknown node kind";
        return MC_ERR_ARG;
    }
};

}  // namespace
}  // namespace mc

int32_t mc_tape::push(mc::Node&& n) {
    nodes.push_back(std::move(n));
    return (int32_t)nodes.size() - 1;
}

int mc_tape::compile(int32_t root, double slack, mc::Program& out, std::string& err) const {
    if (!valid(root)) { err = "invalid root node"; return MC_ERR_ARG; }
    out = mc::Program();
    mc::Lowering lw{*this, out, err};
    int rc = lw.emit(root, slack);
    if (rc != MC_OK) return rc;
    out.words.push_back(MC_HDR(MC_OP_END, 0, 0, 0, 0));
    out.value_depth = lw.max_depth;
    out.point_depth = lw.max_pdepth;
    out.n_decisions = lw.next_dec;
    if (out.value_depth > MC_MAX_VALUE_DEPTH) {
        err = "SDF tree needs a deeper value stack than the device evaluator has";
        return MC_ERR_LIMIT;
    }
    if (out.point_depth > MC_MAX_POINT_DEPTH) {
        err = "affine transforms nested deeper than the device evaluator supports";
        return MC_ERR_LIMIT;
    }
    if (out.words.size() > MC_MAX_PROGRAM_WORDS || out.n_decisions > MC_MAX_DECISIONS) {
        err = "SDF program too large for the device evaluator";
        return MC_ERR_LIMIT;
    }
    // instruction table
    out.word_inst.assign(out.words.size(), 0xffff);
    for (size_t pc = 0; pc < out.words.size();) {
        if (out.inst_start.size() < 0xffff) out.word_inst[pc] = (uint16_t)out.inst_start.size();
        out.inst_start.push_back((uint32_t)pc);
        pc += (size_t)mc_op_words(MC_HDR_OP(out.words[pc]));
    }
    return MC_OK;
}
