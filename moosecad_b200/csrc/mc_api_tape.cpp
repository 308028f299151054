// C ABI — scene lowering entry points (include/moosecad_b200.h, "Scene lowering").
// One call per node created in luascad::geom (luascad/src/lib.rs:89-281).
#include <cmath>
#include <cstring>
#include <limits>
#include <string>

#include "../../include/moosecad_b200.h"
#include "mc_error.hpp"
#include "mc_tape.hpp"

using namespace mc;

namespace mc {
thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
}  // namespace mc

extern "C" {

const char* mc_last_error(void) { return mc::g_last_error.c_str(); }
int mc_abi_version(void) { return MC_ABI_VERSION; }

mc_tape* mc_tape_new(void) { return new mc_tape(); }
void mc_tape_free(mc_tape* t) { delete t; }
int32_t mc_tape_num_nodes(const mc_tape* t) { return t ? (int32_t)t->nodes.size() : MC_ERR_ARG; }

int32_t mc_tape_plane(mc_tape* t, int axis, int negative, double d) {
    if (!t || axis < 0 || axis > 2) return fail(MC_ERR_ARG, "mc_tape_plane: bad argument");
    Node n;
    n.kind = NK_PLANE;
    n.axis = axis;
    n.inverted = negative != 0;
    n.p[0] = d;
    n.bbox = box_infinity();
    if (n.inverted) n.bbox.lo[axis] = -d; else n.bbox.hi[axis] = d;
    return t->push(std::move(n));
}

static double norm3(const double v[3]) { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

int32_t mc_tape_plane_hesian(mc_tape* t, double nx, double ny, double nz, double p) {
    if (!t) return fail(MC_ERR_ARG, "mc_tape_plane_hesian: null tape");
    Node n;
    n.kind = NK_NPLANE;
    const double v[3] = {nx, ny, nz};
    const double l = norm3(v);
    n.p[0] = nx / l; n.p[1] = ny / l; n.p[2] = nz / l; n.p[3] = p;
    n.bbox = box_infinity();
    return t->push(std::move(n));
}

int32_t mc_tape_plane_3_points(mc_tape* t, const double a[3], const double b[3], const double c[3]) {
    if (!t || !a || !b || !c) return fail(MC_ERR_ARG, "mc_tape_plane_3_points: bad argument");
    const double u[3] = {a[0] - c[0], a[1] - c[1], a[2] - c[2]};
    const double w[3] = {b[0] - c[0], b[1] - c[1], b[2] - c[2]};
    const double v[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    const double l = norm3(v);
    Node n;
    n.kind = NK_NPLANE;
    n.p[0] = v[0] / l; n.p[1] = v[1] / l; n.p[2] = v[2] / l;
    n.p[3] = (n.p[0] * a[0] + n.p[1] * a[1]) + n.p[2] * a[2];
    n.bbox = box_infinity();
    return t->push(std::move(n));
}

int32_t mc_tape_sphere(mc_tape* t, double r) {
    if (!t) return fail(MC_ERR_ARG, "mc_tape_sphere: null tape");
    Node n;
    n.kind = NK_SPHERE;
    n.p[0] = r;
    n.bbox = {{-r, -r, -r}, {r, r, r}};
    return t->push(std::move(n));
}

int32_t mc_tape_i_cylinder(mc_tape* t, double r) {
    if (!t) return fail(MC_ERR_ARG, "mc_tape_i_cylinder: null tape");
    const double inf = std::numeric_limits<double>::infinity();
    Node n;
    n.kind = NK_CYL;
    n.p[0] = r;
    n.bbox = {{-r, -r, -inf}, {r, r, inf}};
    return t->push(std::move(n));
}

int32_t mc_tape_i_cone(mc_tape* t, double slope, double offset) {
    if (!t) return fail(MC_ERR_ARG, "mc_tape_i_cone: null tape");
    Node n;
    n.kind = NK_CONE;
    n.p[0] = slope;
    n.p[1] = offset;
    n.p[2] = 1.0 / std::sqrt(slope * slope + 1.0);  // distance_multiplier
    n.bbox = box_infinity();
    return t->push(std::move(n));
}

int mc_tape_set_bbox(mc_tape* t, int32_t node, const double b[6]) {
    if (!t || !t->valid(node) || !b) return fail(MC_ERR_ARG, "mc_tape_set_bbox: bad argument");
    for (int k = 0; k < 3; ++k) {
        t->nodes[(size_t)node].bbox.lo[k] = b[k];
        t->nodes[(size_t)node].bbox.hi[k] = b[3 + k];
    }
    return MC_OK;
}

int32_t mc_tape_unit_sphere_sq(mc_tape* t) {
    if (!t) return fail(MC_ERR_ARG, "mc_tape_unit_sphere_sq: null tape");
    Node n;
    n.kind = NK_USQ;
    n.bbox = {{-1, -1, -1}, {1, 1, 1}};
    return t->push(std::move(n));
}

static int32_t copy_node(mc_tape* t, int32_t id) {
    Node n = t->nodes[(size_t)id];
    return t->push(std::move(n));
}

static int32_t make_boolean(mc_tape* t, NodeKind kind, const int32_t* ch, int n, double r, bool negate_rest,
                            const char* who) {
    if (!t || !ch || n <= 0) return fail(MC_ERR_ARG, std::string(who) + ": bad argument");
    for (int i = 0; i < n; ++i)
        if (!t->valid(ch[i])) return fail(MC_ERR_ARG, std::string(who) + ": invalid child id");
    if (n == 1) return copy_node(t, ch[0]);  // from_vec with one element returns it unchanged
    Node b;
    b.kind = kind;
    b.r = r;
    for (int i = 0; i < n; ++i) {
        int32_t c = ch[i];
        if (negate_rest && i > 0) {
            Node neg;
            neg.kind = NK_NEG;
            neg.children = {c};
            neg.bbox = box_infinity();
            c = t->push(std::move(neg));
        }
        b.children.push_back(c);
    }
    if (kind == NK_UNION) {
        Box acc = box_neg_infinity();
        for (int32_t c : b.children) {
            const Box& cb = t->nodes[(size_t)c].bbox;
            for (int k = 0; k < 3; ++k) {
                acc.lo[k] = std::fmin(acc.lo[k], cb.lo[k]);
                acc.hi[k] = std::fmax(acc.hi[k], cb.hi[k]);
            }
        }
        const double d = r * (double)0.2f;  // "dilate by some factor of r", f32 literal widened
        for (int k = 0; k < 3; ++k) { acc.lo[k] -= d; acc.hi[k] += d; }
        b.bbox = acc;
    } else {
        Box acc = box_infinity();
        for (int32_t c : b.children) {
            const Box& cb = t->nodes[(size_t)c].bbox;
            for (int k = 0; k < 3; ++k) {
                acc.lo[k] = std::fmax(acc.lo[k], cb.lo[k]);
                acc.hi[k] = std::fmin(acc.hi[k], cb.hi[k]);
            }
        }
        b.bbox = acc;
    }
    return t->push(std::move(b));
}

int32_t mc_tape_union(mc_tape* t, const int32_t* ch, int n, double r) {
    return make_boolean(t, NK_UNION, ch, n, r, false, "mc_tape_union");
}
int32_t mc_tape_intersection(mc_tape* t, const int32_t* ch, int n, double r) {
    return make_boolean(t, NK_INTER, ch, n, r, false, "mc_tape_intersection");
}
int32_t mc_tape_difference(mc_tape* t, const int32_t* ch, int n, double r) {
    return make_boolean(t, NK_INTER, ch, n, r, true, "mc_tape_difference");
}
int32_t mc_tape_negation(mc_tape* t, int32_t child) {
    if (!t || !t->valid(child)) return fail(MC_ERR_ARG, "mc_tape_negation: bad argument");
    Node neg;
    neg.kind = NK_NEG;
    neg.children = {child};
    neg.bbox = box_infinity();
    return t->push(std::move(neg));
}

// AffineTransformer::new_with_scaler: bbox = inner.bbox().transform(M^-1)
static int32_t make_affine(mc_tape* t, int32_t inner, const Mat4& m, double scale_min, const char* who) {
    Mat4 inv;
    if (!mat_inverse(m, inv)) return fail(MC_ERR_SINGULAR, std::string(who) + ": matrix not invertible");
    Node a;
    a.kind = NK_AFFINE;
    a.children = {inner};
    a.m = m;
    a.scale_min = scale_min;
    a.bbox = box_transform(t->nodes[(size_t)inner].bbox, inv);
    return t->push(std::move(a));
}

int32_t mc_tape_translate(mc_tape* t, int32_t node, double x, double y, double z) {
    if (!t || !t->valid(node)) return fail(MC_ERR_ARG, "mc_tape_translate: bad argument");
    const double s[3] = {-x, -y, -z};
    const Node& n = t->nodes[(size_t)node];
    if (n.kind == NK_AFFINE)
        return make_affine(t, n.children[0], mat_append_translation(n.m, s), n.scale_min, "mc_tape_translate");
    return make_affine(t, node, mat_append_translation(mat_identity(), s), 1.0, "mc_tape_translate");
}

int32_t mc_tape_rotate(mc_tape* t, int32_t node, double x, double y, double z) {
    if (!t || !t->valid(node)) return fail(MC_ERR_ARG, "mc_tape_rotate: bad argument");
    const Node& n = t->nodes[(size_t)node];
    const Mat4 e = mat_euler(x, y, z);
    if (n.kind == NK_AFFINE)
        return make_affine(t, n.children[0], mat_mul(n.m, e), n.scale_min, "mc_tape_rotate");
    return make_affine(t, node, e, 1.0, "mc_tape_rotate");
}

int32_t mc_tape_scale(mc_tape* t, int32_t node, double x, double y, double z) {
    if (!t || !t->valid(node)) return fail(MC_ERR_ARG, "mc_tape_scale: bad argument");
    const double s[3] = {1.0 / x, 1.0 / y, 1.0 / z};
    const double smin = std::fmin(x, std::fmin(y, z));
    const Node& n = t->nodes[(size_t)node];
    if (n.kind == NK_AFFINE)
        return make_affine(t, n.children[0], mat_append_scaling(n.m, s), n.scale_min * smin, "mc_tape_scale");
    return make_affine(t, node, mat_append_scaling(mat_identity(), s), smin, "mc_tape_scale");
}

int32_t mc_tape_bend(mc_tape*, int32_t, double) {
    return fail(MC_ERR_UNSUPPORTED, "geom.bend (implicit3d Bender) is outside the accelerated path");
}
int32_t mc_tape_twist(mc_tape*, int32_t, double) {
    return fail(MC_ERR_UNSUPPORTED, "geom.twist (implicit3d Twister) is outside the accelerated path");
}

int mc_tape_set_parameters(mc_tape* t, double fade_range, double r_multiplier) {
    if (!t) return fail(MC_ERR_ARG, "mc_tape_set_parameters: null tape");
    t->fade_range = fade_range;  // only feeds normal(), which the mesher never calls
    t->r_multiplier = r_multiplier;
    return MC_OK;
}

int mc_tape_bbox(const mc_tape* t, int32_t node, double out[6]) {
    if (!t || !t->valid(node) || !out) return fail(MC_ERR_ARG, "mc_tape_bbox: bad argument");
    const Box& b = t->nodes[(size_t)node].bbox;
    for (int k = 0; k < 3; ++k) { out[k] = b.lo[k]; out[3 + k] = b.hi[k]; }
    return MC_OK;
}

int mc_tape_program_info(const mc_tape* t, int32_t root, double slack, uint64_t* words, int32_t* value_depth,
                         int32_t* point_depth, uint64_t* flops) {
    if (!t) return fail(MC_ERR_ARG, "mc_tape_program_info: null tape");
    Program p;
    std::string err;
    int rc = t->compile(root, slack, p, err);
    if (rc != MC_OK) return fail(rc, err);
    if (words) *words = p.words.size();
    if (value_depth) *value_depth = p.value_depth;
    if (point_depth) *point_depth = p.point_depth;
    if (flops) *flops = p.flops;
    return MC_OK;
}

int64_t mc_tape_program_words(const mc_tape* t, int32_t root, double slack, uint64_t* out, uint64_t capacity) {
    if (!t) return fail(MC_ERR_ARG, "mc_tape_program_words: null tape");
    Program p;
    std::string err;
    int rc = t->compile(root, slack, p, err);
    if (rc != MC_OK) return fail(rc, err);
    if (out) {
        if (capacity < p.words.size()) return fail(MC_ERR_ARG, "mc_tape_program_words: buffer too small");
        for (size_t i = 0; i < p.words.size(); ++i) out[i] = p.words[i];
    }
    return (int64_t)p.words.size();
}

int mc_tape_program_hash(const mc_tape* t, int32_t root, double slack, uint64_t* hash) {
    if (!t || !hash) return fail(MC_ERR_ARG, "mc_tape_program_hash: bad argument");
    Program p;
    std::string err;
    int rc = t->compile(root, slack, p, err);
    if (rc != MC_OK) return fail(rc, err);
    // content hash of everything the device sees: the program words and the root bounding box
    auto mix = [](uint64_t x) {
        x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
        x ^= x >> 27; x *= 0x94d049bb133111ebull;
        x ^= x >> 31;
        return x;
    };
    uint64_t h = 0x9e3779b97f4a7c15ull + p.words.size();
    for (uint64_t w : p.words) h = mix(h + w);
    const Box& bb = t->nodes[(size_t)root].bbox;
    for (int k = 0; k < 3; ++k) {
        uint64_t lo, hi;
        std::memcpy(&lo, &bb.lo[k], 8);
        std::memcpy(&hi, &bb.hi[k], 8);
        h = mix(mix(h + lo) + hi);
    }
    *hash = h;
    return MC_OK;
}

}  // extern "C"
