// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call it, and only as the checker / the CPU baseline.
//
// Single-threaded C++17 restatement of moosecad's meshing hot path,
// keeping the reference's algorithm AND data structures (hash-map vertex / cut-edge
// dedup, recursive boxed SDF tree with a vector per boolean node, swap_remove, ...):
//   tessellate3d/src/lib.rs:36-454   IsosurfaceStuffing (new, cut_lattice, warp_vertices,
//                                    cut_edge, trim_spikes, remove_exterior_tets,
//                                    check_for_inverted_tets, tessellate, Tile)
//   mesh/src/lib.rs:20-24,124-196,292-357,625-679   flip, add_*, volume, aspect_ratio,
//                                    compact, boundary, Tet4 edge/face tables
//   src/main.rs:23-40,45-48,53-58,86-106  ObjectAdaptor, defaults, set_parameters, side-sets
//   luascad/src/lib.rs:89-281        tree shapes built by geom.* (cube, cylinder, booleans)
//
// PARITY STATUS: "parity unpinned" at the implicit3d / bbox / nalgebra boundary.
// The SDF arithmetic lives in crates that are NOT in /root/reference (implicit3d 0.14.2,
// bbox 0.11.2, nalgebra 0.22.1 — Cargo.lock:745-755,120-129,1107-1119) and no Rust
// toolchain exists here, so those parts restate the crates' published algorithm as
// recorded in SURVEY.md Appendix A; each such block is tagged [A].  The reference-owned
// stuffing code is pinned by the reference's own known-answer tests (tests/test_oracle_*.py:
// test_cut_edge, test_tile_volume, test_compact_mesh, test_aspect_ratio,
// test_bcc_tet4_boundary, test_convert_mesh, luascad test_primitives' two asserts).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fno-fast-math).

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "oracle_math.h"

namespace orc {

static int g_libm_mode = 1;  // see oracle_math.h

// Named assumptions about the un-vendored crates, switchable so that the sensitivity
// of every parity configuration to each recollection can be measured (scripts/assumption_sensitivity.py).
// Value 0 is what the oracle (and the device path) implement; the alternatives are the other plausible readings.
enum Assumption : int {
    A_UNION_BBOX_DILATION = 0,   // 0: r * 0.2f32 as f64   1: r * 0.2 (f64)   2: r   3: none
    A_INTERSECTION_BBOX_DILATION,// 0: not dilated         1: dilated like the union
    A_EULER_ORDER,               // 0: Rz(yaw) Ry(pitch) Rx(roll) (nalgebra from_euler_angles)   1: Rx Ry Rz
    A_SMOOTH_LIMIT,              // 0: children with x < m + r enter the exponential sum   1: x < m + exact_range
    A_BOOLEAN_SLACK,             // 0: children get slack + r                             1: slack unchanged
    A_COUNT
};
static int g_assume[A_COUNT] = {0, 0, 0, 0, 0};
static inline double union_dilation(double r) {
    switch (g_assume[A_UNION_BBOX_DILATION]) {
        case 1: return r * 0.2;
        case 2: return r;
        case 3: return 0.0;
        default: return r * (double)0.2f;
    }
}
static inline double m_exp(double x) { return g_libm_mode ? portable_exp(x) : std::exp(x); }
static inline double m_ln(double x) { return g_libm_mode ? portable_log(x) : std::log(x); }

static const double INF = std::numeric_limits<double>::infinity();

struct V3 {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
// [A] nalgebra Vector3 dot / norm: (x*x + y*y) + z*z
static inline double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Rust f64::max / f64::min (num_traits Float) = IEEE maxNum / minNum
static inline double fmax_(double a, double b) { return std::fmax(a, b); }
static inline double fmin_(double a, double b) { return std::fmin(a, b); }

// ---- [A] bbox 0.11.2 -------------------------------------------------------------------
struct BBox {
    V3 min, max;
    static BBox infinity() { return {{-INF, -INF, -INF}, {INF, INF, INF}}; }
    static BBox neg_infinity() { return {{INF, INF, INF}, {-INF, -INF, -INF}}; }
    V3 dim() const { return max - min; }
    double distance(V3 p) const {
        double xval = fmax_(p.x - max.x, min.x - p.x);
        double yval = fmax_(p.y - max.y, min.y - p.y);
        double zval = fmax_(p.z - max.z, min.z - p.z);
        return fmax_(xval, fmax_(yval, zval));
    }
    BBox union_(const BBox& o) const {
        return {{fmin_(min.x, o.min.x), fmin_(min.y, o.min.y), fmin_(min.z, o.min.z)},
                {fmax_(max.x, o.max.x), fmax_(max.y, o.max.y), fmax_(max.z, o.max.z)}};
    }
    BBox intersection(const BBox& o) const {
        return {{fmax_(min.x, o.min.x), fmax_(min.y, o.min.y), fmax_(min.z, o.min.z)},
                {fmin_(max.x, o.max.x), fmin_(max.y, o.max.y), fmin_(max.z, o.max.z)}};
    }
    BBox dilate(double d) const {
        return {{min.x - d, min.y - d, min.z - d}, {max.x + d, max.y + d, max.z + d}};
    }
    void insert(V3 p) {
        min = {fmin_(min.x, p.x), fmin_(min.y, p.y), fmin_(min.z, p.z)};
        max = {fmax_(max.x, p.x), fmax_(max.y, p.y), fmax_(max.z, p.z)};
    }
};

// ---- [A] nalgebra 0.22.1 Matrix4 (column-major storage m[col*4+row]) ---------------------
struct M4 {
    double m[16];
    double& at(int r, int c) { return m[c * 4 + r]; }
    double at(int r, int c) const { return m[c * 4 + r]; }
    static M4 identity() {
        M4 r{};
        for (int i = 0; i < 4; ++i) r.at(i, i) = 1.0;
        return r;
    }
    // self * rhs; gemv column accumulation: ((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + a_i3 b_3j
    M4 mul(const M4& b) const {
        M4 r{};
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) {
                double acc = at(i, 0) * b.at(0, j);
                for (int k = 1; k < 4; ++k) acc = at(i, k) * b.at(k, j) + acc;
                r.at(i, j) = acc;
            }
        return r;
    }
    // append_translation: M[i][j] += M[3][j] * shift[i]
    M4 append_translation(V3 s) const {
        M4 r = *this;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) r.at(i, j) += r.at(3, j) * s[i];
        return r;
    }
    // append_nonuniform_scaling: row i *= s[i]
    M4 append_nonuniform_scaling(V3 s) const {
        M4 r = *this;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) r.at(i, j) *= s[i];
        return r;
    }
    // transform_point: (M3*p + t) / n, n = M[3,0..3].p + M[3,3] (skipped when n == 0)
    V3 transform_point(V3 p) const {
        double n = ((at(3, 0) * p.x + at(3, 1) * p.y) + at(3, 2) * p.z) + at(3, 3);
        V3 r;
        double* o = &r.x;
        for (int i = 0; i < 3; ++i) {
            double acc = at(i, 0) * p.x;
            acc = at(i, 1) * p.y + acc;
            acc = at(i, 2) * p.z + acc;
            o[i] = acc + at(i, 3);
        }
        if (n != 0.0) r = r / n;
        return r;
    }
    // try_inverse: 4x4 cofactor expansion, out = cofactor * (1/det)
    bool try_inverse(M4& out) const {
        const double* a = m;
        double* o = out.m;
        o[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] +
               a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
        o[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] -
               a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
        o[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] +
               a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
        o[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] -
               a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
        o[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] -
               a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
        o[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] +
               a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
        o[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] -
               a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
        o[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] +
               a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
        o[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] +
               a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
        o[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] -
               a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
        o[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] +
                a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
        o[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] -
                a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
        o[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] -
                a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
        o[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] +
                a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
        o[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] -
                a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
        o[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] +
                a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
        double det = a[0] * o[0] + a[1] * o[4] + a[2] * o[8] + a[3] * o[12];
        if (det == 0.0) return false;
        double inv_det = 1.0 / det;
        for (int i = 0; i < 16; ++i) o[i] *= inv_det;
        return true;
    }
};

// [A] bbox.transform(M): box of the 8 transformed corners
static BBox bbox_transform(const BBox& b, const M4& m) {
    BBox r = BBox::neg_infinity();
    for (int c = 0; c < 8; ++c) {
        V3 p = {(c & 1) ? b.max.x : b.min.x, (c & 2) ? b.max.y : b.min.y,
                (c & 4) ? b.max.z : b.min.z};
        r.insert(m.transform_point(p));
    }
    return r;
}

// [A] Rotation3::from_euler_angles(roll, pitch, yaw).to_homogeneous()
static M4 euler_matrix(V3 r) {
    double sr = std::sin(r.x), cr = std::cos(r.x);
    double sp = std::sin(r.y), cp = std::cos(r.y);
    double sy = std::sin(r.z), cy = std::cos(r.z);
    M4 m = M4::identity();
    if (g_assume[A_EULER_ORDER]) {  // [A] the other composition order: Rx(roll) Ry(pitch) Rz(yaw)
        m.at(0, 0) = cp * cy;                m.at(0, 1) = -cp * sy;               m.at(0, 2) = sp;
        m.at(1, 0) = sr * sp * cy + cr * sy; m.at(1, 1) = -sr * sp * sy + cr * cy; m.at(1, 2) = -sr * cp;
        m.at(2, 0) = -cr * sp * cy + sr * sy; m.at(2, 1) = cr * sp * sy + sr * cy; m.at(2, 2) = cr * cp;
        return m;
    }
    m.at(0, 0) = cy * cp;
    m.at(0, 1) = cy * sp * sr - sy * cr;
    m.at(0, 2) = cy * sp * cr + sy * sr;
    m.at(1, 0) = sy * cp;
    m.at(1, 1) = sy * sp * sr + cy * cr;
    m.at(1, 2) = sy * sp * cr - cy * sr;
    m.at(2, 0) = -sp;
    m.at(2, 1) = cp * sr;
    m.at(2, 2) = cp * cr;
    return m;
}

// ---- [A] implicit3d 0.14.2 objects -----------------------------------------------------
struct Object;
using Obj = std::shared_ptr<Object>;

struct Object {
    BBox bbox_ = BBox::infinity();
    virtual ~Object() {}
    virtual double approx_value(V3 p, double slack) const = 0;
    virtual void set_parameters(double /*fade_range*/, double /*r_multiplier*/) {}
    virtual Obj clone() const = 0;
    // affine ops return a fresh AffineTransformer wrapping a clone (Object::translate etc.)
    virtual Obj translate(V3 v) const;
    virtual Obj rotate(V3 r) const;
    virtual Obj scale(V3 s) const;
};

// rvmin / rvmax: exponential smooth min / max, only evaluated when two children are close
static double rvmin(const std::vector<double>& v, double r, double exact_range) {
    bool close_min = false;
    double minimum = INF;
    for (double x : v) {
        if (x < minimum) {
            close_min = (minimum - x) < exact_range;
            minimum = x;
        } else if ((x - minimum) < exact_range) {
            close_min = true;
        }
    }
    if (!close_min) return minimum;
    double min_plus_r = minimum + (g_assume[A_SMOOTH_LIMIT] ? exact_range : r);   // [A] A_SMOOTH_LIMIT
    double r4 = r / 4.0;
    double exp_sum = 0.0;
    for (double x : v)
        if (x < min_plus_r) exp_sum = exp_sum + m_exp(-x / r4);
    return m_ln(exp_sum) * -r4;
}
static double rvmax(const std::vector<double>& v, double r, double exact_range) {
    bool close_max = false;
    double maximum = -INF;
    for (double x : v) {
        if (x > maximum) {
            close_max = (x - maximum) < exact_range;
            maximum = x;
        } else if ((maximum - x) < exact_range) {
            close_max = true;
        }
    }
    if (!close_max) return maximum;
    double max_minus_r = maximum - (g_assume[A_SMOOTH_LIMIT] ? exact_range : r);  // [A] A_SMOOTH_LIMIT
    double r4 = r / 4.0;
    double exp_sum = 0.0;
    for (double x : v)
        if (x > max_minus_r) exp_sum = exp_sum + m_exp(x / r4);
    return m_ln(exp_sum) * r4;
}

struct Plane : Object {  // PlaneX/Y/Z (inverted=false) and PlaneNegX/Y/Z (inverted=true)
    int axis;
    bool inverted;
    double d;
    Plane(int axis_, bool inv, double d_) : axis(axis_), inverted(inv), d(d_) {
        double* lo = &bbox_.min.x;
        double* hi = &bbox_.max.x;
        if (inverted) lo[axis] = -d; else hi[axis] = d;
    }
    double approx_value(V3 p, double) const override {
        double q = p[axis];
        double ap = std::fabs(q);
        bool sign_positive = !std::signbit(q);
        if (sign_positive != inverted) return ap - d;
        return -ap - d;
    }
    Obj clone() const override { return std::make_shared<Plane>(*this); }
};

struct NormalPlane : Object {
    V3 n;
    double p0;
    NormalPlane(V3 n_, double p_) : n(n_), p0(p_) {}
    double approx_value(V3 p, double) const override { return dot(p, n) - p0; }
    Obj clone() const override { return std::make_shared<NormalPlane>(*this); }
};

struct Sphere : Object {
    double r;
    explicit Sphere(double r_) : r(r_) { bbox_ = {{-r, -r, -r}, {r, r, r}}; }
    double approx_value(V3 p, double slack) const override {
        double approx = bbox_.distance(p);
        if (approx <= slack) return norm(p) - r;
        return approx;
    }
    Obj clone() const override { return std::make_shared<Sphere>(*this); }
};

struct Cylinder : Object {
    double r;
    explicit Cylinder(double r_) : r(r_) { bbox_ = {{-r, -r, -INF}, {r, r, INF}}; }
    double approx_value(V3 p, double slack) const override {
        double approx = bbox_.distance(p);
        if (approx <= slack) return std::sqrt(p.x * p.x + p.y * p.y) - r;
        return approx;
    }
    Obj clone() const override { return std::make_shared<Cylinder>(*this); }
};

struct Cone : Object {
    double slope, offset, distance_multiplier;
    Cone(double s, double o)
        : slope(s), offset(o), distance_multiplier(1.0 / std::sqrt(s * s + 1.0)) {}
    double approx_value(V3 p, double) const override {
        double radius = std::fabs(slope * (p.z + offset));
        return (std::sqrt(p.x * p.x + p.y * p.y) - radius) * distance_multiplier;
    }
    Obj clone() const override { return std::make_shared<Cone>(*this); }
};

// Test fixture of the reference's own tests: tessellate3d/src/lib.rs:460-484 `UnitSphere`,
// value = |p|^2 - 1 (magnitude_squared), bbox [-1,1]^3, no bbox guard.
struct UnitSphereSq : Object {
    UnitSphereSq() { bbox_ = {{-1, -1, -1}, {1, 1, 1}}; }
    double approx_value(V3 p, double) const override { return dot(p, p) - 1.0; }
    Obj clone() const override { return std::make_shared<UnitSphereSq>(*this); }
};

struct Negation : Object {
    Obj o;
    explicit Negation(Obj o_) : o(std::move(o_)) {}
    double approx_value(V3 p, double slack) const override { return -o->approx_value(p, slack); }
    void set_parameters(double f, double m) override { o->set_parameters(f, m); }
    Obj clone() const override { return std::make_shared<Negation>(o->clone()); }
};

struct Boolean : Object {
    bool is_union;
    std::vector<Obj> objs;
    double r, exact_range;
    Boolean(bool u, std::vector<Obj> v, double r_) : is_union(u), objs(std::move(v)), r(r_) {
        exact_range = r * 1.0;  // R_MULTIPLIER default
        if (is_union) {
            BBox b = BBox::neg_infinity();
            for (auto& o : objs) b = b.union_(o->bbox_);
            bbox_ = b.dilate(union_dilation(r));               // [A] A_UNION_BBOX_DILATION
        } else {
            BBox b = BBox::infinity();
            for (auto& o : objs) b = b.intersection(o->bbox_);
            bbox_ = g_assume[A_INTERSECTION_BBOX_DILATION] ? b.dilate(union_dilation(r)) : b;   // [A]
        }
    }
    double approx_value(V3 p, double slack) const override {
        double approx = bbox_.distance(p);
        if (approx <= slack) {
            std::vector<double> vals;  // the reference collects into a Vec per call
            vals.reserve(objs.size());
            for (auto& o : objs) vals.push_back(o->approx_value(p, g_assume[A_BOOLEAN_SLACK] ? slack : slack + r));   // [A]
            return is_union ? rvmin(vals, r, exact_range) : rvmax(vals, r, exact_range);
        }
        return approx;
    }
    void set_parameters(double f, double m) override {
        exact_range = r * m;
        for (auto& o : objs) o->set_parameters(f, m);
    }
    Obj clone() const override {
        std::vector<Obj> v;
        for (auto& o : objs) v.push_back(o->clone());
        auto b = std::make_shared<Boolean>(is_union, std::move(v), r);
        b->exact_range = exact_range;
        b->bbox_ = bbox_;
        return b;
    }
};

struct Affine : Object {
    Obj o;
    M4 t;
    double scale_min;
    bool ok = true;
    Affine(Obj o_, const M4& t_, double smin) : o(std::move(o_)), t(t_), scale_min(smin) {
        M4 inv;
        if (t.try_inverse(inv)) bbox_ = bbox_transform(o->bbox_, inv);
        else ok = false;  // reference panics
    }
    double approx_value(V3 p, double slack) const override {
        double approx = bbox_.distance(p);
        if (approx <= slack)
            return o->approx_value(t.transform_point(p), slack / scale_min) * scale_min;
        return approx;
    }
    void set_parameters(double f, double m) override { o->set_parameters(f, m); }
    Obj clone() const override { return std::make_shared<Affine>(o->clone(), t, scale_min); }
    Obj translate(V3 v) const override {
        return std::make_shared<Affine>(o->clone(), t.append_translation({-v.x, -v.y, -v.z}),
                                        scale_min);
    }
    Obj rotate(V3 r) const override {
        return std::make_shared<Affine>(o->clone(), t.mul(euler_matrix(r)), scale_min);
    }
    Obj scale(V3 s) const override {
        return std::make_shared<Affine>(
            o->clone(), t.append_nonuniform_scaling({1.0 / s.x, 1.0 / s.y, 1.0 / s.z}),
            scale_min * fmin_(s.x, fmin_(s.y, s.z)));
    }
};

Obj Object::translate(V3 v) const {
    return std::make_shared<Affine>(clone(),
                                    M4::identity().append_translation({-v.x, -v.y, -v.z}), 1.0);
}
Obj Object::rotate(V3 r) const { return std::make_shared<Affine>(clone(), euler_matrix(r), 1.0); }
Obj Object::scale(V3 s) const {
    return std::make_shared<Affine>(
        clone(), M4::identity().append_nonuniform_scaling({1.0 / s.x, 1.0 / s.y, 1.0 / s.z}),
        fmin_(s.x, fmin_(s.y, s.z)));
}

// ---- mesh crate (mesh/src/lib.rs) -------------------------------------------------------
using Tet = std::array<size_t, 4>;
static const int TET_EDGE[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};  // :651
static const int TET_FACE[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};      // :662

struct Mesh {
    std::vector<V3> coords;
    std::vector<Tet> elems;
    size_t add_vertex(V3 p) { coords.push_back(p); return coords.size() - 1; }   // :124-128
    void add_elem(Tet t) { elems.push_back(t); }                                 // :159-165
    void elem_swap_remove(size_t i) { elems[i] = elems.back(); elems.pop_back(); }  // :191
    // :292-298, nalgebra Matrix3::determinant (columns a-d, b-d, c-d) [A]
    double volume(const Tet& t) const {
        V3 c0 = coords[t[0]] - coords[t[3]];
        V3 c1 = coords[t[1]] - coords[t[3]];
        V3 c2 = coords[t[2]] - coords[t[3]];
        double m11 = c0.x, m21 = c0.y, m31 = c0.z;
        double m12 = c1.x, m22 = c1.y, m32 = c1.z;
        double m13 = c2.x, m23 = c2.y, m33 = c2.z;
        double minor_m12_m23 = m22 * m33 - m32 * m23;
        double minor_m11_m23 = m21 * m33 - m31 * m23;
        double minor_m11_m22 = m21 * m32 - m31 * m22;
        double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
        return det / 6.0;
    }
    V3 diff(size_t a, size_t b) const { return coords[b] - coords[a]; }  // :220-224
    // :300-316
    double aspect_ratio(const Tet& t) const {
        double max_length2 = 0.0, sum_normals = 0.0;
        V3 ab = diff(t[0], t[1]), ac = diff(t[0], t[2]), ad = diff(t[0], t[3]);
        double alpha = std::fabs(dot(ab, cross(ac, ad)));
        for (int f = 0; f < 4; ++f) {
            V3 a = coords[t[TET_FACE[f][0]]], b = coords[t[TET_FACE[f][1]]],
               c = coords[t[TET_FACE[f][2]]];
            sum_normals += norm(cross(b - a, c - a));
        }
        for (int e = 0; e < 6; ++e) {
            V3 d = diff(t[TET_EDGE[e][0]], t[TET_EDGE[e][1]]);
            max_length2 = fmax_(max_length2, dot(d, d));
        }
        double max_length = std::sqrt(max_length2);
        double radius = alpha / sum_normals;
        return 1.0 / (max_length / radius / (2.0 * std::sqrt(6.0)));
    }
    // :318-339
    void compact() {
        std::vector<int64_t> remap(coords.size(), -1);
        int64_t nv = 0;
        for (auto& e : elems)
            for (auto& n : e) {
                if (remap[n] < 0) remap[n] = nv++;
                n = (size_t)remap[n];
            }
        std::vector<V3> nc((size_t)nv, V3{0, 0, 0});
        for (size_t i = 0; i < coords.size(); ++i)
            if (remap[i] >= 0) nc[(size_t)remap[i]] = coords[i];
        coords.swap(nc);
    }
};

struct TripleHash {
    size_t operator()(const std::array<size_t, 3>& k) const {
        uint64_t h = 1469598103934665603ull;
        for (size_t v : k) { h ^= v; h *= 1099511628211ull; h ^= h >> 29; }
        return (size_t)h;
    }
};
struct PairHash {
    size_t operator()(const std::pair<size_t, size_t>& k) const {
        uint64_t h = k.first * 0x9E3779B97F4A7C15ull;
        h ^= (k.second + 0x7F4A7C159E3779B9ull) + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

// mesh/src/lib.rs:341-357: faces toggled in/out of a hash map keyed by the sorted node triple
static std::vector<std::pair<size_t, size_t>> boundary(const Mesh& m) {
    std::unordered_map<std::array<size_t, 3>, std::pair<size_t, size_t>, TripleHash> side_set;
    side_set.reserve(m.elems.size() * 2);
    for (size_t e = 0; e < m.elems.size(); ++e)
        for (size_t s = 0; s < 4; ++s) {
            std::array<size_t, 3> nodes = {m.elems[e][TET_FACE[s][0]], m.elems[e][TET_FACE[s][1]],
                                           m.elems[e][TET_FACE[s][2]]};
            std::sort(nodes.begin(), nodes.end());
            auto it = side_set.find(nodes);
            if (it != side_set.end()) side_set.erase(it);
            else side_set.emplace(nodes, std::make_pair(e, s));
        }
    std::vector<std::pair<size_t, size_t>> out;
    out.reserve(side_set.size());
    for (auto& kv : side_set) out.push_back(kv.second);
    std::sort(out.begin(), out.end());  // the reference's HashSet order is arbitrary
    return out;
}

// ---- tessellate3d (tessellate3d/src/lib.rs) ----------------------------------------------
static const int TILE_LATTICE[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                       {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};  // :437-446
static const int TILE_TETS[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4},
                                    {2, 6, 5, 7}, {0, 2, 7, 3}};                      // :447-453

struct KeyHash {
    size_t operator()(const std::array<size_t, 3>& k) const { return TripleHash()(k); }
};

struct Stats {
    uint64_t dim[3] = {0, 0, 0};
    uint64_t ntets_stage[3] = {0, 0, 0};  // after cut_lattice, before trim (same), after trim
    uint64_t stats[7] = {0, 0, 0, 0, 0, 0, 0};
    uint64_t n_warped = 0, n_cut = 0, n_value_calls = 0, n_bisect_iters = 0;
    uint64_t n_exterior_removed = 0, n_exterior_skipped_by_swap = 0, inverted = 0;
    double ar_min = 1.0, ar_max = 0.0;
    double t_cut = 0, t_warp = 0, t_trim = 0, t_ext = 0, t_compact = 0, t_check = 0,
           t_boundary = 0, t_sidesets = 0;
};

struct Stuffing {
    const Object* function;
    double value_slack;  // ObjectAdaptor.tessellation_resolution, src/main.rs:33-35
    double resolution, error_threshold, warp_threshold;
    size_t dim[3];
    Mesh mesh;
    std::vector<double> phi;
    std::unordered_map<std::pair<size_t, size_t>, size_t, PairHash> cut_map;
    Stats st;
    bool ok = true;
    uint64_t max_bisect = 100000;

    double value(V3 p) {
        ++st.n_value_calls;
        return function->approx_value(p, value_slack);
    }

    // :36-52
    Stuffing(const Object* f, double slack, double res, double rel_err)
        : function(f), value_slack(slack), resolution(res), error_threshold(res * rel_err),
          warp_threshold(0.3) {
        V3 d = f->bbox_.dim();
        for (int k = 0; k < 3; ++k) {
            double c = std::ceil(d[k] / res);
            if (!(c >= 0.0) || !(c < 1.8446744073709552e19)) { ok = false; dim[k] = 0; }  // to_usize().unwrap()
            else dim[k] = (size_t)c;
        }
    }

    // :55-135
    void cut_lattice() {
        const BBox& bbox = function->bbox_;
        V3 origin = bbox.min;
        std::unordered_map<std::array<size_t, 3>, size_t, KeyHash> lattice_map;
        bool fx = true;
        for (size_t x = 0; x < dim[0]; ++x) {
            fx = !fx;
            bool fy = true;
            for (size_t y = 0; y < dim[1]; ++y) {
                fy = !fy;
                bool fz = true;
                for (size_t z = 0; z < dim[2]; ++z) {
                    fz = !fz;
                    for (int t = 0; t < 5; ++t) {
                        std::array<size_t, 3> coordsi[4];
                        V3 coordss[4];
                        for (int n = 0; n < 4; ++n) {
                            const int* l = TILE_LATTICE[TILE_TETS[t][n]];
                            size_t cx = (size_t)(fx ? 1 - l[0] : l[0]) + x;
                            size_t cy = (size_t)(fy ? 1 - l[1] : l[1]) + y;
                            size_t cz = (size_t)(fz ? 1 - l[2] : l[2]) + z;
                            coordsi[n] = {cx, cy, cz};
                            coordss[n] = {(double)cx * resolution + origin.x,
                                          (double)cy * resolution + origin.y,
                                          (double)cz * resolution + origin.z};
                        }
                        bool intersects_f = false;
                        for (int n = 0; n < 4; ++n)
                            if (value(coordss[n]) <= 0.0) { intersects_f = true; break; }
                        if (!intersects_f) continue;
                        Tet tet{};
                        for (int n = 0; n < 4; ++n) {
                            auto it = lattice_map.find(coordsi[n]);
                            if (it != lattice_map.end()) {
                                tet[n] = it->second;
                            } else {
                                tet[n] = mesh.coords.size();
                                lattice_map.emplace(coordsi[n], tet[n]);
                                mesh.add_vertex(coordss[n]);
                                phi.push_back(value(coordss[n]));
                            }
                        }
                        if (fx ^ fy ^ fz) std::swap(tet[0], tet[1]);
                        mesh.add_elem(tet);
                    }
                }
            }
        }
    }

    // :138-181
    void warp_vertices() {
        size_t size = mesh.coords.size();
        std::vector<char> has(size, 0);
        std::vector<double> warp(size, 0.0);
        std::vector<V3> delta(size, V3{0, 0, 0});
        for (const Tet& tet : mesh.elems) {
            for (int e = 0; e < 6; ++e) {
                size_t n0 = tet[TET_EDGE[e][0]], n1 = tet[TET_EDGE[e][1]];
                double phi0 = phi[n0], phi1 = phi[n1];
                V3 x0 = mesh.coords[n0], x1 = mesh.coords[n1];
                if ((phi0 > 0.0 && phi1 > 0.0) || (phi0 < 0.0 && phi1 < 0.0)) continue;
                double alpha = phi0 / (phi0 - phi1);
                if (alpha < warp_threshold) {
                    double d = alpha * norm(x1 - x0);
                    if (!has[n0] || d < warp[n0]) {
                        has[n0] = 1; warp[n0] = d;
                        delta[n0] = (x1 - x0) * alpha;
                    }
                } else if (alpha > 1.0 - warp_threshold) {
                    double d = (1.0 - alpha) * norm(x1 - x0);
                    if (!has[n1] || d < warp[n1]) {
                        has[n1] = 1; warp[n1] = d;
                        delta[n1] = (x0 - x1) * (1.0 - alpha);
                    }
                }
            }
        }
        for (size_t i = 0; i < size; ++i)
            if (has[i]) {
                mesh.coords[i] = mesh.coords[i] + delta[i];
                phi[i] = 0.0;
                ++st.n_warped;
            }
    }

    // :185-217
    size_t cut_edge(size_t a, size_t b) {
        std::pair<size_t, size_t> edge(std::min(a, b), std::max(a, b));
        auto it = cut_map.find(edge);
        if (it != cut_map.end()) return it->second;
        double phia = phi[a];
        V3 pa = mesh.coords[a], pb = mesh.coords[b];
        V3 c;
        uint64_t iters = 0;
        for (;;) {
            c = (pa + pb) * 0.5;
            double p = value(c);
            ++st.n_bisect_iters;
            if (std::fabs(p) <= error_threshold) break;
            else if (p > 0.0 && phia > 0.0) { pa = c; phia = p; }
            else pb = c;
            if (++iters > max_bisect) { ok = false; break; }  // reference would spin forever
        }
        size_t index = mesh.coords.size();
        mesh.add_vertex(c);
        phi.push_back(0.0);
        cut_map.emplace(edge, index);
        ++st.n_cut;
        return index;
    }

    // :242-367
    void trim_spikes() {
        struct NewTet { Tet t; bool flipped; };
        std::vector<NewTet> new_tets;
        size_t ti = 0;
        st.ntets_stage[1] = mesh.elems.size();
        while (ti < mesh.elems.size()) {
            size_t a = mesh.elems[ti][0], b = mesh.elems[ti][1], c = mesh.elems[ti][2],
                   d = mesh.elems[ti][3];
            if (phi[a] <= 0.0 && phi[b] <= 0.0 && phi[c] <= 0.0 && phi[d] <= 0.0) { ++ti; continue; }
            bool flipped = false;
            if (phi[a] < phi[b] || (phi[a] == phi[b] && a < b)) { std::swap(a, b); flipped = !flipped; }
            if (phi[c] < phi[d] || (phi[c] == phi[d] && c < d)) { std::swap(c, d); flipped = !flipped; }
            if (phi[a] < phi[c] || (phi[a] == phi[c] && a < c)) { std::swap(a, c); flipped = !flipped; }
            if (phi[b] < phi[d] || (phi[b] == phi[d] && b < d)) { std::swap(b, d); flipped = !flipped; }
            if (phi[b] < phi[c] || (phi[b] == phi[c] && b < c)) { std::swap(b, c); flipped = !flipped; }
            if (phi[d] == 0.0) {
                mesh.elem_swap_remove(ti);
                st.stats[0]++;
            } else if (phi[c] > 0.0) {
                mesh.elem_swap_remove(ti);
                st.stats[1]++;
                size_t ad = cut_edge(a, d), bd = cut_edge(b, d), cd = cut_edge(c, d);
                new_tets.push_back({{ad, bd, cd, d}, flipped});
            } else if (phi[b] < 0.0) {
                mesh.elem_swap_remove(ti);
                st.stats[2]++;
                size_t ab = cut_edge(a, b), ac = cut_edge(a, c), ad = cut_edge(a, d);
                new_tets.push_back({{ab, b, c, d}, flipped});
                new_tets.push_back({{c, ab, ac, d}, !flipped});
                new_tets.push_back({{ab, d, ac, ad}, flipped});
            } else if (phi[b] > 0.0 && phi[c] < 0.0) {
                mesh.elem_swap_remove(ti);
                st.stats[3]++;
                size_t ac = cut_edge(a, c), ad = cut_edge(a, d), bc = cut_edge(b, c),
                       bd = cut_edge(b, d);
                new_tets.push_back({{ac, bd, c, d}, flipped});
                new_tets.push_back({{ac, bd, bc, d}, flipped});
                new_tets.push_back({{ac, ad, bd, d}, flipped});
            } else if (phi[b] == 0.0 && phi[c] == 0.0) {
                mesh.elem_swap_remove(ti);
                st.stats[4]++;
                size_t ad = cut_edge(a, d);
                new_tets.push_back({{ad, b, c, d}, flipped});
            } else if (phi[b] == 0.0) {
                mesh.elem_swap_remove(ti);
                st.stats[5]++;
                size_t ac = cut_edge(a, c), ad = cut_edge(a, d);
                new_tets.push_back({{b, ac, c, d}, !flipped});
                new_tets.push_back({{ad, b, ac, d}, flipped});
            } else {
                mesh.elem_swap_remove(ti);
                st.stats[6]++;
                size_t ad = cut_edge(a, d), bd = cut_edge(b, d);
                new_tets.push_back({{ad, bd, c, d}, flipped});
            }
        }
        n_new_tets = new_tets.size();
        for (auto& nt : new_tets) {
            Tet t = nt.t;
            if (nt.flipped) std::swap(t[0], t[1]);
            mesh.add_elem(t);
        }
        st.ntets_stage[2] = mesh.elems.size();
    }
    size_t n_new_tets = 0;

    // :370-398 — literal, including `i += 1` after swap_remove (the swapped-in tet is skipped)
    void remove_exterior_tets() {
        size_t i = 0;
        while (i < mesh.elems.size()) {
            const Tet& e = mesh.elems[i];
            double pa = phi[e[0]], pb = phi[e[1]], pc = phi[e[2]], pd = phi[e[3]];
            if (pa == 0.0 && pb == 0.0 && pc == 0.0 && pd == 0.0) {
                V3 p = ((mesh.coords[e[0]] + mesh.coords[e[1]]) + mesh.coords[e[2]]) +
                       mesh.coords[e[3]];
                if (value(p / 4.0) > 0.0) {
                    // would the swapped-in element itself have been a candidate?
                    const Tet& last = mesh.elems.back();
                    if (&last != &e && phi[last[0]] == 0.0 && phi[last[1]] == 0.0 &&
                        phi[last[2]] == 0.0 && phi[last[3]] == 0.0)
                        ++st.n_exterior_skipped_by_swap;
                    mesh.elem_swap_remove(i);
                    ++st.n_exterior_removed;
                }
            }
            i += 1;
        }
    }

    // :400-418
    void check_for_inverted_tets() {
        double mn = 1.0, mx = 0.0;
        for (const Tet& t : mesh.elems) {
            double v = mesh.volume(t);
            double ar = mesh.aspect_ratio(t);
            if (v <= 0.0) ++st.inverted;
            mn = fmin_(mn, ar);
            mx = fmax_(mx, ar);
        }
        st.ar_min = mn;
        st.ar_max = mx;
    }

    static double now() {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
            .count();
    }

    // :420-431
    void tessellate() {
        st.dim[0] = dim[0]; st.dim[1] = dim[1]; st.dim[2] = dim[2];
        double t0 = now();
        cut_lattice();
        st.ntets_stage[0] = mesh.elems.size();
        double t1 = now();
        warp_vertices();
        double t2 = now();
        trim_spikes();
        double t3 = now();
        remove_exterior_tets();
        double t4 = now();
        mesh.compact();
        double t5 = now();
        check_for_inverted_tets();
        double t6 = now();
        st.t_cut = t1 - t0; st.t_warp = t2 - t1; st.t_trim = t3 - t2; st.t_ext = t4 - t3;
        st.t_compact = t5 - t4; st.t_check = t6 - t5;
    }
};

// ---- scene builder used through the C entry points ---------------------------------------
struct Scene {
    std::vector<Obj> nodes;
    int push(Obj o) { nodes.push_back(std::move(o)); return (int)nodes.size() - 1; }
};

struct MeshResult {
    Mesh mesh;
    Stats st;
    bool ok = true;
    std::vector<std::pair<size_t, size_t>> surface;
    bool have_surface = false;
    std::vector<std::vector<std::pair<size_t, size_t>>> side_sets;
};

}  // namespace orc

using namespace orc;

extern "C" {

void orc_set_libm_mode(int mode) { g_libm_mode = mode; }
// named Appendix-A assumptions (see enum Assumption); returns the previous value, -1 for a bad id
int orc_set_assumption(int id, int value) {
    if (id < 0 || id >= A_COUNT) return -1;
    const int old = g_assume[id];
    g_assume[id] = value;
    return old;
}
int orc_num_assumptions() { return A_COUNT; }
int orc_get_libm_mode() { return g_libm_mode; }
double orc_exp(double x) { return m_exp(x); }
double orc_ln(double x) { return m_ln(x); }

void* orc_scene_new() { return new Scene(); }
void orc_scene_free(void* s) { delete (Scene*)s; }

// luascad/src/lib.rs:89-123
int orc_plane(void* s, int axis, int negative, double d) {
    return ((Scene*)s)->push(std::make_shared<Plane>(axis, negative != 0, d));
}
// :125-132 — NormalPlane::from_normal_and_p normalises n [A]
int orc_plane_hesian(void* s, double nx, double ny, double nz, double p) {
    V3 n = {nx, ny, nz};
    double l = norm(n);
    return ((Scene*)s)->push(std::make_shared<NormalPlane>(n / l, p));
}
// :134-152 — from_3_points [A]
int orc_plane_3_points(void* s, const double* a, const double* b, const double* c) {
    V3 A = {a[0], a[1], a[2]}, B = {b[0], b[1], b[2]}, C = {c[0], c[1], c[2]};
    V3 v = cross(A - C, B - C);
    V3 n = v / norm(v);
    return ((Scene*)s)->push(std::make_shared<NormalPlane>(n, dot(n, A)));
}
int orc_sphere(void* s, double r) { return ((Scene*)s)->push(std::make_shared<Sphere>(r)); }  // :154
int orc_i_cylinder(void* s, double r) { return ((Scene*)s)->push(std::make_shared<Cylinder>(r)); }  // :160
int orc_i_cone(void* s, double slope, double offset) {  // :166, :200
    return ((Scene*)s)->push(std::make_shared<Cone>(slope, offset));
}
int orc_unit_sphere_sq(void* s) { return ((Scene*)s)->push(std::make_shared<UnitSphereSq>()); }
int orc_set_bbox(void* s, int node, const double* b) {  // Cone::set_bbox, :202-205
    ((Scene*)s)->nodes[node]->bbox_ = {{b[0], b[1], b[2]}, {b[3], b[4], b[5]}};
    return 0;
}
static std::vector<Obj> clones(Scene* sc, const int* ch, int n) {
    std::vector<Obj> v;
    for (int i = 0; i < n; ++i) v.push_back(sc->nodes[ch[i]]->clone());
    return v;
}
// Union::from_vec / Intersection::from_vec / difference_from_vec [A]; 1 child -> the child
int orc_union(void* s, const int* ch, int n, double r) {
    Scene* sc = (Scene*)s;
    if (n <= 0) return -1;
    if (n == 1) return sc->push(sc->nodes[ch[0]]->clone());
    return sc->push(std::make_shared<Boolean>(true, clones(sc, ch, n), r));
}
int orc_intersection(void* s, const int* ch, int n, double r) {
    Scene* sc = (Scene*)s;
    if (n <= 0) return -1;
    if (n == 1) return sc->push(sc->nodes[ch[0]]->clone());
    return sc->push(std::make_shared<Boolean>(false, clones(sc, ch, n), r));
}
int orc_difference(void* s, const int* ch, int n, double r) {
    Scene* sc = (Scene*)s;
    if (n <= 0) return -1;
    if (n == 1) return sc->push(sc->nodes[ch[0]]->clone());
    std::vector<Obj> v = clones(sc, ch, n);
    for (int i = 1; i < n; ++i) v[i] = std::make_shared<Negation>(v[i]);
    return sc->push(std::make_shared<Boolean>(false, std::move(v), r));
}
int orc_negation(void* s, int node) {
    Scene* sc = (Scene*)s;
    return sc->push(std::make_shared<Negation>(sc->nodes[node]->clone()));
}
int orc_translate(void* s, int node, double x, double y, double z) {  // :221-225
    Scene* sc = (Scene*)s;
    Obj o = sc->nodes[node]->translate({x, y, z});
    if (!std::static_pointer_cast<Affine>(o)->ok) return -2;
    return sc->push(o);
}
int orc_rotate(void* s, int node, double x, double y, double z) {  // :227-231
    Scene* sc = (Scene*)s;
    Obj o = sc->nodes[node]->rotate({x, y, z});
    if (!std::static_pointer_cast<Affine>(o)->ok) return -2;
    return sc->push(o);
}
int orc_scale(void* s, int node, double x, double y, double z) {  // :233-237
    Scene* sc = (Scene*)s;
    Obj o = sc->nodes[node]->scale({x, y, z});
    if (!std::static_pointer_cast<Affine>(o)->ok) return -2;
    return sc->push(o);
}
// src/main.rs:53-58 (applied to every output object)
void orc_set_parameters(void* s, int node, double fade_range, double r_multiplier) {
    ((Scene*)s)->nodes[node]->set_parameters(fade_range, r_multiplier);
}
void orc_bbox(void* s, int node, double* out) {
    const BBox& b = ((Scene*)s)->nodes[node]->bbox_;
    out[0] = b.min.x; out[1] = b.min.y; out[2] = b.min.z;
    out[3] = b.max.x; out[4] = b.max.y; out[5] = b.max.z;
}
double orc_approx_value(void* s, int node, double x, double y, double z, double slack) {
    return ((Scene*)s)->nodes[node]->approx_value({x, y, z}, slack);
}
void orc_eval_points(void* s, int node, double slack, const double* pts, uint64_t n, double* out) {
    const Object* o = ((Scene*)s)->nodes[node].get();
    for (uint64_t i = 0; i < n; ++i)
        out[i] = o->approx_value({pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, slack);
}
// the lattice SDF field in the library's layout (x fastest, z slowest), positions as
// tessellate3d/src/lib.rs:84-103
void orc_eval_lattice(void* s, int node, double slack, double res, const uint64_t* dim, double* out) {
    const Object* o = ((Scene*)s)->nodes[node].get();
    V3 origin = o->bbox_.min;
    uint64_t i = 0;
    for (uint64_t z = 0; z <= dim[2]; ++z)
        for (uint64_t y = 0; y <= dim[1]; ++y)
            for (uint64_t x = 0; x <= dim[0]; ++x)
                out[i++] = o->approx_value({(double)x * res + origin.x, (double)y * res + origin.y,
                                            (double)z * res + origin.z}, slack);
}

// src/main.rs:74-82: IsosurfaceStuffing::new(&ObjectAdaptor{implicit, res}, res, err).tessellate()
void* orc_tessellate(void* s, int node, double resolution, double relative_error) {
    const Object* o = ((Scene*)s)->nodes[node].get();
    Stuffing st(o, resolution, resolution, relative_error);
    auto* r = new MeshResult();
    if (!st.ok) { r->ok = false; return r; }
    st.tessellate();
    r->ok = st.ok;
    r->mesh = std::move(st.mesh);
    r->st = st.st;
    return r;
}
void orc_mesh_free(void* m) { delete (MeshResult*)m; }
int orc_mesh_ok(void* m) { return ((MeshResult*)m)->ok ? 1 : 0; }
uint64_t orc_mesh_num_vertices(void* m) { return ((MeshResult*)m)->mesh.coords.size(); }
uint64_t orc_mesh_num_tets(void* m) { return ((MeshResult*)m)->mesh.elems.size(); }
void orc_mesh_copy_coords(void* m, double* out) {
    auto& c = ((MeshResult*)m)->mesh.coords;
    for (size_t i = 0; i < c.size(); ++i) { out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].z; }
}
void orc_mesh_copy_tets(void* m, uint64_t* out) {
    auto& e = ((MeshResult*)m)->mesh.elems;
    for (size_t i = 0; i < e.size(); ++i) for (int k = 0; k < 4; ++k) out[4 * i + k] = e[i][k];
}
// out[0..3]=dim, [3..6]=ntets_stage, [6..13]=stats, [13]=n_warped, [14]=n_cut, [15]=n_value_calls,
// [16]=n_bisect_iters, [17]=n_exterior_removed, [18]=n_exterior_skipped_by_swap, [19]=inverted
void orc_mesh_stats(void* m, uint64_t* out, double* ar_minmax, double* times) {
    const Stats& st = ((MeshResult*)m)->st;
    for (int i = 0; i < 3; ++i) out[i] = st.dim[i];
    for (int i = 0; i < 3; ++i) out[3 + i] = st.ntets_stage[i];
    for (int i = 0; i < 7; ++i) out[6 + i] = st.stats[i];
    out[13] = st.n_warped; out[14] = st.n_cut; out[15] = st.n_value_calls;
    out[16] = st.n_bisect_iters; out[17] = st.n_exterior_removed;
    out[18] = st.n_exterior_skipped_by_swap; out[19] = st.inverted;
    ar_minmax[0] = st.ar_min; ar_minmax[1] = st.ar_max;
    times[0] = st.t_cut; times[1] = st.t_warp; times[2] = st.t_trim; times[3] = st.t_ext;
    times[4] = st.t_compact; times[5] = st.t_check; times[6] = st.t_boundary; times[7] = st.t_sidesets;
}
// src/main.rs:86
uint64_t orc_mesh_boundary(void* m) {
    auto* r = (MeshResult*)m;
    if (!r->have_surface) {
        double t0 = Stuffing::now();
        r->surface = boundary(r->mesh);
        r->st.t_boundary = Stuffing::now() - t0;
        r->have_surface = true;
    }
    return r->surface.size();
}
void orc_mesh_copy_boundary(void* m, uint64_t* out) {
    auto* r = (MeshResult*)m;
    for (size_t i = 0; i < r->surface.size(); ++i) { out[2 * i] = r->surface[i].first; out[2 * i + 1] = r->surface[i].second; }
}
// src/main.rs:87-106: a surface side joins the set iff all three nodes have
// |approx_value(p, EPSILON)| <= EPSILON.  Returns the side-set index.
int orc_mesh_add_sideset(void* m, void* s, int node) {
    auto* r = (MeshResult*)m;
    orc_mesh_boundary(m);
    const Object* o = ((Scene*)s)->nodes[node].get();
    const double EPS = std::numeric_limits<double>::epsilon();
    double t0 = Stuffing::now();
    std::vector<std::pair<size_t, size_t>> set;
    for (auto& es : r->surface) {
        bool add = true;
        for (int k = 0; k < 3; ++k) {
            size_t n = r->mesh.elems[es.first][TET_FACE[es.second][k]];
            double phi = o->approx_value(r->mesh.coords[n], EPS);
            if (phi > EPS || phi < -EPS) { add = false; break; }
        }
        if (add) set.push_back(es);
    }
    r->st.t_sidesets += Stuffing::now() - t0;
    r->side_sets.push_back(std::move(set));
    return (int)r->side_sets.size() - 1;
}
uint64_t orc_mesh_sideset_len(void* m, int i) { return ((MeshResult*)m)->side_sets[i].size(); }
void orc_mesh_copy_sideset(void* m, int i, uint64_t* out) {
    auto& v = ((MeshResult*)m)->side_sets[i];
    for (size_t k = 0; k < v.size(); ++k) { out[2 * k] = v[k].first; out[2 * k + 1] = v[k].second; }
}

// ---- mesh-crate level entry points used by the known-answer tests -------------------------
void* orc_rawmesh_new(const double* coords, uint64_t nv, const uint64_t* tets, uint64_t nt) {
    auto* r = new MeshResult();
    for (uint64_t i = 0; i < nv; ++i) r->mesh.add_vertex({coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]});
    for (uint64_t i = 0; i < nt; ++i) r->mesh.add_elem({(size_t)tets[4 * i], (size_t)tets[4 * i + 1], (size_t)tets[4 * i + 2], (size_t)tets[4 * i + 3]});
    return r;
}
void orc_rawmesh_compact(void* m) { ((MeshResult*)m)->mesh.compact(); }
double orc_rawmesh_volume(void* m, uint64_t e) { auto* r = (MeshResult*)m; return r->mesh.volume(r->mesh.elems[e]); }
double orc_rawmesh_aspect_ratio(void* m, uint64_t e) { auto* r = (MeshResult*)m; return r->mesh.aspect_ratio(r->mesh.elems[e]); }
// mesh.normal(&tet.side(s)) = (b-a) x (c-a), mesh/src/lib.rs:285-290
void orc_rawmesh_face_normal(void* m, uint64_t e, int s, double* out) {
    auto* r = (MeshResult*)m;
    const Tet& t = r->mesh.elems[e];
    V3 a = r->mesh.coords[t[TET_FACE[s][0]]], b = r->mesh.coords[t[TET_FACE[s][1]]], c = r->mesh.coords[t[TET_FACE[s][2]]];
    V3 n = cross(b - a, c - a);
    out[0] = n.x; out[1] = n.y; out[2] = n.z;
}
// tessellate3d/src/lib.rs:492-507 test_cut_edge: two vertices, phi from the function,
// cut_edge(0,1).  Returns the new vertex index, writes its position.
int64_t orc_cut_edge_test(void* s, int node, double resolution, double relative_error,
                          const double* v0, const double* v1, double* out_pos, double* out_phi_new) {
    const Object* o = ((Scene*)s)->nodes[node].get();
    Stuffing st(o, resolution, resolution, relative_error);
    st.mesh.add_vertex({v0[0], v0[1], v0[2]});
    st.mesh.add_vertex({v1[0], v1[1], v1[2]});
    st.phi = {st.value(st.mesh.coords[0]), st.value(st.mesh.coords[1])};
    size_t idx = st.cut_edge(0, 1);
    out_pos[0] = st.mesh.coords[idx].x; out_pos[1] = st.mesh.coords[idx].y; out_pos[2] = st.mesh.coords[idx].z;
    *out_phi_new = st.phi[idx];
    return (int64_t)idx;
}
const int* orc_tile_lattice() { return &TILE_LATTICE[0][0]; }
const int* orc_tile_tets() { return &TILE_TETS[0][0]; }

}  // extern "C"
