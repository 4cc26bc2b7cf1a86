// lyon_oracle.hpp -- CPU restatement of lyon's curve flattening and StrokeTessellator.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
// lyon_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may build, link or call it.  The product (lyon_b200) never does.
//
// It restates, scalar and single-threaded, the algorithms of nical/lyon @ 8071ec06:
//   crates/geom/src/quadratic_bezier.rs, cubic_bezier.rs, line.rs   (flattening, line math)
//   crates/tessellation/src/stroke.rs, math_utils.rs                (StrokeTessellator)
// plus the handful of euclid 0.22.9 vector/angle routines those call (euclid's source is
// not in /root/reference; its formulas are restated from the published crate, see
// DESIGN.md "Oracle").  Every function cites the reference file:line it follows.
//
// PARITY PINNING: the Rust reference cannot be built in this image (no cargo/rustc).  The
// oracle is pinned against every known-answer the reference's own tests hold for this path
// (tests/test_oracle_known_answers.py: test_square 12/16 triangles, test_empty_caps 0/8,
// test_line_width corners, correct_miter_clip_length exact max-y, compute_normal vectors,
// flatten monotonic-t / last==to / straight-line==1 segment ...).  There are no golden
// vertex/index arrays in the reference, so beyond those answers parity with real lyon is
// "partial": see DESIGN.md.
//
// Build with:  g++ -O2 -ffp-contract=off -fno-fast-math   (rustc never contracts a*b+c).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace lyon_oracle {

// ---------------------------------------------------------------------------------------
// euclid 0.22.9 restatement (Vector2D / Point2D / Angle / Trig::fast_atan2)
// ---------------------------------------------------------------------------------------
struct V2 {
    float x, y;
};
static inline V2 v2(float x, float y) { return V2{x, y}; }
static inline V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
static inline V2 operator-(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
static inline V2 operator-(V2 a) { return V2{-a.x, -a.y}; }
static inline V2 operator*(V2 a, float s) { return V2{a.x * s, a.y * s}; }
static inline V2 operator/(V2 a, float s) { return V2{a.x / s, a.y / s}; }
// euclid Vector2D::dot / cross / square_length / length / normalize
static inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline float sqlen(V2 a) { return a.x * a.x + a.y * a.y; }
static inline float length(V2 a) { return sqrtf(sqlen(a)); }
static inline V2 normalize(V2 a) { return a / length(a); }
// euclid Point2D::lerp : (1-t)*a + t*b per component
static inline V2 lerp(V2 a, V2 b, float t) {
    float one_t = 1.0f - t;
    return V2{one_t * a.x + t * b.x, one_t * a.y + t * b.y};
}
static inline bool is_nan(float v) { return v != v; }

static const float PI_F = 3.14159274101257324f;       // core::f32::consts::PI
static const float FRAC_PI_2_F = 1.57079637050628662f;  // core::f32::consts::FRAC_PI_2

// euclid Trig::fast_atan2 for f32 (polynomial approximation, NOT libm atan2).
static inline float fast_atan2(float y, float x) {
    float x_abs = fabsf(x);
    float y_abs = fabsf(y);
    float a = fminf(x_abs, y_abs) / fmaxf(x_abs, y_abs);
    float s = a * a;
    float result = ((-0.0464964749f * s + 0.15931422f) * s - 0.327622764f) * s * a + a;
    if (y_abs > x_abs) result = FRAC_PI_2_F - result;
    if (x < 0.0f) result = PI_F - result;
    if (y < 0.0f) result = -result;
    return result;
}
// euclid Vector2D::angle_from_x_axis
static inline float angle_from_x_axis(V2 v) { return fast_atan2(v.y, v.x); }
// euclid Angle::angle_to : d = (to - self) % 2pi ; 2*d % 2pi - d      (% is fmodf)
static inline float angle_to(float self, float to) {
    float max = PI_F * 2.0f;
    float d = fmodf(to - self, max);
    return fmodf(2.0f * d, max) - d;
}

// libm calls (Rust f32::{sin,cos,acos,asin,log2} -> platform libm).  Evaluated in double and
// rounded once so that host and device agree (see DESIGN.md "Transcendentals").
// Every call can be recorded (function code + argument) for tests/test_libm_pinning.py, which compares this
// convention with the platform's own f32 routines (what rustc-built lyon would call) on the arguments that
// actually occur: 0 cos, 1 sin (arc vertices, variable-width side normals), 2 acos, 3 asin, 4 log2, 5 cos / 6 sin of
// find_sharp_turn's rotation.
struct LibmTrace {
    std::vector<uint8_t> fn;
    std::vector<float> arg;
};
static thread_local LibmTrace* g_libm_trace = nullptr;
static inline void libm_note(uint8_t fn, float a) {
    if (g_libm_trace) {
        g_libm_trace->fn.push_back(fn);
        g_libm_trace->arg.push_back(a);
    }
}
// g_libm_platform_f32 (tests only, oracle_set_libm_mode): call the platform's f32 routines instead -- what a rustc build
// of lyon does on this host -- so that the whole oracle can be run under both conventions and the meshes compared.
static thread_local bool g_libm_platform_f32 = false;
static inline float libm_cos(float a, uint8_t code = 0) { libm_note(code, a); return g_libm_platform_f32 ? cosf(a) : (float)cos((double)a); }
static inline float libm_sin(float a, uint8_t code = 1) { libm_note(code, a); return g_libm_platform_f32 ? sinf(a) : (float)sin((double)a); }
static inline float libm_acos(float a) { libm_note(2, a); return g_libm_platform_f32 ? acosf(a) : (float)acos((double)a); }
static inline float libm_asin(float a) { libm_note(3, a); return g_libm_platform_f32 ? asinf(a) : (float)asin((double)a); }
static inline float libm_log2(float a) { libm_note(4, a); return g_libm_platform_f32 ? log2f(a) : (float)log2((double)a); }

// Rust `x as u32` (saturating; NaN -> 0).
static inline uint32_t sat_u32(float v) {
    if (!(v > 0.0f)) return 0u;  // NaN, negative, zero
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
// num-traits ToPrimitive::to_u32 for f32, then `.unwrap_or(1)` as used by the flatteners.
static inline uint32_t to_u32_or_1(float n) {
    if (n > -1.0f && n < 4294967296.0f) return (uint32_t)n;  // truncation toward zero
    return 1u;                                               // NaN / inf / out of range -> None
}

// ---------------------------------------------------------------------------------------
// lyon_geom flattening
// ---------------------------------------------------------------------------------------
// crates/geom/src/quadratic_bezier.rs:956-973 and cubic_bezier.rs:1385-1402 (shared tail).
static inline float wang_segments_from_err4(float err4) {
    static const float LUT[24] = {1.0f,      16.0f,     81.0f,     256.0f,    625.0f,    1296.0f,
                                  2401.0f,   4096.0f,   6561.0f,   10000.0f,  14641.0f,  20736.0f,
                                  28561.0f,  38416.0f,  50625.0f,  65536.0f,  83521.0f,  104976.0f,
                                  130321.0f, 160000.0f, 194481.0f, 234256.0f, 279841.0f, 331776.0f};
    if (err4 <= 331776.0f) {
        for (int i = 0; i < 24; ++i)
            if (err4 <= LUT[i]) return (float)(i + 1);
    }
    return fmaxf(sqrtf(sqrtf(err4)), 1.0f);
}
// crates/geom/src/quadratic_bezier.rs:943-974
static inline float quad_wang(V2 from, V2 ctrl, V2 to, float tol) {
    V2 l = (from - ctrl * 2.0f + to) * 2.0f;
    float d = 1.0f / (8.0f * tol);
    float err4 = dot(l, l) * d * d;
    return wang_segments_from_err4(err4);
}
// crates/geom/src/cubic_bezier.rs:1369-1403
static inline float cubic_wang(V2 from, V2 c1, V2 c2, V2 to, float tol) {
    V2 v1 = (from - c1 * 2.0f + c2) * 6.0f;
    V2 v2 = (c1 - c2 * 2.0f + to) * 6.0f;
    float l = fmaxf(dot(v1, v1), dot(v2, v2));
    float d = 1.0f / (8.0f * tol);
    float err4 = l * d * d;
    return wang_segments_from_err4(err4);
}

struct QuadPoly {
    V2 a0, a1, a2;
};
struct CubicPoly {
    V2 a0, a1, a2, a3;
};
// crates/geom/src/quadratic_bezier.rs:891-900
static inline QuadPoly quad_poly(V2 from, V2 ctrl, V2 to) {
    QuadPoly p;
    p.a0 = from;
    p.a1 = (ctrl - from) * 2.0f;
    p.a2 = from + to - ctrl * 2.0f;
    return p;
}
// crates/geom/src/quadratic_bezier.rs:927-939 (Horner; v += a*t is mul then add, unfused)
static inline V2 quad_poly_sample(const QuadPoly& p, float t) {
    V2 v = p.a0;
    float t2 = t;
    v = v + p.a1 * t2;
    t2 = t2 * t;
    v = v + p.a2 * t2;
    return v;
}
// crates/geom/src/cubic_bezier.rs:1318-1325
static inline CubicPoly cubic_poly(V2 from, V2 c1, V2 c2, V2 to) {
    CubicPoly p;
    p.a0 = from;
    p.a1 = (c1 - from) * 3.0f;
    p.a2 = from * 3.0f - c1 * 6.0f + c2 * 3.0f;
    p.a3 = to - from + (c1 - c2) * 3.0f;
    return p;
}
// crates/geom/src/cubic_bezier.rs:1351-1365
static inline V2 cubic_poly_sample(const CubicPoly& p, float t) {
    V2 v = p.a0;
    float t2 = t;
    v = v + p.a1 * t2;
    t2 = t2 * t;
    v = v + p.a2 * t2;
    t2 = t2 * t;
    v = v + p.a3 * t2;
    return v;
}

// crates/geom/src/quadratic_bezier.rs:368-391.  cb(to, t_end)
template <class F>
static inline void quad_for_each_flattened_with_t(V2 from, V2 ctrl, V2 to, float tol, F&& cb) {
    float n = quad_wang(from, ctrl, to, tol);
    float step = 1.0f / n;
    float t0 = 0.0f;
    QuadPoly poly = quad_poly(from, ctrl, to);
    uint32_t cnt = to_u32_or_1(n) - 1u;
    for (uint32_t i = 0; i < cnt; ++i) {
        float t1 = t0 + step;
        cb(quad_poly_sample(poly, t1), t1);
        t0 = t1;
    }
    cb(to, 1.0f);
}
// crates/geom/src/cubic_bezier.rs:590-614
template <class F>
static inline void cubic_for_each_flattened_with_t(V2 from, V2 c1, V2 c2, V2 to, float tol, F&& cb) {
    float n = cubic_wang(from, c1, c2, to, tol);
    float step = 1.0f / n;
    float t0 = 0.0f;
    CubicPoly poly = cubic_poly(from, c1, c2, to);
    uint32_t cnt = to_u32_or_1(n) - 1u;
    for (uint32_t i = 0; i < cnt; ++i) {
        float t1 = t0 + step;
        cb(cubic_poly_sample(poly, t1), t1);
        t0 = t1;
    }
    cb(to, 1.0f);
}

// crates/geom/src/quadratic_bezier.rs:35-41 (Bernstein form, left-to-right association)
static inline V2 quad_sample(V2 from, V2 ctrl, V2 to, float t) {
    float t2 = t * t;
    float one_t = 1.0f - t;
    float one_t2 = one_t * one_t;
    return from * one_t2 + ctrl * 2.0f * one_t * t + to * t2;
}
// crates/geom/src/quadratic_bezier.rs:180-191 ; returns false for None
static inline bool quad_local_x_extremum_t(V2 from, V2 ctrl, V2 to, float* t_out) {
    float div = from.x - 2.0f * ctrl.x + to.x;
    if (div == 0.0f) return false;
    float t = (from.x - ctrl.x) / div;
    if (t > 0.0f && t < 1.0f) {
        *t_out = t;
        return true;
    }
    return false;
}
// crates/tessellation/src/stroke.rs:3178-3212 ; returns false for None
static inline bool find_sharp_turn(V2 from, V2 ctrl, V2 to, float* t_out) {
    V2 baseline = to - from;
    V2 v = ctrl - from;
    V2 n = v2(-baseline.y, baseline.x);
    float v_dot_b = dot(v, baseline);
    float v_dot_n = dot(v, n);
    V2 long_axis;
    if ((v_dot_b >= 0.0f && v_dot_b <= dot(baseline, baseline)) ||
        fabsf(v_dot_n) * 2.0f >= fabsf(v_dot_b)) {
        if (sqlen(baseline) * 30.0f > sqlen(v)) return false;
        long_axis = v;
    } else {
        long_axis = baseline;
    }
    // euclid Rotation2D::new(-angle).transform_vector: (x*cos - y*sin, y*cos + x*sin)
    float ang = -angle_from_x_axis(long_axis);
    float s = libm_sin(ang, 6), c = libm_cos(ang, 5);  // traced apart: this rotation decides t_split, hence flatten counts
    V2 rctrl = v2(v.x * c - v.y * s, v.y * c + v.x * s);
    V2 rto = v2(baseline.x * c - baseline.y * s, baseline.y * c + baseline.x * s);
    return quad_local_x_extremum_t(v2(0.0f, 0.0f), rctrl, rto, t_out);
}
// crates/tessellation/src/stroke.rs:2755-2779 ; cb(position, t, is_flattening_step)
template <class F>
static inline void flatten_quad(V2 from, V2 ctrl, V2 to, float tol, F&& cb) {
    float t_split;
    if (find_sharp_turn(from, ctrl, to, &t_split)) {
        // QuadraticBezierSegment::split, quadratic_bezier.rs:208-223
        V2 split_point = quad_sample(from, ctrl, to, t_split);
        V2 b_ctrl = lerp(from, ctrl, t_split);
        V2 a_ctrl = lerp(ctrl, to, t_split);
        quad_for_each_flattened_with_t(from, b_ctrl, split_point, tol, [&](V2 p, float t) {
            bool fs = t != 1.0f;
            cb(p, t * t_split, fs);
        });
        quad_for_each_flattened_with_t(split_point, a_ctrl, to, tol, [&](V2 p, float t) {
            bool fs = t != 1.0f;
            cb(p, t_split + t * (1.0f - t_split), fs);
        });
    } else {
        quad_for_each_flattened_with_t(from, ctrl, to, tol, [&](V2 p, float t) { cb(p, t, t != 1.0f); });
    }
}

// crates/geom/src/line.rs:562-578 in f64 (Scalar::EPSILON for f64 = 1e-8, lib.rs:224)
struct LineD {
    double px, py, vx, vy;
};
static inline LineD line_d(V2 p, V2 v) { return LineD{(double)p.x, (double)p.y, (double)v.x, (double)v.y}; }
static inline bool line_intersection(const LineD& a, const LineD& b, V2* out) {
    double det = a.vx * b.vy - a.vy * b.vx;
    if (fabs(det) <= 1e-8) return false;
    double inv_det = 1.0 / det;
    double a2x = a.px + a.vx, a2y = a.py + a.vy;
    double b2x = b.px + b.vx, b2y = b.py + b.vy;
    double ca = a.px * a2y - a.py * a2x;
    double cb = b.px * b2y - b.py * b2x;
    double x = (cb * a.vx - ca * b.vx) * inv_det;
    double y = (cb * a.vy - ca * b.vy) * inv_det;
    *out = v2((float)x, (float)y);
    return true;
}
// crates/geom/src/utils.rs:15-17
static inline V2 tangent(V2 v) { return v2(-v.y, v.x); }

// crates/tessellation/src/math_utils.rs:16-40
static inline V2 compute_normal(V2 v1, V2 v2_) {
    const float epsilon = 1e-4f;
    V2 n1 = v2(-v1.y, v1.x);
    V2 v12 = v1 + v2_;
    if (sqlen(v12) < epsilon) return v2(0.0f, 0.0f);
    V2 tg = normalize(v12);
    V2 n = v2(-tg.y, tg.x);
    float inv_len = dot(n, n1);
    if (fabsf(inv_len) < epsilon) return n1;
    return n / inv_len;
}

// ---------------------------------------------------------------------------------------
// Options / vertex output (same layout as include/lyon_cuda.h)
// ---------------------------------------------------------------------------------------
enum : uint8_t { CAP_BUTT = 0, CAP_SQUARE = 1, CAP_ROUND = 2 };              // crates/path/src/lib.rs:170-174
enum : uint8_t { JOIN_MITER = 0, JOIN_MITER_CLIP = 1, JOIN_ROUND = 2, JOIN_BEVEL = 3 };  // lib.rs:193-198

struct Options {
    uint8_t start_cap = CAP_BUTT, end_cap = CAP_BUTT, line_join = JOIN_MITER, _pad = 0;
    float line_width = 1.0f;
    int32_t variable_line_width = -1;  // -1 = None
    float miter_limit = 4.0f;
    float tolerance = 0.1f;
};

static const uint32_t SIDE_NEG_BIT = 0x80000000u;
// StrokeVertexData mirror (stroke.rs:2656-2665), 32 bytes.
struct Vertex {
    float px, py;   // position_on_path
    float nx, ny;   // normal
    float advancement;
    float half_width;
    uint32_t src;   // bits 0..30: endpoint id of VertexSource (`to` for Edge); bit 31: Side::Negative
    float t;        // 1.0 => VertexSource::Endpoint, else VertexSource::Edge{t}
};

enum : int { OK = 0, ERR_TOO_MANY_VERTICES = -2 };

// BuffersBuilder<_, u32> mirror (geometry_builder.rs:300-351, 451-508, MaxIndex :562-603)
struct Buffers {
    std::vector<Vertex> vertices;
    std::vector<uint32_t> vertex_from;  // per vertex: VertexSource::Edge.from (what interpolated_attributes reads, stroke.rs:2735)
    std::vector<uint32_t> indices;
    uint64_t max_index = 0xFFFFFFFEull;  // <u32 as MaxIndex>::MAX: the largest id that may be handed out
    size_t first_vertex = 0, first_index = 0;
    void begin_geometry() {
        first_vertex = vertices.size();
        first_index = indices.size();
    }
    void abort_geometry() {
        vertices.resize(first_vertex);
        vertex_from.resize(first_vertex);
        indices.resize(first_index);
    }
    bool add_vertex(const Vertex& v, uint32_t* id) {
        if ((uint64_t)vertices.size() > max_index) return false;  // geometry_builder.rs:342-346
        *id = (uint32_t)vertices.size();
        vertices.push_back(v);
        return true;
    }
    void add_triangle(uint32_t a, uint32_t b, uint32_t c) {
        indices.push_back(a);
        indices.push_back(b);
        indices.push_back(c);
    }
};

// ---------------------------------------------------------------------------------------
// StrokeTessellator restatement
// ---------------------------------------------------------------------------------------
static const int SIDE_POSITIVE = 0;
static const int SIDE_NEGATIVE = 1;

struct SidePoints {  // stroke.rs:311-317
    V2 prev, next;
    bool has_single;
    V2 single;
    uint32_t prev_vertex, next_vertex;
};
struct Source {
    uint32_t id;    // Endpoint id, or Edge.to
    float t;        // 1.0 for Endpoint
    uint32_t from;  // Edge.from (lib.rs:282-291); unused for Endpoint sources
};
struct EndpointData {  // stroke.rs:320-352
    V2 position;
    float half_width;
    float advancement;
    uint8_t line_join;
    Source src;
    SidePoints side_points[2];
    bool fold[2];
    bool is_flattening_step;
};
static inline EndpointData endpoint_default() {
    const float nan = std::numeric_limits<float>::quiet_NaN();
    EndpointData e;
    e.position = v2(0.0f, 0.0f);
    e.half_width = nan;
    e.advancement = nan;
    e.line_join = JOIN_MITER;
    e.src = Source{0xFFFFFFFFu, 1.0f, 0u};
    for (int s = 0; s < 2; ++s) {
        e.side_points[s].prev = v2(nan, nan);
        e.side_points[s].next = v2(nan, nan);
        e.side_points[s].has_single = false;
        e.side_points[s].single = v2(nan, nan);
        e.side_points[s].prev_vertex = 0xFFFFFFFFu;
        e.side_points[s].next_vertex = 0xFFFFFFFFu;
    }
    e.fold[0] = e.fold[1] = false;
    e.is_flattening_step = false;
    return e;
}

struct VertexState {  // StrokeVertexData, stroke.rs:2656-2665
    V2 position_on_path;
    float half_width;
    V2 normal;
    float advancement;
    bool side_negative;
    Source src;
};

struct PointBuffer {  // stroke.rs:2371-2457
    EndpointData points[3];
    int start = 0, count = 0;
    void push(const EndpointData& p) {
        if (count < 3) {
            points[count] = p;
            count += 1;
            return;
        }
        points[start] = p;
        start += 1;
        if (start == 3) start = 0;
    }
    void replace_last(const EndpointData& p) {
        int idx = start;
        if (idx == 0) idx = count;
        points[idx - 1] = p;
    }
    void clear() {
        count = 0;
        start = 0;
    }
    EndpointData& get(int idx) { return points[(idx + start) % 3]; }
    EndpointData& last() { return get(count - 1); }
    EndpointData& last_but_one() { return points[(start + count - 2) % 3]; }
};

// stroke.rs:2320-2322
static inline bool miter_limit_is_exceeded(V2 normal, float miter_limit) {
    return sqlen(normal) > miter_limit * miter_limit * 4.0f;
}
static inline float side_sign(int side) { return side == SIDE_NEGATIVE ? -1.0f : 1.0f; }  // :2324-2330
// stroke.rs:2749-2753
static inline float circle_flattening_step(float radius, float tolerance) {
    tolerance = fminf(tolerance, radius);
    return 2.0f * libm_acos((radius - tolerance) / radius);
}
// stroke.rs:2279-2315
static inline void get_clip_intersections(V2 previous_normal, V2 next_normal, V2 normal, float clip_distance,
                                          V2* i1, V2* i2) {
    LineD clip_line = line_d(normalize(normal) * clip_distance, tangent(normal));
    LineD prev_line = line_d(previous_normal, tangent(previous_normal));
    LineD next_line = line_d(next_normal, tangent(next_normal));
    if (!line_intersection(clip_line, prev_line, i1)) *i1 = normal;
    if (!line_intersection(clip_line, next_line, i2)) *i2 = normal;
}

template <class Out>
class Stroker {
public:
    Stroker(const Options& o, Out& out) : options(o), output(out) {  // stroke.rs:591-631
        output.begin_geometry();
        square_merge_threshold =
            fmaxf(fminf(o.tolerance * o.tolerance * 0.5f, o.line_width * o.line_width * 0.05f), 1e-8f);
        vertex.position_on_path = v2(0.0f, 0.0f);
        vertex.normal = v2(0.0f, 0.0f);
        vertex.half_width = o.line_width * 0.5f;
        vertex.advancement = 0.0f;
        vertex.side_negative = true;
        vertex.src = Source{0xFFFFFFFFu, 1.0f, 0u};
    }

    // One path = one `StrokeTessellator::tessellate_path` call (stroke.rs:164-177):
    // num_attributes > 0 -> id-based events (endpoint ids = offsets in the point array),
    // else PathEvent iteration with sequential ids (stroke.rs:872-952).
    // verbs: {LineTo=0,QuadraticTo=1,CubicTo=2,Begin=3,Close=4,End=5} (path.rs:29-36)
    int tessellate_path(const uint8_t* verbs, uint32_t n_verbs, const float* pts, uint32_t num_attributes) {
        const uint32_t a_stride = (num_attributes + 1) / 2;  // path.rs:838 div_ceil(2)
        const bool with_ids = num_attributes > 0;
        const bool vw = with_ids && options.variable_line_width >= 0;
        if (!with_ids) options.variable_line_width = -1;  // stroke.rs:878
        const float base_width = options.line_width;
        const int attr_idx = options.variable_line_width;
        uint32_t cur = 0;        // cursor in point array (in points)
        uint32_t seq_id = 0;     // tessellate_fw id counter
        uint32_t cur_ep = 0xFFFFFFFFu;  // current endpoint offset (id iter)
        const float nan = std::numeric_limits<float>::quiet_NaN();
        V2 current_position = v2(nan, nan);
        auto P = [&](uint32_t i) { return v2(pts[2 * i], pts[2 * i + 1]); };
        auto attr = [&](uint32_t ep) { return pts[2 * (ep + 1) + attr_idx]; };  // path.rs:1266-1282
        for (uint32_t vi = 0; vi < n_verbs; ++vi) {
            uint8_t verb = verbs[vi];
            switch (verb) {
                case 3: {  // Begin
                    V2 at = P(cur);
                    uint32_t ep = cur;
                    cur += 1 + a_stride;
                    current_position = at;
                    cur_ep = ep;
                    may_need_empty_cap = false;
                    EndpointData d = endpoint_default();
                    d.position = at;
                    d.advancement = sub_path_start_advancement;
                    d.line_join = options.line_join;
                    d.src = Source{with_ids ? ep : seq_id, 1.0f, 0u};
                    if (vw) {
                        d.half_width = base_width * attr(ep) * 0.5f;  // stroke.rs:671
                        step(d);
                    } else {
                        d.half_width = options.line_width * 0.5f;
                        fixed_width_step(d);
                    }
                    seq_id += 1;
                    break;
                }
                case 0: {  // LineTo
                    V2 to = P(cur);
                    uint32_t ep = cur;
                    cur += 1 + a_stride;
                    current_position = to;
                    cur_ep = ep;
                    EndpointData d = endpoint_default();
                    d.position = to;
                    d.line_join = options.line_join;
                    d.src = Source{with_ids ? ep : seq_id, 1.0f, 0u};
                    if (vw) {
                        d.half_width = base_width * attr(ep) * 0.5f;  // stroke.rs:689
                        step(d);
                    } else {
                        d.half_width = options.line_width * 0.5f;
                        fixed_width_step(d);
                    }
                    seq_id += 1;
                    break;
                }
                case 1: {  // QuadraticTo
                    V2 ctrl = P(cur);
                    V2 to = P(cur + 1);
                    uint32_t ep = cur + 1;
                    cur += 2 + a_stride;
                    V2 from = current_position;
                    uint32_t from_ep = cur_ep;
                    current_position = to;
                    cur_ep = ep;
                    uint32_t to_id = with_ids ? ep : seq_id;
                    uint32_t from_id = with_ids ? from_ep : seq_id - 1u;  // the previous event's id (stroke.rs:905-923)
                    float start_width = 0.0f, end_width = 0.0f;
                    if (vw) {
                        start_width = base_width * attr(from_ep);  // stroke.rs:705-706
                        end_width = base_width * attr(ep);
                    }
                    const float hw_fixed = options.line_width * 0.5f;
                    flatten_quad(from, ctrl, to, options.tolerance, [&](V2 p, float t, bool fs) {
                        EndpointData d = endpoint_default();
                        d.position = p;
                        d.line_join = options.line_join;
                        d.src = Source{to_id, t, from_id};  // t == 1.0 -> Endpoint (stroke.rs:1009-1017)
                        d.is_flattening_step = fs;
                        if (vw) {
                            d.half_width = (start_width * (1.0f - t) + end_width * t) * 0.5f;  // :1022
                            step(d);
                        } else {
                            d.half_width = hw_fixed;
                            fixed_width_step(d);
                        }
                    });
                    seq_id += 1;
                    break;
                }
                case 2: {  // CubicTo
                    V2 c1 = P(cur);
                    V2 c2 = P(cur + 1);
                    V2 to = P(cur + 2);
                    uint32_t ep = cur + 2;
                    cur += 3 + a_stride;
                    V2 from = current_position;
                    uint32_t from_ep = cur_ep;
                    current_position = to;
                    cur_ep = ep;
                    uint32_t to_id = with_ids ? ep : seq_id;
                    uint32_t from_id = with_ids ? from_ep : seq_id - 1u;  // the previous event's id (stroke.rs:905-923)
                    float start_width = 0.0f, end_width = 0.0f;
                    if (vw) {
                        start_width = base_width * attr(from_ep);  // stroke.rs:731-732
                        end_width = base_width * attr(ep);
                    }
                    const float hw_fixed = options.line_width * 0.5f;
                    cubic_for_each_flattened_with_t(from, c1, c2, to, options.tolerance, [&](V2 p, float t) {
                        bool fs = t != 1.0f;  // stroke.rs:1044 / :1154
                        EndpointData d = endpoint_default();
                        d.position = p;
                        d.line_join = options.line_join;
                        d.src = Source{to_id, t, from_id};
                        d.is_flattening_step = fs;
                        if (vw) {
                            d.half_width = (start_width * (1.0f - t) + end_width * t) * 0.5f;  // :1058
                            step(d);
                        } else {
                            d.half_width = hw_fixed;
                            fixed_width_step(d);
                        }
                    });
                    seq_id += 1;
                    break;
                }
                case 4:  // Close: the first point (and attributes) is stored again (path.rs:537-539)
                    cur += 1 + a_stride;
                    end(true);
                    break;
                case 5:  // End
                    end(false);
                    break;
                default:
                    break;
            }
            if (error != OK) {  // stroke.rs:944-947
                output.abort_geometry();
                return error;
            }
        }
        return OK;  // build(): end_geometry is a no-op for BuffersBuilder
    }

    Options options;

private:
    Out& output;
    int error = OK;
    VertexState vertex;
    PointBuffer point_buffer;
    EndpointData firsts[2];
    int n_firsts = 0;
    float sub_path_start_advancement = 0.0f;
    float square_merge_threshold;
    bool may_need_empty_cap = false;
    bool vw_mode() const { return options.variable_line_width >= 0; }

    void set_error(int e) {
        if (error == OK) error = e;
    }
    bool emit_vertex(uint32_t* id) {
        Vertex v;
        v.px = vertex.position_on_path.x;
        v.py = vertex.position_on_path.y;
        v.nx = vertex.normal.x;
        v.ny = vertex.normal.y;
        v.advancement = vertex.advancement;
        v.half_width = vertex.half_width;
        v.src = (vertex.src.id & 0x7FFFFFFFu) | (vertex.side_negative ? SIDE_NEG_BIT : 0u);
        v.t = vertex.src.t;
        if (!output.add_vertex(v, id)) return false;
        output.vertex_from.push_back(vertex.src.from);
        return true;
    }
    bool points_are_too_close(V2 p0, V2 p1) const {  // stroke.rs:1318-1320
        return sqlen(p0 - p1) < square_merge_threshold;
    }

    void fixed_width_step(const EndpointData& next) {  // stroke.rs:1635-1639
        bool added;
        if (!fixed_width_step_impl(next, &added)) set_error(ERR_TOO_MANY_VERTICES);
    }
    void step(const EndpointData& next) {  // stroke.rs:1515-1519
        bool added;
        if (!step_impl(next, &added)) set_error(ERR_TOO_MANY_VERTICES);
    }

    void end(bool close) {  // stroke.rs:1179-1193
        may_need_empty_cap |= close && point_buffer.count == 1;
        bool ok;
        if (close && point_buffer.count > 2)
            ok = close_subpath();
        else
            ok = end_with_caps();
        if (!ok) set_error(ERR_TOO_MANY_VERTICES);
        point_buffer.clear();
        n_firsts = 0;
    }

    bool close_subpath() {  // stroke.rs:1206-1279 (count > 2 guaranteed by the caller)
        EndpointData p = firsts[0];
        float advancement = p.advancement;
        p.advancement = std::numeric_limits<float>::quiet_NaN();
        bool segment_added;
        if (vw_mode()) {
            if (!step_impl(p, &segment_added)) return false;
        } else {
            if (!fixed_width_step_impl(p, &segment_added)) return false;
        }
        if (!segment_added) point_buffer.last().position = p.position;  // :1228-1238
        if (n_firsts >= 2) {
            EndpointData p2 = firsts[1];
            bool dummy;
            if (vw_mode()) {
                if (!step_impl(p2, &dummy)) return false;
            } else {
                if (!fixed_width_step_impl(p2, &dummy)) return false;
            }
            EndpointData& p0 = point_buffer.last_but_one();
            EndpointData& p1 = point_buffer.last();
            vertex.src = p0.src;
            vertex.position_on_path = p0.position;
            vertex.half_width = p0.half_width;
            vertex.advancement = advancement;
            for (int side = 0; side < 2; ++side) {
                vertex.side_negative = side != SIDE_POSITIVE;
                if (p0.side_points[side].has_single)
                    vertex.normal = (p0.side_points[side].single - p0.position) / p0.half_width;
                else
                    vertex.normal = (p0.side_points[side].next - p0.position) / p0.half_width;
                uint32_t id;
                if (!emit_vertex(&id)) return false;
                p0.side_points[side].next_vertex = id;
            }
            add_edge_triangles(p0, p1);
        }
        return true;
    }

    bool tessellate_empty_cap() {  // stroke.rs:1282-1316
        EndpointData& point = point_buffer.get(0);
        vertex.advancement = point.advancement;
        vertex.src = point.src;
        vertex.half_width = point.half_width;
        if (options.start_cap == CAP_SQUARE) return tessellate_empty_square_cap(point.position);
        if (options.start_cap == CAP_ROUND) return tessellate_empty_round_cap(point.position);
        return true;
    }

    bool end_with_caps() {  // stroke.rs:1322-1378
        int count = point_buffer.count;
        if (may_need_empty_cap && count == 1) return tessellate_empty_cap();
        if (count >= 2) {
            EndpointData p0 = point_buffer.last_but_one();
            EndpointData p1 = point_buffer.last();
            if (!vw_mode()) {
                V2 tg = normalize(p1.position - p0.position);
                V2 n = v2(-tg.y, tg.x) * p1.half_width;
                p1.side_points[SIDE_POSITIVE].prev = p1.position + n;
                p1.side_points[SIDE_NEGATIVE].prev = p1.position - n;
            }
            bool is_first = count == 2;
            if (!tessellate_last_edge(p0, p1, is_first)) return false;
            sub_path_start_advancement = p1.advancement;
            if (count > 2) {
                p0 = firsts[0];
                p1 = firsts[1];
            }
            if (!tessellate_first_edge(p0, p1)) return false;
        }
        return true;
    }

    bool step_impl(EndpointData next, bool* added) {  // stroke.rs:1380-1512
        int count = point_buffer.count;
        if (count > 0 && points_are_too_close(point_buffer.last().position, next.position)) {
            if (count == 1) may_need_empty_cap = true;
            *added = false;
            return true;
        }
        if (count > 0) {
            EndpointData& join = point_buffer.last();
            if (!join.is_flattening_step || !next.is_flattening_step) compute_edge_attachment_positions(join, next);
        }
        bool skip = false;
        if (count > 1) {
            EndpointData& prev = point_buffer.last_but_one();
            EndpointData& join = point_buffer.last();
            vertex.src = join.src;
            vertex.position_on_path = join.position;
            vertex.half_width = join.half_width;
            vertex.advancement = join.advancement;
            bool fast_path = false;
            if (join.is_flattening_step) {
                V2 v0 = join.position - prev.position;
                V2 v1 = next.position - join.position;
                fast_path = dot(v0, v1) > 0.0f;
            }
            if (fast_path) {
                join.line_join = JOIN_MITER;
                if (!flattened_step(prev, join, next, &skip)) return false;
            } else {
                compute_join_side_positions(prev, join, next, options.miter_limit, SIDE_POSITIVE);
                compute_join_side_positions(prev, join, next, options.miter_limit, SIDE_NEGATIVE);
                if (join.side_points[SIDE_POSITIVE].has_single) join.fold[SIDE_NEGATIVE] = false;
                if (join.side_points[SIDE_NEGATIVE].has_single) join.fold[SIDE_POSITIVE] = false;
                if (!add_join_base_vertices(join, true)) return false;
                if (!add_join_base_vertices(join, false)) return false;
            }
            if (!skip) {
                if (count > 2) add_edge_triangles(prev, join);
                if (!tessellate_join(join)) return false;
                if (count == 2) {
                    firsts[0] = prev;
                    firsts[1] = join;
                    n_firsts = 2;
                }
            }
        }
        if (skip)
            point_buffer.replace_last(next);
        else
            point_buffer.push(next);
        *added = true;
        return true;
    }

    bool fixed_width_step_impl(EndpointData next, bool* added) {  // stroke.rs:1521-1632
        int count = point_buffer.count;
        if (count > 0) {
            if (points_are_too_close(point_buffer.last().position, next.position)) {
                if (count == 1) may_need_empty_cap = true;
                *added = false;
                return true;
            }
            if (count == 1) {
                EndpointData& first = point_buffer.last();
                V2 edge = next.position - first.position;
                float len = length(edge);
                if (is_nan(next.advancement)) next.advancement = first.advancement + len;
                V2 tg = edge / len;
                V2 n = v2(-tg.y, tg.x) * next.half_width;
                first.side_points[SIDE_POSITIVE].next = first.position + n;
                first.side_points[SIDE_NEGATIVE].next = first.position - n;
            }
        }
        if (count > 1) {
            EndpointData& prev = point_buffer.last_but_one();
            EndpointData& join = point_buffer.last();
            vertex.src = join.src;
            vertex.position_on_path = join.position;
            vertex.half_width = join.half_width;
            bool fast_path = false;
            if (join.is_flattening_step) {
                V2 v0 = join.position - prev.position;
                V2 v1 = next.position - join.position;
                fast_path = dot(v0, v1) > 0.0f;
            }
            if (fast_path) {
                join.line_join = JOIN_MITER;
                bool skip_ignored;
                if (!flattened_step(prev, join, next, &skip_ignored)) return false;
            } else {
                compute_join_side_positions_fixed_width(prev, join, next, options.miter_limit);
                if (!add_join_base_vertices(join, true)) return false;
                if (!add_join_base_vertices(join, false)) return false;
            }
            if (count > 2) add_edge_triangles(prev, join);
            if (!tessellate_join(join)) return false;
            if (count == 2) {
                firsts[0] = prev;
                firsts[1] = join;
                n_firsts = 2;
            }
        }
        point_buffer.push(next);
        *added = true;
        return true;
    }

    // stroke.rs:1643-1733
    void compute_join_side_positions_fixed_width(const EndpointData& prev, EndpointData& join,
                                                 const EndpointData& next, float miter_limit) {
        V2 prev_tangent = join.position - prev.position;
        V2 next_tangent = next.position - join.position;
        float prev_length = length(prev_tangent);
        float next_length = length(next_tangent);
        prev_tangent = prev_tangent / prev_length;
        next_tangent = next_tangent / next_length;
        if (is_nan(join.advancement)) join.advancement = prev.advancement + prev_length;
        vertex.advancement = join.advancement;
        V2 normal = compute_normal(prev_tangent, next_tangent);
        int front_side;
        V2 front_normal;
        if (cross(prev_tangent, next_tangent) >= 0.0f) {
            front_side = SIDE_NEGATIVE;
            front_normal = -normal;
        } else {
            front_side = SIDE_POSITIVE;
            front_normal = normal;
        }
        int back_side = 1 - front_side;
        V2 extruded_normal = front_normal * vertex.half_width;
        bool unclipped_miter = (join.line_join == JOIN_MITER || join.line_join == JOIN_MITER_CLIP) &&
                               !miter_limit_is_exceeded(front_normal, miter_limit);
        bool fold = false;
        bool angle_is_sharp = dot(next_tangent, prev_tangent) < 0.0f;
        if (!unclipped_miter && angle_is_sharp) {
            float d_next = dot(extruded_normal, -next_tangent) - next_length;
            float d_prev = dot(extruded_normal, prev_tangent) - prev_length;
            if (fminf(d_next, d_prev) > 0.0f || sqlen(normal) < 1e-5f) {
                join.fold[front_side] = true;
                fold = true;
            }
        }
        V2 n0 = v2(-prev_tangent.y, prev_tangent.x) * vertex.half_width;
        V2 n1 = v2(-next_tangent.y, next_tangent.x) * vertex.half_width;
        join.side_points[SIDE_POSITIVE].prev = join.position + n0;
        join.side_points[SIDE_POSITIVE].next = join.position + n1;
        join.side_points[SIDE_NEGATIVE].prev = join.position - n0;
        join.side_points[SIDE_NEGATIVE].next = join.position - n1;
        if (!fold) {
            V2 miter_pos[2] = {join.position + normal * vertex.half_width, join.position - normal * vertex.half_width};
            join.side_points[back_side].has_single = true;
            join.side_points[back_side].single = miter_pos[back_side];
            if (unclipped_miter) {
                join.side_points[front_side].has_single = true;
                join.side_points[front_side].single = miter_pos[front_side];
            } else if (join.line_join == JOIN_MITER_CLIP) {
                V2 c0 = join.side_points[front_side].prev - join.position;
                V2 c1 = join.side_points[front_side].next - join.position;
                V2 pn, nn;
                get_clip_intersections(c0, c1, front_normal, miter_limit * vertex.half_width, &pn, &nn);
                join.side_points[front_side].prev = join.position + pn;
                join.side_points[front_side].next = join.position + nn;
            }
        }
    }

    // stroke.rs:1743-1806
    bool flattened_step(EndpointData& prev, EndpointData& join, EndpointData& next, bool* skip) {
        V2 prev_edge = join.position - prev.position;
        float prev_length = length(prev_edge);
        V2 prev_tangent = prev_edge / prev_length;
        V2 next_edge = next.position - join.position;
        float next_length = length(next_edge);
        V2 next_tangent = next_edge / next_length;
        V2 normal = compute_normal(prev_tangent, next_tangent);
        if (is_nan(join.advancement)) join.advancement = prev.advancement + prev_length;
        if (is_nan(next.advancement)) next.advancement = join.advancement + next_length;
        vertex.advancement = join.advancement;
        V2 p0 = join.position + normal * vertex.half_width;
        join.side_points[SIDE_POSITIVE].prev = p0;
        join.side_points[SIDE_POSITIVE].next = p0;
        join.side_points[SIDE_POSITIVE].has_single = true;
        join.side_points[SIDE_POSITIVE].single = p0;
        V2 p1 = join.position - normal * vertex.half_width;
        join.side_points[SIDE_NEGATIVE].prev = p1;
        join.side_points[SIDE_NEGATIVE].next = p1;
        join.side_points[SIDE_NEGATIVE].has_single = true;
        join.side_points[SIDE_NEGATIVE].single = p1;
        V2 v0 = p0 - prev.side_points[SIDE_POSITIVE].next;
        V2 v1 = p1 - prev.side_points[SIDE_NEGATIVE].next;
        if (dot(prev_edge, v0) < 0.0f && dot(prev_edge, v1) < 0.0f) {
            *skip = true;
            return true;
        }
        *skip = false;
        vertex.normal = normal;
        vertex.side_negative = false;
        uint32_t pos_vertex, neg_vertex;
        if (!emit_vertex(&pos_vertex)) return false;
        vertex.normal = -normal;
        vertex.side_negative = true;
        if (!emit_vertex(&neg_vertex)) return false;
        join.side_points[SIDE_POSITIVE].prev_vertex = pos_vertex;
        join.side_points[SIDE_POSITIVE].next_vertex = pos_vertex;
        join.side_points[SIDE_NEGATIVE].prev_vertex = neg_vertex;
        join.side_points[SIDE_NEGATIVE].next_vertex = neg_vertex;
        return true;
    }

    // stroke.rs:1809-1836
    void compute_edge_attachment_positions(EndpointData& p0, EndpointData& p1) {
        V2 edge = p1.position - p0.position;
        float d = length(edge);
        float edge_angle = angle_from_x_axis(edge);
        float sin_vwidth_angle = (p1.half_width - p0.half_width) / d;
        float vwidth_angle = libm_asin(sin_vwidth_angle);
        if (is_nan(vwidth_angle)) vwidth_angle = 0.0f;
        compute_side_attachment_positions(p0, p1, edge_angle, vwidth_angle, SIDE_POSITIVE);
        compute_side_attachment_positions(p0, p1, edge_angle, vwidth_angle, SIDE_NEGATIVE);
        if (is_nan(p1.advancement)) p1.advancement = p0.advancement + d;
    }
    // stroke.rs:1838-1866
    void compute_side_attachment_positions(EndpointData& p0, EndpointData& p1, float edge_angle, float vwidth_angle,
                                           int side) {
        float nl = side_sign(side);
        float normal_angle = edge_angle + nl * (PI_F * 0.5f + vwidth_angle);
        V2 normal = v2(libm_cos(normal_angle), libm_sin(normal_angle));
        p0.side_points[side].next = p0.position + normal * p0.half_width;
        p1.side_points[side].prev = p1.position + normal * p1.half_width;
    }

    // stroke.rs:1869-1904
    void add_edge_triangles(const EndpointData& p0, const EndpointData& p1) {
        uint32_t p0_neg = p0.side_points[SIDE_NEGATIVE].next_vertex;
        uint32_t p0_pos = p0.side_points[SIDE_POSITIVE].next_vertex;
        uint32_t p1_neg = p1.side_points[SIDE_NEGATIVE].prev_vertex;
        uint32_t p1_pos = p1.side_points[SIDE_POSITIVE].prev_vertex;
        if (p0.fold[SIDE_POSITIVE]) p0_neg = p0.side_points[SIDE_POSITIVE].prev_vertex;
        if (p0.fold[SIDE_NEGATIVE]) p0_pos = p0.side_points[SIDE_NEGATIVE].prev_vertex;
        if (p1.fold[SIDE_POSITIVE]) p1_neg = p1.side_points[SIDE_POSITIVE].next_vertex;
        if (p1.fold[SIDE_NEGATIVE]) p1_pos = p1.side_points[SIDE_NEGATIVE].next_vertex;
        if (p0_neg == p1_pos) return;
        if (p0_neg != p0_pos && p0_pos != p1_pos) output.add_triangle(p0_neg, p0_pos, p1_pos);
        if (p0_neg != p1_neg && p1_pos != p1_neg) output.add_triangle(p0_neg, p1_pos, p1_neg);
    }

    // stroke.rs:1907-1965
    bool tessellate_join(EndpointData& join) {
        bool need[2] = {!join.side_points[SIDE_POSITIVE].has_single && !join.fold[SIDE_NEGATIVE],
                        !join.side_points[SIDE_NEGATIVE].has_single && !join.fold[SIDE_POSITIVE]};
        if (!join.fold[SIDE_POSITIVE] && !join.fold[SIDE_NEGATIVE]) {
            const SidePoints& sp = join.side_points[SIDE_POSITIVE];
            const SidePoints& sn = join.side_points[SIDE_NEGATIVE];
            if (need[0] && need[1]) {
                output.add_triangle(sp.prev_vertex, sp.next_vertex, sn.next_vertex);
                output.add_triangle(sp.prev_vertex, sn.next_vertex, sn.prev_vertex);
            } else if (!need[0] && need[1]) {
                output.add_triangle(sn.prev_vertex, sp.prev_vertex, sn.next_vertex);
            } else if (need[0] && !need[1]) {
                output.add_triangle(sn.prev_vertex, sp.prev_vertex, sp.next_vertex);
            }
        }
        for (int side = 0; side < 2; ++side) {
            if (!need[side]) continue;
            if (join.line_join == JOIN_ROUND)
                if (!tessellate_round_join(join, side)) return false;
        }
        return true;
    }

    // stroke.rs:1968-2021
    bool tessellate_round_join(EndpointData& join, int side) {
        V2 center = join.position;
        float radius = join.half_width;
        V2 start_normal = join.side_points[side].prev - center;
        V2 end_normal = join.side_points[side].next - center;
        uint32_t start_vertex = join.side_points[side].prev_vertex;
        uint32_t end_vertex = join.side_points[side].next_vertex;
        float angle_sign = side == SIDE_NEGATIVE ? 1.0f : -1.0f;
        float start_angle = angle_from_x_axis(start_normal);
        float diff = angle_to(start_angle, angle_from_x_axis(end_normal));
        if (diff * angle_sign < 0.0f) diff = angle_sign * (2.0f * PI_F - fabsf(diff));
        float end_angle = start_angle + diff;
        if (side == SIDE_NEGATIVE) {
            float ta = start_angle;
            start_angle = end_angle;
            end_angle = ta;
            uint32_t tv = start_vertex;
            start_vertex = end_vertex;
            end_vertex = tv;
        }
        float stp = circle_flattening_step(radius, options.tolerance);
        float num_segments = ceilf(fabsf(diff) / stp);
        uint32_t num_subdivisions = sat_u32(roundf(libm_log2(num_segments)));
        vertex.side_negative = side != SIDE_POSITIVE;
        return tessellate_arc(start_angle, end_angle, start_vertex, end_vertex, num_subdivisions);
    }

    // stroke.rs:2024-2055
    bool add_join_base_vertices(EndpointData& join, bool negative) {
        vertex.side_negative = negative;
        int side = negative ? SIDE_NEGATIVE : SIDE_POSITIVE;
        SidePoints& sp = join.side_points[side];
        if (sp.has_single) {
            vertex.normal = (sp.single - join.position) / join.half_width;
            uint32_t id;
            if (!emit_vertex(&id)) return false;
            sp.prev_vertex = id;
            sp.next_vertex = id;
        } else {
            uint32_t a, b;
            vertex.normal = (sp.prev - join.position) / join.half_width;
            if (!emit_vertex(&a)) return false;
            vertex.normal = (sp.next - join.position) / join.half_width;
            if (!emit_vertex(&b)) return false;
            sp.prev_vertex = a;
            sp.next_vertex = b;
        }
        return true;
    }

    // stroke.rs:2060-2128
    void compute_join_side_positions(const EndpointData& prev, EndpointData& join, const EndpointData& next,
                                     float miter_limit, int side) {
        float sign = side_sign(side);
        V2 v0 = normalize(join.side_points[side].prev - prev.side_points[side].next);
        V2 v1 = normalize(next.side_points[side].prev - join.side_points[side].next);
        bool inward = cross(v0, v1) * sign > 0.0f;
        bool forward = dot(v0, v1) > 0.0f;
        V2 normal = compute_normal(v0, v1) * sign;
        V2 path_v0 = normalize(join.position - prev.position);
        V2 path_v1 = normalize(next.position - join.position);
        bool normal_same_side = dot(v0 + v1, path_v0 + path_v1) >= 0.0f;
        bool angle_is_sharp = inward && !forward && normal_same_side;
        if (angle_is_sharp) {
            V2 extruded_normal = normal * join.half_width;
            float prev_length = join.advancement - prev.advancement;
            float next_length = next.advancement - join.advancement;
            float d_next = dot(extruded_normal, v1) - next_length;
            float d_prev = dot(extruded_normal, -v0) - prev_length;
            if (fminf(d_next, d_prev) > 0.0f || sqlen(normal) < 1e-5f) join.fold[side] = true;
        }
        bool concave = inward && normal_same_side && !join.fold[side];
        if (concave || ((join.line_join == JOIN_MITER || join.line_join == JOIN_MITER_CLIP) &&
                        !miter_limit_is_exceeded(normal, miter_limit))) {
            join.side_points[side].has_single = true;
            join.side_points[side].single = join.position + normal * join.half_width;
        } else if (join.line_join == JOIN_MITER_CLIP) {
            V2 n0 = join.side_points[side].prev - join.position;
            V2 n1 = join.side_points[side].next - join.position;
            V2 pn, nn;
            get_clip_intersections(n0, n1, normal, miter_limit * join.half_width, &pn, &nn);
            join.side_points[side].prev = join.position + pn;
            join.side_points[side].next = join.position + nn;
        }
    }

    // stroke.rs:2130-2206
    bool tessellate_last_edge(const EndpointData& p0, EndpointData& p1, bool is_first_edge) {
        V2 v = p1.position - p0.position;
        p1.advancement = p0.advancement + length(v);
        vertex.src = p1.src;
        vertex.position_on_path = p1.position;
        vertex.advancement = p1.advancement;
        vertex.half_width = p1.half_width;
        for (int side = 0; side < 2; ++side) {
            V2 side_position = p1.side_points[side].prev;
            bool has_clip = options.end_cap == CAP_SQUARE || options.end_cap == CAP_BUTT;
            float clip = options.end_cap == CAP_SQUARE ? p1.half_width : 0.0f;
            if (has_clip) {
                V2 normal = normalize(p1.position - p0.position);
                LineD clip_line = line_d(p1.position + normal * clip, tangent(normal));
                LineD side_line = line_d(side_position, side_position - p0.side_points[side].next);
                V2 inter;
                if (!line_intersection(clip_line, side_line, &inter)) inter = p1.side_points[side].prev;
                p1.side_points[side].prev = inter;
            }
            vertex.side_negative = side != SIDE_POSITIVE;
            vertex.normal = (p1.side_points[side].prev - p1.position) / p1.half_width;
            uint32_t id;
            if (!emit_vertex(&id)) return false;
            p1.side_points[side].prev_vertex = id;
        }
        if (!is_first_edge) add_edge_triangles(p0, p1);
        if (options.end_cap == CAP_ROUND) {
            if (!tessellate_round_cap(p1.position, p1.half_width, p1.side_points[SIDE_POSITIVE].prev - p1.position,
                                      p1.side_points[SIDE_POSITIVE].prev_vertex,
                                      p1.side_points[SIDE_NEGATIVE].prev_vertex, v, false))
                return false;
        }
        return true;
    }

    // stroke.rs:2208-2276
    bool tessellate_first_edge(EndpointData& first, const EndpointData& second) {
        vertex.src = first.src;
        vertex.position_on_path = first.position;
        vertex.advancement = first.advancement;
        vertex.half_width = first.half_width;
        for (int side = 0; side < 2; ++side) {
            V2 side_position = first.side_points[side].next;
            bool has_clip = options.start_cap == CAP_SQUARE || options.start_cap == CAP_BUTT;
            float clip = options.start_cap == CAP_SQUARE ? first.half_width : 0.0f;
            if (has_clip) {
                V2 normal = normalize(first.position - second.position);
                LineD clip_line = line_d(first.position + normal * clip, tangent(normal));
                LineD side_line = line_d(side_position, side_position - second.side_points[side].prev);
                V2 inter;
                if (!line_intersection(clip_line, side_line, &inter)) inter = first.side_points[side].next;
                side_position = inter;
            }
            vertex.side_negative = side != SIDE_POSITIVE;
            vertex.normal = (side_position - first.position) / first.half_width;
            uint32_t id;
            if (!emit_vertex(&id)) return false;
            first.side_points[side].next_vertex = id;
        }
        add_edge_triangles(first, second);
        if (options.start_cap == CAP_ROUND) {
            return tessellate_round_cap(first.position, first.half_width,
                                        first.side_points[SIDE_NEGATIVE].next - first.position,
                                        first.side_points[SIDE_NEGATIVE].next_vertex,
                                        first.side_points[SIDE_POSITIVE].next_vertex,
                                        first.position - second.position, true);
        }
        return true;
    }

    // stroke.rs:2459-2524
    bool tessellate_round_cap(V2 center, float radius, V2 start_normal, uint32_t start_vertex, uint32_t end_vertex,
                              V2 edge_normal, bool is_start) {
        if (radius < options.tolerance) return true;
        bool first_side_positive = is_start ^ (cross(edge_normal, start_normal) >= 0.0f);
        float start_angle = angle_from_x_axis(start_normal);
        float diff = angle_to(start_angle, angle_from_x_axis(edge_normal));
        float mid_angle = start_angle + diff;
        float end_angle = mid_angle + diff;
        float stp = circle_flattening_step(radius, options.tolerance);
        float num_segments = ceilf(fabsf(diff) / stp);
        uint32_t num_subdivisions = sat_u32(roundf(libm_log2(num_segments)));
        vertex.position_on_path = center;
        vertex.half_width = radius;
        vertex.side_negative = !first_side_positive;
        vertex.normal = normalize(edge_normal);
        uint32_t mid_vertex;
        if (!emit_vertex(&mid_vertex)) return false;
        output.add_triangle(start_vertex, mid_vertex, end_vertex);
        if (!tessellate_arc(start_angle, mid_angle, start_vertex, mid_vertex, num_subdivisions)) return false;
        vertex.side_negative = first_side_positive;  // first_side.opposite()
        return tessellate_arc(mid_angle, end_angle, mid_vertex, end_vertex, num_subdivisions);
    }

    // stroke.rs:2526-2558
    bool tessellate_empty_square_cap(V2 position) {
        vertex.position_on_path = position;
        uint32_t a, b, c, d;
        vertex.normal = v2(1.0f, 1.0f);
        vertex.side_negative = true;
        if (!emit_vertex(&a)) return false;
        vertex.normal = v2(1.0f, -1.0f);
        vertex.side_negative = false;
        if (!emit_vertex(&b)) return false;
        vertex.normal = v2(-1.0f, -1.0f);
        vertex.side_negative = false;
        if (!emit_vertex(&c)) return false;
        vertex.normal = v2(-1.0f, 1.0f);
        vertex.side_negative = true;
        if (!emit_vertex(&d)) return false;
        output.add_triangle(a, b, c);
        output.add_triangle(a, c, d);
        return true;
    }

    // stroke.rs:2560-2609
    bool tessellate_empty_round_cap(V2 center) {
        float radius = vertex.half_width;
        vertex.position_on_path = center;
        uint32_t left_id, right_id;
        vertex.normal = v2(-1.0f, 0.0f);
        vertex.side_negative = false;
        if (!emit_vertex(&left_id)) return false;
        vertex.normal = v2(1.0f, 0.0f);
        vertex.side_negative = true;
        if (!emit_vertex(&right_id)) return false;
        if (!tessellate_round_cap(center, radius, v2(-1.0f, 0.0f), left_id, right_id, v2(0.0f, 1.0f), true))
            return false;
        return tessellate_round_cap(center, radius, v2(1.0f, 0.0f), right_id, left_id, v2(0.0f, -1.0f), false);
    }

    // stroke.rs:2612-2653 (pre-order recursion)
    bool tessellate_arc(float a0, float a1, uint32_t va, uint32_t vb, uint32_t num_recursions) {
        if (num_recursions == 0) return true;
        float mid_angle = (a0 + a1) * 0.5f;
        vertex.normal = v2(libm_cos(mid_angle), libm_sin(mid_angle));
        uint32_t id;
        if (!emit_vertex(&id)) return false;
        output.add_triangle(va, id, vb);
        if (!tessellate_arc(a0, mid_angle, va, id, num_recursions - 1)) return false;
        return tessellate_arc(mid_angle, a1, id, vb, num_recursions - 1);
    }
};

// Convenience: tessellate one path into `out` (one StrokeTessellator::tessellate_path call).
static inline int stroke_path(const uint8_t* verbs, uint32_t n_verbs, const float* pts, uint32_t num_attributes,
                              const Options& opts, Buffers& out) {
    Stroker<Buffers> s(opts, out);
    return s.tessellate_path(verbs, n_verbs, pts, num_attributes);
}

}  // namespace lyon_oracle
