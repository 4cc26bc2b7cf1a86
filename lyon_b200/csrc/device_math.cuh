// device_math.cuh -- f32 vector math and lyon_geom flattening primitives for sm_100a.
//
// Every routine keeps the reference's f32 operation order (mul then add, never fused: this file must
// be compiled with -fmad=false, IEEE div/sqrt) so that flatten step counts and every branch taken by
// the stroker agree bit-for-bit with lyon.  Citations are file:line under nical/lyon @ 8071ec06.
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define LYC_HD __host__ __device__ __forceinline__
// Real function calls for the heavy, rarely taken routines (caps, closing joins, arcs): the tile kernel's code must
// stay small enough for the instruction cache.
#define LYC_NOINLINE static __host__ __device__ __noinline__
#else
// Host-only build (tests/emul): the same source compiled by g++ -ffp-contract=off emulates the kernels.
#define LYC_HD inline
#define LYC_NOINLINE static inline
#endif

#if defined(__FAST_MATH__) || defined(__USE_FAST_MATH__)
#error "lyon_cuda kernels must not be compiled with --use_fast_math"
#endif

namespace lyc {

// 8-byte aligned: one 64-bit load / store per point (LDS.64 / LDG.64); every point array handed to the library is at least
// 8-byte aligned (the C entry points check).
struct alignas(8) V2 {
    float x, y;
};
LYC_HD V2 v2(float x, float y) { return V2{x, y}; }
LYC_HD V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
LYC_HD V2 operator-(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
LYC_HD V2 operator-(V2 a) { return V2{-a.x, -a.y}; }
LYC_HD V2 operator*(V2 a, float s) { return V2{a.x * s, a.y * s}; }
LYC_HD V2 operator/(V2 a, float s) { return V2{a.x / s, a.y / s}; }
// euclid 0.22 Vector2D::{dot,cross,square_length,length,normalize}
LYC_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
LYC_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
LYC_HD float sqlen(V2 a) { return a.x * a.x + a.y * a.y; }
LYC_HD float length(V2 a) { return sqrtf(sqlen(a)); }
LYC_HD V2 normalize(V2 a) { return a / length(a); }
// euclid Point2D::lerp
LYC_HD V2 lerp(V2 a, V2 b, float t) {
    float one_t = 1.0f - t;
    return V2{one_t * a.x + t * b.x, one_t * a.y + t * b.y};
}
LYC_HD bool is_nan(float v) { return v != v; }

#define LYC_PI_F 3.14159274101257324f
#define LYC_FRAC_PI_2_F 1.57079637050628662f

// euclid Trig::fast_atan2 (f32 polynomial approximation)
LYC_HD float fast_atan2(float y, float x) {
    float x_abs = fabsf(x);
    float y_abs = fabsf(y);
    float a = fminf(x_abs, y_abs) / fmaxf(x_abs, y_abs);
    float s = a * a;
    float result = ((-0.0464964749f * s + 0.15931422f) * s - 0.327622764f) * s * a + a;
    if (y_abs > x_abs) result = LYC_FRAC_PI_2_F - result;
    if (x < 0.0f) result = LYC_PI_F - result;
    if (y < 0.0f) result = -result;
    return result;
}
LYC_HD float angle_from_x_axis(V2 v) { return fast_atan2(v.y, v.x); }
// fmodf(x, m) for m > 0, exact like the libm routine.  |x| < m returns x; m <= |x| < 2m returns x -/+ m, which is
// exact (Sterbenz: m <= |x| <= 2m); everything else (incl. NaN/inf) goes to fmodf.
LYC_NOINLINE float fmod_rare(float x, float m) { return fmodf(x, m); }
LYC_HD float fmod_small(float x, float m) {
    float ax = fabsf(x);
    if (ax < m) return x;
    if (ax < 2.0f * m) return x < 0.0f ? x + m : x - m;
    return fmod_rare(x, m);
}
// euclid Angle::angle_to
LYC_HD float angle_to(float self, float to) {
    float mx = LYC_PI_F * 2.0f;
    float d = fmod_small(to - self, mx);
    return fmod_small(2.0f * d, mx) - d;
}
// Rust f32::{cos,sin,acos,asin,log2} call the platform libm; both sides evaluate in double and round
// once (DESIGN.md "Transcendentals").
LYC_NOINLINE float libm_cos(float a) { return (float)cos((double)a); }
LYC_NOINLINE float libm_sin(float a) { return (float)sin((double)a); }
LYC_NOINLINE float libm_acos(float a) { return (float)acos((double)a); }
LYC_NOINLINE float libm_asin(float a) { return (float)asin((double)a); }
LYC_NOINLINE float libm_log2(float a) { return (float)log2((double)a); }
LYC_NOINLINE V2 libm_cos_sin(float a) {  // (cos a, sin a)
    double s, c;
    sincos((double)a, &s, &c);
    return V2{(float)c, (float)s};
}

// Rust `as u32` (saturating) and num-traits `to_u32().unwrap_or(1)`
LYC_HD uint32_t sat_u32(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
LYC_HD uint32_t to_u32_or_1(float n) {
    if (n > -1.0f && n < 4294967296.0f) return (uint32_t)n;
    return 1u;
}

// crates/geom/src/cubic_bezier.rs:1385-1402 / quadratic_bezier.rs:956-973: Wang count from err^4.
// The LUT holds i^4 for i = 1..24; the linear search `first i with err4 <= i^4` is evaluated as a
// branch-free count of LUT entries strictly below err4 (identical result, no divergence).
LYC_HD float wang_segments_from_err4(float err4) {
    if (err4 <= 331776.0f) {
        int c = 0;
#pragma unroll
        for (int i = 1; i <= 24; ++i) {
            float q = (float)(i * i) * (float)(i * i);  // exact in f32 (<= 331776)
            c += (err4 > q) ? 1 : 0;
        }
        return (float)(c + 1);
    }
    return fmaxf(sqrtf(sqrtf(err4)), 1.0f);
}
// quadratic_bezier.rs:943-974
LYC_HD float quad_wang(V2 from, V2 ctrl, V2 to, float tol) {
    V2 l = (from - ctrl * 2.0f + to) * 2.0f;
    float d = 1.0f / (8.0f * tol);
    float err4 = dot(l, l) * d * d;
    return wang_segments_from_err4(err4);
}
// cubic_bezier.rs:1369-1403
LYC_HD float cubic_wang(V2 from, V2 c1, V2 c2, V2 to, float tol) {
    V2 a = (from - c1 * 2.0f + c2) * 6.0f;
    V2 b = (c1 - c2 * 2.0f + to) * 6.0f;
    float l = fmaxf(dot(a, a), dot(b, b));
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
// quadratic_bezier.rs:891-900
LYC_HD QuadPoly quad_poly(V2 from, V2 ctrl, V2 to) {
    QuadPoly p;
    p.a0 = from;
    p.a1 = (ctrl - from) * 2.0f;
    p.a2 = from + to - ctrl * 2.0f;
    return p;
}
// quadratic_bezier.rs:927-939
LYC_HD V2 quad_poly_sample(const QuadPoly& p, float t) {
    V2 v = p.a0;
    float t2 = t;
    v = v + p.a1 * t2;
    t2 = t2 * t;
    v = v + p.a2 * t2;
    return v;
}
// cubic_bezier.rs:1318-1325
LYC_HD CubicPoly cubic_poly(V2 from, V2 c1, V2 c2, V2 to) {
    CubicPoly p;
    p.a0 = from;
    p.a1 = (c1 - from) * 3.0f;
    p.a2 = from * 3.0f - c1 * 6.0f + c2 * 3.0f;
    p.a3 = to - from + (c1 - c2) * 3.0f;
    return p;
}
// cubic_bezier.rs:1351-1365
LYC_HD V2 cubic_poly_sample(const CubicPoly& p, float t) {
    V2 v = p.a0;
    float t2 = t;
    v = v + p.a1 * t2;
    t2 = t2 * t;
    v = v + p.a2 * t2;
    t2 = t2 * t;
    v = v + p.a3 * t2;
    return v;
}
// The reference accumulates t (t1 = t0 + step, quadratic_bezier.rs:382-388 / cubic_bezier.rs:605-611);
// k*step differs from the k-fold sum in f32, so point k replays k+1 additions.
LYC_HD float accumulate_t(float step, uint32_t k_plus_1) {
    float t = 0.0f;
    for (uint32_t i = 0; i < k_plus_1; ++i) t = t + step;
    return t;
}

// quadratic_bezier.rs:35-41
LYC_HD V2 quad_sample(V2 from, V2 ctrl, V2 to, float t) {
    float t2 = t * t;
    float one_t = 1.0f - t;
    float one_t2 = one_t * one_t;
    return from * one_t2 + ctrl * 2.0f * one_t * t + to * t2;
}
// quadratic_bezier.rs:180-191
LYC_HD bool quad_local_x_extremum_t(V2 from, V2 ctrl, V2 to, float* t_out) {
    float div = from.x - 2.0f * ctrl.x + to.x;
    if (div == 0.0f) return false;
    float t = (from.x - ctrl.x) / div;
    if (t > 0.0f && t < 1.0f) {
        *t_out = t;
        return true;
    }
    return false;
}
// crates/tessellation/src/stroke.rs:3178-3212
// find_sharp_turn (stroke.rs:3178-3212) in two steps: the cheap test that sends most quadratics away (false), and the
// search itself (rotation onto the long axis, local x extremum) for the candidates it leaves.
LYC_HD bool sharp_turn_candidate(V2 from, V2 ctrl, V2 to, V2* long_axis) {
    V2 baseline = to - from;
    V2 v = ctrl - from;
    V2 n = v2(-baseline.y, baseline.x);
    float v_dot_b = dot(v, baseline);
    float v_dot_n = dot(v, n);
    if ((v_dot_b >= 0.0f && v_dot_b <= dot(baseline, baseline)) || fabsf(v_dot_n) * 2.0f >= fabsf(v_dot_b)) {
        if (sqlen(baseline) * 30.0f > sqlen(v)) return false;
        *long_axis = v;
    } else {
        *long_axis = baseline;
    }
    return true;
}
LYC_HD bool find_sharp_turn(V2 from, V2 ctrl, V2 to, float* t_out) {
    V2 long_axis;
    if (!sharp_turn_candidate(from, ctrl, to, &long_axis)) return false;
    V2 baseline = to - from;
    V2 v = ctrl - from;
    float ang = -angle_from_x_axis(long_axis);
    const V2 cs = libm_cos_sin(ang);  // one double-precision sincos: the same two values as cos() and sin()
    float s = cs.y, c = cs.x;
    V2 rctrl = v2(v.x * c - v.y * s, v.y * c + v.x * s);
    V2 rto = v2(baseline.x * c - baseline.y * s, baseline.y * c + baseline.x * s);
    return quad_local_x_extremum_t(v2(0.0f, 0.0f), rctrl, rto, t_out);
}

// crates/geom/src/line.rs:562-578 evaluated in f64 (f64 EPSILON 1e-8, lib.rs:224); result narrowed to f32.
struct LineD {
    double px, py, vx, vy;
};
LYC_HD LineD line_d(V2 p, V2 v) {
    return LineD{(double)p.x, (double)p.y, (double)v.x, (double)v.y};
}
LYC_HD bool line_intersection(const LineD& a, const LineD& b, V2* out) {
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
LYC_HD V2 tangent(V2 v) { return v2(-v.y, v.x); }  // geom/src/utils.rs:15-17

// crates/tessellation/src/math_utils.rs:16-40
LYC_HD V2 compute_normal(V2 v1, V2 v2_) {
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

}  // namespace lyc
