// stroke_core.cuh -- per-item logic of the batched stroke tessellator (host+device).
//
// The GPU formulation of lyon's streaming StrokeTessellator (crates/tessellation/src/stroke.rs):
//
//   verbs ──K1 seg_setup──▶ per-verb point offset + number of FED SLOTS (Wang counts)   [scan]
//   slots ──K2 point_gen──▶ per-slot position / t / edge length / "maybe too close" flag
//   paths ──K3 path_walk──▶ sequential-in-path f32 chains: merge filter, advancement, sub-path records
//   slots ──K4 count──────▶ vertices / triangles owned by each slot                       [scan]
//   slots ──K5 emit───────▶ StrokeVertexGpu[] and u32 indices
//
// A SLOT is one call of `fixed_width_step` (stroke.rs:1635) -- a Begin point, a line endpoint, one
// flattened curve point -- or one `end()` (stroke.rs:1179).  Vertex ids are handed out in the exact
// order of the reference's add_stroke_vertex calls: the join at kept point i is owned by slot i, and
// everything `end()` emits (last/first edge + caps, or the closing joins) is owned by the End slot.
//
// Everything here is `LYC_HD` so tests/emul can run the very same code on the CPU.
#pragma once
#include "device_math.cuh"

namespace lyc {

enum : uint8_t { VERB_LINE = 0, VERB_QUAD = 1, VERB_CUBIC = 2, VERB_BEGIN = 3, VERB_CLOSE = 4, VERB_END = 5 };
enum : uint8_t { CAP_BUTT = 0, CAP_SQUARE = 1, CAP_ROUND = 2 };
enum : uint8_t { JOIN_MITER = 0, JOIN_MITER_CLIP = 1, JOIN_ROUND = 2, JOIN_BEVEL = 3 };

static const uint32_t SEG_FS_BIT = 0x80000000u;     // slot is a flattening step (stroke.rs:1044)
static const uint32_t SEG_END_BIT = 0x40000000u;    // slot belongs to an End/Close verb
static const uint32_t SEG_BEGIN_BIT = 0x20000000u;  // slot belongs to a Begin verb
static const uint32_t SEG_MASK = 0x1FFFFFFFu;
static const uint32_t SIDE_NEG_BIT = 0x80000000u;
static const int SIDE_POSITIVE = 0;
static const int SIDE_NEGATIVE = 1;
static const uint32_t MAX_ARC_SUBDIVISIONS = 20;  // reference recursion is unbounded; 2^20 vertices per arc is our cap

struct StrokeParams {
    uint8_t start_cap, end_cap, line_join, _pad;
    float line_width, half_width, miter_limit, tolerance;
    float arc_step;   // circle_flattening_step(half_width, tolerance): constant for a fixed line width (host computed)
    float merge_thr;  // square_merge_threshold (stroke.rs:603-605)
    float flag_thr;   // 5 * merge_thr: conservative pre-filter, see point_gen
    uint32_t num_attributes, attr_stride;
    int32_t vw_attr;
};

struct GpuVertex {  // == lyon_stroke_vertex
    float px, py, nx, ny, advancement, half_width;
    uint32_t src;
    float t;
};

struct Views {
    const uint8_t* verbs;
    const V2* points;
    uint32_t n_verbs;
    // per verb (n_verbs + 1)
    uint32_t* seg_pt;    // index of the verb's first stored point
    uint32_t* seg_off;   // first slot of the verb
    float* seg_tsplit;   // quadratic sharp-turn split parameter, or -1
    uint32_t* seg_id;    // EndpointId of the verb's endpoint (written by path_walk)
    // per slot
    V2* pos;
    float* t;
    uint32_t* segfs;
    float* len;   // |P_s - P_{s-1}|; sign bit set when the pair MAY be within the merge threshold
    float* adv;   // advancement; NaN = point dropped by the merge filter
    uint32_t* voff;
    uint32_t* toff;
    uint32_t n_slots;
    // accessors shared with TileViews (the per-item code below is written against these only)
    LYC_HD V2 P(uint32_t s) const { return pos[s]; }
    LYC_HD float A(uint32_t s) const { return adv[s]; }
    LYC_HD float T(uint32_t s) const { return t[s]; }
    LYC_HD uint32_t F(uint32_t s) const { return segfs[s]; }  // SEG_*_BIT flags in the top bits
    LYC_HD uint32_t SRC(uint32_t s) const { return seg_id[segfs[s] & SEG_MASK]; }
    LYC_HD uint32_t VO(uint32_t s) const { return voff[s]; }
    LYC_HD bool D(uint32_t s) const { return adv[s] != adv[s]; }  // dropped by the merge filter: NaN advancement
};

// The same slot arrays for ONE TILE of whole paths held in shared memory by the fused kernel (stroke_tile.cuh):
// slot indices are tile-relative, flags are one byte (SEG_*_BIT >> 24), the endpoint id is stored per slot.
static const uint8_t TF_FS = 0x80, TF_END = 0x40, TF_BEGIN = 0x20, TF_CLOSE = 0x10, TF_FIRST = 0x08, TF_EVAL = 0x04, TF_CUBIC = 0x02,
                     TF_LAST = 0x01;
// CTA tiles (TileViews) use two of the low bits for the merge filter: FLAGGED = the point MAY be within the merge distance of
// its predecessor (written with the slot, never changed), DROPPED = the merge filter dropped it.  While the runs are being
// resolved (tile_resolve_run) the dropped points are collected in `dropmask`, one bit per slot set with an atomic OR, so the
// flag bytes stay read-only in that phase and its threads share nothing they write (compute-sanitizer racecheck clean);
// tile_merge_dropped then copies the bits into the flag bytes, where every later phase reads them with the byte it loads
// anyway.
static const uint8_t TF_FLAGGED = 0x02, TF_DROPPED = 0x04;
#ifdef __CUDA_ARCH__
#define LYC_ATOMIC_OR_U32(ptr, v) atomicOr((ptr), (v))
#else
#define LYC_ATOMIC_OR_U32(ptr, v) (*(ptr) |= (v))
#endif
// One slot array of a CTA tile.  Under nvcc it is a 32-bit byte offset into the CTA's dynamic shared memory: the accessors
// go through the `extern __shared__` symbol, so every access compiles to LDS / STS with 32-bit address arithmetic (a
// pointer kept in a struct that is passed to out-of-line routines degrades to generic LD / ST with 64-bit address
// arithmetic: 15 % of the tile kernel's instructions in round 1).  On the CPU (tests/emul) it is a plain pointer.
#ifdef __CUDACC__
extern __shared__ __align__(16) unsigned char lyc_tile_smem[];
template <class T>
struct TArr {
    uint32_t off;
    __host__ __device__ __forceinline__ T* ptr() const {
#ifdef __CUDA_ARCH__
        return reinterpret_cast<T*>(lyc_tile_smem + off);
#else
        return nullptr;
#endif
    }
    __host__ __device__ __forceinline__ T& operator[](uint32_t i) const { return ptr()[i]; }
    __host__ __device__ __forceinline__ T* operator+(uint32_t i) const { return ptr() + i; }
    __host__ __device__ __forceinline__ explicit operator bool() const { return true; }
    template <class U>
    __host__ __device__ __forceinline__ TArr<U> as() const {
        TArr<U> r;
        r.off = off;
        return r;
    }
};
#else
template <class T>
struct TArr {
    T* p;
    TArr& operator=(T* q) {
        p = q;
        return *this;
    }
    T* ptr() const { return p; }
    T& operator[](uint32_t i) const { return p[i]; }
    T* operator+(uint32_t i) const { return p + i; }
    explicit operator bool() const { return p != nullptr; }
    template <class U>
    TArr<U> as() const {
        TArr<U> r;
        r.p = reinterpret_cast<U*>(p);
        return r;
    }
};
#endif
struct TileViews {
    TArr<V2> pos;
    TArr<float> t;
    TArr<float> adv;
    TArr<uint32_t> src;   // EndpointId of the slot's vertex source
    TArr<uint8_t> flags;  // TF_*
    TArr<uint32_t> voff;
    TArr<uint32_t> toff;
    TArr<uint8_t> jcode;
    TArr<uint32_t> dropmask;  // one bit per slot: dropped by the merge filter (until tile_merge_dropped)
    uint32_t n_slots;
    LYC_HD V2 P(uint32_t s) const { return pos[s]; }
    LYC_HD float A(uint32_t s) const { return adv[s]; }
    LYC_HD float T(uint32_t s) const { return t[s]; }
    LYC_HD uint32_t F(uint32_t s) const { return (uint32_t)flags[s] << 24; }
    LYC_HD uint32_t SRC(uint32_t s) const { return src[s]; }
    LYC_HD uint32_t VO(uint32_t s) const { return voff[s]; }
    LYC_HD bool D(uint32_t s) const { return (flags[s] & TF_DROPPED) != 0; }  // dropped by the merge filter
    LYC_HD TArr<float> len() const { return adv; }  // |P - prev| until the advancement chain turns it into adv, in place
};

LYC_HD uint32_t verb_point_count(uint8_t verb, uint32_t a) {
    switch (verb) {
        case VERB_BEGIN:
        case VERB_LINE:
        case VERB_CLOSE: return 1u + a;
        case VERB_QUAD: return 2u + a;
        case VERB_CUBIC: return 3u + a;
        default: return 0u;
    }
}
LYC_HD V2 seg_from_point(const V2* points, uint32_t pt, uint32_t a) {
    uint32_t back = 1u + a;
    return points[pt >= back ? pt - back : 0u];
}

// K1: number of slots of verb v (Wang counts of cubic_bezier.rs:590-603 / quadratic_bezier.rs:368-380,
// quads possibly split in two by flatten_quad, stroke.rs:2755-2779).
LYC_HD uint32_t seg_slot_count(uint8_t verb, const V2* points, uint32_t pt, uint32_t a, float tol, float* tsplit_out) {
    *tsplit_out = -1.0f;
    if (verb == VERB_CUBIC) {
        V2 from = seg_from_point(points, pt, a);
        return to_u32_or_1(cubic_wang(from, points[pt], points[pt + 1], points[pt + 2], tol));
    }
    if (verb == VERB_QUAD) {
        V2 from = seg_from_point(points, pt, a);
        V2 ctrl = points[pt], to = points[pt + 1];
        float ts;
        if (find_sharp_turn(from, ctrl, to, &ts)) {
            *tsplit_out = ts;
            V2 sp = quad_sample(from, ctrl, to, ts);  // QuadraticBezierSegment::split, quadratic_bezier.rs:208-223
            uint32_t n1 = to_u32_or_1(quad_wang(from, lerp(from, ctrl, ts), sp, tol));
            uint32_t n2 = to_u32_or_1(quad_wang(sp, lerp(ctrl, to, ts), to, tol));
            return n1 + n2;
        }
        return to_u32_or_1(quad_wang(from, ctrl, to, tol));
    }
    return 1u;  // Begin, Line, End, Close
}

struct SlotPoint {
    V2 pos, prev;
    float t;
    bool fs, has_prev;
};
// One flattened point of a quadratic (the k-th callback of for_each_flattened_with_t) and its predecessor.
LYC_HD void quad_point(V2 from, V2 ctrl, V2 to, float tol, uint32_t k, V2* pos, V2* prev, float* t_out) {
    float n = quad_wang(from, ctrl, to, tol);
    float step = 1.0f / n;
    uint32_t cnt = to_u32_or_1(n);
    QuadPoly poly = quad_poly(from, ctrl, to);
    float tp = 0.0f, tc = 0.0f;
    uint32_t adds = (k + 1u < cnt) ? k + 1u : k;  // the last point is the exact endpoint (t == 1.0)
    for (uint32_t i = 0; i < adds; ++i) {
        tp = tc;
        tc = tc + step;
    }
    if (k + 1u < cnt) {
        *pos = quad_poly_sample(poly, tc);
        *t_out = tc;
        *prev = (k == 0u) ? from : quad_poly_sample(poly, tp);
    } else {
        *pos = to;
        *t_out = 1.0f;
        *prev = (k == 0u) ? from : quad_poly_sample(poly, tc);
    }
}
LYC_HD void cubic_point(V2 from, V2 c1, V2 c2, V2 to, float tol, uint32_t k, V2* pos, V2* prev, float* t_out) {
    float n = cubic_wang(from, c1, c2, to, tol);
    float step = 1.0f / n;
    uint32_t cnt = to_u32_or_1(n);
    CubicPoly poly = cubic_poly(from, c1, c2, to);
    float tp = 0.0f, tc = 0.0f;
    uint32_t adds = (k + 1u < cnt) ? k + 1u : k;
    for (uint32_t i = 0; i < adds; ++i) {
        tp = tc;
        tc = tc + step;
    }
    if (k + 1u < cnt) {
        *pos = cubic_poly_sample(poly, tc);
        *t_out = tc;
        *prev = (k == 0u) ? from : cubic_poly_sample(poly, tp);
    } else {
        *pos = to;
        *t_out = 1.0f;
        *prev = (k == 0u) ? from : cubic_poly_sample(poly, tc);
    }
}

// K2: evaluate slot k of verb v.
LYC_HD SlotPoint eval_slot(uint8_t verb, const V2* points, uint32_t pt, uint32_t a, float tol, float tsplit, uint32_t k) {
    SlotPoint r;
    r.t = 1.0f;
    r.fs = false;
    r.has_prev = true;
    r.prev = v2(0.0f, 0.0f);
    switch (verb) {
        case VERB_BEGIN:
            r.pos = points[pt];
            r.has_prev = false;
            break;
        case VERB_LINE:
            r.pos = points[pt];
            r.prev = seg_from_point(points, pt, a);
            break;
        case VERB_CUBIC: {
            V2 from = seg_from_point(points, pt, a);
            cubic_point(from, points[pt], points[pt + 1], points[pt + 2], tol, k, &r.pos, &r.prev, &r.t);
            r.fs = r.t != 1.0f;
            break;
        }
        case VERB_QUAD: {
            V2 from = seg_from_point(points, pt, a);
            V2 ctrl = points[pt], to = points[pt + 1];
            if (tsplit >= 0.0f) {
                V2 sp = quad_sample(from, ctrl, to, tsplit);
                V2 bc = lerp(from, ctrl, tsplit);
                V2 ac = lerp(ctrl, to, tsplit);
                uint32_t n1 = to_u32_or_1(quad_wang(from, bc, sp, tol));
                float ts;
                if (k < n1) {
                    quad_point(from, bc, sp, tol, k, &r.pos, &r.prev, &ts);
                    r.fs = ts != 1.0f;
                    r.t = ts * tsplit;  // stroke.rs:2764
                } else {
                    quad_point(sp, ac, to, tol, k - n1, &r.pos, &r.prev, &ts);
                    r.fs = ts != 1.0f;
                    r.t = tsplit + ts * (1.0f - tsplit);  // stroke.rs:2770
                }
            } else {
                quad_point(from, ctrl, to, tol, k, &r.pos, &r.prev, &r.t);
                r.fs = r.t != 1.0f;
            }
            break;
        }
        default:  // End / Close: not a point
            r.pos = v2(0.0f, 0.0f);
            r.has_prev = false;
            break;
    }
    return r;
}

// K2 per slot: writes pos/t/segfs/len.  The merge filter of fixed_width_step_impl (stroke.rs:1530-1536)
// compares against the last KEPT point, which is sequential; here we only pre-classify: a pair further
// apart than sqrt(5)x the merge distance can never be merged whatever was dropped before it (triangle
// inequality with margin), so path_walk only has to look at the flagged pairs.
LYC_HD void point_gen_slot(const Views& w, const StrokeParams& p, uint32_t v, uint32_t s, uint32_t k) {
    uint8_t verb = w.verbs[v];
    if (verb == VERB_END || verb == VERB_CLOSE) {
        w.segfs[s] = v | SEG_END_BIT;
        w.t[s] = 1.0f;
        w.len[s] = 0.0f;
        return;  // pos[s] / adv[s] are written by path_walk (sub-path record)
    }
    SlotPoint r = eval_slot(verb, w.points, w.seg_pt[v], p.attr_stride, p.tolerance, w.seg_tsplit[v], k);
    w.pos[s] = r.pos;
    w.t[s] = r.t;
    w.segfs[s] = v | (r.fs ? SEG_FS_BIT : 0u) | (verb == VERB_BEGIN ? SEG_BEGIN_BIT : 0u);
    float l = 0.0f;
    if (r.has_prev) {
        V2 d = r.pos - r.prev;
        float d2 = sqlen(d);
        l = sqrtf(d2);
        if (!(d2 >= p.flag_thr)) l = -l;  // also flags NaN distances; -0.0f keeps the flag for zero length
    }
    w.len[s] = l;
}
LYC_HD bool len_flagged(float l) { return (l != l) || signbit(l); }

// Sub-path record stored in pos[end_slot] by path_walk.
LYC_HD uint32_t f2u(float f) {
    union {
        float f;
        uint32_t u;
    } c;
    c.f = f;
    return c.u;
}
LYC_HD float u2f(uint32_t u) {
    union {
        float f;
        uint32_t u;
    } c;
    c.u = u;
    return c.f;
}
static const uint32_t REC_MNEC = 4u, REC_CLOSE = 8u;

// K3: one path, sequential.  Restates the parts of the stroker that are loop-carried in f32:
//   * merge filter against the last kept point          (stroke.rs:1318-1320, 1530-1536)
//   * advancement chain adv += |edge|                     (stroke.rs:1546-1550, 1657-1661, 1759-1769, 2140)
//   * sub_path_start_advancement carry (open sub-paths only, stroke.rs:1359)
//   * may_need_empty_cap                                   (stroke.rs:1180, 1533)
//   * endpoint ids: sequential per event (stroke.rs:882-937) or point-array offsets (path.rs:1197-1263)
LYC_HD void path_walk(const Views& w, const StrokeParams& p, uint32_t vb, uint32_t ve) {
    const float nan = u2f(0x7FC00000u);
    float carry = 0.0f, adv = 0.0f;
    uint32_t evt = 0, begin_slot = 0, last_kept = 0, m = 0;
    bool mnec = false;
    const bool with_ids = p.num_attributes > 0;
    const uint32_t pt_base = vb < ve ? w.seg_pt[vb] : 0u;
    for (uint32_t v = vb; v < ve; ++v) {
        uint8_t verb = w.verbs[v];
        uint32_t s0 = w.seg_off[v], s1 = w.seg_off[v + 1];
        if (verb == VERB_BEGIN) {
            begin_slot = s0;
            last_kept = s0;
            m = 1;
            mnec = false;
            adv = carry;
            w.adv[s0] = adv;
            w.seg_id[v] = with_ids ? (w.seg_pt[v] - pt_base) : evt;
            evt++;
        } else if (verb == VERB_END || verb == VERB_CLOSE) {
            bool close = verb == VERB_CLOSE;
            if (close && m == 1) mnec = true;
            V2 rec;
            rec.x = u2f(begin_slot);
            rec.y = u2f((m > 3u ? 3u : m) | (mnec ? REC_MNEC : 0u) | (close ? REC_CLOSE : 0u));
            w.pos[s0] = rec;
            w.adv[s0] = nan;
            if (!(close && m > 2u) && m >= 2u) carry = adv;
            m = 0;
        } else {
            uint32_t npts = verb == VERB_LINE ? 1u : (verb == VERB_QUAD ? 2u : 3u);
            w.seg_id[v] = with_ids ? (w.seg_pt[v] + npts - 1u - pt_base) : evt;
            evt++;
            for (uint32_t s = s0; s < s1; ++s) {
                float l = w.len[s];
                if (len_flagged(l) || last_kept != s - 1u) {
                    V2 pl = w.pos[last_kept], ps = w.pos[s];
                    if (sqlen(pl - ps) < p.merge_thr) {
                        w.adv[s] = nan;
                        if (m == 1u) mnec = true;
                        continue;
                    }
                    l = length(ps - pl);
                }
                adv = adv + l;
                w.adv[s] = adv;
                last_kept = s;
                m++;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Join geometry (fixed line width)
// ---------------------------------------------------------------------------------------------------------
// The two sides are indexed [SIDE_POSITIVE] / [SIDE_NEGATIVE]; all indexing below is by compile-time constants
// (loops over sides are fully unrolled) so the struct lives in registers on the device.
struct Join {
    bool fast;
    bool fold[2];
    bool has_single[2];
    uint8_t line_join;
    V2 single[2], prev[2], next[2];
    V2 normal;  // fast path vertex normal
    uint32_t prev_v[2], next_v[2];
};
LYC_HD bool miter_limit_is_exceeded(V2 normal, float miter_limit) {  // stroke.rs:2320-2322
    return sqlen(normal) > miter_limit * miter_limit * 4.0f;
}
LYC_HD float circle_flattening_step(float radius, float tolerance) {  // stroke.rs:2749-2753
    tolerance = fminf(tolerance, radius);
    return 2.0f * libm_acos((radius - tolerance) / radius);
}
LYC_HD void get_clip_intersections(V2 previous_normal, V2 next_normal, V2 normal, float clip_distance, V2* i1,
                                   V2* i2) {  // stroke.rs:2279-2315
    LineD clip_line = line_d(normalize(normal) * clip_distance, tangent(normal));
    LineD prev_line = line_d(previous_normal, tangent(previous_normal));
    LineD next_line = line_d(next_normal, tangent(next_normal));
    if (!line_intersection(clip_line, prev_line, i1)) *i1 = normal;
    if (!line_intersection(clip_line, next_line, i2)) *i2 = normal;
}

// flattened_step's miter normal (stroke.rs:1751-1757)
LYC_HD V2 fast_join_normal(V2 pp, V2 pj, V2 pn) {
    V2 prev_edge = pj - pp;
    V2 next_edge = pn - pj;
    V2 prev_tangent = prev_edge / length(prev_edge);
    V2 next_tangent = next_edge / length(next_edge);
    return compute_normal(prev_tangent, next_tangent);
}
LYC_HD bool join_is_fast(V2 pp, V2 pj, V2 pn, bool join_fs) {  // stroke.rs:1568-1574
    return join_fs && dot(pj - pp, pn - pj) > 0.0f;
}

// fixed_width_step_impl's choice between flattened_step (stroke.rs:1743-1806) and
// compute_join_side_positions_fixed_width (stroke.rs:1643-1733) for the join at `pj`.
LYC_HD void compute_join_fw(Join& j, V2 pp, V2 pj, V2 pn, bool join_fs, const StrokeParams& o) {
    const float hw = o.half_width;
    V2 prev_edge = pj - pp;
    V2 next_edge = pn - pj;
    float prev_length = length(prev_edge);
    float next_length = length(next_edge);
    V2 prev_tangent = prev_edge / prev_length;
    V2 next_tangent = next_edge / next_length;
    V2 normal = compute_normal(prev_tangent, next_tangent);
    j.fold[0] = j.fold[1] = false;
    j.fast = join_fs && dot(prev_edge, next_edge) > 0.0f;  // stroke.rs:1568-1574
    j.normal = normal;
    if (j.fast) {
        j.line_join = JOIN_MITER;
        V2 p0 = pj + normal * hw;
        V2 p1 = pj - normal * hw;
        j.prev[SIDE_POSITIVE] = j.next[SIDE_POSITIVE] = j.single[SIDE_POSITIVE] = p0;
        j.prev[SIDE_NEGATIVE] = j.next[SIDE_NEGATIVE] = j.single[SIDE_NEGATIVE] = p1;
        j.has_single[0] = j.has_single[1] = true;
        return;
    }
    j.line_join = o.line_join;
    j.has_single[0] = j.has_single[1] = false;
    // front side = the outside of the turn (stroke.rs:1665-1669)
    const bool front_neg = cross(prev_tangent, next_tangent) >= 0.0f;
    V2 front_normal = front_neg ? -normal : normal;
    V2 extruded_normal = front_normal * hw;
    bool unclipped_miter =
        (j.line_join == JOIN_MITER || j.line_join == JOIN_MITER_CLIP) && !miter_limit_is_exceeded(front_normal, o.miter_limit);
    bool fold = false;
    bool angle_is_sharp = dot(next_tangent, prev_tangent) < 0.0f;
    if (!unclipped_miter && angle_is_sharp) {
        float d_next = dot(extruded_normal, -next_tangent) - next_length;
        float d_prev = dot(extruded_normal, prev_tangent) - prev_length;
        if (fminf(d_next, d_prev) > 0.0f || sqlen(normal) < 1e-5f) {
            j.fold[SIDE_NEGATIVE] = front_neg;
            j.fold[SIDE_POSITIVE] = !front_neg;
            fold = true;
        }
    }
    V2 n0 = v2(-prev_tangent.y, prev_tangent.x) * hw;
    V2 n1 = v2(-next_tangent.y, next_tangent.x) * hw;
    j.prev[SIDE_POSITIVE] = pj + n0;
    j.next[SIDE_POSITIVE] = pj + n1;
    j.prev[SIDE_NEGATIVE] = pj - n0;
    j.next[SIDE_NEGATIVE] = pj - n1;
    j.single[0] = j.single[1] = pj;
    if (!fold) {
        V2 mp = pj + normal * hw;
        V2 mn = pj - normal * hw;
        // the back (inner) side always collapses to one vertex; the front side only for an unclipped miter
        j.has_single[SIDE_POSITIVE] = front_neg || unclipped_miter;
        j.has_single[SIDE_NEGATIVE] = !front_neg || unclipped_miter;
        if (j.has_single[SIDE_POSITIVE]) j.single[SIDE_POSITIVE] = mp;
        if (j.has_single[SIDE_NEGATIVE]) j.single[SIDE_NEGATIVE] = mn;
        if (!unclipped_miter && j.line_join == JOIN_MITER_CLIP) {
            V2 c0 = (front_neg ? j.prev[SIDE_NEGATIVE] : j.prev[SIDE_POSITIVE]) - pj;
            V2 c1 = (front_neg ? j.next[SIDE_NEGATIVE] : j.next[SIDE_POSITIVE]) - pj;
            V2 pn_, nn_;
            get_clip_intersections(c0, c1, front_normal, o.miter_limit * hw, &pn_, &nn_);
            if (front_neg) {
                j.prev[SIDE_NEGATIVE] = pj + pn_;
                j.next[SIDE_NEGATIVE] = pj + nn_;
            } else {
                j.prev[SIDE_POSITIVE] = pj + pn_;
                j.next[SIDE_POSITIVE] = pj + nn_;
            }
        }
    }
}
LYC_NOINLINE void compute_join_fw_nl(Join& j, V2 pp, V2 pj, V2 pn, bool join_fs, const StrokeParams& o) {
    compute_join_fw(j, pp, pj, pn, join_fs, o);
}
// Vertex ids of a join whose first vertex is `base`, in add_stroke_vertex order:
// fast path Positive then Negative (stroke.rs:1791-1797); otherwise Negative side then Positive side,
// each one vertex (single) or two (prev, next) (stroke.rs:1595-1608, 2024-2055).
LYC_HD uint32_t join_assign_ids(Join& j, uint32_t base) {
    if (j.fast) {
        j.prev_v[SIDE_POSITIVE] = j.next_v[SIDE_POSITIVE] = base;
        j.prev_v[SIDE_NEGATIVE] = j.next_v[SIDE_NEGATIVE] = base + 1u;
        return 2u;
    }
    const uint32_t nneg = j.has_single[SIDE_NEGATIVE] ? 1u : 2u;
    const uint32_t npos = j.has_single[SIDE_POSITIVE] ? 1u : 2u;
    j.prev_v[SIDE_NEGATIVE] = base;
    j.next_v[SIDE_NEGATIVE] = base + nneg - 1u;
    j.prev_v[SIDE_POSITIVE] = base + nneg;
    j.next_v[SIDE_POSITIVE] = base + nneg + npos - 1u;
    return nneg + npos;
}

// Per-slot join layout code written by the count pass and read back by the emit pass (so that the edge
// triangles towards the previous join do not have to recompute that join): which vertices the join owns.
static const uint8_t JC_FAST = 1, JC_FOLD_POS = 2, JC_FOLD_NEG = 4, JC_SINGLE_POS = 8, JC_SINGLE_NEG = 16;
LYC_HD uint8_t join_encode(const Join& j) {
    return (uint8_t)((j.fast ? JC_FAST : 0) | (j.fold[SIDE_POSITIVE] ? JC_FOLD_POS : 0) | (j.fold[SIDE_NEGATIVE] ? JC_FOLD_NEG : 0) |
                     (j.has_single[SIDE_POSITIVE] ? JC_SINGLE_POS : 0) | (j.has_single[SIDE_NEGATIVE] ? JC_SINGLE_NEG : 0));
}
// The CTA-tile kernels keep a shorter code: all the NEXT slot's edge triangles need of this join are the two vertices the
// edge leaves from (add_edge_triangles' p0_neg / p0_pos with the fold remaps applied), as offsets 0..3 from the join's first
// vertex: bit 0 JC_FAST, bits 1-2 the Negative-side exit vertex, bits 3-4 the Positive-side one.  `j` carries ids relative to 0.
static const uint8_t JX_FAST = (uint8_t)(JC_FAST | (1u << 1) | (0u << 3));  // flattened_step: Positive = 0, Negative = 1
LYC_HD uint8_t join_exit_code(const Join& j) {
    const uint32_t en = j.fold[SIDE_POSITIVE] ? j.prev_v[SIDE_POSITIVE] : j.next_v[SIDE_NEGATIVE];
    const uint32_t ep = j.fold[SIDE_NEGATIVE] ? j.prev_v[SIDE_NEGATIVE] : j.next_v[SIDE_POSITIVE];
    return (uint8_t)((j.fast ? JC_FAST : 0) | ((en & 3u) << 1) | ((ep & 3u) << 3));
}
// add_edge_triangles (stroke.rs:1869-1904) from an exit code: for a fixed line width the vertices of two joins are all
// distinct, so both triangles are always there.
template <class Sink>
LYC_HD void add_edge_triangles_x(Sink& sink, uint32_t p0_base, uint8_t p0_code, const Join& b) {
    const uint32_t p0_neg = p0_base + ((p0_code >> 1) & 3u), p0_pos = p0_base + ((p0_code >> 3) & 3u);
    uint32_t p1_neg = b.prev_v[SIDE_NEGATIVE], p1_pos = b.prev_v[SIDE_POSITIVE];
    if (b.fold[SIDE_POSITIVE]) p1_neg = b.next_v[SIDE_POSITIVE];
    if (b.fold[SIDE_NEGATIVE]) p1_pos = b.next_v[SIDE_NEGATIVE];
    sink.add_triangle(p0_neg, p0_pos, p1_pos);
    sink.add_triangle(p0_neg, p1_pos, p1_neg);
}
LYC_HD void join_decode_ids(Join& j, uint8_t code, uint32_t base) {  // fills only what add_edge_triangles reads
    j.fast = (code & JC_FAST) != 0;
    j.fold[SIDE_POSITIVE] = (code & JC_FOLD_POS) != 0;
    j.fold[SIDE_NEGATIVE] = (code & JC_FOLD_NEG) != 0;
    j.has_single[SIDE_POSITIVE] = (code & JC_SINGLE_POS) != 0;
    j.has_single[SIDE_NEGATIVE] = (code & JC_SINGLE_NEG) != 0;
    join_assign_ids(j, base);
}

// ---------------------------------------------------------------------------------------------------------
// Sinks: the count pass and the emit pass run the SAME emission code against different sinks.
// ---------------------------------------------------------------------------------------------------------
struct VertexTemplate {  // StrokeVertexData minus normal/side (stroke.rs:2656-2665)
    V2 position_on_path;
    float half_width, advancement;
    uint32_t src;  // endpoint id
    float t;
};
struct CountSink {
    uint32_t nv, nt, base;
    LYC_HD uint32_t next_id() const { return base + nv; }
    LYC_HD uint32_t add_vertex(const VertexTemplate&, V2, bool) { return base + nv++; }
    LYC_HD void add_triangle(uint32_t, uint32_t, uint32_t) { nt++; }
    LYC_HD void arc(const VertexTemplate&, bool, float, float, uint32_t, uint32_t, uint32_t total) {
        nv += total;
        nt += total;
    }
    static const bool kCounting = true;
};
// One node of tessellate_arc's recursion (stroke.rs:2612-2653).  The recursion emits 2^k - 1 vertices in
// PRE-ORDER, each followed by its triangle (va, v, vb).  Node j re-derives its angle interval and corner ids by
// descending from the root with the same sequence of f32 midpoints the recursion computes, so any thread can
// produce any arc vertex independently.
struct ArcNode {
    float mid;
    uint32_t id, ida, idb;
};
LYC_HD ArcNode arc_node(float a0, float a1, uint32_t va, uint32_t vb, uint32_t total, uint32_t first_id, uint32_t j) {
    float lo = a0, hi = a1;
    uint32_t ida = va, idb = vb, base = 0u, size = total;
    ArcNode n;
    for (;;) {
        n.mid = (lo + hi) * 0.5f;
        n.id = first_id + base;
        if (j == base) break;
        uint32_t half = (size - 1u) >> 1;
        if (j <= base + half) {
            hi = n.mid;
            idb = n.id;
            base += 1u;
        } else {
            lo = n.mid;
            ida = n.id;
            base += 1u + half;
        }
        size = half;
    }
    n.ida = ida;
    n.idb = idb;
    return n;
}
LYC_HD V2 arc_normal(float mid) { return libm_cos_sin(mid); }  // vector(mid.cos(), mid.sin()), stroke.rs:2627
struct EmitSink {
    GpuVertex* vtx;  // first vertex owned by this slot (global memory)
    uint32_t* idx;   // first index owned by this slot
    uint32_t nv, nt, base;
    LYC_HD uint32_t next_id() const { return base + nv; }
    LYC_HD uint32_t add_vertex(const VertexTemplate& vt, V2 normal, bool negative) {
#ifdef __CUDA_ARCH__
        float4* dst = reinterpret_cast<float4*>(vtx + nv);
        dst[0] = make_float4(vt.position_on_path.x, vt.position_on_path.y, normal.x, normal.y);
        dst[1] = make_float4(vt.advancement, vt.half_width,
                             __uint_as_float((vt.src & 0x7FFFFFFFu) | (negative ? SIDE_NEG_BIT : 0u)), vt.t);
#else
        GpuVertex g;
        g.px = vt.position_on_path.x;
        g.py = vt.position_on_path.y;
        g.nx = normal.x;
        g.ny = normal.y;
        g.advancement = vt.advancement;
        g.half_width = vt.half_width;
        g.src = (vt.src & 0x7FFFFFFFu) | (negative ? SIDE_NEG_BIT : 0u);
        g.t = vt.t;
        vtx[nv] = g;
#endif
        return base + nv++;
    }
    LYC_HD void add_triangle(uint32_t a, uint32_t b, uint32_t c) {
        uint32_t* d = idx + 3u * (size_t)nt;
        d[0] = a;
        d[1] = b;
        d[2] = c;
        nt++;
    }
    LYC_HD void arc(const VertexTemplate& vt, bool negative, float a0, float a1, uint32_t va, uint32_t vb, uint32_t total);
    static const bool kCounting = false;
};

// All vertices and triangles of one arc, written at vtx[0..total) / idx[0..3*total) (one copy of this code per kernel).
LYC_NOINLINE void arc_emit_nl(GpuVertex* vtx, uint32_t* idx, uint32_t first_id, VertexTemplate vt, bool negative, float a0,
                              float a1, uint32_t va, uint32_t vb, uint32_t total) {
    EmitSink s;
    s.vtx = vtx;
    s.idx = idx;
    s.nv = s.nt = 0u;
    s.base = first_id;
    for (uint32_t j = 0; j < total; ++j) {
        ArcNode n = arc_node(a0, a1, va, vb, total, first_id, j);
        s.add_vertex(vt, arc_normal(n.mid), negative);
        s.add_triangle(n.ida, n.id, n.idb);
    }
}
LYC_HD void EmitSink::arc(const VertexTemplate& vt, bool negative, float a0, float a1, uint32_t va, uint32_t vb, uint32_t total) {
    arc_emit_nl(vtx + nv, idx + 3u * (size_t)nt, base + nv, vt, negative, a0, a1, va, vb, total);
    nv += total;
    nt += total;
}

// The sink of the tile kernels' fast-join pass: one 256-bit store per 32-byte vertex (SASS STG.E.ENL2.256) instead of two
// 128-bit ones; the vertex array must be 32-byte aligned (checked by the C entry points).  Used ONLY where the store is
// inlined into the kernel body: ptxas 12.9 miscompiles the same PTX inside out-of-line device functions (a 32-bit STG,
// seen in round 2; tools/check_sass.py looks for that pattern in the library that ships).
struct EmitSinkWide : EmitSink {
    LYC_HD uint32_t add_vertex(const VertexTemplate& vt, V2 normal, bool negative) {
#ifdef __CUDA_ARCH__
        const float srcbits = __uint_as_float((vt.src & 0x7FFFFFFFu) | (negative ? SIDE_NEG_BIT : 0u));
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(__cvta_generic_to_global(vtx + nv)),
                     "f"(vt.position_on_path.x), "f"(vt.position_on_path.y), "f"(normal.x), "f"(normal.y), "f"(vt.advancement),
                     "f"(vt.half_width), "f"(srcbits), "f"(vt.t)
                     : "memory");
        return base + nv++;
#else
        return EmitSink::add_vertex(vt, normal, negative);
#endif
    }
};

// The sink of the tile kernels' dense join pass: the same bytes as EmitSink, in fewer and fuller store transactions (the
// emit phases are bound by the SM's store path, which handles one 32-byte sector per transaction however little of it a
// store fills).  A vertex is ONE 256-bit store; the three indices of a triangle (12 bytes, so every other triangle starts
// off an 8-byte boundary) go out as 64-bit stores: the odd word waits in `pend` for the first index of the next triangle.
// flush() writes a word that is still waiting; it must run before anybody else continues at idx + 3 * nt (arc()) and when
// the slot is done.
struct EmitSinkPacked : EmitSink {
    uint32_t pend;
    bool has_pend;
#ifdef __CUDA_ARCH__
    __device__ __forceinline__ uint32_t add_vertex(const VertexTemplate& vt, V2 normal, bool negative) {
        const float srcbits = __uint_as_float((vt.src & 0x7FFFFFFFu) | (negative ? SIDE_NEG_BIT : 0u));
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(__cvta_generic_to_global(vtx + nv)),
                     "f"(vt.position_on_path.x), "f"(vt.position_on_path.y), "f"(normal.x), "f"(normal.y), "f"(vt.advancement),
                     "f"(vt.half_width), "f"(srcbits), "f"(vt.t)
                     : "memory");
        return base + nv++;
    }
    __device__ __forceinline__ void add_triangle(uint32_t a, uint32_t b, uint32_t c) {
        uint32_t* d = idx + 3u * (size_t)nt;
        if (has_pend) {  // the word in front of d is waiting, and it is 8-byte aligned
            *reinterpret_cast<uint2*>(d - 1) = make_uint2(pend, a);
            *reinterpret_cast<uint2*>(d + 1) = make_uint2(b, c);
            has_pend = false;
        } else if ((reinterpret_cast<uintptr_t>(d) & 7u) == 0u) {
            *reinterpret_cast<uint2*>(d) = make_uint2(a, b);
            pend = c;
            has_pend = true;
        } else {
            d[0] = a;
            *reinterpret_cast<uint2*>(d + 1) = make_uint2(b, c);
        }
        nt++;
    }
    __device__ __forceinline__ void flush() {
        if (has_pend) idx[3u * (size_t)nt - 1u] = pend;
        has_pend = false;
    }
    __device__ __forceinline__ void arc(const VertexTemplate& vt, bool negative, float a0, float a1, uint32_t va, uint32_t vb,
                                        uint32_t total);
#else
    LYC_HD void flush() {}
#endif
};
#ifdef __CUDA_ARCH__
// All vertices and triangles of one arc through the packed sink (one copy of this code per kernel).
__device__ __noinline__ void arc_emit_packed_nl(GpuVertex* vtx, uint32_t* idx, uint32_t first_id, VertexTemplate vt, bool negative,
                                                float a0, float a1, uint32_t va, uint32_t vb, uint32_t total) {
    EmitSinkPacked s;
    s.vtx = vtx;
    s.idx = idx;
    s.nv = s.nt = 0u;
    s.base = first_id;
    s.has_pend = false;
    s.pend = 0u;
    for (uint32_t j = 0; j < total; ++j) {
        ArcNode n = arc_node(a0, a1, va, vb, total, first_id, j);
        // (two 128-bit stores: in this out-of-line function ptxas 12.9 turned the 256-bit store into a 32-bit one in one
        // of the kernel instantiations -- see EmitSinkWide)
        s.EmitSink::add_vertex(vt, arc_normal(n.mid), negative);
        s.add_triangle(n.ida, n.id, n.idb);
    }
    s.flush();
}
__device__ __forceinline__ void EmitSinkPacked::arc(const VertexTemplate& vt, bool negative, float a0, float a1, uint32_t va,
                                                    uint32_t vb, uint32_t total) {
    flush();
    arc_emit_packed_nl(vtx + nv, idx + 3u * (size_t)nt, base + nv, vt, negative, a0, a1, va, vb, total);
    nv += total;
    nt += total;
}
#endif

// stroke.rs:1869-1904.  (ids of p0 after its join, ids of p1 before its join, with the fold remaps)
template <class Sink>
LYC_HD void add_edge_triangles(Sink& sink, const uint32_t p0_prev_v[2], const uint32_t p0_next_v[2], const bool p0_fold[2],
                               const uint32_t p1_prev_v[2], const uint32_t p1_next_v[2], const bool p1_fold[2]) {
    uint32_t p0_neg = p0_next_v[SIDE_NEGATIVE];
    uint32_t p0_pos = p0_next_v[SIDE_POSITIVE];
    uint32_t p1_neg = p1_prev_v[SIDE_NEGATIVE];
    uint32_t p1_pos = p1_prev_v[SIDE_POSITIVE];
    if (p0_fold[SIDE_POSITIVE]) p0_neg = p0_prev_v[SIDE_POSITIVE];
    if (p0_fold[SIDE_NEGATIVE]) p0_pos = p0_prev_v[SIDE_NEGATIVE];
    if (p1_fold[SIDE_POSITIVE]) p1_neg = p1_next_v[SIDE_POSITIVE];
    if (p1_fold[SIDE_NEGATIVE]) p1_pos = p1_next_v[SIDE_NEGATIVE];
    if (p0_neg == p1_pos) return;
    if (p0_neg != p0_pos && p0_pos != p1_pos) sink.add_triangle(p0_neg, p0_pos, p1_pos);
    if (p0_neg != p1_neg && p1_pos != p1_neg) sink.add_triangle(p0_neg, p1_pos, p1_neg);
}
template <class Sink>
LYC_HD void add_edge_triangles(Sink& sink, const Join& a, const Join& b) {
    add_edge_triangles(sink, a.prev_v, a.next_v, a.fold, b.prev_v, b.next_v, b.fold);
}

template <class Sink>
LYC_HD void emit_arc(Sink& sink, const VertexTemplate& vt, bool negative, float a0, float a1, uint32_t va, uint32_t vb,
                     uint32_t k) {
    if (k == 0u) return;
    if (k > MAX_ARC_SUBDIVISIONS) k = MAX_ARC_SUBDIVISIONS;
    sink.arc(vt, negative, a0, a1, va, vb, (1u << k) - 1u);
}
// `num_segments.log2().round() as u32` (stroke.rs:2004, 2490) for the integer-valued, non-negative float that
// ceil() produced.  For N < 2^16 the rounding is decided exactly in integers: round(log2 N) = f + [N^2 >= 2^(2f+1)]
// with f = floor(log2 N) (log2 N is never within float rounding of f + 0.5 for such N; checked exhaustively by
// tests/test_emulation_parity.py against the libm formulation).  Larger N, NaN and inf take the libm route.
LYC_HD uint32_t round_log2_segments(float num_segments) {
    if (num_segments >= 1.0f && num_segments < 65536.0f) {
        uint32_t n = (uint32_t)num_segments;
        uint32_t f = 0;
        while ((2u << f) <= n) ++f;  // f = floor(log2 n)
        return f + ((n * n >= (2u << (2u * f))) ? 1u : 0u);
    }
    return sat_u32(roundf(libm_log2(num_segments)));
}
LYC_HD uint32_t arc_subdivisions(float diff, float radius, const StrokeParams& o) {  // stroke.rs:2002-2004, 2488-2490
    float stp = radius == o.half_width ? o.arc_step : circle_flattening_step(radius, o.tolerance);
    float num_segments = ceilf(fabsf(diff) / stp);
    return round_log2_segments(num_segments);
}

// tessellate_round_join (stroke.rs:1968-2021).  `negative` is a run-time value: in a warp the front side of the joins
// is Positive for some lanes and Negative for the others, and one code path keeps them all busy (the side's fields are
// picked with selects, so the Join stays in registers).
template <class Sink>
LYC_HD void tessellate_round_join(Sink& sink, const VertexTemplate& vt, const Join& j, const StrokeParams& o, bool negative) {
    V2 center = vt.position_on_path;
    float radius = vt.half_width;
    V2 start_normal = (negative ? j.prev[SIDE_NEGATIVE] : j.prev[SIDE_POSITIVE]) - center;
    V2 end_normal = (negative ? j.next[SIDE_NEGATIVE] : j.next[SIDE_POSITIVE]) - center;
    uint32_t start_vertex = negative ? j.prev_v[SIDE_NEGATIVE] : j.prev_v[SIDE_POSITIVE];
    uint32_t end_vertex = negative ? j.next_v[SIDE_NEGATIVE] : j.next_v[SIDE_POSITIVE];
    float angle_sign = negative ? 1.0f : -1.0f;
    float start_angle = angle_from_x_axis(start_normal);
    float diff = angle_to(start_angle, angle_from_x_axis(end_normal));
    if (diff * angle_sign < 0.0f) diff = angle_sign * (2.0f * LYC_PI_F - fabsf(diff));
    float end_angle = start_angle + diff;
    if (negative) {
        float ta = start_angle;
        start_angle = end_angle;
        end_angle = ta;
        uint32_t tv = start_vertex;
        start_vertex = end_vertex;
        end_vertex = tv;
    }
    uint32_t k = arc_subdivisions(diff, radius, o);
    emit_arc(sink, vt, negative, start_angle, end_angle, start_vertex, end_vertex, k);
}

// tessellate_join (stroke.rs:1907-1965)
template <class Sink>
LYC_HD void tessellate_join(Sink& sink, const VertexTemplate& vt, const Join& j, const StrokeParams& o) {
    bool need0 = !j.has_single[SIDE_POSITIVE] && !j.fold[SIDE_NEGATIVE];
    bool need1 = !j.has_single[SIDE_NEGATIVE] && !j.fold[SIDE_POSITIVE];
    if (!j.fold[SIDE_POSITIVE] && !j.fold[SIDE_NEGATIVE]) {
        if (need0 && need1) {
            sink.add_triangle(j.prev_v[SIDE_POSITIVE], j.next_v[SIDE_POSITIVE], j.next_v[SIDE_NEGATIVE]);
            sink.add_triangle(j.prev_v[SIDE_POSITIVE], j.next_v[SIDE_NEGATIVE], j.prev_v[SIDE_NEGATIVE]);
        } else if (!need0 && need1) {
            sink.add_triangle(j.prev_v[SIDE_NEGATIVE], j.prev_v[SIDE_POSITIVE], j.next_v[SIDE_NEGATIVE]);
        } else if (need0 && !need1) {
            sink.add_triangle(j.prev_v[SIDE_NEGATIVE], j.prev_v[SIDE_POSITIVE], j.next_v[SIDE_POSITIVE]);
        }
    }
    if (j.line_join == JOIN_ROUND) {
        if (need0 != need1) {  // the usual case: only the front side
            tessellate_round_join(sink, vt, j, o, need1);
        } else if (need0) {  // Positive first, then Negative (stroke.rs:1954-1963)
            tessellate_round_join(sink, vt, j, o, false);
            tessellate_round_join(sink, vt, j, o, true);
        }
    }
}

// Vertices of a join, in the reference's order.  Fills j.prev_v / j.next_v.
template <class Sink>
LYC_HD void emit_join_vertices(Sink& sink, const VertexTemplate& vt, Join& j) {
    if (j.fast) {  // flattened_step, stroke.rs:1791-1803
        uint32_t pv = sink.add_vertex(vt, j.normal, false);
        uint32_t nv = sink.add_vertex(vt, -j.normal, true);
        j.prev_v[SIDE_POSITIVE] = j.next_v[SIDE_POSITIVE] = pv;
        j.prev_v[SIDE_NEGATIVE] = j.next_v[SIDE_NEGATIVE] = nv;
        return;
    }
    // add_join_base_vertices, stroke.rs:2024-2055: Negative side first, then Positive; a side is one vertex (single) or
    // two (prev, next).  Written as "vertex 0, 1, 2[, 3]" with the position picked by selects, so that the lanes of a
    // warp -- whose joins differ in which side is the single one -- all emit their i-th vertex together.
    const uint32_t base = sink.next_id();
    const bool sn = j.has_single[SIDE_NEGATIVE], sp = j.has_single[SIDE_POSITIVE];
    const V2 c = vt.position_on_path;
    const float hw = vt.half_width;
    const V2 n0 = sn ? j.single[SIDE_NEGATIVE] : j.prev[SIDE_NEGATIVE];
    const V2 p0 = sp ? j.single[SIDE_POSITIVE] : j.prev[SIDE_POSITIVE];
    sink.add_vertex(vt, (n0 - c) / hw, true);
    {
        const V2 q = sn ? p0 : j.next[SIDE_NEGATIVE];
        sink.add_vertex(vt, (q - c) / hw, !sn);
    }
    if (!sn || !sp) {
        const V2 q = sn ? j.next[SIDE_POSITIVE] : p0;
        sink.add_vertex(vt, (q - c) / hw, false);
    }
    if (!sn && !sp) sink.add_vertex(vt, (j.next[SIDE_POSITIVE] - c) / hw, false);
    join_assign_ids(j, base);
}

// tessellate_round_cap (stroke.rs:2459-2524)
template <class Sink>
LYC_HD void tessellate_round_cap(Sink& sink, VertexTemplate vt, V2 center, float radius, V2 start_normal,
                                 uint32_t start_vertex, uint32_t end_vertex, V2 edge_normal, bool is_start,
                                 const StrokeParams& o) {
    if (radius < o.tolerance) return;
    bool first_side_positive = is_start ^ (cross(edge_normal, start_normal) >= 0.0f);
    float start_angle = angle_from_x_axis(start_normal);
    float diff = angle_to(start_angle, angle_from_x_axis(edge_normal));
    float mid_angle = start_angle + diff;
    float end_angle = mid_angle + diff;
    uint32_t k = arc_subdivisions(diff, radius, o);
    vt.position_on_path = center;
    vt.half_width = radius;
    uint32_t mid_vertex = sink.add_vertex(vt, normalize(edge_normal), !first_side_positive);
    sink.add_triangle(start_vertex, mid_vertex, end_vertex);
    emit_arc(sink, vt, !first_side_positive, start_angle, mid_angle, start_vertex, mid_vertex, k);
    emit_arc(sink, vt, first_side_positive, mid_angle, end_angle, mid_vertex, end_vertex, k);
}

// Out-of-line entry points used by everything `end()` emits (one End slot per sub-path: few, heavy).
template <class Sink>
LYC_NOINLINE void tessellate_join_nl(Sink& sink, const VertexTemplate& vt, const Join& j, const StrokeParams& o) {
    tessellate_join(sink, vt, j, o);
}
template <class Sink>
LYC_NOINLINE void emit_join_vertices_nl(Sink& sink, const VertexTemplate& vt, Join& j) {
    emit_join_vertices(sink, vt, j);
}
template <class Sink>
LYC_NOINLINE void tessellate_round_cap_nl(Sink& sink, VertexTemplate vt, V2 center, float radius, V2 start_normal,
                                          uint32_t start_vertex, uint32_t end_vertex, V2 edge_normal, bool is_start,
                                          const StrokeParams& o) {
    tessellate_round_cap(sink, vt, center, radius, start_normal, start_vertex, end_vertex, edge_normal, is_start, o);
}

// ---------------------------------------------------------------------------------------------------------
// Slot neighbourhood helpers (W = Views or TileViews)
// ---------------------------------------------------------------------------------------------------------
template <class W>
LYC_HD bool slot_dropped(const W& w, uint32_t s) { return w.D(s); }
template <class W>
LYC_HD uint32_t prev_kept(const W& w, uint32_t s) {  // s is a kept non-Begin point: a kept predecessor exists
    uint32_t j = s - 1u;
    while (slot_dropped(w, j)) --j;
    return j;
}
template <class W>
LYC_HD uint32_t next_kept_or_end(const W& w, uint32_t s) {  // returns the End slot when s is the last kept point
    uint32_t j = s + 1u;
    while (!(w.F(j) & SEG_END_BIT) && slot_dropped(w, j)) ++j;
    return j;
}
template <class W>
LYC_HD VertexTemplate slot_vertex_template(const W& w, const StrokeParams& o, uint32_t s) {
    VertexTemplate vt;
    vt.position_on_path = w.P(s);
    vt.half_width = o.half_width;
    vt.advancement = w.A(s);
    vt.src = w.SRC(s);
    vt.t = w.T(s);
    return vt;
}
// Join at kept interior point `s` as computed while the stream was running (not by end()).
template <class W>
LYC_HD void stream_join(const W& w, const StrokeParams& o, uint32_t s, uint32_t pk, uint32_t nk, Join& j) {
    compute_join_fw(j, w.P(pk), w.P(s), w.P(nk), (w.F(s) & SEG_FS_BIT) != 0u, o);
}
template <class W>
LYC_HD void stream_join_nl(const W& w, const StrokeParams& o, uint32_t s, uint32_t pk, uint32_t nk, Join& j) {
    compute_join_fw_nl(j, w.P(pk), w.P(s), w.P(nk), (w.F(s) & SEG_FS_BIT) != 0u, o);
}
// Base vertex id of slot s. Counting pass: ids are never compared across slots in a way that matters for
// fixed width, any distinct bases do.
template <class Sink, class W>
LYC_HD uint32_t slot_vertex_base(const W& w, uint32_t s, uint32_t global_base) {
    return Sink::kCounting ? (0x40000000u + (s & 0xFFFFFu) * 16u) : (global_base + w.VO(s));
}

// ---------------------------------------------------------------------------------------------------------
// K4/K5 per slot
// ---------------------------------------------------------------------------------------------------------
template <class Sink, class W>
LYC_HD void stroke_end_slot(const W& w, const StrokeParams& o, uint32_t s, uint32_t global_base, Sink& sink);

// Edge triangles between the previous stream join (at kept slot pk, not a Begin) and join j at slot s.
// With a jcode array (emit pass) the previous join's vertex layout is decoded, otherwise it is recomputed.
template <class Sink, class W, class JC>
LYC_HD void edge_to_previous_join(const W& w, const StrokeParams& o, uint32_t s, uint32_t pk, uint32_t global_base,
                                  JC jcode, const Join& j, Sink& sink) {
    Join jp;
    if (jcode) {
        join_decode_ids(jp, jcode[pk], slot_vertex_base<Sink>(w, pk, global_base));
    } else {
        uint32_t pk2 = prev_kept(w, pk);
        stream_join(w, o, pk, pk2, s, jp);
        join_assign_ids(jp, slot_vertex_base<Sink>(w, pk, global_base));
    }
    add_edge_triangles(sink, jp, j);
}

// Slot classes.  The join at kept point i is owned by slot i; everything end() emits by the End slot.
enum { KIND_NONE = 0, KIND_JOIN = 1, KIND_END = 2 };
struct SlotClass {
    int kind;
    uint32_t sf, pk, nk;
};
template <class W>
LYC_HD SlotClass classify_slot(const W& w, uint32_t s) {
    SlotClass c;
    c.kind = KIND_NONE;
    c.pk = c.nk = 0u;
    c.sf = 0u;
    if (s >= w.n_slots) return c;
    c.sf = w.F(s);
    if (c.sf & SEG_END_BIT) {
        c.kind = KIND_END;
        return c;
    }
    if ((c.sf & SEG_BEGIN_BIT) || slot_dropped(w, s)) return c;
    c.nk = next_kept_or_end(w, s);
    if (w.F(c.nk) & SEG_END_BIT) return c;  // last kept point: owned by the End slot
    c.pk = prev_kept(w, s);
    c.kind = KIND_JOIN;
    return c;
}
// Count pass of a join that does not take the flattened_step fast path.
template <bool XC = false, class W>
LYC_HD void count_full_join(const W& w, const StrokeParams& o, uint32_t s, uint32_t pk, uint32_t nk, uint32_t* nv,
                            uint32_t* nt, uint8_t* code) {
    Join j;
    stream_join(w, o, s, pk, nk, j);
    CountSink sink;
    sink.nv = sink.nt = 0u;
    sink.base = 0u;
    VertexTemplate vt;
    vt.position_on_path = w.P(s);
    vt.half_width = o.half_width;
    vt.advancement = 0.0f;
    vt.src = 0u;
    vt.t = 1.0f;
    emit_join_vertices(sink, vt, j);
    if (!(w.F(pk) & SEG_BEGIN_BIT)) sink.nt += 2u;  // add_edge_triangles: always two for a fixed width
    tessellate_join(sink, vt, j, o);
    *nv = sink.nv;
    *nt = sink.nt;
    *code = XC ? join_exit_code(j) : join_encode(j);
}
// Emit pass of a fast join (flattened_step, stroke.rs:1743-1806).
template <class Sink, class W, class JC>
LYC_HD void emit_fast_join(const W& w, const StrokeParams& o, uint32_t s, uint32_t pk, uint32_t nk, uint32_t global_base,
                           JC jcode, Sink& sink) {
    const V2 normal = fast_join_normal(w.P(pk), w.P(s), w.P(nk));
    VertexTemplate vt = slot_vertex_template(w, o, s);
    Join j;
    j.fast = true;
    j.fold[0] = j.fold[1] = false;
    j.prev_v[SIDE_POSITIVE] = j.next_v[SIDE_POSITIVE] = sink.add_vertex(vt, normal, false);
    j.prev_v[SIDE_NEGATIVE] = j.next_v[SIDE_NEGATIVE] = sink.add_vertex(vt, -normal, true);
    if (!(w.F(pk) & SEG_BEGIN_BIT)) {
        Join jp;
        join_decode_ids(jp, jcode[pk], global_base + w.VO(pk));
        add_edge_triangles(sink, jp, j);
    }
}
// Emit pass of any other stream join.
template <bool XC = false, class Sink, class W, class JC>
LYC_HD void emit_full_join(const W& w, const StrokeParams& o, uint32_t s, uint32_t pk, uint32_t nk, uint32_t global_base,
                           JC jcode, Sink& sink) {
    Join j;
    stream_join(w, o, s, pk, nk, j);
    VertexTemplate vt = slot_vertex_template(w, o, s);
    emit_join_vertices(sink, vt, j);
    if (!(w.F(pk) & SEG_BEGIN_BIT)) {
        if (XC) add_edge_triangles_x(sink, global_base + w.VO(pk), jcode[pk], j);
        else edge_to_previous_join(w, o, s, pk, global_base, jcode, j, sink);
    }
    tessellate_join(sink, vt, j, o);
}

template <class Sink, class W>
LYC_HD void stroke_slot(const W& w, const StrokeParams& o, uint32_t s, uint32_t global_base, Sink& sink,
                        const uint8_t* jcode = nullptr, uint8_t* jcode_out = nullptr) {
    uint32_t sf = w.F(s);
    if (sf & SEG_END_BIT) {
        stroke_end_slot(w, o, s, global_base, sink);
        return;
    }
    if ((sf & SEG_BEGIN_BIT) || slot_dropped(w, s)) return;
    uint32_t nk = next_kept_or_end(w, s);
    if (w.F(nk) & SEG_END_BIT) return;  // last kept point: owned by the End slot
    uint32_t pk = prev_kept(w, s);
    Join j;
    stream_join(w, o, s, pk, nk, j);
    VertexTemplate vt = slot_vertex_template(w, o, s);
    if (jcode_out) jcode_out[s] = join_encode(j);
    emit_join_vertices(sink, vt, j);
    if (!(w.F(pk) & SEG_BEGIN_BIT))  // count > 2: edge between the previous join and this one (stroke.rs:1611-1613)
        edge_to_previous_join(w, o, s, pk, global_base, jcode, j, sink);
    tessellate_join(sink, vt, j, o);
}

// Side points of the first point of a sub-path, computed when the second kept point arrives
// (fixed_width_step_impl count == 1, stroke.rs:1538-1556).
LYC_HD void first_point_side_next(V2 p0, V2 p1, float hw, V2 next[2]) {
    V2 edge = p1 - p0;
    float len = length(edge);
    V2 tg = edge / len;
    V2 n = v2(-tg.y, tg.x) * hw;
    next[SIDE_POSITIVE] = p0 + n;
    next[SIDE_NEGATIVE] = p0 - n;
}

// tessellate_empty_square_cap / tessellate_empty_round_cap (stroke.rs:2526-2609) via tessellate_empty_cap (:1282-1316)
template <class Sink>
LYC_HD void tessellate_empty_cap(Sink& sink, VertexTemplate vt, const StrokeParams& o) {
    V2 c = vt.position_on_path;
    if (o.start_cap == CAP_SQUARE) {
        uint32_t a = sink.add_vertex(vt, v2(1.0f, 1.0f), true);
        uint32_t b = sink.add_vertex(vt, v2(1.0f, -1.0f), false);
        uint32_t cc = sink.add_vertex(vt, v2(-1.0f, -1.0f), false);
        uint32_t d = sink.add_vertex(vt, v2(-1.0f, 1.0f), true);
        sink.add_triangle(a, b, cc);
        sink.add_triangle(a, cc, d);
    } else if (o.start_cap == CAP_ROUND) {
        float radius = vt.half_width;
        uint32_t left_id = sink.add_vertex(vt, v2(-1.0f, 0.0f), false);
        uint32_t right_id = sink.add_vertex(vt, v2(1.0f, 0.0f), true);
        tessellate_round_cap_nl(sink, vt, c, radius, v2(-1.0f, 0.0f), left_id, right_id, v2(0.0f, 1.0f), true, o);
        tessellate_round_cap_nl(sink, vt, c, radius, v2(1.0f, 0.0f), right_id, left_id, v2(0.0f, -1.0f), false, o);
    }
}

// One end of an open sub-path: the side position clipped against the cap line in f64
// (tessellate_last_edge / tessellate_first_edge, stroke.rs:2160-2180, 2235-2255).
LYC_NOINLINE V2 clip_cap_side(uint8_t cap, V2 end_point, V2 other_point, V2 side_position, V2 towards, float hw) {
    if (cap == CAP_SQUARE || cap == CAP_BUTT) {
        float clip = cap == CAP_SQUARE ? hw : 0.0f;
        V2 normal = normalize(end_point - other_point);
        LineD clip_line = line_d(end_point + normal * clip, tangent(normal));
        LineD side_line = line_d(side_position, side_position - towards);
        V2 inter;
        if (line_intersection(clip_line, side_line, &inter)) return inter;
    }
    return side_position;
}

// Everything `end()` emits for one sub-path (stroke.rs:1179-1193): end_with_caps (:1322-1378) or close (:1206-1279).
template <class Sink, class W>
LYC_HD void stroke_end_slot(const W& w, const StrokeParams& o, uint32_t s, uint32_t global_base, Sink& sink) {
    const float hw = o.half_width;
    V2 rec = w.P(s);
    uint32_t b0 = f2u(rec.x);
    uint32_t flags = f2u(rec.y);
    uint32_t m = flags & 3u;
    bool mnec = (flags & REC_MNEC) != 0u;
    bool close = (flags & REC_CLOSE) != 0u;
    const bool no_fold[2] = {false, false};

    if (m == 1u) {
        if (mnec) tessellate_empty_cap(sink, slot_vertex_template(w, o, b0), o);
        return;
    }
    if (m < 2u) return;

    uint32_t s_p1 = next_kept_or_end(w, b0);  // second kept point
    uint32_t s_l1 = s - 1u;                   // last kept point
    while (slot_dropped(w, s_l1)) --s_l1;
    uint32_t s_l2 = prev_kept(w, s_l1);
    V2 P0 = w.P(b0), P1 = w.P(s_p1);

    if (!(close && m > 2u)) {
        // ---- end_with_caps ---------------------------------------------------------------------------------
        V2 pl2 = w.P(s_l2), pl1 = w.P(s_l1);
        // p0 = second to last point: its `next` side points, vertex ids and fold flags
        Join j0;
        if (m == 2u) {
            first_point_side_next(P0, P1, hw, j0.next);
            j0.fold[0] = j0.fold[1] = false;
            j0.prev_v[0] = j0.prev_v[1] = j0.next_v[0] = j0.next_v[1] = 0xFFFFFFFFu;
        } else {
            uint32_t s_l3 = prev_kept(w, s_l2);
            stream_join_nl(w, o, s_l2, s_l3, s_l1, j0);
            join_assign_ids(j0, slot_vertex_base<Sink>(w, s_l2, global_base));
        }
        // p1 = last point: side points from the edge direction (stroke.rs:1338-1346)
        V2 p1_prev[2];
        {
            V2 tg = normalize(pl1 - pl2);
            V2 n = v2(-tg.y, tg.x) * hw;
            p1_prev[SIDE_POSITIVE] = pl1 + n;
            p1_prev[SIDE_NEGATIVE] = pl1 - n;
        }
        // tessellate_last_edge (stroke.rs:2130-2206)
        VertexTemplate vt = slot_vertex_template(w, o, s_l1);  // advancement = chain value = p0.adv + |v|
        V2 v = pl1 - pl2;
        uint32_t p1_prev_v[2];
        p1_prev[SIDE_POSITIVE] = clip_cap_side(o.end_cap, pl1, pl2, p1_prev[SIDE_POSITIVE], j0.next[SIDE_POSITIVE], hw);
        p1_prev_v[SIDE_POSITIVE] = sink.add_vertex(vt, (p1_prev[SIDE_POSITIVE] - pl1) / hw, false);
        p1_prev[SIDE_NEGATIVE] = clip_cap_side(o.end_cap, pl1, pl2, p1_prev[SIDE_NEGATIVE], j0.next[SIDE_NEGATIVE], hw);
        p1_prev_v[SIDE_NEGATIVE] = sink.add_vertex(vt, (p1_prev[SIDE_NEGATIVE] - pl1) / hw, true);
        if (m > 2u) add_edge_triangles(sink, j0.prev_v, j0.next_v, j0.fold, p1_prev_v, p1_prev_v, no_fold);
        if (o.end_cap == CAP_ROUND)
            tessellate_round_cap_nl(sink, vt, pl1, hw, p1_prev[SIDE_POSITIVE] - pl1, p1_prev_v[SIDE_POSITIVE],
                                 p1_prev_v[SIDE_NEGATIVE], v, false, o);
        // tessellate_first_edge (stroke.rs:2208-2276): first = P0, second = P1 (firsts[1] when count > 2)
        V2 first_next[2];
        first_point_side_next(P0, P1, hw, first_next);
        V2 second_prev[2];
        uint32_t second_prev_v[2], second_next_v[2];
        bool second_fold[2] = {false, false};
        if (m == 2u) {
            second_prev[0] = p1_prev[0];
            second_prev[1] = p1_prev[1];
            second_prev_v[0] = second_next_v[0] = p1_prev_v[0];
            second_prev_v[1] = second_next_v[1] = p1_prev_v[1];
        } else {
            uint32_t s_p2 = next_kept_or_end(w, s_p1);
            Join j1;
            stream_join_nl(w, o, s_p1, b0, s_p2, j1);
            join_assign_ids(j1, slot_vertex_base<Sink>(w, s_p1, global_base));
            second_prev[0] = j1.prev[0];
            second_prev[1] = j1.prev[1];
            second_prev_v[0] = j1.prev_v[0];
            second_prev_v[1] = j1.prev_v[1];
            second_next_v[0] = j1.next_v[0];
            second_next_v[1] = j1.next_v[1];
            second_fold[0] = j1.fold[0];
            second_fold[1] = j1.fold[1];
        }
        VertexTemplate vf = slot_vertex_template(w, o, b0);
        uint32_t first_next_v[2];
        {
            V2 sp = clip_cap_side(o.start_cap, P0, P1, first_next[SIDE_POSITIVE], second_prev[SIDE_POSITIVE], hw);
            first_next_v[SIDE_POSITIVE] = sink.add_vertex(vf, (sp - P0) / hw, false);
            V2 sn = clip_cap_side(o.start_cap, P0, P1, first_next[SIDE_NEGATIVE], second_prev[SIDE_NEGATIVE], hw);
            first_next_v[SIDE_NEGATIVE] = sink.add_vertex(vf, (sn - P0) / hw, true);
        }
        {
            const uint32_t unset[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
            add_edge_triangles(sink, unset, first_next_v, no_fold, second_prev_v, second_next_v, second_fold);
        }
        if (o.start_cap == CAP_ROUND)
            tessellate_round_cap_nl(sink, vf, P0, hw, first_next[SIDE_NEGATIVE] - P0, first_next_v[SIDE_NEGATIVE],
                                 first_next_v[SIDE_POSITIVE], P0 - P1, true, o);
        return;
    }

    // ---- close (count > 2) -------------------------------------------------------------------------------------
    // firsts[0] = P0 (advancement reset to NaN), firsts[1] = P1 with the state of its stream join.
    uint32_t s_p2 = next_kept_or_end(w, s_p1);
    Join j1;  // firsts[1]
    stream_join_nl(w, o, s_p1, b0, s_p2, j1);
    join_assign_ids(j1, slot_vertex_base<Sink>(w, s_p1, global_base));
    const float start_advancement = w.A(b0);
    V2 A = w.P(s_l2), B = w.P(s_l1);
    uint32_t s_l3 = prev_kept(w, s_l2);  // exists: m > 2
    Join ja;                             // join at A = second to last kept point (already emitted by its slot)
    stream_join_nl(w, o, s_l2, s_l3, s_l1, ja);
    join_assign_ids(ja, slot_vertex_base<Sink>(w, s_l2, global_base));
    const bool b_fs = (w.F(s_l1) & SEG_FS_BIT) != 0u;

    bool segment_added = !(sqlen(B - P0) < o.merge_thr);  // step(firsts[0]) merge test (stroke.rs:1531)
    Join jc;       // the join the two extra vertices are re-created on ("p0" at stroke.rs:1248)
    VertexTemplate vc;
    if (segment_added) {
        // step(P0'): join at B between A and P0
        Join jb;
        compute_join_fw_nl(jb, A, B, P0, b_fs, o);
        VertexTemplate vb = slot_vertex_template(w, o, s_l1);
        emit_join_vertices_nl(sink, vb, jb);
        add_edge_triangles(sink, ja, jb);
        tessellate_join_nl(sink, vb, jb, o);
        // step(P1'): join at P0' between B and P1; P0' is never a flattening step
        compute_join_fw_nl(jc, B, P0, P1, false, o);
        vc = slot_vertex_template(w, o, b0);
        vc.advancement = vb.advancement + length(P0 - B);
        emit_join_vertices_nl(sink, vc, jc);
        add_edge_triangles(sink, jb, jc);
        tessellate_join_nl(sink, vc, jc, o);
    } else {
        // The closing point merged into B: B takes P0's position (stroke.rs:1228-1238), then step(P1').
        compute_join_fw_nl(jc, A, P0, P1, b_fs, o);
        vc = slot_vertex_template(w, o, s_l1);
        vc.position_on_path = P0;
        // B's advancement was already fixed iff the join at A took the fast path (flattened_step sets
        // next.advancement, stroke.rs:1765-1769); otherwise it is computed now from the moved position.
        if (!ja.fast) vc.advancement = w.A(s_l2) + length(P0 - A);
        emit_join_vertices_nl(sink, vc, jc);
        add_edge_triangles(sink, ja, jc);
        tessellate_join_nl(sink, vc, jc, o);
    }
    // Re-create the two vertices towards the second endpoint with the sub-path start advancement (stroke.rs:1248-1275)
    vc.advancement = start_advancement;
    {
        V2 rp = jc.has_single[SIDE_POSITIVE] ? jc.single[SIDE_POSITIVE] : jc.next[SIDE_POSITIVE];
        jc.next_v[SIDE_POSITIVE] = sink.add_vertex(vc, (rp - vc.position_on_path) / vc.half_width, false);
        V2 rn = jc.has_single[SIDE_NEGATIVE] ? jc.single[SIDE_NEGATIVE] : jc.next[SIDE_NEGATIVE];
        jc.next_v[SIDE_NEGATIVE] = sink.add_vertex(vc, (rn - vc.position_on_path) / vc.half_width, true);
    }
    add_edge_triangles(sink, jc, j1);
}

}  // namespace lyc
