// stroke_wtile.cuh -- per-item logic of the WARP-TILE stroke pipeline (host+device).
//
// The warp-tile kernel (wtile_kernels.cuh) gives every WARP its own tile of whole paths in shared memory and runs all
// phases warp-synchronously: no block barrier, no shared counters, sixteen tiles in flight per SM instead of two.  A tile
// keeps 20 bytes per slot; what the CTA-tile kernel stored per slot and this one does not is re-derived on demand:
//
//   slot -> verb     one bit per slot marks the first slot of a verb (`head`), `vword` counts the marks before each
//                    32-slot word: verb_of(s) = vword[s / 32] + popc(head[s / 32] up to s) - 1
//   t, EndpointId    per VERB (`vstep` = 1 / Wang count, `vsrc`): t of slot k replays k + 1 additions, exactly the chain
//                    for_each_flattened_with_t runs (cubic_bezier.rs:605-611, quadratic_bezier.rs:382-388)
//   output offsets   16 + 16 bits per slot, relative to the vertex / triangle base of the slot's 32-slot chunk
//
// Everything `end()` emits (stroke.rs:1179-1193) leaves the tile kernel as a 128-byte EndRecord and is tessellated by a
// dense thread-per-record kernel (k_stroke_ends): one End slot per sub-path would otherwise run on ONE lane of its warp.
// Like stroke_core.cuh / stroke_tile.cuh this file is LYC_HD so that tests/emul runs the very same code on the CPU.
#pragma once
#include "stroke_tile.cuh"

namespace lyc {

// Plan of a quadratic that flatten_quad splits at a sharp turn (stroke.rs:2755-2779), written by the pre-pass
// (k_path_slots) in path order: path_split_begin[path] + ordinal of the split quadratic within its path.
struct SplitRec {
    float tsplit, n_a, n_b, _pad;
};
static const uint8_t VK_SPLIT = 0x80;  // per-verb kind byte: the verb is a split quadratic (low bits: the verb code)

// flags of a warp-tile slot: TF_FS / TF_END / TF_BEGIN / TF_CLOSE as in TileViews, and for T():
//   TF_EVAL            sample of a curve's polynomial form (t = accumulated parameter), TF_LAST: the exact endpoint (t = 1)
//   TF_FIRST, no EVAL  slot of a split quadratic (t from the SplitRec)
struct WTileViews {
    V2* pos;
    float* adv;      // advancement; NaN = dropped by the merge filter
    uint32_t* off;   // low 16 bits: vertex offset, high 16 bits: triangle offset, relative to the chunk bases
                     // (per-slot COUNTS in the same packing between the count and the scan phase; `len` before that)
    uint8_t* flags;  // TF_*
    uint8_t* jcode;
    uint32_t* cbase_v;  // per 32-slot chunk: vertices / triangles of the tile before the chunk
    uint32_t* cbase_t;
    const uint32_t* head;   // 1 bit per slot: first slot of its verb
    const uint16_t* vword;  // per 32-slot word: verbs starting before the word
    const uint16_t* vs0;    // per verb (tile-relative index): first slot; vs0[n_verbs] = n_slots
    const float* vstep;     // per verb: 1 / Wang count; split quadratic: bits = index of its SplitRec
    const uint32_t* vsrc;   // per verb: EndpointId
    const uint8_t* vkind;   // per verb: verb code | VK_SPLIT
    const SplitRec* split_rec;
    uint32_t n_slots;
    LYC_HD V2 P(uint32_t s) const { return pos[s]; }
    LYC_HD float A(uint32_t s) const { return adv[s]; }
    LYC_HD uint32_t F(uint32_t s) const { return (uint32_t)flags[s] << 24; }
    LYC_HD float* len() const { return reinterpret_cast<float*>(off); }
    LYC_HD uint32_t verb_of(uint32_t s) const {
        const uint32_t below = head[s >> 5] & (0xFFFFFFFFu >> (31u - (s & 31u)));
#ifdef __CUDA_ARCH__
        return (uint32_t)vword[s >> 5] + (uint32_t)__popc(below) - 1u;
#else
        return (uint32_t)vword[s >> 5] + (uint32_t)__builtin_popcount(below) - 1u;
#endif
    }
    LYC_HD uint32_t SRC(uint32_t s) const { return vsrc[verb_of(s)]; }
    LYC_HD uint32_t VO(uint32_t s) const { return cbase_v[s >> 5] + (off[s] & 0xFFFFu); }
    LYC_HD bool D(uint32_t s) const {
        const float a = A(s);
        return a != a;
    }
    LYC_HD uint32_t TO(uint32_t s) const { return cbase_t[s >> 5] + (off[s] >> 16); }
    LYC_HD float T(uint32_t s) const;
};

// t of slot j of a split quadratic, mapped back onto the whole curve (stroke.rs:2764 / 2770)
LYC_NOINLINE float wt_split_slot_t(SplitRec r, uint32_t j) {
    const uint32_t cnt_a = to_u32_or_1(r.n_a), cnt_b = to_u32_or_1(r.n_b);
    const bool hb = j >= cnt_a;
    const uint32_t k = hb ? j - cnt_a : j, cnt = hb ? cnt_b : cnt_a;
    const float ts = k + 1u < cnt ? accumulate_t(1.0f / (hb ? r.n_b : r.n_a), k + 1u) : 1.0f;
    return hb ? r.tsplit + ts * (1.0f - r.tsplit) : ts * r.tsplit;
}
LYC_HD float WTileViews::T(uint32_t s) const {
    const uint8_t fl = flags[s];
    if (fl & TF_EVAL) {
        if (fl & TF_LAST) return 1.0f;
        const uint32_t v = verb_of(s);
        return accumulate_t(vstep[v], s - (uint32_t)vs0[v] + 1u);
    }
    if ((fl & TF_FIRST) && !(fl & (TF_BEGIN | TF_END))) {
        const uint32_t v = verb_of(s);
        return wt_split_slot_t(split_rec[f2u(vstep[v])], s - (uint32_t)vs0[v]);
    }
    return 1.0f;  // Begin, line endpoints (End slots are never asked)
}

// Slot j of a split quadratic: position, the previous fed point, flags, and the "maybe mergeable" length.  Same
// arithmetic as tile_gen_split_quad (stroke_tile.cuh) with the plan read from the pre-pass instead of recomputed.
template <class W>
LYC_HD float wt_gen_split_slot(const W& w, const StrokeParams& p, const V2* points, uint32_t pt, uint32_t s0, SplitRec r,
                               uint32_t j) {
    const V2 from = seg_from_point(points, pt, p.attr_stride);
    const V2 ctrl = points[pt], to = points[pt + 1];
    const float tsplit = r.tsplit;
    const V2 sp = quad_sample(from, ctrl, to, tsplit);  // QuadraticBezierSegment::split, quadratic_bezier.rs:208-223
    const uint32_t cnt_a = to_u32_or_1(r.n_a), cnt_b = to_u32_or_1(r.n_b);
    const bool hb = j >= cnt_a;
    const uint32_t k = hb ? j - cnt_a : j, cnt = hb ? cnt_b : cnt_a;
    const float step = 1.0f / (hb ? r.n_b : r.n_a);
    const QuadPoly poly = hb ? quad_poly(sp, lerp(ctrl, to, tsplit), to) : quad_poly(from, lerp(from, ctrl, tsplit), sp);
    float t = 0.0f;
    for (uint32_t i = 0; i < k; ++i) t = t + step;
    const V2 prev = k == 0u ? (hb ? sp : from) : quad_poly_sample(poly, t);
    V2 pos;
    float ts;
    if (k + 1u < cnt) {
        ts = t + step;
        pos = quad_poly_sample(poly, ts);
    } else {
        ts = 1.0f;
        pos = hb ? to : sp;
    }
    w.pos[s0 + j] = pos;
    w.flags[s0 + j] = (uint8_t)((ts != 1.0f ? TF_FS : 0) | TF_FIRST);
    w.len()[s0 + j] = tile_len_flagged(p, pos, prev);
    return hb ? tsplit + ts * (1.0f - tsplit) : ts * tsplit;  // t on the whole curve (stroke.rs:2764 / 2770)
}
// The same for the CTA-tile kernel, which keeps t and the EndpointId per slot.
LYC_HD void tile_gen_split_slot(const TileViews& w, const StrokeParams& p, const V2* points, uint32_t pt, uint32_t s0, uint32_t src,
                                SplitRec r, uint32_t j) {
    const float t = wt_gen_split_slot(w, p, points, pt, s0, r, j);
    w.t[s0 + j] = t;
    w.src[s0 + j] = src;
    // CTA tiles keep the "maybe mergeable" flag in the flag byte and the length unsigned (tile_resolve_run)
    const float l = w.len()[s0 + j];
    w.flags[s0 + j] = (uint8_t)((w.flags[s0 + j] & TF_FS) | (len_flagged(l) ? TF_FLAGGED : 0));
    w.len()[s0 + j] = fabsf(l);
}

// One ordinary slot: slot k of `cnt` of a verb of kind `kind` whose first stored point is `pt`.  Writes pos and flags;
// returns through *from_out the point the verb starts from (the previous fed point of the verb's first slot).  The
// caller computes the edge length once the previous slot's position is visible (it is pos[s - 1] for k > 0).
template <class W>
LYC_HD void wt_gen_slot(const W& w, const StrokeParams& p, const V2* points, uint8_t kind, uint32_t pt, uint32_t k, uint32_t cnt,
                        float step, uint32_t s, V2* from_out) {
    if (kind == VERB_END || kind == VERB_CLOSE) {
        w.flags[s] = (uint8_t)(TF_END | (kind == VERB_CLOSE ? TF_CLOSE : 0));
        *from_out = v2(0.0f, 0.0f);
        return;  // pos / adv: the sub-path record, written by the walk
    }
    const bool curve = kind == VERB_QUAD || kind == VERB_CUBIC, cubic = kind == VERB_CUBIC;
    const uint32_t npts = cubic ? 3u : (curve ? 2u : 1u);
    const V2 to = points[pt + npts - 1u];
    const V2 from = kind == VERB_BEGIN ? to : seg_from_point(points, pt, p.attr_stride);
    *from_out = from;
    const bool last = k + 1u >= cnt;
    V2 pos = to;
    uint8_t fl = kind == VERB_BEGIN ? TF_BEGIN : 0;
    if (curve) {
        float t = 1.0f;
        if (!last) {
            t = accumulate_t(step, k + 1u);
            const V2 c1 = points[pt];
            V2 a1, a2, a3 = v2(0.0f, 0.0f);
            if (cubic) {
                const CubicPoly q = cubic_poly(from, c1, points[pt + 1], to);
                a1 = q.a1;
                a2 = q.a2;
                a3 = q.a3;
            } else {
                const QuadPoly q = quad_poly(from, c1, to);
                a1 = q.a1;
                a2 = q.a2;
            }
            // Horner as the reference has it (cubic_bezier.rs:1351-1365 / quadratic_bezier.rs:927-939)
            float u = t;
            V2 v = from + a1 * u;
            u = u * t;
            v = v + a2 * u;
            if (cubic) {
                u = u * t;
                v = v + a3 * u;
            }
            pos = v;
        }
        fl = (uint8_t)(TF_EVAL | (cubic ? TF_CUBIC : 0) | (t != 1.0f ? TF_FS : 0) | (last ? TF_LAST : 0));
    }
    w.pos[s] = pos;
    w.flags[s] = fl;
}

// ---- End slots -----------------------------------------------------------------------------------------------------
// Vertices (= triangles) a round cap adds: tessellate_round_cap, stroke.rs:2459-2524.
LYC_HD uint32_t round_cap_count(const StrokeParams& o, float radius, V2 start_normal, V2 edge_normal) {
    if (radius < o.tolerance) return 0u;
    const float start_angle = angle_from_x_axis(start_normal);
    const float diff = angle_to(start_angle, angle_from_x_axis(edge_normal));
    uint32_t k = arc_subdivisions(diff, radius, o);
    if (k > MAX_ARC_SUBDIVISIONS) k = MAX_ARC_SUBDIVISIONS;
    const uint32_t arc = k == 0u ? 0u : (1u << k) - 1u;
    return 1u + 2u * arc;
}
template <class W>
LYC_NOINLINE void end_slot_count_generic(const W& w, const StrokeParams& o, uint32_t s, uint32_t* nv, uint32_t* nt) {
    CountSink sink;
    sink.nv = sink.nt = 0u;
    sink.base = 0u;
    stroke_end_slot(w, o, s, 0u, sink);
    *nv = sink.nv;
    *nt = sink.nt;
}
// What stroke_end_slot emits for End slot s, without running it for the common case -- an OPEN sub-path with at least
// two kept points (end_with_caps, stroke.rs:1322-1378): two vertices per end, two edge triangles towards each
// neighbouring join, and the round caps, whose size depends on one angle each.  Everything else (closed sub-paths,
// empty caps) runs the real routine against a counting sink.
template <class W>
LYC_HD void end_slot_count(const W& w, const StrokeParams& o, uint32_t s, uint32_t* nv, uint32_t* nt) {
    const V2 rec = w.P(s);
    const uint32_t b0 = f2u(rec.x), fl = f2u(rec.y), m = fl & 3u;
    const bool mnec = (fl & REC_MNEC) != 0u, close = (fl & REC_CLOSE) != 0u;
    *nv = *nt = 0u;
    if (m < 2u) {
        if (m == 1u && mnec) end_slot_count_generic(w, o, s, nv, nt);
        return;
    }
    if (close && m > 2u) {
        end_slot_count_generic(w, o, s, nv, nt);
        return;
    }
    const float hw = o.half_width;
    uint32_t v = 4u, t = (m > 2u ? 2u : 0u) + 2u;
    if (o.end_cap == CAP_ROUND || o.start_cap == CAP_ROUND) {
        const uint32_t s_p1 = next_kept_or_end(w, b0);
        uint32_t s_l1 = s - 1u;
        while (slot_dropped(w, s_l1)) --s_l1;
        const uint32_t s_l2 = prev_kept(w, s_l1);
        if (o.end_cap == CAP_ROUND) {
            const V2 pl2 = w.P(s_l2), pl1 = w.P(s_l1);
            const V2 tg = normalize(pl1 - pl2);
            const V2 n = v2(-tg.y, tg.x) * hw;
            const uint32_t c = round_cap_count(o, hw, (pl1 + n) - pl1, pl1 - pl2);
            v += c;
            t += c;
        }
        if (o.start_cap == CAP_ROUND) {
            const V2 P0 = w.P(b0), P1 = w.P(s_p1);
            V2 fn[2];
            first_point_side_next(P0, P1, hw, fn);
            const uint32_t c = round_cap_count(o, hw, fn[SIDE_NEGATIVE] - P0, P0 - P1);
            v += c;
            t += c;
        }
    }
    *nv = v;
    *nt = t;
}

// What k_stroke_ends needs to run stroke_end_slot for one End slot outside its tile: the first three and the last
// three kept points of the sub-path (they overlap when it has fewer than six), the vertex ids of the two stream joins
// end() refers to, and where the output goes.  Slot numbering of the view: 0 = Begin point, 1 / 2 = second / third kept
// point, 3 / 4 / 5 = third-to-last / second-to-last / last kept point, 6 = the End slot.
struct EndRecord {
    V2 P[6];                 //  48
    float adv0, adv4, adv5;  //  12  advancement of slots 0, 4, 5
    uint32_t vo1, vo4;       //   8  absolute id of the first vertex of the joins at slots 1 and 4
    uint32_t src0, src5;     //   8  EndpointIds
    float t0, t5;            //   8
    uint8_t f[6];            //   6  TF_* of the six slots
    uint8_t rec, _pad;       //   2  m | REC_MNEC | REC_CLOSE
    uint32_t first_id;       //   4  absolute id of the End slot's first vertex
    uint32_t nv, nt;         //   8  what the count pass promised
    unsigned long long vtx_index, tri_index;  // 16  where its vertices / triangles go in the output arrays
    uint32_t _pad2[2];
};
static_assert(sizeof(EndRecord) == 128, "EndRecord is 128 bytes");

template <class W>
LYC_HD void end_record_fill(const W& w, uint32_t s, EndRecord* r) {
    const V2 rec = w.P(s);
    const uint32_t b0 = f2u(rec.x), fl = f2u(rec.y), m = fl & 3u;
    uint32_t q[6];
    q[0] = q[1] = q[2] = q[3] = q[4] = q[5] = b0;
    if (m >= 2u) {
        q[1] = next_kept_or_end(w, b0);
        q[2] = m > 2u ? next_kept_or_end(w, q[1]) : q[1];
        uint32_t l1 = s - 1u;
        while (slot_dropped(w, l1)) --l1;
        q[5] = l1;
        q[4] = prev_kept(w, l1);
        q[3] = m > 2u ? prev_kept(w, q[4]) : q[4];
    }
    for (int i = 0; i < 6; ++i) {
        r->P[i] = w.P(q[i]);
        r->f[i] = (uint8_t)(w.F(q[i]) >> 24);
    }
    r->adv0 = w.A(q[0]);
    r->adv4 = w.A(q[4]);
    r->adv5 = w.A(q[5]);
    r->src0 = w.SRC(q[0]);
    r->src5 = w.SRC(q[5]);
    r->t0 = w.T(q[0]);
    r->t5 = w.T(q[5]);
    r->rec = (uint8_t)fl;
    r->_pad = 0;
    r->_pad2[0] = r->_pad2[1] = 0u;
}
struct EndRecView {
    const EndRecord* r;
    uint32_t n_slots;
    LYC_HD V2 P(uint32_t s) const {
        if (s < 6u) return r->P[s];
        return v2(u2f(0u), u2f((uint32_t)r->rec));  // the sub-path record of the End slot: begin slot 0
    }
    LYC_HD float A(uint32_t s) const {
        if (s == 0u) return r->adv0;
        if (s == 4u) return r->adv4;
        if (s == 5u) return r->adv5;
        return s < 6u ? 0.0f : u2f(0x7FC00000u);
    }
    LYC_HD float T(uint32_t s) const { return s == 0u ? r->t0 : (s == 5u ? r->t5 : 1.0f); }
    LYC_HD uint32_t SRC(uint32_t s) const { return s == 0u ? r->src0 : r->src5; }
    LYC_HD uint32_t F(uint32_t s) const { return s < 6u ? (uint32_t)r->f[s] << 24 : (uint32_t)TF_END << 24; }
    LYC_HD uint32_t VO(uint32_t s) const { return s == 1u ? r->vo1 : r->vo4; }
    LYC_HD bool D(uint32_t s) const {
        const float a = A(s);
        return a != a;
    }
};

}  // namespace lyc
