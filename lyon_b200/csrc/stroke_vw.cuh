// stroke_vw.cuh -- variable line width (StrokeOptions::variable_line_width, stroke.rs:653-767): one THREAD per path.
//
// With a per-endpoint width the stroker's state stops being local: a flattened step may be SKIPPED when both of its
// side points fall behind the previous ones (flattened_step, stroke.rs:1771-1789 -- `replace_last`), which changes the
// neighbours of every later join, and each side of a join is mitred, folded or clipped on its own from the side
// points of three consecutive kept points (compute_join_side_positions, stroke.rs:2060-2128).  This file therefore
// runs the reference's streaming state machine as it is, one path per thread, twice: once against CountSink (exact
// vertex / triangle counts per path), once against EmitSink at the scanned offsets.  Paths are independent, so a
// batch of many paths (the C4 map-style workload: 100k polylines) still fills the machine.  Join triangles, arcs, caps
// and the f64 cap clipping are the fixed-width routines of stroke_core.cuh (they take the radius from the vertex).
//
// Everything is LYC_HD: tests/emul runs the same code on the CPU against the oracle.
#pragma once
#include "stroke_core.cuh"

namespace lyc {

struct VwPoint {  // EndpointData (stroke.rs:320-352); the side points, folds and vertex ids live in `j`
    V2 pos;
    float hw, adv;
    uint32_t src;
    float t;
    bool fs;  // is_flattening_step
    Join j;
};
LYC_HD VwPoint vw_point_default(uint8_t line_join) {
    const float nan = u2f(0x7FC00000u);
    VwPoint e;
    e.pos = v2(0.0f, 0.0f);
    e.hw = nan;
    e.adv = nan;
    e.src = 0xFFFFFFFFu;
    e.t = 1.0f;
    e.fs = false;
    e.j.fast = false;
    e.j.line_join = line_join;
    e.j.normal = v2(0.0f, 0.0f);
    for (int s = 0; s < 2; ++s) {
        e.j.prev[s] = e.j.next[s] = e.j.single[s] = v2(nan, nan);
        e.j.has_single[s] = false;
        e.j.fold[s] = false;
        e.j.prev_v[s] = e.j.next_v[s] = 0xFFFFFFFFu;
    }
    return e;
}
LYC_HD VertexTemplate vw_vertex(const VwPoint& p) {
    VertexTemplate vt;
    vt.position_on_path = p.pos;
    vt.half_width = p.hw;
    vt.advancement = p.adv;
    vt.src = p.src;
    vt.t = p.t;
    return vt;
}

// stroke.rs:1809-1866: the side lines of an edge whose width changes tilt by asin((hw1 - hw0) / d)
LYC_HD void vw_edge_attachment(VwPoint& p0, VwPoint& p1) {
    const V2 edge = p1.pos - p0.pos;
    const float d = length(edge);
    const float edge_angle = angle_from_x_axis(edge);
    float vwidth_angle = libm_asin((p1.hw - p0.hw) / d);
    if (is_nan(vwidth_angle)) vwidth_angle = 0.0f;
    for (int side = 0; side < 2; ++side) {
        const float nl = side == SIDE_NEGATIVE ? -1.0f : 1.0f;
        const float normal_angle = edge_angle + nl * (LYC_PI_F * 0.5f + vwidth_angle);
        const V2 normal = v2(libm_cos(normal_angle), libm_sin(normal_angle));
        p0.j.next[side] = p0.pos + normal * p0.hw;
        p1.j.prev[side] = p1.pos + normal * p1.hw;
    }
    if (is_nan(p1.adv)) p1.adv = p0.adv + d;
}

// stroke.rs:2060-2128, one side of a join
LYC_HD void vw_join_side(const VwPoint& prev, VwPoint& join, const VwPoint& next, float miter_limit, int side) {
    const float sign = side == SIDE_NEGATIVE ? -1.0f : 1.0f;
    const V2 v0 = normalize(join.j.prev[side] - prev.j.next[side]);
    const V2 v1 = normalize(next.j.prev[side] - join.j.next[side]);
    const bool inward = cross(v0, v1) * sign > 0.0f;
    const bool forward = dot(v0, v1) > 0.0f;
    const V2 normal = compute_normal(v0, v1) * sign;
    const V2 path_v0 = normalize(join.pos - prev.pos);
    const V2 path_v1 = normalize(next.pos - join.pos);
    const bool normal_same_side = dot(v0 + v1, path_v0 + path_v1) >= 0.0f;
    const bool angle_is_sharp = inward && !forward && normal_same_side;
    if (angle_is_sharp) {
        const V2 extruded_normal = normal * join.hw;
        const float prev_length = join.adv - prev.adv;
        const float next_length = next.adv - join.adv;
        const float d_next = dot(extruded_normal, v1) - next_length;
        const float d_prev = dot(extruded_normal, -v0) - prev_length;
        if (fminf(d_next, d_prev) > 0.0f || sqlen(normal) < 1e-5f) join.j.fold[side] = true;
    }
    const bool concave = inward && normal_same_side && !join.j.fold[side];
    if (concave || ((join.j.line_join == JOIN_MITER || join.j.line_join == JOIN_MITER_CLIP) &&
                    !miter_limit_is_exceeded(normal, miter_limit))) {
        join.j.has_single[side] = true;
        join.j.single[side] = join.pos + normal * join.hw;
    } else if (join.j.line_join == JOIN_MITER_CLIP) {
        const V2 n0 = join.j.prev[side] - join.pos;
        const V2 n1 = join.j.next[side] - join.pos;
        V2 pn, nn;
        get_clip_intersections(n0, n1, normal, miter_limit * join.hw, &pn, &nn);
        join.j.prev[side] = join.pos + pn;
        join.j.next[side] = join.pos + nn;
    }
}

template <class Sink>
struct VwStroker {
    const StrokeParams& o;
    Sink& sink;
    // PointBuffer (stroke.rs:2371-2457) keeps the last three points in a ring, but only the last two are ever read
    // (`prev` and `join` of a step, the two ends of the last edge); two named points stay in registers where a ring
    // indexed by (idx + start) % 3 lives in local memory.  `count` saturates at 3 like the reference's.
    VwPoint p_mid, p_new;
    int count;
    VwPoint firsts[2];
    int n_firsts;
    float sub_path_start_advancement;
    bool may_need_empty_cap;

    LYC_HD VwStroker(const StrokeParams& o_, Sink& s_) : o(o_), sink(s_) {
        count = 0;
        n_firsts = 0;
        sub_path_start_advancement = 0.0f;
        may_need_empty_cap = false;
    }
    LYC_HD VwPoint& last() { return p_new; }
    LYC_HD VwPoint& last_but_one() { return p_mid; }
    LYC_HD void push(const VwPoint& p) {
        p_mid = p_new;
        p_new = p;
        if (count < 3) count += 1;
    }
    LYC_HD void replace_last(const VwPoint& p) { p_new = p; }

    // flattened_step (stroke.rs:1743-1806); returns the skip flag
    LYC_HD bool flattened_step(VwPoint& prev, VwPoint& join, VwPoint& next) {
        const V2 prev_edge = join.pos - prev.pos;
        const float prev_length = length(prev_edge);
        const V2 prev_tangent = prev_edge / prev_length;
        const V2 next_edge = next.pos - join.pos;
        const float next_length = length(next_edge);
        const V2 next_tangent = next_edge / next_length;
        const V2 normal = compute_normal(prev_tangent, next_tangent);
        if (is_nan(join.adv)) join.adv = prev.adv + prev_length;
        if (is_nan(next.adv)) next.adv = join.adv + next_length;
        const V2 p0 = join.pos + normal * join.hw;
        const V2 p1 = join.pos - normal * join.hw;
        join.j.prev[SIDE_POSITIVE] = join.j.next[SIDE_POSITIVE] = join.j.single[SIDE_POSITIVE] = p0;
        join.j.prev[SIDE_NEGATIVE] = join.j.next[SIDE_NEGATIVE] = join.j.single[SIDE_NEGATIVE] = p1;
        join.j.has_single[0] = join.j.has_single[1] = true;
        const V2 v0 = p0 - prev.j.next[SIDE_POSITIVE];
        const V2 v1 = p1 - prev.j.next[SIDE_NEGATIVE];
        if (dot(prev_edge, v0) < 0.0f && dot(prev_edge, v1) < 0.0f) return true;
        join.j.fast = true;
        join.j.normal = normal;
        emit_join_vertices(sink, vw_vertex(join), join.j);
        join.j.fast = false;
        return false;
    }

    // step_impl (stroke.rs:1380-1512); returns whether the point was added
    LYC_HD bool step(VwPoint next) {
        if (count > 0 && sqlen(last().pos - next.pos) < o.merge_thr) {
            if (count == 1) may_need_empty_cap = true;
            return false;
        }
        if (count > 0) {
            VwPoint& join = last();
            if (!join.fs || !next.fs) vw_edge_attachment(join, next);
        }
        bool skip = false;
        if (count > 1) {
            VwPoint& prev = last_but_one();
            VwPoint& join = last();
            bool fast_path = false;
            if (join.fs) fast_path = dot(join.pos - prev.pos, next.pos - join.pos) > 0.0f;
            if (fast_path) {
                join.j.line_join = JOIN_MITER;
                skip = flattened_step(prev, join, next);
            } else {
                vw_join_side(prev, join, next, o.miter_limit, SIDE_POSITIVE);
                vw_join_side(prev, join, next, o.miter_limit, SIDE_NEGATIVE);
                if (join.j.has_single[SIDE_POSITIVE]) join.j.fold[SIDE_NEGATIVE] = false;
                if (join.j.has_single[SIDE_NEGATIVE]) join.j.fold[SIDE_POSITIVE] = false;
                emit_join_vertices(sink, vw_vertex(join), join.j);
            }
            if (!skip) {
                if (count > 2) add_edge_triangles(sink, prev.j, join.j);
                tessellate_join(sink, vw_vertex(join), join.j, o);
                if (count == 2) {
                    firsts[0] = prev;
                    firsts[1] = join;
                    n_firsts = 2;
                }
            }
        }
        if (skip)
            replace_last(next);
        else
            push(next);
        return true;
    }

    // tessellate_last_edge (stroke.rs:2130-2206)
    LYC_HD void last_edge(const VwPoint& p0, VwPoint& p1, bool is_first_edge) {
        const V2 v = p1.pos - p0.pos;
        p1.adv = p0.adv + length(v);
        const VertexTemplate vt = vw_vertex(p1);
        p1.j.prev[SIDE_POSITIVE] = clip_cap_side(o.end_cap, p1.pos, p0.pos, p1.j.prev[SIDE_POSITIVE], p0.j.next[SIDE_POSITIVE], p1.hw);
        p1.j.prev_v[SIDE_POSITIVE] = sink.add_vertex(vt, (p1.j.prev[SIDE_POSITIVE] - p1.pos) / p1.hw, false);
        p1.j.prev[SIDE_NEGATIVE] = clip_cap_side(o.end_cap, p1.pos, p0.pos, p1.j.prev[SIDE_NEGATIVE], p0.j.next[SIDE_NEGATIVE], p1.hw);
        p1.j.prev_v[SIDE_NEGATIVE] = sink.add_vertex(vt, (p1.j.prev[SIDE_NEGATIVE] - p1.pos) / p1.hw, true);
        if (!is_first_edge) add_edge_triangles(sink, p0.j, p1.j);
        if (o.end_cap == CAP_ROUND)
            tessellate_round_cap_nl(sink, vt, p1.pos, p1.hw, p1.j.prev[SIDE_POSITIVE] - p1.pos, p1.j.prev_v[SIDE_POSITIVE],
                                    p1.j.prev_v[SIDE_NEGATIVE], v, false, o);
    }
    // tessellate_first_edge (stroke.rs:2208-2276)
    LYC_HD void first_edge(VwPoint& first, const VwPoint& second) {
        const VertexTemplate vt = vw_vertex(first);
        const V2 sp = clip_cap_side(o.start_cap, first.pos, second.pos, first.j.next[SIDE_POSITIVE], second.j.prev[SIDE_POSITIVE], first.hw);
        first.j.next_v[SIDE_POSITIVE] = sink.add_vertex(vt, (sp - first.pos) / first.hw, false);
        const V2 sn = clip_cap_side(o.start_cap, first.pos, second.pos, first.j.next[SIDE_NEGATIVE], second.j.prev[SIDE_NEGATIVE], first.hw);
        first.j.next_v[SIDE_NEGATIVE] = sink.add_vertex(vt, (sn - first.pos) / first.hw, true);
        add_edge_triangles(sink, first.j, second.j);
        if (o.start_cap == CAP_ROUND)
            tessellate_round_cap_nl(sink, vt, first.pos, first.hw, first.j.next[SIDE_NEGATIVE] - first.pos,
                                    first.j.next_v[SIDE_NEGATIVE], first.j.next_v[SIDE_POSITIVE], first.pos - second.pos, true, o);
    }
    // end_with_caps (stroke.rs:1322-1378)
    LYC_HD void end_with_caps() {
        if (may_need_empty_cap && count == 1) {
            tessellate_empty_cap(sink, vw_vertex(last()), o);  // count == 1
            return;
        }
        if (count >= 2) {
            VwPoint p0 = last_but_one();
            VwPoint p1 = last();
            last_edge(p0, p1, count == 2);
            sub_path_start_advancement = p1.adv;
            if (count > 2) {
                p0 = firsts[0];
                p1 = firsts[1];
            }
            first_edge(p0, p1);
        }
    }
    // close (stroke.rs:1206-1279), count > 2
    LYC_HD void close_subpath() {
        VwPoint p = firsts[0];
        const float advancement = p.adv;
        p.adv = u2f(0x7FC00000u);
        const bool segment_added = step(p);
        if (!segment_added) last().pos = p.pos;  // stroke.rs:1228-1238
        if (n_firsts >= 2) {
            step(firsts[1]);
            VwPoint& p0 = last_but_one();
            VwPoint& p1 = last();
            VertexTemplate vt = vw_vertex(p0);
            vt.advancement = advancement;
            const V2 rp = p0.j.has_single[SIDE_POSITIVE] ? p0.j.single[SIDE_POSITIVE] : p0.j.next[SIDE_POSITIVE];
            p0.j.next_v[SIDE_POSITIVE] = sink.add_vertex(vt, (rp - p0.pos) / p0.hw, false);
            const V2 rn = p0.j.has_single[SIDE_NEGATIVE] ? p0.j.single[SIDE_NEGATIVE] : p0.j.next[SIDE_NEGATIVE];
            p0.j.next_v[SIDE_NEGATIVE] = sink.add_vertex(vt, (rn - p0.pos) / p0.hw, true);
            add_edge_triangles(sink, p0.j, p1.j);
        }
    }
    // end (stroke.rs:1179-1193)
    LYC_HD void end(bool close) {
        may_need_empty_cap = may_need_empty_cap || (close && count == 1);
        if (close && count > 2)
            close_subpath();
        else
            end_with_caps();
        count = 0;
        n_firsts = 0;
    }
};

// tessellate_with_ids_vw (stroke.rs:653-767) over the verbs [vb, ve) of one path whose first stored point is pt_path.
// Endpoint ids are offsets in the path's point array (path.rs:1197-1263); the width attribute of an endpoint sits in
// the attribute slots stored right after it (path.rs:1266-1282).
template <class Sink>
LYC_HD void vw_stroke_path(const uint8_t* verbs, uint32_t vb, uint32_t ve, const V2* points, uint32_t pt_path, const StrokeParams& o,
                           Sink& sink) {
    VwStroker<Sink> st(o, sink);
    const uint32_t a = o.attr_stride;
    const float* fpts = reinterpret_cast<const float*>(points);
    const float base_width = o.line_width;
    uint32_t cur = pt_path, cur_ep = pt_path;
    for (uint32_t v = vb; v < ve; ++v) {
        const uint8_t verb = verbs[v];
        if (verb == VERB_BEGIN || verb == VERB_LINE) {
            const uint32_t ep = cur;
            cur += 1u + a;
            cur_ep = ep;
            VwPoint d = vw_point_default(o.line_join);
            d.pos = points[ep];
            d.src = ep - pt_path;
            d.hw = base_width * fpts[2u * (ep + 1u) + (uint32_t)o.vw_attr] * 0.5f;  // stroke.rs:671, 689
            if (verb == VERB_BEGIN) {
                st.may_need_empty_cap = false;
                d.adv = st.sub_path_start_advancement;
            }
            st.step(d);
        } else if (verb == VERB_QUAD || verb == VERB_CUBIC) {
            const bool cubic = verb == VERB_CUBIC;
            const uint32_t pt = cur;
            const uint32_t ep = cur + (cubic ? 2u : 1u);
            cur += (cubic ? 3u : 2u) + a;
            const uint32_t from_ep = cur_ep;
            cur_ep = ep;
            const float start_width = base_width * fpts[2u * (from_ep + 1u) + (uint32_t)o.vw_attr];  // stroke.rs:705-706, 731-732
            const float end_width = base_width * fpts[2u * (ep + 1u) + (uint32_t)o.vw_attr];
            // the flattened points of the curve, in order (eval_slot replays the reference's flattening: Wang count,
            // accumulated t, exact last point; quadratics possibly split in two by flatten_quad, stroke.rs:2755-2779)
            float tsplit = -1.0f;
            const uint32_t nslots = seg_slot_count(verb, points, pt, a, o.tolerance, &tsplit);
            for (uint32_t k = 0; k < nslots; ++k) {
                const SlotPoint r = eval_slot(verb, points, pt, a, o.tolerance, tsplit, k);
                VwPoint d = vw_point_default(o.line_join);
                d.pos = r.pos;
                d.src = ep - pt_path;
                d.t = r.t;
                d.fs = r.fs;
                d.hw = (start_width * (1.0f - r.t) + end_width * r.t) * 0.5f;  // stroke.rs:1022, 1058
                st.step(d);
            }
        } else if (verb == VERB_CLOSE) {
            cur += 1u + a;  // the first point (and its attributes) is stored again (path.rs:537-539)
            st.end(true);
        } else {
            st.end(false);
        }
    }
}

}  // namespace lyc
