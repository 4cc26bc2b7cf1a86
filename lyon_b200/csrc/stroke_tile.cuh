// stroke_tile.cuh -- per-item logic of the FUSED tile pipeline (host+device).
//
// The fused kernel (tile_kernels.cuh) keeps a tile of WHOLE paths in shared memory from the first byte read to
// the last vertex written.  These are the per-verb / per-path routines it runs on that tile; like everything in
// stroke_core.cuh they are LYC_HD so tests/emul can run the very same code on the CPU.
//
//   plan_verb        Wang counts of one verb (how many slots it feeds)                 [phase 1, thread per verb]
//   tile_gen_verb    all slots of one verb, generated INCREMENTALLY (t1 = t0 + step)   [phase 1, thread per verb]
//   tile_path_walk   merge filter + advancement chain + endpoint ids of one path       [phase 2, thread per path]
#pragma once
#include "stroke_core.cuh"

namespace lyc {

struct VerbPlan {
    uint32_t nslots;
    float n_a, n_b;  // Wang segment counts (f32, possibly non-integer: cubic_bezier.rs:1401) of the curve / the two halves
    float tsplit;    // flatten_quad's sharp-turn split parameter or -1 (stroke.rs:2755-2779)
};

LYC_HD VerbPlan plan_verb(uint8_t verb, const V2* points, uint32_t pt, uint32_t a, float tol) {
    VerbPlan r;
    r.nslots = 1u;  // Begin, Line, End, Close
    r.n_a = r.n_b = 1.0f;
    r.tsplit = -1.0f;
    if (verb == VERB_CUBIC) {
        V2 from = seg_from_point(points, pt, a);
        r.n_a = cubic_wang(from, points[pt], points[pt + 1], points[pt + 2], tol);
        r.nslots = to_u32_or_1(r.n_a);
    } else if (verb == VERB_QUAD) {
        V2 from = seg_from_point(points, pt, a);
        V2 ctrl = points[pt], to = points[pt + 1];
        float ts;
        if (find_sharp_turn(from, ctrl, to, &ts)) {
            r.tsplit = ts;
            V2 sp = quad_sample(from, ctrl, to, ts);  // QuadraticBezierSegment::split, quadratic_bezier.rs:208-223
            r.n_a = quad_wang(from, lerp(from, ctrl, ts), sp, tol);
            r.n_b = quad_wang(sp, lerp(ctrl, to, ts), to, tol);
            r.nslots = to_u32_or_1(r.n_a) + to_u32_or_1(r.n_b);
        } else {
            r.n_a = quad_wang(from, ctrl, to, tol);
            r.nslots = to_u32_or_1(r.n_a);
        }
    }
    return r;
}

// One fed point: position, curve parameter, flags, endpoint id and |P - prev| with the "maybe mergeable" sign flag
// (see point_gen_slot in stroke_core.cuh for the flag's meaning).
LYC_HD void tile_put_slot(const TileViews& w, const StrokeParams& p, uint32_t s, V2 pos, V2 prev, bool has_prev, float t,
                          uint8_t fl, uint32_t src) {
    w.pos[s] = pos;
    w.t[s] = t;
    w.flags[s] = fl;
    w.src[s] = src;
    float l = 0.0f;
    if (has_prev) {
        V2 d = pos - prev;
        float d2 = sqlen(d);
        l = sqrtf(d2);
        if (!(d2 >= p.flag_thr)) l = -l;
    }
    w.len()[s] = l;
}

// for_each_flattened_with_t of one quadratic (quadratic_bezier.rs:368-391) feeding slots s0.. .  With REMAP the
// curve is one half of a quad split at `tsplit` by flatten_quad and t is mapped back onto the whole curve
// (stroke.rs:2764 / 2770); is_flattening_step is decided on the un-mapped t of the half (stroke.rs:2765, 2771).
template <bool REMAP>
LYC_HD uint32_t tile_gen_quad(const TileViews& w, const StrokeParams& p, uint32_t s0, V2 from, V2 ctrl, V2 to, float n,
                              bool remap_is_first_half, float tsplit, uint32_t src, uint8_t first_flag) {
    const float step = 1.0f / n;
    const uint32_t cnt = to_u32_or_1(n);
    const QuadPoly poly = quad_poly(from, ctrl, to);
    float t = 0.0f;
    V2 prev = from;
    for (uint32_t k = 0; k < cnt; ++k) {
        V2 pos;
        float ts;
        if (k + 1u < cnt) {
            t = t + step;
            ts = t;
            pos = quad_poly_sample(poly, t);
        } else {
            ts = 1.0f;
            pos = to;
        }
        float t_out = ts;
        if (REMAP) t_out = remap_is_first_half ? ts * tsplit : tsplit + ts * (1.0f - tsplit);
        uint8_t fl = (ts != 1.0f ? TF_FS : 0) | (k == 0u ? first_flag : 0);
        tile_put_slot(w, p, s0 + k, pos, prev, true, t_out, fl, src);
        prev = pos;
    }
    return cnt;
}

// Phase 1: every slot of verb `verb` (first stored point `pt`, first slot s0).  `src` is the tile-relative index of
// the verb's endpoint in the point array (turned into an EndpointId by tile_path_walk).
LYC_HD void tile_gen_verb(const TileViews& w, const StrokeParams& p, uint8_t verb, const V2* points, uint32_t pt,
                          const VerbPlan& plan, uint32_t s0, uint32_t src) {
    const uint32_t a = p.attr_stride;
    switch (verb) {
        case VERB_BEGIN:
            tile_put_slot(w, p, s0, points[pt], v2(0.0f, 0.0f), false, 1.0f, TF_BEGIN, src);
            break;
        case VERB_LINE:
            tile_put_slot(w, p, s0, points[pt], seg_from_point(points, pt, a), true, 1.0f, TF_FIRST, src);
            break;
        case VERB_CUBIC: {  // cubic_bezier.rs:590-614
            const V2 from = seg_from_point(points, pt, a);
            const V2 to = points[pt + 2];
            const float step = 1.0f / plan.n_a;
            const uint32_t cnt = to_u32_or_1(plan.n_a);
            const CubicPoly poly = cubic_poly(from, points[pt], points[pt + 1], to);
            float t = 0.0f;
            V2 prev = from;
            for (uint32_t k = 0; k < cnt; ++k) {
                V2 pos;
                float ts;
                if (k + 1u < cnt) {
                    t = t + step;
                    ts = t;
                    pos = cubic_poly_sample(poly, t);
                } else {
                    ts = 1.0f;
                    pos = to;
                }
                uint8_t fl = (ts != 1.0f ? TF_FS : 0) | (k == 0u ? TF_FIRST : 0);
                tile_put_slot(w, p, s0 + k, pos, prev, true, ts, fl, src);
                prev = pos;
            }
            break;
        }
        case VERB_QUAD: {
            const V2 from = seg_from_point(points, pt, a);
            const V2 ctrl = points[pt], to = points[pt + 1];
            if (plan.tsplit >= 0.0f) {
                const float ts = plan.tsplit;
                const V2 sp = quad_sample(from, ctrl, to, ts);
                uint32_t n1 = tile_gen_quad<true>(w, p, s0, from, lerp(from, ctrl, ts), sp, plan.n_a, true, ts, src, TF_FIRST);
                tile_gen_quad<true>(w, p, s0 + n1, sp, lerp(ctrl, to, ts), to, plan.n_b, false, ts, src, 0);
            } else {
                tile_gen_quad<false>(w, p, s0, from, ctrl, to, plan.n_a, false, 0.0f, src, TF_FIRST);
            }
            break;
        }
        default:  // End / Close: not a point; pos/adv are written by tile_path_walk (sub-path record)
            w.flags[s0] = TF_END | (verb == VERB_CLOSE ? TF_CLOSE : 0);
            w.t[s0] = 1.0f;
            w.src[s0] = 0u;
            w.len()[s0] = 0.0f;
            break;
    }
}

// Phase 2: one path = slots [sa, sb) of the tile, sequential.  Same chains as path_walk (stroke_core.cuh):
// merge filter against the last kept point (stroke.rs:1318-1320, 1530-1536), advancement (stroke.rs:1546-1550,
// 1657-1661, 1759-1769, 2140) with the open-sub-path carry (stroke.rs:1359), may_need_empty_cap (stroke.rs:1180,
// 1533) and endpoint ids: sequential per event (stroke.rs:882-937) or point-array offsets (path.rs:1197-1263).
// `pt_rel` = index of the path's first stored point relative to the tile's first point.
LYC_HD void tile_path_walk(const TileViews& w, const StrokeParams& p, uint32_t sa, uint32_t sb, uint32_t pt_rel) {
    const float nan = u2f(0x7FC00000u);
    float carry = 0.0f, adv = 0.0f;
    uint32_t evt = 0, begin_slot = sa, last_kept = sa, m = 0, cur_id = 0;
    bool mnec = false;
    const bool with_ids = p.num_attributes > 0;
    float* len = w.len();
    for (uint32_t s = sa; s < sb; ++s) {
        const uint8_t fl = w.flags[s];
        if (fl & TF_BEGIN) {
            begin_slot = s;
            last_kept = s;
            m = 1;
            mnec = false;
            adv = carry;
            w.adv[s] = adv;
            w.src[s] = with_ids ? (w.src[s] - pt_rel) : evt;
            evt++;
        } else if (fl & TF_END) {
            const bool close = (fl & TF_CLOSE) != 0;
            if (close && m == 1u) mnec = true;
            V2 rec;
            rec.x = u2f(begin_slot);
            rec.y = u2f((m > 3u ? 3u : m) | (mnec ? REC_MNEC : 0u) | (close ? REC_CLOSE : 0u));
            w.pos[s] = rec;
            w.adv[s] = nan;
            if (!(close && m > 2u) && m >= 2u) carry = adv;
            m = 0;
        } else {
            if (fl & TF_FIRST) {
                cur_id = with_ids ? (w.src[s] - pt_rel) : evt;
                evt++;
            }
            w.src[s] = cur_id;
            float l = len[s];
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

}  // namespace lyc
