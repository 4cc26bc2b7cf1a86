// stroke_tile.cuh -- per-item logic of the FUSED tile pipeline (host+device).
//
// The fused kernel (tile_kernels.cuh) keeps a tile of WHOLE paths in shared memory from the first byte read to
// the last vertex written.  These are the per-verb / per-path routines it runs on that tile; like everything in
// stroke_core.cuh they are LYC_HD so tests/emul can run the very same code on the CPU.
//
//   plan_verb        Wang counts of one verb (how many slots it feeds)                 [phase 1a, thread per verb]
//   tile_gen_verb    slots of one verb; for curves only the accumulated t of each      [phase 1a, thread per verb]
//   tile_eval_slot   position + edge length of one curve slot from its t               [phase 1b, thread per slot]
//   tile_path_walk   merge filter + advancement chain of one path                      [phase 2,  thread per path]
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
LYC_HD float tile_len_flagged(const StrokeParams& p, V2 pos, V2 prev) {
    V2 d = pos - prev;
    float d2 = sqlen(d);
    float l = sqrtf(d2);
    if (!(d2 >= p.flag_thr)) l = -l;  // also flags NaN distances; -0.0f keeps the flag for zero length
    return l;
}
LYC_HD void tile_put_slot(const TileViews& w, const StrokeParams& p, uint32_t s, V2 pos, V2 prev, bool has_prev, float t,
                          uint8_t fl, uint32_t src) {
    w.pos[s] = pos;
    w.t[s] = t;
    w.flags[s] = fl;
    w.src[s] = src;
    w.len()[s] = has_prev ? tile_len_flagged(p, pos, prev) : 0.0f;
}

// for_each_flattened_with_t of one quadratic (quadratic_bezier.rs:368-391) feeding slots s0.. .  The curve is one
// half of a quad split at `tsplit` by flatten_quad and t is mapped back onto the whole curve (stroke.rs:2764 /
// 2770); is_flattening_step is decided on the un-mapped t of the half (stroke.rs:2765, 2771).
LYC_HD uint32_t tile_gen_quad_half(const TileViews& w, const StrokeParams& p, uint32_t s0, V2 from, V2 ctrl, V2 to, float n,
                                   bool first_half, float tsplit, uint32_t src, uint8_t first_flag) {
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
        float t_out = first_half ? ts * tsplit : tsplit + ts * (1.0f - tsplit);
        uint8_t fl = (ts != 1.0f ? TF_FS : 0) | (k == 0u ? first_flag : 0);
        tile_put_slot(w, p, s0 + k, pos, prev, true, t_out, fl, src);
        prev = pos;
    }
    return cnt;
}

// One quadratic that flatten_quad splits at a sharp turn (stroke.rs:2755-2779), its slots dealt round-robin: this call
// writes slots j0, j0 + stride, ... of the verb (tile kernel: one WARP per split quadratic, j0 = lane, stride = 32, instead
// of one lane of the verb phase running both halves alone; tests/emul: every j0 of a stride).  A slot replays the
// accumulated parameter of its half from t = 0 (one FADD per earlier slot, continuing from the caller's previous slot
// while it stays in the same half) and re-evaluates its predecessor's position from the predecessor's t, so the slots
// are independent of each other and bit-identical to the sequential tile_gen_quad_half.
LYC_HD void tile_gen_split_quad(const TileViews& w, const StrokeParams& p, const V2* points, uint32_t pt, uint32_t s0,
                                uint32_t src, uint32_t j0, uint32_t stride) {
    const VerbPlan plan = plan_verb(VERB_QUAD, points, pt, p.attr_stride, p.tolerance);
    if (!(plan.tsplit >= 0.0f)) return;  // the pre-pass found a split: cannot happen
    const V2 from = seg_from_point(points, pt, p.attr_stride);
    const V2 ctrl = points[pt], to = points[pt + 1];
    const float tsplit = plan.tsplit;
    const V2 sp = quad_sample(from, ctrl, to, tsplit);
    const uint32_t cnt_a = to_u32_or_1(plan.n_a), cnt_b = to_u32_or_1(plan.n_b);
    const QuadPoly poly_a = quad_poly(from, lerp(from, ctrl, tsplit), sp), poly_b = quad_poly(sp, lerp(ctrl, to, tsplit), to);
    const float step_a = 1.0f / plan.n_a, step_b = 1.0f / plan.n_b;
    const unsigned long long total = (unsigned long long)cnt_a + cnt_b;
    bool in_b = false;
    uint32_t k_done = 0u;  // additions replayed so far in the current half
    float t = 0.0f;        // the half's accumulated parameter after k_done additions
    for (unsigned long long j = j0; j < total; j += stride) {
        const bool hb = j >= cnt_a;
        const uint32_t k = (uint32_t)(hb ? j - cnt_a : j);
        if (hb != in_b) {
            in_b = hb;
            k_done = 0u;
            t = 0.0f;
        }
        const float step = hb ? step_b : step_a;
        const uint32_t cnt = hb ? cnt_b : cnt_a;
        while (k_done < k) {
            t = t + step;
            ++k_done;
        }
        const V2 prev = k == 0u ? (hb ? sp : from) : quad_poly_sample(hb ? poly_b : poly_a, t);
        V2 pos;
        float ts;
        if (k + 1u < cnt) {
            ts = t + step;
            pos = quad_poly_sample(hb ? poly_b : poly_a, ts);
        } else {
            ts = 1.0f;
            pos = hb ? to : sp;
        }
        const float t_out = hb ? tsplit + ts * (1.0f - tsplit) : ts * tsplit;
        tile_put_slot(w, p, s0 + (uint32_t)j, pos, prev, true, t_out, ts != 1.0f ? TF_FS : 0, src);
    }
}

// Phase 1a (thread per verb): every slot of verb `verb` (first stored point `pt`, first slot s0, EndpointId `src`).
// Only the SEQUENTIAL part of a verb runs here -- the accumulated parameter t1 = t0 + step of
// for_each_flattened_with_t (cubic_bezier.rs:605-611, quadratic_bezier.rs:382-388) -- and every slot is left with its
// kind in the low flag bits and the verb's point index parked in `adv`; phase 1b computes positions and edge lengths
// with one thread per slot (no divergence between verbs of different kinds and counts).  Low flag bits for phase 1b:
//   TF_EVAL [| TF_CUBIC] [| TF_FIRST] [| TF_LAST]   curve sample        TF_CUBIC alone   line endpoint
//   TF_BEGIN / TF_END (high bits)                    Begin / End slot    none             finished here (split quadratic)
// The rare quadratics split by flatten_quad (stroke.rs:2755-2779) are finished here.
LYC_HD void tile_gen_verb(const TileViews& w, const StrokeParams& p, uint8_t verb, const V2* points, uint32_t pt,
                          const VerbPlan& plan, uint32_t s0, uint32_t src) {
    const uint32_t a = p.attr_stride;
    if (verb == VERB_QUAD && plan.tsplit >= 0.0f) {
        const V2 from = seg_from_point(points, pt, a);
        const V2 ctrl = points[pt], to = points[pt + 1];
        const float ts = plan.tsplit;
        const V2 sp = quad_sample(from, ctrl, to, ts);
        uint32_t n1 = tile_gen_quad_half(w, p, s0, from, lerp(from, ctrl, ts), sp, plan.n_a, true, ts, src, 0);
        tile_gen_quad_half(w, p, s0 + n1, sp, lerp(ctrl, to, ts), to, plan.n_b, false, ts, src, 0);
        return;
    }
    const bool curve = verb == VERB_CUBIC || verb == VERB_QUAD;
    const uint32_t cnt = curve ? to_u32_or_1(plan.n_a) : 1u;
    const float step = curve ? 1.0f / plan.n_a : 0.0f;
    uint8_t kind;
    switch (verb) {
        case VERB_BEGIN: kind = TF_BEGIN; break;
        case VERB_LINE: kind = TF_CUBIC; break;
        case VERB_CUBIC: kind = TF_EVAL | TF_CUBIC; break;
        case VERB_QUAD: kind = TF_EVAL; break;
        case VERB_CLOSE: kind = TF_END | TF_CLOSE; break;
        default: kind = TF_END; break;
    }
    const float ptbits = u2f(pt);
    float t = 0.0f;
    for (uint32_t k = 0; k < cnt; ++k) {
        float ts = 1.0f;
        if (k + 1u < cnt) {
            t = t + step;
            ts = t;
        }
        uint8_t fl = kind;
        if (curve) fl |= (ts != 1.0f ? TF_FS : 0) | (k == 0u ? TF_FIRST : 0) | (k + 1u < cnt ? 0 : TF_LAST);
        w.t[s0 + k] = ts;
        w.flags[s0 + k] = fl;
        w.src[s0 + k] = src;
        w.adv[s0 + k] = ptbits;
    }
}

// Phase 1a, complete (thread per verb): every slot of verb `verb` finished in one loop -- accumulated parameter, position
// from the polynomial form held in registers (cubic_bezier.rs:1318-1365 / quadratic_bezier.rs:891-939; the last point is
// the exact endpoint), flags, EndpointId and |P - prev| with the "maybe mergeable" flag.  The per-verb loop costs a FADD, a
// Horner evaluation and a square root per point; evaluating slot by slot afterwards (tile_eval_slot) re-derives the
// polynomial and re-loads the control points for every point, four times the instructions.  Split quadratics are not
// handled here (tile_gen_split_slot).  `n` = the verb's Wang count (pre-pass).
LYC_HD void tile_gen_verb_full(const TileViews& w, const StrokeParams& p, uint8_t verb, const V2* points, uint32_t pt, float n,
                               uint32_t s0, uint32_t src, uint32_t begin_slot = 0u) {
    if (verb == VERB_END || verb == VERB_CLOSE) {
        w.t[s0] = 1.0f;
        w.flags[s0] = (uint8_t)(TF_END | (verb == VERB_CLOSE ? TF_CLOSE : 0));
        w.src[s0] = src;
        w.len()[s0] = 0.0f;
        // the sub-path record: the Begin slot now, the kept-point count and may_need_empty_cap once the merge filter has
        // run (tile_end_record_finish)
        w.pos[s0] = v2(u2f(begin_slot), u2f(verb == VERB_CLOSE ? REC_CLOSE : 0u));
        return;
    }
    const bool curve = verb == VERB_QUAD || verb == VERB_CUBIC, cubic = verb == VERB_CUBIC;
    const uint32_t npts = cubic ? 3u : (curve ? 2u : 1u);
    const V2 to = points[pt + npts - 1u];
    if (verb == VERB_BEGIN) {
        w.pos[s0] = to;
        w.t[s0] = 1.0f;
        w.flags[s0] = TF_BEGIN;
        w.src[s0] = src;
        w.len()[s0] = 0.0f;
        return;
    }
    const V2 from = seg_from_point(points, pt, p.attr_stride);
    uint32_t cnt = 1u;
    float step = 0.0f;
    V2 a1 = v2(0.0f, 0.0f), a2 = a1, a3 = a1;
    if (curve) {
        cnt = to_u32_or_1(n);
        step = 1.0f / n;
        const V2 c1 = points[pt];
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
    }
    float t = 0.0f;
    V2 prev = from;
    for (uint32_t k = 0; k < cnt; ++k) {
        V2 pos = to;
        float ts = 1.0f;
        if (k + 1u < cnt) {
            t = t + step;
            ts = t;
            // Horner as the reference has it: v = a0; v += a1*t; v += a2*t^2; [v += a3*t^3]
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
        const uint32_t s = s0 + k;
        // |P - prev| and the "maybe mergeable" flag (point_gen_slot in stroke_core.cuh): a pair further apart than sqrt(5)x
        // the merge distance can never be merged; NaN distances are flagged too
        const V2 d = pos - prev;
        const float d2 = sqlen(d);
        w.pos[s] = pos;
        w.t[s] = ts;
        w.flags[s] = (uint8_t)((ts != 1.0f ? TF_FS : 0) | (!(d2 >= p.flag_thr) ? TF_FLAGGED : 0));
        w.src[s] = src;
        w.len()[s] = sqrtf(d2);
        prev = pos;
    }
}

// Phase 1b (thread per slot): position of a slot and its distance to the previous fed point.  Curve samples come from
// the polynomial form (cubic_bezier.rs:1318-1365 / quadratic_bezier.rs:891-939; the last point of a curve is the exact
// endpoint); the previous sample of the same curve is re-evaluated rather than read, so the phase needs no ordering.
LYC_HD void tile_eval_slot(const TileViews& w, const StrokeParams& p, const V2* points, uint32_t s) {
    const uint8_t fl = w.flags[s];
    if (fl & TF_END) {
        w.len()[s] = 0.0f;  // pos/adv are written by tile_path_walk (sub-path record)
        return;
    }
    if (!(fl & (TF_EVAL | TF_BEGIN | TF_CUBIC))) return;  // finished in phase 1a (split quadratic)
    // One code path for Begin, line endpoints and curve samples: the point is either the verb's last stored point
    // (`to`) or a sample of the polynomial form; only the coefficients differ between quadratics and cubics.
    const uint32_t pt = f2u(w.adv[s]);
    const bool curve = (fl & TF_EVAL) != 0, cubic = curve && (fl & TF_CUBIC) != 0;
    const uint32_t npts = cubic ? 3u : (curve ? 2u : 1u);
    const V2 to = points[pt + npts - 1u];
    const V2 from = (fl & TF_BEGIN) ? to : seg_from_point(points, pt, p.attr_stride);
    const bool first = !curve || (fl & TF_FIRST) != 0, last = !curve || (fl & TF_LAST) != 0;
    V2 pos = to, prev = from;
    if (!(first && last)) {
        const float t = w.t[s];
        const float tp = first ? 0.0f : w.t[s - 1u];
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
        // Horner as the reference has it: v = a0; v += a1*t; v += a2*t^2; [v += a3*t^3] -- for t and for the previous t
        float u = t, up = tp;
        V2 v = from + a1 * u, vp = from + a1 * up;
        u = u * t;
        up = up * tp;
        v = v + a2 * u;
        vp = vp + a2 * up;
        if (cubic) {
            u = u * t;
            up = up * tp;
            v = v + a3 * u;
            vp = vp + a3 * up;
        }
        if (!last) pos = v;
        if (!first) prev = vp;
    }
    w.pos[s] = pos;
    w.len()[s] = (fl & TF_BEGIN) ? 0.0f : tile_len_flagged(p, pos, prev);
}
// Phase 1c: a slot is "special" for the walk when it is not an ordinary kept point: Begin, End, or possibly within
// the merge distance of its predecessor.
template <class W>
LYC_HD bool tile_slot_special(const W& w, uint32_t s) {
    return (w.flags[s] & (TF_BEGIN | TF_END)) != 0 || len_flagged(w.len()[s]);
}

// Phase 1d: a flagged ("maybe mergeable") slot whose predecessor is certainly kept -- an ordinary point or a Begin -- is
// compared by the merge filter (stroke.rs:1530-1536) against exactly that predecessor, so the decision is local.  When
// the pair is NOT within the merge distance the slot is an ordinary kept point after all: its length loses the flag
// and the caller clears its special bit, so it no longer interrupts the walk's 32-point steps.  Slots may be resolved
// concurrently: a neighbour seen still flagged only leaves this slot to the walk's general step (same result).
template <class W>
LYC_HD bool tile_resolve_flagged(const W& w, const StrokeParams& p, uint32_t s) {
    if (s == 0u) return false;
    if (w.flags[s] & (TF_BEGIN | TF_END)) return false;
    if (!len_flagged(w.len()[s])) return false;
    const uint8_t fp = w.flags[s - 1u];
    if (fp & TF_END) return false;
    if (!(fp & TF_BEGIN) && len_flagged(w.len()[s - 1u])) return false;
    const V2 pl = w.pos[s - 1u], ps = w.pos[s];
    if (!(sqlen(pl - ps) >= p.merge_thr)) return false;  // merged (or NaN): the walk decides
    w.len()[s] = length(ps - pl);
    return true;
}

// Phase 2: one path = slots [sa, sb) of the tile, sequential.  Same chains as path_walk (stroke_core.cuh):
// merge filter against the last kept point (stroke.rs:1318-1320, 1530-1536), advancement (stroke.rs:1546-1550,
// 1657-1661, 1759-1769, 2140) with the open-sub-path carry (stroke.rs:1359), may_need_empty_cap (stroke.rs:1180,
// 1533).  `special` holds one bit per slot (phase 1c): aligned groups of 32 ordinary points -- the overwhelmingly
// common case -- reduce to the bare f32 addition chain on vector loads/stores.
template <class W, class SP>
LYC_HD void tile_path_walk(const W& w, const StrokeParams& p, uint32_t sa, uint32_t sb, SP special) {
    const float nan = u2f(0x7FC00000u);
    float carry = 0.0f, adv = 0.0f;
    uint32_t begin_slot = sa, last_kept = sa, m = 0;
    bool mnec = false;
    const auto len = w.len();
    uint32_t s = sa;
    while (s < sb) {
#ifdef __CUDA_ARCH__
        // 32 ordinary points: eight 16-byte loads first, then the 32 dependent additions, then eight 16-byte stores --
        // two instructions per point, nothing but the additions on the critical path.
        if ((s & 3u) == 0u && s + 32u <= sb && m >= 1u && last_kept == s - 1u) {
            const uint32_t sh = s & 31u;
            uint32_t bits = special[s >> 5] >> sh;
            if (sh) bits |= special[(s >> 5) + 1u] << (32u - sh);
            if (bits == 0u) {
                float4 l[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) l[q] = *reinterpret_cast<const float4*>(&len[s + 4 * q]);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    l[q].x = adv = adv + l[q].x;
                    l[q].y = adv = adv + l[q].y;
                    l[q].z = adv = adv + l[q].z;
                    l[q].w = adv = adv + l[q].w;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(&w.adv[s + 4 * q]) = l[q];
                m += 32u;
                last_kept = s + 31u;
                s += 32u;
                continue;
            }
        }
        // A shorter run of ordinary points: the bare f32 addition chain, four slots (one 16-byte load / store) at a time.  The
        // next quad's lengths are loaded before this quad's results are stored (the arrays may alias as far as the
        // compiler knows), so the loads stay off the dependent chain.
        if ((s & 3u) == 0u && s + 4u <= sb && m >= 1u && last_kept == s - 1u) {
            uint32_t word = special[s >> 5] >> (s & 31u);
            uint32_t left = 32u - (s & 31u);
            if ((word & 0xFu) == 0u) {
                float4 l = *reinterpret_cast<const float4*>(&len[s]);
                const uint32_t s_in = s;
                for (;;) {
                    const uint32_t s2 = s + 4u;
                    bool more = false;
                    float4 ln = l;
                    if (s2 + 4u <= sb) {
                        word >>= 4;
                        left -= 4u;
                        if (left == 0u) {
                            word = special[s2 >> 5];
                            left = 32u;
                        }
                        more = (word & 0xFu) == 0u;
                        ln = *reinterpret_cast<const float4*>(&len[s2]);
                    }
                    float4 a;
                    a.x = adv = adv + l.x;
                    a.y = adv = adv + l.y;
                    a.z = adv = adv + l.z;
                    a.w = adv = adv + l.w;
                    *reinterpret_cast<float4*>(&w.adv[s]) = a;
                    s = s2;
                    if (!more) break;
                    l = ln;
                }
                m += s - s_in;
                last_kept = s - 1u;
                continue;
            }
        }
#else
        if (m >= 1u && last_kept == s - 1u && !((special[s >> 5] >> (s & 31u)) & 1u)) {
            adv = adv + len[s];
            w.adv[s] = adv;
            m++;
            last_kept = s;
            ++s;
            continue;
        }
#endif
        const uint8_t fl = w.flags[s];
        if (fl & TF_BEGIN) {
            begin_slot = s;
            last_kept = s;
            m = 1;
            mnec = false;
            adv = carry;
            w.adv[s] = adv;
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
            float l = len[s];
            bool keep = true;
            if (len_flagged(l) || last_kept != s - 1u) {
                V2 pl = w.pos[last_kept], ps = w.pos[s];
                if (sqlen(pl - ps) < p.merge_thr) {
                    w.adv[s] = nan;
                    if (m == 1u) mnec = true;
                    keep = false;
                } else {
                    l = length(ps - pl);
                }
            }
            if (keep) {
                adv = adv + l;
                w.adv[s] = adv;
                last_kept = s;
                m++;
            }
        }
        ++s;
    }
}

// ---- the merge filter and the advancement chain of the CTA-tile kernel ------------------------------------------------
// The merge filter of fixed_width_step_impl (stroke.rs:1530-1536) compares a point with the last KEPT point, which is
// sequential -- but only inside a RUN of consecutive flagged points: an unflagged point is certainly kept (it is further
// than sqrt(5)x the merge distance from its predecessor, which in turn is within the merge distance of whatever point
// is the last kept one), so the point before a run is kept whatever happened earlier.  One thread resolves one run, all
// runs of the tile at once (tile kernel: a thread per slot looks for run starts): a dropped point gets its bit in `dropmask` and the
// length -0.0f (adv + -0.0f == adv: the advancement chain then needs no branch), a kept one the distance to the last kept
// point; the unflagged point right after a dropped one is re-measured against the last kept point as the reference does
// (last_kept != s - 1).  `s` must be a run start: flagged, its predecessor not flagged (an ordinary point or a Begin).
template <class W>
LYC_HD bool tile_run_start(const W& w, uint32_t s) {
    const uint8_t fl = w.flags[s];
    if (!(fl & TF_FLAGGED) || (fl & (TF_BEGIN | TF_END)) || s == 0u) return false;
    return !(w.flags[s - 1u] & TF_FLAGGED);
}
template <class W>
LYC_HD void tile_resolve_run(const W& w, const StrokeParams& p, uint32_t s, uint32_t n_slots) {
    const auto len = w.len();
    uint32_t last_kept = s - 1u;
    bool prev_flagged = true;  // (of slot r - 1, for r > s)
    for (uint32_t r = s; r < n_slots; ++r) {
        const uint8_t fl = w.flags[r];
        if (fl & (TF_BEGIN | TF_END)) break;
        const bool flagged = (fl & TF_FLAGGED) != 0;
        // the run goes on while flagged points follow each other, and through the point after a dropped one; a flagged
        // point behind a kept unflagged one starts a run of its own (another thread's)
        if (r != s && !((flagged && prev_flagged) || last_kept != r - 1u)) break;
        const V2 pl = w.pos[last_kept], ps = w.pos[r];
        if (sqlen(pl - ps) < p.merge_thr) {
            LYC_ATOMIC_OR_U32(&w.dropmask[r >> 5], 1u << (r & 31u));
            len[r] = -0.0f;
        } else {
            len[r] = length(ps - pl);
            last_kept = r;
        }
        prev_flagged = flagged;
    }
}

// After the runs: the dropped bit of slot s moves from the mask into its flag byte (thread per slot; a warp of 32
// consecutive slots shares one mask word and mostly finds it empty).
template <class W>
LYC_HD void tile_merge_dropped(const W& w, uint32_t s) {
    if ((w.dropmask[s >> 5] >> (s & 31u)) & 1u) w.flags[s] = (uint8_t)(w.flags[s] | TF_DROPPED);
}

// The sub-path record of End slot `e` (written by the verb phase with the Begin slot and the closed bit) gets what depends on
// the merge filter: the number of kept points capped at 3, and may_need_empty_cap (stroke.rs:1180, 1533: a point was dropped
// while only the Begin point was kept -- i.e. the point after the Begin was dropped -- or close() on a single point).
template <class W>
LYC_HD void tile_end_record_finish(const W& w, uint32_t e) {
    V2 rec = w.pos[e];
    const uint32_t b0 = f2u(rec.x);
    const bool close = (f2u(rec.y) & REC_CLOSE) != 0u;
    uint32_t m = (b0 < e && (w.flags[b0] & TF_BEGIN)) ? 1u : 0u;
    for (uint32_t r = b0 + 1u; r < e && m < 3u && m >= 1u; ++r)
        if (!(w.flags[r] & (TF_BEGIN | TF_END)) && !w.D(r)) ++m;
    bool mnec = m >= 1u && b0 + 1u < e && w.D(b0 + 1u);
    if (close && m == 1u) mnec = true;
    rec.y = u2f(m | (mnec ? REC_MNEC : 0u) | (close ? REC_CLOSE : 0u));
    w.pos[e] = rec;
}

// The advancement chain of one path = slots [sa, sb) of the tile, after tile_resolve_run and tile_end_record_finish:
// adv += |edge| over every point of a sub-path (stroke.rs:1546-1550, 1657-1661, 1759-1769, 2140) with the open-sub-path
// carry (stroke.rs:1359).  In place: the slot's length becomes its advancement.  The chain itself is branch-free: 32 points
// per step on vector loads / stores (eight 16-byte loads, 32 dependent additions, eight 16-byte stores), then 4, then 1.
// Nothing that decides vertex or triangle COUNTS depends on it, so the tile kernel runs it after its counts are published.
template <class W>
LYC_HD void tile_path_chain(const W& w, const StrokeParams& p, uint32_t sa, uint32_t sb) {
    (void)p;
    const float nan = u2f(0x7FC00000u);
    const auto len = w.len();
    float carry = 0.0f, adv = 0.0f;
    uint32_t s = sa;
    while (s < sb) {
        const uint8_t fl = w.flags[s];
        if (fl & TF_END) {
            const uint32_t r = f2u(w.pos[s].y), m = r & 3u;
            const bool close = (r & REC_CLOSE) != 0u;
            w.adv[s] = nan;
            if (!(close && m > 2u) && m >= 2u) carry = adv;
            ++s;
            continue;
        }
        if (fl & TF_BEGIN) {
            adv = carry;
            w.adv[s] = adv;
            ++s;
            continue;
        }
#ifdef __CUDA_ARCH__
        if ((s & 3u) == 0u && s + 32u < sb) {  // 32 slots, none of them a Begin / End?
            uint32_t any = 0u;                 // (the flag bytes of a tile are 4-byte aligned, not more)
#pragma unroll
            for (int q = 0; q < 8; ++q) any |= *reinterpret_cast<const uint32_t*>(&w.flags[s + 4 * q]);
            if (!(any & (0x01010101u * (uint32_t)(TF_BEGIN | TF_END)))) {
                float4 l[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) l[q] = *reinterpret_cast<const float4*>(&len[s + 4 * q]);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    l[q].x = adv = adv + l[q].x;
                    l[q].y = adv = adv + l[q].y;
                    l[q].z = adv = adv + l[q].z;
                    l[q].w = adv = adv + l[q].w;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(&w.adv[s + 4 * q]) = l[q];
                s += 32u;
                continue;
            }
        }
        if ((s & 3u) == 0u && s + 4u < sb) {
            const uint32_t f4 = *reinterpret_cast<const uint32_t*>(&w.flags[s]);
            if (!(f4 & (0x01010101u * (uint32_t)(TF_BEGIN | TF_END)))) {
                float4 l = *reinterpret_cast<const float4*>(&len[s]);
                l.x = adv = adv + l.x;
                l.y = adv = adv + l.y;
                l.z = adv = adv + l.z;
                l.w = adv = adv + l.w;
                *reinterpret_cast<float4*>(&w.adv[s]) = l;
                s += 4u;
                continue;
            }
        }
#endif
        adv = adv + len[s];
        w.adv[s] = adv;
        ++s;
    }
}

}  // namespace lyc
