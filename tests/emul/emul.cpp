// emul.cpp -- CPU emulation of the CUDA pipeline, for the GPU-less build container.
// TEST INFRASTRUCTURE ONLY.  It runs the *same* per-item functions the kernels run (stroke_core.cuh,
// device_math.cuh compiled by g++ -ffp-contract=off) with sequential loops in place of the grid and plain
// prefix sums in place of the decoupled look-back scans, so that the device logic can be checked against
// the oracle without a GPU.  The product never links this; on a GPU box the -m gpu tests exercise the real
// kernels through the C ABI.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/lyon_cuda.h"
#include "../../lyon_b200/csrc/stroke_tile.cuh"
#include "../../lyon_b200/csrc/stroke_wtile.cuh"
#include "../../lyon_b200/csrc/stroke_vw.cuh"

using namespace lyc;

static StrokeParams make_params(const lyon_stroke_options& o, uint32_t num_attributes) {
    StrokeParams p{};
    p.start_cap = o.start_cap;
    p.end_cap = o.end_cap;
    p.line_join = o.line_join;
    p.line_width = o.line_width;
    p.half_width = o.line_width * 0.5f;
    p.miter_limit = o.miter_limit;
    p.tolerance = o.tolerance;
    p.merge_thr = fmaxf(fminf(o.tolerance * o.tolerance * 0.5f, o.line_width * o.line_width * 0.05f), 1e-8f);
    p.flag_thr = p.merge_thr * 5.0f;
    p.arc_step = circle_flattening_step(p.half_width, p.tolerance);  // host libm, same as the oracle
    p.num_attributes = num_attributes;
    p.attr_stride = (num_attributes + 1u) / 2u;
    p.vw_attr = num_attributes > 0 ? o.variable_line_width : -1;
    return p;
}

extern "C" int emul_stroke_batch(const lyon_path_batch* b, const lyon_stroke_options* o, uint64_t vertex_base,
                                 lyon_stroke_vertex** out_v, uint64_t* n_v, uint32_t** out_i, uint64_t* n_i,
                                 uint64_t* path_vertex_end, uint64_t* path_index_end) {
    *out_v = nullptr;
    *out_i = nullptr;
    *n_v = *n_i = 0;
    StrokeParams p = make_params(*o, b->num_attributes);
    const uint32_t NV = b->n_verbs;
    std::vector<uint32_t> seg_pt(NV + 1), seg_off(NV + 1), seg_id(NV + 1);
    std::vector<float> seg_tsplit(NV + 1);
    Views w{};
    w.verbs = b->verbs;
    w.points = reinterpret_cast<const V2*>(b->points);
    w.n_verbs = NV;
    w.seg_pt = seg_pt.data();
    w.seg_off = seg_off.data();
    w.seg_tsplit = seg_tsplit.data();
    w.seg_id = seg_id.data();
    // K1
    uint64_t pt = 0, so = 0;
    for (uint32_t v = 0; v < NV; ++v) {
        uint8_t verb = b->verbs[v];
        seg_pt[v] = (uint32_t)pt;
        seg_off[v] = (uint32_t)so;
        float ts;
        so += seg_slot_count(verb, w.points, (uint32_t)pt, p.attr_stride, p.tolerance, &ts);
        seg_tsplit[v] = ts;
        pt += verb_point_count(verb, p.attr_stride);
    }
    seg_pt[NV] = (uint32_t)pt;
    seg_off[NV] = (uint32_t)so;
    if (so >= 0xFFFFFFF0ull) return LYON_ERR_TOO_MANY_VERTICES;
    const uint32_t NS = (uint32_t)so;
    std::vector<V2> pos(NS + 1);
    std::vector<float> t(NS + 1), len(NS + 1), adv(NS + 1);
    std::vector<uint32_t> segfs(NS + 1), voff(NS + 1), toff(NS + 1);
    std::vector<uint8_t> jcode(NS + 1, 0);
    w.pos = pos.data();
    w.t = t.data();
    w.segfs = segfs.data();
    w.len = len.data();
    w.adv = adv.data();
    w.voff = voff.data();
    w.toff = toff.data();
    w.n_slots = NS;
    // K2
    for (uint32_t v = 0; v < NV; ++v)
        for (uint32_t s = seg_off[v]; s < seg_off[v + 1]; ++s) point_gen_slot(w, p, v, s, s - seg_off[v]);
    // K3
    for (uint32_t path = 0; path < b->n_paths; ++path) path_walk(w, p, b->path_verb_begin[path], b->path_verb_begin[path + 1]);
    // K4
    uint64_t tv = 0, tt = 0;
    for (uint32_t s = 0; s < NS; ++s) {
        CountSink sink;
        sink.nv = sink.nt = 0;
        sink.base = 0;
        stroke_slot(w, p, s, 0u, sink, nullptr, jcode.data());
        voff[s] = (uint32_t)tv;
        toff[s] = (uint32_t)tt;
        tv += sink.nv;
        tt += sink.nt;
    }
    voff[NS] = (uint32_t)tv;
    toff[NS] = (uint32_t)tt;
    if (tv > 0xFFFFFFFFull || tt > 0xFFFFFFFFull) return LYON_ERR_TOO_MANY_VERTICES;
    if (tv && vertex_base + tv - 1ull > 0xFFFFFFFEull) return LYON_ERR_TOO_MANY_VERTICES;
    // K5
    GpuVertex* V = (GpuVertex*)malloc((tv ? tv : 1) * sizeof(GpuVertex));
    uint32_t* I = (uint32_t*)malloc((tt ? tt : 1) * 12);
    int status = LYON_OK;
    for (uint32_t s = 0; s < NS; ++s) {
        uint32_t nv = voff[s + 1] - voff[s], nt = toff[s + 1] - toff[s];
        if ((nv | nt) == 0) continue;
        EmitSink sink;
        sink.vtx = V + voff[s];
        sink.idx = I + 3ull * toff[s];
        sink.nv = sink.nt = 0;
        sink.base = (uint32_t)vertex_base + voff[s];
        stroke_slot(w, p, s, (uint32_t)vertex_base, sink, jcode.data(), nullptr);
        if (sink.nv != nv || sink.nt != nt) status = LYON_ERR_INTERNAL;
    }
    for (uint32_t path = 0; path < b->n_paths; ++path) {
        uint32_t s_end = seg_off[b->path_verb_begin[path + 1]];
        if (path_vertex_end) path_vertex_end[path] = voff[s_end];
        if (path_index_end) path_index_end[path] = 3ull * toff[s_end];
    }
    *out_v = (lyon_stroke_vertex*)V;
    *out_i = I;
    *n_v = tv;
    *n_i = 3ull * tt;
    return status;
}
// The FUSED tile pipeline (tile_kernels.cuh) with the whole batch as one tile: the same per-item routines
// (stroke_tile.cuh + stroke_core.cuh against TileViews), sequential loops in place of the block.
extern "C" int emul_stroke_batch_tile(const lyon_path_batch* b, const lyon_stroke_options* o, uint64_t vertex_base,
                                      lyon_stroke_vertex** out_v, uint64_t* n_v, uint32_t** out_i, uint64_t* n_i,
                                      uint64_t* path_vertex_end, uint64_t* path_index_end) {
    *out_v = nullptr;
    *out_i = nullptr;
    *n_v = *n_i = 0;
    StrokeParams p = make_params(*o, b->num_attributes);
    const uint32_t NV = b->n_verbs;
    const V2* points = reinterpret_cast<const V2*>(b->points);
    // phase 1a: plans + offsets
    std::vector<VerbPlan> plans(NV);
    std::vector<uint32_t> seg_pt(NV + 1), seg_off(NV + 1);
    uint64_t pt = 0, so = 0;
    for (uint32_t v = 0; v < NV; ++v) {
        seg_pt[v] = (uint32_t)pt;
        seg_off[v] = (uint32_t)so;
        plans[v] = plan_verb(b->verbs[v], points, (uint32_t)pt, p.attr_stride, p.tolerance);
        so += plans[v].nslots;
        pt += verb_point_count(b->verbs[v], p.attr_stride);
    }
    seg_pt[NV] = (uint32_t)pt;
    seg_off[NV] = (uint32_t)so;
    if (so >= 0xFFFFFFF0ull) return LYON_ERR_TOO_MANY_VERTICES;
    const uint32_t NS = (uint32_t)so;
    std::vector<V2> pos(NS + 1);
    std::vector<float> t(NS + 1), adv(NS + 1);
    std::vector<uint32_t> src(NS + 1), voff(NS + 1), toff(NS + 1);
    std::vector<uint8_t> flags(NS + 1, TF_END), jcode(NS + 1, 0);
    TileViews w{};
    w.pos = pos.data();
    w.t = t.data();
    w.adv = adv.data();
    w.src = src.data();
    w.flags = flags.data();
    w.voff = voff.data();
    w.toff = toff.data();
    w.jcode = jcode.data();
    std::vector<uint32_t> dropmask(NS / 32 + 2, 0u);
    w.dropmask = dropmask.data();
    w.n_slots = NS;
    // phase 1a: slots of every verb (curves: accumulated t only); EndpointIds per path
    std::vector<uint32_t> verb_path(NV, 0), evt_before(NV + 1, 0);
    for (uint32_t path = 0; path < b->n_paths; ++path)
        for (uint32_t v = b->path_verb_begin[path]; v < b->path_verb_begin[path + 1]; ++v) verb_path[v] = path;
    for (uint32_t v = 0; v < NV; ++v)
        evt_before[v + 1] = evt_before[v] + ((b->verbs[v] != VERB_END && b->verbs[v] != VERB_CLOSE) ? 1u : 0u);
    uint32_t begin_slot = 0u;
    for (uint32_t v = 0; v < NV; ++v) {
        uint8_t verb = b->verbs[v];
        uint32_t ncurve = verb == VERB_QUAD ? 2u : (verb == VERB_CUBIC ? 3u : 1u);
        uint32_t vb = b->path_verb_begin[verb_path[v]];
        if (v == vb) begin_slot = seg_off[v];
        uint32_t id = p.num_attributes > 0 ? (seg_pt[v] - seg_pt[vb]) + ncurve - 1u : evt_before[v] - evt_before[vb];
        if (verb == VERB_QUAD && plans[v].tsplit >= 0.0f) {  // split quadratic: a lane per slot from the pre-pass's plan
            const SplitRec rec{plans[v].tsplit, plans[v].n_a, plans[v].n_b, 0.0f};
            for (uint32_t j = 0; j < plans[v].nslots; ++j) tile_gen_split_slot(w, p, points, seg_pt[v], seg_off[v], id, rec, j);
            continue;
        }
        if (verb == VERB_BEGIN) begin_slot = seg_off[v];
        tile_gen_verb_full(w, p, verb, points, seg_pt[v], plans[v].n_a, seg_off[v], id, begin_slot);
    }
    // phase 1c: the merge filter, run by run (on the device all runs at once, in any order: here odd starts first)
    for (uint32_t pass = 0; pass < 2u; ++pass)
        for (uint32_t s = 0; s < NS; ++s)
            if ((s & 1u) != pass && tile_run_start(w, s)) tile_resolve_run(w, p, s, NS);
    for (uint32_t s = 0; s < NS; ++s) tile_merge_dropped(w, s);
    // phase 3: classify + count
    std::vector<uint32_t> cnt_v(NS + 1, 0), cnt_t(NS + 1, 0);
    for (uint32_t s = 0; s < NS; ++s) {
        SlotClass c = classify_slot(w, s);
        uint32_t nv = 0, nt = 0;
        uint8_t code = 0;
        if (c.kind == KIND_JOIN) {
            if (join_is_fast(w.P(c.pk), w.P(s), w.P(c.nk), (c.sf & SEG_FS_BIT) != 0u)) {
                nv = 2u;
                nt = (w.F(c.pk) & SEG_BEGIN_BIT) ? 0u : 2u;
                code = JC_FAST | JC_SINGLE_POS | JC_SINGLE_NEG;
            } else {
                count_full_join(w, p, s, c.pk, c.nk, &nv, &nt, &code);
            }
        } else if (c.kind == KIND_END) {
            tile_end_record_finish(w, s);
            end_slot_count(w, p, s, &nv, &nt);
            uint32_t gv, gt;
            end_slot_count_generic(w, p, s, &gv, &gt);
            if (gv != nv || gt != nt) return LYON_ERR_INTERNAL;  // the cheap End count must agree with the real routine
        }
        cnt_v[s] = nv;
        cnt_t[s] = nt;
        jcode[s] = code;
    }
    uint64_t tv = 0, tt = 0;
    for (uint32_t s = 0; s < NS; ++s) {
        voff[s] = (uint32_t)tv;
        toff[s] = (uint32_t)tt;
        tv += cnt_v[s];
        tt += cnt_t[s];
    }
    voff[NS] = (uint32_t)tv;
    toff[NS] = (uint32_t)tt;
    if (tv > 0xFFFFFFFFull || tt > 0xFFFFFFFFull) return LYON_ERR_TOO_MANY_VERTICES;
    if (tv && vertex_base + tv - 1ull > 0xFFFFFFFEull) return LYON_ERR_TOO_MANY_VERTICES;
    // phase 2 (no count depends on it; the kernel runs it while the look-back waits): advancement chain, per path
    for (uint32_t path = 0; path < b->n_paths; ++path) {
        uint32_t vb = b->path_verb_begin[path], ve = b->path_verb_begin[path + 1];
        tile_path_chain(w, p, seg_off[vb], seg_off[ve]);
    }
    // phase 4: emit
    GpuVertex* V = (GpuVertex*)malloc((tv ? tv : 1) * sizeof(GpuVertex));
    uint32_t* I = (uint32_t*)malloc((tt ? tt : 1) * 12);
    int status = LYON_OK;
    for (uint32_t s = 0; s < NS; ++s) {
        uint32_t nv = cnt_v[s], nt = cnt_t[s];
        if ((nv | nt) == 0) continue;
        EmitSink sink;
        sink.vtx = V + voff[s];
        sink.idx = I + 3ull * toff[s];
        sink.nv = sink.nt = 0;
        sink.base = (uint32_t)vertex_base + voff[s];
        if (w.F(s) & SEG_END_BIT) {  // as the kernels do: an EndRecord, then stroke_end_slot on the record alone
            EndRecord rec;
            end_record_fill(w, s, &rec);
            const V2 sub = w.P(s);
            const uint32_t b0 = f2u(sub.x), m = f2u(sub.y) & 3u;
            uint32_t q1 = b0, q4 = b0;
            if (m >= 2u) {
                q1 = next_kept_or_end(w, b0);
                uint32_t l1 = s - 1u;
                while (slot_dropped(w, l1)) --l1;
                q4 = prev_kept(w, l1);
            }
            rec.vo1 = (uint32_t)vertex_base + w.VO(q1);
            rec.vo4 = (uint32_t)vertex_base + w.VO(q4);
            rec.first_id = sink.base;
            EndRecView rv;
            rv.r = &rec;
            rv.n_slots = 7u;
            stroke_end_slot(rv, p, 6u, 0u, sink);
        } else {
            uint32_t pk = prev_kept(w, s), nk = next_kept_or_end(w, s);
            if (jcode[s] & JC_FAST)
                emit_fast_join(w, p, s, pk, nk, (uint32_t)vertex_base, jcode.data(), sink);
            else
                emit_full_join(w, p, s, pk, nk, (uint32_t)vertex_base, jcode.data(), sink);
        }
        if (sink.nv != nv || sink.nt != nt) status = LYON_ERR_INTERNAL;
    }
    for (uint32_t path = 0; path < b->n_paths; ++path) {
        uint32_t s_end = seg_off[b->path_verb_begin[path + 1]];
        if (path_vertex_end) path_vertex_end[path] = voff[s_end];
        if (path_index_end) path_index_end[path] = 3ull * toff[s_end];
    }
    *out_v = (lyon_stroke_vertex*)V;
    *out_i = I;
    *n_v = tv;
    *n_i = 3ull * tt;
    return status;
}
// The WARP-TILE pipeline (wtile_kernels.cuh), one tile per path: the per-item routines of stroke_wtile.cuh against
// WTileViews -- slot -> verb lookup through the head mask, t / EndpointId re-derived per verb, 16-bit chunk-relative
// offsets, the cheap End-slot count (checked here against the real routine), End slots emitted from EndRecords.
extern "C" int emul_stroke_batch_wtile(const lyon_path_batch* b, const lyon_stroke_options* o, uint64_t vertex_base,
                                       lyon_stroke_vertex** out_v, uint64_t* n_v, uint32_t** out_i, uint64_t* n_i,
                                       uint64_t* path_vertex_end, uint64_t* path_index_end) {
    *out_v = nullptr;
    *out_i = nullptr;
    *n_v = *n_i = 0;
    StrokeParams p = make_params(*o, b->num_attributes);
    const V2* points = reinterpret_cast<const V2*>(b->points);
    std::vector<GpuVertex> V;
    std::vector<uint32_t> I;
    int status = LYON_OK;
    uint64_t V0 = 0, T0 = 0;
    std::vector<SplitRec> split_rec;
    for (uint32_t path = 0; path < b->n_paths; ++path) {
        const uint32_t va = b->path_verb_begin[path], vb = b->path_verb_begin[path + 1], nverbs = vb - va;
        const uint32_t pt_path = b->path_point_begin[path];
        // pre-pass + setup: per verb
        std::vector<uint16_t> vs0(nverbs + 1), vword;
        std::vector<uint32_t> vsrc(nverbs + 1), vptg(nverbs + 1);
        std::vector<float> vstep(nverbs + 1);
        std::vector<uint8_t> vkind(nverbs + 1);
        uint32_t pt = pt_path, so = 0, ev = 0;
        split_rec.clear();
        for (uint32_t v = va; v < vb; ++v) {
            const uint32_t vi = v - va;
            const uint8_t verb = b->verbs[v];
            const VerbPlan plan = plan_verb(verb, points, pt, p.attr_stride, p.tolerance);
            const bool is_ev = verb != VERB_END && verb != VERB_CLOSE;
            const uint32_t ncurve = verb == VERB_QUAD ? 2u : (verb == VERB_CUBIC ? 3u : 1u);
            vs0[vi] = (uint16_t)so;
            vptg[vi] = pt;
            vsrc[vi] = !is_ev ? 0u : (p.num_attributes > 0 ? (pt - pt_path) + ncurve - 1u : ev);
            vkind[vi] = verb;
            vstep[vi] = 0.0f;
            if (verb == VERB_QUAD || verb == VERB_CUBIC) {
                if (plan.tsplit >= 0.0f) {
                    vkind[vi] = (uint8_t)(verb | VK_SPLIT);
                    vstep[vi] = u2f((uint32_t)split_rec.size());
                    split_rec.push_back(SplitRec{plan.tsplit, plan.n_a, plan.n_b, 0.0f});
                } else {
                    vstep[vi] = 1.0f / plan.n_a;
                }
            }
            so += plan.nslots;
            pt += verb_point_count(verb, p.attr_stride);
            ev += is_ev ? 1u : 0u;
        }
        if (so > 60000u) return LYON_ERR_INTERNAL;  // the emulation keeps one path per warp tile
        const uint32_t NS = so;
        vs0[nverbs] = (uint16_t)NS;
        const uint32_t NW = NS / 32 + 2;
        std::vector<uint32_t> head(NW, 0), special(NW, 0), cbase_v(NW, 0), cbase_t(NW, 0), off(NS + 40, 0);
        vword.assign(NW, 0);
        for (uint32_t vi = 0; vi < nverbs; ++vi) head[vs0[vi] >> 5] |= 1u << (vs0[vi] & 31u);
        for (uint32_t i = 1; i < NW; ++i) vword[i] = (uint16_t)(vword[i - 1] + __builtin_popcount(head[i - 1]));
        std::vector<V2> pos(NS + 40);
        std::vector<float> adv(NS + 40);
        std::vector<uint8_t> flags(NS + 40, TF_END), jcode(NS + 40, 0);
        WTileViews w{};
        w.pos = pos.data();
        w.adv = adv.data();
        w.off = off.data();
        w.flags = flags.data();
        w.jcode = jcode.data();
        w.cbase_v = cbase_v.data();
        w.cbase_t = cbase_t.data();
        w.head = head.data();
        w.vword = vword.data();
        w.vs0 = vs0.data();
        w.vstep = vstep.data();
        w.vsrc = vsrc.data();
        w.vkind = vkind.data();
        w.split_rec = split_rec.data();
        w.n_slots = NS;
        // setup: split quadratics, a "lane" per slot
        for (uint32_t vi = 0; vi < nverbs; ++vi)
            if (vkind[vi] & VK_SPLIT)
                for (uint32_t k = 0; k < (uint32_t)(vs0[vi + 1] - vs0[vi]); ++k)
                    wt_gen_split_slot(w, p, points, vptg[vi], vs0[vi], split_rec[f2u(vstep[vi])], k);
        // gen
        std::vector<V2> from(NS + 1);
        for (uint32_t s = 0; s < NS; ++s) {
            const uint32_t vi = w.verb_of(s);
            if (vkind[vi] & VK_SPLIT) continue;
            wt_gen_slot(w, p, points, vkind[vi], vptg[vi], s - vs0[vi], (uint32_t)(vs0[vi + 1] - vs0[vi]), vstep[vi], s, &from[s]);
        }
        for (uint32_t s = 0; s < NS; ++s) {
            const uint32_t vi = w.verb_of(s);
            const uint8_t kind = vkind[vi];
            if (kind & VK_SPLIT) continue;
            float l = 0.0f;
            if (kind != VERB_BEGIN && kind != VERB_END && kind != VERB_CLOSE)
                l = tile_len_flagged(p, w.pos[s], s - vs0[vi] == 0u ? from[s] : w.pos[s - 1u]);
            w.len()[s] = l;
        }
        for (uint32_t s = 0; s < NS; ++s)
            if (tile_slot_special(w, s)) special[s >> 5] |= 1u << (s & 31u);
        for (uint32_t pass = 0; pass < 2u; ++pass)
            for (uint32_t s = 0; s < NS; ++s)
                if (((special[s >> 5] >> (s & 31u)) & 1u) && ((s & 1u) == pass) && (s & 31u) != 0u && tile_resolve_flagged(w, p, s))
                    special[s >> 5] &= ~(1u << (s & 31u));
        // walk
        if (NS) tile_path_walk(w, p, 0u, NS, special.data());
        // classify + count (packed 16 + 16 bit counts in `off`)
        std::vector<uint32_t> ends;
        for (uint32_t s = 0; s < NS; ++s) {
            const SlotClass c = classify_slot(w, s);
            uint32_t nv = 0, nt = 0;
            uint8_t code = 0;
            if (c.kind == KIND_JOIN) {
                if (join_is_fast(w.P(c.pk), w.P(s), w.P(c.nk), (c.sf & SEG_FS_BIT) != 0u)) {
                    nv = 2u;
                    nt = (w.F(c.pk) & SEG_BEGIN_BIT) ? 0u : 2u;
                    code = JC_FAST | JC_SINGLE_POS | JC_SINGLE_NEG;
                } else {
                    count_full_join(w, p, s, c.pk, c.nk, &nv, &nt, &code);
                }
            } else if (c.kind == KIND_END) {
                end_slot_count(w, p, s, &nv, &nt);
                uint32_t gv, gt;
                end_slot_count_generic(w, p, s, &gv, &gt);
                if (gv != nv || gt != nt) status = LYON_ERR_INTERNAL;  // the cheap End count must agree with the real routine
                ends.push_back(s);
            }
            if ((nv | nt) > 0xFFFFu) return LYON_ERR_INTERNAL;
            off[s] = nv | (nt << 16);
            jcode[s] = code;
        }
        // scan: chunk bases + 16-bit relative offsets
        uint32_t run_v = 0, run_t = 0;
        for (uint32_t c = 0; c <= NS; c += 32) {
            cbase_v[c >> 5] = run_v;
            cbase_t[c >> 5] = run_t;
            uint32_t ev_ = 0, et_ = 0;
            for (uint32_t s = c; s < c + 32 && s <= NS; ++s) {
                const uint32_t cnt = s < NS ? off[s] : 0u;
                if ((ev_ | et_) > 0xFFFFu) return LYON_ERR_INTERNAL;
                off[s] = ev_ | (et_ << 16);
                ev_ += cnt & 0xFFFFu;
                et_ += cnt >> 16;
            }
            run_v += ev_;
            run_t += et_;
        }
        const uint64_t tv = V0 + run_v, tt = T0 + run_t;
        if (tv > 0xFFFFFFFFull || tt > 0xFFFFFFFFull) return LYON_ERR_TOO_MANY_VERTICES;
        if (tv && vertex_base + tv - 1ull > 0xFFFFFFFEull) return LYON_ERR_TOO_MANY_VERTICES;
        V.resize(tv);
        I.resize(3 * tt);
        const uint32_t gbase = (uint32_t)vertex_base + (uint32_t)V0;
        GpuVertex* gv = V.data() + V0;
        uint32_t* gi = I.data() + 3 * T0;
        for (uint32_t s = 0; s < NS; ++s) {
            if (w.F(s) & SEG_END_BIT) continue;
            const uint32_t vo = w.VO(s), to = w.TO(s), nv = w.VO(s + 1) - vo, nt = w.TO(s + 1) - to;
            if ((nv | nt) == 0) continue;
            EmitSink sink;
            sink.vtx = gv + vo;
            sink.idx = gi + 3ull * to;
            sink.nv = sink.nt = 0;
            sink.base = gbase + vo;
            const uint32_t pk = prev_kept(w, s), nk = next_kept_or_end(w, s);
            if (jcode[s] & JC_FAST)
                emit_fast_join(w, p, s, pk, nk, gbase, jcode.data(), sink);
            else
                emit_full_join(w, p, s, pk, nk, gbase, jcode.data(), sink);
            if (sink.nv != nv || sink.nt != nt) status = LYON_ERR_INTERNAL;
        }
        // End slots: record, then the End kernel's routine on the record alone
        for (uint32_t s : ends) {
            const uint32_t vo = w.VO(s), to = w.TO(s), nv = w.VO(s + 1) - vo, nt = w.TO(s + 1) - to;
            if ((nv | nt) == 0) continue;
            EndRecord rec;
            end_record_fill(w, s, &rec);
            const V2 sub = w.P(s);
            const uint32_t b0 = f2u(sub.x), m = f2u(sub.y) & 3u;
            uint32_t q1 = b0, q4 = b0;
            if (m >= 2u) {
                q1 = next_kept_or_end(w, b0);
                uint32_t l1 = s - 1u;
                while (slot_dropped(w, l1)) --l1;
                q4 = prev_kept(w, l1);
            }
            rec.vo1 = gbase + w.VO(q1);
            rec.vo4 = gbase + w.VO(q4);
            rec.first_id = gbase + vo;
            rec.nv = nv;
            rec.nt = nt;
            rec.vtx_index = V0 + vo;
            rec.tri_index = T0 + to;
            EndRecView rv;
            rv.r = &rec;
            rv.n_slots = 7u;
            EmitSink sink;
            sink.vtx = V.data() + rec.vtx_index;
            sink.idx = I.data() + 3ull * rec.tri_index;
            sink.nv = sink.nt = 0;
            sink.base = rec.first_id;
            stroke_end_slot(rv, p, 6u, 0u, sink);
            if (sink.nv != rec.nv || sink.nt != rec.nt) status = LYON_ERR_INTERNAL;
        }
        V0 = tv;
        T0 = tt;
        if (path_vertex_end) path_vertex_end[path] = V0;
        if (path_index_end) path_index_end[path] = 3ull * T0;
    }
    GpuVertex* Vout = (GpuVertex*)malloc((V.size() ? V.size() : 1) * sizeof(GpuVertex));
    uint32_t* Iout = (uint32_t*)malloc((I.size() ? I.size() : 1) * 4);
    if (!V.empty()) memcpy(Vout, V.data(), V.size() * sizeof(GpuVertex));
    if (!I.empty()) memcpy(Iout, I.data(), I.size() * 4);
    *out_v = (lyon_stroke_vertex*)Vout;
    *out_i = Iout;
    *n_v = V.size();
    *n_i = I.size();
    return status;
}
extern "C" void emul_free(void* p) { free(p); }
extern "C" uint32_t emul_round_log2_segments(float n) { return round_log2_segments(n); }
extern "C" float emul_angle_to(float a, float b) { return angle_to(a, b); }

// Flatten a batch of standalone curves with the device routines (k_flat_count / k_flat_emit logic).
extern "C" uint64_t emul_flatten(const float* ctrl, uint64_t n, int npts, float tol, uint64_t* offsets, float* points, float* t_out) {
    uint64_t off = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const float* c = ctrl + 2 * npts * i;
        uint32_t cnt = npts == 4 ? to_u32_or_1(cubic_wang(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), v2(c[6], c[7]), tol))
                                 : to_u32_or_1(quad_wang(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), tol));
        offsets[i] = off;
        if (points) {
            for (uint32_t k = 0; k < cnt; ++k) {
                V2 pos, prev;
                float t;
                if (npts == 4)
                    cubic_point(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), v2(c[6], c[7]), tol, k, &pos, &prev, &t);
                else
                    quad_point(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), tol, k, &pos, &prev, &t);
                points[2 * (off + k)] = pos.x;
                points[2 * (off + k) + 1] = pos.y;
                if (t_out) t_out[off + k] = t;
            }
        }
        off += cnt;
    }
    offsets[n] = off;
    return off;
}

// Variable line width (stroke_vw.cuh): the per-path state machine the k_vw_count / k_vw_emit threads run, paths in a loop.
extern "C" int emul_stroke_batch_vw(const lyon_path_batch* b, const lyon_stroke_options* o, uint64_t vertex_base,
                                    lyon_stroke_vertex** out_v, uint64_t* n_v, uint32_t** out_i, uint64_t* n_i,
                                    uint64_t* path_vertex_end, uint64_t* path_index_end) {
    *out_v = nullptr;
    *out_i = nullptr;
    *n_v = *n_i = 0;
    StrokeParams p = make_params(*o, b->num_attributes);
    const V2* points = reinterpret_cast<const V2*>(b->points);
    std::vector<uint64_t> v0(b->n_paths + 1, 0), t0(b->n_paths + 1, 0);
    for (uint32_t path = 0; path < b->n_paths; ++path) {
        CountSink sink;
        sink.nv = sink.nt = 0;
        sink.base = 0;
        vw_stroke_path(b->verbs, b->path_verb_begin[path], b->path_verb_begin[path + 1], points, b->path_point_begin[path], p, sink);
        v0[path + 1] = v0[path] + sink.nv;
        t0[path + 1] = t0[path] + sink.nt;
    }
    const uint64_t tv = v0[b->n_paths], tt = t0[b->n_paths];
    if (tv > 0xFFFFFFFFull || tt > 0xFFFFFFFFull) return LYON_ERR_TOO_MANY_VERTICES;
    if (tv && vertex_base + tv - 1ull > 0xFFFFFFFEull) return LYON_ERR_TOO_MANY_VERTICES;
    GpuVertex* V = (GpuVertex*)malloc((tv ? tv : 1) * sizeof(GpuVertex));
    uint32_t* I = (uint32_t*)malloc((tt ? tt : 1) * 12);
    int status = LYON_OK;
    for (uint32_t path = 0; path < b->n_paths; ++path) {
        EmitSink sink;
        sink.vtx = V + v0[path];
        sink.idx = I + 3ull * t0[path];
        sink.nv = sink.nt = 0;
        sink.base = (uint32_t)vertex_base + (uint32_t)v0[path];
        vw_stroke_path(b->verbs, b->path_verb_begin[path], b->path_verb_begin[path + 1], points, b->path_point_begin[path], p, sink);
        if (sink.nv != v0[path + 1] - v0[path] || sink.nt != t0[path + 1] - t0[path]) status = LYON_ERR_INTERNAL;
        if (path_vertex_end) path_vertex_end[path] = v0[path + 1];
        if (path_index_end) path_index_end[path] = 3ull * t0[path + 1];
    }
    *out_v = (lyon_stroke_vertex*)V;
    *out_i = I;
    *n_v = tv;
    *n_i = 3ull * tt;
    return status;
}
