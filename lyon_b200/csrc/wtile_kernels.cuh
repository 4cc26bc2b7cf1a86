// wtile_kernels.cuh -- the WARP-TILE stroke pipeline.  Included by lyon_cuda.cu after tile_kernels.cuh (same anonymous
// namespace: it reuses TileArgs, the TMA helpers, the look-back and k_path_slots / k_scan_paths / k_tile_pack).
//
//   k_stroke_wtile  every WARP owns a tile of whole paths in its slice of shared memory and runs the phases below
//                   warp-synchronously (no __syncthreads, no shared counters); tiles are taken in ticket order.
//        0  stage    verbs, Wang counts, points of the tile: TMA bulk copies onto a per-warp mbarrier
//        1  setup    lane / verb: point + slot offsets by shuffle scans, EndpointIds, 1 / Wang count, verb heads;
//                    split quadratics (flatten_quad) generated right here, a lane per slot, from the pre-pass's plan
//        2  gen      lane / slot: slot -> verb by popc over the head mask, t replayed, Horner; edge length; "special" mask
//        3  walk     lane / path: merge filter + advancement chain (sequential f32, stroke.rs:1530-1550)
//        4  count    lane / slot classify; fast joins counted on the spot, other joins compacted (ballot) and counted
//                    densely; End slots counted by end_slot_count
//        5  scan     lane / slot: 16-bit offsets relative to per-chunk bases; tile totals
//        6  base     decoupled look-back over tile aggregates (COUNT / FUSED) or read back (EMIT)
//        7  emit     fast joins lane / slot, other joins lane / list entry: 32-byte vertex stores straight to HBM;
//                    End slots leave as EndRecords
//   k_stroke_ends   thread / EndRecord: everything end() emits (caps, closing joins), dense
//
// Shared memory per warp: 20 B per slot + 13 B per verb (+ 8 B per point beyond 3 B per slot), see wt_layout.
#pragma once

struct WtArgs {
    uint32_t s_cap, v_cap, p_cap;  // per-warp tile capacities: slots, verbs, points
    uint32_t warp_bytes;           // shared memory per warp
};
enum { TILE_ERR_WT_OVERFLOW = 8u };  // a count or chunk-relative offset left 16 bits: the batch is redone by the CTA-tile pipeline

struct WtLayout {
    uint32_t pos, adv, off, list, flags, jcode, cbase_v, cbase_t, special, head, vword, vs0, vpt, vsrc, vstep, vkind, pts, mbar, total;
};
__host__ __device__ inline uint32_t wt_up16(uint32_t x) { return (x + 15u) & ~15u; }
__host__ __device__ inline WtLayout wt_layout(uint32_t s_cap, uint32_t v_cap, uint32_t p_cap) {
    const uint32_t S1 = s_cap + 36u;          // + sentinel slot, + a chunk of slack for the scan's last iteration
    const uint32_t W1 = S1 / 32u + 2u;        // 32-slot words / chunks
    const uint32_t V1 = v_cap + 20u;          // + sentinel, + 16-byte alignment slack of the staged range
    const uint32_t P1 = p_cap + 4u;
    WtLayout L;
    uint32_t o = 0;
    L.pos = o;     o = wt_up16(o + S1 * 8u);
    L.adv = o;     o = wt_up16(o + S1 * 4u);
    L.off = o;     o = wt_up16(o + S1 * 4u);
    L.flags = o;   o = wt_up16(o + S1);
    L.list = o;    o = wt_up16(o + S1 * 2u);  // list + jcode are first written by the count phase: until then the
    L.jcode = o;   o = wt_up16(o + S1);       // staged points live there (when they fit)
    const uint32_t late = o - L.list;
    L.cbase_v = o; o = wt_up16(o + W1 * 4u);
    L.cbase_t = o; o = wt_up16(o + W1 * 4u);
    L.special = o; o = wt_up16(o + W1 * 4u);
    L.head = o;    o = wt_up16(o + W1 * 4u);
    L.vword = o;   o = wt_up16(o + W1 * 2u);
    L.vs0 = o;     o = wt_up16(o + V1 * 2u);
    L.vpt = o;     o = wt_up16(o + V1 * 2u);
    L.vsrc = o;    o = wt_up16(o + V1 * 4u);
    L.vstep = o;   o = wt_up16(o + V1 * 4u);
    L.vkind = o;   o = wt_up16(o + V1);
    if (P1 * 8u <= late) {
        L.pts = L.list;
    } else {
        L.pts = o;
        o = wt_up16(o + P1 * 8u);
    }
    L.mbar = o;    o = wt_up16(o + 8u);
    L.total = (o + 127u) & ~127u;
    return L;
}

// One staged input range of a warp's tile: elements [lo, hi) of `src` land at dst[0 .. hi - a0), a0 = lo rounded down to
// a 16-byte boundary; the aligned bulk [a0, a0 + bulk) is one TMA copy (issued by lane 0), the ragged tail is copied by
// the lanes.  A source that is not 16-byte aligned is copied by the lanes entirely.
struct WtStage {
    uint32_t a0, bulk;
};
template <uint32_t ESZ>
__device__ __forceinline__ WtStage wt_stage_plan(const void* src, uint32_t lo, uint32_t hi) {
    constexpr uint32_t PER16 = 16u / ESZ;
    WtStage r;
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0u;
    r.a0 = aligned ? (lo & ~(PER16 - 1u)) : lo;
    r.bulk = aligned ? ((hi - r.a0) & ~(PER16 - 1u)) : 0u;
    return r;
}
template <class T>
__device__ __forceinline__ void wt_stage_tail(T* dst, const T* src, WtStage st, uint32_t hi, uint32_t lane) {
    for (uint32_t i = st.a0 + st.bulk + lane; i < hi; i += 32u) dst[i - st.a0] = src[i];
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t x, unsigned lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    return x;
}

#ifdef LYON_PHASE_TIMING
#define WPT(i)                                              \
    do {                                                    \
        __syncwarp();                                       \
        const long long _n = clock64();                     \
        pt_acc[i] += (unsigned long long)(_n - pt_last);    \
        pt_last = _n;                                       \
    } while (0)
#else
#define WPT(i)
#endif

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_stroke_wtile(TileArgs A, WtArgs X, StrokeParams p) {
    extern __shared__ __align__(128) unsigned char wt_smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    unsigned char* const base = wt_smem + (size_t)(threadIdx.x >> 5) * X.warp_bytes;
    const WtLayout L = wt_layout(X.s_cap, X.v_cap, X.p_cap);
    uint16_t* const list = reinterpret_cast<uint16_t*>(base + L.list);
    uint32_t* const special = reinterpret_cast<uint32_t*>(base + L.special);
    uint32_t* const head = reinterpret_cast<uint32_t*>(base + L.head);
    uint16_t* const vword = reinterpret_cast<uint16_t*>(base + L.vword);
    uint16_t* const vs0 = reinterpret_cast<uint16_t*>(base + L.vs0);
    uint16_t* const vpt = reinterpret_cast<uint16_t*>(base + L.vpt);
    uint32_t* const vsrc = reinterpret_cast<uint32_t*>(base + L.vsrc);
    float* const vstep_raw = reinterpret_cast<float*>(base + L.vstep);
    uint8_t* const vkind_raw = base + L.vkind;
    V2* const pts_raw = reinterpret_cast<V2*>(base + L.pts);
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(base + L.mbar);
    WTileViews w;
    w.pos = reinterpret_cast<V2*>(base + L.pos);
    w.adv = reinterpret_cast<float*>(base + L.adv);
    w.off = reinterpret_cast<uint32_t*>(base + L.off);
    w.flags = base + L.flags;
    w.jcode = base + L.jcode;
    w.cbase_v = reinterpret_cast<uint32_t*>(base + L.cbase_v);
    w.cbase_t = reinterpret_cast<uint32_t*>(base + L.cbase_t);
    w.head = head;
    w.vword = vword;
    w.vs0 = vs0;
    w.vsrc = vsrc;
    w.split_rec = A.split_rec;
    const bool with_ids = p.num_attributes > 0;
    uint32_t parity = 0u;
    if (lane == 0u) mbar_init(mbar, 1u);
    __syncwarp();
#ifdef LYON_PHASE_TIMING
    unsigned long long pt_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long pt_last = clock64();
#endif

    for (;;) {
        uint32_t tile = 0u;
        if (lane == 0u) tile = atomicAdd(A.ticket, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= A.n_tiles) break;
        // ---- tile header: path range, slot / verb / point ranges (two dependent round trips) ----------------------------
        uint32_t pa = 0u;
        if (lane < 2u) pa = A.tile_first[tile + lane];
        const uint32_t pb = __shfl_sync(FULL, pa, 1);
        pa = __shfl_sync(FULL, pa, 0);
        const uint32_t npaths = pb - pa;
        unsigned long long sl = 0ull;
        uint32_t hv = 0u;
        if (lane < 2u) sl = A.path_slot_begin[lane ? pb : pa];
        else if (lane < 4u) hv = A.path_verb_begin[lane == 3u ? pb : pa];
        else if (lane < 6u) hv = A.path_point_begin[lane == 5u ? pb : pa];
        const unsigned long long slot0 = __shfl_sync(FULL, sl, 0);
        const unsigned long long nslots64 = __shfl_sync(FULL, sl, 1) - slot0;
        const uint32_t va = __shfl_sync(FULL, hv, 2), vb = __shfl_sync(FULL, hv, 3);
        const uint32_t pt0 = __shfl_sync(FULL, hv, 4), pt1 = __shfl_sync(FULL, hv, 5);
        const uint32_t back = 1u + p.attr_stride;
        const uint32_t ptlo = pt0 >= back ? pt0 - back : 0u;  // seg_from_point may reach one endpoint back
        bool bad = nslots64 > (unsigned long long)X.s_cap || vb - va > X.v_cap || pt1 - ptlo + 2u > X.p_cap || vb < va || pt1 < pt0;
        const uint32_t nslots = bad ? 0u : (uint32_t)nslots64;
        const uint32_t nverbs = bad ? 0u : vb - va;
        const uint32_t nwords = (nslots + 32u) >> 5;  // words covering slots 0 .. nslots (the sentinel included)
        w.n_slots = nslots;
        WPT(0);

        // ---- phase 0: stage the tile's inputs ----------------------------------------------------------------------------
        const uint8_t* vkind = vkind_raw;  // indexed by GLOBAL verb / point index from here on
        float* vstep = vstep_raw;
        const V2* pts = pts_raw;
        if (nverbs) {
            const WtStage sk = wt_stage_plan<1>(A.verbs, va, vb), sn = wt_stage_plan<4>(A.verb_n, va, vb),
                          sp = wt_stage_plan<8>(A.points, ptlo, pt1);
            const uint32_t tx = sk.bulk + sn.bulk * 4u + sp.bulk * 8u;
            if (lane == 0u && tx) {
                // the staging regions were last touched through the generic proxy (previous tile)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(mbar, tx);
                if (sk.bulk) tma_bulk_g2s(vkind_raw, A.verbs + sk.a0, sk.bulk, mbar);
                if (sn.bulk) tma_bulk_g2s(vstep_raw, A.verb_n + sn.a0, sn.bulk * 4u, mbar);
                if (sp.bulk) tma_bulk_g2s(pts_raw, A.points + sp.a0, sp.bulk * 8u, mbar);
            }
            wt_stage_tail(vkind_raw, A.verbs, sk, vb, lane);
            wt_stage_tail(vstep_raw, A.verb_n, sn, vb, lane);
            wt_stage_tail(pts_raw, A.points, sp, pt1, lane);
            vkind = vkind_raw - sk.a0;
            vstep = vstep_raw - sn.a0;
            pts = pts_raw - sp.a0;
            if (tx) {  // uniform: every lane waits for the bytes to land
                mbar_wait(mbar, parity);
                parity ^= 1u;
            }
            __syncwarp();
        }
        uint8_t* const vkind_w = const_cast<uint8_t*>(vkind);
        w.vkind = vkind + va;
        w.vstep = vstep + va;
        for (uint32_t i = lane; i < nwords + 1u; i += 32u) head[i] = 0u;
        __syncwarp();
        WPT(1);

        // ---- phase 1: setup, lane / verb ---------------------------------------------------------------------------------
        bool inconsistent = false;
        for (uint32_t i0 = 0u; i0 < npaths && !bad; i0 += 32u) {
            // path table entries of up to 32 paths at once
            const uint32_t ip = i0 + lane;
            uint32_t t_vb = 0u, t_ve = 0u, t_pt = 0u, t_sb = 0u, t_se = 0u, t_sp = 0u;
            if (ip < npaths) {
                t_vb = A.path_verb_begin[pa + ip];
                t_ve = A.path_verb_begin[pa + ip + 1u];
                t_pt = A.path_point_begin[pa + ip];
                t_sb = (uint32_t)(A.path_slot_begin[pa + ip] - slot0);
                t_se = (uint32_t)(A.path_slot_begin[pa + ip + 1u] - slot0);
                t_sp = A.path_split_begin[pa + ip];
            }
            const uint32_t nhere = min(32u, npaths - i0);
            for (uint32_t j = 0u; j < nhere; ++j) {
                const uint32_t pvb = __shfl_sync(FULL, t_vb, j), pve = __shfl_sync(FULL, t_ve, j);
                const uint32_t pt_path = __shfl_sync(FULL, t_pt, j);
                uint32_t slot_run = __shfl_sync(FULL, t_sb, j);
                const uint32_t slot_end = __shfl_sync(FULL, t_se, j);
                uint32_t split_run = __shfl_sync(FULL, t_sp, j);
                uint32_t pt_run = pt_path, ev_run = 0u;
                for (uint32_t v0 = pvb; v0 < pve; v0 += 32u) {
                    const uint32_t v = v0 + lane;
                    const bool valid = v < pve;
                    const uint8_t kind = valid ? vkind[v] : (uint8_t)VERB_END;
                    const uint32_t npts = valid ? verb_point_count(kind, p.attr_stride) : 0u;
                    const bool is_ev = valid && kind != VERB_END && kind != VERB_CLOSE;
                    const bool curve = valid && (kind == VERB_QUAD || kind == VERB_CUBIC);
                    const uint32_t inc_p = warp_inclusive_scan(npts, lane);
                    const unsigned evm = __ballot_sync(FULL, is_ev);
                    const uint32_t pt = pt_run + inc_p - npts;
                    const uint32_t ev = ev_run + (uint32_t)__popc(evm & lt);
                    const float raw = curve ? vstep[v] : 1.0f;
                    const bool split = curve && raw < 0.0f && raw > -16777216.0f;  // -(slots) of a split quadratic
                    uint32_t n = 0u;
                    if (valid) n = split ? (uint32_t)(-raw) : (curve ? to_u32_or_1(raw) : 1u);
                    if (n > X.s_cap) n = X.s_cap + 1u;  // keeps the warp sum inside u32
                    const uint32_t inc_s = warp_inclusive_scan(n, lane);
                    const uint32_t s0 = slot_run + inc_s - n;
                    const unsigned spm = __ballot_sync(FULL, split);
                    const uint32_t sidx = split_run + (uint32_t)__popc(spm & lt);
                    const uint32_t ncurve = kind == VERB_QUAD ? 2u : (kind == VERB_CUBIC ? 3u : 1u);
                    // EndpointId: offset in the path's point array (path.rs:1197-1263) or sequential per event (stroke.rs:882-937)
                    const uint32_t id = !is_ev ? 0u : (with_ids ? (pt - pt_path) + ncurve - 1u : ev);
                    if (valid) {
                        if (s0 + n <= slot_end && n >= 1u) {
                            const uint32_t vi = v - va;
                            vs0[vi] = (uint16_t)s0;
                            vpt[vi] = (uint16_t)(pt - ptlo);
                            vsrc[vi] = id;
                            if (curve) vstep[v] = split ? u2f(sidx) : 1.0f / raw;
                            if (split) vkind_w[v] = (uint8_t)(kind | VK_SPLIT);
                            atomicOr(&head[s0 >> 5], 1u << (s0 & 31u));
                        } else {
                            inconsistent = true;
                        }
                    }
                    // split quadratics of this round: the whole warp generates one after the other, a lane per slot
                    unsigned todo = __ballot_sync(FULL, split && !inconsistent);
                    while (todo) {
                        const int src_lane = __ffs((int)todo) - 1;
                        todo &= todo - 1u;
                        const uint32_t q_pt = __shfl_sync(FULL, pt, src_lane), q_s0 = __shfl_sync(FULL, s0, src_lane);
                        const uint32_t q_n = __shfl_sync(FULL, n, src_lane), q_idx = __shfl_sync(FULL, sidx, src_lane);
                        const float4 rr = *reinterpret_cast<const float4*>(&A.split_rec[q_idx]);
                        SplitRec r;
                        r.tsplit = rr.x;
                        r.n_a = rr.y;
                        r.n_b = rr.z;
                        if (to_u32_or_1(r.n_a) + to_u32_or_1(r.n_b) != q_n) {
                            inconsistent = true;
                        } else {
                            for (uint32_t k = lane; k < q_n; k += 32u) wt_gen_split_slot(w, p, pts, q_pt, q_s0, r, k);
                        }
                    }
                    pt_run += __shfl_sync(FULL, inc_p, 31);
                    slot_run += __shfl_sync(FULL, inc_s, 31);
                    ev_run += (uint32_t)__popc(evm);
                    split_run += (uint32_t)__popc(spm);
                }
                if (lane == 0u && slot_run != slot_end) inconsistent = true;
            }
        }
        if (__any_sync(FULL, inconsistent)) {
            if (lane == 0u) atomicOr(&A.totals->error, TILE_ERR_INCONSISTENT);
            bad = true;
        }
        if (bad && lane == 0u && !(nslots64 == 0ull)) atomicOr(&A.totals->error, TILE_ERR_INCONSISTENT);
        __syncwarp();
        // verbs starting before each 32-slot word; the sentinel verb
        if (!bad) {
            uint32_t run = 0u;
            for (uint32_t i0 = 0u; i0 < nwords; i0 += 32u) {
                const uint32_t i = i0 + lane;
                const uint32_t c = i < nwords ? (uint32_t)__popc(head[i]) : 0u;
                const uint32_t inc = warp_inclusive_scan(c, lane);
                if (i < nwords) vword[i] = (uint16_t)(run + inc - c);
                run += __shfl_sync(FULL, inc, 31);
            }
            if (lane == 0u) vs0[nverbs] = (uint16_t)nslots;
            if (run != nverbs) bad = true;  // uniform
        }
        __syncwarp();
        WPT(2);

        unsigned long long agg_v = 0ull, agg_t = 0ull;
        uint32_t n_full = 0u, n_end = 0u;
        if (!bad && nslots) {
            // ---- phase 2: gen, lane / slot -------------------------------------------------------------------------------
            float* const len = w.len();
            for (uint32_t c = 0u; c < nslots; c += 32u) {
                const uint32_t s = c + lane;
                const bool in = s < nslots;
                uint32_t vi = 0u, k = 0u;
                uint8_t kind = VERB_END;
                bool ordinary = false;
                V2 from = v2(0.0f, 0.0f);
                if (in) {
                    vi = w.verb_of(s);
                    kind = w.vkind[vi];
                    ordinary = !(kind & VK_SPLIT);
                    k = s - (uint32_t)vs0[vi];
                    if (ordinary)
                        wt_gen_slot(w, p, pts, kind, ptlo + (uint32_t)vpt[vi], k, (uint32_t)vs0[vi + 1u] - (uint32_t)vs0[vi],
                                    w.vstep[vi], s, &from);
                }
                __syncwarp();
                bool sp = false;
                if (in) {
                    if (ordinary) {
                        float l = 0.0f;
                        if (kind != VERB_BEGIN && kind != VERB_END && kind != VERB_CLOSE)
                            l = tile_len_flagged(p, w.pos[s], k == 0u ? from : w.pos[s - 1u]);
                        len[s] = l;
                    }
                }
                __syncwarp();
                if (in) {
                    sp = tile_slot_special(w, s);
                    // a flagged slot next to a certainly-kept predecessor and beyond the merge distance is an ordinary point
                    // after all (the predecessor's slot was written by the neighbouring lane; lane 0 stays with the walk)
                    if (sp && lane != 0u && tile_resolve_flagged(w, p, s)) sp = false;
                }
                const unsigned bits = __ballot_sync(FULL, sp);
                if (lane == 0u) special[c >> 5] = bits;
            }
            if (lane == 0u) special[nslots >> 5] = (nslots & 31u) ? special[nslots >> 5] : 0u;
            __syncwarp();
            WPT(3);
            // ---- phase 3: walk, lane / path ------------------------------------------------------------------------------
            for (uint32_t i = lane; i < npaths; i += 32u) {
                const uint32_t sa = (uint32_t)(A.path_slot_begin[pa + i] - slot0);
                const uint32_t sb = (uint32_t)(A.path_slot_begin[pa + i + 1u] - slot0);
                if (sa < sb) tile_path_walk(w, p, sa, sb, special);
            }
            __syncwarp();
            WPT(4);
            // ---- phase 4: classify + count.  Fast joins on the spot; the other joins are compacted at the front of `list`,
            // End slots at its back, and counted densely afterwards.
            bool overflow = false;
            for (uint32_t c = 0u; c < nslots; c += 32u) {
                const uint32_t s = c + lane;
                uint32_t cnt = 0u;
                uint8_t code = 0;
                bool is_full = false, is_end = false;
                if (s < nslots) {
                    const SlotClass k = classify_slot(w, s);
                    if (k.kind == KIND_JOIN) {
                        if (join_is_fast(w.P(k.pk), w.P(s), w.P(k.nk), (k.sf & SEG_FS_BIT) != 0u)) {
                            // flattened_step: Positive, Negative (stroke.rs:1791-1797), two edge triangles unless the
                            // previous point is the sub-path's first
                            cnt = 2u | ((w.F(k.pk) & SEG_BEGIN_BIT) ? 0u : (2u << 16));
                            code = JC_FAST | JC_SINGLE_POS | JC_SINGLE_NEG;
                        } else {
                            is_full = true;
                        }
                    } else if (k.kind == KIND_END) {
                        is_end = true;
                    }
                    w.off[s] = cnt;  // aliased `len` until the walk finished
                    w.jcode[s] = code;
                }
                const unsigned mf = __ballot_sync(FULL, is_full), me = __ballot_sync(FULL, is_end);
                if (is_full) list[n_full + (uint32_t)__popc(mf & lt)] = (uint16_t)s;
                if (is_end) list[X.s_cap - 1u - (n_end + (uint32_t)__popc(me & lt))] = (uint16_t)s;
                n_full += (uint32_t)__popc(mf);
                n_end += (uint32_t)__popc(me);
            }
            __syncwarp();
            WPT(5);
            for (uint32_t r = lane; r < n_full; r += 32u) {
                const uint32_t s = list[r];
                const uint32_t pk = prev_kept(w, s), nk = next_kept_or_end(w, s);
                uint32_t nv, nt;
                uint8_t code;
                count_full_join(w, p, s, pk, nk, &nv, &nt, &code);
                overflow |= (nv | nt) > 0xFFFFu;
                w.off[s] = (nv & 0xFFFFu) | (nt << 16);
                w.jcode[s] = code;
            }
            for (uint32_t r = lane; r < n_end; r += 32u) {
                const uint32_t s = list[X.s_cap - 1u - r];
                uint32_t nv, nt;
                end_slot_count(w, p, s, &nv, &nt);
                overflow |= (nv | nt) > 0xFFFFu;
                w.off[s] = (nv & 0xFFFFu) | (nt << 16);
            }
            __syncwarp();
            WPT(6);
            // ---- phase 5: scan.  Per chunk of 32 slots: base = the tile's vertices / triangles before the chunk, per slot the
            // 16-bit offsets relative to it.  Slot `nslots` (the sentinel) gets the totals.
            {
                uint32_t run_v = 0u, run_t = 0u;
                for (uint32_t c = 0u; c <= nslots; c += 32u) {
                    const uint32_t s = c + lane;
                    const uint32_t cnt = s < nslots ? w.off[s] : 0u;
                    const uint32_t cv = cnt & 0xFFFFu, ct = cnt >> 16;
                    const uint32_t iv = warp_inclusive_scan(cv, lane), it = warp_inclusive_scan(ct, lane);
                    const uint32_t ev = iv - cv, et = it - ct;
                    overflow |= (ev | et) > 0xFFFFu;
                    if (s <= nslots) w.off[s] = (ev & 0xFFFFu) | (et << 16);
                    if (lane == 0u) {
                        w.cbase_v[c >> 5] = run_v;
                        w.cbase_t[c >> 5] = run_t;
                    }
                    run_v += __shfl_sync(FULL, iv, 31);
                    run_t += __shfl_sync(FULL, it, 31);
                }
                agg_v = run_v;
                agg_t = run_t;
            }
            if (__any_sync(FULL, overflow)) {
                if (lane == 0u) atomicOr(&A.totals->error, TILE_ERR_WT_OVERFLOW);
                bad = true;
                agg_v = agg_t = 0ull;
            }
            __syncwarp();
        }
        WPT(7);
        // ---- phase 6: the tile's global base --------------------------------------------------------------------------
        unsigned long long V0, T0;
        if (MODE == TILE_MODE_EMIT) {
            V0 = A.tile_base_v[tile];
            T0 = A.tile_base_t[tile];
        } else {
            uint64_t ev, et;
            lookback_exclusive_pair(A.desc_v, A.desc_t, tile, agg_v, agg_t, &ev, &et);
            V0 = ev;
            T0 = et;
            if (lane == 0u) {
                if (MODE == TILE_MODE_COUNT) {
                    A.tile_base_v[tile] = ev;
                    A.tile_base_t[tile] = et;
                }
                if (tile == A.n_tiles - 1u) {
                    A.totals->vertices = ev + agg_v;
                    A.totals->triangles = et + agg_t;
                }
            }
        }
        WPT(8);
        if (A.path_vertex_end || A.path_index_end) {
            for (uint32_t i = lane; i < npaths; i += 32u) {
                const uint32_t sb = (bad || !nslots) ? 0u : (uint32_t)(A.path_slot_begin[pa + i + 1u] - slot0);
                const unsigned long long ve = V0 + ((bad || !nslots) ? 0u : w.VO(sb)), te = T0 + ((bad || !nslots) ? 0u : w.TO(sb));
                if (A.path_vertex_end) A.path_vertex_end[pa + i] = ve;
                if (A.path_index_end) A.path_index_end[pa + i] = 3ull * te;
            }
        }
        // ---- phase 7: emit ----------------------------------------------------------------------------------------------
        bool do_emit = MODE != TILE_MODE_COUNT && !bad && nslots && (agg_v | agg_t) != 0ull;
        if (MODE == TILE_MODE_FUSED && do_emit) {
            if (V0 + agg_v > A.cap_v || T0 + agg_t > A.cap_t || A.vertex_base + V0 + agg_v - 1ull > 0xFFFFFFFEull) {
                do_emit = false;
                if (lane == 0u) atomicOr(&A.totals->overflow, 1u);
            }
        }
        if (do_emit) {
            const uint32_t gbase = A.vertex_base + (uint32_t)V0;  // global id of the tile's first vertex
            GpuVertex* const gv = A.vertices + V0;
            uint32_t* const gi = A.indices + 3ull * T0;
            bool mismatch = false;
            for (uint32_t c = 0u; c < nslots; c += 32u) {
                const uint32_t s = c + lane;
                if (s >= nslots || !(w.jcode[s] & JC_FAST)) continue;
                const uint32_t vo = w.VO(s), to = w.TO(s);
                EmitSinkWide sink;  // one 256-bit store per vertex
                sink.vtx = gv + vo;
                sink.idx = gi + 3ull * to;
                sink.nv = sink.nt = 0u;
                sink.base = gbase + vo;
                const uint32_t pk = prev_kept(w, s), nk = next_kept_or_end(w, s);
                emit_fast_join(w, p, s, pk, nk, gbase, w.jcode, sink);
                mismatch |= sink.nv != w.VO(s + 1u) - vo || sink.nt != w.TO(s + 1u) - to;
            }
            WPT(9);
            for (uint32_t r = lane; r < n_full; r += 32u) {
                const uint32_t s = list[r];
                const uint32_t vo = w.VO(s), to = w.TO(s);
                const uint32_t nv = w.VO(s + 1u) - vo, nt = w.TO(s + 1u) - to;
                if ((nv | nt) == 0u) continue;
                EmitSink sink;
                sink.vtx = gv + vo;
                sink.idx = gi + 3ull * to;
                sink.nv = sink.nt = 0u;
                sink.base = gbase + vo;
                const uint32_t pk = prev_kept(w, s), nk = next_kept_or_end(w, s);
                emit_full_join(w, p, s, pk, nk, gbase, w.jcode, sink);
                mismatch |= sink.nv != nv || sink.nt != nt;
            }
            WPT(10);
            // End slots with output leave the tile as records for k_stroke_ends
            for (uint32_t r0 = 0u; r0 < n_end; r0 += 32u) {
                const uint32_t r = r0 + lane;
                uint32_t s = 0u, vo = 0u, to = 0u, nv = 0u, nt = 0u;
                if (r < n_end) {
                    s = list[X.s_cap - 1u - r];
                    vo = w.VO(s);
                    to = w.TO(s);
                    nv = w.VO(s + 1u) - vo;
                    nt = w.TO(s + 1u) - to;
                }
                const bool has = (nv | nt) != 0u;
                const unsigned hm = __ballot_sync(FULL, has);
                uint32_t slot = 0u;
                if (lane == 0u && hm) slot = atomicAdd(A.n_end_rec, (uint32_t)__popc(hm));
                slot = __shfl_sync(FULL, slot, 0) + (uint32_t)__popc(hm & lt);
                if (has) {
                    if (slot < A.end_rec_cap) {
                        EndRecord rec;
                        end_record_fill(w, s, &rec);
                        // the two stream joins end() refers to: at the second kept point and at the second-to-last one
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
                        const uint4* srcv = reinterpret_cast<const uint4*>(&rec);
                        uint4* dstv = reinterpret_cast<uint4*>(&A.end_rec[slot]);
#pragma unroll
                        for (int q = 0; q < 8; ++q) dstv[q] = srcv[q];
                    } else {
                        mismatch = true;
                    }
                }
            }
            if (mismatch) atomicOr(&A.totals->error, TILE_ERR_EMIT_MISMATCH);
        }
        __syncwarp();
        WPT(11);
    }
#ifdef LYON_PHASE_TIMING
    if (lane == 0u)
        for (int i = 0; i < 12; ++i) atomicAdd(&g_phase_cycles[i], pt_acc[i]);
#endif
}

// Everything end() emits, one thread per EndRecord (dense: caps with their arcs, closing joins).
__global__ void __launch_bounds__(128) k_stroke_ends(const EndRecord* __restrict__ recs, const uint32_t* __restrict__ n_recs,
                                                     uint32_t cap, StrokeParams p, GpuVertex* vertices, uint32_t* indices,
                                                     TileTotals* totals) {
    const uint32_t n = min(*n_recs, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        EndRecord rec;
        {
            const uint4* srcv = reinterpret_cast<const uint4*>(&recs[i]);
            uint4* dstv = reinterpret_cast<uint4*>(&rec);
#pragma unroll
            for (int q = 0; q < 8; ++q) dstv[q] = srcv[q];
        }
        EndRecView w;
        w.r = &rec;
        w.n_slots = 7u;
        EmitSink sink;
        sink.vtx = vertices + rec.vtx_index;
        sink.idx = indices + 3ull * rec.tri_index;
        sink.nv = sink.nt = 0u;
        sink.base = rec.first_id;
        stroke_end_slot(w, p, 6u, 0u, sink);
        if (sink.nv != rec.nv || sink.nt != rec.nt) atomicOr(&totals->error, TILE_ERR_EMIT_MISMATCH);
    }
}
