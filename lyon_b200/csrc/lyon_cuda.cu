// lyon_cuda.cu -- CUDA kernels (sm_100a) and the C ABI of include/lyon_cuda.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false (see __graft_entry__.build()).
// -fmad=false is part of the numerical contract: lyon (rustc) never contracts a*b+c, and vertex/index
// topology depends on f32 compares of such expressions (hard part 1).
//
// Kernel inventory (the per-item logic lives in stroke_core.cuh / device_math.cuh):
//   k_seg_setup     1 thread / verb        two chained decoupled-look-back scans (points, slots)
//   k_point_gen     1 thread / slot        segment tiles, smem binary search, coalesced slot writes
//   k_path_walk     1 thread / path        sequential f32 chains (merge filter, advancement)
//   k_classify_count / k_full_count / k_end_count / k_scan_offsets   count pass split by slot population + in-place scan
//   k_emit_fast / k_emit_full / k_emit_end                            emit pass, StrokeVertexGpu (2 x 16 B stores) + u32 indices
//   k_path_ranges   1 thread / path        per-path output ranges
//   k_flat_count / k_flat_emit             standalone curve batches (lyon_geom flattening)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/lyon_cuda.h"
#include "stroke_tile.cuh"
#include "stroke_wtile.cuh"
#include "stroke_vw.cuh"

using namespace lyc;

static_assert(sizeof(GpuVertex) == 32 && sizeof(lyon_stroke_vertex) == 32, "StrokeVertexGpu must be 32 bytes");

namespace {

constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;

// ---------------------------------------------------------------------------------------------------------
// Decoupled look-back (single pass chained scan).  Descriptor word: [63:62] status, [61:0] value.
// ---------------------------------------------------------------------------------------------------------
constexpr uint64_t ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, VAL_MASK = (1ull << 62) - 1ull;

#ifdef LYON_PHASE_TIMING
__device__ unsigned long long g_lookback_stats[4];
#endif
__device__ __forceinline__ uint64_t ld_desc(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Called by all 32 lanes of ONE warp. Publishes `aggregate` for `tile`, returns the exclusive prefix.
// `published`: the caller already made the aggregate visible with lookback_publish (so that work which does not need
// the prefix can run between the two).
__device__ __forceinline__ void lookback_publish(uint64_t* desc, uint32_t tile, uint64_t aggregate) {
    st_desc(&desc[tile], (tile == 0u ? ST_PREFIX : ST_AGG) | aggregate);
}
__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t* desc, uint32_t tile, uint64_t aggregate, bool published = false) {
    const unsigned lane = threadIdx.x & 31u;
    if (tile == 0u) {
        if (lane == 0u && !published) st_desc(&desc[0], ST_PREFIX | aggregate);
        return 0ull;
    }
    if (lane == 0u && !published) st_desc(&desc[tile], ST_AGG | aggregate);
    uint64_t excl = 0ull;
    int64_t look = (int64_t)tile - 1;
    for (;;) {
        int64_t idx = look - (int64_t)lane;
        uint64_t d = ST_PREFIX;  // tiles before the first one: prefix 0
#ifdef LYON_PHASE_TIMING
        if (lane == 0u) atomicAdd(&g_lookback_stats[0], 1ull);  // windows
#endif
        if (idx >= 0) {
            d = ld_desc(&desc[idx]);
            while ((d >> 62) == 0ull) {
#ifdef LYON_PHASE_TIMING
                atomicAdd(&g_lookback_stats[1], 1ull);  // polls of a descriptor that was not ready
#endif
                __nanosleep(64);
                d = ld_desc(&desc[idx]);
            }
        }
        unsigned pmask = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
        uint64_t val = d & VAL_MASK;
        if (pmask) {
            unsigned first = (unsigned)__ffs((int)pmask) - 1u;
            excl += warp_sum_u64(lane <= first ? val : 0ull);
            break;
        }
        excl += warp_sum_u64(val);
        look -= 32;
    }
    if (lane == 0u) st_desc(&desc[tile], ST_PREFIX | (excl + aggregate));
    return excl;
}
// The two look-backs of a stroke tile (vertex and triangle descriptors) advanced together: both descriptors of a
// predecessor are loaded in the same round trip.  Each array keeps its own state (a predecessor's two descriptors are
// two stores and may be observed in different states).  Called by all 32 lanes of ONE warp.
__device__ __forceinline__ void lookback_exclusive_pair(uint64_t* desc_a, uint64_t* desc_b, uint32_t tile, uint64_t agg_a,
                                                        uint64_t agg_b, uint64_t* excl_a, uint64_t* excl_b,
                                                        bool published = false) {
    const unsigned lane = threadIdx.x & 31u;
    if (tile == 0u) {
        if (lane == 0u && !published) {
            st_desc(&desc_a[0], ST_PREFIX | agg_a);
            st_desc(&desc_b[0], ST_PREFIX | agg_b);
        }
        *excl_a = *excl_b = 0ull;
        return;
    }
    if (lane == 0u && !published) {
        st_desc(&desc_a[tile], ST_AGG | agg_a);
        st_desc(&desc_b[tile], ST_AGG | agg_b);
    }
    uint64_t ea = 0ull, eb = 0ull;
    bool done_a = false, done_b = false;
    int64_t look = (int64_t)tile - 1;
    for (;;) {
        const int64_t idx = look - (int64_t)lane;
        uint64_t da = ST_PREFIX, db = ST_PREFIX;  // tiles before the first one: prefix 0
        if (idx >= 0) {
            if (!done_a) da = ld_desc(&desc_a[idx]);
            if (!done_b) db = ld_desc(&desc_b[idx]);
            while ((da >> 62) == 0ull || (db >> 62) == 0ull) {
                __nanosleep(64);
                if ((da >> 62) == 0ull) da = ld_desc(&desc_a[idx]);
                if ((db >> 62) == 0ull) db = ld_desc(&desc_b[idx]);
            }
        }
        if (!done_a) {
            const unsigned pmask = __ballot_sync(0xffffffffu, (da >> 62) == 2ull);
            const uint64_t val = da & VAL_MASK;
            if (pmask) {
                const unsigned first = (unsigned)__ffs((int)pmask) - 1u;
                ea += warp_sum_u64(lane <= first ? val : 0ull);
                done_a = true;
            } else {
                ea += warp_sum_u64(val);
            }
        }
        if (!done_b) {
            const unsigned pmask = __ballot_sync(0xffffffffu, (db >> 62) == 2ull);
            const uint64_t val = db & VAL_MASK;
            if (pmask) {
                const unsigned first = (unsigned)__ffs((int)pmask) - 1u;
                eb += warp_sum_u64(lane <= first ? val : 0ull);
                done_b = true;
            } else {
                eb += warp_sum_u64(val);
            }
        }
        if (done_a && done_b) break;
        look -= 32;
    }
    if (lane == 0u) {
        st_desc(&desc_a[tile], ST_PREFIX | (ea + agg_a));
        st_desc(&desc_b[tile], ST_PREFIX | (eb + agg_b));
    }
    *excl_a = ea;
    *excl_b = eb;
}
// Block-wide exclusive scan of one u32 per thread. Returns the thread's exclusive prefix; *total = block sum.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t* s_warp /*[WARPS+1]*/, uint32_t* total) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += y;
    }
    if (lane == 31u) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0u) {
        uint32_t w = lane < (unsigned)WARPS ? s_warp[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < WARPS; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (unsigned)o) winc += y;
        }
        if (lane < (unsigned)WARPS) s_warp[lane] = winc - w;
        if (lane == (unsigned)WARPS - 1u) s_warp[WARPS] = winc;
    }
    __syncthreads();
    uint32_t r = s_warp[wid] + inc - x;
    *total = s_warp[WARPS];
    __syncthreads();
    return r;
}

struct Totals {
    unsigned long long points, slots, vertices, triangles, flat_points, subpaths;
    unsigned int error;
    unsigned int _pad;
};

// ---------------------------------------------------------------------------------------------------------
// K1: per verb -- stored-point offset (scan #1), then slot count from the Wang formula (scan #2).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_seg_setup(Views w, StrokeParams p, uint32_t* ticket, uint64_t* desc_pts,
                                                     uint64_t* desc_slots, uint64_t* desc_sp, Totals* totals) {
    __shared__ uint32_t s_warp[WARPS + 1];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t v = tile * BLOCK + threadIdx.x;
    const bool valid = v < w.n_verbs;
    const uint8_t verb = valid ? w.verbs[v] : (uint8_t)VERB_END;
    const uint32_t npts = valid ? verb_point_count(verb, p.attr_stride) : 0u;
    uint32_t agg;
    uint32_t local = block_exclusive_scan(npts, s_warp, &agg);
    if (threadIdx.x < 32) {
        uint64_t e = lookback_exclusive(desc_pts, tile, agg);
        if (threadIdx.x == 0) s_base = e;
    }
    __syncthreads();
    const unsigned long long pt64 = s_base + local;
    const uint32_t pt = (uint32_t)pt64;
    float tsplit = -1.0f;
    uint32_t nslots = 0u;
    if (valid) {
        nslots = seg_slot_count(verb, w.points, pt, p.attr_stride, p.tolerance, &tsplit);
        w.seg_pt[v] = pt;
        w.seg_tsplit[v] = tsplit;
    }
    __syncthreads();
    local = block_exclusive_scan(nslots, s_warp, &agg);
    // A single verb can ask for up to 2^32-1 slots; sum in 64 bits so the host can reject the batch.
    unsigned long long agg64 = agg;
    {
        // detect u32 wrap inside the block scan: recompute the block sum in 64 bits (rare, cheap)
        unsigned long long x = nslots;
        x = warp_sum_u64(x);
        __shared__ unsigned long long s_sum[WARPS];
        if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int i = 0; i < WARPS; ++i) t += s_sum[i];
            s_sum[0] = t;
        }
        __syncthreads();
        agg64 = s_sum[0];
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        uint64_t e = lookback_exclusive(desc_slots, tile, agg64);
        if (threadIdx.x == 0) s_base = e;
    }
    __syncthreads();
    const unsigned long long so = s_base + local;
    const unsigned long long slots_total = s_base + agg64;
    if (valid) w.seg_off[v] = (uint32_t)so;
    __syncthreads();
    // scan #3: sub-path index of every End/Close verb (stored in the verb's otherwise unused tsplit field)
    const uint32_t is_end = (valid && (verb == VERB_END || verb == VERB_CLOSE)) ? 1u : 0u;
    local = block_exclusive_scan(is_end, s_warp, &agg);
    if (threadIdx.x < 32) {
        uint64_t e = lookback_exclusive(desc_sp, tile, agg);
        if (threadIdx.x == 0) s_base = e;
    }
    __syncthreads();
    if (is_end) w.seg_tsplit[v] = __uint_as_float((uint32_t)(s_base + local));
    if (v == w.n_verbs - 1u) {
        w.seg_pt[w.n_verbs] = (uint32_t)(pt64 + npts);
        w.seg_off[w.n_verbs] = (uint32_t)slots_total;
        totals->points = pt64 + npts;
        totals->slots = slots_total;
        totals->subpaths = s_base + local + is_end;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2: per slot.  A block owns a tile of BLOCK verbs and walks their slots BLOCK at a time, so global
// writes are slot-contiguous (coalesced) whatever the per-curve flatten counts are.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_point_gen(Views w, StrokeParams p, uint32_t* __restrict__ sp_end_slot) {
    __shared__ uint32_t s_off[BLOCK + 1];
    const uint32_t v0 = blockIdx.x * BLOCK;
    const uint32_t nv = min((uint32_t)BLOCK, w.n_verbs - v0);
    for (uint32_t i = threadIdx.x; i <= nv; i += BLOCK) s_off[i] = w.seg_off[v0 + i];
    __syncthreads();
    const uint32_t base = s_off[0];
    const uint32_t total = s_off[nv] - base;
    for (uint32_t i = threadIdx.x; i < total; i += BLOCK) {
        const uint32_t s = base + i;
        uint32_t lo = 0, hi = nv;  // last j in [0, nv) with s_off[j] <= s
        while (hi - lo > 1u) {
            uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] <= s) lo = mid; else hi = mid;
        }
        point_gen_slot(w, p, v0 + lo, s, s - s_off[lo]);
        const uint8_t verb = w.verbs[v0 + lo];
        if (verb == VERB_END || verb == VERB_CLOSE) sp_end_slot[__float_as_uint(w.seg_tsplit[v0 + lo])] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3: per path sequential walk.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_path_walk(Views w, StrokeParams p, const uint32_t* __restrict__ path_verb_begin,
                                                   uint32_t n_paths) {
    const uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
    if (path >= n_paths) return;
    path_walk(w, p, path_verb_begin[path], path_verb_begin[path + 1]);
}

// ---------------------------------------------------------------------------------------------------------
// K4 (count) and K5 (emit).  Slots fall in three populations with very different costs, so each gets its own
// dense launch instead of diverging inside one kernel:
//   fast joins  (~75 % of slots: joins inside flattened curves)  thread per slot, a few dozen flops
//   full joins  (real corners: miter/bevel/round logic, arcs)    compacted list built by the classify pass
//   End slots   (caps or closing joins, one per sub-path)         list indexed by sub-path (built by K1/K2)
// count:  k_classify_count -> k_full_count -> k_end_count -> k_scan_offsets (in-place, decoupled look-back)
// emit:   k_emit_fast -> k_emit_full -> k_emit_end
// ---------------------------------------------------------------------------------------------------------
// counts of fast joins are written directly; full joins are appended to `full_list` (warp-aggregated atomics).
__global__ void __launch_bounds__(BLOCK) k_classify_count(Views w, StrokeParams p, uint8_t* __restrict__ jcode,
                                                          uint32_t* __restrict__ full_list, uint32_t* full_count) {
    const uint32_t s = blockIdx.x * BLOCK + threadIdx.x;
    uint32_t nv = 0u, nt = 0u;
    uint8_t code = 0;
    bool is_full = false;
    SlotClass c = classify_slot(w, s);
    if (c.kind == KIND_JOIN) {
        if (join_is_fast(w.pos[c.pk], w.pos[s], w.pos[c.nk], (c.sf & SEG_FS_BIT) != 0u)) {
            nv = 2u;  // flattened_step: Positive, Negative (stroke.rs:1791-1797)
            nt = (w.segfs[c.pk] & SEG_BEGIN_BIT) ? 0u : 2u;
            code = JC_FAST | JC_SINGLE_POS | JC_SINGLE_NEG;
        } else {
            is_full = true;
        }
    }
    if (s < w.n_slots) {
        w.voff[s] = nv;
        w.toff[s] = nt;
        jcode[s] = code;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, is_full);
    if (mask) {
        const unsigned lane = threadIdx.x & 31u;
        uint32_t base = 0u;
        if (lane == (unsigned)(__ffs((int)mask) - 1)) base = atomicAdd(full_count, (uint32_t)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, __ffs((int)mask) - 1);
        if (is_full) full_list[base + __popc(mask & ((1u << lane) - 1u))] = s;
    }
}
__global__ void __launch_bounds__(128) k_full_count(Views w, StrokeParams p, uint8_t* __restrict__ jcode,
                                                    const uint32_t* __restrict__ full_list, const uint32_t* full_count) {
    const uint32_t n = *full_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = full_list[i];
        const uint32_t pk = prev_kept(w, s), nk = next_kept_or_end(w, s);
        uint32_t nv, nt;
        uint8_t code;
        count_full_join(w, p, s, pk, nk, &nv, &nt, &code);
        w.voff[s] = nv;
        w.toff[s] = nt;
        jcode[s] = code;
    }
}
__global__ void __launch_bounds__(128) k_end_count(Views w, StrokeParams p, const uint32_t* __restrict__ sp_end_slot,
                                                   uint32_t n_subpaths) {
    const uint32_t sp = blockIdx.x * blockDim.x + threadIdx.x;
    if (sp >= n_subpaths) return;
    const uint32_t s = sp_end_slot[sp];
    CountSink sink;
    sink.nv = sink.nt = 0u;
    sink.base = 0u;
    stroke_end_slot(w, p, s, 0u, sink);
    w.voff[s] = sink.nv;
    w.toff[s] = sink.nt;
}
// In-place exclusive scan of the per-slot counts in voff / toff (4 slots per thread).
constexpr int SCAN_ITEMS = 4;
__global__ void __launch_bounds__(BLOCK) k_scan_offsets(Views w, uint32_t* ticket, uint64_t* desc_v, uint64_t* desc_t,
                                                        Totals* totals) {
    __shared__ uint32_t s_warp[WARPS + 1];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base_v, s_base_t;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t s0 = (tile * BLOCK + threadIdx.x) * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], t[SCAN_ITEMS];
    if (s0 + SCAN_ITEMS <= w.n_slots) {  // buffers are 16-byte aligned and s0 is a multiple of 4
        const uint4 a = *reinterpret_cast<const uint4*>(w.voff + s0);
        const uint4 b = *reinterpret_cast<const uint4*>(w.toff + s0);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        t[0] = b.x; t[1] = b.y; t[2] = b.z; t[3] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = (s0 + k < w.n_slots) ? w.voff[s0 + k] : 0u;
            t[k] = (s0 + k < w.n_slots) ? w.toff[s0 + k] : 0u;
        }
    }
    const uint32_t sum_v = v[0] + v[1] + v[2] + v[3], sum_t = t[0] + t[1] + t[2] + t[3];
    uint32_t agg_v, agg_t;
    uint32_t lv = block_exclusive_scan(sum_v, s_warp, &agg_v);
    uint32_t lt = block_exclusive_scan(sum_t, s_warp, &agg_t);
    if (threadIdx.x < 32) {
        uint64_t ev = lookback_exclusive(desc_v, tile, agg_v);
        uint64_t et = lookback_exclusive(desc_t, tile, agg_t);
        if (threadIdx.x == 0) {
            s_base_v = ev;
            s_base_t = et;
        }
    }
    __syncthreads();
    unsigned long long bv = s_base_v + lv, bt = s_base_t + lt;
    uint32_t ov[SCAN_ITEMS], ot[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        ov[k] = (uint32_t)bv;
        ot[k] = (uint32_t)bt;
        bv += v[k];
        bt += t[k];
    }
    if (s0 + SCAN_ITEMS <= w.n_slots) {
        *reinterpret_cast<uint4*>(w.voff + s0) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
        *reinterpret_cast<uint4*>(w.toff + s0) = make_uint4(ot[0], ot[1], ot[2], ot[3]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (s0 + k < w.n_slots) {
                w.voff[s0 + k] = ov[k];
                w.toff[s0 + k] = ot[k];
            }
    }
    if (s0 < w.n_slots && s0 + SCAN_ITEMS >= w.n_slots) {  // the thread owning the last slot
        w.voff[w.n_slots] = (uint32_t)bv;
        w.toff[w.n_slots] = (uint32_t)bt;
        totals->vertices = bv;
        totals->triangles = bt;
    }
}

// ---- emit ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_emit_fast(Views w, StrokeParams p, const uint8_t* __restrict__ jcode,
                                                     GpuVertex* __restrict__ vertices, uint32_t* __restrict__ indices,
                                                     uint32_t vertex_base, Totals* totals) {
    const uint32_t s = blockIdx.x * BLOCK + threadIdx.x;
    if (s >= w.n_slots) return;
    if (!(jcode[s] & JC_FAST)) return;
    const uint32_t pk = prev_kept(w, s), nk = next_kept_or_end(w, s);
    const uint32_t vo = w.voff[s], to = w.toff[s];
    EmitSink sink;
    sink.vtx = vertices + vo;
    sink.idx = indices + 3ull * to;
    sink.nv = sink.nt = 0u;
    sink.base = vertex_base + vo;
    emit_fast_join(w, p, s, pk, nk, vertex_base, jcode, sink);
    if ((sink.nv != w.voff[s + 1] - vo) || (sink.nt != w.toff[s + 1] - to)) atomicOr(&totals->error, 1u);
}
__global__ void __launch_bounds__(128) k_emit_full(Views w, StrokeParams p, const uint8_t* __restrict__ jcode,
                                                   const uint32_t* __restrict__ full_list, const uint32_t* full_count,
                                                   GpuVertex* __restrict__ vertices, uint32_t* __restrict__ indices,
                                                   uint32_t vertex_base, Totals* totals) {
    const uint32_t n = *full_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = full_list[i];
        const uint32_t pk = prev_kept(w, s), nk = next_kept_or_end(w, s);
        const uint32_t vo = w.voff[s], to = w.toff[s];
        EmitSink sink;
        sink.vtx = vertices + vo;
        sink.idx = indices + 3ull * to;
        sink.nv = sink.nt = 0u;
        sink.base = vertex_base + vo;
        emit_full_join(w, p, s, pk, nk, vertex_base, jcode, sink);
        if ((sink.nv != w.voff[s + 1] - vo) || (sink.nt != w.toff[s + 1] - to)) atomicOr(&totals->error, 1u);
    }
}
__global__ void __launch_bounds__(128) k_emit_end(Views w, StrokeParams p, const uint32_t* __restrict__ sp_end_slot,
                                                  uint32_t n_subpaths, GpuVertex* __restrict__ vertices,
                                                  uint32_t* __restrict__ indices, uint32_t vertex_base, Totals* totals) {
    const uint32_t sp = blockIdx.x * blockDim.x + threadIdx.x;
    if (sp >= n_subpaths) return;
    const uint32_t s = sp_end_slot[sp];
    const uint32_t vo = w.voff[s], to = w.toff[s];
    const uint32_t nv = w.voff[s + 1] - vo, nt = w.toff[s + 1] - to;
    if ((nv | nt) == 0u) return;
    EmitSink sink;
    sink.vtx = vertices + vo;
    sink.idx = indices + 3ull * to;
    sink.nv = sink.nt = 0u;
    sink.base = vertex_base + vo;
    stroke_end_slot(w, p, s, vertex_base, sink);
    if (sink.nv != nv || sink.nt != nt) atomicOr(&totals->error, 1u);
}

__global__ void __launch_bounds__(128) k_path_ranges(Views w, const uint32_t* __restrict__ path_verb_begin, uint32_t n_paths,
                                                     unsigned long long* path_vertex_end, unsigned long long* path_index_end,
                                                     const Totals* totals) {
    const uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
    if (path >= n_paths) return;
    // first slot after the path; offsets are u32-wrapped, rebuild 64-bit values monotonically is not needed:
    // totals are checked < 2^32 for vertices; triangles likewise (see plan()).
    const uint32_t s_end = w.seg_off[path_verb_begin[path + 1]];
    if (path_vertex_end) path_vertex_end[path] = w.voff[s_end];
    if (path_index_end) path_index_end[path] = 3ull * w.toff[s_end];
    (void)totals;
}

// ---------------------------------------------------------------------------------------------------------
// Standalone curve batches: CubicBezierSegment / QuadraticBezierSegment ::for_each_flattened.
// ---------------------------------------------------------------------------------------------------------
template <int NPTS>  // 4 = cubic, 3 = quadratic
__device__ __forceinline__ uint32_t curve_count(const float* __restrict__ c, float tol) {
    if (NPTS == 4)
        return to_u32_or_1(cubic_wang(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), v2(c[6], c[7]), tol));
    return to_u32_or_1(quad_wang(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), tol));
}
template <int NPTS>
__global__ void __launch_bounds__(BLOCK) k_flat_count(const float* __restrict__ ctrl, unsigned long long n, float tol,
                                                      unsigned long long* __restrict__ offsets, uint32_t* ticket,
                                                      uint64_t* desc, Totals* totals) {
    __shared__ uint32_t s_warp[WARPS + 1];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const unsigned long long i = (unsigned long long)tile * BLOCK + threadIdx.x;
    uint32_t cnt = 0u;
    if (i < n) {
        float c[2 * NPTS];
        const float2* src = reinterpret_cast<const float2*>(ctrl) + i * NPTS;
#pragma unroll
        for (int k = 0; k < NPTS; ++k) {
            float2 q = src[k];
            c[2 * k] = q.x;
            c[2 * k + 1] = q.y;
        }
        cnt = curve_count<NPTS>(c, tol);
    }
    // counts are < 2^32 each; a block of 256 may exceed u32 only for absurd inputs -> 64-bit aggregate
    unsigned long long x = warp_sum_u64(cnt);
    __shared__ unsigned long long s_sum[WARPS];
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = x;
    uint32_t agg32;
    uint32_t local = block_exclusive_scan(cnt, s_warp, &agg32);
    unsigned long long agg = 0;
    for (int k = 0; k < WARPS; ++k) agg += s_sum[k];
    if (threadIdx.x < 32) {
        uint64_t e = lookback_exclusive(desc, tile, agg);
        if (threadIdx.x == 0) s_base = e;
    }
    __syncthreads();
    if (i < n) offsets[i] = s_base + local;
    if (i == n - 1ull) {
        offsets[n] = s_base + local + cnt;
        totals->flat_points = s_base + local + cnt;
    }
}
// Emit pass of the standalone curve flattener.  A block owns FLAT_CURVES consecutive curves, one per thread, and moves
// their output points through a shared-memory window:
//   phase A  thread / curve   the thread keeps its curve's polynomial form (cubic_bezier.rs:1318-1325) in registers and
//                             walks its points in order: t1 = t0 + step accumulated in f32 (cubic_bezier.rs:605-611,
//                             quadratic_bezier.rs:382-388), Horner evaluation (cubic_bezier.rs:1351-1365), the last
//                             point is the exact endpoint.  Points go to the window at their output position.
//   phase B  thread / point   the window leaves the SM as contiguous 8-byte stores (a warp = 256 contiguous bytes).
// HBM traffic: control points and offsets once, every output point once.
#ifndef LYON_FLAT_CURVES
#define LYON_FLAT_CURVES 128   // measured on C2 (10 M cubics): 256 / 8192: 0.83 ms, 128 / 3584: 0.72 ms, 64 / 1792: 0.76 ms
#define LYON_FLAT_WINDOW 3584
#endif
constexpr int FLAT_CURVES = LYON_FLAT_CURVES;
constexpr int FLAT_WINDOW = LYON_FLAT_WINDOW;  // points per window (half of it when the curve parameters are written too)
template <int NPTS>
struct FlatSmem {
    unsigned long long off[FLAT_CURVES + 1];
    float2 pts[FLAT_WINDOW];  // with t output: first half points, second half (as float) the parameters
};
// FUSED: one pass.  The block takes a ticket, counts its curves (Wang), gets the position of its first point by
// decoupled look-back over the block totals, writes the offsets it was given to compute, and emits -- unless the
// caller's point buffer is too small, in which case it only counts (lyon_cuda_flatten_*_fused then reports the size).
struct FlatFused {
    unsigned long long* offsets_out;  // [n + 1]
    unsigned long long capacity;      // points the output buffers can hold
    uint32_t* ticket;
    uint64_t* desc;
    Totals* totals;
};
template <int NPTS, bool FUSED, bool WITH_T>
__global__ void __launch_bounds__(FLAT_CURVES) k_flat_emit(const float* __restrict__ ctrl, unsigned long long n, float tol,
                                                           const unsigned long long* __restrict__ offsets,
                                                           float2* __restrict__ out, float* __restrict__ t_out, FlatFused fu) {
    extern __shared__ __align__(16) unsigned char flat_smem_raw[];
    FlatSmem<NPTS>& sm = *reinterpret_cast<FlatSmem<NPTS>*>(flat_smem_raw);
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_wsum[FLAT_CURVES / 32], s_base;
    const uint32_t tid = threadIdx.x;
    uint32_t tile = blockIdx.x;
    if (FUSED) {
        if (tid == 0) s_tile = atomicAdd(fu.ticket, 1u);
        __syncthreads();
        tile = s_tile;
    }
    const unsigned long long c0 = (unsigned long long)tile * FLAT_CURVES;
    const uint32_t nc = (uint32_t)min((unsigned long long)FLAT_CURVES, n - c0);
    if (!FUSED)
        for (uint32_t i = tid; i <= nc; i += FLAT_CURVES) sm.off[i] = offsets[c0 + i];
    // stage the control points through the (not yet used) window: coalesced 4-byte loads
    float* stage = reinterpret_cast<float*>(sm.pts);
    for (uint32_t i = tid; i < nc * 2u * NPTS; i += FLAT_CURVES) stage[i] = ctrl[c0 * 2ull * NPTS + i];
    __syncthreads();
    float step = 0.0f;
    uint32_t cnt = 0u;
    V2 a0 = v2(0.0f, 0.0f), a1 = a0, a2 = a0, a3 = a0, to = a0;
    if (tid < nc) {
        const float* c = stage + tid * 2u * NPTS;
        float nseg;
        if (NPTS == 4) {
            const V2 from = v2(c[0], c[1]), c1 = v2(c[2], c[3]), c2 = v2(c[4], c[5]);
            to = v2(c[6], c[7]);
            nseg = cubic_wang(from, c1, c2, to, tol);
            const CubicPoly q = cubic_poly(from, c1, c2, to);
            a0 = q.a0; a1 = q.a1; a2 = q.a2; a3 = q.a3;
        } else {
            const V2 from = v2(c[0], c[1]), cc = v2(c[2], c[3]);
            to = v2(c[4], c[5]);
            nseg = quad_wang(from, cc, to, tol);
            const QuadPoly q = quad_poly(from, cc, to);
            a0 = q.a0; a1 = q.a1; a2 = q.a2;
        }
        step = 1.0f / nseg;
        cnt = to_u32_or_1(nseg);
    }
    unsigned long long base = 0ull, total = 0ull, my0 = 0ull;  // my0: block-relative index of my first point
    if (FUSED) {
        // exclusive scan of the counts inside the block (64-bit); the block total is published at once, the look-back
        // for the block's base runs after the first window has been computed (phase A does not need it)
        const unsigned lane = tid & 31u, wid = tid >> 5;
        unsigned long long inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += y;
        }
        if (lane == 31u) s_wsum[wid] = inc;
        __syncthreads();  // also: `stage` (= the window) is free again
        unsigned long long wbase = 0ull;
        for (int k = 0; k < FLAT_CURVES / 32; ++k) {
            if (k < (int)wid) wbase += s_wsum[k];
            total += s_wsum[k];
        }
        if (tid == 0) lookback_publish(fu.desc, tile, total);
        my0 = wbase + inc - cnt;
    } else {
        __syncthreads();  // `stage` (= the window) is free again
        base = sm.off[0];
        total = sm.off[nc] - base;
        my0 = tid < nc ? sm.off[tid] - base : 0ull;
    }
    constexpr uint32_t window = WITH_T ? FLAT_WINDOW / 2 : FLAT_WINDOW;
    float* const win_t = reinterpret_cast<float*>(sm.pts + FLAT_WINDOW / 2);
    float2* gout = out + base;
    float* gt = WITH_T ? t_out + base : nullptr;
    bool resolved = !FUSED;
    float t = 0.0f;
    uint32_t k = 0u;
    for (unsigned long long w0 = 0; w0 < total; w0 += window) {
        const unsigned long long w1 = min(total, w0 + (unsigned long long)window);
        // ---- phase A: my points [my0 + k, my0 + cnt) that fall into the window [w0, w1).  The sample loop is branch-free
        // (the exact endpoint is peeled off) and indexes the window with 32-bit arithmetic: a third of the instructions
        // per point of the straightforward loop, which was what bound the kernel.
        {
            const unsigned long long pos0 = my0 + k, pend = min(my0 + (unsigned long long)cnt, w1);
            if (k < cnt && pos0 < pend) {
                const uint32_t i0 = (uint32_t)(pos0 - w0), nk = (uint32_t)(pend - pos0);
                const bool has_last = k + nk == cnt;
                const uint32_t ninter = has_last ? nk - 1u : nk;
                float2* const wp = sm.pts + i0;
                float* const wt = win_t + i0;
#pragma unroll 4
                for (uint32_t j = 0; j < ninter; ++j) {
                    t = t + step;
                    float u = t;
                    V2 pos = a0 + a1 * u;
                    u = u * t;
                    pos = pos + a2 * u;
                    if (NPTS == 4) {
                        u = u * t;
                        pos = pos + a3 * u;
                    }
                    wp[j] = make_float2(pos.x, pos.y);
                    if (WITH_T) wt[j] = t;
                }
                if (has_last) {
                    wp[ninter] = make_float2(to.x, to.y);
                    if (WITH_T) wt[ninter] = 1.0f;
                }
                k += nk;
            }
        }
        if (FUSED && !resolved) {
            // the block's base: look-back by warp 0 (the first window is already computed), then offsets out
            if (tid < 32) {
                const uint64_t e = lookback_exclusive(fu.desc, tile, total, true);
                if (tid == 0) s_base = e;
            }
            __syncthreads();
            base = s_base;
            resolved = true;
            if (tid < nc) fu.offsets_out[c0 + tid] = base + my0;
            if (tid == nc - 1u && c0 + nc == n) {
                fu.offsets_out[n] = base + total;
                fu.totals->flat_points = base + total;
            }
            if (base + total > fu.capacity) {
                if (tid == 0) atomicOr(&fu.totals->error, 0x100u);  // point buffer too small: counted, not emitted
                return;
            }
            gout = out + base;
            gt = WITH_T ? t_out + base : nullptr;
        }
        __syncthreads();
        // ---- phase B: the window leaves the SM as 16-byte stores (two points each; the odd point in front of the first
        // 16-byte boundary and the one behind the last go out on their own)
        const uint32_t wn = (uint32_t)(w1 - w0);
        {
            float2* const g = gout + w0;
            const uint32_t head = (uint32_t)((reinterpret_cast<uintptr_t>(g) >> 3) & 1u) & (wn ? 1u : 0u);
            if (tid == 0 && head) __stcs(g, sm.pts[0]);
            const uint32_t pairs = (wn - head) >> 1;
            for (uint32_t i = tid; i < pairs; i += FLAT_CURVES) {
                const float2 p0 = sm.pts[head + 2u * i], p1 = sm.pts[head + 2u * i + 1u];
                __stcs(reinterpret_cast<float4*>(g + head) + i, make_float4(p0.x, p0.y, p1.x, p1.y));
            }
            if (tid == FLAT_CURVES - 1 && ((wn - head) & 1u)) __stcs(g + wn - 1u, sm.pts[wn - 1u]);
            if (WITH_T)
                for (uint32_t i = tid; i < wn; i += FLAT_CURVES) __stcs(gt + w0 + i, win_t[i]);
        }
        __syncthreads();
    }
}

// The same emit pass with WARP-owned tiles (experiment, LYON_FLAT=warp): a warp owns 32 consecutive curves and a window of
// its own, so nothing in the kernel is block-wide -- k_flat_emit spends a third of its warp time in the two block barriers
// of every window.  FUSED: a warp takes a ticket, scans its 32 counts with shuffles, publishes the total, computes its
// first window and only then resolves its base by decoupled look-back over the warp totals (one descriptor per 32
// curves).  Measured on C2: 0.88 ms (640-point windows) against 0.72 ms for the block tiles -- eight times the tickets
// and look-backs cost more than the barriers; the scattered window stores of phase A (bank conflicts) are the same.
constexpr int FLATW_WARPS = 4;        // warps per block (the block is only a container)
#ifndef LYON_FLATW_WINDOW
#define LYON_FLATW_WINDOW 640
#endif
constexpr int FLATW_WINDOW = LYON_FLATW_WINDOW;  // points per warp window (half of it when the curve parameters are written too)
template <int NPTS, bool FUSED, bool WITH_T>
__global__ void __launch_bounds__(FLATW_WARPS * 32) k_flat_emit_w(const float* __restrict__ ctrl, unsigned long long n, float tol,
                                                                 const unsigned long long* __restrict__ offsets,
                                                                 float2* __restrict__ out, float* __restrict__ t_out, FlatFused fu) {
    __shared__ __align__(16) float2 s_win[FLATW_WARPS][FLATW_WINDOW];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    float2* const win = s_win[wid];
    uint32_t tile = blockIdx.x * FLATW_WARPS + wid;
    if (FUSED) {
        if (lane == 0u) tile = atomicAdd(fu.ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
    }
    const unsigned long long c0 = (unsigned long long)tile * 32ull;
    if (c0 >= n) return;
    const uint32_t nc = (uint32_t)min(32ull, n - c0);
    // the warp's control points (nc * NPTS * 8 contiguous bytes) through its window: coalesced 4-byte loads
    static_assert(FLATW_WINDOW * 2 >= 32 * 2 * NPTS, "window too small to stage the control points");
    float* const stage = reinterpret_cast<float*>(win);
    for (uint32_t i = lane; i < nc * 2u * NPTS; i += 32u) stage[i] = ctrl[c0 * 2ull * NPTS + i];
    __syncwarp();
    float step = 0.0f;
    uint32_t cnt = 0u;
    V2 a0 = v2(0.0f, 0.0f), a1 = a0, a2 = a0, a3 = a0, to = a0;
    if (lane < nc) {
        const float* c = stage + lane * 2u * NPTS;
        float nseg;
        if (NPTS == 4) {
            const V2 from = v2(c[0], c[1]), c1 = v2(c[2], c[3]), c2 = v2(c[4], c[5]);
            to = v2(c[6], c[7]);
            nseg = cubic_wang(from, c1, c2, to, tol);
            const CubicPoly q = cubic_poly(from, c1, c2, to);
            a0 = q.a0; a1 = q.a1; a2 = q.a2; a3 = q.a3;
        } else {
            const V2 from = v2(c[0], c[1]), cc = v2(c[2], c[3]);
            to = v2(c[4], c[5]);
            nseg = quad_wang(from, cc, to, tol);
            const QuadPoly q = quad_poly(from, cc, to);
            a0 = q.a0; a1 = q.a1; a2 = q.a2;
        }
        step = 1.0f / nseg;
        cnt = to_u32_or_1(nseg);
    }
    __syncwarp();  // `stage` (= the window) is free again
    unsigned long long base = 0ull, total = 0ull, my0 = 0ull;  // my0: warp-relative index of my first point
    if (FUSED) {
        unsigned long long inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += y;
        }
        total = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 0u) lookback_publish(fu.desc, tile, total);
        my0 = inc - cnt;
    } else {
        const unsigned long long o0 = offsets[c0], o1 = offsets[c0 + nc];
        base = o0;
        total = o1 - o0;
        my0 = lane < nc ? offsets[c0 + lane] - o0 : 0ull;
    }
    constexpr uint32_t window = WITH_T ? FLATW_WINDOW / 2 : FLATW_WINDOW;
    float* const win_t = reinterpret_cast<float*>(win + FLATW_WINDOW / 2);
    float2* gout = out + base;
    float* gt = WITH_T ? t_out + base : nullptr;
    bool resolved = !FUSED;
    float t = 0.0f;
    uint32_t k = 0u;
    for (unsigned long long w0 = 0; w0 < total; w0 += window) {
        const unsigned long long w1 = min(total, w0 + (unsigned long long)window);
        {  // ---- phase A (lane / curve): my points that fall into the window [w0, w1)
            const unsigned long long pos0 = my0 + k, pend = min(my0 + (unsigned long long)cnt, w1);
            if (k < cnt && pos0 < pend) {
                const uint32_t i0 = (uint32_t)(pos0 - w0), nk = (uint32_t)(pend - pos0);
                const bool has_last = k + nk == cnt;
                const uint32_t ninter = has_last ? nk - 1u : nk;
                float2* const wp = win + i0;
                float* const wt = win_t + i0;
#pragma unroll 4
                for (uint32_t j = 0; j < ninter; ++j) {
                    t = t + step;
                    float u = t;
                    V2 pos = a0 + a1 * u;
                    u = u * t;
                    pos = pos + a2 * u;
                    if (NPTS == 4) {
                        u = u * t;
                        pos = pos + a3 * u;
                    }
                    wp[j] = make_float2(pos.x, pos.y);
                    if (WITH_T) wt[j] = t;
                }
                if (has_last) {
                    wp[ninter] = make_float2(to.x, to.y);
                    if (WITH_T) wt[ninter] = 1.0f;
                }
                k += nk;
            }
        }
        if (FUSED && !resolved) {
            base = lookback_exclusive(fu.desc, tile, total, true);
            base = __shfl_sync(0xffffffffu, base, 0);
            resolved = true;
            if (lane < nc) fu.offsets_out[c0 + lane] = base + my0;
            if (lane == nc - 1u && c0 + nc == n) {
                fu.offsets_out[n] = base + total;
                fu.totals->flat_points = base + total;
            }
            if (base + total > fu.capacity) {
                if (lane == 0u) atomicOr(&fu.totals->error, 0x100u);  // point buffer too small: counted, not emitted
                return;
            }
            gout = out + base;
            gt = WITH_T ? t_out + base : nullptr;
        }
        __syncwarp();
        // ---- phase B (lane / two points): the window leaves the SM as 16-byte stores, 512 contiguous bytes per instruction
        const uint32_t wn = (uint32_t)(w1 - w0);
        {
            float2* const g = gout + w0;
            const uint32_t head = (uint32_t)((reinterpret_cast<uintptr_t>(g) >> 3) & 1u) & (wn ? 1u : 0u);
            if (lane == 0u && head) __stcs(g, win[0]);
            const uint32_t pairs = (wn - head) >> 1;
            for (uint32_t i = lane; i < pairs; i += 32u) {
                const float2 p0 = win[head + 2u * i], p1 = win[head + 2u * i + 1u];
                __stcs(reinterpret_cast<float4*>(g + head) + i, make_float4(p0.x, p0.y, p1.x, p1.y));
            }
            if (lane == 31u && ((wn - head) & 1u)) __stcs(g + wn - 1u, win[wn - 1u]);
            if (WITH_T)
                for (uint32_t i = lane; i < wn; i += 32u) __stcs(gt + w0 + i, win_t[i]);
        }
        __syncwarp();
    }
    if (FUSED && !resolved) {  // (a warp without points cannot happen: every curve has at least one; kept for safety)
        base = lookback_exclusive(fu.desc, tile, total, true);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < nc) fu.offsets_out[c0 + lane] = base + my0;
        if (lane == nc - 1u && c0 + nc == n) {
            fu.offsets_out[n] = base + total;
            fu.totals->flat_points = base + total;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Path-level flattening: the `Flattened` adapter (crates/path/src/iterator.rs:125-226) over a whole batch.  Every
// QuadraticTo / CubicTo becomes the LineTo sequence of its flattened points (for_each_flattened: plain Wang counts, no
// sharp-turn split -- that one belongs to the stroker); Begin / LineTo / Close / End pass through.  Custom attributes
// (two per point slot after every endpoint, path.rs:650-665) are interpolated along the curves as the `Flattened` builder
// adapter does (crates/path/src/private.rs:56-122): prev[i] * (1 - t) + attr[i] * t, the curve's own where t == 1.0.
// One warp per path in both passes; a lane owns a verb and writes its points in order.
// ---------------------------------------------------------------------------------------------------------
struct PfRec {  // per input verb, written by the count pass
    float n;          // Wang count of a curve (f32, possibly non-integer), 1 otherwise
    uint32_t voff;    // first output verb, relative to the path's first output verb
    uint32_t poff;    // first output point, relative to the path's first output point
    uint32_t pt;      // first stored input point of the verb
};
__device__ __forceinline__ uint32_t pf_count(uint8_t verb, const V2* points, uint32_t pt, uint32_t a, float tol, float* n_out) {
    float n = 1.0f;
    if (verb == VERB_CUBIC) n = cubic_wang(seg_from_point(points, pt, a), points[pt], points[pt + 1], points[pt + 2], tol);
    else if (verb == VERB_QUAD) n = quad_wang(seg_from_point(points, pt, a), points[pt], points[pt + 1], tol);
    *n_out = n;
    return (verb == VERB_CUBIC || verb == VERB_QUAD) ? to_u32_or_1(n) : 1u;
}
__global__ void __launch_bounds__(256) k_pf_count(const uint8_t* __restrict__ verbs, const V2* __restrict__ points,
                                                  const uint32_t* __restrict__ path_verb_begin,
                                                  const uint32_t* __restrict__ path_point_begin, uint32_t n_paths,
                                                  uint32_t attr_stride, uint32_t pstride, float tol, PfRec* __restrict__ rec,
                                                  uint32_t* __restrict__ path_nv,
                                                  uint32_t* __restrict__ path_np, Totals* totals) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t path = warp; path < n_paths; path += nwarps) {
        const uint32_t vb = path_verb_begin[path], ve = path_verb_begin[path + 1];
        uint32_t pt_run = path_point_begin[path];
        unsigned long long v_run = 0ull, p_run = 0ull;
        for (uint32_t v0 = vb; v0 < ve; v0 += 32u) {
            const uint32_t v = v0 + lane;
            const bool valid = v < ve;
            const uint8_t verb = valid ? verbs[v] : (uint8_t)VERB_END;
            const uint32_t npts = valid ? verb_point_count(verb, attr_stride) : 0u;
            uint32_t inc = npts;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (unsigned)o) inc += y;
            }
            const uint32_t pt = pt_run + inc - npts;
            float n = 1.0f;
            const uint32_t cnt = valid ? pf_count(verb, points, pt, attr_stride, tol, &n) : 0u;
            // End carries no point; every other output verb one endpoint and its attribute slots
            const uint32_t pcnt = (valid && verb == VERB_END) ? 0u : cnt * pstride;  // (pstride 1: PathMeasurements edges)
            unsigned long long iv = cnt, ip = pcnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xffffffffu, iv, o), z = __shfl_up_sync(0xffffffffu, ip, o);
                if (lane >= (unsigned)o) {
                    iv += y;
                    ip += z;
                }
            }
            if (valid) {
                *reinterpret_cast<uint4*>(rec + v) =
                    make_uint4(__float_as_uint(n), (uint32_t)(v_run + iv - cnt), (uint32_t)(p_run + ip - pcnt), pt);
            }
            pt_run += __shfl_sync(0xffffffffu, inc, 31);
            v_run += __shfl_sync(0xffffffffu, iv, 31);
            p_run += __shfl_sync(0xffffffffu, ip, 31);
        }
        if (lane == 0u) {
            if (v_run > 0xFFFFFFFFull || p_run > 0xFFFFFFFFull) atomicOr(&totals->error, 2u);
            path_nv[path] = (uint32_t)v_run;
            path_np[path] = (uint32_t)p_run;
            if (pt_run != path_point_begin[path + 1]) atomicOr(&totals->error, 4u);
        }
    }
}
__global__ void __launch_bounds__(256) k_pf_emit(const uint8_t* __restrict__ verbs, const V2* __restrict__ points,
                                                 const uint32_t* __restrict__ path_verb_begin,
                                                 const uint32_t* __restrict__ path_point_begin, uint32_t n_paths,
                                                 uint32_t num_attributes, const PfRec* __restrict__ rec,
                                                 const unsigned long long* __restrict__ out_v0,
                                                 const unsigned long long* __restrict__ out_p0, uint8_t* __restrict__ out_verbs,
                                                 V2* __restrict__ out_points, uint32_t* __restrict__ out_path_verb_begin,
                                                 uint32_t* __restrict__ out_path_point_begin) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t path = warp; path < n_paths; path += nwarps) {
        const uint32_t vb = path_verb_begin[path], ve = path_verb_begin[path + 1];
        const unsigned long long v0 = out_v0[path], p0 = out_p0[path];
        if (lane == 0u) {
            out_path_verb_begin[path] = (uint32_t)v0;
            out_path_point_begin[path] = (uint32_t)p0;
            if (path + 1u == n_paths) {
                out_path_verb_begin[n_paths] = (uint32_t)out_v0[n_paths];
                out_path_point_begin[n_paths] = (uint32_t)out_p0[n_paths];
            }
        }
        const uint32_t A = num_attributes, a = (A + 1u) / 2u;
        for (uint32_t v = vb + lane; v < ve; v += 32u) {
            const uint8_t verb = verbs[v];
            const uint4 q = *reinterpret_cast<const uint4*>(rec + v);
            const float n = __uint_as_float(q.x);
            uint8_t* ov = out_verbs + v0 + q.y;
            V2* op = out_points + p0 + q.z;
            const uint32_t pt = q.w;
            if (verb == VERB_CUBIC || verb == VERB_QUAD) {
                const bool cubic = verb == VERB_CUBIC;
                const V2 from = seg_from_point(points, pt, a);
                const uint32_t nc = cubic ? 2u : 1u;
                const V2 to = points[pt + nc];
                // attributes of the curve's two endpoints: the slots after `to`, and the slots right before the control points
                const float* at = reinterpret_cast<const float*>(points + pt + nc + 1u);
                const float* ap = reinterpret_cast<const float*>(points + (pt >= a ? pt - a : 0u));
                V2 a1, a2, a3 = v2(0.0f, 0.0f);
                if (cubic) {
                    const CubicPoly c = cubic_poly(from, points[pt], points[pt + 1], to);
                    a1 = c.a1; a2 = c.a2; a3 = c.a3;
                } else {
                    const QuadPoly c = quad_poly(from, points[pt], to);
                    a1 = c.a1; a2 = c.a2;
                }
                const uint32_t cnt = to_u32_or_1(n);
                const float step = 1.0f / n;
                float t = 0.0f;
                for (uint32_t k = 0; k < cnt; ++k) {
                    V2 pos = to;
                    if (k + 1u < cnt) {
                        t = t + step;
                        float u = t;
                        pos = from + a1 * u;
                        u = u * t;
                        pos = pos + a2 * u;
                        if (cubic) {
                            u = u * t;
                            pos = pos + a3 * u;
                        }
                    }
                    ov[k] = VERB_LINE;
                    V2* o = op + (size_t)k * (1u + a);
                    o[0] = pos;
                    if (a) {
                        const float tt = k + 1u < cnt ? t : 1.0f;
                        float* oa = reinterpret_cast<float*>(o + 1);
                        for (uint32_t i = 0; i < 2u * a; ++i)
                            oa[i] = i >= A ? 0.0f : (tt == 1.0f ? at[i] : ap[i] * (1.0f - tt) + at[i] * tt);
                    }
                }
            } else {
                ov[0] = verb;
                if (verb != VERB_END)  // Begin / LineTo / Close (the first point stored again), with the attribute slots
                    for (uint32_t i = 0; i <= a; ++i) op[i] = points[pt + i];
            }
        }
    }
}

// PathMeasurements::initialize (crates/algorithms/src/measure.rs:170-257) for every path of a batch: one Edge {distance,
// index, t} per Begin (distance unchanged, t 1), per LineTo (distance += |from - to|), per flattened segment of a curve
// (for_each_flattened_with_t: distance += |segment|, t = t.end) and per Close (distance += |last - first|); `index` is
// the event's position in the path.  One warp per path: a lane writes the lengths of its verb's edges, then the warp
// replays the f32 running sum in event order -- 32 edges per step, every lane following the same chain through
// shuffles and keeping the value at its own position (the sum is sequential by definition: another order rounds differently).
__global__ void __launch_bounds__(256) k_pm_emit(const uint8_t* __restrict__ verbs, const V2* __restrict__ points,
                                                 const uint32_t* __restrict__ path_verb_begin, uint32_t n_paths, uint32_t attr_stride,
                                                 const PfRec* __restrict__ rec, const unsigned long long* __restrict__ edge0,
                                                 float* __restrict__ distance, uint32_t* __restrict__ index, float* __restrict__ t_out,
                                                 unsigned long long* __restrict__ path_edge_begin, float* __restrict__ path_length) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t a = attr_stride;
    for (uint32_t path = warp; path < n_paths; path += nwarps) {
        const uint32_t vb = path_verb_begin[path], ve = path_verb_begin[path + 1];
        const unsigned long long e0 = edge0[path], e1 = edge0[path + 1];
        if (lane == 0u) {
            path_edge_begin[path] = e0;
            if (path + 1u == n_paths) path_edge_begin[n_paths] = e1;
        }
        for (uint32_t v = vb + lane; v < ve; v += 32u) {
            const uint8_t verb = verbs[v];
            if (verb == VERB_END) continue;  // an open sub-path's End has no edge
            const uint4 q = *reinterpret_cast<const uint4*>(rec + v);
            const float n = __uint_as_float(q.x);
            const unsigned long long e = e0 + q.z;
            const uint32_t pt = q.w, ev = v - vb;
            if (verb == VERB_CUBIC || verb == VERB_QUAD) {
                const bool cubic = verb == VERB_CUBIC;
                const V2 from = seg_from_point(points, pt, a);
                const V2 to = points[pt + (cubic ? 2u : 1u)];
                V2 a1, a2, a3 = v2(0.0f, 0.0f);
                if (cubic) {
                    const CubicPoly c = cubic_poly(from, points[pt], points[pt + 1], to);
                    a1 = c.a1; a2 = c.a2; a3 = c.a3;
                } else {
                    const QuadPoly c = quad_poly(from, points[pt], to);
                    a1 = c.a1; a2 = c.a2;
                }
                const uint32_t cnt = to_u32_or_1(n);
                const float step = 1.0f / n;
                float t = 0.0f;
                V2 prev = from;
                for (uint32_t k = 0; k < cnt; ++k) {
                    V2 pos = to;
                    float te = 1.0f;
                    if (k + 1u < cnt) {
                        t = t + step;
                        te = t;
                        float u = t;
                        pos = from + a1 * u;
                        u = u * t;
                        pos = pos + a2 * u;
                        if (cubic) {
                            u = u * t;
                            pos = pos + a3 * u;
                        }
                    }
                    distance[e + k] = length(pos - prev);
                    index[e + k] = ev;
                    t_out[e + k] = te;
                    prev = pos;
                }
            } else {
                // Begin: distance unchanged (+0 is exact); LineTo / Close: |from - to| with `to` the stored point
                distance[e] = verb == VERB_BEGIN ? 0.0f : length(seg_from_point(points, pt, a) - points[pt]);
                index[e] = ev;
                t_out[e] = 1.0f;
            }
        }
        __threadfence_block();
        __syncwarp();
        float acc = 0.0f;
        for (unsigned long long b = e0; b < e1; b += 32ull) {
            const bool in = b + lane < e1;
            const float x = in ? distance[b + lane] : 0.0f;
            float mine = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                acc = acc + __shfl_sync(0xffffffffu, x, j);  // (+0 beyond the end: exact)
                if ((int)lane == j) mine = acc;
            }
            if (in) distance[b + lane] = mine;
        }
        if (lane == 0u) path_length[path] = acc;
        __syncwarp();
    }
}

// Global index rebasing of a mesh emitted with vertex_base 0 (multi-GPU: a rank learns the number of vertices emitted
// by the ranks before it only after its own single-pass emit): indices[i] += base, 16 bytes per thread and iteration.
// ---- interpolated attributes (StrokeVertex::interpolated_attributes, stroke.rs:2728-2747) ------------------------------
// An Edge source names the curve's two endpoints; the vertex keeps `to` and t, and `from` is the endpoint stored before
// the curve's control points: to - (control points + 1) - attribute slots.  k_attr_curve_kind marks every curve endpoint
// with its verb's point count (2 quadratic, 3 cubic), one warp per path like the slot pre-pass.
__global__ void __launch_bounds__(256) k_attr_curve_kind(const uint8_t* __restrict__ verbs, const uint32_t* __restrict__ path_verb_begin,
                                                         const uint32_t* __restrict__ path_point_begin, uint32_t n_paths,
                                                         uint32_t attr_stride, uint8_t* __restrict__ kind) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t path = warp; path < n_paths; path += nwarps) {
        const uint32_t vb = path_verb_begin[path], ve = path_verb_begin[path + 1];
        uint32_t pt_run = path_point_begin[path];
        for (uint32_t v0 = vb; v0 < ve; v0 += 32u) {
            const uint32_t v = v0 + lane;
            const uint8_t verb = v < ve ? verbs[v] : (uint8_t)VERB_END;
            const uint32_t npts = v < ve ? verb_point_count(verb, attr_stride) : 0u;
            uint32_t inc = npts;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (unsigned)o) inc += y;
            }
            const uint32_t pt = pt_run + inc - npts;
            if (verb == VERB_QUAD) kind[pt + 1u] = 2;
            else if (verb == VERB_CUBIC) kind[pt + 2u] = 3;
            pt_run += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
}
// One thread per vertex: its path by binary search in the per-path vertex ranges, then Endpoint{id} -> the endpoint's
// stored attributes (path.rs:650-665), Edge{from, to, t} -> a[i] * (1 - t) + b[i] * t (unfused, like the reference).
__global__ void __launch_bounds__(256) k_interpolate_attributes(const GpuVertex* __restrict__ vertices, unsigned long long n_vertices,
                                                                const unsigned long long* __restrict__ path_vertex_end, uint32_t n_paths,
                                                                const float* __restrict__ points,
                                                                const uint32_t* __restrict__ path_point_begin,
                                                                const uint8_t* __restrict__ kind, uint32_t num_attributes,
                                                                uint32_t attr_stride, float* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vertices; v += stride) {
        uint32_t lo = 0u, hi = n_paths;  // first path whose range ends beyond v
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (path_vertex_end[mid] <= v) lo = mid + 1u; else hi = mid;
        }
        float* o = out + v * num_attributes;
        if (lo >= n_paths) {
            for (uint32_t i = 0; i < num_attributes; ++i) o[i] = 0.0f;
            continue;
        }
        const GpuVertex vx = vertices[v];
        const unsigned long long to = (unsigned long long)path_point_begin[lo] + (vx.src & 0x7FFFFFFFu);
        const float* b = points + 2ull * (to + 1ull);
        if (vx.t == 1.0f) {
            for (uint32_t i = 0; i < num_attributes; ++i) o[i] = b[i];
        } else {
            const unsigned long long from = to - kind[to] - attr_stride;
            const float* a = points + 2ull * (from + 1ull);
            const float t = vx.t, one_t = 1.0f - t;
            for (uint32_t i = 0; i < num_attributes; ++i) o[i] = a[i] * one_t + b[i] * t;
        }
    }
}

__global__ void __launch_bounds__(256) k_rebase_indices(uint32_t* __restrict__ idx, unsigned long long n, uint32_t base) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long head = min(n, (unsigned long long)((16u - ((uintptr_t)idx & 15u)) & 15u) / 4ull);  // to 16-byte alignment
    if (i < head) idx[i] += base;
    uint4* v = reinterpret_cast<uint4*>(idx + head);
    const unsigned long long nv = (n - head) / 4ull;
    for (unsigned long long k = i; k < nv; k += stride) {
        uint4 q = v[k];
        q.x += base;
        q.y += base;
        q.z += base;
        q.w += base;
        v[k] = q;
    }
    const unsigned long long tail0 = head + nv * 4ull;
    if (tail0 + i < n && i < 4ull) idx[tail0 + i] += base;
}

// ---------------------------------------------------------------------------------------------------------
// Variable line width (stroke_vw.cuh): one thread per path runs the streaming state machine, once to count, once to emit.
// ---------------------------------------------------------------------------------------------------------
#ifndef LYON_VW_MINB
#define LYON_VW_MINB 3  // resident 128-thread blocks per SM the variable-width kernels are compiled for (168 registers; C4: 28.6 ms uncapped, 22.2 ms at 3, 27.0 ms at 4)
#endif
__global__ void __launch_bounds__(128, LYON_VW_MINB) k_vw_count(const uint8_t* __restrict__ verbs, const V2* __restrict__ points,
                                                  const uint32_t* __restrict__ path_verb_begin,
                                                  const uint32_t* __restrict__ path_point_begin, uint32_t n_paths, StrokeParams p,
                                                  uint32_t* __restrict__ path_nv, uint32_t* __restrict__ path_nt) {
    const uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
    if (path >= n_paths) return;
    CountSink sink;
    sink.nv = sink.nt = 0u;
    sink.base = 0u;
    vw_stroke_path(verbs, path_verb_begin[path], path_verb_begin[path + 1], points, path_point_begin[path], p, sink);
    path_nv[path] = sink.nv;
    path_nt[path] = sink.nt;
}
// Exclusive scan of u32 counts into u64 offsets (n + 1 entries), decoupled look-back over tiles of BLOCK * 4 items.
__global__ void __launch_bounds__(BLOCK) k_scan_counts(const uint32_t* __restrict__ in, uint32_t n,
                                                       unsigned long long* __restrict__ out, uint32_t* ticket, uint64_t* desc) {
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base, s_sum[WARPS];
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t i0 = (tile * BLOCK + threadIdx.x) * 4u;
    uint32_t x[4];
    unsigned long long sum = 0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        x[k] = (i0 + k < n) ? in[i0 + k] : 0u;
        sum += x[k];
    }
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += y;
    }
    if (lane == 31u) s_sum[wid] = inc;
    __syncthreads();
    unsigned long long wbase = 0ull, agg = 0ull;
    for (int k = 0; k < WARPS; ++k) {
        if (k < (int)wid) wbase += s_sum[k];
        agg += s_sum[k];
    }
    if (threadIdx.x < 32) {
        const uint64_t e = lookback_exclusive(desc, tile, agg);
        if (threadIdx.x == 0) s_base = e;
    }
    __syncthreads();
    unsigned long long run = s_base + wbase + inc - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n) out[i0 + k] = run;
        run += x[k];
        if (i0 + k + 1 == n) out[n] = run;
    }
}
__global__ void __launch_bounds__(128, LYON_VW_MINB) k_vw_emit(const uint8_t* __restrict__ verbs, const V2* __restrict__ points,
                                                 const uint32_t* __restrict__ path_verb_begin,
                                                 const uint32_t* __restrict__ path_point_begin, uint32_t n_paths, StrokeParams p,
                                                 const unsigned long long* __restrict__ path_v0,
                                                 const unsigned long long* __restrict__ path_t0, GpuVertex* __restrict__ vertices,
                                                 uint32_t* __restrict__ indices, uint32_t vertex_base,
                                                 unsigned long long* __restrict__ path_vertex_end,
                                                 unsigned long long* __restrict__ path_index_end, Totals* totals) {
    const uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
    if (path >= n_paths) return;
    const unsigned long long v0 = path_v0[path], t0 = path_t0[path];
    EmitSink sink;
    sink.vtx = vertices + v0;
    sink.idx = indices + 3ull * t0;
    sink.nv = sink.nt = 0u;
    sink.base = vertex_base + (uint32_t)v0;
    vw_stroke_path(verbs, path_verb_begin[path], path_verb_begin[path + 1], points, path_point_begin[path], p, sink);
    if (sink.nv != path_v0[path + 1] - v0 || sink.nt != path_t0[path + 1] - t0) atomicOr(&totals->error, 1u);
    if (path_vertex_end) path_vertex_end[path] = v0 + sink.nv;
    if (path_index_end) path_index_end[path] = 3ull * (t0 + sink.nt);
}

#include "tile_kernels.cuh"
#include "wtile_kernels.cuh"

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// Context / workspace
// ---------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};
struct lyon_cuda_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    // workspace
    DevBuf seg_pt, seg_off, seg_tsplit, seg_id, pos, t, segfs, len, adv, voff, toff, jcode, full_list, sp_end_slot, desc, misc;
    uint32_t n_subpaths = 0;
    unsigned persistent_blocks = 148 * 8;
    int num_sms = 148;
    size_t max_smem_optin = 0;
    Totals* h_totals = nullptr;  // pinned
    // fused tile pipeline
    DevBuf verb_n, path_slots, path_events, path_slot_begin, path_event_begin, tile_first, tile_desc, tile_base, tile_misc;
    // warp-tile pipeline (wtile_kernels.cuh)
    DevBuf split_rec, path_split_begin, end_rec;
    uint32_t split_cap = 0, end_rec_cap = 0;
    bool tile_is_wtile = false;   // the tile table of the last pre-pass is for k_stroke_wtile
    bool tile_has_split_recs = false;  // the pre-pass kept the plan of every split quadratic
    int wt_warps = 0;             // warps per CTA (= tiles in flight per SM)
    uint32_t wt_s_cap = 0, wt_v_cap = 0, wt_p_cap = 0, wt_bytes = 0;
    bool flat_block_tiles = true;  // LYON_FLAT=warp: the warp-tile flatten kernel (k_flat_emit_w; measured slower) instead of block tiles
    int wt_mode = 0;              // 1 (LYON_PIPE=wtile / set_pipeline 3): warp tiles whenever they fit; 0: CTA tiles
    TileTotals* h_tile_totals = nullptr;  // pinned
    bool tile_planned = false;            // the last plan went through the tile pipeline
    bool vw_planned = false;              // the last plan went through the variable-width (thread per path) kernels
    DevBuf vw_nv, vw_nt, vw_v0, vw_t0;
    DevBuf attr_kind;  // lyon_cuda_interpolate_attributes: per point, the point count of the curve ending there
    DevBuf pf_rec;  // path-level flattening: per-verb records (counts reuse vw_nv / vw_nt / vw_v0 / vw_t0)
    bool pf_planned = false;
    bool pf_measure = false;  // ... by lyon_cuda_measure_paths_count (edges) rather than lyon_cuda_flatten_paths_count
    lyon_path_batch pf_batch{};
    lyon_path_batch vw_batch{};
    lyon_path_batch tile_batch{};
    uint32_t tile_n_tiles = 0, tile_s_cap = 0, tile_w_v = 0, tile_w_t = 0, tile_quantum = 0;
    int tile_nt = 256;
    int tile_force_general = 0;  // test hook: lyon_cuda_set_pipeline
    // optional device timing of the dominant kernel (lyon_cuda_enable_timing)
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    // state of the last plan
    bool planned = false;
    Views views{};
    StrokeParams params{};
    const uint32_t* d_path_verb_begin = nullptr;
    uint32_t n_paths = 0;
    uint64_t n_vertices = 0, n_triangles = 0;
    size_t ws_bytes = 0;
    struct HostPipe* host_pipe = nullptr;  // lyon_cuda_tessellate_paths_into (host_pipeline.cuh)
    struct CommState* comm = nullptr;      // lyon_cuda_comm_init (sharded.cuh)
};
static void host_pipe_destroy(struct HostPipe* hp);
static void comm_state_destroy(struct CommState* s);

namespace {

#define CU_TRY(ctx, expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);               \
            return _e == cudaErrorMemoryAllocation ? LYON_ERR_OUT_OF_MEMORY : LYON_ERR_CUDA; \
        }                                                                                  \
    } while (0)

int ensure(lyon_cuda_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return LYON_OK;
    if (b.p) {
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        CU_TRY(c, cudaFree(b.p));
        c->ws_bytes -= b.cap;
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;  // grow-only with a little slack
    CU_TRY(c, cudaMalloc(&b.p, want));
    b.cap = want;
    c->ws_bytes += want;
    return LYON_OK;
}
#define TRY(expr)                  \
    do {                           \
        int _s = (expr);           \
        if (_s != LYON_OK) return _s; \
    } while (0)

inline unsigned grid_for(uint64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// misc buffer layout: [0] Totals, [256..] tickets (4 x u32)
Totals* d_totals(lyon_cuda_ctx* c) { return reinterpret_cast<Totals*>(c->misc.p); }
uint32_t* d_ticket(lyon_cuda_ctx* c, int i) { return reinterpret_cast<uint32_t*>(static_cast<char*>(c->misc.p) + 256) + i; }

int fetch_totals(lyon_cuda_ctx* c) {
    CU_TRY(c, cudaMemcpyAsync(c->h_totals, d_totals(c), sizeof(Totals), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return LYON_OK;
}

StrokeParams make_params(const lyon_stroke_options& o, uint32_t num_attributes) {
    StrokeParams p{};
    p.start_cap = o.start_cap;
    p.end_cap = o.end_cap;
    p.line_join = o.line_join;
    p.line_width = o.line_width;
    p.half_width = o.line_width * 0.5f;
    p.miter_limit = o.miter_limit;
    p.tolerance = o.tolerance;
    // stroke.rs:603-605
    p.merge_thr = fmaxf(fminf(o.tolerance * o.tolerance * 0.5f, o.line_width * o.line_width * 0.05f), 1e-8f);
    p.flag_thr = p.merge_thr * 5.0f;
    p.arc_step = circle_flattening_step(p.half_width, p.tolerance);  // host libm, same as the oracle
    p.num_attributes = num_attributes;
    p.attr_stride = (num_attributes + 1u) / 2u;
    p.vw_attr = num_attributes > 0 ? o.variable_line_width : -1;  // tessellate() forces it off, stroke.rs:878
    return p;
}

template <int NPTS>
int flatten_count(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol, uint64_t* d_offsets, uint64_t* total) {
    if (!c || !total || (n && (!d_ctrl || !d_offsets))) return LYON_ERR_INVALID_ARGUMENT;
    if (!(tol > 0.0f)) return LYON_ERR_UNSUPPORTED_PARAMETER;  // ToleranceIsNaN / non-positive
    *total = 0;
    if (n == 0) {
        if (d_offsets) CU_TRY(c, cudaMemsetAsync(d_offsets, 0, 8, c->stream));
        return LYON_OK;
    }
    const unsigned tiles = grid_for(n, BLOCK);
    TRY(ensure(c, c->misc, 512));
    TRY(ensure(c, c->desc, (size_t)tiles * 8));
    CU_TRY(c, cudaMemsetAsync(c->misc.p, 0, 512, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->desc.p, 0, (size_t)tiles * 8, c->stream));
    k_flat_count<NPTS><<<tiles, BLOCK, 0, c->stream>>>(d_ctrl, n, tol, (unsigned long long*)d_offsets, d_ticket(c, 0),
                                                      (uint64_t*)c->desc.p, d_totals(c));
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    TRY(fetch_totals(c));
    *total = c->h_totals->flat_points;
    return LYON_OK;
}
template <int NPTS>
int flatten_emit(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol, const uint64_t* d_offsets, uint64_t total,
                 float* d_points, float* d_t) {
    if (!c || (n && (!d_ctrl || !d_offsets)) || (total && !d_points)) return LYON_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(d_points) | reinterpret_cast<uintptr_t>(d_ctrl)) & 7u) return LYON_ERR_INVALID_ARGUMENT;
    if (n == 0 || total == 0) return LYON_OK;
    if (!c->flat_block_tiles) {  // warp-owned tiles (experiment)
        const unsigned blocks = grid_for(n, 32 * FLATW_WARPS);
        if (d_t)
            k_flat_emit_w<NPTS, false, true><<<blocks, FLATW_WARPS * 32, 0, c->stream>>>(
                d_ctrl, n, tol, (const unsigned long long*)d_offsets, (float2*)d_points, d_t, FlatFused{});
        else
            k_flat_emit_w<NPTS, false, false><<<blocks, FLATW_WARPS * 32, 0, c->stream>>>(
                d_ctrl, n, tol, (const unsigned long long*)d_offsets, (float2*)d_points, d_t, FlatFused{});
        c->launches++;
        CU_TRY(c, cudaGetLastError());
        return LYON_OK;
    }
    const size_t smem = sizeof(FlatSmem<NPTS>);
    if (d_t) {
        CU_TRY(c, (cudaFuncSetAttribute(k_flat_emit<NPTS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        k_flat_emit<NPTS, false, true><<<grid_for(n, FLAT_CURVES), FLAT_CURVES, smem, c->stream>>>(
            d_ctrl, n, tol, (const unsigned long long*)d_offsets, (float2*)d_points, d_t, FlatFused{});
    } else {
        CU_TRY(c, (cudaFuncSetAttribute(k_flat_emit<NPTS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        k_flat_emit<NPTS, false, false><<<grid_for(n, FLAT_CURVES), FLAT_CURVES, smem, c->stream>>>(
            d_ctrl, n, tol, (const unsigned long long*)d_offsets, (float2*)d_points, d_t, FlatFused{});
    }
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}
// One pass: count + look-back + emit into caller-kept buffers (see k_flat_emit<NPTS, FUSED = true, WITH_T>).
template <int NPTS>
int flatten_fused(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol, uint64_t* d_offsets, float* d_points,
                  uint64_t capacity, float* d_t, uint64_t* total) {
    if (!c || !total || (n && (!d_ctrl || !d_offsets)) || (capacity && !d_points)) return LYON_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(d_points) | reinterpret_cast<uintptr_t>(d_ctrl)) & 7u) return LYON_ERR_INVALID_ARGUMENT;
    if (!(tol > 0.0f)) return LYON_ERR_UNSUPPORTED_PARAMETER;  // ToleranceIsNaN / non-positive
    *total = 0;
    if (n == 0) {
        if (d_offsets) CU_TRY(c, cudaMemsetAsync(d_offsets, 0, 8, c->stream));
        return LYON_OK;
    }
    const bool wtiles = !c->flat_block_tiles;  // warp-owned tiles (experiment): one look-back descriptor per 32 curves
    const unsigned tiles = wtiles ? grid_for(n, 32) : grid_for(n, FLAT_CURVES);
    TRY(ensure(c, c->misc, 512));
    TRY(ensure(c, c->desc, (size_t)tiles * 8));
    CU_TRY(c, cudaMemsetAsync(c->misc.p, 0, 512, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->desc.p, 0, (size_t)tiles * 8, c->stream));
    const size_t smem = sizeof(FlatSmem<NPTS>);
    if (!wtiles) {
        CU_TRY(c, (cudaFuncSetAttribute(k_flat_emit<NPTS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        CU_TRY(c, (cudaFuncSetAttribute(k_flat_emit<NPTS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    }
    FlatFused fu{(unsigned long long*)d_offsets, capacity, d_ticket(c, 0), (uint64_t*)c->desc.p, d_totals(c)};
    if (c->timing) CU_TRY(c, cudaEventRecord(c->ev0, c->stream));
    if (wtiles) {
        const unsigned blocks = grid_for(n, 32 * FLATW_WARPS);
        if (d_t) k_flat_emit_w<NPTS, true, true><<<blocks, FLATW_WARPS * 32, 0, c->stream>>>(d_ctrl, n, tol, nullptr, (float2*)d_points, d_t, fu);
        else k_flat_emit_w<NPTS, true, false><<<blocks, FLATW_WARPS * 32, 0, c->stream>>>(d_ctrl, n, tol, nullptr, (float2*)d_points, d_t, fu);
    } else if (d_t) k_flat_emit<NPTS, true, true><<<tiles, FLAT_CURVES, smem, c->stream>>>(d_ctrl, n, tol, nullptr, (float2*)d_points, d_t, fu);
    else k_flat_emit<NPTS, true, false><<<tiles, FLAT_CURVES, smem, c->stream>>>(d_ctrl, n, tol, nullptr, (float2*)d_points, d_t, fu);
    if (c->timing) {
        CU_TRY(c, cudaEventRecord(c->ev1, c->stream));
        c->ev_valid = true;
    }
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    TRY(fetch_totals(c));
    *total = c->h_totals->flat_points;
    const bool too_small = (c->h_totals->error & 0x100u) != 0u;
    CU_TRY(c, cudaMemsetAsync(&d_totals(c)->error, 0, 4, c->stream));  // the flag is not an emit/count mismatch
    return too_small ? LYON_ERR_BUFFER_TOO_SMALL : LYON_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Fused tile pipeline, host side
// ---------------------------------------------------------------------------------------------------------

// tile_misc layout: [0] TileTotals, [256..] tickets (4 x u32)
TileTotals* d_tile_totals(lyon_cuda_ctx* c) { return reinterpret_cast<TileTotals*>(c->tile_misc.p); }
uint32_t* d_tile_ticket(lyon_cuda_ctx* c, int i) { return reinterpret_cast<uint32_t*>(static_cast<char*>(c->tile_misc.p) + 256) + i; }

int fetch_tile_totals(lyon_cuda_ctx* c) {
    CU_TRY(c, cudaMemcpyAsync(c->h_tile_totals, d_tile_totals(c), sizeof(TileTotals), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return LYON_OK;
}

struct TileGeom {
    int nt;                   // threads per CTA (128 / 256 / 512)
    uint32_t s_cap, w_v, w_t;  // slot capacity (w_v / w_t: unused staging sizes, kept 0)
};
// Tile geometry for a batch whose longest path feeds `max_path` slots and that feeds `total` slots in all.
// Default: 256 threads and 3520 slots per CTA, two CTAs per SM (each hides the other's serial phases; measured
// best on C3, see profiles/).  Longer paths: one 512-thread CTA per SM with all its shared memory (~7000 slots).
// The tile quantum shrinks for small batches so that every SM still gets a few tiles.  LYON_TILE=nt,cap,0 overrides.
#ifdef LYON_DEV_NT
#define LYON_DEV_NT_OR_DEFAULT LYON_DEV_NT
#else
#define LYON_DEV_NT_OR_DEFAULT 256
#endif
#ifdef LYON_DEV_BUILD  // experiment builds launch one fixed instantiation (LYON_DEV_NT): LYON_TILE may name any warp multiple
bool dev_nt_ok(unsigned nt) { return nt >= 64u && nt <= 1024u && (nt & 31u) == 0u; }
#else
bool dev_nt_ok(unsigned) { return false; }
#endif
bool tile_config(const lyon_cuda_ctx* c, uint32_t max_path, unsigned long long total, TileGeom* g, uint32_t* quantum) {
    TileGeom pick{};
    bool found = false;
    if (const char* e = getenv("LYON_TILE")) {
        unsigned nt = 0, cap = 0, wv = 0;
        if (sscanf(e, "%u,%u,%u", &nt, &cap, &wv) == 3 && (nt == 128 || nt == 256 || nt == 257 || nt == 512 || dev_nt_ok(nt)) && cap >= 64) {
            pick = TileGeom{(int)nt, cap & ~31u, 0u, 0u};
            found = max_path <= pick.s_cap;
        }
    }
    if (!found && max_path * 2u <= 3520u) {
        pick = TileGeom{256, 3520u, 0u, 0u};
        found = true;
    }
    if (!found) {
        uint32_t cap = (uint32_t)((c->max_smem_optin - 2048) / 33u) & ~31u;
        if (cap > 65000u) cap = 65000u & ~31u;
        while (cap >= 256u && tile_smem_bytes(cap, 0, 0) > c->max_smem_optin) cap -= 32u;
        if (cap < 256u || max_path > cap) return false;
        pick = TileGeom{512, cap, 0u, 0u};
    }
    if (tile_smem_bytes(pick.s_cap, pick.w_v, pick.w_t) > c->max_smem_optin) return false;
    *g = pick;
    uint32_t q = pick.s_cap - (max_path ? max_path : 1u) + 1u;
    const unsigned long long spread = total / (4ull * (unsigned long long)c->num_sms) + 1ull;  // >= 4 tiles per SM
    if (spread < q) q = (uint32_t)(spread < 64ull ? 64ull : spread);
    if (q > pick.s_cap - (max_path ? max_path : 1u) + 1u) q = pick.s_cap - (max_path ? max_path : 1u) + 1u;
    *quantum = q;
    return true;
}

// Warp-tile geometry: the largest number of warps per SM (16 .. 6) whose share of the shared memory holds the longest
// path of the batch, with the rest of the share given to slots (several short paths per tile) at the batch's average
// verbs-per-slot and points-per-slot ratios.  LYON_WT=warps,s_cap,v_cap,p_cap overrides.
bool wtile_config(lyon_cuda_ctx* c, const lyon_path_batch* b, const TileTotals& t) {
    if (c->wt_mode != 1) return false;  // experiment, measured 3x slower than CTA tiles on C3 (profiles/r02_experiments.md)
    if (t.wt_flags & WT_FLAG_NO_SPLIT_RECS) return false;
    const uint32_t back = 1u + (b->num_attributes + 1u) / 2u;
    const uint32_t need_s = t.max_path_slots ? t.max_path_slots : 1u, need_v = t.max_path_verbs ? t.max_path_verbs : 1u;
    const uint32_t need_p = t.max_path_points + back + 2u;
    if (need_s > 60000u || need_p > 60000u) return false;
    if (const char* e = getenv("LYON_WT")) {
        unsigned wq = 0, sc = 0, vc = 0, pc = 0;
        if (sscanf(e, "%u,%u,%u,%u", &wq, &sc, &vc, &pc) == 4 && wq >= 1 && wq <= 16 && sc >= need_s && vc >= need_v && pc >= need_p &&
            sc <= 60000u && pc <= 60000u) {
            const WtLayout L = wt_layout(sc, vc, pc);
            if ((size_t)L.total * wq <= c->max_smem_optin) {
                c->wt_warps = (int)wq;
                c->wt_s_cap = sc;
                c->wt_v_cap = vc;
                c->wt_p_cap = pc;
                c->wt_bytes = L.total;
                return true;
            }
        }
    }
    const double total = t.total_slots ? (double)t.total_slots : 1.0;
    const double rv = (double)b->n_verbs / total, rp = (double)b->n_points / total;
    for (int wq = 16; wq >= 6; --wq) {
        const uint32_t budget = (uint32_t)(c->max_smem_optin / (size_t)wq) & ~127u;
        uint32_t sc = (budget / 20u) & ~31u;
        if (sc > 60000u) sc = 60000u & ~31u;
        for (; sc >= need_s && sc >= 64u; sc -= 32u) {
            uint32_t vc = (uint32_t)((double)sc * rv) + 8u, pc = (uint32_t)((double)sc * rp) + back + 8u;
            if (vc > sc) vc = sc;
            if (vc < need_v) vc = need_v;
            if (pc < need_p) pc = need_p;
            if (pc > 60000u) continue;
            const WtLayout L = wt_layout(sc, vc, pc);
            if (L.total <= budget) {
                c->wt_warps = wq;
                c->wt_s_cap = sc;
                c->wt_v_cap = vc;
                c->wt_p_cap = pc;
                c->wt_bytes = L.total;
                return true;
            }
        }
    }
    return false;
}

// Pre-pass: per-path slot counts, their scan, the tile table.  *usable = the batch fits a tile pipeline; which one is
// left in c->tile_is_wtile (warp tiles when `allow_wtile` and the longest path fits a warp's share of shared memory).
int tile_prepass(lyon_cuda_ctx* c, const lyon_path_batch* b, bool* usable, bool allow_wtile = true) {
    *usable = false;
    c->tile_is_wtile = false;
    if (c->tile_force_general || !b->path_point_begin || b->n_paths == 0 || b->n_verbs == 0) return LYON_OK;
    const uint32_t np = b->n_paths;
    TRY(ensure(c, c->verb_n, ((size_t)b->n_verbs + 1) * 4));
    TRY(ensure(c, c->path_slots, ((size_t)np + 1) * 4));
    TRY(ensure(c, c->path_events, ((size_t)np + 1) * 4));
    TRY(ensure(c, c->path_slot_begin, ((size_t)np + 1) * 8));
    TRY(ensure(c, c->path_event_begin, ((size_t)np + 1) * 4));
    TRY(ensure(c, c->path_split_begin, ((size_t)np + 1) * 4));
    // plans of the split quadratics: one verb in eight is plenty (C3: one in thirty); more => CTA-tile pipeline
    const size_t split_cap = (size_t)b->n_verbs / 8u + 1024u;
    TRY(ensure(c, c->split_rec, split_cap * sizeof(SplitRec)));
    c->split_cap = (uint32_t)(split_cap > 0x7FFFFFFFull ? 0x7FFFFFFFull : split_cap);
    TRY(ensure(c, c->tile_misc, 512));
    const unsigned ptiles = grid_for(np, BLOCK * PSCAN_ITEMS);
    TRY(ensure(c, c->tile_desc, (size_t)ptiles * 16));
    CU_TRY(c, cudaMemsetAsync(c->tile_misc.p, 0, 512, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->tile_desc.p, 0, (size_t)ptiles * 16, c->stream));
    const unsigned warps_needed = np;
    unsigned blocks = (warps_needed + 7u) / 8u;
    if (blocks > (unsigned)c->num_sms * 16u) blocks = (unsigned)c->num_sms * 16u;
    k_path_slots<<<blocks, 256, 0, c->stream>>>(b->verbs, reinterpret_cast<const V2*>(b->points), b->path_verb_begin,
                                               b->path_point_begin, np, c->params, (uint32_t*)c->path_slots.p,
                                               (uint32_t*)c->path_events.p, (float*)c->verb_n.p, (SplitRec*)c->split_rec.p,
                                               c->split_cap, (uint32_t*)c->path_split_begin.p, d_tile_totals(c));
    k_scan_paths<<<ptiles, BLOCK, 0, c->stream>>>((const uint32_t*)c->path_slots.p, (const uint32_t*)c->path_events.p, np,
                                                 (unsigned long long*)c->path_slot_begin.p, (uint32_t*)c->path_event_begin.p,
                                                 d_tile_ticket(c, 0), (uint64_t*)c->tile_desc.p,
                                                 (uint64_t*)c->tile_desc.p + ptiles, d_tile_totals(c));
    c->launches += 2;
    CU_TRY(c, cudaGetLastError());
    TRY(fetch_tile_totals(c));
    if (c->h_tile_totals->error & TILE_ERR_POINTS) {
        c->err = "path_point_begin does not match the points the verbs consume";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    const unsigned long long total = c->h_tile_totals->total_slots;
    const uint32_t max_path = c->h_tile_totals->max_path_slots ? c->h_tile_totals->max_path_slots : 1u;
    uint32_t s_cap = 0, cap = 0, group = 0, v_cap = 0, p_cap = 0;
    c->tile_has_split_recs = !(c->h_tile_totals->wt_flags & WT_FLAG_NO_SPLIT_RECS);
    {
        const unsigned long long n_ends = (unsigned long long)b->n_verbs - c->h_tile_totals->total_events;
        TRY(ensure(c, c->end_rec, (size_t)(n_ends + 1ull) * sizeof(EndRecord)));
        c->end_rec_cap = (uint32_t)(n_ends > 0xFFFFFFFEull ? 0xFFFFFFFEull : n_ends);
    }
    if (allow_wtile && wtile_config(c, b, *c->h_tile_totals)) {
        c->tile_is_wtile = true;
        s_cap = cap = c->wt_s_cap;
        // small batches: shrink the tiles so that every warp of every SM still gets a few
        const unsigned long long spread = total / (4ull * (unsigned long long)c->num_sms * (unsigned long long)c->wt_warps) + 1ull;
        if (spread < cap) cap = (uint32_t)(spread < max_path ? max_path : spread);
        v_cap = c->wt_v_cap;
        p_cap = c->wt_p_cap - (1u + (b->num_attributes + 1u) / 2u) - 2u;
        c->tile_w_v = c->tile_w_t = c->tile_quantum = 0;
    } else {
        TileGeom g;
        uint32_t q;
        if (!tile_config(c, c->h_tile_totals->max_path_slots, total, &g, &q)) return LYON_OK;
        s_cap = g.s_cap;
        c->tile_nt = g.nt;
        c->tile_w_v = g.w_v;
        c->tile_w_t = g.w_t;
        c->tile_quantum = q;
        cap = q + max_path - 1u;  // what a tile of the quantum `q` could hold at most (q shrinks for small batches)
        if (cap > s_cap) cap = s_cap;
#ifdef LYON_DEV_BUILD
        group = (uint32_t)LYON_DEV_NT_OR_DEFAULT / 32u;
#else
        group = (uint32_t)(g.nt & ~31) / 32u;
#endif
        if (const char* e = getenv("LYON_TILE_GROUP")) group = (uint32_t)atoi(e);  // experiments: 0 switches the trimming off
    }
    // tile table: greedy packing inside coarse blocks (k_tile_pack), two passes around a scan of the per-block counts.
    // coarse block = what one thread packs: 32 tiles' worth of slots, fewer for small batches (the packing of a block is
    // sequential, and a batch of a few hundred tiles should still spread over many threads)
    unsigned long long qc = 32ull * cap;
    if (const char* e = getenv("LYON_PACK_QC")) qc = (unsigned long long)std::max(2, atoi(e)) * cap;  // experiments
    if (total / 512ull < qc) qc = total / 512ull;
    if (qc < 2ull * cap) qc = 2ull * cap;
    const unsigned long long n_blocks64 = total / qc + 1ull;
    if (n_blocks64 > 0x3FFFFFFFull) return LYON_OK;
    const uint32_t n_blocks = (uint32_t)n_blocks64;
    TRY(ensure(c, c->vw_nv, ((size_t)n_blocks + 1) * 4));
    TRY(ensure(c, c->vw_v0, ((size_t)n_blocks + 1) * 8));
    const unsigned btiles = grid_for(n_blocks, BLOCK * 4);
    TRY(ensure(c, c->desc, (size_t)btiles * 8));
    TRY(ensure(c, c->misc, 512));
    CU_TRY(c, cudaMemsetAsync(c->misc.p, 0, 512, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->desc.p, 0, (size_t)btiles * 8, c->stream));
    k_tile_pack<<<grid_for(n_blocks, 128), 128, 0, c->stream>>>((const unsigned long long*)c->path_slot_begin.p, np, qc, cap, group, n_blocks,
                                                               (uint32_t*)c->vw_nv.p, nullptr, nullptr, b->path_verb_begin,
                                                               b->path_point_begin, v_cap, p_cap);
    k_scan_counts<<<btiles, BLOCK, 0, c->stream>>>((const uint32_t*)c->vw_nv.p, n_blocks, (unsigned long long*)c->vw_v0.p,
                                                  d_ticket(c, 0), (uint64_t*)c->desc.p);
    unsigned long long n_tiles64 = 0ull;
    CU_TRY(c, cudaMemcpyAsync(&n_tiles64, (unsigned long long*)c->vw_v0.p + n_blocks, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->launches += 2;
    if (n_tiles64 == 0ull || n_tiles64 > 0x7FFFFFFFull) return LYON_OK;
    c->tile_n_tiles = (uint32_t)n_tiles64;
    c->tile_s_cap = s_cap;
    TRY(ensure(c, c->tile_first, ((size_t)c->tile_n_tiles + 1) * 4));
    TRY(ensure(c, c->tile_desc, (size_t)c->tile_n_tiles * 16));
    TRY(ensure(c, c->tile_base, (size_t)c->tile_n_tiles * 16));
    k_tile_pack<<<grid_for(n_blocks, 128), 128, 0, c->stream>>>((const unsigned long long*)c->path_slot_begin.p, np, qc, cap, group, n_blocks,
                                                               nullptr, (const unsigned long long*)c->vw_v0.p, (uint32_t*)c->tile_first.p,
                                                               b->path_verb_begin, b->path_point_begin, v_cap, p_cap);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    c->tile_batch = *b;
    c->vw_planned = false;  // the pre-pass reuses the variable-width / path-flattening count buffers
    c->pf_planned = false;
    *usable = true;
    return LYON_OK;
}

// Everything end() emits: the EndRecords the tile kernel left behind, one thread each.
int ends_launch(lyon_cuda_ctx* c, const TileArgs& a) {
    if (!c->end_rec_cap) return LYON_OK;
    unsigned eb = (c->end_rec_cap + 127u) / 128u;
    if (eb > (unsigned)c->num_sms * 16u) eb = (unsigned)c->num_sms * 16u;
    k_stroke_ends<<<eb, 128, 0, c->stream>>>(a.end_rec, a.n_end_rec, a.end_rec_cap, c->params, a.vertices, a.indices, a.totals);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}

template <int NT, int MODE, int MINB>
int tile_launch_nt(lyon_cuda_ctx* c, const TileArgs& a) {
    size_t smem = tile_smem_bytes(a.s_cap, a.w_v, a.w_t);
#ifdef LYON_DEV_BUILD  // experiments: LYON_TILE_PERSM=1 pads the allocation so that a single CTA fits an SM
    if (const char* e = getenv("LYON_TILE_PERSM"))
        if (atoi(e) == 1 && smem < (size_t)c->max_smem_optin / 2 + 2048) smem = (size_t)c->max_smem_optin / 2 + 2048;
#endif
    auto kern = k_stroke_tile<NT, MODE, MINB>;
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) {
        c->err = "tile kernel does not fit on an SM";
        return LYON_ERR_INTERNAL;
    }
    unsigned grid = (unsigned)c->num_sms * (unsigned)per_sm;
    if (grid > a.n_tiles) grid = a.n_tiles;
    CU_TRY(c, cudaMemsetAsync(a.ticket, 0, 4, c->stream));
    CU_TRY(c, cudaMemsetAsync(a.n_end_rec, 0, 4, c->stream));
    if (MODE != TILE_MODE_EMIT) CU_TRY(c, cudaMemsetAsync(c->tile_desc.p, 0, (size_t)c->tile_n_tiles * 16, c->stream));
    if (c->timing) CU_TRY(c, cudaEventRecord(c->ev0, c->stream));
    kern<<<grid, NT, smem, c->stream>>>(a, c->params);
    if (MODE != TILE_MODE_COUNT) TRY(ends_launch(c, a));
    if (c->timing) {
        CU_TRY(c, cudaEventRecord(c->ev1, c->stream));
        c->ev_valid = true;
    }
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}

// Warp-tile launch: one CTA of wt_warps warps per SM, every warp a tile of its own; then the End records.
template <int MODE>
int wtile_launch(lyon_cuda_ctx* c, const TileArgs& a) {
    WtArgs x{};
    x.s_cap = c->wt_s_cap;
    x.v_cap = c->wt_v_cap;
    x.p_cap = c->wt_p_cap;
    x.warp_bytes = c->wt_bytes;
    const size_t smem = (size_t)c->wt_bytes * (size_t)c->wt_warps;
    auto kern = k_stroke_wtile<MODE>;
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)c->num_sms;
    const unsigned want = (a.n_tiles + (unsigned)c->wt_warps - 1u) / (unsigned)c->wt_warps;
    if (grid > want) grid = want;
    CU_TRY(c, cudaMemsetAsync(a.ticket, 0, 4, c->stream));
    CU_TRY(c, cudaMemsetAsync(a.n_end_rec, 0, 4, c->stream));
    if (MODE != TILE_MODE_EMIT) CU_TRY(c, cudaMemsetAsync(c->tile_desc.p, 0, (size_t)c->tile_n_tiles * 16, c->stream));
    if (c->timing) CU_TRY(c, cudaEventRecord(c->ev0, c->stream));
    kern<<<grid, c->wt_warps * 32, smem, c->stream>>>(a, x, c->params);
    c->launches++;
    if (MODE != TILE_MODE_COUNT) TRY(ends_launch(c, a));
    if (c->timing) {
        CU_TRY(c, cudaEventRecord(c->ev1, c->stream));
        c->ev_valid = true;
    }
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}

template <int MODE>
int tile_launch(lyon_cuda_ctx* c, GpuVertex* d_vertices, uint32_t* d_indices, uint64_t cap_v, uint64_t cap_t, uint64_t vertex_base,
                uint64_t* d_path_vertex_end, uint64_t* d_path_index_end) {
    const lyon_path_batch& b = c->tile_batch;
    TileArgs a{};
    a.verbs = b.verbs;
    a.points = reinterpret_cast<const V2*>(b.points);
    a.verb_n = (const float*)c->verb_n.p;
    a.path_verb_begin = b.path_verb_begin;
    a.path_point_begin = b.path_point_begin;
    a.path_slot_begin = (const unsigned long long*)c->path_slot_begin.p;
    a.path_event_begin = (const uint32_t*)c->path_event_begin.p;
    a.tile_first = (const uint32_t*)c->tile_first.p;
    a.n_tiles = c->tile_n_tiles;
    a.n_paths = b.n_paths;
    a.s_cap = c->tile_s_cap;
    a.w_v = c->tile_w_v;
    a.w_t = c->tile_w_t;
    a.vertices = d_vertices;
    a.indices = d_indices;
    a.cap_v = cap_v;
    a.cap_t = cap_t;
    a.vertex_base = (uint32_t)vertex_base;
    a.desc_v = (uint64_t*)c->tile_desc.p;
    a.desc_t = (uint64_t*)c->tile_desc.p + c->tile_n_tiles;
    a.tile_base_v = (unsigned long long*)c->tile_base.p;
    a.tile_base_t = (unsigned long long*)c->tile_base.p + c->tile_n_tiles;
    a.ticket = d_tile_ticket(c, 1);
    a.totals = d_tile_totals(c);
    a.path_vertex_end = (unsigned long long*)d_path_vertex_end;
    a.path_index_end = (unsigned long long*)d_path_index_end;
    a.path_split_begin = c->tile_has_split_recs ? (const uint32_t*)c->path_split_begin.p : nullptr;
    a.split_rec = c->tile_has_split_recs ? (const SplitRec*)c->split_rec.p : nullptr;
    a.end_rec = (EndRecord*)c->end_rec.p;
    a.n_end_rec = &d_tile_totals(c)->n_end_rec;
    a.end_rec_cap = c->end_rec_cap;
    if (c->tile_is_wtile) return wtile_launch<MODE>(c, a);
#ifdef LYON_DEV_BUILD  // experiments only: one instantiation, compiles 4x faster
#ifndef LYON_DEV_NT
#define LYON_DEV_NT 256
#define LYON_DEV_MINB 2
#endif
    return tile_launch_nt<LYON_DEV_NT, MODE, LYON_DEV_MINB>(c, a);
#else
    switch (c->tile_nt) {
        case 128: return tile_launch_nt<128, MODE, 4>(c, a);
        case 512: return tile_launch_nt<512, MODE, 1>(c, a);
        case 257: return tile_launch_nt<256, MODE, 3>(c, a);
        default: return tile_launch_nt<256, MODE, 2>(c, a);
    }
#endif
}

int tile_error(lyon_cuda_ctx* c) {
    if (c->h_tile_totals->error) {
        c->err = (c->h_tile_totals->error & TILE_ERR_EMIT_MISMATCH) ? "internal: emit pass diverged from the count pass"
                                                                     : "internal: tile pipeline found inconsistent slot counts";
        return LYON_ERR_INTERNAL;
    }
    return LYON_OK;
}

// Variable line width: count per path, scan, (later) emit.  See stroke_vw.cuh.
int vw_plan(lyon_cuda_ctx* c, const lyon_path_batch* b, const lyon_stroke_options* o, uint64_t* n_vertices, uint64_t* n_indices) {
    if ((uint32_t)o->variable_line_width >= b->num_attributes) return LYON_ERR_INVALID_ARGUMENT;
    c->n_paths = b->n_paths;
    c->n_vertices = c->n_triangles = 0;
    c->d_path_verb_begin = b->path_verb_begin;
    c->params = make_params(*o, b->num_attributes);
    c->views = Views{};
    c->tile_planned = false;
    if (b->n_verbs == 0 || b->n_paths == 0) {
        c->planned = true;
        return LYON_OK;
    }
    if (!b->verbs || !b->points || !b->path_verb_begin || !b->path_point_begin) return LYON_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(b->points) & 7u) return LYON_ERR_INVALID_ARGUMENT;  // points are read as 64-bit pairs
    const uint32_t np = b->n_paths;
    TRY(ensure(c, c->vw_nv, ((size_t)np + 1) * 4));
    TRY(ensure(c, c->vw_nt, ((size_t)np + 1) * 4));
    TRY(ensure(c, c->vw_v0, ((size_t)np + 1) * 8));
    TRY(ensure(c, c->vw_t0, ((size_t)np + 1) * 8));
    TRY(ensure(c, c->misc, 512));
    const unsigned tiles = grid_for(np, BLOCK * 4);
    TRY(ensure(c, c->desc, (size_t)tiles * 16));
    CU_TRY(c, cudaMemsetAsync(c->misc.p, 0, 512, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->desc.p, 0, (size_t)tiles * 16, c->stream));
    if (c->timing) CU_TRY(c, cudaEventRecord(c->ev0, c->stream));
    k_vw_count<<<grid_for(np, 128), 128, 0, c->stream>>>(b->verbs, reinterpret_cast<const V2*>(b->points), b->path_verb_begin,
                                                        b->path_point_begin, np, c->params, (uint32_t*)c->vw_nv.p,
                                                        (uint32_t*)c->vw_nt.p);
    if (c->timing) {
        CU_TRY(c, cudaEventRecord(c->ev1, c->stream));
        c->ev_valid = true;
    }
    k_scan_counts<<<tiles, BLOCK, 0, c->stream>>>((const uint32_t*)c->vw_nv.p, np, (unsigned long long*)c->vw_v0.p, d_ticket(c, 0),
                                                 (uint64_t*)c->desc.p);
    k_scan_counts<<<tiles, BLOCK, 0, c->stream>>>((const uint32_t*)c->vw_nt.p, np, (unsigned long long*)c->vw_t0.p, d_ticket(c, 1),
                                                 (uint64_t*)c->desc.p + tiles);
    c->launches += 3;
    CU_TRY(c, cudaGetLastError());
    unsigned long long tot[2] = {0ull, 0ull};
    CU_TRY(c, cudaMemcpyAsync(&tot[0], (unsigned long long*)c->vw_v0.p + np, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(&tot[1], (unsigned long long*)c->vw_t0.p + np, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->n_vertices = tot[0];
    c->n_triangles = tot[1];
    if (c->n_vertices > 0xFFFFFFFFull || c->n_triangles > 0xFFFFFFFFull) return LYON_ERR_TOO_MANY_VERTICES;
    *n_vertices = c->n_vertices;
    *n_indices = 3ull * c->n_triangles;
    c->vw_batch = *b;
    c->vw_planned = true;
    c->planned = true;
    return LYON_OK;
}
int vw_emit(lyon_cuda_ctx* c, GpuVertex* d_vertices, uint32_t* d_indices, uint64_t vertex_base, uint64_t* d_path_vertex_end,
            uint64_t* d_path_index_end) {
    const lyon_path_batch& b = c->vw_batch;
    if ((c->n_vertices && !d_vertices) || (c->n_triangles && !d_indices)) return LYON_ERR_INVALID_ARGUMENT;
    if (c->timing) CU_TRY(c, cudaEventRecord(c->ev0, c->stream));
    k_vw_emit<<<grid_for(b.n_paths, 128), 128, 0, c->stream>>>(
        b.verbs, reinterpret_cast<const V2*>(b.points), b.path_verb_begin, b.path_point_begin, b.n_paths, c->params,
        (const unsigned long long*)c->vw_v0.p, (const unsigned long long*)c->vw_t0.p, d_vertices, d_indices, (uint32_t)vertex_base,
        (unsigned long long*)d_path_vertex_end, (unsigned long long*)d_path_index_end, d_totals(c));
    if (c->timing) {
        CU_TRY(c, cudaEventRecord(c->ev1, c->stream));
        c->ev_valid = true;
    }
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
// Every entry point runs on the context's device whatever device is current in the calling thread, and leaves the
// caller's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// The two-phase entry points (stroke plan -> emit, flatten_paths count -> emit) keep their plan in workspace that the
// other pipelines reuse (tickets, descriptors, count / offset buffers).  Any call that may overwrite that workspace
// drops the plans it does not own, so a stale emit fails with LYON_ERR_INVALID_ARGUMENT instead of reading another
// pipeline's data.
enum { KEEP_NONE = 0, KEEP_STROKE = 1, KEEP_PATH_FLATTEN = 2 };
static void drop_plans(lyon_cuda_ctx* c, int keep) {
    if (!(keep & KEEP_STROKE)) c->planned = c->tile_planned = c->vw_planned = false;
    if (!(keep & KEEP_PATH_FLATTEN)) c->pf_planned = false;
}

extern "C" {

int lyon_cuda_create(int device, void* stream, lyon_cuda_ctx** out) {
    if (!out) return LYON_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return LYON_ERR_CUDA;
    DeviceGuard device_guard(device);
    int now = -1;
    if (cudaGetDevice(&now) != cudaSuccess || now != device) return LYON_ERR_CUDA;
    lyon_cuda_ctx* c = new lyon_cuda_ctx();
    c->device = device;
    if (const char* e = getenv("LYON_PIPE")) c->wt_mode = strcmp(e, "wtile") == 0 ? 1 : 0;  // experiments: warp tiles
    if (const char* e = getenv("LYON_FLAT")) c->flat_block_tiles = strcmp(e, "warp") != 0;  // experiments: warp-tile flatten
    c->stream = (cudaStream_t)stream;  // NULL is CUDA's default stream
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
            c->persistent_blocks = (unsigned)prop.multiProcessorCount * 8u;
            c->num_sms = prop.multiProcessorCount;
            c->max_smem_optin = prop.sharedMemPerBlockOptin;
        }
    }
    if (cudaMallocHost((void**)&c->h_totals, sizeof(Totals)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_tile_totals, sizeof(TileTotals)) != cudaSuccess) {
        if (c->h_totals) cudaFreeHost(c->h_totals);
        delete c;
        return LYON_ERR_CUDA;
    }
    *out = c;
    return LYON_OK;
}

void lyon_cuda_destroy(lyon_cuda_ctx* c) {
    if (!c) return;
    DeviceGuard device_guard(c->device);
    host_pipe_destroy(c->host_pipe);
    c->host_pipe = nullptr;
    comm_state_destroy(c->comm);
    c->comm = nullptr;
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->seg_pt, &c->seg_off, &c->seg_tsplit, &c->seg_id, &c->pos, &c->t,   &c->segfs,
                      &c->len,    &c->adv,     &c->voff,       &c->toff,   &c->jcode, &c->full_list, &c->sp_end_slot, &c->desc, &c->misc,
                      &c->pf_rec, &c->attr_kind, &c->vw_nv, &c->vw_nt, &c->vw_v0, &c->vw_t0, &c->verb_n, &c->path_slots, &c->path_events, &c->path_slot_begin, &c->path_event_begin, &c->tile_first, &c->tile_desc, &c->tile_base, &c->tile_misc};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->h_totals) cudaFreeHost(c->h_totals);
    if (c->h_tile_totals) cudaFreeHost(c->h_tile_totals);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    delete c;
}

int lyon_cuda_set_stream(lyon_cuda_ctx* c, void* stream) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (!c) return LYON_ERR_INVALID_ARGUMENT;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->stream = (cudaStream_t)stream;
    return LYON_OK;
}
int lyon_cuda_synchronize(lyon_cuda_ctx* c) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (!c) return LYON_ERR_INVALID_ARGUMENT;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return LYON_OK;
}
const char* lyon_cuda_last_error(const lyon_cuda_ctx* c) { return c ? c->err.c_str() : "null context"; }
uint64_t lyon_cuda_launch_count(const lyon_cuda_ctx* c) { return c ? c->launches : 0; }
uint64_t lyon_cuda_workspace_bytes(const lyon_cuda_ctx* c) { return c ? c->ws_bytes : 0; }

int lyon_cuda_flatten_cubics_count(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol, uint64_t* d_offsets,
                                   uint64_t* total) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    return flatten_count<4>(c, d_ctrl, n, tol, d_offsets, total);
}
int lyon_cuda_flatten_cubics_emit(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol, const uint64_t* d_offsets,
                                  uint64_t total, float* d_points, float* d_t) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    return flatten_emit<4>(c, d_ctrl, n, tol, d_offsets, total, d_points, d_t);
}
int lyon_cuda_flatten_quadratics_count(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol, uint64_t* d_offsets,
                                       uint64_t* total) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    return flatten_count<3>(c, d_ctrl, n, tol, d_offsets, total);
}
int lyon_cuda_flatten_quadratics_emit(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol,
                                      const uint64_t* d_offsets, uint64_t total, float* d_points, float* d_t) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    return flatten_emit<3>(c, d_ctrl, n, tol, d_offsets, total, d_points, d_t);
}

int lyon_cuda_flatten_cubics_fused(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol, uint64_t* d_offsets,
                                   float* d_points, uint64_t point_capacity, float* d_t, uint64_t* total_points) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    return flatten_fused<4>(c, d_ctrl, n, tol, d_offsets, d_points, point_capacity, d_t, total_points);
}
int lyon_cuda_flatten_quadratics_fused(lyon_cuda_ctx* c, const float* d_ctrl, uint64_t n, float tol, uint64_t* d_offsets,
                                       float* d_points, uint64_t point_capacity, float* d_t, uint64_t* total_points) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    return flatten_fused<3>(c, d_ctrl, n, tol, d_offsets, d_points, point_capacity, d_t, total_points);
}

// Count pass shared by lyon_cuda_flatten_paths_count and lyon_cuda_measure_paths_count (`measure`: the second number is the
// number of PathMeasurements edges instead of output points).
static int pf_count_impl(lyon_cuda_ctx* c, const lyon_path_batch* b, float tol, bool measure, uint64_t* n_out_verbs,
                         uint64_t* n_out_points) {
    *n_out_verbs = *n_out_points = 0;
    c->pf_planned = false;
    c->pf_measure = measure;
    if (!(tol > 0.0f)) return LYON_ERR_UNSUPPORTED_PARAMETER;  // ToleranceIsNaN / non-positive
    c->pf_batch = *b;
    if (b->n_paths == 0 || b->n_verbs == 0) {
        c->pf_planned = true;
        return LYON_OK;
    }
    if (!b->verbs || !b->points || !b->path_verb_begin || !b->path_point_begin) return LYON_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(b->points) & 7u) return LYON_ERR_INVALID_ARGUMENT;  // points are read as 64-bit pairs
    const uint32_t np = b->n_paths;
    TRY(ensure(c, c->pf_rec, ((size_t)b->n_verbs + 1) * sizeof(PfRec)));
    TRY(ensure(c, c->vw_nv, ((size_t)np + 1) * 4));
    TRY(ensure(c, c->vw_nt, ((size_t)np + 1) * 4));
    TRY(ensure(c, c->vw_v0, ((size_t)np + 1) * 8));
    TRY(ensure(c, c->vw_t0, ((size_t)np + 1) * 8));
    TRY(ensure(c, c->misc, 512));
    const unsigned tiles = grid_for(np, BLOCK * 4);
    TRY(ensure(c, c->desc, (size_t)tiles * 16));
    CU_TRY(c, cudaMemsetAsync(c->misc.p, 0, 512, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->desc.p, 0, (size_t)tiles * 16, c->stream));
    unsigned blocks = (np + 7u) / 8u;
    if (blocks > (unsigned)c->num_sms * 16u) blocks = (unsigned)c->num_sms * 16u;
    k_pf_count<<<blocks, 256, 0, c->stream>>>(b->verbs, reinterpret_cast<const V2*>(b->points), b->path_verb_begin,
                                             b->path_point_begin, np, (b->num_attributes + 1u) / 2u,
                                             measure ? 1u : 1u + (b->num_attributes + 1u) / 2u, tol, (PfRec*)c->pf_rec.p,
                                             (uint32_t*)c->vw_nv.p,
                                             (uint32_t*)c->vw_nt.p, d_totals(c));
    k_scan_counts<<<tiles, BLOCK, 0, c->stream>>>((const uint32_t*)c->vw_nv.p, np, (unsigned long long*)c->vw_v0.p, d_ticket(c, 0),
                                                 (uint64_t*)c->desc.p);
    k_scan_counts<<<tiles, BLOCK, 0, c->stream>>>((const uint32_t*)c->vw_nt.p, np, (unsigned long long*)c->vw_t0.p, d_ticket(c, 1),
                                                 (uint64_t*)c->desc.p + tiles);
    c->launches += 3;
    CU_TRY(c, cudaGetLastError());
    unsigned long long tot[2] = {0ull, 0ull};
    CU_TRY(c, cudaMemcpyAsync(&tot[0], (unsigned long long*)c->vw_v0.p + np, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(&tot[1], (unsigned long long*)c->vw_t0.p + np, 8, cudaMemcpyDeviceToHost, c->stream));
    TRY(fetch_totals(c));
    const unsigned err = c->h_totals->error;
    CU_TRY(c, cudaMemsetAsync(&d_totals(c)->error, 0, 4, c->stream));
    if (err & 4u) {
        c->err = "path_point_begin does not match the points the verbs consume";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    if ((err & 2u) || tot[0] > 0xFFFFFFFFull || tot[1] > 0xFFFFFFFFull) {
        c->err = "flattened batch does not fit 32-bit verb / point ranges";
        return LYON_ERR_TOO_MANY_VERTICES;
    }
    *n_out_verbs = tot[0];
    *n_out_points = tot[1];
    c->pf_planned = true;
    return LYON_OK;
}
int lyon_cuda_flatten_paths_count(lyon_cuda_ctx* c, const lyon_path_batch* b, float tol, uint64_t* n_out_verbs,
                                  uint64_t* n_out_points) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    if (!c || !b || !n_out_verbs || !n_out_points) return LYON_ERR_INVALID_ARGUMENT;
    return pf_count_impl(c, b, tol, false, n_out_verbs, n_out_points);
}
int lyon_cuda_measure_paths_count(lyon_cuda_ctx* c, const lyon_path_batch* b, float tol, uint64_t* n_edges) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    if (!c || !b || !n_edges) return LYON_ERR_INVALID_ARGUMENT;
    uint64_t n_verbs = 0;
    // `let tolerance = tolerance.max(1e-4)` (measure.rs:176; f32::max ignores a NaN operand)
    return pf_count_impl(c, b, tol > 1e-4f ? tol : 1e-4f, true, &n_verbs, n_edges);
}
int lyon_cuda_measure_paths_emit(lyon_cuda_ctx* c, float* d_distance, uint32_t* d_event_index, float* d_t,
                                 uint64_t* d_path_edge_begin, float* d_path_length) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_PATH_FLATTEN);
    if (!c || !d_path_edge_begin) return LYON_ERR_INVALID_ARGUMENT;
    if (!c->pf_planned || !c->pf_measure) {
        c->err = "lyon_cuda_measure_paths_emit without a successful lyon_cuda_measure_paths_count";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    const lyon_path_batch& b = c->pf_batch;
    if (b.n_paths == 0 || b.n_verbs == 0) {
        CU_TRY(c, cudaMemsetAsync(d_path_edge_begin, 0, 8ull * ((size_t)b.n_paths + 1), c->stream));
        if (d_path_length && b.n_paths) CU_TRY(c, cudaMemsetAsync(d_path_length, 0, 4ull * (size_t)b.n_paths, c->stream));
        return LYON_OK;
    }
    if (!d_distance || !d_event_index || !d_t || !d_path_length) return LYON_ERR_INVALID_ARGUMENT;
    unsigned blocks = (b.n_paths + 7u) / 8u;
    if (blocks > (unsigned)c->num_sms * 16u) blocks = (unsigned)c->num_sms * 16u;
    k_pm_emit<<<blocks, 256, 0, c->stream>>>(b.verbs, reinterpret_cast<const V2*>(b.points), b.path_verb_begin, b.n_paths,
                                            (b.num_attributes + 1u) / 2u, (const PfRec*)c->pf_rec.p,
                                            (const unsigned long long*)c->vw_t0.p, d_distance, d_event_index, d_t,
                                            reinterpret_cast<unsigned long long*>(d_path_edge_begin), d_path_length);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}
int lyon_cuda_flatten_paths_emit(lyon_cuda_ctx* c, uint8_t* d_out_verbs, float* d_out_points, uint32_t* d_out_path_verb_begin,
                                 uint32_t* d_out_path_point_begin) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_PATH_FLATTEN);
    if (!c || !d_out_path_verb_begin || !d_out_path_point_begin) return LYON_ERR_INVALID_ARGUMENT;
    if (!c->pf_planned || c->pf_measure) {
        c->err = "lyon_cuda_flatten_paths_emit without a successful lyon_cuda_flatten_paths_count";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    const lyon_path_batch& b = c->pf_batch;
    if (b.n_paths == 0 || b.n_verbs == 0) {
        CU_TRY(c, cudaMemsetAsync(d_out_path_verb_begin, 0, 4ull * ((size_t)b.n_paths + 1), c->stream));
        CU_TRY(c, cudaMemsetAsync(d_out_path_point_begin, 0, 4ull * ((size_t)b.n_paths + 1), c->stream));
        return LYON_OK;
    }
    if (!d_out_verbs || !d_out_points) return LYON_ERR_INVALID_ARGUMENT;
    unsigned blocks = (b.n_paths + 7u) / 8u;
    if (blocks > (unsigned)c->num_sms * 16u) blocks = (unsigned)c->num_sms * 16u;
    k_pf_emit<<<blocks, 256, 0, c->stream>>>(b.verbs, reinterpret_cast<const V2*>(b.points), b.path_verb_begin, b.path_point_begin,
                                            b.n_paths, b.num_attributes, (const PfRec*)c->pf_rec.p,
                                            (const unsigned long long*)c->vw_v0.p,
                                            (const unsigned long long*)c->vw_t0.p, d_out_verbs, reinterpret_cast<V2*>(d_out_points),
                                            d_out_path_verb_begin, d_out_path_point_begin);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}

int lyon_cuda_stroke_plan(lyon_cuda_ctx* c, const lyon_path_batch* b, const lyon_stroke_options* o, uint64_t* n_vertices,
                          uint64_t* n_indices) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    if (!c || !b || !o || !n_vertices || !n_indices) return LYON_ERR_INVALID_ARGUMENT;
    *n_vertices = *n_indices = 0;
    c->planned = false;
    if (o->tolerance != o->tolerance) return LYON_ERR_UNSUPPORTED_PARAMETER;  // ToleranceIsNaN
    if (o->start_cap > 2 || o->end_cap > 2 || o->line_join > 3) return LYON_ERR_INVALID_ARGUMENT;
    c->vw_planned = false;
    if (o->variable_line_width >= 0 && b->num_attributes > 0) return vw_plan(c, b, o, n_vertices, n_indices);
    if (b->n_verbs >= SEG_MASK) return LYON_ERR_INVALID_ARGUMENT;
    c->n_paths = b->n_paths;
    c->n_vertices = c->n_triangles = 0;
    c->d_path_verb_begin = b->path_verb_begin;
    c->params = make_params(*o, b->num_attributes);
    Views& w = c->views;
    w = Views{};
    w.n_verbs = b->n_verbs;
    if (b->n_verbs == 0 || b->n_paths == 0) {
        c->planned = true;
        return LYON_OK;
    }
    if (!b->verbs || !b->points || !b->path_verb_begin) return LYON_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(b->points) & 7u) return LYON_ERR_INVALID_ARGUMENT;  // points are read as 64-bit pairs
    c->tile_planned = false;
    {
        bool usable = false;
        TRY(tile_prepass(c, b, &usable));
        if (usable) {
            TRY((tile_launch<TILE_MODE_COUNT>(c, nullptr, nullptr, 0, 0, 0, nullptr, nullptr)));
            TRY(fetch_tile_totals(c));
            if (c->tile_is_wtile && (c->h_tile_totals->error & TILE_ERR_WT_OVERFLOW)) {
                // a join or cap with more than 65535 vertices: warp tiles keep 16-bit offsets, CTA tiles do not
                TRY(tile_prepass(c, b, &usable, false));
                if (usable) {
                    TRY((tile_launch<TILE_MODE_COUNT>(c, nullptr, nullptr, 0, 0, 0, nullptr, nullptr)));
                    TRY(fetch_tile_totals(c));
                }
            }
        }
        if (usable) {
            TRY(tile_error(c));
            c->n_vertices = c->h_tile_totals->vertices;
            c->n_triangles = c->h_tile_totals->triangles;
            if (c->n_vertices > 0xFFFFFFFFull || c->n_triangles > 0xFFFFFFFFull) return LYON_ERR_TOO_MANY_VERTICES;
            *n_vertices = c->n_vertices;
            *n_indices = 3ull * c->n_triangles;
            c->tile_planned = true;
            c->planned = true;
            return LYON_OK;
        }
    }
    const size_t nv1 = (size_t)b->n_verbs + 1;
    TRY(ensure(c, c->seg_pt, nv1 * 4));
    TRY(ensure(c, c->seg_off, nv1 * 4));
    TRY(ensure(c, c->seg_tsplit, nv1 * 4));
    TRY(ensure(c, c->seg_id, nv1 * 4));
    TRY(ensure(c, c->misc, 512));
    const unsigned vtiles = grid_for(b->n_verbs, BLOCK);
    TRY(ensure(c, c->desc, (size_t)vtiles * 24));
    w.verbs = b->verbs;
    w.points = reinterpret_cast<const V2*>(b->points);
    w.seg_pt = (uint32_t*)c->seg_pt.p;
    w.seg_off = (uint32_t*)c->seg_off.p;
    w.seg_tsplit = (float*)c->seg_tsplit.p;
    w.seg_id = (uint32_t*)c->seg_id.p;
    CU_TRY(c, cudaMemsetAsync(c->misc.p, 0, 512, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->desc.p, 0, (size_t)vtiles * 24, c->stream));
    k_seg_setup<<<vtiles, BLOCK, 0, c->stream>>>(w, c->params, d_ticket(c, 0), (uint64_t*)c->desc.p,
                                                (uint64_t*)c->desc.p + vtiles, (uint64_t*)c->desc.p + 2 * (size_t)vtiles,
                                                d_totals(c));
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    TRY(fetch_totals(c));
    const uint64_t n_slots = c->h_totals->slots;
    if (c->h_totals->points > b->n_points) {
        c->err = "verbs consume more points than the batch holds";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    if (n_slots >= 0xFFFFFFF0ull) return LYON_ERR_TOO_MANY_VERTICES;  // flatten-count explosion past u32 ids
    w.n_slots = (uint32_t)n_slots;
    const size_t ns1 = (size_t)n_slots + 1;
    TRY(ensure(c, c->pos, ns1 * 8));
    TRY(ensure(c, c->t, ns1 * 4));
    TRY(ensure(c, c->segfs, ns1 * 4));
    TRY(ensure(c, c->len, ns1 * 4));
    TRY(ensure(c, c->adv, ns1 * 4));
    TRY(ensure(c, c->voff, ns1 * 4));
    TRY(ensure(c, c->toff, ns1 * 4));
    TRY(ensure(c, c->jcode, ns1));
    TRY(ensure(c, c->full_list, ns1 * 4));  // worst case: every slot is a full join
    c->n_subpaths = (uint32_t)c->h_totals->subpaths;
    TRY(ensure(c, c->sp_end_slot, ((size_t)c->n_subpaths + 1) * 4));
    const unsigned stiles = grid_for(n_slots, BLOCK * SCAN_ITEMS);
    TRY(ensure(c, c->desc, (size_t)stiles * 16));
    w.pos = (V2*)c->pos.p;
    w.t = (float*)c->t.p;
    w.segfs = (uint32_t*)c->segfs.p;
    w.len = (float*)c->len.p;
    w.adv = (float*)c->adv.p;
    w.voff = (uint32_t*)c->voff.p;
    w.toff = (uint32_t*)c->toff.p;
    uint32_t* sp_end = (uint32_t*)c->sp_end_slot.p;
    uint32_t* full_list = (uint32_t*)c->full_list.p;
    uint32_t* full_count = d_ticket(c, 2);
    k_point_gen<<<vtiles, BLOCK, 0, c->stream>>>(w, c->params, sp_end);
    k_path_walk<<<grid_for(b->n_paths, 128), 128, 0, c->stream>>>(w, c->params, b->path_verb_begin, b->n_paths);
    k_classify_count<<<grid_for(n_slots, BLOCK), BLOCK, 0, c->stream>>>(w, c->params, (uint8_t*)c->jcode.p, full_list, full_count);
    k_full_count<<<c->persistent_blocks, 128, 0, c->stream>>>(w, c->params, (uint8_t*)c->jcode.p, full_list, full_count);
    if (c->n_subpaths)
        k_end_count<<<grid_for(c->n_subpaths, 128), 128, 0, c->stream>>>(w, c->params, sp_end, c->n_subpaths);
    CU_TRY(c, cudaMemsetAsync(c->desc.p, 0, (size_t)stiles * 16, c->stream));
    k_scan_offsets<<<stiles, BLOCK, 0, c->stream>>>(w, d_ticket(c, 1), (uint64_t*)c->desc.p, (uint64_t*)c->desc.p + stiles,
                                                   d_totals(c));
    c->launches += c->n_subpaths ? 6 : 5;
    CU_TRY(c, cudaGetLastError());
    TRY(fetch_totals(c));
    c->n_vertices = c->h_totals->vertices;
    c->n_triangles = c->h_totals->triangles;
    if (c->n_vertices > 0xFFFFFFFFull || c->n_triangles > 0xFFFFFFFFull) return LYON_ERR_TOO_MANY_VERTICES;
    *n_vertices = c->n_vertices;
    *n_indices = 3ull * c->n_triangles;
    c->planned = true;
    return LYON_OK;
}

int lyon_cuda_stroke_emit(lyon_cuda_ctx* c, lyon_stroke_vertex* d_vertices, uint32_t* d_indices, uint64_t vertex_base,
                          uint64_t* d_path_vertex_end, uint64_t* d_path_index_end) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_STROKE);
    if (!c) return LYON_ERR_INVALID_ARGUMENT;
    if (!c->planned) {
        c->err = "lyon_cuda_stroke_emit without a successful lyon_cuda_stroke_plan";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    // ids handed out are vertex_base .. vertex_base + V - 1 and must be <= u32::MAX - 1 (geometry_builder.rs:342-346)
    if (c->n_vertices && vertex_base + c->n_vertices - 1ull > 0xFFFFFFFEull) return LYON_ERR_TOO_MANY_VERTICES;
    if (reinterpret_cast<uintptr_t>(d_vertices) & 31u) {
        c->err = "d_vertices must be 32-byte aligned (vertices leave the SM as 256-bit stores)";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    if (c->vw_planned) return vw_emit(c, (GpuVertex*)d_vertices, d_indices, vertex_base, d_path_vertex_end, d_path_index_end);
    if (c->tile_planned) {
        if ((c->n_vertices && !d_vertices) || (c->n_triangles && !d_indices)) return LYON_ERR_INVALID_ARGUMENT;
        return tile_launch<TILE_MODE_EMIT>(c, (GpuVertex*)d_vertices, d_indices, c->n_vertices, c->n_triangles, vertex_base,
                                           d_path_vertex_end, d_path_index_end);
    }
    if (c->views.n_slots == 0 || c->views.n_verbs == 0) {
        if (c->n_paths) {
            if (d_path_vertex_end) CU_TRY(c, cudaMemsetAsync(d_path_vertex_end, 0, 8ull * c->n_paths, c->stream));
            if (d_path_index_end) CU_TRY(c, cudaMemsetAsync(d_path_index_end, 0, 8ull * c->n_paths, c->stream));
        }
        return LYON_OK;
    }
    if ((c->n_vertices && !d_vertices) || (c->n_triangles && !d_indices)) return LYON_ERR_INVALID_ARGUMENT;
    k_emit_fast<<<grid_for(c->views.n_slots, BLOCK), BLOCK, 0, c->stream>>>(
        c->views, c->params, (const uint8_t*)c->jcode.p, (GpuVertex*)d_vertices, d_indices, (uint32_t)vertex_base, d_totals(c));
    k_emit_full<<<c->persistent_blocks, 128, 0, c->stream>>>(c->views, c->params, (const uint8_t*)c->jcode.p,
                                                             (const uint32_t*)c->full_list.p, d_ticket(c, 2), (GpuVertex*)d_vertices,
                                                             d_indices, (uint32_t)vertex_base, d_totals(c));
    c->launches += 2;
    if (c->n_subpaths) {
        k_emit_end<<<grid_for(c->n_subpaths, 128), 128, 0, c->stream>>>(c->views, c->params, (const uint32_t*)c->sp_end_slot.p,
                                                                         c->n_subpaths, (GpuVertex*)d_vertices, d_indices,
                                                                         (uint32_t)vertex_base, d_totals(c));
        c->launches++;
    }
    if (d_path_vertex_end || d_path_index_end) {
        k_path_ranges<<<grid_for(c->n_paths, 128), 128, 0, c->stream>>>(c->views, c->d_path_verb_begin, c->n_paths,
                                                                       (unsigned long long*)d_path_vertex_end,
                                                                       (unsigned long long*)d_path_index_end, d_totals(c));
        c->launches++;
    }
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}

// Reads back the device-side consistency flag of the last emit (plan/emit divergence => LYON_ERR_INTERNAL).
int lyon_cuda_check(lyon_cuda_ctx* c) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (!c) return LYON_ERR_INVALID_ARGUMENT;
    if (c->tile_misc.p) {
        TRY(fetch_tile_totals(c));
        TRY(tile_error(c));
    }
    if (!c->misc.p) return LYON_OK;
    TRY(fetch_totals(c));
    if (c->h_totals->error) {
        c->err = "internal: emit pass diverged from the count pass";
        return LYON_ERR_INTERNAL;
    }
    return LYON_OK;
}

int lyon_cuda_enable_timing(lyon_cuda_ctx* c, int on) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (!c) return LYON_ERR_INVALID_ARGUMENT;
    if (on && !c->ev0) {
        CU_TRY(c, cudaEventCreate(&c->ev0));
        CU_TRY(c, cudaEventCreate(&c->ev1));
    }
    c->timing = on != 0;
    c->ev_valid = false;
    return LYON_OK;
}
int lyon_cuda_last_kernel_ms(lyon_cuda_ctx* c, float* ms) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (!c || !ms) return LYON_ERR_INVALID_ARGUMENT;
    *ms = 0.0f;
    if (!c->ev_valid) return LYON_ERR_INVALID_ARGUMENT;
    CU_TRY(c, cudaEventSynchronize(c->ev1));
    CU_TRY(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return LYON_OK;
}

int lyon_cuda_rebase_indices(lyon_cuda_ctx* c, uint32_t* d_indices, uint64_t n_indices, uint64_t n_vertices, uint64_t vertex_base) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (!c || (n_indices && !d_indices)) return LYON_ERR_INVALID_ARGUMENT;
    // ids handed out become vertex_base .. vertex_base + V - 1 and must stay <= u32::MAX - 1 (geometry_builder.rs:342-346)
    if (n_vertices && vertex_base + n_vertices - 1ull > 0xFFFFFFFEull) return LYON_ERR_TOO_MANY_VERTICES;
    if (n_indices == 0 || vertex_base == 0) return LYON_OK;
    unsigned blocks = grid_for((n_indices + 3) / 4, 256);
    const unsigned cap = (unsigned)c->num_sms * 32u;
    if (blocks > cap) blocks = cap;
    k_rebase_indices<<<blocks, 256, 0, c->stream>>>(d_indices, n_indices, (uint32_t)vertex_base);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}

int lyon_cuda_interpolate_attributes(lyon_cuda_ctx* c, const lyon_path_batch* b, const lyon_stroke_vertex* d_vertices,
                                     uint64_t n_vertices, const uint64_t* d_path_vertex_end, float* d_out) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    if (!c || !b) return LYON_ERR_INVALID_ARGUMENT;
    if (b->num_attributes == 0 || n_vertices == 0) return LYON_OK;  // nothing to interpolate
    if (!d_vertices || !d_path_vertex_end || !d_out || !b->verbs || !b->points || !b->path_verb_begin || !b->path_point_begin ||
        b->n_paths == 0)
        return LYON_ERR_INVALID_ARGUMENT;
    const uint32_t attr_stride = (b->num_attributes + 1u) / 2u;
    TRY(ensure(c, c->attr_kind, (size_t)b->n_points + 16));
    CU_TRY(c, cudaMemsetAsync(c->attr_kind.p, 0, (size_t)b->n_points + 16, c->stream));
    unsigned blocks = (b->n_paths + 7u) / 8u;
    if (blocks > (unsigned)c->num_sms * 16u) blocks = (unsigned)c->num_sms * 16u;
    k_attr_curve_kind<<<blocks, 256, 0, c->stream>>>(b->verbs, b->path_verb_begin, b->path_point_begin, b->n_paths, attr_stride,
                                                    (uint8_t*)c->attr_kind.p);
    unsigned vblocks = grid_for(n_vertices, 256);
    const unsigned cap = (unsigned)c->num_sms * 32u;
    if (vblocks > cap) vblocks = cap;
    k_interpolate_attributes<<<vblocks, 256, 0, c->stream>>>(reinterpret_cast<const GpuVertex*>(d_vertices), n_vertices,
                                                            (const unsigned long long*)d_path_vertex_end, b->n_paths, b->points,
                                                            b->path_point_begin, (const uint8_t*)c->attr_kind.p, b->num_attributes,
                                                            attr_stride, d_out);
    c->launches += 2;
    CU_TRY(c, cudaGetLastError());
    return LYON_OK;
}

int lyon_cuda_set_pipeline(lyon_cuda_ctx* c, int pipeline) {
    if (!c || pipeline < 0 || pipeline > 2) return LYON_ERR_INVALID_ARGUMENT;
    c->tile_force_general = pipeline == 1;
    c->wt_mode = pipeline == 2 ? 1 : 0;  // 2: warp-owned tiles instead of CTA-owned ones (experiment)
    if (pipeline == 0)
        if (const char* e = getenv("LYON_PIPE")) c->wt_mode = strcmp(e, "wtile") == 0 ? 1 : 0;
    return LYON_OK;
}

int lyon_cuda_stroke_fused(lyon_cuda_ctx* c, const lyon_path_batch* b, const lyon_stroke_options* o, lyon_stroke_vertex* d_vertices,
                           uint64_t vertex_capacity, uint32_t* d_indices, uint64_t index_capacity, uint64_t vertex_base,
                           uint64_t* d_path_vertex_end, uint64_t* d_path_index_end, uint64_t* n_vertices, uint64_t* n_indices) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (c) drop_plans(c, KEEP_NONE);
    if (!c || !b || !o || !n_vertices || !n_indices) return LYON_ERR_INVALID_ARGUMENT;
    *n_vertices = *n_indices = 0;
    c->planned = false;
    c->tile_planned = false;
    if (o->tolerance != o->tolerance) return LYON_ERR_UNSUPPORTED_PARAMETER;  // ToleranceIsNaN
    if (o->start_cap > 2 || o->end_cap > 2 || o->line_join > 3) return LYON_ERR_INVALID_ARGUMENT;
    if ((vertex_capacity && !d_vertices) || (index_capacity && !d_indices)) return LYON_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(d_vertices) & 31u) {
        c->err = "d_vertices must be 32-byte aligned (vertices leave the SM as 256-bit stores)";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    bool usable = false;
    const bool vw = o->variable_line_width >= 0 && b->num_attributes > 0;  // thread-per-path kernels: plan, then emit
    if (!vw && b->n_paths && b->n_verbs) {
        if (!b->verbs || !b->points || !b->path_verb_begin) return LYON_ERR_INVALID_ARGUMENT;
        if (reinterpret_cast<uintptr_t>(b->points) & 7u) return LYON_ERR_INVALID_ARGUMENT;  // points are read as 64-bit pairs
        c->params = make_params(*o, b->num_attributes);
        TRY(tile_prepass(c, b, &usable));
    }
    if (!usable) {
        // paths too long for a shared-memory tile (or no path_point_begin): two-pass general pipeline
        uint64_t nv = 0, ni = 0;
        TRY(lyon_cuda_stroke_plan(c, b, o, &nv, &ni));
        *n_vertices = nv;
        *n_indices = ni;
        if (nv > vertex_capacity || ni > index_capacity) return LYON_ERR_BUFFER_TOO_SMALL;
        return lyon_cuda_stroke_emit(c, d_vertices, d_indices, vertex_base, d_path_vertex_end, d_path_index_end);
    }
    TRY((tile_launch<TILE_MODE_FUSED>(c, (GpuVertex*)d_vertices, d_indices, vertex_capacity, index_capacity / 3ull, vertex_base,
                                      d_path_vertex_end, d_path_index_end)));
    TRY(fetch_tile_totals(c));
    if (c->tile_is_wtile && (c->h_tile_totals->error & TILE_ERR_WT_OVERFLOW)) {
        // a join or cap with more than 65535 vertices: warp tiles keep 16-bit offsets, CTA tiles do not
        TRY(tile_prepass(c, b, &usable, false));
        if (!usable) {
            uint64_t nv = 0, ni = 0;
            c->tile_force_general ^= 2;  // plan through the general pipeline
            const int st = lyon_cuda_stroke_plan(c, b, o, &nv, &ni);
            c->tile_force_general ^= 2;
            TRY(st);
            *n_vertices = nv;
            *n_indices = ni;
            if (nv > vertex_capacity || ni > index_capacity) return LYON_ERR_BUFFER_TOO_SMALL;
            return lyon_cuda_stroke_emit(c, d_vertices, d_indices, vertex_base, d_path_vertex_end, d_path_index_end);
        }
        TRY((tile_launch<TILE_MODE_FUSED>(c, (GpuVertex*)d_vertices, d_indices, vertex_capacity, index_capacity / 3ull, vertex_base,
                                          d_path_vertex_end, d_path_index_end)));
        TRY(fetch_tile_totals(c));
    }
    TRY(tile_error(c));
    const uint64_t nv = c->h_tile_totals->vertices, nt = c->h_tile_totals->triangles;
    if (nv > 0xFFFFFFFFull || nt > 0xFFFFFFFFull || (nv && vertex_base + nv - 1ull > 0xFFFFFFFEull)) return LYON_ERR_TOO_MANY_VERTICES;
    *n_vertices = nv;
    *n_indices = 3ull * nt;
    if (c->h_tile_totals->overflow) return LYON_ERR_BUFFER_TOO_SMALL;
    return LYON_OK;
}

int lyon_cuda_tessellate_paths(lyon_cuda_ctx* c, const lyon_path_batch* hb, const lyon_stroke_options* o,
                               lyon_stroke_vertex** h_vertices, uint64_t* n_vertices, uint32_t** h_indices,
                               uint64_t* n_indices) {
    DeviceGuard device_guard(c ? c->device : -1);
    if (!c || !hb || !o || !h_vertices || !n_vertices || !h_indices || !n_indices) return LYON_ERR_INVALID_ARGUMENT;
    *h_vertices = nullptr;
    *h_indices = nullptr;
    *n_vertices = *n_indices = 0;
    if (hb->n_paths == 0 || hb->n_verbs == 0) return LYON_OK;
    const size_t a_verbs = ((size_t)hb->n_verbs + 255) & ~(size_t)255;
    const size_t b_points = (size_t)hb->n_points * 8;
    const size_t b_ranges = ((size_t)hb->n_paths + 1) * 4;
    const size_t a_ranges = (b_ranges + 255) & ~(size_t)255;
    char* d_in = nullptr;
    CU_TRY(c, cudaMalloc((void**)&d_in, a_verbs + b_points + 256 + 2 * a_ranges));
    lyon_path_batch db = *hb;
    db.verbs = (const uint8_t*)d_in;
    db.points = (const float*)(d_in + a_verbs);
    db.path_verb_begin = (const uint32_t*)(d_in + a_verbs + ((b_points + 255) & ~(size_t)255));
    db.path_point_begin = hb->path_point_begin ? (const uint32_t*)((const char*)db.path_verb_begin + a_ranges) : nullptr;
    int st = LYON_OK;
    void *dv = nullptr, *di = nullptr;
    do {
        if (cudaMemcpyAsync((void*)db.verbs, hb->verbs, hb->n_verbs, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync((void*)db.points, hb->points, b_points, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync((void*)db.path_verb_begin, hb->path_verb_begin, b_ranges, cudaMemcpyHostToDevice, c->stream) !=
                cudaSuccess ||
            (db.path_point_begin && cudaMemcpyAsync((void*)db.path_point_begin, hb->path_point_begin, b_ranges,
                                                    cudaMemcpyHostToDevice, c->stream) != cudaSuccess)) {
            st = LYON_ERR_CUDA;
            break;
        }
        uint64_t nv = 0, ni = 0;
        st = lyon_cuda_stroke_plan(c, &db, o, &nv, &ni);
        if (st != LYON_OK) break;
        if (nv == 0) break;
        if (cudaMalloc(&dv, nv * 32) != cudaSuccess || cudaMalloc(&di, (ni ? ni : 1) * 4) != cudaSuccess) {
            st = LYON_ERR_OUT_OF_MEMORY;
            break;
        }
        st = lyon_cuda_stroke_emit(c, (lyon_stroke_vertex*)dv, (uint32_t*)di, 0, nullptr, nullptr);
        if (st != LYON_OK) break;
        *h_vertices = (lyon_stroke_vertex*)malloc(nv * 32);
        *h_indices = (uint32_t*)malloc((ni ? ni : 1) * 4);
        if (!*h_vertices || !*h_indices) {
            st = LYON_ERR_OUT_OF_MEMORY;
            break;
        }
        if (cudaMemcpyAsync(*h_vertices, dv, nv * 32, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(*h_indices, di, ni * 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) {
            st = LYON_ERR_CUDA;
            break;
        }
        st = lyon_cuda_check(c);
        if (st != LYON_OK) break;
        *n_vertices = nv;
        *n_indices = ni;
    } while (0);
    cudaStreamSynchronize(c->stream);
    if (dv) cudaFree(dv);
    if (di) cudaFree(di);
    cudaFree(d_in);
    c->planned = false;  // the device batch is gone
    if (st != LYON_OK) {
        free(*h_vertices);
        free(*h_indices);
        *h_vertices = nullptr;
        *h_indices = nullptr;
        *n_vertices = *n_indices = 0;
    }
    return st;
}
void lyon_cuda_free_host(void* p) { free(p); }

}  // extern "C"

#include "host_pipeline.cuh"
#include "sharded.cuh"

#ifdef LYON_PHASE_TIMING
// Debug builds only (tools/phase_profile.py): read and reset the per-phase cycle counters of k_stroke_tile.
extern "C" int lyon_cuda_debug_phase_cycles(lyon_cuda_ctx* c, uint64_t* out16) {
    if (!c || !out16) return LYON_ERR_INVALID_ARGUMENT;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaMemcpyFromSymbol(out16, g_phase_cycles, 16 * sizeof(uint64_t)));
    {
        unsigned long long st[4];
        CU_TRY(c, cudaMemcpyFromSymbol(st, g_lookback_stats, sizeof(st)));
        out16[14] = st[0];
        out16[15] = st[1];
        static const unsigned long long z[4] = {0, 0, 0, 0};
        CU_TRY(c, cudaMemcpyToSymbol(g_lookback_stats, z, sizeof(z)));
    }
    static const uint64_t zero[16] = {0};
    CU_TRY(c, cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)));
    return LYON_OK;
}
#endif
