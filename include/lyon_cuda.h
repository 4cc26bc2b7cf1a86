/* lyon_cuda.h -- C ABI of the B200-native batched path flattener and stroke tessellator.
 *
 * This is the drop-in boundary for lyon's data-parallel hot path.  Nothing like
 * it exists in the reference (lyon is single-process Rust); each entry point names the reference
 * interface it stands in for (file:line under nical/lyon @ 8071ec06).  Plain pointers and sizes only.
 *
 * Conventions
 *   - A context owns one CUDA device, one stream and a grow-only workspace; it is externally
 *     synchronised (like `&mut StrokeTessellator`, stroke.rs:122-127).  Distinct contexts are independent.
 *   - All functions return LYON_OK (0) or a negative lyon_status.  On error output counts are 0
 *     (the reference rolls its buffers back with abort_geometry, geometry_builder.rs:478-481).
 *   - Pointers named d_* are DEVICE pointers; h_* are HOST pointers.
 *   - There is no CPU fallback: without a CUDA device lyon_cuda_create fails.
 */
#ifndef LYON_CUDA_H
#define LYON_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lyon_cuda_ctx lyon_cuda_ctx;

typedef enum {
    LYON_OK = 0,
    LYON_ERR_UNSUPPORTED_PARAMETER = -1, /* TessellationError::UnsupportedParamater (error.rs) */
    LYON_ERR_TOO_MANY_VERTICES = -2,     /* GeometryBuilderError::TooManyVertices (geometry_builder.rs:342-346) */
    LYON_ERR_INTERNAL = -3,              /* TessellationError::Internal */
    LYON_ERR_CUDA = -4,                  /* CUDA runtime failure (no reference analogue) */
    LYON_ERR_INVALID_ARGUMENT = -5,
    LYON_ERR_OUT_OF_MEMORY = -6,
    LYON_ERR_BUFFER_TOO_SMALL = -7       /* lyon_cuda_stroke_fused: output capacity exceeded; required sizes returned */
} lyon_status;

/* Path verbs, the encoding of lyon_path's private `Verb` (crates/path/src/path.rs:29-36). */
enum { LYON_VERB_LINE_TO = 0, LYON_VERB_QUADRATIC_TO = 1, LYON_VERB_CUBIC_TO = 2, LYON_VERB_BEGIN = 3, LYON_VERB_CLOSE = 4, LYON_VERB_END = 5 };
/* LineCap / LineJoin in declaration order (crates/path/src/lib.rs:170-174, :193-198). */
enum { LYON_CAP_BUTT = 0, LYON_CAP_SQUARE = 1, LYON_CAP_ROUND = 2 };
enum { LYON_JOIN_MITER = 0, LYON_JOIN_MITER_CLIP = 1, LYON_JOIN_ROUND = 2, LYON_JOIN_BEVEL = 3 };

/* #[repr(C)] mirror of StrokeOptions (crates/tessellation/src/lib.rs:315-353). */
typedef struct {
    uint8_t start_cap, end_cap, line_join, _pad;
    float line_width;
    int32_t variable_line_width; /* Option<AttributeIndex>: -1 = None */
    float miter_limit;
    float tolerance;
} lyon_stroke_options;

/* StrokeVertexGpu: everything lyon's StrokeVertex exposes (stroke.rs:2656-2747), 32 bytes.  DEVICE vertex arrays
 * handed to lyon_cuda_stroke_emit / lyon_cuda_stroke_fused must be 32-byte aligned (every cudaMalloc'ed pointer is):
 * the kernels write a vertex with one 256-bit store; a misaligned array is rejected with LYON_ERR_INVALID_ARGUMENT.
 *   position()  = position_on_path + normal * half_width      (stroke.rs:2676-2678)
 *   side        = (src & 0x80000000) ? Side::Negative : Side::Positive
 *   source      = t == 1.0 ? VertexSource::Endpoint{ id: src & 0x7fffffff }
 *                          : VertexSource::Edge{ to: src & 0x7fffffff, t }   (lib.rs:282-291)          */
typedef struct {
    float position_on_path[2];
    float normal[2];
    float advancement;
    float half_width;
    uint32_t src;
    float t;
} lyon_stroke_vertex;

/* A batch of paths in lyon_path's own storage layout, concatenated (path.rs:75-79; PathBuffer,
 * path_buffer.rs:15-27): verbs[u8], points[f32 xy pairs incl. control points and, after every endpoint,
 * ceil(num_attributes/2) attribute slots], and N+1 begin offsets per path (in verbs / in points). */
typedef struct {
    const uint8_t* verbs;
    const float* points; /* 8-byte aligned */
    const uint32_t* path_verb_begin;
    const uint32_t* path_point_begin;
    uint32_t n_paths;
    uint32_t num_attributes;
    uint32_t n_verbs;  /* == path_verb_begin[n_paths]; path_verb_begin[0] must be 0 */
    uint32_t n_points; /* == path_point_begin[n_paths]; path_point_begin[0] must be 0 */
} lyon_path_batch;

/* ---- context ------------------------------------------------------------------------------------- */
/* `stream` is the cudaStream_t every kernel and copy of the context is issued on; NULL is CUDA's default
 * stream.  The context never creates or destroys streams. */
int lyon_cuda_create(int device, void* stream, lyon_cuda_ctx** out);
void lyon_cuda_destroy(lyon_cuda_ctx* ctx);
int lyon_cuda_set_stream(lyon_cuda_ctx* ctx, void* stream);
int lyon_cuda_synchronize(lyon_cuda_ctx* ctx);
const char* lyon_cuda_last_error(const lyon_cuda_ctx* ctx);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t lyon_cuda_launch_count(const lyon_cuda_ctx* ctx);
/* Bytes currently held by the context's workspace. */
uint64_t lyon_cuda_workspace_bytes(const lyon_cuda_ctx* ctx);

/* ---- curve flattening (lyon_geom) ------------------------------------------------------------------ */
/* Batched CubicBezierSegment::for_each_flattened / ::flattened(tolerance) (cubic_bezier.rs:563-614,
 * 1405-1454).  d_ctrl: n x 8 floats (from, ctrl1, ctrl2, to).  Pass 1 counts: d_offsets (n+1 u64, device)
 * receives the exclusive scan of per-curve point counts; *total_points (host) the grand total. */
int lyon_cuda_flatten_cubics_count(lyon_cuda_ctx* ctx, const float* d_ctrl, uint64_t n, float tolerance,
                                   uint64_t* d_offsets, uint64_t* total_points);
/* Pass 2 emits the points after `from` of every curve, curve after curve, into d_points (xy pairs);
 * d_t (may be NULL) receives the curve parameter of each point (for_each_flattened_with_t). */
int lyon_cuda_flatten_cubics_emit(lyon_cuda_ctx* ctx, const float* d_ctrl, uint64_t n, float tolerance,
                                  const uint64_t* d_offsets, uint64_t total_points, float* d_points, float* d_t);
/* Same for QuadraticBezierSegment (quadratic_bezier.rs:337-391, 979-1031). d_ctrl: n x 6 floats. */
int lyon_cuda_flatten_quadratics_count(lyon_cuda_ctx* ctx, const float* d_ctrl, uint64_t n, float tolerance,
                                       uint64_t* d_offsets, uint64_t* total_points);
int lyon_cuda_flatten_quadratics_emit(lyon_cuda_ctx* ctx, const float* d_ctrl, uint64_t n, float tolerance,
                                      const uint64_t* d_offsets, uint64_t total_points, float* d_points, float* d_t);

/* Single pass: count, scan and emit in one kernel into buffers the caller keeps across calls (the steady-state form of
 * the two calls above: `Vec`s that have reached their capacity).  d_offsets (n+1 u64) is written, d_points / d_t hold
 * point_capacity points.  *total_points receives the exact total; when it exceeds point_capacity nothing is emitted
 * and LYON_ERR_BUFFER_TOO_SMALL is returned (grow the buffers and call again). */
int lyon_cuda_flatten_cubics_fused(lyon_cuda_ctx* ctx, const float* d_ctrl, uint64_t n, float tolerance, uint64_t* d_offsets,
                                   float* d_points, uint64_t point_capacity, float* d_t, uint64_t* total_points);
int lyon_cuda_flatten_quadratics_fused(lyon_cuda_ctx* ctx, const float* d_ctrl, uint64_t n, float tolerance, uint64_t* d_offsets,
                                       float* d_points, uint64_t point_capacity, float* d_t, uint64_t* total_points);

/* Path-level flattening: the `Flattened` adapter, `path.iter().flattened(tolerance)` (crates/path/src/iterator.rs:125-226,
 * PathEvent::Quadratic / Cubic -> the Line events of `segment.flattened(tolerance)`), for every path of a batch.  The
 * output is a batch in the same encoding (verbs Begin / LineTo / Close / End only; Close stores the first point again)
 * and can be fed back to any entry point.  Custom attributes are interpolated along the curves like the `Flattened`
 * builder adapter does (crates/path/src/private.rs:56-122: prev[i] * (1 - t) + attr[i] * t, the curve's own attributes
 * where t == 1.0) and stored after every output endpoint.  Pass 1 returns the sizes; pass 2 writes
 * d_out_verbs[n_out_verbs], d_out_points[2 * n_out_points] and the two n_paths + 1 range arrays. */
int lyon_cuda_flatten_paths_count(lyon_cuda_ctx* ctx, const lyon_path_batch* d_batch, float tolerance, uint64_t* n_out_verbs,
                                  uint64_t* n_out_points);
int lyon_cuda_flatten_paths_emit(lyon_cuda_ctx* ctx, uint8_t* d_out_verbs, float* d_out_points, uint32_t* d_out_path_verb_begin,
                                 uint32_t* d_out_path_point_begin);

/* PathMeasurements::from_path (crates/algorithms/src/measure.rs:140-260) for every path of a batch -- a consumer of
 * flattening: one Edge {distance, index, t} per Begin (distance unchanged, t 1), per LineTo (distance += |from - to|), per
 * flattened segment of a curve (for_each_flattened_with_t at max(tolerance, 1e-4): distance += |segment|, t = t.end) and per
 * Close (distance += |last - first|); `index` is the event's position in its path, the distance runs over the whole path
 * in f32, event order.  Pass 1 returns the number of edges; pass 2 writes d_distance / d_event_index / d_t (n_edges
 * each), d_path_edge_begin (n_paths + 1, u64) and d_path_length (n_paths: PathMeasurements::length, measure.rs:262-268). */
int lyon_cuda_measure_paths_count(lyon_cuda_ctx* ctx, const lyon_path_batch* d_batch, float tolerance, uint64_t* n_edges);
int lyon_cuda_measure_paths_emit(lyon_cuda_ctx* ctx, float* d_distance, uint32_t* d_event_index, float* d_t,
                                 uint64_t* d_path_edge_begin, float* d_path_length);

/* ---- stroke tessellation (lyon_tessellation::StrokeTessellator) ------------------------------------ */
/* Equivalent of, for every path p of the batch in order,
 *     StrokeTessellator::new().tessellate_path(p, options, &mut BuffersBuilder::new(&mut out, ctor))
 * (stroke.rs:164-177) on ONE VertexBuffers<StrokeVertexGpu, u32>, i.e. indices are absolute
 * (geometry_builder.rs:342-349).
 *
 * Step 1, plan: uploads nothing; `batch` pointers are DEVICE pointers.  Runs the count passes and the
 * scans and returns the exact output sizes.  Fails with LYON_ERR_TOO_MANY_VERTICES when
 * vertex_base + n_vertices - 1 > u32::MAX - 1. */
int lyon_cuda_stroke_plan(lyon_cuda_ctx* ctx, const lyon_path_batch* d_batch, const lyon_stroke_options* options,
                          uint64_t* n_vertices, uint64_t* n_indices);
/* Step 2, emit: writes n_vertices vertices and n_indices u32 indices of the planned batch.
 * vertex_base is added to every index (BuffersBuilder::with_vertex_offset, geometry_builder.rs:325-329;
 * multi-GPU global rebasing).  d_path_vertex_end / d_path_index_end (may be NULL): per path, the vertex /
 * index count of the batch up to and including that path (u64, relative to this batch). */
int lyon_cuda_stroke_emit(lyon_cuda_ctx* ctx, lyon_stroke_vertex* d_vertices, uint32_t* d_indices,
                          uint64_t vertex_base, uint64_t* d_path_vertex_end, uint64_t* d_path_index_end);

/* Single pass: count, scan and emit fused in one kernel (decoupled look-back assigns every tile of paths its
 * output base while the tile is still in shared memory), for callers that keep output buffers around like
 * bench/tess keeps its VertexBuffers (bench/tess/src/main.rs:186-196).  Capacities are in vertices / in u32
 * indices.  *n_vertices / *n_indices always receive the exact sizes of the result; when a capacity is too small
 * nothing useful is written and LYON_ERR_BUFFER_TOO_SMALL is returned (grow the buffers and call again).
 * Needs d_batch->path_point_begin.  Batches holding a path too long for a shared-memory tile (> ~3000 flattened
 * points) run the two-pass plan/emit pipeline inside this call. */
int lyon_cuda_stroke_fused(lyon_cuda_ctx* ctx, const lyon_path_batch* d_batch, const lyon_stroke_options* options,
                           lyon_stroke_vertex* d_vertices, uint64_t vertex_capacity, uint32_t* d_indices,
                           uint64_t index_capacity, uint64_t vertex_base, uint64_t* d_path_vertex_end,
                           uint64_t* d_path_index_end, uint64_t* n_vertices, uint64_t* n_indices);

/* Adds vertex_base to every index of a mesh that was emitted with vertex_base 0: BuffersBuilder::with_vertex_offset
 * (geometry_builder.rs:325-329) applied after the fact.  Multi-GPU use: every rank runs the single-pass
 * lyon_cuda_stroke_fused on its slice, the ranks all-gather {V_r, I_r}, and rank r rebases by sum(V_q, q < r) -- one
 * streaming pass over the indices instead of a separate count pass before the emit.  n_vertices is the number of
 * vertices of the mesh (TooManyVertices when vertex_base + n_vertices - 1 would pass u32::MAX - 1). */
int lyon_cuda_rebase_indices(lyon_cuda_ctx* ctx, uint32_t* d_indices, uint64_t n_indices, uint64_t n_vertices,
                             uint64_t vertex_base);

/* StrokeVertex::interpolated_attributes (stroke.rs:2728-2747) for every vertex of a mesh tessellated from d_batch (paths
 * built with num_attributes > 0, ids = offsets in each path's point array, stroke.rs:769-869): what a custom
 * StrokeVertexConstructor reads (geometry_builder.rs:398-433).  d_out[v * num_attributes + i]: the stored attributes of
 * the endpoint for VertexSource::Endpoint{id} (Path as AttributeStore, path.rs:650-665), a[i] * (1 - t) + b[i] * t between
 * the curve's two endpoints for VertexSource::Edge{from, to, t}.  d_path_vertex_end: the per-path vertex ranges that
 * stroke_emit / stroke_fused returned (relative to d_vertices[0]).  A batch without attributes is a no-op. */
int lyon_cuda_interpolate_attributes(lyon_cuda_ctx* ctx, const lyon_path_batch* d_batch, const lyon_stroke_vertex* d_vertices,
                                     uint64_t n_vertices, const uint64_t* d_path_vertex_end, float* d_out);

/* Pipeline selection for plan/emit/fused: 0 = automatic (shared-memory tile pipeline whenever every path of the
 * batch fits a tile), 1 = always the general multi-kernel pipeline (any path length).  Results are identical. */
int lyon_cuda_set_pipeline(lyon_cuda_ctx* ctx, int pipeline);

/* Device timing of the dominant kernel (the tile kernel of plan / emit / fused): when enabled the context brackets
 * that launch with CUDA events on its stream; lyon_cuda_last_kernel_ms returns the duration of the last one. */
int lyon_cuda_enable_timing(lyon_cuda_ctx* ctx, int on);
int lyon_cuda_last_kernel_ms(lyon_cuda_ctx* ctx, float* ms);

/* Synchronises and reads back the device-side consistency flag of the last emit: LYON_ERR_INTERNAL if the
 * emit pass ever produced a different number of vertices/triangles for a slot than the count pass planned. */
int lyon_cuda_check(lyon_cuda_ctx* ctx);

/* Host-buffer convenience (what a lyon user calls): upload the batch, plan, allocate, emit, download.
 * *h_vertices / *h_indices are malloc'ed by the library (free with lyon_cuda_free_host). */
int lyon_cuda_tessellate_paths(lyon_cuda_ctx* ctx, const lyon_path_batch* h_batch,
                               const lyon_stroke_options* options, lyon_stroke_vertex** h_vertices,
                               uint64_t* n_vertices, uint32_t** h_indices, uint64_t* n_indices);
void lyon_cuda_free_host(void* p);

/* The same call for production use: caller-provided host buffers (reused across calls like bench/tess reuses its
 * VertexBuffers, bench/tess/src/main.rs:186-196; pinned memory from lyon_cuda_host_alloc makes the copies asynchronous),
 * the batch cut into chunks of `chunk_paths` whole paths (0 = 32768) that are uploaded, tessellated and downloaded on two
 * alternating streams, so the download of one chunk overlaps the upload and the kernels of the next.  Chunk c is emitted
 * with vertex_base = the vertices of the chunks before it: the concatenated buffers equal the result of one
 * StrokeTessellator::tessellate_path loop over one VertexBuffers (stroke.rs:164-177, geometry_builder.rs:325-349).
 * Capacities are in vertices / in u32 indices.  *n_vertices / *n_indices always receive the exact sizes; when a capacity
 * is too small the whole batch is still counted and LYON_ERR_BUFFER_TOO_SMALL is returned (grow and call again).
 * Needs h_batch->path_point_begin. */
int lyon_cuda_tessellate_paths_into(lyon_cuda_ctx* ctx, const lyon_path_batch* h_batch, const lyon_stroke_options* options,
                                    lyon_stroke_vertex* h_vertices, uint64_t vertex_capacity, uint32_t* h_indices,
                                    uint64_t index_capacity, uint32_t chunk_paths, uint64_t* n_vertices, uint64_t* n_indices);
/* Page-locked host memory for the buffers above (cudaHostAlloc / cudaFreeHost). */
int lyon_cuda_host_alloc(void** p, size_t bytes);
void lyon_cuda_host_free(void* p);

/* ---- multi-GPU: one process (or thread) per GPU, one context each ------------------------------------------------
 * Paths are independent (a fresh StrokeBuilderImpl per tessellate call, stroke.rs:133-134), so the batch is cut into
 * nranks contiguous path ranges and every rank strokes its own.  The only exchange is an all-gather of the per-rank
 * totals {V_r, I_r} over NCCL; rank r's indices are rebased by sum(V_q, q < r) (BuffersBuilder::with_vertex_offset,
 * geometry_builder.rs:325-329), so the concatenation of the per-rank buffers is the single-GPU result bit for bit.
 * NCCL is loaded at run time (dlopen "libnccl.so.2", or the path in $LYON_NCCL_LIB). */
#define LYON_CUDA_COMM_ID_BYTES 128
/* Rank 0 creates the 128-byte id (ncclGetUniqueId) and the application hands it to every rank. */
int lyon_cuda_comm_unique_id(uint8_t* id128);
/* Collective: every rank calls it with the same id (ncclCommInitRank on the context's device). */
int lyon_cuda_comm_init(lyon_cuda_ctx* ctx, const uint8_t* id128, int rank, int nranks);
int lyon_cuda_comm_destroy(lyon_cuda_ctx* ctx);
/* Collective: lyon_cuda_stroke_fused on this rank's slice `d_batch` with slice-local ids, the all-gather on the context's
 * stream straight behind the kernels, then the index rebase with the base read from the gathered totals on the device
 * (no host round trip in between).  *vertex_base receives sum(V_q, q < rank); h_totals (may be NULL) receives
 * {V_q, I_q} for q = 0 .. nranks-1 (2 * nranks u64).  A rank whose slice fails still takes part in the exchange: it
 * returns its own error, the other ranks return LYON_ERR_INTERNAL naming it (LYON_ERR_BUFFER_TOO_SMALL on every rank when
 * a rank's buffers were too small: all ranks call again).  LYON_ERR_TOO_MANY_VERTICES when the job
 * as a whole emits more than u32::MAX vertices (geometry_builder.rs:578-581). */
int lyon_cuda_stroke_fused_sharded(lyon_cuda_ctx* ctx, const lyon_path_batch* d_batch, const lyon_stroke_options* options,
                                   lyon_stroke_vertex* d_vertices, uint64_t vertex_capacity, uint32_t* d_indices,
                                   uint64_t index_capacity, uint64_t* d_path_vertex_end, uint64_t* d_path_index_end,
                                   uint64_t* n_vertices, uint64_t* n_indices, uint64_t* vertex_base, uint64_t* h_totals);

#ifdef __cplusplus
}
#endif
#endif /* LYON_CUDA_H */
