//! Raw bindings of include/lyon_cuda.h (kept in declaration order).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct lyon_cuda_ctx {
    _private: [u8; 0],
}

pub const LYON_OK: c_int = 0;
pub const LYON_ERR_UNSUPPORTED_PARAMETER: c_int = -1;
pub const LYON_ERR_TOO_MANY_VERTICES: c_int = -2;
pub const LYON_ERR_INTERNAL: c_int = -3;
pub const LYON_ERR_CUDA: c_int = -4;
pub const LYON_ERR_INVALID_ARGUMENT: c_int = -5;
pub const LYON_ERR_OUT_OF_MEMORY: c_int = -6;
pub const LYON_ERR_BUFFER_TOO_SMALL: c_int = -7;

/// `#[repr(C)]` mirror of `StrokeOptions` (crates/tessellation/src/lib.rs:315-353).
#[repr(C)]
#[derive(Copy, Clone, Debug)]
pub struct lyon_stroke_options {
    pub start_cap: u8,
    pub end_cap: u8,
    pub line_join: u8,
    pub _pad: u8,
    pub line_width: f32,
    pub variable_line_width: i32,
    pub miter_limit: f32,
    pub tolerance: f32,
}

/// `StrokeVertexGpu`: everything `StrokeVertex` exposes (crates/tessellation/src/stroke.rs:2656-2747).
/// 32 bytes, natural alignment 4 like the C typedef; DEVICE vertex arrays handed to the library must be 32-byte aligned
/// (the kernels write a vertex with one 256-bit store; the entry points check and return LYON_ERR_INVALID_ARGUMENT).
#[repr(C)]
#[derive(Copy, Clone, Debug, Default, PartialEq)]
pub struct lyon_stroke_vertex {
    pub position_on_path: [f32; 2],
    pub normal: [f32; 2],
    pub advancement: f32,
    pub half_width: f32,
    pub src: u32,
    pub t: f32,
}

#[repr(C)]
pub struct lyon_path_batch {
    pub verbs: *const u8,
    pub points: *const f32,
    pub path_verb_begin: *const u32,
    pub path_point_begin: *const u32,
    pub n_paths: u32,
    pub num_attributes: u32,
    pub n_verbs: u32,
    pub n_points: u32,
}

extern "C" {
    pub fn lyon_cuda_create(device: c_int, stream: *mut c_void, out: *mut *mut lyon_cuda_ctx) -> c_int;
    pub fn lyon_cuda_destroy(ctx: *mut lyon_cuda_ctx);
    pub fn lyon_cuda_set_stream(ctx: *mut lyon_cuda_ctx, stream: *mut c_void) -> c_int;
    pub fn lyon_cuda_synchronize(ctx: *mut lyon_cuda_ctx) -> c_int;
    pub fn lyon_cuda_last_error(ctx: *const lyon_cuda_ctx) -> *const c_char;
    pub fn lyon_cuda_launch_count(ctx: *const lyon_cuda_ctx) -> u64;
    pub fn lyon_cuda_workspace_bytes(ctx: *const lyon_cuda_ctx) -> u64;
    pub fn lyon_cuda_flatten_cubics_count(ctx: *mut lyon_cuda_ctx, d_ctrl: *const f32, n: u64, tolerance: f32,
                                          d_offsets: *mut u64, total_points: *mut u64) -> c_int;
    pub fn lyon_cuda_flatten_cubics_emit(ctx: *mut lyon_cuda_ctx, d_ctrl: *const f32, n: u64, tolerance: f32,
                                         d_offsets: *const u64, total_points: u64, d_points: *mut f32,
                                         d_t: *mut f32) -> c_int;
    pub fn lyon_cuda_flatten_quadratics_count(ctx: *mut lyon_cuda_ctx, d_ctrl: *const f32, n: u64, tolerance: f32,
                                              d_offsets: *mut u64, total_points: *mut u64) -> c_int;
    pub fn lyon_cuda_flatten_quadratics_emit(ctx: *mut lyon_cuda_ctx, d_ctrl: *const f32, n: u64, tolerance: f32,
                                             d_offsets: *const u64, total_points: u64, d_points: *mut f32,
                                             d_t: *mut f32) -> c_int;
    pub fn lyon_cuda_stroke_plan(ctx: *mut lyon_cuda_ctx, d_batch: *const lyon_path_batch,
                                 options: *const lyon_stroke_options, n_vertices: *mut u64,
                                 n_indices: *mut u64) -> c_int;
    pub fn lyon_cuda_stroke_emit(ctx: *mut lyon_cuda_ctx, d_vertices: *mut lyon_stroke_vertex, d_indices: *mut u32,
                                 vertex_base: u64, d_path_vertex_end: *mut u64, d_path_index_end: *mut u64) -> c_int;
    pub fn lyon_cuda_stroke_fused(ctx: *mut lyon_cuda_ctx, d_batch: *const lyon_path_batch,
                                  options: *const lyon_stroke_options, d_vertices: *mut lyon_stroke_vertex,
                                  vertex_capacity: u64, d_indices: *mut u32, index_capacity: u64, vertex_base: u64,
                                  d_path_vertex_end: *mut u64, d_path_index_end: *mut u64, n_vertices: *mut u64,
                                  n_indices: *mut u64) -> c_int;
    pub fn lyon_cuda_flatten_cubics_fused(ctx: *mut lyon_cuda_ctx, d_ctrl: *const f32, n: u64, tolerance: f32,
                                          d_offsets: *mut u64, d_points: *mut f32, point_capacity: u64, d_t: *mut f32,
                                          total_points: *mut u64) -> c_int;
    pub fn lyon_cuda_flatten_quadratics_fused(ctx: *mut lyon_cuda_ctx, d_ctrl: *const f32, n: u64, tolerance: f32,
                                              d_offsets: *mut u64, d_points: *mut f32, point_capacity: u64, d_t: *mut f32,
                                              total_points: *mut u64) -> c_int;
    pub fn lyon_cuda_flatten_paths_count(ctx: *mut lyon_cuda_ctx, d_batch: *const lyon_path_batch, tolerance: f32,
                                         n_out_verbs: *mut u64, n_out_points: *mut u64) -> c_int;
    pub fn lyon_cuda_flatten_paths_emit(ctx: *mut lyon_cuda_ctx, d_out_verbs: *mut u8, d_out_points: *mut f32,
                                        d_out_path_verb_begin: *mut u32, d_out_path_point_begin: *mut u32) -> c_int;
    pub fn lyon_cuda_measure_paths_count(ctx: *mut lyon_cuda_ctx, d_batch: *const lyon_path_batch, tolerance: f32,
                                         n_edges: *mut u64) -> c_int;
    pub fn lyon_cuda_measure_paths_emit(ctx: *mut lyon_cuda_ctx, d_distance: *mut f32, d_event_index: *mut u32, d_t: *mut f32,
                                        d_path_edge_begin: *mut u64, d_path_length: *mut f32) -> c_int;
    pub fn lyon_cuda_rebase_indices(ctx: *mut lyon_cuda_ctx, d_indices: *mut u32, n_indices: u64, n_vertices: u64,
                                    vertex_base: u64) -> c_int;
    pub fn lyon_cuda_interpolate_attributes(ctx: *mut lyon_cuda_ctx, d_batch: *const lyon_path_batch,
                                            d_vertices: *const lyon_stroke_vertex, n_vertices: u64,
                                            d_path_vertex_end: *const u64, d_out: *mut f32) -> c_int;
    pub fn lyon_cuda_set_pipeline(ctx: *mut lyon_cuda_ctx, pipeline: c_int) -> c_int;
    pub fn lyon_cuda_enable_timing(ctx: *mut lyon_cuda_ctx, on: c_int) -> c_int;
    pub fn lyon_cuda_last_kernel_ms(ctx: *mut lyon_cuda_ctx, ms: *mut f32) -> c_int;
    pub fn lyon_cuda_check(ctx: *mut lyon_cuda_ctx) -> c_int;
    pub fn lyon_cuda_tessellate_paths(ctx: *mut lyon_cuda_ctx, h_batch: *const lyon_path_batch,
                                      options: *const lyon_stroke_options, h_vertices: *mut *mut lyon_stroke_vertex,
                                      n_vertices: *mut u64, h_indices: *mut *mut u32, n_indices: *mut u64) -> c_int;
    pub fn lyon_cuda_free_host(p: *mut c_void);
    pub fn lyon_cuda_tessellate_paths_into(ctx: *mut lyon_cuda_ctx, h_batch: *const lyon_path_batch,
                                           options: *const lyon_stroke_options, h_vertices: *mut lyon_stroke_vertex,
                                           vertex_capacity: u64, h_indices: *mut u32, index_capacity: u64,
                                           chunk_paths: u32, n_vertices: *mut u64, n_indices: *mut u64) -> c_int;
    pub fn lyon_cuda_host_alloc(p: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn lyon_cuda_host_free(p: *mut c_void);
    pub fn lyon_cuda_comm_unique_id(id128: *mut u8) -> c_int;
    pub fn lyon_cuda_comm_init(ctx: *mut lyon_cuda_ctx, id128: *const u8, rank: c_int, nranks: c_int) -> c_int;
    pub fn lyon_cuda_comm_destroy(ctx: *mut lyon_cuda_ctx) -> c_int;
    pub fn lyon_cuda_stroke_fused_sharded(ctx: *mut lyon_cuda_ctx, d_batch: *const lyon_path_batch,
                                          options: *const lyon_stroke_options, d_vertices: *mut lyon_stroke_vertex,
                                          vertex_capacity: u64, d_indices: *mut u32, index_capacity: u64,
                                          d_path_vertex_end: *mut u64, d_path_index_end: *mut u64, n_vertices: *mut u64,
                                          n_indices: *mut u64, vertex_base: *mut u64, h_totals: *mut u64) -> c_int;
}
