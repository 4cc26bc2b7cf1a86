//! `lyon_cuda::tessellate_paths(&[Path], &StrokeOptions) -> VertexBuffers<StrokeVertexGpu, u32>`
//!
//! The result equals
//! ```ignore
//! let mut out = VertexBuffers::new();
//! for p in paths {
//!     StrokeTessellator::new().tessellate_path(p, options, &mut BuffersBuilder::new(&mut out, ctor))?;
//! }
//! ```
//! (crates/tessellation/src/stroke.rs:164-177, geometry_builder.rs:300-351): vertex ids are absolute.
//!
//! `Path`'s point/verb arrays are private (crates/path/src/path.rs:76-78), so packing walks
//! `Path::iter()` / `Path::id_iter()`; the packed layout is lyon's own storage layout concatenated
//! (see include/lyon_cuda.h: lyon_path_batch).
pub mod ffi;

use lyon_path::{Event, LineCap, LineJoin, Path};
use lyon_tessellation::{GeometryBuilderError, StrokeOptions, TessellationError, VertexBuffers};
use std::ptr;

pub use ffi::lyon_stroke_vertex as StrokeVertexGpu;

const VERB_LINE: u8 = 0;
const VERB_QUADRATIC: u8 = 1;
const VERB_CUBIC: u8 = 2;
const VERB_BEGIN: u8 = 3;
const VERB_CLOSE: u8 = 4;
const VERB_END: u8 = 5;

/// A batch of paths in the SoA layout the device consumes.
#[derive(Default)]
pub struct PathBatch {
    pub verbs: Vec<u8>,
    pub points: Vec<[f32; 2]>,
    pub path_verb_begin: Vec<u32>,
    pub path_point_begin: Vec<u32>,
    pub num_attributes: u32,
}

impl PathBatch {
    pub fn from_paths(paths: &[Path]) -> Self {
        let num_attributes = paths.first().map(|p| p.num_attributes() as u32).unwrap_or(0);
        let mut b = PathBatch { num_attributes, ..Default::default() };
        b.path_verb_begin.push(0);
        b.path_point_begin.push(0);
        let push_attrs = |pts: &mut Vec<[f32; 2]>, attrs: &[f32]| {
            for pair in attrs.chunks(2) {
                pts.push([pair[0], *pair.get(1).unwrap_or(&0.0)]);
            }
        };
        for path in paths {
            assert_eq!(path.num_attributes() as u32, num_attributes, "all paths of a batch share num_attributes");
            for evt in path.id_iter() {
                match evt {
                    Event::Begin { at } => {
                        b.verbs.push(VERB_BEGIN);
                        b.points.push(path[at].to_array());
                        push_attrs(&mut b.points, path.attributes(at));
                    }
                    Event::Line { to, .. } => {
                        b.verbs.push(VERB_LINE);
                        b.points.push(path[to].to_array());
                        push_attrs(&mut b.points, path.attributes(to));
                    }
                    Event::Quadratic { ctrl, to, .. } => {
                        b.verbs.push(VERB_QUADRATIC);
                        b.points.push(path[ctrl].to_array());
                        b.points.push(path[to].to_array());
                        push_attrs(&mut b.points, path.attributes(to));
                    }
                    Event::Cubic { ctrl1, ctrl2, to, .. } => {
                        b.verbs.push(VERB_CUBIC);
                        b.points.push(path[ctrl1].to_array());
                        b.points.push(path[ctrl2].to_array());
                        b.points.push(path[to].to_array());
                        push_attrs(&mut b.points, path.attributes(to));
                    }
                    Event::End { first, close, .. } => {
                        if close {
                            // Close stores the first endpoint (and its attributes) again, path.rs:537-539
                            b.verbs.push(VERB_CLOSE);
                            b.points.push(path[first].to_array());
                            push_attrs(&mut b.points, path.attributes(first));
                        } else {
                            b.verbs.push(VERB_END);
                        }
                    }
                }
            }
            b.path_verb_begin.push(b.verbs.len() as u32);
            b.path_point_begin.push(b.points.len() as u32);
        }
        b
    }
}

fn cap_code(c: LineCap) -> u8 {
    match c {
        LineCap::Butt => 0,
        LineCap::Square => 1,
        LineCap::Round => 2,
    }
}
fn join_code(j: LineJoin) -> u8 {
    match j {
        LineJoin::Miter => 0,
        LineJoin::MiterClip => 1,
        LineJoin::Round => 2,
        LineJoin::Bevel => 3,
    }
}

pub fn options_to_c(o: &StrokeOptions) -> ffi::lyon_stroke_options {
    ffi::lyon_stroke_options {
        start_cap: cap_code(o.start_cap),
        end_cap: cap_code(o.end_cap),
        line_join: join_code(o.line_join),
        _pad: 0,
        line_width: o.line_width,
        variable_line_width: o.variable_line_width.map(|i| i as i32).unwrap_or(-1),
        miter_limit: o.miter_limit,
        tolerance: o.tolerance,
    }
}

fn status_to_error(code: i32) -> TessellationError {
    match code {
        ffi::LYON_ERR_TOO_MANY_VERTICES => TessellationError::GeometryBuilder(GeometryBuilderError::TooManyVertices),
        ffi::LYON_ERR_UNSUPPORTED_PARAMETER => {
            TessellationError::UnsupportedParamater(lyon_tessellation::UnsupportedParamater::ToleranceIsNaN)
        }
        _ => TessellationError::Internal(lyon_tessellation::InternalError::ErrorCode(code as i16)),
    }
}

/// One CUDA context (device + stream + grow-only workspace); `&mut self` like `StrokeTessellator`.
pub struct CudaTessellator {
    ctx: *mut ffi::lyon_cuda_ctx,
}

impl CudaTessellator {
    pub fn new(device: i32) -> Result<Self, TessellationError> {
        let mut ctx = ptr::null_mut();
        let st = unsafe { ffi::lyon_cuda_create(device, ptr::null_mut(), &mut ctx) };
        if st != ffi::LYON_OK {
            return Err(status_to_error(st));
        }
        Ok(CudaTessellator { ctx })
    }

    pub fn tessellate_batch(&mut self, batch: &PathBatch, options: &StrokeOptions)
        -> Result<VertexBuffers<StrokeVertexGpu, u32>, TessellationError> {
        let hb = ffi::lyon_path_batch {
            verbs: batch.verbs.as_ptr(),
            points: batch.points.as_ptr() as *const f32,
            path_verb_begin: batch.path_verb_begin.as_ptr(),
            path_point_begin: batch.path_point_begin.as_ptr(),
            n_paths: (batch.path_verb_begin.len() - 1) as u32,
            num_attributes: batch.num_attributes,
            n_verbs: batch.verbs.len() as u32,
            n_points: batch.points.len() as u32,
        };
        let co = options_to_c(options);
        // `lyon_cuda_tessellate_paths_into`: chunked, double-buffered upload / kernels / download into buffers the caller
        // owns.  The first call with empty buffers returns the exact sizes (LYON_ERR_BUFFER_TOO_SMALL), the second fills
        // them; a caller that keeps its VertexBuffers across frames (bench/tess/src/main.rs:186-196) pays for one call.
        let mut out: VertexBuffers<StrokeVertexGpu, u32> = VertexBuffers::new();
        let (mut nv, mut ni) = (0u64, 0u64);
        for _ in 0..2 {
            let st = unsafe {
                ffi::lyon_cuda_tessellate_paths_into(self.ctx, &hb, &co, out.vertices.as_mut_ptr(), out.vertices.capacity() as u64,
                                                     out.indices.as_mut_ptr(), out.indices.capacity() as u64, 0, &mut nv, &mut ni)
            };
            if st == ffi::LYON_ERR_BUFFER_TOO_SMALL {
                out.vertices.reserve_exact(nv as usize);
                out.indices.reserve_exact(ni as usize);
                continue;
            }
            if st != ffi::LYON_OK {
                return Err(status_to_error(st));
            }
            // the library wrote nv vertices / ni indices into the reserved capacity (nothing when they are 0)
            unsafe {
                out.vertices.set_len(nv as usize);
                out.indices.set_len(ni as usize);
            }
            return Ok(out);
        }
        Err(status_to_error(ffi::LYON_ERR_BUFFER_TOO_SMALL))
    }
}

impl Drop for CudaTessellator {
    fn drop(&mut self) {
        unsafe { ffi::lyon_cuda_destroy(self.ctx) }
    }
}

/// The batch entry point named by the north star.
pub fn tessellate_paths(paths: &[Path], options: &StrokeOptions)
    -> Result<VertexBuffers<StrokeVertexGpu, u32>, TessellationError> {
    let mut t = CudaTessellator::new(0)?;
    t.tessellate_batch(&PathBatch::from_paths(paths), options)
}
