"""lyon_b200 -- B200-native batched path flattening + stroke tessellation (lyon's hot path).

Host-side mirror of lyon's Path / StrokeOptions / VertexBuffers API over the C ABI in
include/lyon_cuda.h (CUDA kernels in lyon_b200/csrc).  There is no CPU fallback: every compute entry
point raises if the CUDA library or a device is missing.
"""
from .geometry import STROKE_VERTEX_DTYPE, TessellationError, TooManyVertices, VertexBuffers
from .options import LineCap, LineJoin, StrokeOptions
from .path import Path, PathBatch, PathBuilder, SvgPathBuilder, parse_svg_path
from .sharding import ShardedStroker, exchange_totals, partition_by_cost
from .shapes import StrokeBuilder, StrokeTessellator
from .tessellator import (DeviceBatch, DeviceTessellator, HostPipeline, device_mesh_to_host, flatten_cubics, flatten_quadratics,
                          tessellate_paths)

__all__ = [
    "Path", "PathBatch", "PathBuilder", "SvgPathBuilder", "parse_svg_path",
    "LineCap", "LineJoin", "StrokeOptions",
    "VertexBuffers", "STROKE_VERTEX_DTYPE", "TessellationError", "TooManyVertices",
    "DeviceBatch", "DeviceTessellator", "HostPipeline", "StrokeTessellator", "StrokeBuilder", "tessellate_paths",
    "device_mesh_to_host",
    "flatten_cubics", "flatten_quadratics", "ShardedStroker", "exchange_totals", "partition_by_cost",
]
