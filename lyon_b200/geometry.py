"""`VertexBuffers<StrokeVertexGpu, u32>` on the host (reference: geometry_builder.rs:262-286) and the
accessors of lyon's `StrokeVertex` (stroke.rs:2668-2747) evaluated on whole arrays."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# include/lyon_cuda.h: lyon_stroke_vertex (32 bytes)
STROKE_VERTEX_DTYPE = np.dtype(
    [
        ("position_on_path", np.float32, (2,)),
        ("normal", np.float32, (2,)),
        ("advancement", np.float32),
        ("half_width", np.float32),
        ("src", np.uint32),  # bits 0..30 endpoint id (Edge: `to`), bit 31 set = Side::Negative
        ("t", np.float32),  # 1.0 = VertexSource::Endpoint, else VertexSource::Edge{t}
    ]
)
assert STROKE_VERTEX_DTYPE.itemsize == 32

SIDE_NEGATIVE_BIT = np.uint32(0x80000000)


class TessellationError(Exception):
    """lyon's TessellationError (error.rs).  `code` is the negative C-ABI status."""

    def __init__(self, code: int, what: str):
        super().__init__(f"{what} (code {code})")
        self.code = code


class TooManyVertices(TessellationError):
    pass


@dataclass
class VertexBuffers:
    vertices: np.ndarray  # STROKE_VERTEX_DTYPE [V]
    indices: np.ndarray  # uint32 [I]

    @staticmethod
    def empty() -> "VertexBuffers":
        return VertexBuffers(np.zeros(0, STROKE_VERTEX_DTYPE), np.zeros(0, np.uint32))

    def positions(self) -> np.ndarray:
        """StrokeVertex::position() = position_on_path + normal * half_width (stroke.rs:2676-2678), in f32."""
        v = self.vertices
        return (v["position_on_path"] + v["normal"] * v["half_width"][:, None]).astype(np.float32)

    def side(self) -> np.ndarray:
        """Side::to_f32: +1 positive, -1 negative (crates/path/src/lib.rs:249-254)."""
        return np.where(self.vertices["src"] & SIDE_NEGATIVE_BIT, np.float32(-1), np.float32(1))

    def source_endpoint(self) -> np.ndarray:
        return self.vertices["src"] & np.uint32(0x7FFFFFFF)

    def line_width(self) -> np.ndarray:
        return self.vertices["half_width"] * np.float32(2)
