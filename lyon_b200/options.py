"""`StrokeOptions`, `LineCap`, `LineJoin` mirroring lyon (crates/tessellation/src/lib.rs:315-439,
crates/path/src/lib.rs:170-198) and the `#[repr(C)]` struct passed across the C ABI
(include/lyon_cuda.h: lyon_stroke_options)."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, replace
from enum import IntEnum
from typing import Optional


class LineCap(IntEnum):
    Butt = 0
    Square = 1
    Round = 2


class LineJoin(IntEnum):
    Miter = 0
    MiterClip = 1
    Round = 2
    Bevel = 3


class CStrokeOptions(ctypes.Structure):
    _fields_ = [
        ("start_cap", ctypes.c_uint8),
        ("end_cap", ctypes.c_uint8),
        ("line_join", ctypes.c_uint8),
        ("_pad", ctypes.c_uint8),
        ("line_width", ctypes.c_float),
        ("variable_line_width", ctypes.c_int32),
        ("miter_limit", ctypes.c_float),
        ("tolerance", ctypes.c_float),
    ]


@dataclass(frozen=True)
class StrokeOptions:
    start_cap: LineCap = LineCap.Butt
    end_cap: LineCap = LineCap.Butt
    line_join: LineJoin = LineJoin.Miter
    line_width: float = 1.0
    variable_line_width: Optional[int] = None
    miter_limit: float = 4.0
    tolerance: float = 0.1

    MINIMUM_MITER_LIMIT = 1.0
    DEFAULT_MITER_LIMIT = 4.0
    DEFAULT_LINE_WIDTH = 1.0
    DEFAULT_TOLERANCE = 0.1

    @staticmethod
    def default() -> "StrokeOptions":
        return StrokeOptions()

    @staticmethod
    def with_tolerance_only(tolerance: float) -> "StrokeOptions":  # StrokeOptions::tolerance(t)
        return StrokeOptions(tolerance=tolerance)

    def with_tolerance(self, tolerance: float) -> "StrokeOptions":
        return replace(self, tolerance=float(tolerance))

    def with_line_cap(self, cap: LineCap) -> "StrokeOptions":
        return replace(self, start_cap=LineCap(cap), end_cap=LineCap(cap))

    def with_start_cap(self, cap: LineCap) -> "StrokeOptions":
        return replace(self, start_cap=LineCap(cap))

    def with_end_cap(self, cap: LineCap) -> "StrokeOptions":
        return replace(self, end_cap=LineCap(cap))

    def with_line_join(self, join: LineJoin) -> "StrokeOptions":
        return replace(self, line_join=LineJoin(join))

    def with_line_width(self, width: float) -> "StrokeOptions":
        return replace(self, line_width=float(width))

    def with_miter_limit(self, limit: float) -> "StrokeOptions":
        assert limit >= self.MINIMUM_MITER_LIMIT  # lib.rs:422-426
        return replace(self, miter_limit=float(limit))

    def with_variable_line_width(self, idx: int) -> "StrokeOptions":
        return replace(self, variable_line_width=int(idx))

    def to_c(self) -> CStrokeOptions:
        return CStrokeOptions(
            int(self.start_cap),
            int(self.end_cap),
            int(self.line_join),
            0,
            float(self.line_width),
            -1 if self.variable_line_width is None else int(self.variable_line_width),
            float(self.miter_limit),
            float(self.tolerance),
        )
