"""Shared parity assertions: bit-exact topology, positions/normals/advancement within 1e-5 x bbox extent
(BASELINE.json north_star).  In practice the CUDA path reproduces the oracle byte for byte; `exact=True`
asserts that too."""
import numpy as np


def bbox_extent(batch) -> float:
    pts = batch.points
    if pts.shape[0] == 0:
        return 1.0
    return float(max(np.ptp(pts[:, 0]), np.ptp(pts[:, 1]), 1e-6))


def assert_mesh_parity(got, want, extent: float, exact: bool = True, what: str = ""):
    assert len(got.vertices) == len(want.vertices), f"{what}: vertex count {len(got.vertices)} != {len(want.vertices)}"
    assert len(got.indices) == len(want.indices), f"{what}: index count {len(got.indices)} != {len(want.indices)}"
    assert np.array_equal(got.indices, want.indices), f"{what}: index arrays differ"
    tol = 1e-5 * extent  # tolerance stated by north_star
    gv, wv = got.vertices, want.vertices
    if len(gv):
        assert np.abs(got.positions() - want.positions()).max() <= tol, f"{what}: positions"
        assert np.abs(gv["normal"] - wv["normal"]).max() <= tol, f"{what}: normals"
        assert np.abs(gv["advancement"] - wv["advancement"]).max() <= tol, f"{what}: advancement"
        assert np.array_equal(gv["src"], wv["src"]), f"{what}: vertex source / side"
        assert np.array_equal(gv["t"], wv["t"]), f"{what}: source t"
    if exact:
        assert gv.tobytes() == wv.tobytes(), f"{what}: vertices not byte-identical"
