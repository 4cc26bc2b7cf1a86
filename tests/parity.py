"""Shared parity assertions: bit-exact topology, positions/normals/advancement within 1e-5 x bbox extent
(BASELINE.json north_star).  In practice the CUDA path reproduces the oracle byte for byte; `exact=True`
asserts that too."""
import numpy as np


def bbox_extent(batch) -> float:
    pts = batch.points
    if pts.shape[0] == 0:
        return 1.0
    return float(max(np.ptp(pts[:, 0]), np.ptp(pts[:, 1]), 1e-6))


def _close(a, b, tol) -> bool:
    """|a - b| <= tol everywhere, NaN only where the other side has NaN too (degenerate inputs give NaN normals in the
    reference as well: e.g. a zero-length edge normalised by its length)."""
    a, b = np.asarray(a), np.asarray(b)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    d = np.abs(np.where(na, 0, a) - np.where(nb, 0, b))
    return bool(d.size == 0 or d.max() <= tol)


def assert_mesh_parity(got, want, extent: float, exact: bool = True, what: str = ""):
    assert len(got.vertices) == len(want.vertices), f"{what}: vertex count {len(got.vertices)} != {len(want.vertices)}"
    assert len(got.indices) == len(want.indices), f"{what}: index count {len(got.indices)} != {len(want.indices)}"
    assert np.array_equal(got.indices, want.indices), f"{what}: index arrays differ"
    tol = 1e-5 * extent  # tolerance stated by north_star
    gv, wv = got.vertices, want.vertices
    if len(gv):
        assert _close(got.positions(), want.positions(), tol), f"{what}: positions"
        assert _close(gv["normal"], wv["normal"], tol), f"{what}: normals"
        assert _close(gv["advancement"], wv["advancement"], tol), f"{what}: advancement"
        assert np.array_equal(gv["src"], wv["src"]), f"{what}: vertex source / side"
        assert np.array_equal(gv["t"], wv["t"]), f"{what}: source t"
    if exact and len(gv):
        # byte-identical, except that a NaN may carry any sign / payload (x86 and the GPU canonicalise differently)
        gb = np.frombuffer(gv.tobytes(), np.uint32).reshape(-1, 8).copy()
        wb = np.frombuffer(wv.tobytes(), np.uint32).reshape(-1, 8).copy()
        fcols = [0, 1, 2, 3, 4, 5, 7]  # every field but `src` is an f32
        gn = np.isnan(gb[:, fcols].view(np.float32))
        wn = np.isnan(wb[:, fcols].view(np.float32))
        assert np.array_equal(gn, wn), f"{what}: NaN fields differ"
        gf, wf = gb[:, fcols], wb[:, fcols]
        gf[gn] = 0
        wf[wn] = 0
        assert np.array_equal(gf, wf) and np.array_equal(gb[:, 6], wb[:, 6]), f"{what}: vertices not byte-identical"
