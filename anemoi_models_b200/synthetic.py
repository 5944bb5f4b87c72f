"""Deterministic synthetic grids and edge sets for benchmarks and full-size tests (SURVEY.md 8d recipes).

The reference ships no graph builder (graphs come from anemoi-graphs as `sub_graph`), so the named shapes are
re-created by closed-form recipes: octahedral reduced Gaussian grids `oN`, a Fibonacci sphere standing in for
`n320`, encoder edges by cut-off radius (0.6 x the largest nearest-neighbour distance of the dst grid, the
anemoi-graphs default), k-nearest-neighbour edges for decoders / processors.  Host-side numpy/scipy, one-off.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def octahedral_rows(N: int):
    """(latitudes[2N] north->south in radians, points per row[2N]) of the octahedral reduced Gaussian grid oN."""
    x, _ = np.polynomial.legendre.leggauss(2 * N)
    lat = np.arcsin(x)[::-1]  # north -> south
    r = np.arange(2 * N)
    r = np.minimum(r, 2 * N - 1 - r)  # row index counted from the nearer pole
    return lat, 20 + 4 * r


def octahedral_grid(N: int, row_lo: int = 0, row_hi: Optional[int] = None) -> Tuple[np.ndarray, int]:
    """xyz of rows [row_lo, row_hi) of oN (points ordered north->south, west->east) and the global index of the
    first returned point.  o48: 10,944 points, o96: 40,320, o1280: 6,599,680."""
    lat, npts = octahedral_rows(N)
    row_hi = 2 * N if row_hi is None else row_hi
    first = int(npts[:row_lo].sum())
    lats = np.repeat(lat[row_lo:row_hi], npts[row_lo:row_hi])
    lons = np.concatenate([np.arange(n) * (2 * np.pi / n) for n in npts[row_lo:row_hi]]) if row_hi > row_lo else np.zeros(0)
    return latlon_to_xyz(lats, lons), first


def octahedral_size(N: int) -> int:
    return 4 * N * N + 36 * N


def fibonacci_sphere(M: int, lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
    """xyz of points [lo, hi) of an M-point Fibonacci sphere, ordered north->south (index order = latitude order)."""
    hi = M if hi is None else hi
    i = np.arange(lo, hi, dtype=np.float64)
    z = 1.0 - (2.0 * i + 1.0) / M
    phi = i * (np.pi * (3.0 - np.sqrt(5.0)))
    rxy = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    return np.stack([rxy * np.cos(phi), rxy * np.sin(phi), z], axis=1)


def latlon_to_xyz(lat: np.ndarray, lon: np.ndarray) -> np.ndarray:
    return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=1)


def max_nn_distance(xyz: np.ndarray) -> float:
    from scipy.spatial import cKDTree

    d, _ = cKDTree(xyz).query(xyz, k=2, workers=-1)
    return float(d[:, 1].max())


def cutoff_edges(src_xyz: np.ndarray, dst_xyz: np.ndarray, radius: float, src_offset: int = 0, dst_offset: int = 0) -> np.ndarray:
    """edge_index [2,E] int64: every src within chord distance `radius` of a dst; grouped by dst ascending,
    src ascending inside a dst."""
    from scipy.spatial import cKDTree

    tree = cKDTree(src_xyz)
    nbrs = tree.query_ball_point(dst_xyz, radius, return_sorted=True, workers=-1)
    counts = np.fromiter((len(n) for n in nbrs), dtype=np.int64, count=len(nbrs))
    src = np.concatenate([np.asarray(n, dtype=np.int64) for n in nbrs]) if counts.sum() else np.zeros(0, np.int64)
    dst = np.repeat(np.arange(len(nbrs), dtype=np.int64), counts)
    return np.stack([src + src_offset, dst + dst_offset])


def knn_edges(src_xyz: np.ndarray, dst_xyz: np.ndarray, k: int, exclude_self: bool = False) -> np.ndarray:
    """edge_index [2,E]: the k nearest src of every dst (decoder: k=3; processor: k=8 without self loops)."""
    from scipy.spatial import cKDTree

    kk = k + 1 if exclude_self else k
    _, idx = cKDTree(src_xyz).query(dst_xyz, k=kk, workers=-1)
    idx = idx.reshape(len(dst_xyz), kk)
    if exclude_self:
        idx = idx[:, 1:]
    dst = np.repeat(np.arange(len(dst_xyz), dtype=np.int64), k)
    return np.stack([idx.reshape(-1).astype(np.int64), dst])


def encoder_graph(src_points: int, dst_N: int, cutoff: float = 0.6):
    """Whole `fibonacci(src_points) -> o<dst_N>` cut-off encoder graph (headline: 542,080 -> o96)."""
    dst_xyz, _ = octahedral_grid(dst_N)
    radius = cutoff * max_nn_distance(dst_xyz)
    ei = cutoff_edges(fibonacci_sphere(src_points), dst_xyz, radius)
    return ei, src_points, len(dst_xyz), radius


def encoder_graph_band(src_points: int, dst_N: int, parts: int, part: int, cutoff: float = 0.6, bounds=None):
    """Edges (GLOBAL node ids) whose dst lies in shard `part` -- what one rank of a dst-sharded run owns.  The shards are
    `tensor_split(arange(Nd), parts)` (the reference's shapes) unless `bounds` (parts+1 dst cut points) is given.  Only the
    src points of the matching latitude band are generated."""
    from .distributed.shapes import tensor_split_sizes

    dst_xyz, _ = octahedral_grid(dst_N)
    nd = len(dst_xyz)
    radius = cutoff * max_nn_distance(dst_xyz)
    if bounds is None:
        sizes = tensor_split_sizes(nd, parts)
        lo = sum(sizes[:part])
        hi = lo + sizes[part]
    else:
        assert len(bounds) == parts + 1 and bounds[0] == 0 and bounds[-1] == nd
        lo, hi = int(bounds[part]), int(bounds[part + 1])
    band = dst_xyz[lo:hi]
    # Fibonacci index range covering the band's z range plus the cut-off radius
    zmax, zmin = band[:, 2].max() + radius, band[:, 2].min() - radius
    i_lo = max(0, int(np.floor((1.0 - zmax) * src_points / 2.0)) - 1)
    i_hi = min(src_points, int(np.ceil((1.0 - zmin) * src_points / 2.0)) + 1)
    ei = cutoff_edges(fibonacci_sphere(src_points, i_lo, i_hi), band, radius, src_offset=i_lo, dst_offset=lo)
    return ei, src_points, nd, radius


def fibonacci_max_nn_distance(M: int) -> float:
    """largest nearest-neighbour chord distance of the M-point Fibonacci sphere (exact, one KD-tree query over all points)."""
    return max_nn_distance(fibonacci_sphere(M))


def o1280_to_n320_band(parts: int, part: int, src_N: int = 1280, dst_points: int = 542080, cutoff: float = 0.6,
                       dst_bounds=None, radius: Optional[float] = None):
    """BASELINE configs[4] / SURVEY 8d+8e "enc o1280 -> n320": src = octahedral grid o<src_N> (o1280: 6,599,680 points), dst = the
    Fibonacci stand-in for n320 (542,080 points), every src within 0.6 x the largest nearest-neighbour distance of the dst grid.
    Returns the edges (GLOBAL ids, grouped by dst ascending, src ascending inside a dst) into dst shard `part` of `parts`
    (equal-count `tensor_split` shards unless `dst_bounds` gives P+1 cut points), generating only the octahedral rows of the
    matching latitude band: (edge_index [2,E] int64, Ns, Nd, radius).  Whole graph: E = 7,458,659, in-degree 9-45."""
    from .distributed.shapes import tensor_split_sizes

    if radius is None:
        radius = cutoff * fibonacci_max_nn_distance(dst_points)
    if dst_bounds is None:
        sizes = tensor_split_sizes(dst_points, parts)
        lo = sum(sizes[:part])
        hi = lo + sizes[part]
    else:
        lo, hi = int(dst_bounds[part]), int(dst_bounds[part + 1])
    ns = octahedral_size(src_N)
    if hi <= lo:
        return np.zeros((2, 0), np.int64), ns, dst_points, radius
    band = fibonacci_sphere(dst_points, lo, hi)
    zmax, zmin = min(1.0, band[:, 2].max() + radius), max(-1.0, band[:, 2].min() - radius)
    lat, _ = octahedral_rows(src_N)
    z_rows = np.sin(lat)  # north -> south, decreasing
    rows = np.nonzero((z_rows <= zmax + 1e-9) & (z_rows >= zmin - 1e-9))[0]
    if len(rows) == 0:
        return np.zeros((2, 0), np.int64), ns, dst_points, radius
    r_lo, r_hi = max(0, int(rows[0]) - 1), min(2 * src_N, int(rows[-1]) + 2)
    src_xyz, first = octahedral_grid(src_N, r_lo, r_hi)
    ei = cutoff_edges(src_xyz, band, radius, src_offset=first, dst_offset=lo)
    return ei, ns, dst_points, radius


def encoder_work_balanced_bounds(src_points: int, dst_N: int, parts: int, cutoff: float = 0.6, w_edge: float = 1.7,
                                 w_src: float = 1.0, w_dst: float = 0.0):
    """parts+1 dst cut points of the `fibonacci(src_points) -> o<dst_N>` cut-off graph that give every rank the same WORK when
    the src rows follow the dst shards in latitude (`aligned_src_bounds`): work = w_edge * edges + w_src * src rows
    (+ w_dst * dst rows).  The default weights are MEASURED: per-rank kernel times of a 4-GPU run (AB2_TRACE, profiles/r02/
    trace_r02ab_n4.txt) fit 2.08 ns per edge + 1.21 ns per src row -- an edge costs 1.7 src rows, not the 0.5 its share of the
    HBM bytes suggests (the k / v rows of an edge come through L2 once per edge); cut points from the byte weights (3 : 6 : 6)
    were measured SLOWER than equal-count shards.

    Closed form, no graph needed: the Fibonacci points are uniform in z, so the src rows under latitude row i of the octahedral
    grid are src_points * dz_i / 2; a fixed cut-off radius r (chord) covers a cap of area pi r^2, so every dst row has
    ~src_points r^2 / 4 edges.  Equal-COUNT dst shards (`tensor_split`) of an octahedral grid are not equal-area -- the grid is
    denser (per area) towards the equator -- which leaves the polar ranks of 8 with 1.13x the src rows of the single-GPU
    workload.  Cut points are snapped to whole latitude rows (a cut inside a row makes both neighbours need the src rows
    around it)."""
    lat, npts = octahedral_rows(dst_N)
    dst_xyz, _ = octahedral_grid(dst_N)
    radius = cutoff * max_nn_distance(dst_xyz)
    z = np.sin(lat)
    zedge = np.concatenate([[1.0], 0.5 * (z[:-1] + z[1:]), [-1.0]])
    src_under_row = src_points * (zedge[:-1] - zedge[1:]) / 2.0
    deg = src_points * radius * radius / 4.0
    row_cost = (w_edge * deg + w_dst) * npts + w_src * src_under_row
    csum = np.concatenate([[0.0], np.cumsum(row_cost)])  # cost above each row boundary
    first = np.concatenate([[0], np.cumsum(npts)])       # dst index of each row boundary
    cuts = []
    for r in range(1, parts):
        i = int(np.argmin(np.abs(csum - csum[-1] * r / parts)))
        i = min(max(i, len(cuts) + 1), len(npts) - (parts - r))  # strictly increasing, room for the ranks after it
        cuts.append(int(first[i]))
    return [0] + cuts + [int(npts.sum())]


def edge_balanced_bounds(in_degree: np.ndarray, parts: int):
    """P+1 dst cut points that give every rank (nearly) the same number of EDGES instead of the same number of dst rows
    (`tensor_split`, the reference's shard shapes, distributed/shapes.py:19-24).  The shapes argument of the blocks takes
    either; on the o1280 -> n320 graph (in-degree 9-45) equal-count shards leave the busiest of 8 ranks with 1.18x the mean."""
    csum = np.concatenate([[0], np.cumsum(in_degree.astype(np.int64))])
    total = int(csum[-1])
    targets = [(total * r) // parts for r in range(1, parts)]
    cuts = [int(np.searchsorted(csum, t, side="left")) for t in targets]
    return [0] + cuts + [len(in_degree)]


def icosahedron():
    """12 vertices / 20 faces of the unit icosahedron."""
    phi = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, phi, 0], [1, phi, 0], [-1, -phi, 0], [1, -phi, 0], [0, -1, phi], [0, 1, phi], [0, -1, -phi],
                  [0, 1, -phi], [phi, 0, -1], [phi, 0, 1], [-phi, 0, -1], [-phi, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def multiscale_icosahedral_mesh(refinement: int = 6):
    """Multi-scale icosahedral mesh: nodes of the finest level (refinement 6 -> 40,962), edge set = union of the edges of
    every level 0..refinement, both directions (E = 2 * 30 * (4**(refinement+1) - 1) / 3 = 327,660 for refinement 6).
    Returns (xyz[N,3], edge_index[2,E] int64 sorted by dst then src)."""
    verts, faces = icosahedron()
    verts = [tuple(x) for x in verts]
    edges = set()

    def add_face_edges(fs):
        for a, b, c in fs:
            for x, y in ((a, b), (b, c), (c, a)):
                edges.add((int(x), int(y)))
                edges.add((int(y), int(x)))

    add_face_edges(faces)
    for _ in range(refinement):
        mid = {}
        new_faces = []

        def midpoint(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in mid:
                m = np.asarray(verts[a]) + np.asarray(verts[b])
                m /= np.linalg.norm(m)
                verts.append(tuple(m))
                mid[key] = len(verts) - 1
            return mid[key]

        for a, b, c in faces:
            ab, bc, ca = midpoint(a, b), midpoint(b, c), midpoint(c, a)
            new_faces += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        faces = np.asarray(new_faces, dtype=np.int64)
        add_face_edges(faces)
    e = np.asarray(sorted(edges, key=lambda t: (t[1], t[0])), dtype=np.int64).T
    return np.asarray(verts), np.ascontiguousarray(e)
