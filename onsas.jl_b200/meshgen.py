"""Synthetic structured meshes of the reference's example problems.

Host-side input generation only (numpy); nothing here is on the hot path.  Every mesh uses the
reference's own hexahedron -> 6 tetrahedra split of `examples/uniaxial_extension/uniaxial_extension.jl:45-72`
(all six tets share the body diagonal n4-n6, positive volumes, faces conform by translation) and
the dof numbering of `Meshes.jl:85-98` (dof = 3*node + c, 0-based here).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# local corner (dx,dy,dz) of n1..n8 in uniaxial_extension.jl:45-52
_CORNERS = np.array([(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0), (1, 0, 0), (1, 0, 1), (1, 1, 1), (1, 1, 0)], dtype=np.int64)
# tets t1..t6, 1-based corner labels, uniaxial_extension.jl:67-72
_TETS = np.array([(1, 4, 2, 6), (6, 2, 3, 4), (4, 3, 6, 7), (4, 1, 5, 6), (4, 6, 5, 8), (4, 7, 6, 8)], dtype=np.int64) - 1
# boundary triangles per hex face, 1-based corner labels, uniaxial_extension.jl:54-61
_FACE_TRIS = {
    "x1": [(5, 8, 6), (6, 8, 7)],  # loaded face x = Lx (f1, f2)
    "x0": [(4, 1, 2), (4, 2, 3)],  # f3, f4
    "y0": [(6, 2, 1), (6, 1, 5)],  # f5, f6
    "z0": [(1, 4, 5), (4, 8, 5)],  # f7, f8
}


@dataclass
class TetMesh:
    xyz: np.ndarray                     # (n_nodes, 3) float64
    tets: np.ndarray                    # (n_tets, 4) int32, 0-based
    faces: dict = field(default_factory=dict)      # name -> (n, 3) int32 boundary triangles
    node_sets: dict = field(default_factory=dict)  # name -> int64 node ids
    grid: tuple = ()

    @property
    def n_nodes(self):
        return self.xyz.shape[0]

    @property
    def n_tets(self):
        return self.tets.shape[0]


def _hex_tets(corner_ids: np.ndarray) -> np.ndarray:
    """corner_ids: (n_hex, 8) node id of local corners n1..n8 -> (n_hex*6, 4) tets, hex-major."""
    return corner_ids[:, _TETS].reshape(-1, 4)


def box_tet_mesh(nx: int, ny: int, nz: int, Lx: float = 2.0, Ly: float = 1.0, Lz: float = 1.0) -> TetMesh:
    """Box [0,Lx]x[0,Ly]x[0,Lz] with nx*ny*nz hexes; node (i,j,k) id = i + (nx+1)(j + (ny+1)k)."""
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    nid = (i + (nx + 1) * (j + (ny + 1) * k)).astype(np.int64)
    xyz = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    xyz[nid.ravel(), 0] = (i * (Lx / nx)).ravel()
    xyz[nid.ravel(), 1] = (j * (Ly / ny)).ravel()
    xyz[nid.ravel(), 2] = (k * (Lz / nz)).ravel()
    # hexes ordered x-fastest, like the nodes
    hk, hj, hi = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    hi, hj, hk = hi.ravel(), hj.ravel(), hk.ravel()
    corners = np.stack([nid[hi + d[0], hj + d[1], hk + d[2]] for d in _CORNERS], axis=1)
    tets = _hex_tets(corners).astype(np.int32)

    def face(sel, name):
        tri = np.array(_FACE_TRIS[name]) - 1
        return corners[sel][:, tri].reshape(-1, 3).astype(np.int32)

    faces = {"x1": face(hi == nx - 1, "x1"), "x0": face(hi == 0, "x0"), "y0": face(hj == 0, "y0"),
             "z0": face(hk == 0, "z0")}
    node_sets = {"x0": nid[0, :, :].ravel(), "x1": nid[nx, :, :].ravel(), "y0": nid[:, 0, :].ravel(),
                 "z0": nid[:, :, 0].ravel()}
    return TetMesh(xyz, tets, faces, node_sets, (nx, ny, nz))


def cylinder_tet_mesh(nr: int, nt: int, nz: int, Ri: float = 100.0, Re: float = 200.0, Lz: float = 30.0) -> TetMesh:
    """Hollow cylinder of examples/cylinder_internal_pressure (Ri, Re, Lz :14-29) on a structured
    (r, theta, z) grid, periodic in theta; nt must be a multiple of 4 so that nodes lie on the axes.
    Node (i,j,k) id = i + (nr+1)(j + nt*k)."""
    assert nt % 4 == 0 and nt >= 8
    i, j, k = np.meshgrid(np.arange(nr + 1), np.arange(nt), np.arange(nz + 1), indexing="ij")
    nid = (i + (nr + 1) * (j + nt * k)).astype(np.int64)
    r = Ri + (Re - Ri) * i / nr
    th = 2 * np.pi * j / nt
    xyz = np.zeros(((nr + 1) * nt * (nz + 1), 3))
    xyz[nid.ravel(), 0] = (r * np.cos(th)).ravel()
    xyz[nid.ravel(), 1] = (r * np.sin(th)).ravel()
    xyz[nid.ravel(), 2] = (k * (Lz / nz)).ravel()
    # snap the axis nodes exactly onto the axes (cos/sin round-off would leave 1e-14 offsets)
    for jj, (cx, cy) in {0: (1, 0), nt // 4: (0, 1), nt // 2: (-1, 0), 3 * nt // 4: (0, -1)}.items():
        ids = nid[:, jj, :].ravel()
        rr = (Ri + (Re - Ri) * i[:, jj, :] / nr).ravel()
        xyz[ids, 0] = rr * cx
        xyz[ids, 1] = rr * cy
    hk, hj, hi = np.meshgrid(np.arange(nz), np.arange(nt), np.arange(nr), indexing="ij")
    hi, hj, hk = hi.ravel(), hj.ravel(), hk.ravel()
    corners = np.stack([nid[hi + d[0], (hj + d[1]) % nt, hk + d[2]] for d in _CORNERS], axis=1)
    tets = _hex_tets(corners).astype(np.int32)
    # inner surface r = Ri (local x0 face); orient so the normal points to the axis (out of the solid)
    sel = hi == 0
    tri = np.array(_FACE_TRIS["x0"]) - 1
    inner = corners[sel][:, tri].reshape(-1, 3)
    a, b, c = xyz[inner[:, 0]], xyz[inner[:, 1]], xyz[inner[:, 2]]
    nrm = np.cross(b - a, c - a)
    cen = (a + b + c) / 3
    flip = (nrm[:, 0] * cen[:, 0] + nrm[:, 1] * cen[:, 1]) > 0  # pointing away from the axis -> flip
    inner[flip] = inner[flip][:, [0, 2, 1]]
    faces = {"inner": inner.astype(np.int32)}
    node_sets = {
        "z_caps": np.concatenate([nid[:, :, 0].ravel(), nid[:, :, nz].ravel()]),
        # cylinder_mesh.jl:142-154: u_x fixed at the outer nodes on the y axis, u_y at those on the x axis
        "outer_on_y_axis": np.array([nid[nr, nt // 4, 0], nid[nr, nt // 4, nz], nid[nr, 3 * nt // 4, 0], nid[nr, 3 * nt // 4, nz]]),
        "outer_on_x_axis": np.array([nid[nr, 0, 0], nid[nr, 0, nz], nid[nr, nt // 2, 0], nid[nr, nt // 2, nz]]),
    }
    return TetMesh(xyz, tets, faces, node_sets, (nr, nt, nz))


@dataclass
class TrussMesh:
    xyz: np.ndarray      # (n_nodes, dim)
    bars: np.ndarray     # (n_bars, 2) int32
    node_sets: dict = field(default_factory=dict)
    grid: tuple = ()

    @property
    def n_nodes(self):
        return self.xyz.shape[0]

    @property
    def n_bars(self):
        return self.bars.shape[0]


def truss_lattice(nx: int, ny: int, nz: int, L: float = 2.0) -> TrussMesh:
    """Braced cubic space-truss lattice: axis bars, one diagonal per cell face orientation and one body
    diagonal per cell (an unbraced cubic lattice has a singular tangent at U = 0)."""
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    nid = (i + (nx + 1) * (j + (ny + 1) * k)).astype(np.int64)
    xyz = np.zeros((nid.size, 3))
    xyz[nid.ravel()] = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1) * L
    bars = []

    def add(a, b):
        bars.append(np.stack([a.ravel(), b.ravel()], axis=1))

    add(nid[:-1, :, :], nid[1:, :, :])
    add(nid[:, :-1, :], nid[:, 1:, :])
    add(nid[:, :, :-1], nid[:, :, 1:])
    add(nid[:-1, :-1, :], nid[1:, 1:, :])    # xy face diagonals
    add(nid[:-1, :, :-1], nid[1:, :, 1:])    # xz
    add(nid[:, :-1, :-1], nid[:, 1:, 1:])    # yz
    add(nid[:-1, :-1, :-1], nid[1:, 1:, 1:])  # body diagonal
    bars = np.concatenate(bars).astype(np.int32)
    order = np.lexsort((bars[:, 1], bars[:, 0]))  # by first node: keeps the row-owner pairs local
    node_sets = {"x0": nid[0].ravel(), "x1": nid[nx].ravel()}
    return TrussMesh(xyz, bars[order], node_sets, (nx, ny, nz))


# ---------------------------------------------------------------------------------------------
# boundary conditions of the example problems, flattened (reference: StructuralBoundaryConditions.jl:170-220)

def face_areas(xyz: np.ndarray, tri: np.ndarray) -> np.ndarray:
    """|1/2 (x2-x1) x (x3-x1)|  (TriangularFaces.jl:44-54)."""
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    return 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)


def global_face_load(n_nodes: int, xyz: np.ndarray, tri: np.ndarray, traction) -> np.ndarray:
    """F_ext of GlobalLoad(:u, t -> traction) on triangular faces: traction*A/3 on each face node,
    duplicates summed (GlobalLoadBoundaryConditions.jl:50-68, StructuralBoundaryConditions.jl:195-220)."""
    F = np.zeros((n_nodes, 3))
    A = face_areas(xyz, tri)
    contrib = (A / 3.0)[:, None] * np.asarray(traction, dtype=np.float64)[None, :]
    for c in range(3):
        np.add.at(F, tri[:, c], contrib)
    return F.ravel()


def pressure_face_load(n_nodes: int, xyz: np.ndarray, tri: np.ndarray, p: float) -> np.ndarray:
    """F_ext of Pressure(:u, t -> p): -n p A/3 per face node (LocalLoadBoundaryConditions.jl:36-56)."""
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    avec = 0.5 * np.cross(b - a, c - a)  # n * A
    F = np.zeros((n_nodes, 3))
    contrib = -p * avec / 3.0
    for k in range(3):
        np.add.at(F, tri[:, k], contrib)
    return F.ravel()


def free_dofs_from_fixed(n_nodes: int, dim: int, fixed: dict) -> np.ndarray:
    """free dofs = all node dofs in node order minus FixedField dofs (Structures.jl:129-142).
    fixed: component (0-based) -> node ids."""
    mask = np.ones(n_nodes * dim, dtype=bool)
    for comp, nodes in fixed.items():
        mask[np.asarray(nodes, dtype=np.int64) * dim + comp] = False
    return np.nonzero(mask)[0].astype(np.int64)


def uniaxial_fixed(mesh: TetMesh) -> dict:
    """u_x = 0 on x=0, u_y = 0 on y=0, u_z = 0 on z=0 (uniaxial_extension.jl:93-103)."""
    return {0: mesh.node_sets["x0"], 1: mesh.node_sets["y0"], 2: mesh.node_sets["z0"]}


def homogeneous_field(xyz: np.ndarray, alpha: float, beta: float) -> np.ndarray:
    """u = ((alpha-1)x, (beta-1)y, (beta-1)z): the analytic uniaxial solution, exact on any tet mesh."""
    return (xyz * (np.array([alpha, beta, beta]) - 1.0)[None, :]).ravel()
