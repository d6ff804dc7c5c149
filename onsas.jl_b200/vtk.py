"""Result hand-back to disk: `.vtu` / `.pvd` files written straight from the flat per-step arrays
of `Solution` -- the step immediately after the Newton loop.

Mirrors the reference's `write_vtk(sol, filename, time_index; fields)` and `write_vtk(sol, base_filename; fields)`
(src/Interfaces/VTK.jl:209-262): an unstructured grid of VTK_TETRA / VTK_LINE cells (`to_vtkcell_type`, :76-79), the
point vector field "Displacement" with components ux, uy, uz (`FIELD_NAMES`, :158-164) and, per cell, one scalar
array per tensor component named by the reference's labels ("σxx" ... "τyx", "ϵxx" ... "γyx"), each taken at the
(i, j) its label names in `INDEX_MAP` (:122-124, :131-148).  The reference goes through WriteVTK.jl and a
Dictionary-per-element traversal (:180-198); here every array is one numpy slice of the flat result and is written
as raw appended binary (no compression, UInt64 headers), so a 40 M-tet step is a handful of large writes.
"""
from __future__ import annotations

import os
import struct
from typing import Sequence

import numpy as np

VTK_TETRA, VTK_LINE = 10, 3   # VTKCellTypes.VTK_TETRA / VTK_LINE (VTK.jl:76-78)

# VTK.jl:158-164 (component labels) and :122-124 (which (i, j) a label names; 0-based here)
STRESS_LABELS = ["σxx", "σyy", "σzz", "τyz", "τxz", "τxy", "τzy", "τzx", "τyx"]
STRAIN_LABELS = ["ϵxx", "ϵyy", "ϵzz", "γyz", "γxz", "γxy", "γzy", "γzx", "γyx"]
INDEX_MAP = {"xx": (0, 0), "yy": (1, 1), "zz": (2, 2), "xy": (0, 1), "yz": (1, 2), "zx": (2, 0),
             "yx": (1, 0), "zy": (2, 1), "xz": (0, 2)}
DISPLACEMENT_COMPONENTS = ["ux", "uy", "uz"]


def _component(label: str):
    for key, ij in INDEX_MAP.items():
        if key in label:
            return ij
    raise ValueError(f"Unexpected label {label} didnt match INDEX_MAP")   # ArgumentError of VTK.jl:145


class _Appended:
    """Collects the arrays of one file; every DataArray refers to its block by byte offset."""

    def __init__(self):
        self.blocks: list[np.ndarray] = []
        self.offset = 0

    def add(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        off = self.offset
        self.blocks.append(a)
        self.offset += 8 + a.nbytes
        return off

    def write(self, fh):
        fh.write(b'  <AppendedData encoding="raw">\n   _')
        for a in self.blocks:
            fh.write(struct.pack("<Q", a.nbytes))
            fh.write(memoryview(a).cast("B"))
        fh.write(b"\n  </AppendedData>\n")


_VTK_TYPE = {np.dtype("float64"): "Float64", np.dtype("int64"): "Int64", np.dtype("int32"): "Int32", np.dtype("uint8"): "UInt8"}


def _data_array(app: _Appended, a: np.ndarray, name: str, ncomp: int = 1, component_names: Sequence[str] | None = None) -> bytes:
    attrs = f'type="{_VTK_TYPE[a.dtype]}" Name="{name}" NumberOfComponents="{ncomp}" format="appended" offset="{app.add(a)}"'
    if component_names:
        attrs += "".join(f' ComponentName{k}="{c}"' for k, c in enumerate(component_names))
    return f"    <DataArray {attrs}/>\n".encode("utf-8")


def write_vtu(filename: str, xyz: np.ndarray, *, tets=None, trusses=None, point_data: dict | None = None,
              cell_data: dict | None = None) -> str:
    """One `.vtu` file from flat arrays.  `xyz` is (n_nodes, dim); cells are the tets (n, 4) followed by the trusses
    (n, 2) -- the order of the flat results; `point_data` maps name -> (array (n_nodes, k), component names or None),
    `cell_data` maps name -> array (n_cells,).  Returns the path written (".vtu" appended when missing, as WriteVTK does)."""
    if not filename.endswith(".vtu"):
        filename += ".vtu"
    xyz = np.asarray(xyz, dtype=np.float64)
    n_nodes, dim = xyz.shape
    pts = np.zeros((n_nodes, 3))
    pts[:, :dim] = xyz
    tets = np.zeros((0, 4), np.int64) if tets is None else np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    trusses = np.zeros((0, 2), np.int64) if trusses is None else np.asarray(trusses, dtype=np.int64).reshape(-1, 2)
    n_cells = len(tets) + len(trusses)
    conn = np.concatenate([tets.ravel(), trusses.ravel()])
    offsets = np.concatenate([4 * np.arange(1, len(tets) + 1, dtype=np.int64),
                              4 * len(tets) + 2 * np.arange(1, len(trusses) + 1, dtype=np.int64)])
    types = np.concatenate([np.full(len(tets), VTK_TETRA, np.uint8), np.full(len(trusses), VTK_LINE, np.uint8)])
    for name, arr in (cell_data or {}).items():
        arr = np.asarray(arr)
        if arr.shape[0] != n_cells:
            raise ValueError(f"cell data {name}: {arr.shape[0]} values for {n_cells} cells")
        if not np.all(np.isfinite(arr)):
            raise AssertionError(f"cell data {name} holds NaN or Inf")   # the @assert pair of VTK.jl:139-140

    app = _Appended()
    out = [b'<?xml version="1.0" encoding="utf-8"?>\n',
           b'<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n',
           b" <UnstructuredGrid>\n",
           f'  <Piece NumberOfPoints="{n_nodes}" NumberOfCells="{n_cells}">\n'.encode(),
           b"   <Points>\n", _data_array(app, pts, "Points", 3), b"   </Points>\n",
           b"   <Cells>\n", _data_array(app, conn, "connectivity"), _data_array(app, offsets, "offsets"),
           _data_array(app, types, "types"), b"   </Cells>\n",
           b"   <PointData>\n"]
    for name, (arr, comps) in (point_data or {}).items():
        arr = np.asarray(arr, dtype=np.float64).reshape(n_nodes, -1)
        out.append(_data_array(app, arr, name, arr.shape[1], comps))
    out += [b"   </PointData>\n", b"   <CellData>\n"]
    for name, arr in (cell_data or {}).items():
        out.append(_data_array(app, np.asarray(arr, dtype=np.float64), name))
    out += [b"   </CellData>\n", b"  </Piece>\n", b" </UnstructuredGrid>\n"]
    with open(filename, "wb") as fh:
        fh.writelines(out)
        app.write(fh)
        fh.write(b"</VTKFile>\n")
    return filename


def _step_arrays(sol, time_index: int, fields):
    """Point and cell data of stored step `time_index` (1-based like the reference) from the flat Solution arrays."""
    flat = sol.analysis.s.flat
    k = time_index - 1
    if not 0 <= k < len(sol.U):
        raise IndexError(f"time_index {time_index} outside 1:{len(sol.U)}")
    dim = flat.dim
    point_data, cell_data = {}, {}
    if "u" in fields:
        u = np.zeros((flat.n_nodes, 3))
        u[:, :dim] = sol.U[k].reshape(-1, dim)
        point_data["Displacement"] = (u, DISPLACEMENT_COMPONENTS)
    for key, labels, tet_src, truss_src in (("σ", STRESS_LABELS, sol.tet_stress, sol.truss_stress),
                                             ("ϵ", STRAIN_LABELS, sol.tet_strain, sol.truss_strain)):
        if key not in fields:
            continue
        parts = []
        if len(flat.tets):
            parts.append(np.asarray(tet_src[k]).reshape(-1, 9))
        if len(flat.trusses):
            parts.append(np.asarray(truss_src[k]).reshape(-1, 9))
        if not parts:
            continue
        t9 = np.concatenate(parts)          # 3x3 column-major per element: entry (i, j) at i + 3 j
        for label in labels:
            i, j = _component(label)
            cell_data[label] = t9[:, i + 3 * j]
    return point_data, cell_data


DEFAULT_FIELDS = ("u", "σ", "ϵ")   # default_dof_fields(sol) of VTK.jl:147-153 for a displacement-only mesh


def write_vtk(sol, filename: str, time_index: int | None = None, *, fields: Sequence[str] = DEFAULT_FIELDS):
    """write_vtk(sol, filename, time_index; fields) -> path of the `.vtu` (VTK.jl:209-232);
    write_vtk(sol, base_filename; fields) -> path of the `.pvd` collection with one `.vtu` per stored step, the
    time of each step taken from the analysis' load-factor vector (VTK.jl:245-262)."""
    flat = sol.analysis.s.flat
    if time_index is not None:
        point_data, cell_data = _step_arrays(sol, time_index, fields)
        return write_vtu(filename, flat.xyz, tets=flat.tets, trusses=flat.trusses, point_data=point_data, cell_data=cell_data)
    times = list(sol.analysis.load_factors())[:len(sol.U)]
    entries = []
    for idx, t in enumerate(times, start=1):
        path = write_vtk(sol, f"{filename}_timestep_{idx}.vtu", idx, fields=fields)
        entries.append((t, os.path.basename(path)))
    pvd = filename + ".pvd"
    with open(pvd, "w", encoding="utf-8") as fh:
        fh.write('<?xml version="1.0" encoding="utf-8"?>\n<VTKFile type="Collection" version="1.0" byte_order="LittleEndian">\n <Collection>\n')
        for t, name in entries:
            fh.write(f'  <DataSet timestep="{float(t)!r}" part="0" file="{name}"/>\n')
        fh.write(" </Collection>\n</VTKFile>\n")
    return pvd


def read_vtu(filename: str):
    """Minimal reader of the files `write_vtu` produces (tests and round trips): returns (arrays by name, attributes)."""
    raw = open(filename, "rb").read()
    marker = raw.index(b'<AppendedData encoding="raw">')
    start = raw.index(b"_", marker) + 1
    header = raw[:marker].decode("utf-8")
    import re
    arrays, meta = {}, {}
    m = re.search(r'NumberOfPoints="(\d+)" NumberOfCells="(\d+)"', header)
    meta["n_points"], meta["n_cells"] = int(m.group(1)), int(m.group(2))
    inv = {v: k for k, v in _VTK_TYPE.items()}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="([^"]+)" NumberOfComponents="(\d+)" format="appended" offset="(\d+)"([^/]*)/>', header):
        typ, name, ncomp, off = m.group(1), m.group(2), int(m.group(3)), int(m.group(4))
        nbytes = struct.unpack_from("<Q", raw, start + off)[0]
        a = np.frombuffer(raw, dtype=inv[typ], count=nbytes // inv[typ].itemsize, offset=start + off + 8)
        arrays[name] = a.reshape(-1, ncomp) if ncomp > 1 else a
        comps = re.findall(r'ComponentName\d+="([^"]+)"', m.group(5))
        if comps:
            meta.setdefault("component_names", {})[name] = comps
    return arrays, meta
