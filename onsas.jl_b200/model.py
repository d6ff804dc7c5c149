"""Host-side mirror of the reference's model-building API for the hot path: the same names,
argument meaning and error behaviour as ONSAS.jl's Entities / Materials / CrossSections /
BoundaryConditions / Meshes / StructuralModel layers, written in Python because no Julia
toolchain exists in this image (INTEGRATION.md holds the Julia glue for a real ONSAS.jl checkout).

Only what feeds the hot path is here; everything is flattened once into the structure-of-arrays
form the C ABI takes (`Structure.flat`).  Large synthetic meshes skip the per-object layer through
`Structure.from_arrays`.  file:line citations are relative to the reference's src/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np

from . import _lib as L

# ------------------------------------------------------------------------------------------ materials


class AbstractMaterial:
    label: str = ""


class AbstractHyperElasticMaterial(AbstractMaterial):
    pass


def _lame(E, nu):
    return E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))


class SVK(AbstractHyperElasticMaterial):
    """Saint-Venant-Kirchhoff, SVK(lambda, G) or SVK(E=, nu=) (Materials/SVKMaterial.jl:25-54)."""
    kind = L.MAT_SVK

    def __init__(self, lam=None, G=None, rho=None, label="", *, E=None, nu=None):
        if E is not None:
            lam, G = _lame(E, nu)
        self.lam, self.G, self.rho, self.label = float(lam), float(G), rho, str(label)

    def params(self):
        return (self.lam, self.G)

    def lame_parameters(self):
        return self.lam, self.G

    def shear_modulus(self):
        return self.G

    def poisson_ratio(self):
        return self.lam / (2 * (self.lam + self.G))

    def elasticity_modulus(self):
        return self.G * (3 * self.lam + 2 * self.G) / (self.lam + self.G)

    def bulk_modulus(self):
        return self.lam + 2 * self.G / 3


class NeoHookean(AbstractHyperElasticMaterial):
    """NeoHookean(K, G) or NeoHookean(E=, nu=) (Materials/NeoHookeanMaterial.jl:25-56)."""
    kind = L.MAT_NEOHOOKEAN

    def __init__(self, K=None, G=None, rho=None, label="", *, E=None, nu=None):
        if E is not None:
            lam, G = _lame(E, nu)
            K = lam + 2 * G / 3
        self.K, self.G, self.rho, self.label = float(K), float(G), rho, str(label)

    def params(self):
        return (self.K, self.G)

    def bulk_modulus(self):
        return self.K

    def shear_modulus(self):
        return self.G

    def lame_parameters(self):
        return self.K - 2 * self.G / 3, self.G

    def elasticity_modulus(self):
        lam, G = self.lame_parameters()
        return G * (3 * lam + 2 * G) / (lam + G)

    def poisson_ratio(self):
        lam, G = self.lame_parameters()
        return lam / (2 * (lam + G))


class IsotropicLinearElastic(AbstractMaterial):
    """IsotropicLinearElastic(E, nu) or (lam=, G=) (Materials/IsotropicLinearElasticMaterial.jl:22-51)."""
    kind = L.MAT_ISOLINEAR

    def __init__(self, E=None, nu=None, rho=None, label="", *, lam=None, G=None):
        if lam is not None:
            E = G * (3 * lam + 2 * G) / (lam + G)
            nu = lam / (2 * (lam + G))
        self.E, self.nu, self.rho, self.label = float(E), float(nu), rho, str(label)

    def params(self):
        return (self.E, self.nu)

    def elasticity_modulus(self):
        return self.E

    def poisson_ratio(self):
        return self.nu

    def shear_modulus(self):
        return self.E / (2 * (1 + self.nu))

    def bulk_modulus(self):
        return self.E / (3 * (1 - 2 * self.nu))

    def lame_parameters(self):
        return _lame(self.E, self.nu)


# ------------------------------------------------------------------------------------------ cross sections

class Circle:  # CrossSections/Circles.jl
    def __init__(self, d):
        self.d = float(d)

    def area(self):
        return math.pi * self.d ** 2 / 4


class Square:  # CrossSections/Squares.jl
    def __init__(self, a):
        self.a = float(a)

    def area(self):
        return self.a ** 2


class Rectangle:  # CrossSections/Rectangles.jl
    def __init__(self, width_y, width_z):
        self.wy, self.wz = float(width_y), float(width_z)

    def area(self):
        return self.wy * self.wz


class GenericCrossSection:  # CrossSections/GenericCrossSections.jl
    def __init__(self, A, *_):
        self.A = float(A)

    def area(self):
        return self.A


# ------------------------------------------------------------------------------------------ entities

class RotatedEngineeringStrain:  # Trusses.jl:26-38
    code = L.STRAIN_ROTATED_ENGINEERING


class GreenStrain:  # Trusses.jl:40-45
    code = L.STRAIN_GREEN


class Node:
    """Node(x[, y[, z]]) (Entities/Nodes.jl:74-84); dofs: field symbol -> list of 1-based dofs."""

    def __init__(self, *x):
        if len(x) == 1 and not np.isscalar(x[0]):
            x = tuple(x[0])
        assert 1 <= len(x) <= 3, "Only 1D, 2D or 3D nodes are supported."
        self.x = np.asarray(x, dtype=np.float64)
        self.dofs: dict = {}

    @property
    def dim(self):
        return len(self.x)

    def coordinates(self):
        return self.x


class Tetrahedron:
    """Tetrahedron(n1, n2, n3, n4[, label]) (Entities/Tetrahedrons.jl:24-41)."""

    def __init__(self, n1, n2, n3, n4, label=""):
        self.nodes = (n1, n2, n3, n4)
        assert all(n.dim == 3 for n in self.nodes), "Nodes of a tetrahedron element must be 3D."
        self.label = str(label)


class Truss:
    """Truss(n1, n2, cross_section[, strain_model][, label]) (Entities/Trusses.jl:54-90)."""

    def __init__(self, n1, n2, cross_section, strain_model=RotatedEngineeringStrain, label=""):
        if isinstance(strain_model, str):
            strain_model, label = RotatedEngineeringStrain, strain_model
        self.nodes = (n1, n2)
        self.cross_section = cross_section
        self.strain_model = strain_model
        self.label = str(label)


class TriangularFace:
    """TriangularFace(n1, n2, n3[, label]) (Entities/TriangularFaces.jl)."""

    def __init__(self, n1, n2, n3, label=""):
        self.nodes = (n1, n2, n3)
        self.label = str(label)

    def _area_vec(self):  # :44-47
        c = [n.x for n in self.nodes]
        return 0.5 * np.cross(c[1] - c[0], c[2] - c[0])

    def area(self):  # :50-54
        A = float(np.linalg.norm(self._area_vec()))
        if A == 0:
            raise ValueError("Area of TriangularFace is zero. Check that nodes are not aligned.")
        return A

    def normal_direction(self):  # :62-65
        v = self._area_vec()
        return v / np.linalg.norm(v)


# ------------------------------------------------------------------------------------------ boundary conditions

class FixedField:
    """FixedField(:u, [components], label): zero Dirichlet by component, 1-based components
    (BoundaryConditions/FixedFieldBoundaryConditions.jl:23-48)."""

    def __init__(self, field, components, name=""):
        self.field, self.components, self.name = field, [int(c) for c in components], str(name)


class GlobalLoad:
    """GlobalLoad(:u, t -> vector, label) (BoundaryConditions/GlobalLoadBoundaryConditions.jl:21-89)."""

    def __init__(self, field, values: Callable, name=""):
        self.field, self.values, self.name = field, values, str(name)


class Pressure:
    """Pressure(:u, t -> scalar, label): -n p A/3 per face node (LocalLoadBoundaryConditions.jl:21-56)."""

    def __init__(self, field, values: Callable, name=""):
        self.field, self.values, self.name = field, values, str(name)


# ------------------------------------------------------------------------------------------ mesh / structure

class Mesh:
    """Mesh(; nodes, elements, faces) (Meshes/Meshes.jl:158-184)."""

    def __init__(self, nodes=(), elements=(), faces=()):
        self.nodes, self.elements, self.faces = list(nodes), list(elements), list(faces)


def set_dofs(mesh: Mesh, symbol, dofs_per_node: int):
    """set_dofs!(mesh, :u, n): node i (1-based) gets dofs max_dof + (i-1)n+1 .. max_dof + i n (Meshes.jl:85-98)."""
    if any(symbol in n.dofs for n in mesh.nodes):
        raise ValueError(f"Dof symbol {symbol} already exists.")
    max_dof = max((max((max(v) for v in n.dofs.values()), default=0) for n in mesh.nodes), default=0)
    for i, n in enumerate(mesh.nodes, start=1):
        h = max_dof + i * dofs_per_node
        n.dofs[symbol] = list(range(1 + h - dofs_per_node, h + 1))


class StructuralMaterial:
    """StructuralMaterial(mat => [elements], ...) (StructuralModel/StructuralMaterials.jl:25-47);
    iteration order = assembly order."""

    def __init__(self, *pairs):
        if len(pairs) == 1 and isinstance(pairs[0], dict):
            pairs = tuple(pairs[0].items())
        self.pairs = [(m, list(es)) for m, es in pairs]

    def __getitem__(self, label):
        for m, _ in self.pairs:
            if m.label == label:
                return m
        raise KeyError(label)


class StructuralBoundaryCondition:
    """StructuralBoundaryCondition(bc => [entities], ...) (StructuralModel/StructuralBoundaryConditions.jl)."""

    def __init__(self, *pairs):
        if len(pairs) == 1 and isinstance(pairs[0], (dict, list)):
            pairs = tuple(pairs[0].items()) if isinstance(pairs[0], dict) else tuple(pairs[0])
        self.pairs = [(bc, list(ents)) for bc, ents in pairs]

    def fixed(self):
        return [(bc, e) for bc, e in self.pairs if isinstance(bc, FixedField)]

    def loads(self):
        return [(bc, e) for bc, e in self.pairs if isinstance(bc, (GlobalLoad, Pressure))]


@dataclass
class FlatStructure:
    """What the C ABI receives: everything 0-based."""
    xyz: np.ndarray
    dim: int
    tets: np.ndarray
    tet_mat: np.ndarray
    trusses: np.ndarray
    truss_mat: np.ndarray
    truss_area: np.ndarray
    truss_strain: int
    mat_kind: np.ndarray
    mat_params: np.ndarray
    free_dofs: np.ndarray
    fext: Callable[[float], np.ndarray]
    materials: list = field(default_factory=list)
    # device-side loads: [(builder, factor_fn)] with builder(ctx) -> pattern id and factor_fn(t) -> float;
    # None when some load cannot be expressed as a face / nodal pattern (the host-evaluated fext(t) is uploaded then)
    load_patterns: list | None = None

    def apply_loads(self, ctx, t: float):
        """apply!(sa, load_bcs) of one load step (StructuralAnalyses.jl:228-241) on the device when every load is a
        face / nodal pattern: the patterns are built once per context, only the factors p_k(t) cross the boundary."""
        if self.load_patterns is None:
            ctx.set_Fext(self.fext(t))
            return
        if getattr(ctx, "_load_owner", None) is not self:
            ctx.clear_loads()
            for build, _ in self.load_patterns:
                build(ctx)
            ctx._load_owner = self
        ctx.apply_loads([f(t) for _, f in self.load_patterns])

    @property
    def n_nodes(self):
        return self.xyz.shape[0]

    @property
    def n_dofs(self):
        return self.n_nodes * self.dim


class Structure:
    """Structure(mesh, materials, bcs) (StructuralModel/Structures.jl:111-142)."""

    def __init__(self, mesh: Mesh, materials: StructuralMaterial, bcs: StructuralBoundaryCondition):
        self.mesh, self.materials, self.bcs = mesh, materials, bcs
        self.flat = self._flatten()
        self.free_dofs = self.flat.free_dofs + 1  # 1-based, node order, fixed removed (:129-142)

    @classmethod
    def from_arrays(cls, xyz, *, tets=None, tet_mat=None, trusses=None, truss_mat=None, truss_area=None,
                    truss_strain=RotatedEngineeringStrain, materials: Sequence[AbstractMaterial], free_dofs,
                    fext: Callable[[float], np.ndarray]):
        """Array-based constructor for large meshes: no per-entity Python objects."""
        self = cls.__new__(cls)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        if xyz.ndim == 1:
            xyz = xyz.reshape(-1, 1)
        tets = np.zeros((0, 4), np.int32) if tets is None else np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
        trusses = np.zeros((0, 2), np.int32) if trusses is None else np.ascontiguousarray(trusses, np.int32).reshape(-1, 2)
        mats = list(materials)
        self.mesh, self.materials, self.bcs = None, StructuralMaterial(*[(m, []) for m in mats]), None
        self.flat = FlatStructure(
            xyz=xyz, dim=xyz.shape[1], tets=tets,
            tet_mat=np.zeros(len(tets), np.int32) if tet_mat is None else np.ascontiguousarray(tet_mat, np.int32),
            trusses=trusses,
            truss_mat=np.zeros(len(trusses), np.int32) if truss_mat is None else np.ascontiguousarray(truss_mat, np.int32),
            truss_area=np.ones(len(trusses)) if truss_area is None else np.ascontiguousarray(truss_area, np.float64),
            truss_strain=truss_strain.code,
            mat_kind=np.array([m.kind for m in mats], np.int32),
            mat_params=np.array([m.params() for m in mats], np.float64).reshape(-1, 2),
            free_dofs=np.ascontiguousarray(free_dofs, np.int64), fext=fext, materials=mats)
        self.free_dofs = self.flat.free_dofs + 1
        return self

    # ---- flattening of the object graph
    def _flatten(self) -> FlatStructure:
        nodes = self.mesh.nodes
        dim = nodes[0].dim
        idx = {id(n): i for i, n in enumerate(nodes)}
        for i, n in enumerate(nodes):
            if "u" not in n.dofs and ":u" not in n.dofs:
                raise ValueError("Element doesn't have dofs with symbol :u.")  # Entities.jl:163-165
            d = n.dofs.get("u", n.dofs.get(":u"))
            if list(d) != list(range(dim * i + 1, dim * i + dim + 1)):
                raise NotImplementedError("libonsas_cuda needs the set_dofs!(mesh, :u, dim) numbering dof = dim*(i-1)+c")
        xyz = np.array([n.x for n in nodes], dtype=np.float64).reshape(len(nodes), dim)
        mats, tets, tet_mat, trusses, truss_mat, areas = [], [], [], [], [], []
        strain = None
        self._elem_slot = {}  # id(element) -> (family, index)
        for mi, (m, elems) in enumerate(self.materials.pairs):  # assembly order (StaticAnalyses.jl:105-106)
            mats.append(m)
            for e in elems:
                ids = [idx[id(n)] for n in e.nodes]
                if isinstance(e, Tetrahedron):
                    self._elem_slot[id(e)] = (L.FAMILY_TET, len(tets))
                    tets.append(ids)
                    tet_mat.append(mi)
                elif isinstance(e, Truss):
                    if not isinstance(m, AbstractHyperElasticMaterial):
                        raise TypeError("Not implemented.")  # Entities.jl:174-176 fallback
                    if strain is None:
                        strain = e.strain_model
                    elif strain is not e.strain_model:
                        raise NotImplementedError("one strain model per structure")
                    self._elem_slot[id(e)] = (L.FAMILY_TRUSS, len(trusses))
                    trusses.append(ids)
                    truss_mat.append(mi)
                    areas.append(e.cross_section.area())
                else:
                    raise NotImplementedError(f"element type {type(e).__name__} is outside the hot path")
        # fixed dofs (StructuralBoundaryConditions.jl:170-185, FixedFieldBoundaryConditions.jl:36-48)
        fixed = np.zeros(len(nodes) * dim, dtype=bool)
        for bc, ents in self.bcs.fixed():
            for ent in ents:
                for n in (ent.nodes if hasattr(ent, "nodes") else (ent,)):
                    for c in bc.components:
                        fixed[dim * idx[id(n)] + (c - 1)] = True
        free = np.nonzero(~fixed)[0].astype(np.int64)
        loads = self.bcs.loads()

        def fext(t: float) -> np.ndarray:
            """apply!(sa, load_bcs) (StructuralAnalyses.jl:228-241 + StructuralBoundaryConditions.jl:195-220)."""
            F = np.zeros(len(nodes) * dim)
            for bc, ents in loads:
                for ent in ents:
                    if isinstance(bc, Pressure):
                        vec = bc.values(t) * (-ent.normal_direction()) * ent.area()
                        per = vec / len(ent.nodes)  # LocalLoadBoundaryConditions.jl:36-56
                        targets = ent.nodes
                    elif isinstance(ent, Node):
                        per = np.asarray(bc.values(t), dtype=np.float64)  # GlobalLoad on a node :34-48
                        targets = (ent,)
                    elif isinstance(ent, TriangularFace):
                        per = np.asarray(bc.values(t), dtype=np.float64) * ent.area() / len(ent.nodes)  # :50-68
                        targets = ent.nodes
                    else:  # body load on an element :70-89
                        if isinstance(ent, Tetrahedron):
                            X = np.array([n.x for n in ent.nodes])
                            vol = abs(np.linalg.det(np.array([X[0] - X[1], X[3] - X[1], X[2] - X[1]]))) / 6
                        else:
                            vol = ent.cross_section.area() * np.linalg.norm(ent.nodes[1].x - ent.nodes[0].x)
                        per = np.asarray(bc.values(t), dtype=np.float64) * vol / len(ent.nodes)
                        targets = ent.nodes
                    for n in targets:
                        F[dim * idx[id(n)]: dim * idx[id(n)] + dim] += per[:dim]
            return F

        # the same loads as device patterns: one per Pressure BC, one per component of a GlobalLoad BC
        patterns: list | None = []
        for bc, ents in loads:
            ents = list(ents)
            if ents and all(isinstance(e, TriangularFace) for e in ents) and dim == 3:
                tri = np.array([[idx[id(n)] for n in e.nodes] for e in ents], np.int32)
                if isinstance(bc, Pressure):
                    patterns.append((lambda ctx, tri=tri: ctx.add_face_load(tri, 1, [1.0]), lambda t, bc=bc: float(bc.values(t))))
                else:
                    for c in range(dim):
                        e_c = np.eye(3)[c]
                        patterns.append((lambda ctx, tri=tri, e_c=e_c: ctx.add_face_load(tri, 0, e_c),
                                         lambda t, bc=bc, c=c: float(np.asarray(bc.values(t), dtype=np.float64)[c])))
            elif ents and all(isinstance(e, Node) for e in ents) and not isinstance(bc, Pressure):
                ids = np.array([idx[id(n)] for n in ents], np.int32)
                for c in range(dim):
                    e_c = np.eye(3)[c][:dim]
                    patterns.append((lambda ctx, ids=ids, e_c=e_c: ctx.add_nodal_load(ids, e_c),
                                     lambda t, bc=bc, c=c: float(np.asarray(bc.values(t), dtype=np.float64)[c])))
            else:   # body loads on elements etc.: evaluated on the host
                patterns = None
                break

        return FlatStructure(
            load_patterns=patterns,
            xyz=xyz, dim=dim, tets=np.array(tets, np.int32).reshape(-1, 4), tet_mat=np.array(tet_mat, np.int32),
            trusses=np.array(trusses, np.int32).reshape(-1, 2), truss_mat=np.array(truss_mat, np.int32),
            truss_area=np.array(areas, np.float64), truss_strain=(strain or RotatedEngineeringStrain).code,
            mat_kind=np.array([m.kind for m in mats], np.int32),
            mat_params=np.array([m.params() for m in mats], np.float64).reshape(-1, 2), free_dofs=free, fext=fext,
            materials=mats)

    def element_slot(self, e):
        return self._elem_slot[id(e)]

    def node_index(self, n):
        return self.mesh.nodes.index(n)

    @property
    def num_dofs(self):
        return self.flat.n_dofs

    @property
    def num_free_dofs(self):
        return len(self.flat.free_dofs)
