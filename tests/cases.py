"""Shared seeded inputs for the parity tests (same on the CPU and the GPU box)."""
import numpy as np

from onsas_jl_b200 import meshgen as mg
from oracle import oracle as O

MATERIALS = {  # kind, p0, p1
    "svk": (O.MAT_SVK, 0.5769, 0.3846),
    "neo": (O.MAT_NEOHOOKEAN, 0.5769 + 2 * 0.3846 / 3, 0.3846),
    "iso": (O.MAT_ISOLINEAR, 1.0, 0.3),
}


def random_tets(n, seed=20240601):
    """SURVEY.md 8d per-element parity set: unit right tet + vertex jitter U(-0.2,0.2), reject vol <= 0.02,
    u ~ U(-0.3,0.3)^12; numpy.random.default_rng(20240601)."""
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=float)  # positive volume in the reference ordering
    X = np.zeros((0, 4, 3))
    while len(X) < n:
        cand = base[None] + rng.uniform(-0.2, 0.2, (n, 4, 3))
        J = np.stack([cand[:, 0] - cand[:, 1], cand[:, 3] - cand[:, 1], cand[:, 2] - cand[:, 1]], axis=2)
        vol = np.linalg.det(J) / 6
        X = np.concatenate([X, cand[vol > 0.02]])
    X = X[:n]
    u = rng.uniform(-0.3, 0.3, (n, 12))
    return X, u


def random_tet_model(n, mat="svk", seed=20240601):
    """n disconnected random tets as a FlatModel + displacement vector."""
    X, u = random_tets(n, seed)
    kind, p0, p1 = MATERIALS[mat]
    xyz = X.reshape(-1, 3)
    tets = np.arange(4 * n, dtype=np.int32).reshape(n, 4)
    m = O.FlatModel(xyz=xyz, tets=tets, mat_kind=[kind], mat_params=[[p0, p1]], free_dofs=np.arange(12 * n))
    return m, u.reshape(-1)


def box_model(nx, ny, nz, mat="svk", jitter=0.0, seed=1, E=1.0, nu=0.3):
    """Uniaxial box problem of the examples as a FlatModel."""
    mesh = mg.box_tet_mesh(nx, ny, nz)
    xyz = mesh.xyz.copy()
    if jitter:
        rng = np.random.default_rng(seed)
        h = min(2.0 / nx, 1.0 / ny, 1.0 / nz)
        xyz += rng.uniform(-jitter * h, jitter * h, xyz.shape)
    lam, G = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    if mat == "svk":
        kind, p = O.MAT_SVK, (lam, G)
    elif mat == "neo":
        kind, p = O.MAT_NEOHOOKEAN, (E / (3 * (1 - 2 * nu)), G)
    else:
        kind, p = O.MAT_ISOLINEAR, (E, nu)
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    m = O.FlatModel(xyz=xyz, tets=mesh.tets, mat_kind=[kind], mat_params=[p], free_dofs=free)
    return m, mesh


def random_U(m, amp=0.05, seed=7):
    return np.random.default_rng(seed).uniform(-amp, amp, m.n_dofs)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def von_mises_truss(strain):
    """examples/von_misses_truss/von_misses_truss.jl:8-62 as a FlatModel."""
    E, A, theta, Ltr = 210e9, 2.5e-3, 65.0, 2.0
    V, H = Ltr * np.cos(np.deg2rad(theta)), Ltr * np.sin(np.deg2rad(theta))
    xyz = np.array([[0.0, 0.0, 0.0], [V, 0.0, H], [2 * V, 0.0, 0.0]])
    d = np.sqrt(4 * A / np.pi)
    a = np.sqrt(A)
    areas = [np.pi * d ** 2 / 4, a ** 2]  # Circle(d), Square(a)
    G = E / 2  # nu = 0
    m = O.FlatModel(xyz=xyz, dim=3, trusses=[[0, 1], [1, 2]], truss_area=areas, truss_strain=strain,
                    mat_kind=[O.MAT_SVK], mat_params=[[0.0, G]], free_dofs=[3, 5])
    Fk = -1e8

    def fext(t):
        F = np.zeros(9)
        F[5] = Fk * t
        return F
    return m, fext, dict(E=E, A=A, H=H, V=V, L=Ltr, Fk=Fk)


def clamped_truss(N=100):
    """examples/clamped_truss/clamped_truss.jl:11-61 (1D chain, Green strain) as a FlatModel."""
    E, nu, Ltot, A, F = 30e6, 0.3, 200.0, 1.0, 10e6
    xyz = np.linspace(0, Ltot, N + 1).reshape(-1, 1)
    lam, G = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    bars = np.stack([np.arange(N), np.arange(1, N + 1)], axis=1)
    m = O.FlatModel(xyz=xyz, dim=1, trusses=bars, truss_area=np.full(N, A), truss_strain=O.STRAIN_GREEN,
                    mat_kind=[O.MAT_SVK], mat_params=[[lam, G]], free_dofs=np.arange(1, N + 1))

    def fext(t):
        Fv = np.zeros(N + 1)
        Fv[-1] = F * t
        return Fv
    return m, fext, dict(E=E, A=A, L=Ltot, F=F)


def svk_uniaxial_state(traction, E=1.0, nu=0.3):
    """Stretches (alpha, beta) of an SVK bar under the nominal traction P11 = traction with free lateral faces:
    alpha^3 - alpha = 2 traction / E (the reference's load_factors_analytic, examples/uniaxial_extension/
    uniaxial_extension.jl:151-154), beta^2 = 1 - nu (alpha^2 - 1)."""
    a = 1.0 + traction / E
    for _ in range(100):
        f, df = a ** 3 - a - 2.0 * traction / E, 3 * a * a - 1.0
        a, d = a - f / df, f / df
        if abs(d) < 1e-16 * abs(a):
            break
    return float(a), float(np.sqrt(1.0 - nu * (a * a - 1.0)))
