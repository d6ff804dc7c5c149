"""Golden hashes of the host-built tables (csrc/tables.cpp through tests/hostsim) for a set of meshes: structured and random
numbering, trusses in 1-3 D, both families, partial ownership, a 400-bar hub.  tests/test_hostsim.py holds the builder to them.
usage: python scripts/table_hashes.py tests/golden/table_hashes.json   (only after a DELIBERATE change of the table layout)"""
import json, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from onsas_jl_b200 import meshgen as mg
from oracle import oracle as O
from tests import cases
from tests.hostsim import hostsim_py as H

def meshes():
    out = {}
    m, _ = cases.box_model(20, 10, 10, mat="svk")
    out["box_20x10x10"] = (m, None)
    out["box_20x10x10_owned_60pct"] = (m, int(0.6 * m.n_nodes))
    m0, _ = cases.box_model(11, 6, 5, mat="svk", jitter=0.15)
    rng = np.random.default_rng(11)
    n = m0.xyz.shape[0]
    perm = rng.permutation(n)
    xyz = np.empty_like(m0.xyz); xyz[perm] = m0.xyz
    tets = perm[m0.tets][rng.permutation(len(m0.tets))].astype(np.int32)
    out["random_numbering"] = (O.FlatModel(xyz=xyz, tets=tets, mat_kind=[0], mat_params=[[1.0, 1.0]], free_dofs=np.arange(3 * n)), None)
    lat = mg.truss_lattice(12, 6, 5, 2.0)
    out["truss_lattice_3d"] = (O.FlatModel(xyz=lat.xyz, trusses=lat.bars, truss_area=np.ones(lat.n_bars), truss_strain=1, mat_kind=[0], mat_params=[[0.0, 1.0]], free_dofs=np.arange(3 * lat.xyz.shape[0])), None)
    mc, _, _ = cases.clamped_truss(257)
    out["chain_1d"] = (mc, None)
    mv, _, _ = cases.von_mises_truss(1)
    out["von_mises"] = (mv, None)
    mb, _ = cases.box_model(6, 4, 3, mat="neo")
    nb = mb.xyz.shape[0]
    bars = np.stack([np.arange(nb - 1), np.arange(1, nb)], axis=1).astype(np.int32)
    out["mixed_tets_and_trusses"] = (O.FlatModel(xyz=mb.xyz, tets=mb.tets, trusses=bars, truss_area=np.ones(len(bars)), truss_strain=0, mat_kind=[1, 0], mat_params=[[1.0, 0.4], [0.0, 1.0]], tet_mat=np.zeros(len(mb.tets), np.int32), truss_mat=np.ones(len(bars), np.int32), free_dofs=np.arange(3 * nb)), None)
    hub = np.vstack([np.zeros((1, 3)), rng.standard_normal((400, 3))])
    hb = np.stack([np.zeros(400, np.int32), np.arange(1, 401, dtype=np.int32)], axis=1)
    out["hub_400_bars"] = (O.FlatModel(xyz=hub, trusses=hb, truss_area=np.ones(400), truss_strain=1, mat_kind=[0], mat_params=[[0.0, 1.0]], free_dofs=np.arange(3)), None)
    return out

if __name__ == "__main__":
  res = {}
  for name, (m, n_rows) in meshes().items():
    hs = H.HostSim(m, n_rows=n_rows)
    res[name] = {"hash": f"{hs.tables_hash():016x}", "stats": hs.stats()}
  json.dump(res, open(sys.argv[1], "w"), indent=1)
  print(json.dumps(res, indent=1))
