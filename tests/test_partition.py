"""The library's partitioner (csrc/partition.cpp, behind onsas_part_* and inside onsas_finalize_mesh of a multi-device context)
against the numpy restatement in onsas.jl_b200/partition.py, on the CPU: same renumbering, same owned ranges, same local
meshes, same halo plan, and the offsets of a rank's values inside each neighbour's halo."""
import numpy as np
import pytest

from onsas_jl_b200 import meshgen as mg
from onsas_jl_b200 import partition as pt


def _numpy_parts(xyz, n_ranks, tets=None, trusses=None, free=None, truss_area=None):
    order, ranges = pt.rcb_order(xyz, n_ranks)
    out = pt.renumber(order, xyz, tets, trusses)
    xyz2, tets2, trusses2, inv = out
    d = xyz.shape[1]
    gfree = None if free is None else np.sort(inv[free // d] * d + free % d)
    parts = [pt.build_local_part(r, ranges, xyz2, tets=tets2, trusses=trusses2, free_dofs=gfree, truss_area=truss_area)
             for r in range(n_ranks)]
    return order, ranges, parts


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4, 8])
def test_native_partition_matches_numpy_restatement_tets(ob, n_ranks):
    mesh = mg.box_tet_mesh(9, 5, 4, 2.0, 1.0, 1.0)
    rng = np.random.default_rng(5)
    perm = rng.permutation(mesh.n_nodes)                       # an arbitrary caller numbering
    xyz = np.empty_like(mesh.xyz)
    xyz[perm] = mesh.xyz
    tets = perm[mesh.tets].astype(np.int32)
    free = np.sort(rng.choice(mesh.n_nodes * 3, size=mesh.n_nodes * 2, replace=False)).astype(np.int64)
    order, ranges, parts = _numpy_parts(xyz, n_ranks, tets=tets, free=free)
    P = ob.NativePartition(xyz, n_ranks, tets=tets, free_dofs=free)
    seen = np.zeros(mesh.n_nodes, int)
    for r in range(n_ranks):
        ref, sz = parts[r], P.sizes(r)
        assert (sz["n_local"], sz["n_owned"], sz["n_tets"], sz["n_free"]) == (ref.n_local, ref.n_owned, len(ref.tets), len(ref.free_dofs))
        l2g = P.local_to_global(r)
        np.testing.assert_array_equal(l2g, order[ref.local_to_global])          # caller ids, owned first, halo by owner
        seen[l2g[:sz["n_owned"]]] += 1
        np.testing.assert_array_equal(P.local_elements(r), ref.tet_global)
        hp = P.halo_plan(r)
        np.testing.assert_array_equal(hp["nbr_rank"], ref.nbr_rank)
        np.testing.assert_array_equal(hp["send_ptr"], ref.send_ptr)
        np.testing.assert_array_equal(hp["send_nodes"], ref.send_nodes)
        np.testing.assert_array_equal(hp["recv_ptr"], ref.recv_ptr)
        for k, nb in enumerate(ref.nbr_rank):                                    # where my values start inside nb's halo
            j = list(parts[nb].nbr_rank).index(r)
            assert hp["remote_halo_off"][k] == parts[nb].recv_ptr[j]
            # and what I send is exactly what nb expects there, in the same order
            mine = l2g[hp["send_nodes"][hp["send_ptr"][k]:hp["send_ptr"][k + 1]]]
            theirs = P.local_to_global(nb)[parts[nb].n_owned + parts[nb].recv_ptr[j]: parts[nb].n_owned + parts[nb].recv_ptr[j + 1]]
            np.testing.assert_array_equal(mine, theirs)
    assert (seen == 1).all()                                                      # every node owned exactly once
    P.close()


def test_native_partition_trusses_and_errors(ob):
    lat = mg.truss_lattice(6, 3, 3, 2.0)
    order, ranges, parts = _numpy_parts(lat.xyz, 4, trusses=lat.bars, truss_area=np.arange(lat.n_bars, dtype=float))
    P = ob.NativePartition(lat.xyz, 4, trusses=lat.bars, truss_area=np.arange(lat.n_bars, dtype=float))
    for r in range(4):
        np.testing.assert_array_equal(P.local_elements(r, ob.FAMILY_TRUSS), parts[r].truss_global)
        np.testing.assert_array_equal(P.local_to_global(r), order[parts[r].local_to_global])
    with pytest.raises(ob.OnsasError):
        ob.NativePartition(lat.xyz, 17, trusses=lat.bars)                        # at most 16 ranks
    bad = lat.bars.copy()
    bad[0, 0] = lat.n_nodes
    with pytest.raises(ob.OnsasError):
        ob.NativePartition(lat.xyz, 2, trusses=bad)


def test_z_curve_reordering_is_a_local_permutation(ob):
    """reorder = 1 (ONSAS_OPT_REORDER): inside every part the nodes follow the Z-curve of their coordinates -- still a
    permutation with the same owned SETS as reorder = 0, and a randomly numbered mesh gets back the locality of a structured
    one: the nodes an 8-row slice touches stay a small neighbourhood."""
    mesh = mg.box_tet_mesh(16, 16, 16, 1.0, 1.0, 1.0)
    rng = np.random.default_rng(1)
    perm = rng.permutation(mesh.n_nodes)
    xyz = np.empty_like(mesh.xyz)
    xyz[perm] = mesh.xyz
    tets = perm[mesh.tets].astype(np.int32)
    for n_ranks in (1, 4):
        P0 = ob.NativePartition(xyz, n_ranks, tets=tets)
        P1 = ob.NativePartition(xyz, n_ranks, tets=tets, reorder=1)
        for r in range(n_ranks):
            s0, s1 = P0.sizes(r), P1.sizes(r)
            assert s0 == s1
            a, b = P0.local_to_global(r), P1.local_to_global(r)
            no = s0["n_owned"]
            assert np.array_equal(np.sort(a[:no]), np.sort(b[:no])) and np.array_equal(np.sort(a[no:]), np.sort(b[no:]))
            np.testing.assert_array_equal(P0.local_elements(r), P1.local_elements(r))   # element order is untouched
            # locality: mean distance between consecutive owned nodes
            d0 = np.linalg.norm(np.diff(xyz[a[:no]], axis=0), axis=1).mean()
            d1 = np.linalg.norm(np.diff(xyz[b[:no]], axis=0), axis=1).mean()
            assert d1 < 0.25 * d0 and d1 < 3.0 / 16                                     # random numbering: ~0.6; Z-curve: ~1.5 cells
    with pytest.raises(ob.OnsasError):
        ob.NativePartition(xyz, 2, tets=tets, reorder=7)
