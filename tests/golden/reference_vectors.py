"""Literal known-answer vectors copied from the reference's own test files (data, with citations).
All paths relative to /root/reference; the reference prints them to 4-5 digits and checks at rtol 1e-3."""
import numpy as np

# test/entities/tetrahedrons.jl:11-38 -- nodes, SVK parameters, displacements
TET_NODES = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [2, 0, 1]], dtype=float)
TET_LAMBDA, TET_G = 0.5769, 0.3846
TET_U = np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2])
TET_VOLUME = 2 * 1 / 6  # :66, exact ==

# test/entities/tetrahedrons.jl:73-75 ("Values from ONSAS.m")
TET_F_INT = np.array([-0.9160, -1.3446, -1.5253, 0.3319, 0.7067, 0.4415, 0.3120, 0.5210, 0.9390, 0.2720, 0.1169, 0.1448])

# test/entities/tetrahedrons.jl:77-88
TET_K = np.array([
    [2.1635e+00, 7.8458e-01, 8.6150e-01, -9.4812e-01, -4.1633e-01, -2.8172e-01, -9.4668e-01, -2.1522e-01, -4.2675e-01, -2.6874e-01, -1.5304e-01, -1.5304e-01],
    [7.8458e-01, 3.1379e+00, 1.5089e+00, -3.7787e-01, -1.6917e+00, -6.0222e-01, -1.8797e-01, -1.2976e+00, -8.6102e-01, -2.1874e-01, -1.4855e-01, -4.5671e-02],
    [8.6150e-01, 1.5089e+00, 3.2917e+00, -3.0095e-01, -7.2401e-01, -1.1596e+00, -3.4181e-01, -7.3923e-01, -1.9835e+00, -2.1874e-01, -4.5671e-02, -1.4855e-01],
    [-9.4812e-01, -3.7787e-01, -3.0095e-01, 7.0582e-01, 2.4326e-01, 1.8557e-01, 1.4951e-01, 3.4454e-02, 8.8939e-02, 9.2785e-02, 1.0016e-01, 2.6441e-02],
    [-4.1633e-01, -1.6917e+00, -7.2401e-01, 2.4326e-01, 1.2571e+00, 3.0095e-01, 2.6441e-02, 3.6585e-01, 4.0143e-01, 1.4663e-01, 6.8747e-02, 2.1634e-02],
    [-2.8172e-01, -6.0222e-01, -1.1596e+00, 1.8557e-01, 3.0095e-01, 8.2120e-01, 6.0094e-02, 2.8444e-01, 2.9374e-01, 3.6056e-02, 1.6826e-02, 4.4710e-02],
    [-9.4668e-01, -1.8797e-01, -3.4181e-01, 1.4951e-01, 2.6441e-02, 6.0094e-02, 8.8031e-01, 1.5204e-01, 2.0812e-01, -8.3150e-02, 9.4948e-03, 7.3595e-02],
    [-2.1522e-01, -1.2976e+00, -7.3923e-01, 3.4454e-02, 3.6585e-01, 2.8444e-01, 1.5204e-01, 1.0165e+00, 4.7654e-01, 2.8725e-02, -8.4752e-02, -2.1754e-02],
    [-4.2675e-01, -8.6102e-01, -1.9835e+00, 8.8939e-02, 4.0143e-01, 2.9374e-01, 2.0812e-01, 4.7654e-01, 1.7697e+00, 1.2968e-01, -1.6946e-02, -7.9945e-02],
    [-2.6874e-01, -2.1874e-01, -2.1874e-01, 9.2785e-02, 1.4663e-01, 3.6056e-02, -8.3150e-02, 2.8725e-02, 1.2968e-01, 2.5910e-01, 4.3388e-02, 5.3003e-02],
    [-1.5304e-01, -1.4855e-01, -4.5671e-02, 1.0016e-01, 6.8747e-02, 1.6826e-02, 9.4948e-03, -8.4752e-02, -1.6946e-02, 4.3388e-02, 1.6456e-01, 4.5791e-02],
    [-1.5304e-01, -4.5671e-02, -1.4855e-01, 2.6441e-02, 2.1634e-02, 4.4710e-02, 7.3595e-02, -2.1754e-02, -7.9945e-02, 5.3003e-02, 4.5791e-02, 1.8379e-01]])

# test/entities/tetrahedrons.jl:90-92 (named E_e_test there, compared with the returned "strain" = C = F'F)
TET_C = np.array([[1.3675, 0.585, 1.02], [0.585, 1.87, 1.44], [1.02, 1.44, 3.28]])

# test/materials/materials.jl:77-84 -- soft hyperelastic parameters and the test strain
MAT_G, MAT_LAMBDA = 0.3846, 0.5769
MAT_K = MAT_LAMBDA + 2 * MAT_G / 3
MAT_E = np.array([[0.18375, 0.2925, 0.51], [0.2925, 0.435, 0.72], [0.51, 0.72, 1.14]])
# test/materials/materials.jl:112-114
SVK_S = np.array([[1.15596, 0.224991, 0.392292], [0.224991, 1.34922, 0.553824], [0.392292, 0.553824, 1.89151]])
# test/materials/materials.jl:116-121
SVK_D = np.array([[1.3461, 0.5769, 0.5769, 0, 0, 0], [0.5769, 1.3461, 0.5769, 0, 0, 0], [0.5769, 0.5769, 1.3461, 0, 0, 0],
                  [0, 0, 0, 0.3846, 0, 0], [0, 0, 0, 0, 0.3846, 0], [0, 0, 0, 0, 0, 0.3846]])

# test/structural_solvers/structural_solvers.jl:77-110 -- COO triplet order of the Assembler
COO_KE = np.array([[1.0, 2.0], [3.0, 4.0]])
COO_DOFS = [[1, 2], [2, 3]]
COO_I = [1, 2, 1, 2, 2, 3, 2, 3]
COO_J = [1, 1, 2, 2, 2, 2, 3, 3]
COO_KGLOB = np.array([[1.0, 2.0, 0.0], [3.0, 5.0, 2.0], [0.0, 3.0, 4.0]])

# Restatement-derived known answers recorded in SURVEY.md section 4 (numpy, direct solve; NOT Julia-run)
UNIAXIAL_EXTENSION_ITERS = [6, 5, 5, 4, 4, 4, 5, 5]
UNIAXIAL_COMPRESSION_ITERS = [5, 5, 5, 5, 5, 5, 5, 5, 4]
UNIAXIAL_COMPRESSION_ALPHA, UNIAXIAL_COMPRESSION_BETA = 0.47494632337, 1.20848436103
VON_MISES_UK = {"roteng": -0.2418999551, "green": -0.3041777184}
VON_MISES_ITERS = {"roteng": 4, "green": 5}
