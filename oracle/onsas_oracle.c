/*
 * onsas_oracle.c -- CPU restatement of ONSAS.jl's Newton-Raphson hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's shared object, and there only as the checker or the timed CPU baseline.
 * The product path is onsas.jl_b200/csrc (CUDA, sm_100a) and never links this.
 *
 * Parity status: PINNED for the element/material/assembly arithmetic -- checked in
 * tests/test_oracle_golden.py against every literal vector the reference's tests hold
 * for this path (test/entities/tetrahedrons.jl:73-91, test/materials/materials.jl:112-121,
 * test/entities/trusses.jl:52-61,109-117, test/structural_solvers/structural_solvers.jl:77-126)
 * and against the analytic end states of the reference's examples.  The reference
 * itself (Julia) cannot run in this image, so the third-party linear solve
 * (IterativeSolvers.jl 0.9.4 `cg`, LinearSolve.jl 2.39.1 wrapper) is restated from its
 * published algorithm below and is "parity unpinned" at unit level -- see DESIGN.md.
 *
 * All file:line citations are relative to /root/reference/src unless stated otherwise.
 * Conventions: FP64; 3x3 tensors and K_e are column-major (Julia memory
 * order); Voigt order 11,22,33,23,13,12 (Utils.jl:42); element dofs node-major xyz
 * (Entities/Entities.jl:156-171); global dof = dim*node + component (Meshes/Meshes.jl:85-98).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_ERR_NEG_VOLUME 1 /* Tetrahedrons.jl:134-138 ArgumentError */
#define ORC_ERR_BAD_ARG 2
#define ORC_ERR_PATTERN 3

enum { MAT_SVK = 0, MAT_NEOHOOKEAN = 1, MAT_ISOLINEAR = 2 };
enum { STRAIN_ROTENG = 0, STRAIN_GREEN = 1 };

/* column-major 3x3 accessor */
#define M3(A, i, j) ((A)[(i) + 3 * (j)])

static const int VOIGT_I[6] = {0, 1, 2, 1, 0, 0}; /* Utils.jl:42 INDEXES_TO_VOIGT */
static const int VOIGT_J[6] = {0, 1, 2, 2, 2, 1};

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm of bench.py asks for all host cores */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static double det3(const double *A) {
    return M3(A, 0, 0) * (M3(A, 1, 1) * M3(A, 2, 2) - M3(A, 1, 2) * M3(A, 2, 1)) -
           M3(A, 0, 1) * (M3(A, 1, 0) * M3(A, 2, 2) - M3(A, 1, 2) * M3(A, 2, 0)) +
           M3(A, 0, 2) * (M3(A, 1, 0) * M3(A, 2, 1) - M3(A, 1, 1) * M3(A, 2, 0));
}

static void inv3(const double *A, double *Ai, double det) {
    double id = 1.0 / det;
    M3(Ai, 0, 0) = (M3(A, 1, 1) * M3(A, 2, 2) - M3(A, 1, 2) * M3(A, 2, 1)) * id;
    M3(Ai, 0, 1) = (M3(A, 0, 2) * M3(A, 2, 1) - M3(A, 0, 1) * M3(A, 2, 2)) * id;
    M3(Ai, 0, 2) = (M3(A, 0, 1) * M3(A, 1, 2) - M3(A, 0, 2) * M3(A, 1, 1)) * id;
    M3(Ai, 1, 0) = (M3(A, 1, 2) * M3(A, 2, 0) - M3(A, 1, 0) * M3(A, 2, 2)) * id;
    M3(Ai, 1, 1) = (M3(A, 0, 0) * M3(A, 2, 2) - M3(A, 0, 2) * M3(A, 2, 0)) * id;
    M3(Ai, 1, 2) = (M3(A, 0, 2) * M3(A, 1, 0) - M3(A, 0, 0) * M3(A, 1, 2)) * id;
    M3(Ai, 2, 0) = (M3(A, 1, 0) * M3(A, 2, 1) - M3(A, 1, 1) * M3(A, 2, 0)) * id;
    M3(Ai, 2, 1) = (M3(A, 0, 1) * M3(A, 2, 0) - M3(A, 0, 0) * M3(A, 2, 1)) * id;
    M3(Ai, 2, 2) = (M3(A, 0, 0) * M3(A, 1, 1) - M3(A, 0, 1) * M3(A, 1, 0)) * id;
}

/* ------------------------------------------------------------------ materials */

/* Lame parameters from (E, nu): SVKMaterial.jl:50-54, IsotropicLinearElasticMaterial.jl:70-76 */
void orc_lame_from_E_nu(double E, double nu, double *lambda, double *G) {
    *lambda = E * nu / ((1 + nu) * (1 - 2 * nu));
    *G = E / (2 * (1 + nu));
}

/* elasticity modulus of SVK(lambda, G): SVKMaterial.jl:74-77 */
double orc_svk_elasticity_modulus(double lambda, double G) {
    return G * (3 * lambda + 2 * G) / (lambda + G);
}

/* SVKMaterial.jl:89-100: S = lambda tr(E) I + 2 G E;  D[1:3,1:3] = lambda*ones + 2G*I; D[4:6,4:6] = G*I.
 * D is 6x6 column-major.  (The reference leaves the other entries of its shared cache
 * untouched; here they are zero -- SURVEY.md section 7, last bullet.) */
void orc_svk_stress(double lambda, double G, const double *E, double *S, double *D) {
    double tr = M3(E, 0, 0) + M3(E, 1, 1) + M3(E, 2, 2);
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) M3(S, i, j) = lambda * tr * (i == j ? 1.0 : 0.0) + 2 * G * M3(E, i, j);
    memset(D, 0, 36 * sizeof(double));
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) D[i + 6 * j] = lambda + (i == j ? 2 * G : 0.0);
    for (int i = 3; i < 6; ++i) D[i + 6 * i] = G;
}

/* IsotropicLinearElasticMaterial.jl:80-92 (lambda, G from (E, nu) :70-76) */
void orc_isolinear_stress(double Emod, double nu, const double *eps, double *sigma, double *D) {
    double G = Emod / (2 * (1 + nu)); /* shear_modulus :56-59 */
    double lambda = Emod * nu / ((1 + nu) * (1 - 2 * nu));
    orc_svk_stress(lambda, G, eps, sigma, D);
}

/* NeoHookeanMaterial.jl:94-102 (_S_analytic): C = Symmetric(2E + I) (upper triangle),
 * S = G (I - C^-1) + K J (J-1) C^-1, J = sqrt(det C).
 * Tangent: the reference differentiates S with ForwardDiff (NeoHookeanMaterial.jl:115-129),
 * a third-party AD package not vendored under /root/reference (ForwardDiff 0.10.38 per
 * docs/Manifest.toml).  AD of a closed-form function returns its exact derivative, so the
 * oracle uses the closed form  DD = 2 (G - K J (J-1)) C^-1 (.) C^-1 + K J (2J-1) C^-1 (x) C^-1
 * where (A (.) A)_ijkl = (A_ik A_jl + A_il A_jk)/2; with the reference's `voigt(grad, 0.5)`
 * column scaling this is D[voigt(ij), voigt(kl)] = DD_ijkl.
 * tests/test_oracle_golden.py cross-checks it by central differences of S. */
void orc_neohookean_stress(double K, double G, const double *E, double *S, double *D) {
    double C[9], Ci[9];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i <= j; ++i) {
            double c = 2 * M3(E, i, j) + (i == j ? 1.0 : 0.0);
            M3(C, i, j) = c;
            M3(C, j, i) = c;
        }
    double detC = det3(C);
    inv3(C, Ci, detC);
    double J = sqrt(detC);
    double kj = K * J * (J - 1);
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) M3(S, i, j) = G * ((i == j ? 1.0 : 0.0) - M3(Ci, i, j)) + kj * M3(Ci, i, j);
    /* _S_analytic! wraps in Symmetric(): upper triangle mirrored (exactly symmetric here) */
    double ca = 2 * (G - kj), cb = K * J * (2 * J - 1);
    for (int b = 0; b < 6; ++b)
        for (int a = 0; a < 6; ++a) {
            int i = VOIGT_I[a], j = VOIGT_J[a], k = VOIGT_I[b], l = VOIGT_J[b];
            D[a + 6 * b] = ca * 0.5 * (M3(Ci, i, k) * M3(Ci, j, l) + M3(Ci, i, l) * M3(Ci, j, k)) +
                           cb * M3(Ci, i, j) * M3(Ci, k, l);
        }
}

/* strain energies, used only to cross-check S = dPsi/dE in the tests
 * (SVKMaterial.jl:57-60, NeoHookeanMaterial.jl:58-66) */
double orc_strain_energy(int kind, double p0, double p1, const double *E) {
    if (kind == MAT_SVK) {
        double tr = M3(E, 0, 0) + M3(E, 1, 1) + M3(E, 2, 2), tr2 = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) tr2 += M3(E, i, j) * M3(E, j, i);
        return p0 / 2 * tr * tr + p1 * tr2;
    }
    double C[9];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i <= j; ++i) {
            double c = 2 * M3(E, i, j) + (i == j ? 1.0 : 0.0);
            M3(C, i, j) = c;
            M3(C, j, i) = c;
        }
    double J = sqrt(det3(C));
    double I1 = M3(C, 0, 0) + M3(C, 1, 1) + M3(C, 2, 2);
    return p1 / 2 * (I1 - 2 * log(J)) + p0 / 2 * (J - 1) * (J - 1);
}

/* ------------------------------------------------------------------ tetrahedron */

/* Tetrahedrons.jl:269-271 */
static const double DXDZ[3][4] = {{1.0, -1.0, 0.0, 0.0}, {0.0, -1.0, 0.0, 1.0}, {0.0, -1.0, 1.0, 0.0}};

/* Tetrahedrons.jl:140-152 (_B_mat!): B is 6x12 column-major */
static void tet_B(const double *funder /*3x4*/, const double *F, double *B) {
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) {
            int col = 3 * k + c;
            /* rows 1-3: diagm(deriv[:,r])*F' block-> B[r, 3k+c] = deriv[r,k] * F[c,r] */
            for (int r = 0; r < 3; ++r) B[r + 6 * col] = funder[r + 3 * k] * M3(F, c, r);
            B[3 + 6 * col] = funder[1 + 3 * k] * M3(F, c, 2) + funder[2 + 3 * k] * M3(F, c, 1);
            B[4 + 6 * col] = funder[0 + 3 * k] * M3(F, c, 2) + funder[2 + 3 * k] * M3(F, c, 0);
            B[5 + 6 * col] = funder[0 + 3 * k] * M3(F, c, 1) + funder[1 + 3 * k] * M3(F, c, 0);
        }
}

/* Volume of a tet: Tetrahedrons.jl:114-120,129-138.  X is 3x4 column-major. */
double orc_tet_volume(const double *X) {
    double J[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += X[i + 3 * k] * DXDZ[j][k];
            M3(J, i, j) = s;
        }
    return det3(J) / 6.0;
}

/* internal_forces for a Tetrahedron: hyperelastic Tetrahedrons.jl:186-224, linear :240-266.
 * kind/p0/p1: SVK(lambda,G) | NeoHookean(K,G) | IsotropicLinearElastic(E,nu).
 * X 3x4 col-major, u 12 (node-major), out: f[12], Ke[144] col-major, sig[9], eps[9] col-major.
 * "sig" is P = F S and "eps" is C = F'F for hyperelastic (:218-221); Cauchy sigma / small eps for linear (:265). */
int orc_tet_internal_forces(int kind, double p0, double p1, const double *X, const double *u, double *f, double *Ke,
                            double *sig, double *eps) {
    double J[9], Ji[9], funder[12], H[9], F[9], E[9], S[9], D[36], B[72];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += X[i + 3 * k] * DXDZ[j][k];
            M3(J, i, j) = s; /* J = X * dXdz' (:129-131) */
        }
    double detJ = det3(J);
    double vol = detJ / 6.0;
    if (!(vol > 0)) return ORC_ERR_NEG_VOLUME;
    inv3(J, Ji, detJ);
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 4; ++k) {
            double s = 0;
            for (int j = 0; j < 3; ++j) s += M3(Ji, j, i) * DXDZ[j][k];
            funder[i + 3 * k] = s; /* funder = inv(J)' * dXdz (:197) */
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += u[i + 3 * k] * funder[j + 3 * k];
            M3(H, i, j) = s; /* H = U * funder' (:198) */
        }

    if (kind == MAT_ISOLINEAR) {
        double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) M3(eps, i, j) = 0.5 * (M3(H, i, j) + M3(H, j, i)); /* :250 */
        tet_B(funder, I3, B);                                                             /* :251-252 */
        orc_isolinear_stress(p0, p1, eps, sig, D);                                        /* :256 */
    } else {
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) M3(F, i, j) = M3(H, i, j) + (i == j ? 1.0 : 0.0); /* :199 */
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i <= j; ++i) {
                double hh = 0;
                for (int k = 0; k < 3; ++k) hh += M3(H, k, i) * M3(H, k, j);
                double e = 0.5 * (M3(H, i, j) + M3(H, j, i) + hh); /* :201, Symmetric = upper triangle */
                M3(E, i, j) = e;
                M3(E, j, i) = e;
            }
        tet_B(funder, F, B); /* :202 */
        if (kind == MAT_SVK)
            orc_svk_stress(p0, p1, E, S, D); /* :205 */
        else if (kind == MAT_NEOHOOKEAN)
            orc_neohookean_stress(p0, p1, E, S, D);
        else
            return ORC_ERR_BAD_ARG;
    }

    /* Km = Symmetric(B' * D * B * vol)  (:210 / :259) -- upper triangle mirrored */
    double DB[72];
    for (int c = 0; c < 12; ++c)
        for (int r = 0; r < 6; ++r) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += D[r + 6 * k] * B[k + 6 * c];
            DB[r + 6 * c] = s;
        }
    for (int c = 0; c < 12; ++c)
        for (int r = 0; r <= c; ++r) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += B[k + 6 * r] * DB[k + 6 * c];
            s *= vol;
            Ke[r + 12 * c] = s;
            Ke[c + 12 * r] = s;
        }

    if (kind == MAT_ISOLINEAR) {
        for (int r = 0; r < 12; ++r) { /* fint = Ks * u_e (:261) */
            double s = 0;
            for (int c = 0; c < 12; ++c) s += Ke[r + 12 * c] * u[c];
            f[r] = s;
        }
        return ORC_OK;
    }

    /* fint = B' * voigt(S) * vol (:206-207) */
    double Sv[6];
    for (int a = 0; a < 6; ++a) Sv[a] = M3(S, VOIGT_I[a], VOIGT_J[a]);
    for (int c = 0; c < 12; ++c) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += B[k + 6 * c] * Sv[k];
        f[c] = s * vol;
    }
    /* geometric stiffness: kron(funder' S funder vol, I3) (:161-182, :213-215) */
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            double s = 0;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) s += funder[i + 3 * a] * M3(S, i, j) * funder[j + 3 * b];
            s *= vol;
            for (int c = 0; c < 3; ++c) Ke[(3 * a + c) + 12 * (3 * b + c)] += s;
        }
    /* P = F*S (:218), eps = Symmetric(F'F) (:221) */
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += M3(F, i, k) * M3(S, k, j);
            M3(sig, i, j) = s;
        }
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i <= j; ++i) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += M3(F, k, i) * M3(F, k, j);
            M3(eps, i, j) = s;
            M3(eps, j, i) = s;
        }
    return ORC_OK;
}

/* ------------------------------------------------------------------ truss */

/* internal_forces for a Truss: rotated-engineering Trusses.jl:126-155, Green :159-184.
 * Emod = elasticity_modulus(material), A = area(cross_section).  X, u: 2*dim, node-major.
 * out f[2dim], Ke[(2dim)^2] col-major, sig[9]/eps[9] with only [1,1] set (:148-152). */
int orc_truss_internal_forces(int strain_model, int dim, double Emod, double A, const double *X, const double *u,
                              double *f, double *Ke, double *sig, double *eps) {
    if (dim < 1 || dim > 3) return ORC_ERR_BAD_ARG;
    int n = 2 * dim;
    double dref[3] = {0, 0, 0}, ddef[3] = {0, 0, 0};
    double l_ref2 = 0, l_def2 = 0;
    for (int c = 0; c < dim; ++c) { /* _lengths :232-237: Bdif * X = X2 - X1 */
        dref[c] = X[dim + c] - X[c];
        ddef[c] = (X[dim + c] + u[dim + c]) - (X[c] + u[c]); /* _X_rows :214-218 */
        l_ref2 += dref[c] * dref[c];
        l_def2 += ddef[c] * ddef[c];
    }
    double l_ref = sqrt(l_ref2), l_def = sqrt(l_def2);
    memset(sig, 0, 9 * sizeof(double));
    memset(eps, 0, 9 * sizeof(double));
    if (strain_model == STRAIN_ROTENG) {
        double TT[6]; /* TTcl = Bdif' * e1_def = [-e1; e1] (:136-137) */
        for (int c = 0; c < dim; ++c) {
            double e1 = ddef[c] / l_def;
            TT[c] = -e1;
            TT[dim + c] = e1;
        }
        double e = (l_def * l_def - l_ref * l_ref) / (l_ref * (l_ref + l_def)); /* _strain :187-189 */
        double S11 = Emod * e;                                                    /* :141 */
        for (int i = 0; i < n; ++i) f[i] = A * S11 * TT[i];                       /* :142 */
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                /* Ge = Bdif'Bdif = [I -I; -I I] (:207-211) */
                double ge = 0;
                if (i % dim == j % dim) ge = ((i < dim) == (j < dim)) ? 1.0 : -1.0;
                double km = Emod * A / l_ref * (TT[i] * TT[j]);        /* :144 */
                double kg = S11 * A / l_def * (ge - TT[i] * TT[j]);    /* :145 */
                Ke[i + n * j] = km + kg;
            }
        M3(sig, 0, 0) = S11 * l_def / l_ref; /* :151 */
        M3(eps, 0, 0) = e;
    } else if (strain_model == STRAIN_GREEN) {
        double bsum[6]; /* b_ref + b_def = (X_ref + u)' * Ge / l_ref^2 (:221-229, :173) */
        for (int c = 0; c < dim; ++c) {
            double bref = -dref[c] / (l_ref * l_ref);           /* (X' Ge)[c] = X1 - X2 */
            double bdef = -(u[dim + c] - u[c]) / (l_ref * l_ref);
            bsum[c] = bref + bdef;
            bsum[dim + c] = -bref + -bdef;
        }
        double e = (l_def * l_def - l_ref * l_ref) / (2 * l_ref * l_ref); /* _strain :192-194 */
        double S11 = Emod * e;
        for (int i = 0; i < n; ++i) f[i] = A * S11 * l_ref * bsum[i]; /* :174 */
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                double ge = 0;
                if (i % dim == j % dim) ge = ((i < dim) == (j < dim)) ? 1.0 : -1.0;
                Ke[i + n * j] = S11 * A / l_ref * ge + Emod * A * l_ref * (bsum[i] * bsum[j]); /* :176 */
            }
        M3(sig, 0, 0) = S11 * l_def / l_ref; /* :181 */
        M3(eps, 0, 0) = e;
    } else
        return ORC_ERR_BAD_ARG;
    return ORC_OK;
}

/* ------------------------------------------------------------------ batched element evaluation */

/* One call per element family.  xyz: dim x n_nodes col-major (node-interleaved), U: n_dofs,
 * conn: npe x n_elem (0-based node ids), mat_id per element indexing (kind[], params[2*m]).
 * Un-assembled outputs, element-major: f (ndof_e each), K ((ndof_e)^2 col-major each), sig/eps 9 each. */
int orc_eval_tets(int64_t n_elem, const int32_t *conn, const int32_t *mat_id, const int32_t *kind, const double *params,
                  const double *xyz, const double *U, double *f, double *K, double *sig, double *eps) {
    int status = ORC_OK;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n_elem; ++e) {
        double X[12], u[12];
        for (int k = 0; k < 4; ++k) {
            int64_t nd = conn[4 * e + k];
            for (int c = 0; c < 3; ++c) {
                X[c + 3 * k] = xyz[3 * nd + c];
                u[c + 3 * k] = U[3 * nd + c];
            }
        }
        int m = mat_id ? mat_id[e] : 0;
        int st = orc_tet_internal_forces(kind[m], params[2 * m], params[2 * m + 1], X, u, f + 12 * e, K + 144 * e,
                                         sig + 9 * e, eps + 9 * e);
        if (st != ORC_OK) {
#pragma omp critical
            status = st;
        }
    }
    return status;
}

int orc_eval_trusses(int64_t n_elem, int dim, int strain_model, const int32_t *conn, const int32_t *mat_id,
                     const int32_t *kind, const double *params, const double *area, const double *xyz, const double *U,
                     double *f, double *K, double *sig, double *eps) {
    int n = 2 * dim;
    int status = ORC_OK;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n_elem; ++e) {
        double X[6], u[6];
        for (int k = 0; k < 2; ++k) {
            int64_t nd = conn[2 * e + k];
            for (int c = 0; c < dim; ++c) {
                X[c + dim * k] = xyz[dim * nd + c];
                u[c + dim * k] = U[dim * nd + c];
            }
        }
        int m = mat_id ? mat_id[e] : 0;
        if (kind[m] == MAT_ISOLINEAR) { /* Trusses.jl dispatches on AbstractHyperElasticMaterial only */
#pragma omp critical
            status = ORC_ERR_BAD_ARG;
            continue;
        }
        /* elasticity_modulus(m): SVKMaterial.jl:74-77 / NeoHookeanMaterial.jl:84-87 (via lame :68-72) */
        double lam, G;
        if (kind[m] == MAT_SVK) {
            lam = params[2 * m];
            G = params[2 * m + 1];
        } else {
            G = params[2 * m + 1];
            lam = params[2 * m] - 2 * G / 3;
        }
        double Emod = orc_svk_elasticity_modulus(lam, G);
        int st = orc_truss_internal_forces(strain_model, dim, Emod, area[e], X, u, f + n * e, K + n * n * e,
                                           sig + 9 * e, eps + 9 * e);
        if (st != ORC_OK) {
#pragma omp critical
            status = st;
        }
    }
    return status;
}

/* ------------------------------------------------------------------ assembly (reference order) */

/* The reference appends, per element, vec(K_e) column-major with I = row dofs repeated per
 * column and J = the column dof (Assemblers.jl:52-67), then inserts every triplet with
 * K[I,J] += V (Assemblers.jl:84-88) into a SparseMatrixCSC (sorted rows within a column,
 * one binary search per access).  orc_pattern_* builds that fixed sparsity once (what the
 * first end_assemble! leaves behind); orc_assemble_* then repeats the reference's steady
 * state: reset (StaticAnalyses.jl:125-132), serial element loop in the given order
 * (StaticAnalyses.jl:105-118), triplet insertion by binary search.
 *
 * The matrix is stored CSR with sorted columns.  K is structurally symmetric, so CSR of K
 * has the same index arrays as Julia's CSC; inserting (I,J,V) at row I / column J keeps the
 * reference's summation order per entry. */

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* Build CSR pattern of an (n_dofs x n_dofs) matrix from element dof lists.
 * edofs: nde x n_elem (global dofs, element-major). Two-pass API: call with col == NULL to
 * get nnz in *nnz_out and rowptr filled, then with col allocated. */
int orc_pattern_build(int64_t n_dofs, int64_t n_elem, int nde, const int64_t *edofs, int64_t *rowptr, int32_t *col,
                      int64_t *nnz_out) {
    /* count (with duplicates) per row */
    int64_t *cnt = (int64_t *)calloc((size_t)n_dofs + 1, sizeof(int64_t));
    if (!cnt) return ORC_ERR_PATTERN;
    for (int64_t e = 0; e < n_elem; ++e)
        for (int a = 0; a < nde; ++a) cnt[edofs[e * nde + a] + 1] += nde;
    for (int64_t i = 0; i < n_dofs; ++i) cnt[i + 1] += cnt[i];
    int64_t total = cnt[n_dofs];
    int64_t *tmp = (int64_t *)malloc((size_t)(total > 0 ? total : 1) * sizeof(int64_t));
    int64_t *fill = (int64_t *)malloc((size_t)(n_dofs + 1) * sizeof(int64_t));
    if (!tmp || !fill) {
        free(cnt);
        free(tmp);
        free(fill);
        return ORC_ERR_PATTERN;
    }
    memcpy(fill, cnt, (size_t)(n_dofs + 1) * sizeof(int64_t));
    for (int64_t e = 0; e < n_elem; ++e)
        for (int a = 0; a < nde; ++a) {
            int64_t r = edofs[e * nde + a];
            for (int b = 0; b < nde; ++b) tmp[fill[r]++] = edofs[e * nde + b];
        }
    int64_t nnz = 0;
    rowptr[0] = 0;
    for (int64_t i = 0; i < n_dofs; ++i) {
        int64_t s = cnt[i], t = cnt[i + 1];
        qsort(tmp + s, (size_t)(t - s), sizeof(int64_t), cmp_i64);
        int64_t last = -1;
        for (int64_t k = s; k < t; ++k)
            if (tmp[k] != last) {
                last = tmp[k];
                if (col) col[nnz] = (int32_t)last;
                ++nnz;
            }
        rowptr[i + 1] = nnz;
    }
    *nnz_out = nnz;
    free(cnt);
    free(tmp);
    free(fill);
    return ORC_OK;
}

static inline int64_t csr_find(const int64_t *rowptr, const int32_t *col, int64_t r, int64_t c) {
    int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        if (col[mid] < c)
            lo = mid + 1;
        else if (col[mid] > c)
            hi = mid - 1;
        else
            return mid;
    }
    return -1;
}

/* Serial, reference-order assembly of one element family into F_int and CSR values.
 * family: 0 = tets (dim 3), 1 = trusses.  `reset` zeroes F_int / val / nothing else first
 * (reset_assemble!).  COO buffers I,J,V (capacity n_elem*nde*nde) are filled exactly like the
 * reference's Assembler and then inserted; pass NULL to let the function allocate them. */
int orc_assemble_family(int family, int64_t n_elem, int dim, int strain_model, const int32_t *conn,
                        const int32_t *mat_id, const int32_t *kind, const double *params, const double *area,
                        const double *xyz, const double *U, int64_t n_dofs, const int64_t *rowptr, const int32_t *col,
                        double *val, double *F_int, double *sig, double *eps, int reset) {
    int npe = family == 0 ? 4 : 2;
    int nde = npe * dim;
    if (family == 0 && dim != 3) return ORC_ERR_BAD_ARG;
    if (reset) {
        memset(F_int, 0, (size_t)n_dofs * sizeof(double));
        memset(val, 0, (size_t)rowptr[n_dofs] * sizeof(double));
    }
    size_t cap = (size_t)n_elem * nde * nde;
    int64_t *I = (int64_t *)malloc((cap ? cap : 1) * sizeof(int64_t));
    int64_t *Jc = (int64_t *)malloc((cap ? cap : 1) * sizeof(int64_t));
    double *V = (double *)malloc((cap ? cap : 1) * sizeof(double));
    if (!I || !Jc || !V) {
        free(I);
        free(Jc);
        free(V);
        return ORC_ERR_PATTERN;
    }
    int status = ORC_OK;
    size_t nt = 0;
    for (int64_t e = 0; e < n_elem && status == ORC_OK; ++e) {
        double X[12], u[12], f[12], Ke[144];
        int64_t dofs[12];
        for (int k = 0; k < npe; ++k) {
            int64_t nd = conn[npe * e + k];
            for (int c = 0; c < dim; ++c) {
                X[c + dim * k] = xyz[dim * nd + c];
                dofs[c + dim * k] = dim * nd + c; /* local_dofs: node-major (Entities.jl:156-171) */
                u[c + dim * k] = U[dofs[c + dim * k]];
            }
        }
        int m = mat_id ? mat_id[e] : 0;
        if (family == 0)
            status = orc_tet_internal_forces(kind[m], params[2 * m], params[2 * m + 1], X, u, f, Ke, sig + 9 * e,
                                             eps + 9 * e);
        else {
            double lam, G;
            if (kind[m] == MAT_SVK) {
                lam = params[2 * m];
                G = params[2 * m + 1];
            } else if (kind[m] == MAT_NEOHOOKEAN) {
                G = params[2 * m + 1];
                lam = params[2 * m] - 2 * G / 3;
            } else {
                status = ORC_ERR_BAD_ARG;
                break;
            }
            status = orc_truss_internal_forces(strain_model, dim, orc_svk_elasticity_modulus(lam, G), area[e], X, u, f,
                                               Ke, sig + 9 * e, eps + 9 * e);
        }
        if (status != ORC_OK) break;
        for (int a = 0; a < nde; ++a) F_int[dofs[a]] += f[a]; /* StructuralAnalyses.jl:105-107 */
        for (int c = 0; c < nde; ++c)                         /* Assemblers.jl:52-67 */
            for (int r = 0; r < nde; ++r) {
                I[nt] = dofs[r];
                Jc[nt] = dofs[c];
                V[nt] = Ke[r + nde * c];
                ++nt;
            }
    }
    if (status == ORC_OK)
        for (size_t k = 0; k < nt; ++k) { /* Assemblers.jl:84-88 */
            int64_t p = csr_find(rowptr, col, I[k], Jc[k]);
            if (p < 0) {
                status = ORC_ERR_PATTERN;
                break;
            }
            val[p] += V[k];
        }
    free(I);
    free(Jc);
    free(V);
    return status;
}

/* Multi-threaded variant for the "--impl reference, all host threads" CPU arm: OpenMP over
 * elements for evaluation (un-assembled scratch), then a row-parallel gather over a
 * node->element adjacency so the per-entry summation order (ascending element id) and hence
 * the result is identical to orc_assemble_family for a single family.  Tets only. */
int orc_assemble_tets_mt(int64_t n_elem, const int32_t *conn, const int32_t *mat_id, const int32_t *kind,
                         const double *params, const double *xyz, const double *U, int64_t n_nodes,
                         const int64_t *adj_ptr, const int32_t *adj /* elem*4+a, ascending */, const int64_t *rowptr,
                         const int32_t *col, double *val, double *F_int, double *sig, double *eps, double *scratchK,
                         double *scratchf) {
    int status = orc_eval_tets(n_elem, conn, mat_id, kind, params, xyz, U, scratchf, scratchK, sig, eps);
    if (status != ORC_OK) return status;
#pragma omp parallel for schedule(static)
    for (int64_t nd = 0; nd < n_nodes; ++nd) {
        for (int c = 0; c < 3; ++c) {
            int64_t r = 3 * nd + c;
            for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) val[p] = 0.0;
            F_int[r] = 0.0;
        }
        for (int64_t q = adj_ptr[nd]; q < adj_ptr[nd + 1]; ++q) {
            int64_t e = adj[q] >> 2;
            int a = adj[q] & 3;
            const double *Ke = scratchK + 144 * e;
            for (int c = 0; c < 3; ++c) {
                int64_t r = 3 * nd + c;
                F_int[r] += scratchf[12 * e + 3 * a + c];
                for (int b = 0; b < 4; ++b) {
                    int64_t nb = conn[4 * e + b];
                    int64_t p = csr_find(rowptr, col, r, 3 * nb);
                    for (int d = 0; d < 3; ++d) val[p + d] += Ke[(3 * a + c) + 12 * (3 * b + d)];
                }
            }
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ linear solve */

/* y = A[free,free] * x restricted through a mask: the reference extracts A = K[free,free]
 * (NonLinearStaticAnalyses.jl:120).  Here vectors stay n_dofs long and entries at fixed
 * dofs are held at zero, which is the same linear system. */
static void spmv_masked(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const uint8_t *free_mask,
                        const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0;
        if (free_mask[i])
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) s += val[p] * x[col[p]]; /* x is 0 at fixed dofs */
        y[i] = s;
    }
}

static double dot_n(int64_t n, const double *a, const double *b) {
    double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* Conjugate gradients exactly as IterativeSolvers.jl 0.9.4 `cg!` runs it for the reference's
 * call site (NonLinearStaticAnalyses.jl:129-134 through LinearSolve's IterativeSolversJL_CG
 * wrapper; StructuralSolvers.jl:29,229-234): zero initial guess, tolerance =
 * max(reltol*||r0||, abstol), stop when ||r|| <= tolerance or iteration >= maxiter.
 * Third-party, absent from /root/reference: restated from the package's published
 * CGIterable / PCGIterable iteration (src/cg.jl):
 *   CG :  beta = res^2/prev_res^2; u = r + beta u; c = A u; alpha = res^2/(u.c);
 *         x += alpha u; r -= alpha c; prev_res = res; res = ||r||
 *   PCG:  c = Pl \ r; rho_prev = rho; rho = c.r; beta = rho/rho_prev (rho0 = 1, u0 = 0);
 *         u = c + beta u; c = A u; alpha = rho/(u.c); x += alpha u; r -= alpha c; res = ||r||
 * diag == NULL selects the un-preconditioned iteration (the reference default); otherwise
 * Pl = Diagonal(diag) (the Jacobi preconditioner the north star asks for on the GPU).
 * Vectors are n long with zeros at fixed dofs.  Returns iterations in *iters, final ||r|| in *res. */
int orc_cg(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const uint8_t *free_mask,
           const double *b, double *x, const double *diag, double reltol, double abstol, int64_t maxiter,
           int64_t *iters, double *res_out) {
    double *r = (double *)malloc((size_t)n * sizeof(double));
    double *u = (double *)calloc((size_t)n, sizeof(double));
    double *c = (double *)malloc((size_t)n * sizeof(double));
    if (!r || !u || !c) {
        free(r);
        free(u);
        free(c);
        return ORC_ERR_PATTERN;
    }
    for (int64_t i = 0; i < n; ++i) {
        x[i] = 0.0;
        r[i] = free_mask[i] ? b[i] : 0.0;
    }
    double residual = sqrt(dot_n(n, r, r));
    double tol = fmax(reltol * residual, abstol);
    double prev_residual = 1.0, rho = 1.0;
    int64_t it = 0;
    while (!(it >= maxiter || residual <= tol)) {
        double beta;
        if (diag) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) c[i] = free_mask[i] ? r[i] / diag[i] : 0.0;
            double rho_prev = rho;
            rho = dot_n(n, c, r);
            beta = rho / rho_prev;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) u[i] = c[i] + beta * u[i];
        } else {
            beta = residual * residual / (prev_residual * prev_residual);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) u[i] = r[i] + beta * u[i];
        }
        spmv_masked(n, rowptr, col, val, free_mask, u, c);
        double alpha = (diag ? rho : residual * residual) / dot_n(n, u, c);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            x[i] += alpha * u[i];
            r[i] -= alpha * c[i];
        }
        prev_residual = residual;
        residual = sqrt(dot_n(n, r, r));
        ++it;
    }
    *iters = it;
    *res_out = residual;
    free(r);
    free(u);
    free(c);
    return ORC_OK;
}
