// element_math.cuh -- per-element FP64 arithmetic of the ONSAS.jl hot path, written for the
// row-owner assembly scheme: a thread evaluates ONE block-row `a` of an
// element's tangent (its 4 (tet) / 2 (truss) dim x dim blocks), the matching slice f_a of the
// internal force and, when asked, the element's stress / strain output.
//
// Closed forms replace the reference's 6x12 B-matrix products (all are algebraically equal to
// the reference expressions, see the derivations in DESIGN.md section 3.2):
//   reference Tetrahedrons.jl:186-224 (hyperelastic), :240-266 (linear), :140-152 (_B_mat!),
//   :161-182 (geometric stiffness); SVKMaterial.jl:89-100; NeoHookeanMaterial.jl:94-138;
//   IsotropicLinearElasticMaterial.jl:70-92; Trusses.jl:126-237.
//
// The functions are __host__ __device__ so tests/hostsim can compile them with g++ and check
// the arithmetic against the oracle without a GPU.  The shipped library only ever calls them
// from device code.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ONSAS_HD __host__ __device__ __forceinline__
#else
#define ONSAS_HD inline
#endif

namespace onsas {

enum : int { MAT_SVK = 0, MAT_NEOHOOKEAN = 1, MAT_ISOLINEAR = 2, MAT_MIXED = 3 };
enum : int { STRAIN_ROTENG = 0, STRAIN_GREEN = 1 };

// Everything about one tetrahedron that does not depend on the block-row being evaluated.
struct TetCommon {
    double g[4][3];   // material gradients of the shape functions ("funder" columns)
    double vol;       // reference volume, det(J)/6
    double F[3][3];   // deformation gradient (identity-free H for ISOLINEAR: holds H)
    double w[4][3];   // SVK: F g_k ; NeoHookean: h_k = F^-T g_k ; IsoLinear: unused
    double S[6];      // SVK: 2nd Piola-Kirchhoff, Voigt 11,22,33,23,13,12 ; IsoLinear: Cauchy stress
    double b[6];      // SVK: F F^T (sym, same Voigt order)
    double c0, c1, c2;  // material scalars, see tet_common
};

ONSAS_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Shape-function gradients and volume.  With the reference's dX/dzeta (Tetrahedrons.jl:269-271)
// J = [X0-X1, X3-X1, X2-X1] (columns) and funder = inv(J)' * dXdzeta gives
// g0 = row0(inv J), g3 = row1(inv J), g2 = row2(inv J), g1 = -(g0+g2+g3)  (:129-131, :197).
ONSAS_HD void tet_gradients(const double X[4][3], double g[4][3], double& vol) {
    double c0[3], c1[3], c2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        c0[i] = X[0][i] - X[1][i];
        c1[i] = X[3][i] - X[1][i];
        c2[i] = X[2][i] - X[1][i];
    }
    // rows of inv(J) are cross products of the other two columns over det(J)
    double r0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
    double r1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
    double r2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
    double det = c0[0] * r0[0] + c0[1] * r0[1] + c0[2] * r0[2];
    vol = det / 6.0;  // Tetrahedrons.jl:134
    double id = 1.0 / det;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g[0][i] = r0[i] * id;
        g[3][i] = r1[i] * id;
        g[2][i] = r2[i] * id;
        g[1][i] = -(g[0][i] + g[2][i] + g[3][i]);
    }
}

// Kinematics + constitutive scalars shared by all block-rows of the element.
// mat parameters: SVK (lambda, G) | NeoHookean (K, G) | IsotropicLinearElastic (E, nu).
template <int KIND>
ONSAS_HD void tet_common(const double X[4][3], const double U[4][3], double p0, double p1, TetCommon& c) {
    tet_gradients(X, c.g, c.vol);
    // H = U * funder' (:198)
    double H[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            H[i][j] = U[0][i] * c.g[0][j] + U[1][i] * c.g[1][j] + U[2][i] * c.g[2][j] + U[3][i] * c.g[3][j];

    if (KIND == MAT_ISOLINEAR) {
        // eps = (H + H')/2 (:250); sigma = lambda tr(eps) I + 2G eps (IsotropicLinearElasticMaterial.jl:80-92)
        double G = p0 / (2 * (1 + p1));
        double lam = p0 * p1 / ((1 + p1) * (1 - 2 * p1));
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) c.F[i][j] = H[i][j];
        double e[6] = {H[0][0], H[1][1], H[2][2], 0.5 * (H[1][2] + H[2][1]), 0.5 * (H[0][2] + H[2][0]),
                       0.5 * (H[0][1] + H[1][0])};
        double tr = e[0] + e[1] + e[2];
#pragma unroll
        for (int v = 0; v < 3; ++v) c.S[v] = lam * tr + 2 * G * e[v];
#pragma unroll
        for (int v = 3; v < 6; ++v) c.S[v] = 2 * G * e[v];
#pragma unroll
        for (int v = 0; v < 6; ++v) c.b[v] = e[v];  // small strain kept for the output
        c.c0 = c.vol * lam;
        c.c1 = c.vol * G;
        c.c2 = 0.0;
        return;
    }

#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c.F[i][j] = H[i][j] + (i == j ? 1.0 : 0.0);  // :199

    if (KIND == MAT_SVK) {
        // E = (H + H' + H'H)/2 (:201); S = lambda tr(E) I + 2G E (SVKMaterial.jl:89-100)
        const int VI[6] = {0, 1, 2, 1, 0, 0}, VJ[6] = {0, 1, 2, 2, 2, 1};
        double E[6];
#pragma unroll
        for (int v = 0; v < 6; ++v) {
            int i = VI[v], j = VJ[v];
            E[v] = 0.5 * (H[i][j] + H[j][i] + (H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j]));
        }
        double tr = E[0] + E[1] + E[2];
#pragma unroll
        for (int v = 0; v < 3; ++v) c.S[v] = p0 * tr + 2 * p1 * E[v];
#pragma unroll
        for (int v = 3; v < 6; ++v) c.S[v] = 2 * p1 * E[v];
#pragma unroll
        for (int v = 0; v < 6; ++v) {
            int i = VI[v], j = VJ[v];
            c.b[v] = c.F[i][0] * c.F[j][0] + c.F[i][1] * c.F[j][1] + c.F[i][2] * c.F[j][2];  // F F^T
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int i = 0; i < 3; ++i) c.w[k][i] = c.F[i][0] * c.g[k][0] + c.F[i][1] * c.g[k][1] + c.F[i][2] * c.g[k][2];
        c.c0 = c.vol * p0;  // vol*lambda
        c.c1 = c.vol * p1;  // vol*G
        c.c2 = 0.0;
        return;
    }

    // NeoHookean (NeoHookeanMaterial.jl:94-102): C = F'F, J = sqrt(det C) = |det F|,
    // S = G (I - C^-1) + K J (J-1) C^-1.  With h_k = F^-T g_k (spatial shape gradients):
    //   F S g_a = G F g_a + (kj - G) h_a,  g_a.S g_b = G g_a.g_b + (kj - G) h_a.h_b,
    //   K^mat_ab = vol [ K J (2J-1) h_a h_b' + (G - kj)(h_b h_a' + (h_a.h_b) I) ],  kj = K J (J-1)
    // (tangent = exact derivative of S, which is what the reference's ForwardDiff call returns).
    {
        const double(*F)[3] = c.F;
        double cof[3][3];  // cofactor matrix: F^-T = cof / det F
        cof[0][0] = F[1][1] * F[2][2] - F[1][2] * F[2][1];
        cof[0][1] = F[1][2] * F[2][0] - F[1][0] * F[2][2];
        cof[0][2] = F[1][0] * F[2][1] - F[1][1] * F[2][0];
        cof[1][0] = F[0][2] * F[2][1] - F[0][1] * F[2][2];
        cof[1][1] = F[0][0] * F[2][2] - F[0][2] * F[2][0];
        cof[1][2] = F[0][1] * F[2][0] - F[0][0] * F[2][1];
        cof[2][0] = F[0][1] * F[1][2] - F[0][2] * F[1][1];
        cof[2][1] = F[0][2] * F[1][0] - F[0][0] * F[1][2];
        cof[2][2] = F[0][0] * F[1][1] - F[0][1] * F[1][0];
        double detF = F[0][0] * cof[0][0] + F[0][1] * cof[0][1] + F[0][2] * cof[0][2];
        double id = 1.0 / detF;
        double J = fabs(detF);
        double kj = p0 * J * (J - 1);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                c.w[k][i] = (cof[i][0] * c.g[k][0] + cof[i][1] * c.g[k][1] + cof[i][2] * c.g[k][2]) * id;  // h_k
        // F^-T kept in b/S storage for the stress output: S[0..5] + b[0..2] hold the 9 entries row-major
        c.S[0] = cof[0][0] * id; c.S[1] = cof[0][1] * id; c.S[2] = cof[0][2] * id;
        c.S[3] = cof[1][0] * id; c.S[4] = cof[1][1] * id; c.S[5] = cof[1][2] * id;
        c.b[0] = cof[2][0] * id; c.b[1] = cof[2][1] * id; c.b[2] = cof[2][2] * id;
        c.c0 = c.vol * p0 * J * (2 * J - 1);  // vol * K J (2J-1)
        c.c1 = c.vol * p1;                    // vol * G
        c.c2 = c.vol * (kj - p1);             // vol * (kj - G)
    }
}

// Block-row `a` of the element: blk[9*b+3*r+c] = K_e[3a+r, 3b+c] for b = 0..3, f[r] = f_e[3a+r].
// blk / f may point straight into shared memory (entries are written once, as they are formed).
// `a` may be a run-time value (selection is done with predicated moves, no local-memory arrays).
template <int KIND>
ONSAS_HD void tet_row(const TetCommon& c, const double U[4][3], int a, double* blk /*[4*9]*/, double* f /*[3]*/) {
    double ga[3], wa[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ga[i] = a == 0 ? c.g[0][i] : a == 1 ? c.g[1][i] : a == 2 ? c.g[2][i] : c.g[3][i];
        wa[i] = a == 0 ? c.w[0][i] : a == 1 ? c.w[1][i] : a == 2 ? c.w[2][i] : c.w[3][i];
    }
    if (KIND == MAT_ISOLINEAR) {
        // K_ab = vol [ lambda g_a g_b' + G g_b g_a' + G (g_a.g_b) I ]   (B'DB with F = I, :251-259)
        double fl[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            double gg = c.c1 * dot3(ga, c.g[b]);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double k[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    k[q] = c.c0 * ga[r] * c.g[b][q] + c.c1 * c.g[b][r] * ga[q] + (r == q ? gg : 0.0);
                    blk[9 * b + 3 * r + q] = k[q];
                }
                fl[r] += k[0] * U[b][0] + k[1] * U[b][1] + k[2] * U[b][2];  // f = K u_e (:261)
            }
        }
        f[0] = fl[0]; f[1] = fl[1]; f[2] = fl[2];
        return;
    }
    if (KIND == MAT_SVK) {
        // K_ab = vol [ lambda (F g_a)(F g_b)' + G (F g_b)(F g_a)' + G (g_a.g_b) F F' ] + vol (g_a.S g_b) I
        const double* S = c.S;
        double Sga[3] = {S[0] * ga[0] + S[5] * ga[1] + S[4] * ga[2], S[5] * ga[0] + S[1] * ga[1] + S[3] * ga[2],
                         S[4] * ga[0] + S[3] * ga[1] + S[2] * ga[2]};
        // f_a = vol F (S g_a)  (= B_a' voigt(S) vol, :206-207)
#pragma unroll
        for (int r = 0; r < 3; ++r) f[r] = c.vol * (c.F[r][0] * Sga[0] + c.F[r][1] * Sga[1] + c.F[r][2] * Sga[2]);
        double lwa[3] = {c.c0 * wa[0], c.c0 * wa[1], c.c0 * wa[2]};
        double gwa[3] = {c.c1 * wa[0], c.c1 * wa[1], c.c1 * wa[2]};
        const double bm[3][3] = {{c.b[0], c.b[5], c.b[4]}, {c.b[5], c.b[1], c.b[3]}, {c.b[4], c.b[3], c.b[2]}};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            double gg = c.c1 * dot3(ga, c.g[b]);
            double geo = c.vol * dot3(Sga, c.g[b]);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    blk[9 * b + 3 * r + q] = lwa[r] * c.w[b][q] + c.w[b][r] * gwa[q] + gg * bm[r][q] + (r == q ? geo : 0.0);
        }
        return;
    }
    // NeoHookean
    {
        double Fga[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) Fga[r] = c.F[r][0] * ga[0] + c.F[r][1] * ga[1] + c.F[r][2] * ga[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) f[r] = c.c1 * Fga[r] + c.c2 * wa[r];
        double cwa[3] = {c.c0 * wa[0], c.c0 * wa[1], c.c0 * wa[2]};
        double dwa[3] = {-c.c2 * wa[0], -c.c2 * wa[1], -c.c2 * wa[2]};  // vol (G - kj) h_a
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            // material: c0 h_a h_b' + (G-kj)(h_b h_a' + h_a.h_b I); geometric: G g_a.g_b + (kj-G) h_a.h_b
            // -> the h_a.h_b terms cancel: diagonal coefficient = vol G (g_a.g_b)
            double dg = c.c1 * dot3(ga, c.g[b]);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 3; ++q) blk[9 * b + 3 * r + q] = cwa[r] * c.w[b][q] + c.w[b][r] * dwa[q] + (r == q ? dg : 0.0);
        }
    }
}

// Element stress / strain output, 16 doubles: sig 3x3 column-major [0..8], then the symmetric
// strain as (11,22,33,23,13,12) [9..14], [15] = 0.  Hyperelastic: sig = P = F S, strain = C = F'F
// (:218-221); linear: Cauchy sigma, small strain (:265).
template <int KIND>
ONSAS_HD void tet_stress_out(const TetCommon& c, double p0, double p1, double out[16]) {
    const int VI[6] = {0, 1, 2, 1, 0, 0}, VJ[6] = {0, 1, 2, 2, 2, 1};
    if (KIND == MAT_ISOLINEAR) {
        const double* S = c.S;
        out[0] = S[0]; out[1] = S[5]; out[2] = S[4];
        out[3] = S[5]; out[4] = S[1]; out[5] = S[3];
        out[6] = S[4]; out[7] = S[3]; out[8] = S[2];
#pragma unroll
        for (int v = 0; v < 6; ++v) out[9 + v] = c.b[v];
        out[15] = 0.0;
        return;
    }
    if (KIND == MAT_SVK) {
        const double* S = c.S;
        const double Sm[3][3] = {{S[0], S[5], S[4]}, {S[5], S[1], S[3]}, {S[4], S[3], S[2]}};
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) out[i + 3 * j] = c.F[i][0] * Sm[0][j] + c.F[i][1] * Sm[1][j] + c.F[i][2] * Sm[2][j];
    } else {
        // P = F S = G F + (kj - G) F^-T ; c1 = vol G, c2 = vol (kj - G)
        double iv = 1.0 / c.vol;
        double Gm = c.c1 * iv, Km = c.c2 * iv;
        const double FiT[3][3] = {{c.S[0], c.S[1], c.S[2]}, {c.S[3], c.S[4], c.S[5]}, {c.b[0], c.b[1], c.b[2]}};
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) out[i + 3 * j] = Gm * c.F[i][j] + Km * FiT[i][j];
    }
#pragma unroll
    for (int v = 0; v < 6; ++v) {
        int i = VI[v], j = VJ[v];
        out[9 + v] = c.F[0][i] * c.F[0][j] + c.F[1][i] * c.F[1][j] + c.F[2][i] * c.F[2][j];  // C = F'F
    }
    out[15] = 0.0;
    (void)p0;
    (void)p1;
}

// ------------------------------------------------------------------------------------------ truss
// Block-row `a` (0 or 1) of a 2-node truss in DIM dimensions (Trusses.jl:126-184).
// blk[b][DIM*r+c] = K_e[DIM*a+r, DIM*b+c]; f[r] = f_e[DIM*a+r]; se = (P11, eps11) (:148-152).
// Emod = elasticity_modulus(material) (SVKMaterial.jl:74-77), A = area(cross_section).
template <int DIM>
ONSAS_HD void truss_row(int strain_model, const double X[2][3], const double U[2][3], double Emod, double A, int a,
                        double blk[2][9], double f[3], double se[2]) {
    double dref[3], ddef[3], l_ref2 = 0, l_def2 = 0;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        dref[c] = X[1][c] - X[0][c];
        ddef[c] = (X[1][c] + U[1][c]) - (X[0][c] + U[0][c]);
        l_ref2 += dref[c] * dref[c];
        l_def2 += ddef[c] * ddef[c];
    }
    double l_ref = sqrt(l_ref2), l_def = sqrt(l_def2);  // _lengths :232-237
    double sa = a == 0 ? -1.0 : 1.0;                    // sign of node a in Bdif = [-I I]
    if (strain_model == STRAIN_ROTENG) {
        double eps = (l_def * l_def - l_ref * l_ref) / (l_ref * (l_ref + l_def));  // :187-189
        double S11 = Emod * eps;
        double e1[3];
#pragma unroll
        for (int c = 0; c < DIM; ++c) e1[c] = ddef[c] / l_def;
        double km = Emod * A / l_ref, kg = S11 * A / l_def;
#pragma unroll
        for (int r = 0; r < DIM; ++r) f[r] = A * S11 * sa * e1[r];  // :142
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            double sab = sa * (b == 0 ? -1.0 : 1.0);
#pragma unroll
            for (int r = 0; r < DIM; ++r)
#pragma unroll
                for (int q = 0; q < DIM; ++q) {
                    double tt = sab * e1[r] * e1[q];
                    blk[b][DIM * r + q] = km * tt + kg * ((r == q ? sab : 0.0) - tt);  // :144-146
                }
        }
        se[0] = S11 * l_def / l_ref;
        se[1] = eps;
    } else {
        double eps = (l_def * l_def - l_ref * l_ref) / (2 * l_ref * l_ref);  // :192-194
        double S11 = Emod * eps;
        double il2 = 1.0 / (l_ref * l_ref);
        double bs[3];  // b_sum for node 1 (node 0 has the opposite sign): (X+u)'Ge / l_ref^2 (:221-229)
#pragma unroll
        for (int c = 0; c < DIM; ++c) bs[c] = dref[c] * il2 + (U[1][c] - U[0][c]) * il2;
#pragma unroll
        for (int r = 0; r < DIM; ++r) f[r] = A * S11 * l_ref * sa * bs[r];  // :174
        double k1 = S11 * A / l_ref, k2 = Emod * A * l_ref;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            double sab = sa * (b == 0 ? -1.0 : 1.0);
#pragma unroll
            for (int r = 0; r < DIM; ++r)
#pragma unroll
                for (int q = 0; q < DIM; ++q)
                    blk[b][DIM * r + q] = k1 * (r == q ? sab : 0.0) + k2 * sab * bs[r] * bs[q];  // :176
        }
        se[0] = S11 * l_def / l_ref;
        se[1] = eps;
    }
}

// elasticity modulus seen by a truss for material (kind, p0, p1)
ONSAS_HD double truss_modulus(int kind, double p0, double p1) {
    double G = p1;
    double lam = kind == MAT_SVK ? p0 : p0 - 2 * G / 3;  // NeoHookeanMaterial.jl:68-72
    return G * (3 * lam + 2 * G) / (lam + G);             // SVKMaterial.jl:74-77
}

}  // namespace onsas
