/* abi_smoke.c -- TEST ONLY: include/onsas_cuda.h compiled as plain C99 and linked against libonsas_cuda.so, the way a cgo / ccall /
 * JNI binding sees it (no C++ types, no torch).  Without a GPU the life cycle must fail loudly and cleanly: onsas_create returns
 * ONSAS_ERR_CUDA and a message, the null-handle calls return ONSAS_ERR_INVALID_ARG, nothing crashes.  With a GPU it runs the
 * reference's 6-tet uniaxial-extension cube (examples/uniaxial_extension/uniaxial_extension.jl:11-24,45-72) for one Newton step
 * through the C entry points only.  Exit code 0 = every check held. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "onsas_cuda.h"

#define CHECK(cond, what)                                      \
    do {                                                       \
        if (!(cond)) {                                         \
            fprintf(stderr, "abi_smoke: FAILED: %s\n", what); \
            return 1;                                          \
        }                                                      \
    } while (0)

int main(void) {
    onsas_ctx* ctx = NULL;
    int32_t st;
    CHECK(onsas_version() >= 100, "onsas_version");
    CHECK(onsas_create(0, NULL) == ONSAS_ERR_INVALID_ARG, "onsas_create(NULL out pointer) must be an argument error");
    CHECK(onsas_assemble(NULL) == ONSAS_ERR_INVALID_ARG, "a null handle must be an argument error");
    st = onsas_create(0, &ctx);
    if (st != ONSAS_OK) {
        CHECK(st == ONSAS_ERR_CUDA, "without a device onsas_create must return ONSAS_ERR_CUDA");
        CHECK(ctx == NULL, "no context may be handed out on failure");
        printf("abi_smoke: no CUDA device: life cycle fails loudly (status %d), C linkage ok\n", (int)st);
        return 0;
    }
    {
        /* the reference's unit cell: 8 nodes, 6 tets (0-based), SVK E = 1, nu = 0.3 */
        const double xyz[24] = {0, 0, 0, 0, 0, 1, 0, 1, 1, 0, 1, 0, 2, 0, 0, 2, 0, 1, 2, 1, 1, 2, 1, 0};
        const int32_t tets[24] = {0, 3, 1, 5, 5, 1, 2, 3, 3, 2, 5, 6, 3, 0, 4, 5, 3, 5, 4, 7, 3, 6, 5, 7};
        const int32_t kind[1] = {ONSAS_MAT_SVK};
        const double E = 1.0, nu = 0.3;
        const double par[2] = {E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))};
        /* u_x = 0 on x = 0 (nodes 0-3), u_y = 0 on y = 0 (nodes 0, 1, 4, 5), u_z = 0 on z = 0 (nodes 0, 3, 4, 7) */
        int64_t free_dofs[24];
        int64_t nfree = 0;
        double F[24], U[24];
        onsas_step_info info;
        int i;
        for (i = 0; i < 24; ++i) {
            const int n = i / 3, c = i % 3;
            const int fixed = (c == 0 && n <= 3) || (c == 1 && (n == 0 || n == 1 || n == 4 || n == 5)) || (c == 2 && (n == 0 || n == 3 || n == 4 || n == 7));
            if (!fixed) free_dofs[nfree++] = i;
            F[i] = 0.0;
            U[i] = 0.0;
        }
        for (i = 4; i < 8; ++i) F[3 * i] = 0.01;   /* a small pull on the face x = 2 */
        CHECK(onsas_set_nodes(ctx, 8, 8, 3, xyz) == ONSAS_OK, onsas_last_error(ctx));
        CHECK(onsas_set_materials(ctx, 1, kind, par) == ONSAS_OK, onsas_last_error(ctx));
        CHECK(onsas_set_tets(ctx, 6, tets, NULL) == ONSAS_OK, onsas_last_error(ctx));
        CHECK(onsas_set_free_dofs(ctx, nfree, free_dofs, nfree) == ONSAS_OK, onsas_last_error(ctx));
        CHECK(onsas_finalize_mesh(ctx) == ONSAS_OK, onsas_last_error(ctx));
        CHECK(onsas_set_U(ctx, U) == ONSAS_OK, onsas_last_error(ctx));
        CHECK(onsas_set_Fext(ctx, F) == ONSAS_OK, onsas_last_error(ctx));
        memset(&info, 0, sizeof info);
        CHECK(onsas_newton_step(ctx, ONSAS_PRECOND_JACOBI, 1e-12, 0.0, -1, &info) == ONSAS_OK, onsas_last_error(ctx));
        CHECK(onsas_get_U(ctx, U) == ONSAS_OK, onsas_last_error(ctx));
        CHECK(info.cg_iters > 0 && U[12] > 0.0 && fabs(U[0]) == 0.0, "one Newton step must move the loaded face and keep the fixed dofs");
        printf("abi_smoke: one Newton step through the C ABI: %lld CG iterations, u_x(x = 2) = %.6e\n", (long long)info.cg_iters, U[12]);
    }
    CHECK(onsas_destroy(ctx) == ONSAS_OK, "onsas_destroy");
    return 0;
}
