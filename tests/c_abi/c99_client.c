/* Plain C99 client of include/diffopt_b200.h: proves the header is C (not C++) and that a non-Python, non-Julia host can
 * drive the hot path: create -> qp_batch_solve (the reference's KAT 1, test/quadratic_program.jl:232-293) -> destroy.
 * Built by tests/test_abi_cpu.py with `gcc -std=c99 -pedantic -Wall -Werror`.  Exit code 0 = ok (or: no GPU and the
 * library refused cleanly, which is all a CPU box can check), 1 = wrong result, 2 = unexpected error. */
#include <math.h>
#include <stdio.h>

#include "diffopt_b200.h"

int main(void) {
    diffopt_b200_ctx* ctx = NULL;
    int32_t rc = diffopt_b200_create(0, &ctx);
    if (rc != 0) {
        printf("no usable GPU (diffopt_b200_create = %d): nothing to run, there is no CPU fallback\n", (int)rc);
        return (rc == -2 || rc == -4) ? 0 : 2;
    }
    /* column-major data of ONE instance: Q = [4 1; 1 2], G = -I, h = 0, A = [1 1], z = (.25, .75), lam = 0, nu = -2.75 */
    const double Q[4] = {4, 1, 1, 2}, G[4] = {-1, 0, 0, -1}, A[2] = {1, 1}, h[2] = {0, 0};
    const double z[2] = {0.25, 0.75}, lam[2] = {0, 0}, nu[1] = {-2.75}, seed[2] = {1.3, 0.5};
    double rev[5];
    int32_t info[1] = {-1};
    rc = diffopt_b200_qp_batch_solve(ctx, 1, 2, 2, 1, Q, G, A, h, z, lam, nu, NULL, NULL, NULL, NULL, NULL, NULL, seed, NULL, rev,
                                     info, DIFFOPT_B200_HOST);
    if (rc != 0 || info[0] != 0) {
        printf("qp_batch_solve failed: rc = %d (%s)\n", (int)rc, diffopt_b200_last_error(ctx));
        diffopt_b200_destroy(ctx);
        return 2;
    }
    /* expected (grad_z, grad_lam, grad_nu) = (-0.2, 0.2, 0.8, -0.8/3, -0.7) */
    const double want[5] = {-0.2, 0.2, 0.8, -0.8 / 3.0, -0.7};
    int bad = 0, i;
    for (i = 0; i < 5; ++i)
        if (fabs(rev[i] - want[i]) > 1e-9) bad = 1;
    printf("c99 client: rev = %.6f %.6f %.6f %.6f %.6f (%s), %lld kernel launches\n", rev[0], rev[1], rev[2], rev[3], rev[4],
           bad ? "WRONG" : "ok", (long long)diffopt_b200_launch_count(ctx));
    /* forward mode with the direction as sparse triplets (diffopt_b200_coo_batch is a plain C struct): dG = e_1 e_2' given
     * once as (1, 2, 1.0) and once split into two duplicates that must add up -- both calls must agree */
    {
        const int64_t ptr1[2] = {0, 1}, I1[1] = {1}, J1[1] = {2};
        const double V1[1] = {1.0};
        const int64_t ptr2[2] = {0, 2}, I2[2] = {1, 1}, J2[2] = {2, 2};
        const double V2[2] = {0.25, 0.75};
        diffopt_b200_coo_batch g1, g2;
        double f1[5], f2[5];
        g1.ptr = ptr1; g1.I = I1; g1.J = J1; g1.V = V1;
        g2.ptr = ptr2; g2.I = I2; g2.J = J2; g2.V = V2;
        rc = diffopt_b200_qp_batch_solve_coo(ctx, 1, 2, 2, 1, Q, G, A, h, z, lam, nu, NULL, NULL, &g1, NULL, NULL, NULL, NULL, f1, NULL,
                                             info, DIFFOPT_B200_HOST, 0);
        if (rc == 0)
            rc = diffopt_b200_qp_batch_solve_coo(ctx, 1, 2, 2, 1, Q, G, A, h, z, lam, nu, NULL, NULL, &g2, NULL, NULL, NULL, NULL, f2,
                                                 NULL, info, DIFFOPT_B200_HOST, 0);
        if (rc != 0) {
            printf("qp_batch_solve_coo failed: rc = %d (%s)\n", (int)rc, diffopt_b200_last_error(ctx));
            diffopt_b200_destroy(ctx);
            return 2;
        }
        for (i = 0; i < 5; ++i)
            if (fabs(f1[i] - f2[i]) > 1e-12) bad = 1;
        printf("c99 client: sparse-triplet forward dz = %.6f %.6f (%s)\n", f1[0], f1[1], bad ? "WRONG" : "ok");
    }
    diffopt_b200_destroy(ctx);
    return bad;
}
