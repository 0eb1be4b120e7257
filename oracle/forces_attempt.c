/*
 * forces_attempt.c -- ONE logged attempt to run the reference's own solver binary.
 * TEST / MEASUREMENT INFRASTRUCTURE; never part of the product.
 *
 * Linked (oracle/Makefile, target `ref`) against the reference's static archive and model callbacks WHERE THEY LIE:
 *   /root/reference/src/resilient_planner/plan_manage/solver/normal/FORCESNLPsolver_normal/lib/libFORCESNLPsolver_normal.a
 *   .../solver/normal/FORCESNLPsolver_normal_casadi.c, FORCESNLPsolver_normal_casadi2forces.c
 * and called exactly as the planner calls it (plan_manage/src/forces_normal.cpp:30-31,139) on BASELINE config 1
 * (hover at (0,0,1), 1 m/s reference along +x, 6-plane box tightened by the ego ellipsoid).
 * Prints one JSON object: the exit flag (the archive is licence-locked: -100 on any machine but its authors'),
 * iterations and the solver's own solvetime -- the record of WHY bench.py has no ForcesPro timing.
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "FORCESNLPsolver_normal.h"

extern void FORCESNLPsolver_normal_casadi2forces(FORCESNLPsolver_normal_float *x, FORCESNLPsolver_normal_float *y,
    FORCESNLPsolver_normal_float *l, FORCESNLPsolver_normal_float *p, FORCESNLPsolver_normal_float *f,
    FORCESNLPsolver_normal_float *nabla_f, FORCESNLPsolver_normal_float *c, FORCESNLPsolver_normal_float *nabla_c,
    FORCESNLPsolver_normal_float *h, FORCESNLPsolver_normal_float *nabla_h, FORCESNLPsolver_normal_float *hess,
    solver_int32_default stage, solver_int32_default iteration, solver_int32_default threadID);

int main(void)
{
    static FORCESNLPsolver_normal_params params;
    static FORCESNLPsolver_normal_output output;
    static FORCESNLPsolver_normal_info info;
    memset(&params, 0, sizeof params);
    memset(&info, 0, sizeof info);
    params.num_of_threads = 1;                                  /* forces_normal.cpp:31 */
    params.xinit[2] = 1.0;
    const double A[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    const double b[6] = {3, 2, 2, 2, 2, 0}, E[3] = {0.27, 0.27, 0.0425};
    for (int k = 0; k < 20; k++) {
        double *z = params.x0 + 17 * k, *p = params.all_parameters + 130 * k;
        z[3] = z[7] = 7.3; z[10] = 1.0;                          /* initMPCOutput, nmpc_solver.cpp:272-276 */
        p[0] = 0.05 * (k + 1); p[2] = 1.0;
        p[6] = (k == 19) ? 12.0 : 7.0; p[7] = (k == 19) ? 0.5 : 1.0; p[8] = 80.0;   /* rotors_sim.launch:56-66 */
        for (int j = 0; j < 6; j++) {
            double n = 0;
            for (int i = 0; i < 3; i++) { p[10 + 3 * j + i] = A[j][i]; n += E[i] * A[j][i] * E[i] * A[j][i]; }
            p[100 + j] = b[j] - sqrt(n);                         /* forces_normal.cpp:124-125 */
        }
    }
    fflush(stdout);
    const int flag = FORCESNLPsolver_normal_solve(&params, &output, &info, NULL, &FORCESNLPsolver_normal_casadi2forces);
    fflush(stdout);
    printf("\n{\"linked\": true, \"exitflag\": %d, \"it\": %d, \"solvetime\": %.6g, \"library\": "
           "\"solver/normal/FORCESNLPsolver_normal/lib/libFORCESNLPsolver_normal.a\"}\n", flag, info.it, info.solvetime);
    return 0;
}
