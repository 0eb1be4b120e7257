// a planner-side translation unit: only the reference's header names, macros and symbols (forces_normal.cpp:1-35,139)
#include "FORCESNLPsolver_normal.h"
#include "FORCESNLPsolver_final.h"
#include <cstdio>
#include <cstring>
#if SET_PRINTLEVEL_FORCESNLPsolver_normal > 0 && SET_TIMING_FORCESNLPsolver_normal == 1 && MAX_SOC_IT_FORCESNLPsolver_final == 4
static const int maxit = SET_MAXIT_FORCESNLPsolver_normal;
#endif
extern "C" void FORCESNLPsolver_normal_casadi2forces(double*, double*, double*, double*, double*, double*, double*, double*, double*, double*, double*, int, int, int) {}
int main() {
    static FORCESNLPsolver_normal_params p; static FORCESNLPsolver_normal_output o; static FORCESNLPsolver_normal_info info;
    static FORCESNLPsolver_final_params pf; static FORCESNLPsolver_final_output of; static FORCESNLPsolver_final_info inf;
    std::memset(&p, 0, sizeof p); p.num_of_threads = 1;
    for (int k = 0; k < 20; k++) { p.x0[k*17+3] = p.x0[k*17+7] = 7.3; p.x0[k*17+10] = 1.0; double* q = p.all_parameters + k*130; q[0] = 0.05*(k+1); q[2] = 1.0; q[6] = 7; q[7] = 1; q[8] = 80; }
    p.xinit[2] = 1.0;
    std::memcpy(&pf, &p, sizeof p);
    int f1 = FORCESNLPsolver_normal_solve(&p, &o, &info, stdout, &FORCESNLPsolver_normal_casadi2forces);
    int f2 = FORCESNLPsolver_final_solve(&pf, &of, &inf, NULL, (FORCESNLPsolver_final_extfunc)&FORCESNLPsolver_normal_casadi2forces);
    std::printf("exitflags %d %d maxit %d it %d\n", f1, f2, maxit, info.it);
    return (f1 == OPTIMAL_FORCESNLPsolver_normal && f2 == OPTIMAL_FORCESNLPsolver_final) ? 0 : (f1 == LICENSE_ERROR_FORCESNLPsolver_normal ? 100 : 1);
}
