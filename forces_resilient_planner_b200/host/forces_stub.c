/*
 * forces_stub.c -- what goes into libFORCESNLPsolver_normal.a / libFORCESNLPsolver_final.a of the drop-in tree.
 *
 * The reference links its two solvers as static archives BY FILE NAME
 * (/root/reference/src/resilient_planner/plan_manage/CMakeLists.txt:64-65 link_directories, :82-83
 * `libFORCESNLPsolver_normal.a libFORCESNLPsolver_final.a`).  These stubs keep that link line unchanged: each defines
 * the reference's solver symbol (header :321-323) and forwards, on first use, to libnmpc_b200.so -- the CUDA library,
 * which cannot live in a plain archive without adding cudart to the planner's link line -- through dlopen / dlsym.
 *
 *   where the library is looked for:  $NMPC_B200_LIB, then the absolute path recorded when the archive was built
 *                                     (-DNMPC_B200_DEFAULT_LIB), then "libnmpc_b200.so" on the loader's search path
 *   when it cannot be loaded:         the call returns LICENSE_ERROR (-100, "solver not valid on this machine", header
 *                                     :139) and says why on stderr (and on `fs` when given); the planner handles it like
 *                                     any other failed solve (nmpc_solver.cpp:398-421)
 *
 * Compile once per variant: -DNMPC_STUB_VARIANT=normal | final and -DNMPC_STUB_HEADER='"FORCESNLPsolver_<variant>.h"'.
 * Plain C, no CUDA, no C++ runtime.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#define NMPC_CAT_(a, b) a##b
#define NMPC_CAT(a, b) NMPC_CAT_(a, b)
#define NMPC_STR_(a) #a
#define NMPC_STR(a) NMPC_STR_(a)
#ifndef NMPC_STUB_VARIANT
#error "compile with -DNMPC_STUB_VARIANT=normal or -DNMPC_STUB_VARIANT=final"
#endif
#define V NMPC_STUB_VARIANT
#define SOLVER NMPC_CAT(FORCESNLPsolver_, V)

/* the variant's own header (include/FORCESNLPsolver_<variant>.h), named by the build: -DNMPC_STUB_HEADER='"..."' */
#ifndef NMPC_STUB_HEADER
#error "compile with -DNMPC_STUB_HEADER='\"FORCESNLPsolver_<variant>.h\"'"
#endif
#include NMPC_STUB_HEADER

#define T(suffix) NMPC_CAT(SOLVER, suffix)
typedef solver_int32_default (*solve_fn)(T(_params) *, T(_output) *, T(_info) *, FILE *, T(_extfunc));

static solve_fn resolve(FILE *fs)
{
    static solve_fn fn = NULL;
    if (fn) return fn;
    const char *cand[3] = {getenv("NMPC_B200_LIB"),
#ifdef NMPC_B200_DEFAULT_LIB
                           NMPC_B200_DEFAULT_LIB,
#else
                           NULL,
#endif
                           "libnmpc_b200.so"};
    void *h = NULL;
    for (int i = 0; i < 3 && !h; i++)
        if (cand[i] && cand[i][0]) h = dlopen(cand[i], RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        const char *why = dlerror();
        fprintf(stderr, NMPC_STR(SOLVER) "_solve: cannot load libnmpc_b200.so (%s); set NMPC_B200_LIB\n", why ? why : "?");
        if (fs) fprintf(fs, NMPC_STR(SOLVER) "_solve: cannot load libnmpc_b200.so (%s)\n", why ? why : "?");
        return NULL;
    }
    fn = (solve_fn)dlsym(h, "nmpc_forces_" NMPC_STR(V) "_solve");
    if (!fn) fprintf(stderr, NMPC_STR(SOLVER) "_solve: libnmpc_b200.so lacks nmpc_forces_" NMPC_STR(V) "_solve\n");
    return fn;
}

solver_int32_default T(_solve)(T(_params) * params, T(_output) * output, T(_info) * info, FILE *fs, T(_extfunc) extfunc)
{
    solve_fn fn = resolve(fs);
    if (!fn) return NMPC_CAT(LICENSE_ERROR_, SOLVER);
    return fn(params, output, info, fs, extfunc);
}
