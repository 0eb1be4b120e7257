/*
 * nmpc_b200.h -- C ABI of the B200-native batched NMPC solver (libnmpc_b200.so).
 *
 * Plain pointers and sizes only; no torch / CUDA types in any signature (streams travel as void*).
 *
 * Two families of entry points:
 *
 *  (1) the reference's own solver ABI, preserved bit for bit, so that
 *      plan_manage/src/forces_normal.cpp:139 and forces_final.cpp:138 link unchanged:
 *          FORCESNLPsolver_normal_solve / FORCESNLPsolver_final_solve
 *      -> declared in include/FORCESNLPsolver_normal.h and include/FORCESNLPsolver_final.h.
 *
 *  (2) the batched entry points this file declares.  They take the same information as
 *      FORCESNLPsolver_normal_params (reference:
 *      solver/normal/FORCESNLPsolver_normal/include/FORCESNLPsolver_normal.h:153-168) for B
 *      independent problems, in a layout without the 30-row zero padding:
 *
 *        xinit [B][9]           <- params.xinit                     (header :156)
 *        z0    [B][N][17]       <- params.x0                        (header :159)
 *        hdr   [B][N][10]       <- params.all_parameters[k*130+0..9]   ref(3) f_ext(3) w_wp w_in w_rate yaw_ref
 *        rows  [B][N][mcap][4]  <- all_parameters[k*130+10+3j..] and [k*130+100+j]   (a0 a1 a2 b), a.pos <= b
 *        nrows [B][N] int32     <- number of live rows of each stage (<= mcap)
 *        z_out [B][N][17]       -> output.x01 .. x20                (header :173-236)
 *        info_int  [B][4]       -> exitflag (reference codes, header :110-139), iterations, backtracks, 0
 *        info_real [B][8]       -> res_eq, res_ineq, rsnorm, rcompnorm, pobj, mu, alpha_p, alpha_d
 *                                  (the fields of FORCESNLPsolver_normal_info, header :241-301)
 *
 *      exit flags (info_int[b][0]; reference vocabulary, header :110-139):
 *         1  OPTIMAL        all four inf-norms <= tolerance (mpc_generator_normal.m:76-79)
 *         0  MAXITREACHED   opts.maxit iterations without meeting them
 *        -5  FACTORIZATION  non-positive pivot in the KKT factorisation
 *        -6  BADFUNCEVAL    NaN / Inf at the first evaluation (bad input)
 *        -7  NOPROGRESS     NaN / Inf later on, OR the problem has no feasible point because xinit itself violates a
 *                           stage-0 bound or stage-0 corridor row by more than tol_ineq: the stage-0 states are fixed by
 *                           the xinit equality (header :156, mpc_generator_normal.m:50), so no iteration can repair that;
 *                           returned after zero iterations with the violation in info_real[b][1] (res_ineq).
 *      (A feasible xinit makes the stage-0 bounds and rows redundant; they are then not carried as barrier terms.)
 *
 *      variant 0 = "normal" stage/terminal costs, 1 = "final" (terminal velocity cost,
 *      matlab_code/mpc/final/mpc_objectiveN_final.m:26).
 *
 * Return value of every nmpc_* function: 0 on success, <0 on failure
 * (NMPC_ERR_*; nmpc_last_error() gives the text).  Per-problem solver outcomes are in info_int.
 * There is NO CPU fallback: without a CUDA device every solve entry point returns NMPC_ERR_CUDA.
 */
#ifndef NMPC_B200_H
#define NMPC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMPC_ERR_ARG (-11)    /* invalid argument (mirrors PARAM_VALUE_ERROR, header :131) */
#define NMPC_ERR_CUDA (-101)  /* CUDA runtime failure / no device                           */

typedef struct nmpc_opts {
    double mu0;        /* initial barrier parameter                          (default 1.0)   */
    double sigma;      /* centering; <= 0 selects the LOQO centrality rule   (default 0.1)   */
    double mu_floor;   /* lowest barrier target                              (default 1e-5)  */
    double tol_stat;   /* codeoptions.nlp.TolStat  (mpc_generator_normal.m:76)  1e-4         */
    double tol_eq;     /* codeoptions.nlp.TolEq    (:77)                         1e-4         */
    double tol_ineq;   /* codeoptions.nlp.TolIneq  (:78)                         1e-4         */
    double tol_comp;   /* codeoptions.nlp.TolComp  (:79)                         1e-4         */
    double kappa_push; /* push of the initial guess into the bound interior  (default 1e-2)  */
    double s_floor;    /* floor on initial corridor slacks                   (default 1e-2)  */
    int maxit;         /* codeoptions.maxit (:56)                            200             */
    int max_bt;        /* backtracking steps per iteration                   (default 6)     */
    int pc;            /* 1: Mehrotra predictor-corrector -- affine solve, sigma = (mu_aff/mu)^3, second-order
                          corrector through the same factorisation (default 0; fp64 entry points only)   */
    int mixed;         /* *_f32 / *_mixed_* entry points: -1 disables the fp64 re-solve of the problems the mixed-precision
                          kernel does not bring to exit flag 1 (default 0 = re-solve them).  Ignored by the fp64 entry
                          points.  (The CPU oracle shares this struct; there 1 selects its mixed-precision restatement.) */
} nmpc_opts;

void nmpc_default_opts(nmpc_opts *o);
const char *nmpc_last_error(void);
const char *nmpc_version(void);

/* Supported horizons N: 20 (the reference), 40 (BASELINE config 4).  mcap in [0, 32]. */
int nmpc_supported_horizon(int N);
/* dynamic shared memory one problem occupies (bytes): elem_size 8 = the fp64 kernel, 4 = the mixed-precision kernel */
long nmpc_smem_bytes(int N, int mcap, int elem_size);
long nmpc_smem_bytes_pc(int N, int mcap, int elem_size);   /* the fp64 predictor-corrector kernel (opts.pc = 1); elem_size 8 only */

/* ---- device-pointer API: everything already resident in HBM, asynchronous on `stream` --------
 * z0, hdr, rows, nrows and z_out must be 16-byte aligned (TMA bulk copies / 16-byte vector loads);
 * a misaligned pointer is rejected with NMPC_ERR_ARG.                                           */
int nmpc_solve_batch_f64(int B, int N, int mcap, const double *xinit, const double *z0,
                         const double *hdr, const double *rows, const int *nrows, int variant,
                         const nmpc_opts *opts, double *z_out, int *info_int, double *info_real,
                         void *cuda_stream);

/* ---- mixed precision (BASELINE configs 3 and 4, "fp32") -------------------------------------------
 * Same problem, same REFERENCE tolerances (nmpc_default_opts: 1e-4 inf-norms, mpc_generator_normal.m:76-79), same
 * exit codes.  The Newton system (stage Hessians, Jacobians, right-hand side, Riccati recursion, rollout, costates)
 * is formed and solved in single precision in DELTA form; the iterate, the model evaluation, the KKT residuals
 * that decide termination, the step rule and the line search are double precision, so the outer iteration refines
 * the single-precision solve (csrc/nmpc_ipm_mixed.cuh).  Problems the single-precision factorisation cannot carry
 * (non-positive pivot on badly scaled instances, exit -5; or no convergence within 60 iterations) are re-solved by
 * the fp64 kernel on the same stream before the call's work completes; they carry info_int[b][3] = 1.
 *   nmpc_solve_batch_f32        problem data and results are float arrays in HBM
 *   nmpc_solve_batch_mixed_f64  problem data and results are double arrays; optionally the multipliers (as
 *                               nmpc_solve_batch_ex_f64; any may be NULL) and a scheduling order (as
 *                               nmpc_solve_batch_ordered_f64; NULL = natural order)
 * opts->pc is ignored (no predictor-corrector variant).                                                  */
int nmpc_solve_batch_f32(int B, int N, int mcap, const float *xinit, const float *z0,
                         const float *hdr, const float *rows, const int *nrows, int variant,
                         const nmpc_opts *opts, float *z_out, int *info_int, float *info_real,
                         void *cuda_stream);
int nmpc_solve_batch_mixed_f64(int B, int N, int mcap, const double *xinit, const double *z0,
                               const double *hdr, const double *rows, const int *nrows, int variant,
                               const nmpc_opts *opts, double *z_out, int *info_int, double *info_real,
                               double *y_out, double *zl_out, double *zu_out, double *lc_out,
                               const int *order, void *cuda_stream);

/* Low-latency variant of nmpc_solve_batch_mixed_f64 (same arguments, same algorithm, tolerances and exit codes): one
 * warp-GROUP (256 threads) per problem instead of one warp (csrc/nmpc_ipm_group.cuh).  For small fleets -- fewer
 * problems than the GPU has SMs x a few, e.g. BASELINE config 5 at 128 agents per GPU -- where the latency of one
 * solve, not the throughput of thousands, is what the caller waits for: an interior-point iteration takes about half
 * the time of the one-warp kernels'.  With thousands of problems the one-warp kernels are the faster choice.      */
int nmpc_solve_batch_lowlatency_f64(int B, int N, int mcap, const double *xinit, const double *z0,
                                    const double *hdr, const double *rows, const int *nrows, int variant,
                                    const nmpc_opts *opts, double *z_out, int *info_int, double *info_real,
                                    double *y_out, double *zl_out, double *zu_out, double *lc_out,
                                    const int *order, void *cuda_stream);

/* as nmpc_solve_batch_f64, additionally returning the multipliers of the KKT point (any of the
 * four may be NULL): y_out [B][N][13] (c-ordering [x+(9); u(4)], y[0] = 0), zl_out / zu_out
 * [B][N][17], lc_out [B][N][mcap].  Used to check ForcesPro's acceptance test with the
 * reference callbacks (stationarity needs the multipliers).                                    */
int nmpc_solve_batch_ex_f64(int B, int N, int mcap, const double *xinit, const double *z0,
                            const double *hdr, const double *rows, const int *nrows, int variant,
                            const nmpc_opts *opts, double *z_out, int *info_int, double *info_real,
                            double *y_out, double *zl_out, double *zu_out, double *lc_out,
                            void *cuda_stream);

/* as nmpc_solve_batch_f64 with an explicit scheduling order: CTA i solves problem order[i] (a device
 * permutation of 0..B-1; NULL = natural order).  The hardware starts CTAs in index order, so a
 * longest-first order (e.g. by the previous replan's iteration counts in a receding-horizon stream)
 * shortens the tail of the launch; results are written at the problems' own indices.            */
int nmpc_solve_batch_ordered_f64(int B, int N, int mcap, const double *xinit, const double *z0,
                                 const double *hdr, const double *rows, const int *nrows,
                                 int variant, const nmpc_opts *opts, double *z_out, int *info_int,
                                 double *info_real, const int *order, void *cuda_stream);

/* ---- host-pointer API: H2D copies, solve, D2H copies, synchronous ---------------------------
 * Large batches are cut into chunks, one stream each, so that copies and kernels overlap.  Any host memory works;
 * pinned memory (cudaHostAlloc / cudaHostRegister) is faster, and when z_out (and info_int) are pinned and 16-byte
 * aligned the solve kernel stores each solution straight into them as the problem finishes (no copy after the last
 * kernel).  $NMPC_B200_DIRECT_HOST=0 turns that off.                                                          */
int nmpc_solve_batch_host_f64(int B, int N, int mcap, const double *xinit, const double *z0,
                              const double *hdr, const double *rows, const int *nrows, int variant,
                              const nmpc_opts *opts, double *z_out, int *info_int, double *info_real);
/* host-pointer forms of the mixed-precision solve (float arrays / double arrays) */
int nmpc_solve_batch_host_f32(int B, int N, int mcap, const float *xinit, const float *z0,
                              const float *hdr, const float *rows, const int *nrows, int variant,
                              const nmpc_opts *opts, float *z_out, int *info_int, float *info_real);
int nmpc_solve_batch_host_mixed_f64(int B, int N, int mcap, const double *xinit, const double *z0,
                                    const double *hdr, const double *rows, const int *nrows, int variant,
                                    const nmpc_opts *opts, double *z_out, int *info_int, double *info_real);

/* ---- multi-GPU: one process per GPU, contiguous sharding, end-of-batch NCCL all-gather -----------
 * The NMPC instances are independent (single-vehicle planner): the solve needs no collective.  The only exchange is
 * the collation of the results, and it costs no copy: every rank's kernel writes its z straight into its slice of ONE
 * buffer of world * B_local problems, and an in-place ncclAllGather (sendbuff = recvbuff + rank * count) on the same
 * stream completes the other slices.  Allocate that buffer with nmpc_comm_alloc (ncclMemAlloc + ncclCommRegister:
 * NCCL then works on the user buffer itself, NVLS-eligible) or pass any device buffer.
 *   nmpc_comm_unique_id   rank 0 makes the 128-byte id; the host application distributes it (MPI, a file, torchrun's
 *                         store ... -- plumbing, not part of this library)
 *   nmpc_comm_create      collective over all ranks; the CUDA device of the calling thread is the rank's GPU
 *   nmpc_comm_wrap        adopt an ncclComm_t the application already has (same libnccl)
 *   nmpc_solve_batch_sharded_f64 / _f32
 *                         local shard in (B_local problems, layouts as nmpc_solve_batch_f64 / _f32; every rank the same
 *                         B_local, even), z_all [world * B_local][N][17] and info_int_all [world * B_local][4] out on
 *                         every rank, info_real_local [B_local][8] for the own shard; mixed != 0 (f64) selects the
 *                         mixed-precision kernel; _f32 always uses it
 *   nmpc_collate_inplace  the bare in-place all-gather of any buffer of world * bytes_per_rank bytes
 * libnccl.so.2 is resolved at run time (the copy already loaded in the process, else the system's; $NMPC_B200_NCCL
 * overrides), so single-GPU users carry no NCCL dependency.                                                   */
typedef struct nmpc_comm nmpc_comm;
int nmpc_comm_unique_id(char id[128]);
int nmpc_comm_create(int world, int rank, const char id[128], nmpc_comm **comm);
int nmpc_comm_wrap(void *nccl_comm, int world, int rank, nmpc_comm **comm);
int nmpc_comm_destroy(nmpc_comm *comm);
int nmpc_comm_rank(const nmpc_comm *comm);
int nmpc_comm_world(const nmpc_comm *comm);
int nmpc_comm_nccl_version(void);
int nmpc_comm_alloc(nmpc_comm *comm, size_t bytes, void **ptr);
int nmpc_comm_free(nmpc_comm *comm, void *ptr);
int nmpc_solve_batch_sharded_f64(nmpc_comm *comm, int B_local, int N, int mcap, const double *xinit,
                                 const double *z0, const double *hdr, const double *rows, const int *nrows,
                                 int variant, const nmpc_opts *opts, double *z_all, int *info_int_all,
                                 double *info_real_local, int mixed, void *cuda_stream);
int nmpc_solve_batch_sharded_f32(nmpc_comm *comm, int B_local, int N, int mcap, const float *xinit,
                                 const float *z0, const float *hdr, const float *rows, const int *nrows,
                                 int variant, const nmpc_opts *opts, float *z_all, int *info_int_all,
                                 float *info_real_local, void *cuda_stream);
int nmpc_collate_inplace(nmpc_comm *comm, void *buf_all, size_t bytes_per_rank, void *cuda_stream);

/* ---- multi-GPU, fused: collation by peer stores from inside the solve kernel (GPUs of one node, NVLink / NVSwitch) --
 * Every rank owns one allocation [header | z_all | info_int_all] (nmpc_peers_create) and maps the allocation of every
 * other rank (CUDA IPC: nmpc_peers_export -> the application exchanges the 64-byte handles the way it exchanges the
 * NCCL id -> nmpc_peers_connect).  The epilogue of the solve kernel then writes each problem's solution and info
 * integers into its slice of EVERY rank's buffers (TMA bulk stores to peer addresses), so the exchange is spread over
 * the whole kernel and overlaps the arithmetic; what remains of the collation is a barrier kernel on the same stream
 * (release/acquire flags in the peers' headers, epoch kept on the device: CUDA-graph safe).
 *   nmpc_solve_batch_sharded_p2p_f64 / _f32
 *       barrier (every rank is done reading the previous batch) -> solve with peer stores -> barrier (all results
 *       have landed everywhere).  mode: 0 fp64 kernel, 1 mixed precision, 2 low-latency warp-group kernel; _f32 is
 *       always mixed.  Results: nmpc_peers_z() [world * B_local][N][17], nmpc_peers_info() [world * B_local][4] on
 *       every rank.  Every rank must call with the same B_local (even) the same number of times.
 *   nmpc_peers_status   synchronous; non-zero if a barrier timed out (a rank missing for 2 s)
 *   nmpc_peers_destroy  the application synchronises the ranks first (no peer may still be writing)
 * No NCCL involved.  Across nodes, or without P2P access, use nmpc_solve_batch_sharded_* above.              */
typedef struct nmpc_peers nmpc_peers;
int nmpc_peers_create(int world, int rank, size_t z_bytes_all, size_t info_ints_all, nmpc_peers **peers);
int nmpc_peers_export(nmpc_peers *peers, unsigned char handle[64]);
int nmpc_peers_connect(nmpc_peers *peers, const unsigned char *handles /* [world][64], own entry ignored */);
void *nmpc_peers_z(nmpc_peers *peers);
int *nmpc_peers_info(nmpc_peers *peers);
int nmpc_peers_rank(const nmpc_peers *peers);
int nmpc_peers_world(const nmpc_peers *peers);
int nmpc_peers_barrier(nmpc_peers *peers, void *cuda_stream);
int nmpc_peers_status(nmpc_peers *peers);
int nmpc_peers_destroy(nmpc_peers *peers);
int nmpc_solve_batch_sharded_p2p_f64(nmpc_peers *peers, int B_local, int N, int mcap, const double *xinit,
                                     const double *z0, const double *hdr, const double *rows, const int *nrows,
                                     int variant, const nmpc_opts *opts, double *info_real_local, int mode,
                                     void *cuda_stream);
int nmpc_solve_batch_sharded_p2p_f32(nmpc_peers *peers, int B_local, int N, int mcap, const float *xinit,
                                     const float *z0, const float *hdr, const float *rows, const int *nrows,
                                     int variant, const nmpc_opts *opts, float *info_real_local, void *cuda_stream);

/* ---- stand-alone structured KKT factorisation / backsolve (device pointers) ------------------
 * The split the reference binary makes internally (f_17_PD_ldlchol_rowmajor ... vs
 * f_17_ldl_forward_solve_rm / f_13_backward_solve_rm) for the Riccati factor:
 *   phi [B][N][21]  stage Hessian, compact: diag(17) | pos block off-diag (01,02,12) | H[u_i][uprev_i]
 *   jc  [B][N][51]  compact dynamics Jacobian (pos+/vel+ wrt vel, rpy, thrust; vel+ wrt rates)
 *   fac [B][N * nmpc_backsolve_factor_words()]  opaque to the caller; per problem two regions so that each can be
 *       fetched on its own: [P: N x 91, symmetric 13x13 in a bank-conflict-free packed layout]
 *       [N x 113: K_k 52 | Quu^-1 packed 10 | J_k 51].  All pointers 16-byte aligned.
 *   g   [B][N][17], d [B][N][13] (c-ordering, row N-1 unused)  ->  dz [B][N][17], y [B][N][13]
 * solving   min 1/2 dz'Phi dz + g'dz  s.t.  E dz_{k+1} = J_k dz_k + d_k,  dz_0[8:17] = 0.         */
int nmpc_backsolve_factor_words(void);
long nmpc_backsolve_algorithmic_bytes(int N, int elem_size);
int nmpc_riccati_factor_f64(int B, int N, const double *phi, const double *jc, double *fac,
                            int *status, void *cuda_stream);
int nmpc_riccati_factor_f32(int B, int N, const float *phi, const float *jc, float *fac,
                            int *status, void *cuda_stream);
int nmpc_kkt_backsolve_f64(int B, int N, const double *fac, const double *g, const double *d,
                           double *dz, double *y, void *cuda_stream);
int nmpc_kkt_backsolve_f32(int B, int N, const float *fac, const float *g, const float *d,
                           float *dz, float *y, void *cuda_stream);

/* ---- the steps around the solve, device-resident ----------------------
 * nmpc_pack_params_f64: batched body of FORCESNormal::solveNormal (forces_normal.cpp:100-136):
 *   ref_pos [B][N][3], ref_yaw [B][N], ext_acc [B][3], ellipsoid [B][N][9] (E_i row-major),
 *   poly_A [B][P][M][3], poly_b [B][P][M], poly_m [B][P], poly_idx [B][N],
 *   weights5 (HOST) = w_stage_wp, w_stage_input, w_input_rate, w_terminal_wp, w_terminal_input
 *   -> hdr [B][N][10], rows [B][N][mcap][4] with b_j - ||E_i a_j||, nrows [B][N] (truncated at mcap)
 * nmpc_shift_warm_start_f64: z_prev [B][N][17] -> z0 (shifted, last stage duplicated), xinit [B][9]
 *   (forces_normal.cpp:62-97, nmpc_solver.cpp:531-543); wrap_yaw != 0 wraps yaw into (-pi, pi].   */
int nmpc_pack_params_f64(int B, int N, int P, int M, int mcap, const double *ref_pos,
                         const double *ref_yaw, const double *ext_acc, const double *ellipsoid,
                         const double *poly_A, const double *poly_b, const int *poly_m,
                         const int *poly_idx, const double *weights5, double *hdr, double *rows,
                         int *nrows, void *cuda_stream);
int nmpc_shift_warm_start_f64(int B, int N, const double *z_prev, double *xinit, double *z0,
                              int wrap_yaw, void *cuda_stream);
/* NMPCSolver::updateFORCESResults' yaw wrap (nmpc_solver.cpp:531-541) on an adopted plan z [B][N][17], in place:
 * yaw < -PI -> yaw + 2 PI, yaw > PI -> yaw - 2 PI, with the reference's PI = 3.1415926 (:3).               */
int nmpc_wrap_yaw_f64(int B, int N, double *z, void *cuda_stream);

/* ---- result handling of a replan, device-resident (NMPCSolver::solveNMPC, nmpc_solver.cpp:398-427, 363-364) ----
 * z_prev [B][N][17] is the plan in force (mpc_output_).  An agent whose new solve is accepted -- info_int[b][0] == 1, or
 * accept[b] != 0 when the caller supplies its own mask (the reference also tolerates exit flag 0 after more than three
 * replans: host policy, host/forces_wrappers.hpp::SolveAcceptance) -- adopts z_new (yaw wrapped into (-PI, PI] when
 * wrap_yaw != 0, updateFORCESResults :531-541).  A rejected agent takes nothing from the failed solve (its output may
 * be NaN): its plan becomes the cold guess of initMPCOutput (:265-286) at its current state -- odom[b] (9 doubles) when
 * given, else stage 2 of the old plan (where that plan puts the vehicle at the next cycle) -- so that the next shift
 * + solve is a cold start, as in the reference.  cold [B] (may be NULL) receives 1 for the rejected agents.
 * accept, odom may be NULL; info_int may be NULL only if accept is given.                                   */
int nmpc_adopt_plans_f64(int B, int N, const double *z_new, const int *info_int, const int *accept,
                         const double *odom, double *z_prev, int *cold, int wrap_yaw, void *cuda_stream);
/* order [B] <- agent indices sorted by the previous solve's iteration count, longest first (failed agents, which
 * restart cold, first; ties by index): the launch order for nmpc_solve_batch_ordered_f64 in a receding-horizon
 * stream.  Rank by counting out of shared memory (every CTA ranks 128 agents against all keys); B <= 12288.                                      */
int nmpc_rank_longest_first(int B, const int *info_int, int *order, void *cuda_stream);

/* ---- reference sampling + yaw reference, device-resident ---------------
 * Batched NMPCSolver::getCurTraj (plan_manage/src/nmpc_solver.cpp:109-142) + calculate_yaw
 * (:834-862), as called once per stage by setFORCESParams (:484-493):
 *   kino_path [B][P][3] the front-end polyline sampled at Ts (kino_path_, :215), kino_size [B] its
 *   live length (>= 1), t_off [B] = mpc_start_time_ - kino_start_time_ in seconds, last_yaw [B] the
 *   yaw of the previous plan's stage 1 (:486), pos1 [B][3] its position (may be NULL)
 *   -> ref_pos [B][N][3] (interpolated; clamped to the last point), ref_yaw [B][N] (heading towards
 *      the point five samples on, unwrapped against the running value, 0.2/0.8 low-pass),
 *      hard_to_follow [B] (may be NULL): 1 when stage 0's reference is > 1 m from pos1 (:136-140).
 * All device pointers.                                                                          */
int nmpc_sample_reference_f64(int B, int N, int P, double Ts, const double *kino_path,
                              const int *kino_size, const double *t_off, const double *last_yaw,
                              const double *pos1, double *ref_pos, double *ref_yaw,
                              int *hard_to_follow, void *cuda_stream);

/* ---- disturbance-ellipsoid propagation, device-resident ----------------
 * Batched form of the ellipsoid part of NMPCSolver::setFORCESParams (plan_manage/src/nmpc_solver.cpp:
 * 484-521) with updateMatrix (:615-699), eulerToRot (:552-564), getDistrEllipsoid (:567-611) and the
 * 3x3 matrix square root (:511-512):
 *   z [B][N][17] the previous plan (mpc_output_)  ->  ellipsoid [B][N][9], the shape matrices E_i
 *   (ellipsoid_matrices_, row-major) that nmpc_pack_params_f64 takes.  Device pointers.
 * consts == NULL selects nmpc_default_ellipsoid_consts (rotors_sim.launch:53-70, nmpc_utils.h:188).
 * Deviation from the reference: its accumulator `temp` (:573) is uninitialised; here it starts at 0. */
typedef struct nmpc_ellipsoid_consts {
    double mass, drag, ego_r, ego_h, ext_noise_bound, epsilon, Ts;
} nmpc_ellipsoid_consts;
void nmpc_default_ellipsoid_consts(nmpc_ellipsoid_consts *c);
int nmpc_propagate_ellipsoids_f64(int B, int N, const double *z, const nmpc_ellipsoid_consts *consts,
                                  double *ellipsoid, void *cuda_stream);

/* ---- corridor generation + per-stage polytope selection, device-resident --
 * Batched form of the poly_indices / poly_constraints_ part of NMPCSolver::setFORCESParams with
 * getSikangConst (plan_manage/src/nmpc_solver.cpp:288-332, 493-516) and of the DecompROS routines it
 * drives (ThirdParty/DecompROS/decomp_ros_utils/include/decomp_util/{ellipsoid_decomp,line_segment,
 * decomp_base}.h, decomp_geometry/{ellipsoid,polyhedron}.h): per agent, walk the stages; keep the last
 * polytope while the reference point inflated by 1.1 ||E_i a_j|| is inside it, otherwise dilate a new
 * one (ellipsoid fit + half-space carving over the obstacle cloud + local box) around the 0.1 m seed
 * segment along the yaw reference.
 *   cloud [B][M][3] obstacle points per agent (cloud_stride = doubles between agents, 3*M; 0 = one cloud
 *   shared by all, then cloud_n has one entry), cloud_n live points, ref_pos [B][N][3], ref_yaw [B][N],
 *   ellipsoid [B][N][9] (from nmpc_propagate_ellipsoids_f64), bbox3 (HOST, NULL = {2, 2, 1}, :323)
 *   -> poly_A [B][P][R][3], poly_b [B][P][R], poly_m [B][P], poly_idx [B][N]  (the inputs of
 *      nmpc_pack_params_f64 with M := R), n_poly [B], overflow [B] (bit 0: a polytope had more than R
 *      rows, bit 1: more than P polytopes were needed; 0 = exact).  R >= 7.  Device pointers.        */
int nmpc_select_corridors_f64(int B, int N, int M, int P, int R, const double *cloud,
                              long long cloud_stride, const int *cloud_n, const double *ref_pos,
                              const double *ref_yaw, const double *ellipsoid, const double *bbox3,
                              double *poly_A, double *poly_b, int *poly_m, int *poly_idx, int *n_poly,
                              int *overflow, void *cuda_stream);

/* ---- measured CUDA-core FMA peak (TFLOP/s) of the current device, elem_size 8 (fp64) or 4 (fp32):
 * the roofline denominator for the fused solver kernel, which is FMA-issue/latency bound, not
 * HBM bound (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).                       */
int nmpc_fma_peak_probe(int elem_size, double *tflops);

/* ---- device model check: one evaluation of the reference callback per (problem, stage) -------
 * Mirrors FORCESNLPsolver_normal_casadi2forces (solver/normal/FORCESNLPsolver_normal_casadi2forces.c:42-55)
 * for n independent (z[17], p[130], stage) triples in HOST memory; dense COLUMN-major outputs:
 * f[n], grad[n][17], c[n][13], jc[n][13*17], h[n][30], jh[n][30*17].                            */
int nmpc_model_eval_host_f64(int n, const double *z, const double *p, const int *stage, int n_stages,
                             int variant, double *f, double *grad, double *c, double *jc,
                             double *h, double *jh);

#ifdef __cplusplus
}
#endif
#endif
