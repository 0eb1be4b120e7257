// forces_wrappers.hpp -- C++ host side above the C ABI: the reference's solver wrapper classes.
//
// Mirrors resilient_planner::FORCESNormal / FORCESFinal
//   (/root/reference/src/resilient_planner/plan_manage/include/plan_manage/nmpc_utils.h:49-106,
//    src/forces_normal.cpp:36-168, src/forces_final.cpp)
// with the same member names, argument meaning and return convention.  The reference versions need
// Eigen and DecompROS types (neither is installed in this image), so the vector / matrix /
// polytope types are minimal std:: stand-ins with the same accessors the packing loop uses; in the
// planner itself the original forces_normal.cpp compiles unchanged against include/*.h and links
// libnmpc_b200.so -- this header exists so that the flow can be built and tested
// here, and as the batched C++ entry (BatchedNMPC) the reference does not have.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/FORCESNLPsolver_final.h"
#include "../../include/FORCESNLPsolver_normal.h"
#include "../../include/nmpc_b200.h"

namespace resilient_planner {

using Vec3 = std::array<double, 3>;
using Mat3 = std::array<double, 9>;                 // row-major
using StageVec = std::array<double, 17>;
using MPCDeque = std::deque<StageVec>;              // reference: std::deque<Eigen::VectorXd>

struct LinearConstraint3D {                         // decomp_geometry/polyhedron.h:99-149
    std::vector<Vec3> A_;
    std::vector<double> b_;
    const std::vector<Vec3>& A() const { return A_; }
    const std::vector<double>& b() const { return b_; }
};

struct FORCESParams {                               // nmpc_utils.h:49-58
    int num_pre_params = 10;
    int num_const = 30;
    int num_iter = 130;
    int num_var = 17;
    int planning_horizon = 20;
};

namespace detail {
template <class Params>
inline void set_weights(Params& p, const FORCESParams& v, double w_stage_wp, double w_stage_input,
                        double w_input_rate, double w_terminal_wp, double w_terminal_input)
{
    for (int i = 0; i < v.planning_horizon; i++) {
        p.all_parameters[i * v.num_iter + 6] = w_stage_wp;
        p.all_parameters[i * v.num_iter + 7] = w_stage_input;
        p.all_parameters[i * v.num_iter + 8] = w_input_rate;
    }
    p.all_parameters[(v.planning_horizon - 1) * v.num_iter + 6] = w_terminal_wp;
    p.all_parameters[(v.planning_horizon - 1) * v.num_iter + 7] = w_terminal_input;
}

// forces_normal.cpp:62-136: predicted-state xinit, shifted warm start, per-stage parameters with
// corridor tightening b_j - ||E_i a_j||_2 and zero padding to 30 rows (extra rows dropped).
template <class Params>
inline void pack(Params& p, const FORCESParams& v, const MPCDeque& mpc_output, const Vec3& external_acc,
                 const std::vector<Vec3>& ref_total_pos, const std::vector<double>& ref_total_yaw,
                 const std::vector<Mat3>& ellipsoid_matrices, const std::vector<LinearConstraint3D>& poly_constraints,
                 const std::vector<int>& poly_indices)
{
    for (int j = 0; j < 9; j++) p.xinit[j] = mpc_output.at(1)[8 + j];
    for (int i = 0; i < v.planning_horizon; i++) {
        for (int j = 0; j < v.num_var; j++) p.x0[i * v.num_var + j] = mpc_output.at(i + 1)[j];
        double* q = p.all_parameters + i * v.num_iter;
        for (int j = 0; j < 3; j++) { q[j] = ref_total_pos.at(i)[j]; q[3 + j] = external_acc[j]; }
        q[9] = ref_total_yaw.at(i);
        const LinearConstraint3D& lc = poly_constraints.at(poly_indices.at(i));
        const Mat3& E = ellipsoid_matrices.at(i);
        for (int j = 0; j < v.num_const; ++j) {
            double* a = q + v.num_pre_params + j * 3;
            double* b = q + v.num_pre_params + v.num_const * 3 + j;
            if (j < (int)lc.b().size()) {
                const Vec3& r = lc.A()[j];
                a[0] = r[0]; a[1] = r[1]; a[2] = r[2];
                const double e0 = E[0] * r[0] + E[1] * r[1] + E[2] * r[2];
                const double e1 = E[3] * r[0] + E[4] * r[1] + E[5] * r[2];
                const double e2 = E[6] * r[0] + E[7] * r[1] + E[8] * r[2];
                *b = lc.b()[j] - std::sqrt(e0 * e0 + e1 * e1 + e2 * e2);
            } else {
                a[0] = a[1] = a[2] = 0.0;
                *b = 0.0;
            }
        }
    }
}

template <class Output>
inline void unpack(const Output& o, const FORCESParams& v, MPCDeque& mpc_output)
{
    const double* x = o.x01;   // x01 .. x20 are contiguous (2720 B, static-asserted in nmpc_capi.cu)
    for (int k = 0; k < v.planning_horizon; k++)
        for (int j = 0; j < v.num_var; j++) mpc_output.at(k)[j] = x[k * v.num_var + j];
}
}  // namespace detail

class FORCESNormal {
public:
    FORCESNormal() : extfunc_eval_(nullptr) { params_.num_of_threads = 1; }
    FORCESNLPsolver_normal_extfunc extfunc_eval_;   // accepted and ignored by the CUDA solver
    FORCESNLPsolver_normal_output output_{};
    FORCESNLPsolver_normal_params params_{};
    FORCESNLPsolver_normal_info info_{};
    FORCESParams value_;

    void setParasNormal(double w_stage_wp, double w_stage_input, double w_input_rate, double w_terminal_wp,
                        double w_terminal_input)
    {
        detail::set_weights(params_, value_, w_stage_wp, w_stage_input, w_input_rate, w_terminal_wp, w_terminal_input);
    }
    int solveNormal(MPCDeque& mpc_output, Vec3& external_acc, std::vector<Vec3>& ref_total_pos,
                    std::vector<double>& ref_total_yaw, std::vector<Mat3>& ellipsoid_matrices,
                    std::vector<LinearConstraint3D>& poly_constraints, std::vector<int>& poly_indices)
    {
        detail::pack(params_, value_, mpc_output, external_acc, ref_total_pos, ref_total_yaw, ellipsoid_matrices,
                     poly_constraints, poly_indices);
        return FORCESNLPsolver_normal_solve(&params_, &output_, &info_, NULL, extfunc_eval_);
    }
    void updateNormal(MPCDeque& mpc_output) { detail::unpack(output_, value_, mpc_output); }
};

class FORCESFinal {
public:
    FORCESFinal() : extfunc_eval_final_(nullptr) { params_final_.num_of_threads = 1; }
    FORCESNLPsolver_final_extfunc extfunc_eval_final_;
    FORCESNLPsolver_final_output output_final_{};
    FORCESNLPsolver_final_params params_final_{};
    FORCESNLPsolver_final_info info_final_{};
    FORCESParams value_final_;

    void setParasFinal(double w_final_stage_wp, double w_final_stage_input, double w_input_rate,
                       double w_final_terminal_wp, double w_final_terminal_input)
    {
        detail::set_weights(params_final_, value_final_, w_final_stage_wp, w_final_stage_input, w_input_rate,
                            w_final_terminal_wp, w_final_terminal_input);
    }
    int solveFinal(MPCDeque& mpc_output, Vec3& external_acc, std::vector<Vec3>& ref_total_pos,
                   std::vector<double>& ref_total_yaw, std::vector<Mat3>& ellipsoid_matrices,
                   std::vector<LinearConstraint3D>& poly_constraints, std::vector<int>& poly_indices)
    {
        detail::pack(params_final_, value_final_, mpc_output, external_acc, ref_total_pos, ref_total_yaw,
                     ellipsoid_matrices, poly_constraints, poly_indices);
        return FORCESNLPsolver_final_solve(&params_final_, &output_final_, &info_final_, NULL, extfunc_eval_final_);
    }
    void updateFinal(MPCDeque& mpc_output) { detail::unpack(output_final_, value_final_, mpc_output); }
};

// Exit-code acceptance policy of NMPCSolver::solveNMPC (plan_manage/src/nmpc_solver.cpp:398-421):
// only exit code 1 is a success; a maxit exit (0) is tolerated once more than three replans have been
// forced; the third consecutive failure forces a front-end replan.  The next solve starts from the
// cold guess whenever the last exit code was not 1 (:363-364).
struct SolveAcceptance {
    int fail_count = 0, replan_count = 0, last_exit_code = 1;
    bool kino_replan = false;
    bool consume(int exit_code)   // returns update_result: the planner adopts this solve's output
    {
        bool update_result = false;
        last_exit_code = exit_code;
        if (exit_code == 1) {
            fail_count = 0; replan_count = 0; update_result = true;
        } else {
            fail_count += 1;
            if (replan_count > 3 && exit_code == 0) {
                fail_count = 0; replan_count = 0; update_result = true;
            } else if (fail_count > 2) {
                fail_count = 0; replan_count += 1; kino_replan = true;
            }
        }
        return update_result;
    }
    bool next_solve_is_cold(bool initialized_output) const { return !initialized_output || last_exit_code != 1; }
};

// Batched counterpart (new: the reference plans for one vehicle): host-pointer solve of B problems.
class BatchedNMPC {
public:
    BatchedNMPC(int B, int N, int mcap, int variant = 0)
        : B_(B), N_(N), mcap_(mcap), variant_(variant), xinit((size_t)B * 9), z0((size_t)B * N * 17),
          hdr((size_t)B * N * 10), rows((size_t)B * N * mcap * 4), nrows((size_t)B * N), z((size_t)B * N * 17),
          info_int((size_t)B * 4), info_real((size_t)B * 8)
    {
        nmpc_default_opts(&opts);
    }
    void solve()
    {
        const int rc = nmpc_solve_batch_host_f64(B_, N_, mcap_, xinit.data(), z0.data(), hdr.data(), rows.data(),
                                                 nrows.data(), variant_, &opts, z.data(), info_int.data(),
                                                 info_real.data());
        if (rc != 0) throw std::runtime_error(std::string("nmpc_solve_batch_host_f64: ") + nmpc_last_error());
    }
    int exitflag(int b) const { return info_int[(size_t)b * 4]; }
    int iterations(int b) const { return info_int[(size_t)b * 4 + 1]; }
    nmpc_opts opts;
    std::vector<double> xinit, z0, hdr, rows;
    std::vector<int> nrows;
    std::vector<double> z;
    std::vector<int> info_int;
    std::vector<double> info_real;

private:
    int B_, N_, mcap_, variant_;
};

}  // namespace resilient_planner
