// dropin_demo.cpp -- the planner's call sequence (nmpc_solver.cpp:265-286, 384, 421) against
// libnmpc_b200.so: initMPCOutput -> setParasNormal -> solveNormal -> updateNormal, three
// receding-horizon replans, then the same through FORCESFinal.  Prints one line per solve;
// exit status 0 iff every solve returned exitflag 1.   Build: see host/Makefile.
#include <cstdio>

#include "forces_wrappers.hpp"

using namespace resilient_planner;

int main()
{
    FORCESNormal normal;
    FORCESFinal fin;
    normal.setParasNormal(7.0, 1.0, 80.0, 12.0, 0.5);          // launch/rotors_sim.launch:56-66
    fin.setParasFinal(12.0, 1.5, 80.0, 15.0, 0.5);
    // initMPCOutput: hover-ish guess replicated 21 times
    StageVec row{0, 0, 0, 7.3, 0, 0, 0, 7.3, 0, 0, 1.0, 0, 0, 0, 0, 0, 0};
    MPCDeque mpc_output(21, row);
    Vec3 external_acc{0.3, -0.2, 0.0};
    std::vector<Mat3> E(20, Mat3{0.27, 0, 0, 0, 0.27, 0, 0, 0, 0.0425});
    LinearConstraint3D box;
    box.A_ = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    box.b_ = {3.0, 2.0, 2.0, 2.0, 2.0, 0.0};
    std::vector<LinearConstraint3D> polys{box};
    std::vector<int> poly_indices(20, 0);
    int bad = 0;
    for (int replan = 0; replan < 3; replan++) {
        std::vector<Vec3> ref(20);
        std::vector<double> yaw(20, 0.0);
        for (int i = 0; i < 20; i++) ref[i] = {0.05 * (i + 1 + replan), 0.0, 1.0};
        const int flag = normal.solveNormal(mpc_output, external_acc, ref, yaw, E, polys, poly_indices);
        std::printf("normal replan %d: exitflag %d it %d pobj %.6f rsnorm %.2e solvetime %.3f ms\n", replan, flag,
                    normal.info_.it, normal.info_.pobj, normal.info_.rsnorm, 1e3 * normal.info_.solvetime);
        bad += (flag != 1);
        if (flag == 1) {
            normal.updateNormal(mpc_output);
            mpc_output.at(20) = mpc_output.at(19);               // nmpc_solver.cpp:543
        }
    }
    {
        std::vector<Vec3> ref(20, Vec3{0.3, 0.0, 1.0});
        std::vector<double> yaw(20, 0.0);
        const int flag = fin.solveFinal(mpc_output, external_acc, ref, yaw, E, polys, poly_indices);
        std::printf("final: exitflag %d it %d pobj %.6f solvetime %.3f ms\n", flag, fin.info_final_.it,
                    fin.info_final_.pobj, 1e3 * fin.info_final_.solvetime);
        bad += (flag != 1);
        if (flag == 1) fin.updateFinal(mpc_output);
    }
    std::printf("x after final: pos (%.4f %.4f %.4f)\n", mpc_output.at(19)[8], mpc_output.at(19)[9], mpc_output.at(19)[10]);
    return bad;
}
