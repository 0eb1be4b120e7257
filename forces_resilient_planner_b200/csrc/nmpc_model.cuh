// nmpc_model.cuh -- device-side model layer of the NMPC problem (sm_100a).
//
// Replaces, on the device, the reference's per-stage model callback
//   FORCESNLPsolver_{normal,final}_casadi2forces
//   (/root/reference/src/resilient_planner/plan_manage/solver/normal/FORCESNLPsolver_normal_casadi2forces.c:42-245)
// and the CasADi functions it dispatches to (..._casadi.c: objective_{1,2,20}, dynamics_{1,2},
// inequalities_{1,2,20}).  The maths is taken from the MATLAB problem definition
// (matlab_code/setup.m:17-43, dynamics/nonlinear_dynamics.m:21-40, dynamics/transit.m:4-8,
// mpc/mpc_objective1.m:38-48, mpc/final/mpc_objectiveN_final.m:26); ForcesPro's "RK2" is Heun
// (..._casadi.c:238-240,307-311,383-394).  Instead of CasADi's 116 trig evaluations and dense
// 13x17 scatter, one stage costs 12 trig evaluations and the Jacobian is kept in a 51-word
// compact form (its fixed sparsity is 64 nz of which 13 are the constants 1 and h).
#pragma once
#include <cuda_runtime.h>

namespace nmpc {

constexpr int NZ = 17;    // stage vector [u(4) | u_prev(4) | pos vel rpy (9)]
constexpr int NXI = 13;   // equalities per transition, c-ordering [x+(9); u(4)]
constexpr int NJC = 51;   // compact Jacobian words per stage
constexpr int FAC_WORDS = 204;   // stored Riccati factor per stage: P 91 | K 52 | Quu^-1 10 | J 51

// packed-lower index of a symmetric matrix
__host__ __device__ __forceinline__ constexpr int pk(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// Bank-conflict-free packed layout of the symmetric 13x13 cost-to-go matrices in the stored factor:
// PSYM[i][j] = PSYM[j][i] = word (0..90) holding P[i][j].  For every column j the 13 words
// {PSYM[i][j]} are distinct mod 16, so a row-per-lane product y = P x reads 13 different 8-byte banks
// per step (a proper edge colouring of K13 with loops by the 16 banks; scripts/psym_layout.py).
// Rows padded to 16 bytes so that a lane fetches its row with one 16-byte load.
__device__ const unsigned char PSYM[13][16] = {
    { 5,  3, 15,  2,  6, 17, 25, 23, 40, 48, 74, 68, 75, 0, 0, 0},
    { 3, 21,  9, 10,  7, 28, 33, 24, 32, 46, 45, 63, 86, 0, 0, 0},
    {15,  9, 19, 12, 11,  0, 42, 14, 54, 34, 55, 72, 84, 0, 0, 0},
    { 2, 10, 12, 13,  1, 27, 20, 38, 57, 39, 64, 69, 78, 0, 0, 0},
    { 6,  7, 11,  1, 26, 37,  8, 44, 30, 73, 47, 61, 82, 0, 0, 0},
    {17, 28,  0, 27, 37,  4, 22, 58, 18, 29, 89, 67, 87, 0, 0, 0},
    {25, 33, 42, 20,  8, 22, 16, 31, 60, 53, 59, 66, 83, 0, 0, 0},
    {23, 24, 14, 38, 44, 58, 31, 41, 36, 35, 50, 81, 77, 0, 0, 0},
    {40, 32, 54, 57, 30, 18, 60, 36, 49, 43, 51, 71, 79, 0, 0, 0},
    {48, 46, 34, 39, 73, 29, 53, 35, 43, 52, 65, 76, 88, 0, 0, 0},
    {74, 45, 55, 64, 47, 89, 59, 50, 51, 65, 56, 70, 85, 0, 0, 0},
    {68, 63, 72, 69, 61, 67, 66, 81, 71, 76, 70, 62, 90, 0, 0, 0},
    {75, 86, 84, 78, 82, 87, 83, 77, 79, 88, 85, 90, 80, 0, 0, 0},
};

// compact Jacobian layout (row-major 3x3 blocks unless noted)
constexpr int JPV = 0;    // d pos+ / d vel
constexpr int JPR = 9;    // d pos+ / d rpy
constexpr int JPT = 18;   // d pos+ / d thrust   (3)
constexpr int JVV = 21;   // d vel+ / d vel
constexpr int JVR = 30;   // d vel+ / d rpy
constexpr int JVT = 39;   // d vel+ / d thrust   (3)
constexpr int JVW = 42;   // d vel+ / d body rates

template <typename T> struct Const {
    static constexpr T h = T(0.05);
    static constexpr T mass = T(0.745319);
    static constexpr T grav = T(9.81);
    static constexpr T kd = T(0.33);
    static constexpr T hu = T(1e-5);
    static constexpr T pi = T(3.14159265358979323846);
    static constexpr T rate_max = T(3.14159265358979323846 / 2);
    static constexpr T inv_rate2 = T(1.0 / ((3.14159265358979323846 / 2) * (3.14159265358979323846 / 2)));
    static constexpr T thrust_lo = T(0.5 * 9.81 * 0.745319);
    static constexpr T thrust_hi = T(2.0 * 9.81 * 0.745319);
};

// variable bounds (matlab_code/mpc/normal/mpc_generator_normal.m:29-46)
template <typename T> __device__ __forceinline__ T lower_bound(int i)
{
    using C = Const<T>;
    switch (i) {
    case 0: case 1: case 2: case 4: case 5: case 6: return -C::rate_max;
    case 3: case 7: return C::thrust_lo;
    case 8: case 9: return T(-20);
    case 10: return T(0);
    case 11: case 12: case 13: return T(-2);
    case 14: case 15: return T(-0.4) * C::pi;
    default: return T(-2) * C::pi;
    }
}
template <typename T> __device__ __forceinline__ T upper_bound(int i)
{
    using C = Const<T>;
    switch (i) {
    case 0: case 1: case 2: case 4: case 5: case 6: return C::rate_max;
    case 3: case 7: return C::thrust_hi;
    case 8: case 9: return T(20);
    case 10: return T(5);
    case 11: case 12: case 13: return T(2);
    case 14: case 15: return T(0.4) * C::pi;
    default: return T(2) * C::pi;
    }
}

__device__ __forceinline__ void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ void sincos_t(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ double log_t(double x) { return log(x); }
__device__ __forceinline__ float log_t(float x) { return logf(x); }

// sin / cos of the three Euler angles, packed (sr cr sp cp sy cy)
template <typename T> __device__ __forceinline__ void trig3(const T r[3], T sc[6])
{
    sincos_t(r[0], &sc[0], &sc[1]);
    sincos_t(r[1], &sc[2], &sc[3]);
    sincos_t(r[2], &sc[4], &sc[5]);
}
// The same for r + dr with |dr| small, from the values at r by the addition theorems: the Heun step evaluates the
// attitude a second time at rpy + h * rates, and |h * rate| <= 0.05 * pi/2 = 0.079 for every point inside the rate bounds
// (mpc_generator_normal.m:33-35).  Taylor polynomials of sin / cos through x^9 / x^10: truncation < 1e-19 there and
// still < 5e-14 at |dr| = 0.3 (four times the bound) -- 3 x ~20 flops instead of three more sincos calls.
template <typename T> __device__ __forceinline__ void trig3_shifted(const T sc[6], const T dr[3], T out[6])
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const T x = dr[i], x2 = x * x;
        const T sd = x * (T(1) + x2 * (T(-1.0 / 6) + x2 * (T(1.0 / 120) + x2 * (T(-1.0 / 5040) + x2 * T(1.0 / 362880)))));
        const T cd = T(1) + x2 * (T(-0.5) + x2 * (T(1.0 / 24) + x2 * (T(-1.0 / 720) + x2 * (T(1.0 / 40320) + x2 * T(-1.0 / 3628800)))));
        out[2 * i] = sc[2 * i] * cd + sc[2 * i + 1] * sd;
        out[2 * i + 1] = sc[2 * i + 1] * cd - sc[2 * i] * sd;
    }
}

// acc = z_B T/m + f_ext - g e3 - kd (v - z_B (z_B.v));  R diag(kd,kd,0) R' = kd (I - z_B z_B')
// sc = (sin, cos) of roll, pitch, yaw.  JAC: also Av = da/dv, Ar = da/drpy (row-major 3x3), AT = da/dT.
template <typename T, bool JAC>
__device__ __forceinline__ void accel(const T v[3], const T sc[6], T thrust, const T fe[3], T a[3],
                                      T Av[9], T Ar[9], T AT[3])
{
    using C = Const<T>;
    const T sr = sc[0], cr = sc[1], sp = sc[2], cp = sc[3], sy = sc[4], cy = sc[5];
    T zb[3] = {cy * sp * cr + sy * sr, sy * sp * cr - cy * sr, cp * cr};
    T zv = zb[0] * v[0] + zb[1] * v[1] + zb[2] * v[2];
    T tm = thrust * (T(1) / C::mass);
#pragma unroll
    for (int i = 0; i < 3; i++) a[i] = zb[i] * tm + fe[i] - C::kd * (v[i] - zb[i] * zv);
    a[2] -= C::grav;
    if (JAC) {
        T Z[9] = {-cy * sp * sr + sy * cr, cy * cp * cr, -sy * sp * cr + cy * sr,
                  -sy * sp * sr - cy * cr, sy * cp * cr, cy * sp * cr + sy * sr,
                  -cp * sr, -sp * cr, T(0)};
        T vZ[3];
#pragma unroll
        for (int j = 0; j < 3; j++) vZ[j] = v[0] * Z[j] + v[1] * Z[3 + j] + v[2] * Z[6 + j];
        T w = tm + C::kd * zv;
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                Av[3 * i + j] = -C::kd * ((i == j ? T(1) : T(0)) - zb[i] * zb[j]);
                Ar[3 * i + j] = w * Z[3 * i + j] + C::kd * zb[i] * vZ[j];
            }
            AT[i] = zb[i] * (T(1) / C::mass);
        }
    }
}

// c(z) = [Heun_h(x,u;f_ext) (9) ; u (4)]; JAC: compact Jacobian jc[51].
template <typename T, bool JAC>
__device__ __forceinline__ void dynamics(const T z[NZ], const T fe[3], T c[NXI], T jc[NJC])
{
    using C = Const<T>;
    const T h = C::h, hh = T(0.5) * C::h * C::h, h2 = T(0.5) * C::h;
    const T* w = z;
    const T* p = z + 8;
    const T* v = z + 11;
    const T* r = z + 14;
    T a1[3], A1v[9], A1r[9], A1T[3], a2[3], A2v[9], A2r[9], A2T[3];
    T sc1[6], sc2[6];
    trig3<T>(r, sc1);
    accel<T, JAC>(v, sc1, z[3], fe, a1, A1v, A1r, A1T);
    T v2[3], dr[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        v2[i] = v[i] + h * a1[i];
        dr[i] = h * w[i];
    }
    trig3_shifted<T>(sc1, dr, sc2);
    accel<T, JAC>(v2, sc2, z[3], fe, a2, A2v, A2r, A2T);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        c[i] = p[i] + h * v[i] + hh * a1[i];
        c[3 + i] = v[i] + h2 * (a1[i] + a2[i]);
        c[6 + i] = r[i] + h * w[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) c[9 + i] = z[i];
    if (JAC) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                T dv = T(0), dr2 = A2r[3 * i + j];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    dv += A2v[3 * i + k] * ((k == j ? T(1) : T(0)) + h * A1v[3 * k + j]);
                    dr2 += A2v[3 * i + k] * h * A1r[3 * k + j];
                }
                jc[JPV + 3 * i + j] = (i == j ? h : T(0)) + hh * A1v[3 * i + j];
                jc[JPR + 3 * i + j] = hh * A1r[3 * i + j];
                jc[JVV + 3 * i + j] = (i == j ? T(1) : T(0)) + h2 * (A1v[3 * i + j] + dv);
                jc[JVR + 3 * i + j] = h2 * (A1r[3 * i + j] + dr2);
                jc[JVW + 3 * i + j] = hh * A2r[3 * i + j];
            }
            T dT = A2T[i];
#pragma unroll
            for (int k = 0; k < 3; k++) dT += A2v[3 * i + k] * h * A1T[k];
            jc[JPT + i] = hh * A1T[i];
            jc[JVT + i] = h2 * (A1T[i] + dT);
        }
    }
}

// Stage cost (+ gradient when GRAD).  hdr = [ref(3) f_ext(3) w_wp w_in w_rate yaw_ref].
template <typename T, bool GRAD>
__device__ __forceinline__ T objective(const T z[NZ], const T* hdr, bool first, bool final_terminal, T g[NZ])
{
    using C = Const<T>;
    const T wwp = hdr[6], win = hdr[7], wrate = hdr[8];
    T f = T(0);
    if (GRAD) {
#pragma unroll
        for (int i = 0; i < NZ; i++) g[i] = T(0);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        T e = hdr[i] - z[8 + i];
        f += wwp * e * e + win * C::inv_rate2 * z[i] * z[i];
        if (GRAD) {
            g[8 + i] = T(-2) * wwp * e;
            g[i] = T(2) * win * C::inv_rate2 * z[i];
        }
    }
    T ey = hdr[9] - z[16];
    f += T(12) * wwp * ey * ey;
    if (GRAD) g[16] = T(-24) * wwp * ey;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        T du = z[i] - z[4 + i];
        f += wrate * du * du;
        if (GRAD) {
            g[i] += T(2) * wrate * du;
            g[4 + i] = T(-2) * wrate * du;
        }
    }
    if (first) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            f += T(10) * win * z[4 + i] * z[4 + i];
            if (GRAD) g[4 + i] += T(20) * win * z[4 + i];
        }
    }
    if (final_terminal) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            f += T(20) * wwp * z[11 + i] * z[11 + i];
            if (GRAD) g[11 + i] += T(40) * wwp * z[11 + i];
        }
    }
    return f;
}

// diagonal of the constant (Gauss-Newton = exact objective) Hessian; the only off-diagonal
// entries are H[u_i][uprev_i] = -2 w_rate, which the Riccati sweep reads from hdr[8].
template <typename T>
__device__ __forceinline__ T cost_hess_diag(int i, const T* hdr, bool first, bool final_terminal)
{
    using C = Const<T>;
    const T wwp = hdr[6], win = hdr[7], wrate = hdr[8];
    if (i < 4) return T(2) * wrate + (i < 3 ? T(2) * win * C::inv_rate2 : T(0));
    if (i < 8) return T(2) * wrate + ((first && i < 7) ? T(20) * win : T(0));
    if (i < 11) return T(2) * wwp;
    if (i < 14) return final_terminal ? T(40) * wwp : T(0);
    if (i == 16) return T(24) * wwp;
    return T(0);
}

// (J' y)_i for the compact Jacobian; y in c-ordering [x+(9); u(4)].
template <typename T> __device__ __forceinline__ T jt_y(const T* jc, const T* y, int i)
{
    using C = Const<T>;
    if (i < 3) return jc[JVW + i] * y[3] + jc[JVW + 3 + i] * y[4] + jc[JVW + 6 + i] * y[5] + C::h * y[6 + i] + y[9 + i];
    if (i == 3)
        return jc[JPT] * y[0] + jc[JPT + 1] * y[1] + jc[JPT + 2] * y[2] + jc[JVT] * y[3] + jc[JVT + 1] * y[4] +
               jc[JVT + 2] * y[5] + y[12];
    if (i < 8) return T(0);
    if (i < 11) return y[i - 8];
    if (i < 14) {
        int j = i - 11;
        return jc[JPV + j] * y[0] + jc[JPV + 3 + j] * y[1] + jc[JPV + 6 + j] * y[2] + jc[JVV + j] * y[3] +
               jc[JVV + 3 + j] * y[4] + jc[JVV + 6 + j] * y[5];
    }
    int j = i - 14;
    return jc[JPR + j] * y[0] + jc[JPR + 3 + j] * y[1] + jc[JPR + 6 + j] * y[2] + jc[JVR + j] * y[3] +
           jc[JVR + 3 + j] * y[4] + jc[JVR + 6 + j] * y[5] + y[6 + j];
}

}  // namespace nmpc
