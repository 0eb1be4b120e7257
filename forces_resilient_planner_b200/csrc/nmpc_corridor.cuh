// nmpc_corridor.cuh -- safe-corridor generation and per-stage polytope selection (sm_100a).
//
// Batched, device-resident form of the corridor part of NMPCSolver::setFORCESParams with getSikangConst
// (/root/reference/src/resilient_planner/plan_manage/src/nmpc_solver.cpp:288-332, 493-516) and of the
// DecompROS routines it drives, as vendored by the reference under
// /root/reference/src/ThirdParty/DecompROS/decomp_ros_utils/include/ :
//   EllipsoidDecomp::dilate (decomp_util/ellipsoid_decomp.h:76-100), LineSegment::dilate / add_local_bbox /
//   find_ellipsoid<3> (decomp_util/line_segment.h:31-35, 46-88, 137-208), DecompBase::set_obs / find_polyhedron
//   (decomp_util/decomp_base.h:33-38, 66-85), Ellipsoid::closest_point / closest_hyperplane
//   (decomp_geometry/ellipsoid.h:42-60), LinearConstraint(p0, hyperplanes) (decomp_geometry/polyhedron.h:100-120).
//
// Per agent the stages are walked in order (the choice at stage i depends on the polytope made at an
// earlier stage): keep the last polytope while the reference point, inflated by 1.1 ||E_i a_j||, is
// inside; otherwise dilate a new one around the 0.1 m seed segment along the yaw reference.  One CTA
// per agent; what is parallel is the point cloud: every carving step is ONE pass over the agent's
// obstacle points (drop the points the previous plane cut off, find the closest survivor in the
// ellipsoid metric) with a block-wide argmin whose ties go to the lowest index, as the reference's
// sequential scan does.  Output = the poly_A / poly_b / poly_m / poly_idx inputs of pack_params_kernel.
#pragma once
#include <cuda_runtime.h>

namespace nmpc {

struct CorridorParams {
    int B, N, M, P, R;           // agents, stages, max cloud points per agent, max polytopes / agent, max rows / polytope
    const double* cloud;         // [B][M][3], or one shared cloud [M][3] when cloud_stride == 0
    long long cloud_stride;      // doubles between two agents' clouds (M * 3, or 0)
    const int* cloud_n;          // [B] (or [1] when shared): live points
    const double* ref_pos;       // [B][N][3]
    const double* ref_yaw;       // [B][N]
    const double* ellipsoid;     // [B][N][9]  E_i row-major
    double bbox0, bbox1, bbox2;  // local bounding box (nmpc_solver.cpp:323: 2, 2, 1)
    double* poly_A;              // [B][P][R][3]
    double* poly_b;              // [B][P][R]
    int* poly_m;                 // [B][P]   rows stored (carved planes first, then the six box planes)
    int* poly_idx;               // [B][N]
    int* n_poly;                 // [B]
    int* overflow;               // [B]  1: a polytope had more than R rows, 2: more than P polytopes were needed
};

constexpr int COR_THREADS = 128;
constexpr double COR_EPS = 1e-10;      // decomp_basis/data_type.h:129

struct Mat3 { double m[9]; };

__device__ __forceinline__ Mat3 inv3(const Mat3& a)
{
    const double* m = a.m;
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
    Mat3 r;
    r.m[0] = c00 * id; r.m[1] = (m[2] * m[7] - m[1] * m[8]) * id; r.m[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    r.m[3] = c01 * id; r.m[4] = (m[0] * m[8] - m[2] * m[6]) * id; r.m[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    r.m[6] = c02 * id; r.m[7] = (m[1] * m[6] - m[0] * m[7]) * id; r.m[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return r;
}
// R diag(ax) R'
__device__ __forceinline__ Mat3 rdr(const Mat3& R, const double ax[3])
{
    Mat3 c;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c.m[3 * i + j] = R.m[3 * i] * ax[0] * R.m[3 * j] + R.m[3 * i + 1] * ax[1] * R.m[3 * j + 1] + R.m[3 * i + 2] * ax[2] * R.m[3 * j + 2];
    return c;
}
__device__ __forceinline__ double edist(const Mat3& ci, const double d[3], const double p[3])
{
    const double x = p[0] - d[0], y = p[1] - d[1], z = p[2] - d[2];
    const double a = ci.m[0] * x + ci.m[1] * y + ci.m[2] * z, b = ci.m[3] * x + ci.m[4] * y + ci.m[5] * z, c = ci.m[6] * x + ci.m[7] * y + ci.m[8] * z;
    return sqrt(a * a + b * b + c * c);
}

// block-wide argmin of (value, index) pairs, lowest index on ties; every thread gets the result
__device__ __forceinline__ void block_argmin(double& v, int& idx, double* s_v, int* s_i)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (v2 < v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
    }
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_v[w] = v; s_i[w] = idx; }
    __syncthreads();
    v = s_v[0]; idx = s_i[0];
#pragma unroll
    for (int k = 1; k < COR_THREADS / 32; k++)
        if (s_v[k] < v || (s_v[k] == v && s_i[k] < idx)) { v = s_v[k]; idx = s_i[k]; }
}

__global__ void __launch_bounds__(COR_THREADS) corridor_select_kernel(const CorridorParams q)
{
    extern __shared__ unsigned char cor_smem[];
    // flags per cloud point: bit 0 inside the local box, bit 1 "inside the ellipsoid" working set, bit 2 its initial copy,
    // bit 3 remaining for the carving
    unsigned char* flag = cor_smem;
    double* rows = reinterpret_cast<double*>(cor_smem + ((q.M + 15) & ~15));     // last polytope: [R][4] = a0 a1 a2 b
    __shared__ double s_v[COR_THREADS / 32];
    __shared__ int s_i[COR_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (b >= q.B) return;
    const double* cloud = q.cloud + (size_t)b * q.cloud_stride;
    const int n = q.cloud_n[q.cloud_stride ? b : 0] < q.M ? q.cloud_n[q.cloud_stride ? b : 0] : q.M;
    int npoly = 0, last_m = 0, ovf = 0;

    for (int i = 0; i < q.N; i++) {
        const double* rp = q.ref_pos + ((size_t)b * q.N + i) * 3;
        const double ref[3] = {rp[0], rp[1], rp[2]};
        if (npoly > 0) {
            // ---- getSikangConst: is the inflated reference point inside the last polytope? ----
            const double* E = q.ellipsoid + ((size_t)b * q.N + i) * 9;
            int out = 0;
            for (int j = tid; j < last_m; j += COR_THREADS) {
                const double a0 = rows[4 * j], a1 = rows[4 * j + 1], a2 = rows[4 * j + 2];
                const double e0 = E[0] * a0 + E[1] * a1 + E[2] * a2, e1 = E[3] * a0 + E[4] * a1 + E[5] * a2, e2 = E[6] * a0 + E[7] * a1 + E[8] * a2;
                if ((a0 * ref[0] + a1 * ref[1] + a2 * ref[2]) - (rows[4 * j + 3] - 1.1 * sqrt(e0 * e0 + e1 * e1 + e2 * e2)) > 0) out = 1;
            }
            if (!__syncthreads_or(out)) {
                if (tid == 0) q.poly_idx[(size_t)b * q.N + i] = npoly - 1;
                continue;
            }
        }
        if (npoly >= q.P) {             // no room for another polytope: keep using the last one, report it
            ovf |= 2;
            if (tid == 0) q.poly_idx[(size_t)b * q.N + i] = npoly - 1;
            continue;
        }
        // ---- EllipsoidDecomp::dilate on the seed segment p1 = ref, p2 = ref + 0.1 (cos yaw, sin yaw, 0) ----
        const double yaw = q.ref_yaw[(size_t)b * q.N + i];
        double sy, cy;
        sincos(yaw, &sy, &cy);
        const double p1[3] = {ref[0], ref[1], ref[2]}, p2[3] = {ref[0] + 0.1 * cy, ref[1] + 0.1 * sy, ref[2]};
        const double seg[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        const double len = sqrt(seg[0] * seg[0] + seg[1] * seg[1] + seg[2] * seg[2]);
        const double dir[3] = {seg[0] / len, seg[1] / len, seg[2] / len};
        double dh[3] = {dir[1], -dir[0], 0.0};
        {
            double nh = sqrt(dh[0] * dh[0] + dh[1] * dh[1]);
            if (nh == 0) { dh[0] = -1.0; dh[1] = 0.0; nh = 1.0; }
            dh[0] /= nh; dh[1] /= nh;
        }
        const double dv[3] = {dir[1] * dh[2] - dir[2] * dh[1], dir[2] * dh[0] - dir[0] * dh[2], dir[0] * dh[1] - dir[1] * dh[0]};
        // the six box planes (point, outward normal), add_local_bbox order
        double bp[6][3], bn[6][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            bp[0][c] = p1[c] + dh[c] * q.bbox1; bn[0][c] = dh[c];
            bp[1][c] = p1[c] - dh[c] * q.bbox1; bn[1][c] = -dh[c];
            bp[2][c] = p2[c] + dir[c] * q.bbox0; bn[2][c] = dir[c];
            bp[3][c] = p1[c] - dir[c] * q.bbox0; bn[3][c] = -dir[c];
            bp[4][c] = p1[c] + dv[c] * q.bbox2; bn[4][c] = dv[c];
            bp[5][c] = p1[c] - dv[c] * q.bbox2; bn[5][c] = -dv[c];
        }
        // Ri = Rz(yaw_seg) Ry(pitch_seg)   (vec3_to_rotation, zero roll)
        Mat3 Ri;
        {
            const double pit = atan2(-seg[2], sqrt(seg[0] * seg[0] + seg[1] * seg[1])), yw = atan2(seg[1], seg[0]);
            double sp, cp, s2, c2;
            sincos(pit, &sp, &cp); sincos(yw, &s2, &c2);
            Ri.m[0] = c2 * cp; Ri.m[1] = -s2; Ri.m[2] = c2 * sp;
            Ri.m[3] = s2 * cp; Ri.m[4] = c2;  Ri.m[5] = s2 * sp;
            Ri.m[6] = -sp;     Ri.m[7] = 0.0; Ri.m[8] = cp;
        }
        Mat3 Rf = Ri;
        const double f = len / 2;
        double axes[3] = {f, f, f};
        const double d[3] = {(p1[0] + p2[0]) / 2, (p1[1] + p2[1]) / 2, (p1[2] + p2[2]) / 2};
        Mat3 C = rdr(Ri, axes), Ci = inv3(C);
        // ---- set_obs (points inside the local box) and the points inside the initial ellipsoid ----
        int cnt = 0;
        for (int j = tid; j < n; j += COR_THREADS) {
            const double p[3] = {cloud[3 * j], cloud[3 * j + 1], cloud[3 * j + 2]};
            bool in = true;
#pragma unroll
            for (int k = 0; k < 6; k++)
                in = in && (bn[k][0] * (p[0] - bp[k][0]) + bn[k][1] * (p[1] - bp[k][1]) + bn[k][2] * (p[2] - bp[k][2]) <= COR_EPS);
            unsigned char fl = in ? 9 : 0;                       // in the box, remaining for the carving
            if (in && edist(Ci, d, p) <= 1.0) { fl |= 6; cnt = 1; }
            flag[j] = fl;
        }
        // ---- find_ellipsoid<3>: shrink the two short axes until no obstacle is inside ----
        for (int phase = 0; phase < 2; phase++) {
            if (phase == 1) {            // "reset ellipsoid with old axes(2)", points_inside(obs) again
                C = rdr(Rf, axes); Ci = inv3(C);
                cnt = 0;
                for (int j = tid; j < n; j += COR_THREADS) {
                    unsigned char fl = flag[j] & ~2;
                    if (fl & 4) {
                        const double p[3] = {cloud[3 * j], cloud[3 * j + 1], cloud[3 * j + 2]};
                        if (edist(Ci, d, p) <= 1.0) { fl |= 2; cnt = 1; }
                    }
                    flag[j] = fl;
                }
            }
            while (__syncthreads_or(cnt)) {
                double bv = 1e300; int bi = 0x7fffffff;
                for (int j = tid; j < n; j += COR_THREADS)
                    if (flag[j] & 2) {
                        const double p[3] = {cloud[3 * j], cloud[3 * j + 1], cloud[3 * j + 2]};
                        const double dd = edist(Ci, d, p);
                        if (dd < bv) { bv = dd; bi = j; }
                    }
                block_argmin(bv, bi, s_v, s_i);
                const double pw[3] = {cloud[3 * bi] - d[0], cloud[3 * bi + 1] - d[1], cloud[3 * bi + 2] - d[2]};
                if (phase == 0) {
                    double pl[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) pl[c] = Ri.m[c] * pw[0] + Ri.m[3 + c] * pw[1] + Ri.m[6 + c] * pw[2];     // Ri' pw
                    const double roll = atan2(pl[2], pl[1]);
                    double sr, cr;
                    sincos(roll, &sr, &cr);
#pragma unroll
                    for (int r = 0; r < 3; r++) {                // Rf = Ri Rx(roll)
                        Rf.m[3 * r] = Ri.m[3 * r];
                        Rf.m[3 * r + 1] = Ri.m[3 * r + 1] * cr + Ri.m[3 * r + 2] * sr;
                        Rf.m[3 * r + 2] = -Ri.m[3 * r + 1] * sr + Ri.m[3 * r + 2] * cr;
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) pl[c] = Rf.m[c] * pw[0] + Rf.m[3 + c] * pw[1] + Rf.m[6 + c] * pw[2];     // Rf' pw
                    if (pl[0] < axes[0]) {
                        const double r0 = pl[0] / axes[0];
                        axes[1] = fabs(pl[1]) / sqrt(1 - r0 * r0);
                    }
                    const double ax[3] = {axes[0], axes[1], axes[1]};
                    C = rdr(Rf, ax);
                } else {
                    double pl[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) pl[c] = Rf.m[c] * pw[0] + Rf.m[3 + c] * pw[1] + Rf.m[6 + c] * pw[2];
                    const double r0 = pl[0] / axes[0], r1 = pl[1] / axes[1];
                    const double dd = 1 - r0 * r0 - r1 * r1;
                    if (dd > COR_EPS) axes[2] = fabs(pl[2]) / sqrt(dd);
                    C = rdr(Rf, axes);
                }
                Ci = inv3(C);
                cnt = 0;
                for (int j = tid; j < n; j += COR_THREADS)
                    if (flag[j] & 2) {
                        const double p[3] = {cloud[3 * j], cloud[3 * j + 1], cloud[3 * j + 2]};
                        if (1 - edist(Ci, d, p) > COR_EPS) cnt = 1; else flag[j] &= ~2;
                    }
            }
        }
        // ---- find_polyhedron: carve with the tangent plane at the closest remaining point, drop what it cuts off ----
        C = rdr(Rf, axes); Ci = inv3(C);
        int m = 0;
        double pn[3] = {0, 0, 0}, pp[3] = {0, 0, 0};
        bool have_plane = false;
        for (;;) {
            double bv = 1e300; int bi = 0x7fffffff;
            for (int j = tid; j < n; j += COR_THREADS)
                if (flag[j] & 8) {
                    const double p[3] = {cloud[3 * j], cloud[3 * j + 1], cloud[3 * j + 2]};
                    if (have_plane && !(pn[0] * (p[0] - pp[0]) + pn[1] * (p[1] - pp[1]) + pn[2] * (p[2] - pp[2]) < 0)) { flag[j] &= ~8; continue; }
                    const double dd = edist(Ci, d, p);
                    if (dd < bv) { bv = dd; bi = j; }
                }
            block_argmin(bv, bi, s_v, s_i);
            if (bi == 0x7fffffff) break;
            pp[0] = cloud[3 * bi]; pp[1] = cloud[3 * bi + 1]; pp[2] = cloud[3 * bi + 2];
            {
                const double x = pp[0] - d[0], y = pp[1] - d[1], z = pp[2] - d[2];
                // n = Ci Ci' (pt - d)
                const double t0 = Ci.m[0] * x + Ci.m[3] * y + Ci.m[6] * z, t1 = Ci.m[1] * x + Ci.m[4] * y + Ci.m[7] * z, t2 = Ci.m[2] * x + Ci.m[5] * y + Ci.m[8] * z;
                pn[0] = Ci.m[0] * t0 + Ci.m[1] * t1 + Ci.m[2] * t2; pn[1] = Ci.m[3] * t0 + Ci.m[4] * t1 + Ci.m[5] * t2; pn[2] = Ci.m[6] * t0 + Ci.m[7] * t1 + Ci.m[8] * t2;
                const double nn = sqrt(pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]);
                pn[0] /= nn; pn[1] /= nn; pn[2] /= nn;
            }
            have_plane = true;
            if (m < q.R) {
                if (tid == 0) {
                    double c = pp[0] * pn[0] + pp[1] * pn[1] + pp[2] * pn[2];
                    double s = (pn[0] * d[0] + pn[1] * d[1] + pn[2] * d[2]) - c > 0 ? -1.0 : 1.0;      // LinearConstraint: p0 must be inside
                    rows[4 * m] = s * pn[0]; rows[4 * m + 1] = s * pn[1]; rows[4 * m + 2] = s * pn[2]; rows[4 * m + 3] = s * c;
                }
                m++;
            } else {
                ovf |= 1;
            }
        }
        if (tid == 0) {
            for (int k = 0; k < 6; k++) {
                if (m + k >= q.R) break;
                const double c = bp[k][0] * bn[k][0] + bp[k][1] * bn[k][1] + bp[k][2] * bn[k][2];
                const double s = (bn[k][0] * d[0] + bn[k][1] * d[1] + bn[k][2] * d[2]) - c > 0 ? -1.0 : 1.0;
                rows[4 * (m + k)] = s * bn[k][0]; rows[4 * (m + k) + 1] = s * bn[k][1]; rows[4 * (m + k) + 2] = s * bn[k][2];
                rows[4 * (m + k) + 3] = s * c;
            }
        }
        if (m + 6 > q.R) ovf |= 1;
        m = m + 6 < q.R ? m + 6 : q.R;
        __syncthreads();
        // publish polytope `npoly`
        double* gA = q.poly_A + ((size_t)b * q.P + npoly) * q.R * 3;
        double* gb = q.poly_b + ((size_t)b * q.P + npoly) * q.R;
        for (int j = tid; j < q.R; j += COR_THREADS) {
            const bool live = j < m;
            gA[3 * j] = live ? rows[4 * j] : 0.0; gA[3 * j + 1] = live ? rows[4 * j + 1] : 0.0; gA[3 * j + 2] = live ? rows[4 * j + 2] : 0.0;
            gb[j] = live ? rows[4 * j + 3] : 0.0;
        }
        if (tid == 0) { q.poly_m[(size_t)b * q.P + npoly] = m; q.poly_idx[(size_t)b * q.N + i] = npoly; }
        last_m = m;
        npoly++;
        __syncthreads();
    }
    if (tid == 0) {
        q.n_poly[b] = npoly; q.overflow[b] = ovf;
        for (int k = npoly; k < q.P; k++) q.poly_m[(size_t)b * q.P + k] = 0;
    }
}

}  // namespace nmpc
