// K4 -- one reverse-SDE pose update for B copies of one ligand topology, one warp per copy.
// Replaces utils/sampling.py:119-141 (perturbation = c_score*score + c_noise*z) followed by
// modify_conformer_batch (utils/diffusion_utils.py:60-78):
//   axis_angle_to_matrix (utils/geometry.py:39-86)  ->  rigid move about the centroid
//   sequential bond rotations in bond order (utils/torsion.py:75-90)
//   Kabsch alignment of the twisted onto the rigid conformer with the det<0 fix (utils/geometry.py:246-276)
// The reference runs R x (several ATen kernels + bmm) plus a batched cuSOLVER SVD per step; here the
// whole update is one launch, positions staged in shared memory, the 3x3 SVD done in registers
// (cyclic Jacobi on H^T H in double; the smallest singular pair is rebuilt by cross products, which
// is exactly the diag(1,1,-1) reflection fix).
// HBM traffic per copy: 2 * 12 N bytes of positions + (9 + R) floats of scores and noise.
#include "common.cuh"
#include "../../include/cb200.h"

namespace {

constexpr int kWarps = 4;

struct Mat3 { float m[9]; };

// geometry.py:39-86 (pytorch3d axis-angle -> quaternion -> matrix)
__device__ __forceinline__ Mat3 axis_angle_to_matrix(float ax, float ay, float az) {
    const float angle = sqrtf(ax * ax + ay * ay + az * az);
    const float half = 0.5f * angle;
    const float s = fabsf(angle) < 1e-6f ? 0.5f - (angle * angle) / 48.0f : sinf(half) / angle;
    const float r = cosf(half), i = ax * s, j = ay * s, k = az * s;
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    Mat3 o;
    o.m[0] = 1.0f - two_s * (j * j + k * k);
    o.m[1] = two_s * (i * j - k * r);
    o.m[2] = two_s * (i * k + j * r);
    o.m[3] = two_s * (i * j + k * r);
    o.m[4] = 1.0f - two_s * (i * i + k * k);
    o.m[5] = two_s * (j * k - i * r);
    o.m[6] = two_s * (i * k - j * r);
    o.m[7] = two_s * (j * k + i * r);
    o.m[8] = 1.0f - two_s * (i * i + j * j);
    return o;
}

__device__ __forceinline__ void jacobi_rotate(double A[3][3], double V[3][3], int p, int q) {
    if (fabs(A[p][q]) < 1e-300) return;
    const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
    const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    for (int k = 0; k < 3; ++k) {
        const double akp = A[k][p], akq = A[k][q];
        A[k][p] = c * akp - s * akq;
        A[k][q] = s * akp + c * akq;
    }
    for (int k = 0; k < 3; ++k) {
        const double apk = A[p][k], aqk = A[q][k];
        A[p][k] = c * apk - s * aqk;
        A[q][k] = s * apk + c * aqk;
    }
    for (int k = 0; k < 3; ++k) {
        const double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - s * vkq;
        V[k][q] = s * vkp + c * vkq;
    }
}

// Optimal proper rotation R (row-major) minimising |R a - b| given H = sum a b^T  (geometry.py:262-273).
__device__ void kabsch_rotation(const double H[3][3], float R[9]) {
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += H[k][i] * H[k][j];
            A[i][j] = s;  // H^T H
        }
    for (int sweep = 0; sweep < 12; ++sweep) {
        jacobi_rotate(A, V, 0, 1);
        jacobi_rotate(A, V, 0, 2);
        jacobi_rotate(A, V, 1, 2);
    }
    // order the two largest eigenpairs
    int i0 = 0, i1 = 1, i2 = 2;
    if (A[i0][i0] < A[i1][i1]) { int t = i0; i0 = i1; i1 = t; }
    if (A[i0][i0] < A[i2][i2]) { int t = i0; i0 = i2; i2 = t; }
    if (A[i1][i1] < A[i2][i2]) { int t = i1; i1 = i2; i2 = t; }
    double v1[3] = {V[0][i0], V[1][i0], V[2][i0]}, v2[3] = {V[0][i1], V[1][i1], V[2][i1]};
    double u1[3], u2[3];
    for (int i = 0; i < 3; ++i) {
        u1[i] = H[i][0] * v1[0] + H[i][1] * v1[1] + H[i][2] * v1[2];
        u2[i] = H[i][0] * v2[0] + H[i][1] * v2[1] + H[i][2] * v2[2];
    }
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (n1 < 1e-30) {  // H == 0: nothing to align
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
        return;
    }
    for (int i = 0; i < 3; ++i) u1[i] /= n1;
    double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int i = 0; i < 3; ++i) u2[i] -= d * u1[i];
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    if (n2 < 1e-12 * n1) {  // rank one: any unit vector orthogonal to u1 (and to v1 on the other side)
        const int k = fabs(u1[0]) < fabs(u1[1]) ? (fabs(u1[0]) < fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) < fabs(u1[2]) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[k] = 1.0;
        d = u1[k];
        for (int i = 0; i < 3; ++i) u2[i] = e[i] - d * u1[i];
        n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    }
    for (int i = 0; i < 3; ++i) u2[i] /= n2;
    const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
    // H = U S V^T  ->  R = V U^T
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = (float)(v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j]);
}

__global__ void __launch_bounds__(kWarps * 32)
sde_step_kernel(cb_sde_step_args a) {
    extern __shared__ float sm[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kWarps + w;
    if (g >= a.B) return;
    const int N = a.N, R = a.R;
    float* rig = sm + (size_t)w * 6 * N;  // [N][3] rigid-moved conformer
    float* flx = rig + 3 * N;             // [N][3] conformer being twisted
    float* gp = a.pos + (size_t)g * 3 * N;

    // perturbations.  Coefficients by value, or read from device memory (a step loop captured in a CUDA graph keeps its
    // launch parameters: the per-step scalars then live in a device array the host refreshes between replays)
    if (a.coeffs_dev != nullptr) {
        a.c_tr_score = __ldg(a.coeffs_dev + 0); a.c_tr_noise = __ldg(a.coeffs_dev + 1);
        a.c_rot_score = __ldg(a.coeffs_dev + 2); a.c_rot_noise = __ldg(a.coeffs_dev + 3);
        a.c_tor_score = __ldg(a.coeffs_dev + 4); a.c_tor_noise = __ldg(a.coeffs_dev + 5);
    }
    float tr[3], rot[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        tr[k] = a.c_tr_score * a.tr_score[3 * g + k] + (a.z_tr ? a.c_tr_noise * a.z_tr[3 * g + k] : 0.0f);
        rot[k] = a.c_rot_score * a.rot_score[3 * g + k] + (a.z_rot ? a.c_rot_noise * a.z_rot[3 * g + k] : 0.0f);
    }
    // centroid
    float cx = 0, cy = 0, cz = 0;
    for (int n = lane; n < N; n += 32) {
        const float x = gp[3 * n], y = gp[3 * n + 1], z = gp[3 * n + 2];
        flx[3 * n] = x; flx[3 * n + 1] = y; flx[3 * n + 2] = z;
        cx += x; cy += y; cz += z;
    }
    cx = cb_warp_sum(cx) / (float)N;
    cy = cb_warp_sum(cy) / (float)N;
    cz = cb_warp_sum(cz) / (float)N;
    const Mat3 Rm = axis_angle_to_matrix(rot[0], rot[1], rot[2]);
    __syncwarp();
    for (int n = lane; n < N; n += 32) {
        const float x = flx[3 * n] - cx, y = flx[3 * n + 1] - cy, z = flx[3 * n + 2] - cz;
        const float nx = (Rm.m[0] * x + Rm.m[1] * y + Rm.m[2] * z) + tr[0] + cx;
        const float ny = (Rm.m[3] * x + Rm.m[4] * y + Rm.m[5] * z) + tr[1] + cy;
        const float nz = (Rm.m[6] * x + Rm.m[7] * y + Rm.m[8] * z) + tr[2] + cz;
        rig[3 * n] = nx; rig[3 * n + 1] = ny; rig[3 * n + 2] = nz;
        flx[3 * n] = nx; flx[3 * n + 1] = ny; flx[3 * n + 2] = nz;
    }
    __syncwarp();
    if (R == 0 || a.tor_score == nullptr) {
        for (int n = lane; n < 3 * N; n += 32) gp[n] = rig[n];
        return;
    }
    // sequential bond rotations
    for (int b = 0; b < R; ++b) {
        const int u = a.bond_uv[2 * b], v = a.bond_uv[2 * b + 1];
        const float dtau = a.c_tor_score * a.tor_score[(size_t)g * R + b] +
                           (a.z_tor ? a.c_tor_noise * a.z_tor[(size_t)g * R + b] : 0.0f);
        const float px = flx[3 * v], py = flx[3 * v + 1], pz = flx[3 * v + 2];
        float ax = flx[3 * u] - px, ay = flx[3 * u + 1] - py, az = flx[3 * u + 2] - pz;
        const float nrm = sqrtf(ax * ax + ay * ay + az * az);
        ax = ax / nrm * dtau; ay = ay / nrm * dtau; az = az / nrm * dtau;
        const Mat3 T = axis_angle_to_matrix(ax, ay, az);
        __syncwarp();
        const uint8_t* mk = a.mask_rotate + (size_t)b * N;
        for (int n = lane; n < N; n += 32) {
            if (mk[n]) {
                const float x = flx[3 * n] - px, y = flx[3 * n + 1] - py, z = flx[3 * n + 2] - pz;
                flx[3 * n] = (T.m[0] * x + T.m[1] * y + T.m[2] * z) + px;
                flx[3 * n + 1] = (T.m[3] * x + T.m[4] * y + T.m[5] * z) + py;
                flx[3 * n + 2] = (T.m[6] * x + T.m[7] * y + T.m[8] * z) + pz;
            }
        }
        __syncwarp();
    }
    // Kabsch: align flx (A) onto rig (B)
    float ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
    for (int n = lane; n < N; n += 32) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { ca[k] += flx[3 * n + k]; cb[k] += rig[3 * n + k]; }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { ca[k] = cb_warp_sum(ca[k]) / (float)N; cb[k] = cb_warp_sum(cb[k]) / (float)N; }
    float h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int n = lane; n < N; n += 32) {
        float am[3], bm[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { am[k] = flx[3 * n + k] - ca[k]; bm[k] = rig[3 * n + k] - cb[k]; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) h[3 * i + j] = fmaf(am[i], bm[j], h[3 * i + j]);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) h[i] = cb_warp_sum(h[i]);
    float Rk[9];
    {
        double Hd[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Hd[i][j] = (double)h[3 * i + j];
        kabsch_rotation(Hd, Rk);  // every lane computes the same 3x3 (no divergence, no broadcast needed)
    }
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = -(Rk[3 * i] * ca[0] + Rk[3 * i + 1] * ca[1] + Rk[3 * i + 2] * ca[2]) + cb[i];
    for (int n = lane; n < N; n += 32) {
        const float x = flx[3 * n], y = flx[3 * n + 1], z = flx[3 * n + 2];
        gp[3 * n] = (Rk[0] * x + Rk[1] * y + Rk[2] * z) + t[0];
        gp[3 * n + 1] = (Rk[3] * x + Rk[4] * y + Rk[5] * z) + t[1];
        gp[3 * n + 2] = (Rk[6] * x + Rk[7] * y + Rk[8] * z) + t[2];
    }
}

}  // namespace

extern "C" int cb_sde_step(const cb_sde_step_args* a, void* stream) {
    CB_CHECK_ARG(a != nullptr, "cb_sde_step: null args");
    CB_CHECK_ARG(a->B >= 0 && a->N > 0 && a->R >= 0, "cb_sde_step: bad sizes B=%d N=%d R=%d", a->B, a->N, a->R);
    if (a->B == 0) return CB_OK;
    CB_CHECK_ARG(a->pos && a->tr_score && a->rot_score, "cb_sde_step: null pointer");
    CB_CHECK_ARG(a->R == 0 || a->tor_score == nullptr || (a->bond_uv && a->mask_rotate), "cb_sde_step: torsion tables missing");
    const size_t smem = (size_t)kWarps * 6 * a->N * sizeof(float);
    CB_CHECK_ARG(smem <= 200 * 1024, "cb_sde_step: ligand with %d atoms does not fit in shared memory", a->N);
    cudaFuncSetAttribute(sde_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sde_step_kernel<<<cb_div_up(a->B, kWarps), kWarps * 32, smem, (cudaStream_t)stream>>>(*a);
    CB_CHECK_LAUNCH("cb_sde_step");
    return CB_OK;
}
