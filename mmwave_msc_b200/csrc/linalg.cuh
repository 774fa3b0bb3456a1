// Warp-cooperative float64 linear algebra for the 6x6 / 9x9 tracker matrices.
// One warp owns one track; matrices live in shared memory, lanes split the elements.
#pragma once
#include "mmw_internal.cuh"

namespace mmw {

constexpr unsigned kFull = 0xffffffffu;

// Per-warp scratch (doubles): A 36 | B 36 | K 54 | M1 81 | v 8.  M2 = K R (9x6) reuses A|B once S and S^-1 are dead.
constexpr int kWsAug = 0, kWsA = 0, kWsB = 36, kWsK = 72, kWsM1 = 126, kWsM2 = 0, kWsV = 207;
constexpr int kWarpScratch = 216;

__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(kFull, v, src);
}

// Inverse and determinant of a 6x6 row-major matrix by Gauss-Jordan elimination with partial
// pivoting (numpy.linalg.inv / det are LU with partial pivoting: Tracking.py:558-559, filterpy update).
// Lane l < 12 keeps column l of the augmented matrix [A | I] in six registers; pivot search, row swap and the
// multipliers travel by warp shuffle, so there is no shared-memory round trip or __syncwarp inside the six
// elimination steps.  A, Ainv: 36 doubles in shared memory (may not alias).  `aug` is unused (kept for the
// callers' scratch layout).  Returns det(A); every lane returns the same value.
__device__ __forceinline__ double warp_inv6(const double* A, double* Ainv, double* /*aug*/, int lane) {
    const int col = lane < 12 ? lane : 11;          // lanes >= 12 shadow lane 11 (results discarded)
    double r[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) r[i] = col < 6 ? A[i * 6 + col] : (col - 6 == i ? 1.0 : 0.0);
    double det = 1.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        // partial pivoting on column k (held by lane k): first row of maximum magnitude, like LAPACK's idamax
        int p = k;
        double best = fabs(r[k]);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const double v = fabs(r[i]);
            if (v > best) { best = v; p = i; }
        }
        p = __shfl_sync(kFull, p, k);
#pragma unroll
        for (int i = k + 1; i < 6; ++i)
            if (p == i) { const double t = r[k]; r[k] = r[i]; r[i] = t; }
        if (p != k) det = -det;
        const double piv = shfl_d(r[k], k);
        det *= piv;
        r[k] = div_zero_fast(r[k], piv);            // the identity half of [A | I] is mostly exact zeros
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i == k) continue;
            const double f = shfl_d(r[i], k);       // multiplier = column k's entry of row i (before the update)
            r[i] -= f * r[k];
        }
    }
    if (lane >= 6 && lane < 12) {
#pragma unroll
        for (int i = 0; i < 6; ++i) Ainv[i * 6 + (lane - 6)] = r[i];
    }
    __syncwarp();
    return det;
}

// filterpy KalmanFilter.predict with F = KF_F(dt), Q = KF_Q_DISCR(dt) (constants.py:195-215):
//   x = F x ;  P = (F P) F' + Q.  x: 9, P: 81 in shared memory; tmp: 81 doubles scratch.
__device__ __forceinline__ void warp_kf_predict(double* x, double* P, double dt, double q_var, double* tmp,
                                                int lane) {
    const double h = 0.5 * (dt * dt);
    double nx = 0.0;
    if (lane < 9) {
        nx = x[lane];
        if (lane < 6) nx += dt * x[lane + 3];
        if (lane < 3) nx += h * x[lane + 6];
    }
    for (int e = lane; e < 81; e += 32) {
        const int i = e / 9, j = e % 9;
        double v = P[e];
        if (i < 6) v += dt * P[(i + 3) * 9 + j];
        if (i < 3) v += h * P[(i + 6) * 9 + j];
        tmp[e] = v;
    }
    __syncwarp();
    if (lane < 9) x[lane] = nx;
    const double dt2 = dt * dt, dt3 = dt2 * dt, dt4 = dt2 * dt2;
    for (int e = lane; e < 81; e += 32) {
        const int i = e / 9, j = e % 9;
        double v = tmp[e];
        if (j < 6) v += dt * tmp[i * 9 + j + 3];
        if (j < 3) v += h * tmp[i * 9 + j + 6];
        if (i / 3 == j / 3) {   // block_diag(Qw, Qw, Qw): blocks on state indices (0-2), (3-5), (6-8)  (Q10)
            const int a = i % 3, b = j % 3, s = a + b;   // Qw[a][b] depends on a+b only
            double qv;
            if (s == 0) qv = 0.25 * dt4;
            else if (s == 1) qv = 0.5 * dt3;
            else if (s == 2) qv = (a == 1) ? dt2 : 0.5 * dt2;
            else if (s == 3) qv = dt;
            else qv = 1.0;
            v += qv * q_var;
        }
        P[e] = v;
    }
    __syncwarp();
}

// filterpy KalmanFilter.update, Joseph form, H = [I6 0] (Tracking.py:387-393):
//   y = z - x[:6]; S = P[:6,:6] + R; K = P[:, :6] S^-1; x += K y;
//   P = (I-KH) P (I-KH)' + (K R) K'
// then the x[0] nudge of Tracking.py:396-398 when `nudge` is set.
// R: 36 doubles (shared).  ws: per-warp scratch of kWarpScratch doubles.
__device__ __forceinline__ void warp_kf_update(double* x, double* P, const double* z, const double* R, bool nudge,
                                               double nudge_thres, double nudge_gain, double* ws, int lane) {
    double* S = ws + kWsA;
    double* SI = ws + kWsB;
    double* K = ws + kWsK;
    double* M1 = ws + kWsM1;
    double* M2 = ws + kWsM2;
    double* v = ws + kWsV;
    for (int e = lane; e < 36; e += 32) S[e] = P[(e / 6) * 9 + (e % 6)] + R[e];
    if (lane < 6) v[lane] = z[lane] - x[lane];
    __syncwarp();
    warp_inv6(S, SI, ws + kWsAug, lane);
    // K = P[:, :6] SI   (9x6)
    for (int e = lane; e < 54; e += 32) {
        const int i = e / 6, a = e % 6;
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) acc += P[i * 9 + b] * SI[b * 6 + a];
        K[e] = acc;
    }
    __syncwarp();
    // x += K y
    if (lane < 9) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) acc += K[lane * 6 + a] * v[a];
        x[lane] += acc;
    }
    // M1 = (I - K H) P : row i = P[i,:] - sum_{a<6} K[i][a] P[a,:]
    for (int e = lane; e < 81; e += 32) {
        const int i = e / 9, j = e % 9;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double ikh = (k == i ? 1.0 : 0.0) - (k < 6 ? K[i * 6 + k] : 0.0);
            acc += ikh * P[k * 9 + j];
        }
        M1[e] = acc;
    }
    __syncwarp();                                   // every lane is done with SI before M2 overwrites A|B
    // M2 = K R  (9x6)
    for (int e = lane; e < 54; e += 32) {
        const int i = e / 6, a = e % 6;
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) acc += K[i * 6 + b] * R[b * 6 + a];
        M2[e] = acc;
    }
    __syncwarp();
    // P = M1 (I-KH)' + M2 K'
    for (int e = lane; e < 81; e += 32) {
        const int i = e / 9, j = e % 9;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double ikh = (k == j ? 1.0 : 0.0) - (k < 6 ? K[j * 6 + k] : 0.0);
            acc += M1[i * 9 + k] * ikh;
        }
        double acc2 = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) acc2 += M2[i * 6 + a] * K[j * 6 + a];
        P[e] = acc + acc2;
    }
    __syncwarp();
    if (nudge && lane == 0) {
        // `abs(variance.any()) > 0.6` is abs(bool) > 0.6, i.e. true iff z[0] != x[0]  (Q12)
        const double var = z[0] - x[0];
        const double truth = (var != 0.0) ? 1.0 : 0.0;
        if (truth > nudge_thres) x[0] += var * nudge_gain;
    }
    __syncwarp();
}

}  // namespace mmw
