// Float64 linear algebra of the tracker (6x6 / 9x9), laid out for a 128-thread CTA per scene:
//   * element-parallel steps: thread e owns element e of a 9x9 / 9x6 / 6x6 matrix of track j and the CTA walks the
//     scene's tracks; consecutive steps are separated by a block barrier at the call site, so all four warps work
//     on every track instead of one warp per track;
//   * the 6x6 inverses: one HALF warp per matrix (two tracks per warp), a column of [A | I] per lane, no pivot
//     search (the gate matrix and the innovation covariance are symmetric positive definite).
// Every function cites the reference lines it follows; filterpy 1.4.5 semantics are in SURVEY Appendix B.
#pragma once
#include "mmw_internal.cuh"

namespace mmw {

constexpr unsigned kFull = 0xffffffffu;

// Per-thread element coordinates, computed once per kernel: e = threadIdx.x.
struct Elem {
    int e;          // thread index
    int i9, j9;     // row / column of element e of a 9x9 matrix (valid for e < 81)
    int i6, a6;     // row / column of element e of an Rx6 matrix (valid for e < 54; 6x6 for e < 36)
    __device__ __forceinline__ explicit Elem(int tid) : e(tid), i9(tid / 9), j9(tid % 9), i6(tid / 6), a6(tid % 6) {}
};

// ---- Kalman predict (filterpy KalmanFilter.predict with F = KF_F(dt), Q = KF_Q_DISCR(dt); constants.py:195-215,
// Tracking.py:372-385):  x = F x;  P = F P F' + Q.  State order [p(3) v(3) a(3)]: F couples i with i+3 (dt) and
// i+6 (dt^2/2), so element (i, j) of F P F' is a 3 x 3-term sum over P[i + 3a][j + 3b].  Thread e < 81 computes
// its element into a register from the OLD P (kf_predict_elem); after a block barrier the values are stored
// (the caller batches several tracks between the two barriers).
__device__ __forceinline__ double kf_predict_x(const double* x, double dt, int e) {
    double xn = x[e];
    if (e < 6) xn += dt * x[e + 3];
    if (e < 3) xn += (0.5 * (dt * dt)) * x[e + 6];
    return xn;
}
// Q = block_diag(Qw, Qw, Qw) * q_var with Qw = [[dt^4/4, dt^3/2, dt^2/2], [dt^3/2, dt^2, dt], [dt^2/2, dt, 1]] on
// state indices (0-2), (3-5), (6-8): the block order does not match the state order (Q10) -- reproduced.
__device__ __forceinline__ double kf_predict_elem(const double* P, double dt, double q_var, const Elem& t) {
    const double h = 0.5 * (dt * dt);
    const double* p = P + t.e;
    // rows i, i+3, i+6 of P F'
    double v0 = p[0], v1 = 0.0, v2 = 0.0;
    if (t.j9 < 6) v0 += dt * p[3];
    if (t.j9 < 3) v0 += h * p[6];
    if (t.i9 < 6) {
        v1 = p[27];
        if (t.j9 < 6) v1 += dt * p[30];
        if (t.j9 < 3) v1 += h * p[33];
    }
    if (t.i9 < 3) {
        v2 = p[54];
        if (t.j9 < 6) v2 += dt * p[57];
        if (t.j9 < 3) v2 += h * p[60];
    }
    double v = v0 + dt * v1 + h * v2;
    if (t.i9 / 3 == t.j9 / 3) {
        const int a = t.i9 % 3, b = t.j9 % 3, s = a + b;   // Qw[a][b] depends on a + b except the middle entry
        const double dt2 = dt * dt;
        double qv;
        if (s == 0) qv = 0.25 * (dt2 * dt2);
        else if (s == 1) qv = 0.5 * (dt2 * dt);
        else if (s == 2) qv = (a == 1) ? dt2 : 0.5 * dt2;
        else if (s == 3) qv = dt;
        else qv = 1.0;
        v += qv * q_var;
    }
    return v;
}

// ---- 6x6 inverse + determinant, two matrices per warp -------------------------------------------------------
// Half h = lane / 16 of the warp works on matrix h; lane c = lane % 16 < 12 keeps column c of [A | I] in six
// registers.  Gauss-Jordan without row exchanges: A is symmetric positive definite at both call sites (gate matrix
// P[:6,:6] + Rm + G, Tracking.py:551; innovation covariance H P H' + R, filterpy update), so every pivot is
// positive.  The reference inverts with LAPACK (partial pivoting); the results agree to rounding (parity tests hold
// the state to 1e-6 relative and the gate decisions exactly).  Rows are scaled by the pivot's reciprocal; the five
// multipliers of a step travel by shuffle while the reciprocal is being computed.
// A: 36 doubles, row-major (shared memory) or nullptr for an idle half.  On return lane c in [6, 12) of the half
// holds column c - 6 of the inverse in r[0..5]; every lane of the half returns det(A).
__device__ __forceinline__ double inv6_spd_half(const double* A, int lane, double (&r)[6]) {
    const int c = lane & 15;
    const int col = c < 12 ? c : 11;                 // lanes 12..15 shadow lane 11 (results discarded)
#pragma unroll
    for (int i = 0; i < 6; ++i)
        r[i] = (A != nullptr && col < 6) ? A[i * 6 + col] : (col - 6 == i ? 1.0 : 0.0);
    double det = 1.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double piv = __shfl_sync(kFull, r[k], k, 16);
        double f[6];
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i != k) f[i] = __shfl_sync(kFull, r[i], k, 16);     // column k's entry of row i, before the update
        det *= piv;
        const double inv = __drcp_rn(piv);
        r[k] *= inv;
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i != k) r[i] -= f[i] * r[k];
    }
    return det;
}

// ---- Mahalanobis gate (TrackBuffer._calc_dist_fun, Tracking.py:545-572) ---------------------------------------
// Per track the gate needs C^-1, H x and log|det C| (C = P[:6,:6] + Rm + group_disp_est).  They are kept as 28
// doubles: the upper triangle of C^-1, row-major, with the off-diagonal entries doubled (y' C^-1 y in 21 products
// instead of 36), then H x, then log|det C|.
constexpr int kGateWords = 28;
// After inv6_spd_half: lanes [6, 12) of the half store their column of the inverse, lane 0 the log-determinant.
__device__ __forceinline__ void gate_pack_half(double* gp, const double (&r)[6], double det, int lane) {
    const int c = lane & 15;
    if (c >= 6 && c < 12) {
        const int cc = c - 6;
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i <= cc) gp[i * 6 - (i * (i - 1)) / 2 + (cc - i)] = (i == cc) ? r[i] : 2.0 * r[i];
    }
    if (c == 0) gp[27] = log(fabs(det));
}
// d2 = log|det C| + (w - H x)' C^-1 (w - H x)   (Tracking.py:558-560)
__device__ __forceinline__ double gate_score(const double* gp, const double (&w)[6]) {
    double y[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) y[k] = w[k] - gp[21 + k];
    double q = 0.0;
    int p = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double tb = gp[p++] * y[r];
#pragma unroll
        for (int cc = r + 1; cc < 6; ++cc) tb += gp[p++] * y[cc];
        q += tb * y[r];
    }
    return gp[27] + q;
}

// ---- Kalman update (filterpy KalmanFilter.update, Joseph form, H = [I6 0]; Tracking.py:387-398) --------------
//   y = z - x[:6];  S = P[:6,:6] + R;  K = P[:, :6] S^-1;  x += K y;
//   P = (I - K H) P (I - K H)' + (K R) K'
// as element-parallel steps over a track's scratch:  S 36 | R 36  and  M1 81 | K 54 | KR 54.  The functions return
// the thread's element; the caller stores it (and batches several tracks per step so that their loads overlap).
// step 3 (e < 54): K = P[:, :6] S^-1
__device__ __forceinline__ double kf_update_K_elem(const double* P, const double* Sinv, const Elem& t) {
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) acc += P[t.i6 * 9 + b] * Sinv[b * 6 + t.a6];
    return acc;
}
// step 4 (e < 81): M1 = (I - K H) P = P - K P[:6, :]
__device__ __forceinline__ double kf_update_M1_elem(const double* P, const double* K, const Elem& t) {
    double acc = P[t.e];
#pragma unroll
    for (int a = 0; a < 6; ++a) acc -= K[t.i9 * 6 + a] * P[a * 9 + t.j9];
    return acc;
}
// step 4 (e < 54): KR = K R
__device__ __forceinline__ double kf_update_KR_elem(const double* K, const double* R, const Elem& t) {
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) acc += K[t.i6 * 6 + b] * R[b * 6 + t.a6];
    return acc;
}
// x += K (z - x[:6]), row i < 9: returns the new x[i] (the caller writes it after every lane has read the old x).
__device__ __forceinline__ double kf_update_x(const double* x, const double* z, const double* K, int i) {
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) acc += K[i * 6 + a] * (z[a] - x[a]);
    return acc + x[i];
}
// step 5 (e < 81): P = M1 (I - K H)' + KR K' = M1 - M1[:, :6] K' + KR K'
__device__ __forceinline__ double kf_update_P_elem(const double* M1, const double* K, const double* KR, const Elem& t) {
    double acc = M1[t.e];
#pragma unroll
    for (int a = 0; a < 6; ++a) acc -= M1[t.i9 * 9 + a] * K[t.j9 * 6 + a];
    double acc2 = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) acc2 += KR[t.i9 * 6 + a] * K[t.j9 * 6 + a];
    return acc + acc2;
}
// The x[0] nudge of Tracking.py:396-398: `abs(variance.any()) > 0.6` is abs(bool) > 0.6, i.e. true iff
// z[0] != x[0] (Q12); applied when the track got points this frame.
__device__ __forceinline__ void kf_update_nudge(double* x, const double* z, bool lifetime_zero, double nudge_thres,
                                                double nudge_gain) {
    if (!lifetime_zero) return;
    const double var = z[0] - x[0];
    const double truth = (var != 0.0) ? 1.0 : 0.0;
    if (truth > nudge_thres) x[0] += var * nudge_gain;
}

}  // namespace mmw
