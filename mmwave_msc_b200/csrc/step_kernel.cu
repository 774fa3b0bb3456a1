// The fused per-frame tracker step: one CTA per scene runs
//   Utils.normalize_data (Utils.py:342-434) -> TrackBuffer.track (Tracking.py:664-703)
// i.e. transform+bounds filter+ordered compaction, Kalman predict, Mahalanobis gating and association,
// per-track cluster statistics and ring push, track maintenance, Kalman update, push of the unassigned
// points into the scene's global ring, DBSCAN over the fused ring, and track spawning -- with every
// intermediate in shared memory / registers.  HBM traffic per scene-frame is the raw points in, the track
// records in and out, the ring rows written and the fused ring read when DBSCAN runs.
#include "dbscan.cuh"
#include "linalg.cuh"
#include "mmw_internal.cuh"

namespace mmw {

struct SmemLayout {
    // offsets in bytes into dynamic shared memory (kept under 32 KB at N_cap 256 / T_cap 8: 7 CTAs per SM, so the
    // 1024 scenes of the C2 workload are resident in a single wave)
    int vel;         // double [3][ncap]: world-frame velocity columns of the compacted frame (sqrt/div: kept);
                     //   positions are recomputed from craw on use (x is the raw x; y', z' cost 4 mul + 3 add)
    int craw;        // float [ncap*5] compacted raw points of this frame
    int assoc;       // uint8 [ncap]: group of each compacted point = track index, or n_tracks for unassigned
    int dbf;         // float [3][3*ncap]: fp32 world coordinates of the fused ring during DBSCAN; aliases
                     //   vel|craw|assoc, which are dead by then
    int tracks;      // TrackRec [tcap]
    int cinv;        // double [tcap][36]
    int hx;          // double [tcap][6]
    int logdet;      // double [tcap]
    int ws;          // double [warps][kWarpScratch]
    int par, cl;     // int [3*ncap] each: alias ws (dead during DBSCAN/spawn) when they fit, else their own space
    int misc;        // ints
    int total;
};

__host__ __device__ inline SmemLayout make_layout(int ncap, int tcap) {
    SmemLayout L;
    const int n4 = (ncap + 3) & ~3;
    int o = 0;
    L.vel = o;     o += 3 * n4 * 8;
    L.craw = o;    o += n4 * 5 * 4;
    L.assoc = o;   o += n4;
    L.dbf = 0;
    if (o < 36 * n4) o = 36 * n4;                       // (never: 24 + 20 + 1 = 45 bytes per point)
    o = (o + 15) & ~15;
    L.tracks = o;  o += tcap * (int)sizeof(TrackRec);
    L.cinv = o;    o += tcap * 36 * 8;
    L.hx = o;      o += tcap * 6 * 8;
    L.logdet = o;  o += ((tcap + 1) & ~1) * 8;
    L.ws = o;
    const int wsb = kStepWarps * kWarpScratch * 8, dbb = 2 * 3 * n4 * 4;
    o += wsb > dbb ? wsb : dbb;
    L.par = L.ws;  L.cl = L.ws + 3 * n4 * 4;
    L.misc = o;    o += 128 * 4;
    L.total = o;
    return L;
}

int step_smem_bytes(int ncap, int tcap) { return make_layout(ncap, tcap).total; }

// misc int slots
enum { kM = 0, kNFree = 1, kScan = 4 /* .. +kStepWarps+1 */, kOrder = 16 /* .. +32 */,
       kFree = 48 /* .. +32 */ };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// Reduce NV per-lane values across the warp so that lane v (< NV) ends up with the total of value v: at every
// butterfly level a lane keeps one half of its live values and hands the other half to its partner, so the 5 levels
// cost about NV + log2 shuffles instead of 5 * NV.  The order of the additions is fixed (deterministic results).
template <int NV, class Op>
__device__ __forceinline__ double warp_reduce_to_lane(double (&a)[NV], int lane, Op op) {
    static_assert(NV >= 1 && NV <= 32, "one value per lane at most");
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
            if (k < NV) {
                if (k + o < NV) {
                    const double keep = up ? a[k + o] : a[k];
                    const double send = up ? a[k] : a[k + o];
                    a[k] = op(keep, __shfl_xor_sync(kFull, send, o));
                } else {
                    a[k] = op(a[k], __shfl_xor_sync(kFull, a[k], o));
                }
            }
        }
    }
    return a[0];
}

__device__ __forceinline__ const float* uring_frame(const StepArgs& a, int s, int phys) {
    return a.uring + ((size_t)s * kRing + phys) * (size_t)a.cfg.ncap * kRawCols;
}

// One warp turns DBSCAN cluster q of the fused ring into a fresh track (ClusterTrack.__init__, Tracking.py:210-230
// over PointCluster(cluster), :120-136): statistics over the 6 world columns recomputed from the ring's raw rows,
// state x = [centroid, 0, 0, 0], P = KF_P_INIT*I, group dispersion = KF_GROUP_DISP_EST_INIT*I, ring = [first 64
// cluster rows in fused order], keypoints = MODEL_DEFAULT_POSTURE.  `t` may live in shared or global memory.
__device__ __forceinline__ void spawn_track(const StepArgs& a, int s, TrackRec& t, int q, int slot, int track_id,
                                            const int* cl, const int* fcnt, const int* fphys, int lane,
                                            const float* rawc = nullptr /* fused raw rows in shared memory */,
                                            const double* w6 = nullptr /* their world 6-vectors, precomputed */) {
    const DevConfig& c = a.cfg;
    const int tcap = c.tcap;
    int n = 0;
    double sacc[6], mn[6], mx[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { sacc[k] = 0.0; mn[k] = INFINITY; mx[k] = -INFINITY; }
    int b0 = 0;
    float* dst = a.track_ring + (((size_t)s * tcap + slot) * kRing + 0) * (kFeatPts * kRawCols);
    int stored = 0;
    for (int f = 0; f < kRing; ++f) {
        const float* src = rawc != nullptr ? rawc + (size_t)b0 * kRawCols : uring_frame(a, s, fphys[f]);
        for (int i0 = 0; i0 < fcnt[f]; i0 += 32) {
            const int i = i0 + lane;
            const bool in = i < fcnt[f] && cl[b0 + i] == q;
            float r5[kRawCols];
            if (in) {
#pragma unroll
                for (int k = 0; k < kRawCols; ++k) r5[k] = src[i * kRawCols + k];
                double w[6];
                if (w6 != nullptr) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) w[k] = w6[(size_t)(b0 + i) * 6 + k];
                } else {
                    world_from_raw(c, r5[0], r5[1], r5[2], r5[3], w);
                }
                ++n;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    sacc[k] += w[k]; mn[k] = fmin(mn[k], w[k]); mx[k] = fmax(mx[k], w[k]);
                }
            }
            const unsigned m = __ballot_sync(kFull, in);
            const int pos = stored + __popc(m & ((1u << lane) - 1u));
            if (in && pos < kFeatPts) {
#pragma unroll
                for (int k = 0; k < kRawCols; ++k) dst[pos * kRawCols + k] = r5[k];
            }
            stored += __popc(m);
        }
        b0 += fcnt[f];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(kFull, n, o);
    double cen[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        cen[k] = warp_sum(sacc[k]) / (double)n;
        mn[k] = warp_min(mn[k]);
        mx[k] = warp_max(mx[k]);
    }
    for (int e = lane; e < 81; e += 32) t.P[e] = (e / 9 == e % 9) ? c.p_init : 0.0;
    for (int e = lane; e < 36; e += 32) t.G[e] = (e / 6 == e % 6) ? c.g_init : 0.0;
    if (lane >= 6 && lane < 9) t.x[lane] = 0.0;
    if (lane < 6) {
        t.x[lane] = cen[lane];
        t.centroid[lane] = cen[lane];
        t.minv[lane] = mn[lane];
        t.maxv[lane] = mx[lane];
        t.spread[lane] = 0.0;
    }
    if (lane == 0) {
        t.n_est = 0.0;
        t.lifetime = 0.0;
        t.id = track_id;
        t.point_num = n;
        t.is_static = sqrt(cen[3] * cen[3] + cen[4] * cen[4] + cen[5] * cen[5]) < c.vel_thres ? 1 : 0;
        t.slot = slot;
        t.ring_n = 1;
        t.ring_head = 0;
        t.ring_cnt[0] = n < kFeatPts ? n : kFeatPts;
        t.ring_cnt[1] = 0;
        t.ring_cnt[2] = 0;
        t.pad[0] = t.pad[1] = t.pad[2] = 0;
    }
    // keypoints = MODEL_DEFAULT_POSTURE until the first inference (Tracking.py:221, Q25)
    float* kp = a.keypoints + ((size_t)s * tcap + slot) * kKp;
    for (int e = lane; e < kKp; e += 32) kp[e] = a.default_posture[e];
}

// fp32 world coordinates of the fused ring (oldest frame first) into shared memory + the screened predicate over
// them (dbscan.cuh).  Block-cooperative; ends with __syncthreads().
__device__ __forceinline__ NbScreened load_fused_ring(const StepArgs& a, int s, const int* fcnt, const int* fphys,
                                                      float* Xf, int stride, float* rawc = nullptr) {
    const DevConfig& c = a.cfg;
    float* Yf = Xf + stride;
    float* Zf = Yf + stride;
    NbScreened nb{c, Xf, Yf, Zf, {nullptr, nullptr, nullptr}, {0, 0, 0, 0}, c.db_eps, 0.f, 0.f,
                  (float)c.db_range_weight, (float)c.db_z_weight, rawc};
    const float band = 1e-3f * (float)c.db_eps + 1e-4f;
    nb.lo = (float)c.db_eps - band;
    nb.hi = (float)c.db_eps + band;
    int b0 = 0;
    for (int f = 0; f < kRing; ++f) {
        const float* src = uring_frame(a, s, fphys[f]);
        nb.frame[f] = src;
        nb.start[f] = b0;
        for (int i = threadIdx.x; i < fcnt[f]; i += blockDim.x) {
            const float x = src[i * kRawCols + 0], y = src[i * kRawCols + 1], z = src[i * kRawCols + 2];
            double yw, zw;
            world_yz(c, (double)y, (double)z, yw, zw);
            Xf[b0 + i] = x; Yf[b0 + i] = (float)yw; Zf[b0 + i] = (float)zw;
            if (rawc != nullptr) {
                float* r = rawc + (size_t)(b0 + i) * kRawCols;
                r[0] = x; r[1] = y; r[2] = z; r[3] = src[i * kRawCols + 3]; r[4] = src[i * kRawCols + 4];
            }
        }
        b0 += fcnt[f];
    }
    nb.start[kRing] = b0;
    __syncthreads();
    return nb;
}


// Optional per-phase cycle accounting (thread 0 of every CTA, accumulated with one atomic per phase).
#define PHASE_MARK(idx)                                                                  \
    do {                                                                                 \
        if (a.phase_cycles != nullptr && threadIdx.x == 0) {                             \
            const long long now__ = clock64();                                           \
            atomicAdd(&a.phase_cycles[idx], (unsigned long long)(now__ - phase_t0));     \
            phase_t0 = now__;                                                            \
        }                                                                                \
    } while (0)

#ifndef MMW_STEP_MINBLOCKS
#define MMW_STEP_MINBLOCKS 7
#endif
__global__ void __launch_bounds__(kStepThreads, MMW_STEP_MINBLOCKS) step_kernel(const __grid_constant__ StepArgs a) {
    long long phase_t0 = clock64();
    const long long kernel_t0 = phase_t0;
    if (a.phase_cycles != nullptr && threadIdx.x == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        a.phase_cycles[16 + gridDim.x + blockIdx.x] = ns;
    }
    extern __shared__ __align__(16) unsigned char smem[];
    const DevConfig& c = a.cfg;
    const int ncap = c.ncap, tcap = c.tcap;
    const SmemLayout L = make_layout(ncap, tcap);
    double* vel = reinterpret_cast<double*>(smem + L.vel);
    float* craw = reinterpret_cast<float*>(smem + L.craw);
    uint8_t* assoc = smem + L.assoc;
    // world-frame 6-vector of compacted point i: positions recomputed (bit-identical to step 1), velocities stored
    auto load_w = [&](int i, double (&w)[6]) {
        w[0] = (double)craw[i * kRawCols + 0];
        world_yz(c, (double)craw[i * kRawCols + 1], (double)craw[i * kRawCols + 2], w[1], w[2]);
        w[3] = vel[i]; w[4] = vel[ncap + i]; w[5] = vel[2 * ncap + i];
    };
    int* par = reinterpret_cast<int*>(smem + L.par);
    int* cl = reinterpret_cast<int*>(smem + L.cl);
    TrackRec* tr = reinterpret_cast<TrackRec*>(smem + L.tracks);
    double* cinv = reinterpret_cast<double*>(smem + L.cinv);
    double* hx = reinterpret_cast<double*>(smem + L.hx);
    double* logdet = reinterpret_cast<double*>(smem + L.logdet);
    double* wsall = reinterpret_cast<double*>(smem + L.ws);
    int* misc = reinterpret_cast<int*>(smem + L.misc);

    const int s = blockIdx.x;
    if (s >= a.n_scenes) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* ws = wsall + warp * kWarpScratch;

    SceneRec sc = a.scenes[s];           // every thread keeps a copy; thread 0 writes it back
    const int off = a.offsets[s];
    int N = a.offsets[s + 1] - off;
    if (N > ncap) { N = ncap; sc.flags |= MMW_SCENE_POINT_OVERFLOW; }
    const double dt = a.dt[s];

    // Issued first, consumed later: this scene's track records (cp.async -> shared memory, waited for in step 2)
    // and the ring frames DBSCAN will read (L2 prefetch), so their DRAM latency overlaps the point phase.
    const int T0 = sc.n_tracks;
    {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.tracks + (size_t)s * tcap);
        unsigned char* dst = reinterpret_cast<unsigned char*>(tr);
#ifndef MMW_NO_CPASYNC
        for (int i = tid; i < T0 * (int)(sizeof(TrackRec) / 16); i += kStepThreads)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + i * 16)),
                         "l"(src + i * 16)
                         : "memory");
#else
        (void)src; (void)dst;
#endif
        for (int f = 0; f < sc.ring_n; ++f) {
            const int phys = ring_wrap(sc.ring_head + f, c.ring_size);
            const char* fr = reinterpret_cast<const char*>(uring_frame(a, s, phys));
            const int lines = (sc.ring_cnt[phys] * kRawCols * 4 + 127) / 128;
            for (int i = tid; i < lines; i += kStepThreads)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(fr + i * 128));
        }
    }

    // ---- 1. load, transform, bounds filter, ordered compaction (Utils.py:379-432) ------------------
    for (int i = tid; i < N * kRawCols; i += kStepThreads) craw[i] = a.pts[(size_t)off * kRawCols + i];
    __syncthreads();
    int M = 0;
    for (int base = 0; base < N; base += kStepThreads) {
        const int i = base + tid;
        float r5[kRawCols];
        double w[6];
        bool keep = false;
        if (i < N) {
#pragma unroll
            for (int k = 0; k < kRawCols; ++k) r5[k] = craw[i * kRawCols + k];
            world_from_raw(c, r5[0], r5[1], r5[2], r5[3], w);
            keep = (w[2] <= c.z_max) && (w[2] > 0.0) && (w[1] > 0.0);      // Utils.py:423-427
        }
        int tot;
        const int pos = M + block_rank(keep, &tot, misc + kScan);   // syncs: all rows of this chunk are in registers
        if (keep) {
#pragma unroll
            for (int k = 0; k < kRawCols; ++k) craw[pos * kRawCols + k] = r5[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) vel[k * ncap + pos] = w[3 + k];
        }
        M += tot;
        __syncthreads();
    }
    for (int i = M + tid; i < N; i += kStepThreads) a.assoc_out[off + i] = -2;

    sc.last_M = M;
    sc.dbscan_n = -1;
    if (M == 0) {                        // offline_main.py:55: the frame is skipped entirely (Q23)
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (tid == 0) {
            sc.last_ran = 0;
            a.scenes[s] = sc;
            a.pose_cnt[s] = 0;
            atomicAdd(&a.counters[1], (unsigned long long)N);
        }
        return;
    }
    sc.last_ran = 1;

    PHASE_MARK(1);
    // ---- 2. load this scene's track records -----------------------------------------------------------
#ifndef MMW_NO_CPASYNC
    asm volatile("cp.async.wait_all;" ::: "memory");
#else
    {
        const double* src = reinterpret_cast<const double*>(a.tracks + (size_t)s * tcap);
        double* dst = reinterpret_cast<double*>(tr);
        for (int i = tid; i < T0 * kTrackWords; i += kStepThreads) dst[i] = src[i];
    }
#endif
    __syncthreads();

    PHASE_MARK(2);
    // ---- 3. predict (Tracking.py:591-596, Q9) and gate matrices (Tracking.py:545-551) ---------------
    for (int j = warp; j < T0; j += kStepWarps) {
        TrackRec& t = tr[j];
        warp_kf_predict(t.x, t.P, t.lifetime + dt, c.q_var, ws + kWsM1, lane);
        double* C = ws + kWsA;
        for (int e = lane; e < 36; e += 32) {
            const int r = e / 6, q = e % 6;
            double v = t.P[r * 9 + q];
            if (r == q) { const double h = t.spread[r] / 2; v += h * h; }   // get_Rm (361-370)
            C[e] = v + t.G[e];
        }
        if (lane < 6) hx[j * 6 + lane] = t.x[lane];
        __syncwarp();
        const double det = warp_inv6(C, cinv + j * 36, ws + kWsAug, lane);
        if (lane == 0) logdet[j] = log(fabs(det));
    }
    __syncthreads();

    PHASE_MARK(3);
    // ---- 4. gating + association (Tracking.py:553-572, Q14) -------------------------------------------
    // Besides the decision per point this phase builds, in input order, the index list of every group (group j < T0:
    // the points gated into track j; then the unassigned points).  Warp w owns a contiguous run of 32-point chunks
    // and counts its points per group (lane j keeps track j's count); after one barrier every warp knows where its
    // run starts inside each list, places its indices and pushes the raw rows into the rings (Tracking.py:691, 338)
    // with one thread per point.  The per-track statistics below then walk a compact list.
    uint16_t* lst = reinterpret_cast<uint16_t*>(wsall);              // [M]; the warp scratch is idle in phases 4-6
    int* wtot = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(wsall) + 2 * ((ncap + 3) & ~3));   // [33][warps]
    const unsigned ltmask = (1u << lane) - 1u;
    const int nchunk = (M + 31) >> 5, cpw = (nchunk + kStepWarps - 1) / kStepWarps;
    const int ch0 = warp * cpw, ch1 = min(nchunk, ch0 + cpw);
    {
        int cntj = 0, cntu = 0;
        for (int ch = ch0; ch < ch1; ++ch) {
            const int i = ch * 32 + lane;
            int g = 255;
            if (i < M) {
                double p[6];
                load_w(i, p);
                double best = INFINITY;
                int bj = -1;
                for (int j = 0; j < T0; ++j) {
                    double y[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) y[k] = p[k] - hx[j * 6 + k];
                    const double* Ci = cinv + j * 36;
                    double q = 0.0;
#pragma unroll
                    for (int b = 0; b < 6; ++b) {
                        double tb = 0.0;
#pragma unroll
                        for (int k = 0; k < 6; ++k) tb += y[k] * Ci[k * 6 + b];
                        q += tb * y[b];
                    }
                    const double d2 = logdet[j] + q;
                    if (d2 < c.gate && d2 < best) { best = d2; bj = j; }
                }
                g = bj < 0 ? T0 : bj;
                assoc[i] = (uint8_t)g;
                a.assoc_out[off + i] = bj;
            }
            for (int j = 0; j < T0; ++j) {
                const unsigned b = __ballot_sync(kFull, g == j);
                if (lane == j) cntj += __popc(b);
            }
            cntu += __popc(__ballot_sync(kFull, g == T0));
        }
        if (lane < T0) wtot[lane * kStepWarps + warp] = cntj;
        if (lane == 0) wtot[kMaxTcap * kStepWarps + warp] = cntu;
    }
    __syncthreads();
    // lane j: size and start of track j's list, and where this warp's run begins inside it
    int totj = 0, startj, runj = 0, U, runu = 0;
    {
        if (lane < T0)
            for (int w = 0; w < kStepWarps; ++w) {
                const int cw = wtot[lane * kStepWarps + w];
                totj += cw;
                if (w < warp) runj += cw;
            }
        int incl = totj;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += up;
        }
        startj = incl - totj;
        runj += startj;
        U = 0;
        for (int w = 0; w < kStepWarps; ++w) {
            const int cw = wtot[kMaxTcap * kStepWarps + w];
            U += cw;
            if (w < warp) runu += cw;
        }
        runu += __shfl_sync(kFull, incl, 31);                         // the unassigned list follows the tracks' lists
    }
    const int uphys = sc.ring_n >= c.ring_size ? sc.ring_head : ring_wrap(sc.ring_head + sc.ring_n, c.ring_size);
    {
        // BatchedData.add_frame(unassigned): the oldest frame is dropped when the ring is full
        float* udst = const_cast<float*>(uring_frame(a, s, uphys));
        const int ustart = M - U;
        for (int ch = ch0; ch < ch1; ++ch) {
            const int i = ch * 32 + lane;
            const int g = i < M ? (int)assoc[i] : 255;
            int pos = 0;
            for (int j = 0; j < T0; ++j) {
                const unsigned b = __ballot_sync(kFull, g == j);
                const int basej = __shfl_sync(kFull, runj, j);
                if (g == j) pos = basej + __popc(b & ltmask);
                if (lane == j) runj += __popc(b);
            }
            {
                const unsigned b = __ballot_sync(kFull, g == T0);
                if (g == T0) pos = runu + __popc(b & ltmask);
                runu += __popc(b);
            }
            const int gstart = __shfl_sync(kFull, startj, g < 32 ? g : 0);
            if (i < M) {
                lst[pos] = (uint16_t)i;
                float* dst = nullptr;
                if (g == T0) {
                    dst = udst + (size_t)(pos - ustart) * kRawCols;
                } else if (pos - gstart < kFeatPts) {
                    // first 64 associated rows in input order are what format_single_frame can see
                    const TrackRec& t = tr[g];
                    const int phys = t.ring_n >= c.ring_size ? t.ring_head : ring_wrap(t.ring_head + t.ring_n, c.ring_size);
                    dst = a.track_ring + (((size_t)s * tcap + t.slot) * kRing + phys) * (kFeatPts * kRawCols) +
                          (size_t)(pos - gstart) * kRawCols;
                }
                if (dst != nullptr) {
#pragma unroll
                    for (int k = 0; k < kRawCols; ++k) dst[k] = craw[i * kRawCols + k];
                }
            }
        }
    }
    __syncthreads();

    PHASE_MARK(4);
    // ---- 5. per-track association (Tracking.py:648-653, 314-341) ---------------------------------------
    int ring_rows = 0;
    for (int g = warp; g < T0; g += kStepWarps) {
        TrackRec& t = tr[g];
        const int n = __shfl_sync(kFull, totj, g), start = __shfl_sync(kFull, startj, g);
        if (n == 0) {
            if (lane == 0) t.lifetime += dt;                     // update_lifetime(dt) (400-407)
            continue;
        }
        // PointCluster statistics (Tracking.py:120-136): sums in a[0..5]; minima and negated maxima share one array
        double cen[6], mnv, mxv;
        {
            double sa[6], mm[12];
#pragma unroll
            for (int k = 0; k < 6; ++k) { sa[k] = 0.0; mm[k] = INFINITY; mm[6 + k] = INFINITY; }
            for (int k0 = lane; k0 < n; k0 += 32) {
                double w[6];
                load_w((int)lst[start + k0], w);
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    sa[k] += w[k];
                    mm[k] = w[k] < mm[k] ? w[k] : mm[k];             // plain compare-select: no NaNs to honour here
                    mm[6 + k] = -w[k] < mm[6 + k] ? -w[k] : mm[6 + k];
                }
            }
            const double ssum = warp_reduce_to_lane(sa, lane, [](double u, double v) { return u + v; });
            const double smin = warp_reduce_to_lane(mm, lane, [](double u, double v) { return v < u ? v : u; });
#pragma unroll
            for (int k = 0; k < 6; ++k) cen[k] = __shfl_sync(kFull, ssum, k) / (double)n;
            mnv = smin;                                          // lane m < 6: min of column m
            mxv = -__shfl_sync(kFull, smin, (lane + 6) & 31);    // lane m < 6: max of column m
        }
        // dispersion matrix about the centroid (population covariance, _get_D 270-290); entry p lands in lane p
        double cov;
        {
            double acc[21];
#pragma unroll
            for (int p = 0; p < 21; ++p) acc[p] = 0.0;
            for (int k0 = lane; k0 < n; k0 += 32) {
                double d[6];
                load_w((int)lst[start + k0], d);
#pragma unroll
                for (int k = 0; k < 6; ++k) d[k] -= cen[k];
                int p = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int q = r; q < 6; ++q) acc[p++] += d[r] * d[q];
            }
            cov = warp_reduce_to_lane(acc, lane, [](double u, double v) { return u + v; }) / (double)n;
        }
        const int phys = t.ring_n >= c.ring_size ? t.ring_head : ring_wrap(t.ring_head + t.ring_n, c.ring_size);
        double n_est = t.n_est;
        if (c.enable_est) {                                      // _estimate_point_num (232-244)
            if ((double)n > n_est) n_est = (double)n;
            else n_est = (1 - c.a_n) * n_est + c.a_n * (double)n;
        } else {
            n_est = (double)(n > c.est_pointnum ? n : c.est_pointnum);
        }
        __syncwarp();
        if (lane == 0) {
            if (t.ring_n >= c.ring_size) t.ring_head = ring_wrap(t.ring_head + 1, c.ring_size);
            else t.ring_n += 1;
            t.ring_cnt[phys] = n < kFeatPts ? n : kFeatPts;
            t.lifetime = 0.0;
            t.point_num = n;
            t.is_static = sqrt(cen[3] * cen[3] + cen[4] * cen[4] + cen[5] * cen[5]) < c.vel_thres ? 1 : 0;
            t.n_est = n_est;
        }
        if (lane < 6) {                                          // _estimate_measurement_spread (246-268)
            const int m = lane;
            double cm = cen[0];
#pragma unroll
            for (int k = 1; k < 6; ++k) cm = (m == k) ? cen[k] : cm;
            t.centroid[m] = cm;
            t.minv[m] = mnv;
            t.maxv[m] = mxv;
            double spread = mxv - mnv;
            if (n != 1) spread = spread * (double)(n + 1) / (double)(n - 1);
            spread = fmin(2 * c.spread_lim[m], spread);
            spread = fmax(c.spread_lim[m], spread);
            if (spread > t.spread[m]) t.spread[m] = spread;
            else t.spread[m] = (1.0 - c.a_spr) * t.spread[m] + c.a_spr * spread;
        }
        {                                                        // _estimate_group_disp_matrix (292-297)
            const double al = (double)n / n_est;
#pragma unroll
            for (int e0 = 0; e0 < 64; e0 += 32) {
                const int e = e0 + lane;
                int r = e / 6, q = e % 6;
                if (r > q) { const int tmp = r; r = q; q = tmp; }
                const int p = e < 36 ? r * 6 - (r * (r - 1)) / 2 + (q - r) : 0;
                const double dv = __shfl_sync(kFull, cov, p);
                if (e < 36) t.G[e] = (1 - al) * t.G[e] + al * dv;
            }
        }
        if (lane == 0) ring_rows += (n < kFeatPts ? n : kFeatPts);
        __syncwarp();
    }
    __syncthreads();
    {
        if (sc.ring_n >= c.ring_size) sc.ring_head = ring_wrap(sc.ring_head + 1, c.ring_size);
        else sc.ring_n += 1;
        sc.ring_cnt[uphys] = U;
    }

    PHASE_MARK(5);
    // ---- 6. maintenance (Tracking.py:513-528, Q16): drop timed-out tracks, keep list order -----------
    if (tid == 0) {
        int nk = 0, nf = 0;
        for (int k = 0; k < T0; ++k) {
            const double lim = tr[k].is_static ? c.life_sta : c.life_dyn;
            if (tr[k].lifetime > lim) misc[kFree + nf++] = tr[k].slot;     // INACTIVE: dropped from the list
            else misc[kOrder + nk++] = k;
        }
        misc[kM] = nk;
        misc[kNFree] = nf;
    }
    __syncthreads();
    const int T1 = misc[kM];
    for (int k = 0; k < misc[kNFree]; ++k) sc.slot_mask &= ~(1u << misc[kFree + k]);

    PHASE_MARK(6);
    // ---- 7. update every surviving track (Tracking.py:598-603, 387-398; Q11-Q13) ---------------------
    for (int k = warp; k < T1; k += kStepWarps) {
        TrackRec& t = tr[misc[kOrder + k]];
        // _get_Rc (299-312): Rm/N + ((N_est - N)/((N_est - 1) N)) * group_disp_est
        double* Rc = cinv + misc[kOrder + k] * 36;   // the gate matrices are dead by now: reuse as R storage
        const double Nn = (double)t.point_num;
        const double coef = div_zero_fast(t.n_est - Nn, (t.n_est - 1.0) * Nn);     // N_est == N: zero numerator
        for (int e = lane; e < 36; e += 32) {
            const int r = e / 6, q = e % 6;
            double rm = 0.0;
            if (r == q) { const double h = t.spread[r] / 2; rm = h * h; }
            Rc[e] = div_zero_fast(rm, Nn) + coef * t.G[e];                        // rm is zero off the diagonal
        }
        __syncwarp();
        warp_kf_update(t.x, t.P, t.centroid, Rc, t.lifetime == 0.0, c.nudge_thres, c.nudge_gain, ws, lane);
    }
    __syncthreads();

    PHASE_MARK(7);
    // ---- 8. DBSCAN over the fused global ring (Tracking.py:693-700) -----------------------------------
    int B = 0;
    int fcnt[kRing], fphys[kRing];
    for (int f = 0; f < kRing; ++f) {
        fphys[f] = ring_wrap(sc.ring_head + f, c.ring_size);
        fcnt[f] = f < sc.ring_n ? sc.ring_cnt[fphys[f]] : 0;
        B += fcnt[f];
    }
    int ncl = 0;
    const bool run_db = B > 0 && T1 < c.tr_max_tracks;
    // Big fused clouds (someone walked in: hundreds of points) are rare -- a couple of scenes per frame -- but cost
    // several times a normal scene-frame and would set this kernel's duration.  They are handed to
    // dbscan_big_kernel (1024 threads per scene) through a work list; everything else is clustered right here.
    const bool defer = run_db && a.defer_list != nullptr && B > kDeferPoints;
    bool deferred_late = false;
    if (defer) {
        if (tid == 0) a.defer_list[atomicAdd(a.defer_count, 1)] = s;
        sc.dbscan_n = B;
    } else if (run_db) {
        NbScreened nb = load_fused_ring(a, s, fcnt, fphys, reinterpret_cast<float*>(smem + L.dbf), 3 * ncap);
        PHASE_MARK(11);
        // Grid screen first (dbscan.cuh): the residue of a steady scene cannot hold a core point and is all noise.
        // Whatever the screen lets through is most likely a cluster forming and goes straight to dbscan_big_kernel.
        const bool screen = a.defer_list != nullptr && 3 * ncap >= kGridCells;
        if (screen && !dbscan_grid_may_have_core(c, nb.Xf, nb.Yf, B, c.db_min_samples, cl, misc + kScan)) {
            // (the histogram aliased cl; the screen ends with a barrier, and cl[b] is read back by its writer only)
            for (int b = tid; b < B; b += kStepThreads) cl[b] = -1;
            ncl = 0;
        } else if (screen) {
            ncl = -1;
        } else {
            ncl = dbscan_block(nb, B, c.db_min_samples, par, cl, misc + kScan, a.phase_cycles, a.defer_list != nullptr);
        }
        PHASE_MARK(12);
        if (ncl < 0) {
            // A cluster forms in this scene (a couple of scenes per frame): components, border labels and the spawn
            // happen in dbscan_big_kernel.  Finishing them here from the bit rows of the core points was measured:
            // correct, but those CTAs (4 warps sharing an SM with 6 other CTAs) then end 35 us after all others and
            // the step kernel goes from 72 to 107 us -- the separate 512-thread kernel costs 29 us.
            if (tid == 0) a.defer_list[atomicAdd(a.defer_count, 1)] = s;
            ncl = 0;
            deferred_late = true;
        }
        sc.dbscan_n = B;
        if (a.labels_out != nullptr && !deferred_late)
            for (int b = tid; b < B; b += kStepThreads) a.labels_out[(size_t)s * 3 * ncap + b] = cl[b];
    }

    PHASE_MARK(8);
    // ---- 9. spawn one track per cluster, in label order (Tracking.py:576-589, 210-230; Q7, Q20) ------
    int T2 = T1;
    if (ncl > 0) {
        if (T1 + ncl > tcap) { ncl = tcap - T1; sc.flags |= MMW_SCENE_TRACK_OVERFLOW; }
        // smem record indices not used by survivors, physical slots not in use (all threads compute the same)
        unsigned used = 0;
        for (int k = 0; k < T1; ++k) used |= 1u << misc[kOrder + k];
        int newidx[kMaxTcap], newslot[kMaxTcap];
        {
            const unsigned capmask = tcap >= 32 ? 0xffffffffu : ((1u << tcap) - 1u);
            unsigned fr = ~used & capmask, fs = ~sc.slot_mask & capmask;
            for (int q = 0; q < ncl; ++q) {
                newidx[q] = __ffs(fr) - 1; fr &= fr - 1;
                newslot[q] = __ffs(fs) - 1; fs &= fs - 1;
                sc.slot_mask |= 1u << newslot[q];
            }
        }
        for (int q = warp; q < ncl; q += kStepWarps) {
            spawn_track(a, s, tr[newidx[q]], q, newslot[q], sc.next_id + q, cl, fcnt, fphys, lane);
            if (lane == 0) misc[kOrder + T1 + q] = newidx[q];
        }
        sc.next_id += ncl;
        T2 = T1 + ncl;
        sc.ring_n = 0;                   // batch.clear() (Tracking.py:699-700, Q8)
        sc.ring_head = 0;
        sc.ring_cnt[0] = sc.ring_cnt[1] = sc.ring_cnt[2] = 0;
    }
    __syncthreads();

    PHASE_MARK(9);
    pdl_launch_dependents();             // late on purpose: dbscan_big's CTAs would otherwise sit on 16 SMs for the whole step
    // ---- 10. write back: track records in list order, scene record, counters ---------------------------
    {
        double* dstbase = reinterpret_cast<double*>(a.tracks + (size_t)s * tcap);
        for (int i = tid; i < T2 * kTrackWords; i += kStepThreads) {
            const int k = i / kTrackWords, wd = i % kTrackWords;
            dstbase[i] = reinterpret_cast<const double*>(tr + misc[kOrder + k])[wd];
        }
    }
    // ring rows written by all warps of this CTA
    if (lane == 0 && ring_rows) atomicAdd(&a.counters[6], (unsigned long long)ring_rows);
    if (tid == 0) {
        sc.n_tracks = T2;
        a.scenes[s] = sc;
        a.pose_cnt[s] = T2;
        atomicAdd(&a.counters[0], 1ull);
        atomicAdd(&a.counters[1], (unsigned long long)N);
        atomicAdd(&a.counters[2], (unsigned long long)M);
        atomicAdd(&a.counters[3], (unsigned long long)U);
        if (run_db) atomicAdd(&a.counters[4], (unsigned long long)B);
        atomicAdd(&a.counters[5], (unsigned long long)T2);
    }
    PHASE_MARK(10);
    if (a.phase_cycles != nullptr && threadIdx.x == 0)
        a.phase_cycles[16 + blockIdx.x] = (unsigned long long)(clock64() - kernel_t0);   // last frame's cycles of this scene
    if (a.phase_cycles != nullptr && threadIdx.x == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        a.phase_cycles[16 + 2 * gridDim.x + blockIdx.x] = ns;
    }
}

// DBSCAN + spawn for the scenes the step kernel deferred (fused cloud > kDeferPoints): same device functions, one
// CTA of 1024 threads per scene, new tracks written straight to the scene's list in global memory.
constexpr int kBigThreads = 512;
constexpr int kBitsMaxB = 768;        // adjacency bit matrix of up to 768 x 768 (72 KB of shared memory)
__host__ __device__ inline int big_bits_cap(int ncap) { return 3 * ncap < kBitsMaxB ? 3 * ncap : kBitsMaxB; }
// shared-memory layout of dbscan_big_kernel: Xf|Yf|Zf, par, cl, scan, adj, cm, ws, raw rows, (8-byte aligned) w6
constexpr int kBigWsWords = 4 + kMaxCompMasks;    // per-word scratch rows of dbscan_bits_block
__host__ __device__ inline int big_w6_offset(int ncap) {
    const int cap = big_bits_cap(ncap), w = (cap + 31) / 32;
    const int o = 36 * ncap + 2 * 3 * ncap * 4 + 64 * 4 + (cap * w + (1 + kBigWsWords) * w) * 4 + cap * kRawCols * 4;
    return (o + 7) & ~7;
}
__global__ void __launch_bounds__(kBigThreads, 1) dbscan_big_kernel(const __grid_constant__ StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DevConfig& c = a.cfg;
    const int ncap = c.ncap, tcap = c.tcap;
    float* Xf = reinterpret_cast<float*>(smem);
    int* par = reinterpret_cast<int*>(smem + 36 * ncap);
    int* cl = par + 3 * ncap;
    int* scan = cl + 3 * ncap;                               // 64 ints
    // bit-matrix DBSCAN (dbscan.cuh) for fused clouds of up to kBitsMaxB points; larger ones (dense configurations
    // with > 256 points per frame) take the pair-sweep version
    const int bits_cap = big_bits_cap(ncap), bits_w = (bits_cap + 31) >> 5;
    unsigned* adj = reinterpret_cast<unsigned*>(scan + 64);
    unsigned* cm = adj + (size_t)bits_cap * bits_w;
    unsigned* bws = cm + bits_w;                             // kBigWsWords rows of bits_w words
    float* rawc = reinterpret_cast<float*>(bws + kBigWsWords * bits_w);     // bits_cap raw rows
    double* w6 = reinterpret_cast<double*>(smem + big_w6_offset(ncap));   // bits_cap world 6-vectors
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();                          // step_kernel has completed: work list, scene and track records are final
    pdl_launch_dependents();
    const int n_defer = *a.defer_count;
    // debug accounting (mmw_dbscan_big_clocks): cycles of thread 0 per part, summed over the deferred scenes
    unsigned long long* dbg = a.phase_cycles != nullptr ? a.phase_cycles + 16 + 3 * a.n_scenes : nullptr;
    long long dbg_t = clock64();
    auto stamp = [&](int k) {
        if (dbg != nullptr && tid == 0) {
            const long long now = clock64();
            atomicAdd(&dbg[k], (unsigned long long)(now - dbg_t));
            dbg_t = now;
        }
    };
    for (int it = blockIdx.x; it < n_defer; it += gridDim.x) {
        const int s = a.defer_list[it];
        SceneRec sc = a.scenes[s];
        const int T1 = sc.n_tracks;
        int B = 0;
        int fcnt[kRing], fphys[kRing];
        for (int f = 0; f < kRing; ++f) {
            fphys[f] = ring_wrap(sc.ring_head + f, c.ring_size);
            fcnt[f] = f < sc.ring_n ? sc.ring_cnt[fphys[f]] : 0;
            B += fcnt[f];
        }
        const bool bits = B <= bits_cap;
        NbScreened nb = load_fused_ring(a, s, fcnt, fphys, Xf, 3 * ncap, bits ? rawc : nullptr);
        stamp(0);
        int ncl = bits ? dbscan_bits_block(nb, B, c.db_min_samples, adj, cm, bws, par, cl, dbg != nullptr ? dbg - 10 : nullptr)
                       : dbscan_block(nb, B, c.db_min_samples, par, cl, scan, dbg != nullptr ? dbg - 10 : nullptr, false,
                                      true);
        if (dbg != nullptr && tid == 0) dbg_t = clock64();
        if (a.labels_out != nullptr)
            for (int b = tid; b < B; b += kBigThreads) a.labels_out[(size_t)s * 3 * ncap + b] = cl[b];
        if (ncl > 0) {
            if (T1 + ncl > tcap) { ncl = tcap - T1; sc.flags |= MMW_SCENE_TRACK_OVERFLOW; }
            int newslot[kMaxTcap];
            {
                const unsigned capmask = tcap >= 32 ? 0xffffffffu : ((1u << tcap) - 1u);
                unsigned fs = ~sc.slot_mask & capmask;
                for (int q = 0; q < ncl; ++q) {
                    newslot[q] = __ffs(fs) - 1; fs &= fs - 1;
                    sc.slot_mask |= 1u << newslot[q];
                }
            }
            if (bits) {                  // world coordinates and velocities of the clustered points, all threads
                for (int b = tid; b < B; b += kBigThreads)
                    if (cl[b] >= 0) {
                        const float* r = rawc + (size_t)b * kRawCols;
                        double w[6];
                        world_from_raw(c, r[0], r[1], r[2], r[3], w);
#pragma unroll
                        for (int k = 0; k < 6; ++k) w6[(size_t)b * 6 + k] = w[k];
                    }
                __syncthreads();
            }
            for (int q = warp; q < ncl; q += kBigThreads / 32)
                spawn_track(a, s, a.tracks[(size_t)s * tcap + T1 + q], q, newslot[q], sc.next_id + q, cl, fcnt, fphys,
                            lane, bits ? rawc : nullptr, bits ? w6 : nullptr);
            sc.next_id += ncl;
            sc.n_tracks = T1 + ncl;
            sc.ring_n = 0;               // batch.clear() (Tracking.py:699-700, Q8)
            sc.ring_head = 0;
            sc.ring_cnt[0] = sc.ring_cnt[1] = sc.ring_cnt[2] = 0;
            if (tid == 0) atomicAdd(&a.counters[5], (unsigned long long)ncl);
        }
        __syncthreads();
        if (tid == 0) { a.scenes[s] = sc; a.pose_cnt[s] = sc.n_tracks; }
        __syncthreads();
        stamp(1);
        if (dbg != nullptr && tid == 0) { atomicAdd(&dbg[2], 1ull); atomicAdd(&dbg[7], (unsigned long long)B); }
    }
    // every CTA read the work-list length when it started; the last one to finish empties the list for the next step
    // (no memset node between the steps)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.defer_done, 1) == (int)gridDim.x - 1) {
            *a.defer_count = 0;
            *a.defer_done = 0;
        }
    }
}

cudaError_t launch_step(const StepArgs& a, cudaStream_t stream) {
    const int smem = step_smem_bytes(a.cfg.ncap, a.cfg.tcap);
    static int configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(step_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    step_kernel<<<a.n_scenes, kStepThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_dbscan_big(const StepArgs& a, cudaStream_t stream) {
    if (a.defer_count == nullptr) return cudaSuccess;
    cudaError_t e;
    const int big_smem = big_w6_offset(a.cfg.ncap) + big_bits_cap(a.cfg.ncap) * 6 * 8;
    static int big_configured = 0;
    if (big_smem > big_configured) {
        e = cudaFuncSetAttribute(dbscan_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(dbscan_big_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        big_configured = big_smem;
    }
    return launch_pdl(dbscan_big_kernel, dim3(16), dim3(kBigThreads), big_smem, stream, dim3(1, 1, 1), a);
}

}  // namespace mmw
