// The fused per-frame tracker step: one CTA of 128 threads per scene runs
//   Utils.normalize_data (Utils.py:342-434) -> TrackBuffer.track (Tracking.py:664-703)
// i.e. transform + bounds filter + ordered compaction, Kalman predict, Mahalanobis gating and association, per-track
// cluster statistics and ring push, track maintenance, Kalman update, push of the unassigned points into the scene's
// global ring and the grid screen that decides whether DBSCAN has anything to find -- with every intermediate in
// shared memory / registers.  HBM traffic per scene-frame is the raw points in, the track records in and out, the
// ring rows written and the fused ring read by the screen.
//
// How the work is laid out (round 2; profiles/r02_*.md):
//   * inputs arrive by two bulk copies (cp.async.bulk, 1-D TMA) behind one mbarrier: the scene's contiguous block of
//     raw points (as the 16-byte aligned window around it) and its track records;
//   * point phases are thread-per-point, two points per thread per tile of 256 -- transform, filter, gate, list
//     positions, ring pushes all work from registers; only the world 6-vectors of associated points go to shared
//     memory, in list order, for the statistics;
//   * track phases are ELEMENT-parallel (linalg.cuh): thread e owns element e of the 9x9 / 9x6 / 6x6 matrices and
//     the CTA walks the tracks, so all four warps work whatever the number of tracks; the two 6x6 inverses per
//     track run two tracks per warp (one half warp each);
//   * per-track statistics: 39 values (6 sums, 21 products, 6 minima, 6 maxima) x 3 slices of the track's list = 117
//     threads, each a short FMA chain straight from shared memory, combined in a fixed order (deterministic);
//   * clusters that form are left to dbscan_big_kernel through a work list; this kernel only runs the grid screen.
#include "dbscan.cuh"
#include "linalg.cuh"
#include "mmw_internal.cuh"

namespace mmw {

constexpr int kTile = 2 * kStepThreads;          // points per tile: two per thread
constexpr int kChunks = kTile / 32;              // warp chunks of 32 consecutive points in a tile
constexpr int kBw = 72;                          // doubles of per-track scratch B
constexpr int kGinv = 40, kHx = kGinv + 21, kNest = 68;   // inside B from the gate matrices to the statistics (below)
constexpr int kNStat = 39;                       // 6 sums + 21 products + 6 minima + 6 maxima, at B[0..38]
static_assert(kStepThreads == 128, "the element-parallel track phases are written for 128 threads per scene");

// Dynamic shared memory of step_kernel.
//   tracks  TrackRec [tcap]
//   B       double [tcap][72]   gate:   C (0..35) -> statistics totals (0..38) | packed C^-1 (40..60) | H x (61..66)
//                                       | log|det C| (67)
//                               update: S -> S^-1 (0..35) | Rc (36..71)
//   A       double [tcap][81]   F P (predict), (I - K H) P (update)
//   KK      double [tcap][108]  K (0..53) | K Rc (54..107)
//   aliases: the list of world 6-vectors of a tile's associated points (kTile x 6 doubles) lies over A|KK between the
//   gate matrices and the update; the staged raw points lie over KK when a frame is a single tile, in their own
//   space otherwise (they must survive the tile loop); the DBSCAN screen's scratch lies over A|KK at the end.
struct SmemLayout {
    int tracks, B, A, KK, stage, misc, total;
};
constexpr int kMiscInts = 384;
// misc int slots (slot 0-1: the mbarrier)
enum { kT1 = 2, kFreed = 3, kScan = 4 /* 2 ints: dbscan_grid_may_have_core */, kCntK = 8 /* 8 */, kCntU = 16 /* 8 */,
       kBaseG = 24 /* 2 x 32 */, kBaseU = 88 /* 2 */, kOrder = 96 /* 32 */, kCntG = 128 /* 8 x 32 */ };

__host__ __device__ inline SmemLayout make_layout(int ncap, int tcap) {
    SmemLayout L;
    int o = 0;
    L.tracks = o;  o += tcap * (int)sizeof(TrackRec);
    L.B = o;       o += tcap * kBw * 8;
    L.A = o;
    const int a_bytes = (tcap * 81 * 8 + 15) & ~15;      // KK (bulk-copy destination when it stages the points) stays 16-byte aligned
    int kk_bytes = tcap * 108 * 8;
    const int stage_bytes = ((ncap * kRawCols * 4 + 15) & ~15) + 32;
    const bool single = ncap <= kTile;
    if (single && kk_bytes < stage_bytes) kk_bytes = stage_bytes;
    int akk = a_bytes + kk_bytes;
    const int list_bytes = kTile * 6 * 8;
    const int screen_bytes = 2 * 3 * ncap * 4 + kGridCells * 4;   // fp32 world x, y' of the fused ring + histogram
    if (akk < list_bytes) akk = list_bytes;
    if (akk < screen_bytes) akk = screen_bytes;
    akk = (akk + 15) & ~15;
    L.KK = L.A + a_bytes;
    o += akk;
    L.stage = single ? L.KK : o;
    if (!single) o += stage_bytes;
    L.misc = o;    o += kMiscInts * 4;
    L.total = o;
    return L;
}

int step_smem_bytes(int ncap, int tcap) { return make_layout(ncap, tcap).total; }

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

// ---- mbarrier + bulk copy (1-D TMA) ------------------------------------------------------------------
__device__ __forceinline__ uint32_t sk_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sk_smem(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sk_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SK_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SK_DONE;\n\t"
        "bra SK_WAIT_LOOP;\n\t"
        "SK_DONE:\n\t"
        "}" ::"r"(sk_smem(bar)), "r"(parity) : "memory");
}
// bytes: a positive multiple of 16; dst (shared) and src (global) 16-byte aligned
__device__ __forceinline__ void sk_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"(sk_smem(dst)), "l"(src), "r"(bytes), "r"(sk_smem(bar))
                 : "memory");
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
    TrackRec* tr = reinterpret_cast<TrackRec*>(smem + L.tracks);
    double* Bm = reinterpret_cast<double*>(smem + L.B);
    double* Am = reinterpret_cast<double*>(smem + L.A);
    double* KKm = reinterpret_cast<double*>(smem + L.KK);
    double* wl = Am;                                            // list of world 6-vectors (aliases A | KK)
    int* misc = reinterpret_cast<int*>(smem + L.misc);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + L.misc);

    const int s = blockIdx.x;
    if (s >= a.n_scenes) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Elem el(tid);
    const unsigned ltmask = (1u << lane) - 1u;

    // ---- 0. round trip one: scene record, offsets, dt; then the bulk copies ----------------------------------
    if (tid == 0) sk_mbar_init(mbar, 1);
    SceneRec sc = a.scenes[s];           // every thread keeps a copy; thread 0 writes it back
    const int off = a.offsets[s];
    int N = a.offsets[s + 1] - off;
    const double dt = a.dt[s];
    if (N < 0) N = 0;
    if (N > ncap) { N = ncap; sc.flags |= MMW_SCENE_POINT_OVERFLOW; }
    const int T0 = sc.n_tracks;
    __syncthreads();                     // the barrier is initialised
    const size_t byte0 = (size_t)off * (kRawCols * 4);
    const size_t win0 = byte0 & ~(size_t)15;
    const uint32_t win_bytes = (uint32_t)(((byte0 + (size_t)N * (kRawCols * 4) + 15) & ~(size_t)15) - win0);
    const bool bulk_pts = (reinterpret_cast<uintptr_t>(a.pts) & 15) == 0;       // else: plain loads (below)
    const float* stagef = reinterpret_cast<const float*>(smem + L.stage + (byte0 - win0));
    if (tid == 0) {
        const uint32_t tr_bytes = (uint32_t)T0 * (uint32_t)sizeof(TrackRec);
        sk_mbar_expect_tx(mbar, tr_bytes + (bulk_pts ? win_bytes : 0u));
        if (tr_bytes) sk_bulk_g2s(tr, a.tracks + (size_t)s * tcap, tr_bytes, mbar);
        if (bulk_pts && win_bytes)
            sk_bulk_g2s(smem + L.stage, reinterpret_cast<const unsigned char*>(a.pts) + win0, win_bytes, mbar);
    }
    if (!bulk_pts) {
        float* st = const_cast<float*>(stagef);
        for (int i = tid; i < N * kRawCols; i += kStepThreads) st[i] = a.pts[(size_t)off * kRawCols + i];
    }
    // the ring frames the screen will read: L2 prefetch, their DRAM latency overlaps everything below
    for (int f = 0; f < sc.ring_n; ++f) {
        const int phys = ring_wrap(sc.ring_head + f, c.ring_size);
        const char* fr = reinterpret_cast<const char*>(uring_frame(a, s, phys));
        const int lines = (sc.ring_cnt[phys] * kRawCols * 4 + 127) / 128;
        for (int i = tid; i < lines; i += kStepThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(fr + i * 128));
    }
    if (tid < 2 * 32 + 2) misc[kBaseG + tid < kBaseG + 64 ? kBaseG + tid : kBaseU + (tid - 64)] = 0;
    sk_mbar_wait(mbar, 0);
    if (!bulk_pts) __syncthreads();
    PHASE_MARK(1);

    // ---- 1. predict (Tracking.py:591-596, Q9) and gate matrices (Tracking.py:545-551) -----------------------
    // pass 1: A = F P
    for (int j = 0; j < T0; ++j) kf_predict_pass1(tr[j].P, Am + j * 81, tr[j].lifetime + dt, el);
    __syncthreads();
    // pass 2: P = A F' + Q; x = F x; H x for the gate
    for (int j = 0; j < T0; ++j) {
        TrackRec& t = tr[j];
        const double dte = t.lifetime + dt;
        const double x_new = tid < 9 ? kf_predict_x(t.x, dte, tid) : 0.0;
        __syncwarp();                    // warp 0: every lane has read the old x
        kf_predict_pass2(t.x, x_new, Am + j * 81, t.P, dte, c.q_var, el);
        if (tid < 6) Bm[j * kBw + kHx + tid] = x_new;
    }
    __syncthreads();
    PHASE_MARK(2);
    // gate matrix C = P[:6,:6] + Rm + G (get_Rm 361-370)
    for (int j = 0; j < T0; ++j) {
        if (tid < 36) {
            const TrackRec& t = tr[j];
            double v = t.P[el.i6 * 9 + el.a6];
            if (el.i6 == el.a6) { const double h = t.spread[el.i6] / 2; v += h * h; }
            Bm[j * kBw + tid] = v + t.G[tid];
        }
    }
    __syncthreads();
    // inverse + log|det| of the gate matrices, two tracks per warp; C^-1 is stored as its upper triangle with the
    // off-diagonal entries doubled (the quadratic form needs 21 products instead of 36).  The half warp then resets
    // the track's statistics totals, which take the place of C.
    for (int j0 = 0; j0 < T0; j0 += 2 * kStepWarps) {
        const int j = j0 + 2 * warp + (lane >> 4);
        const bool live = j < T0;
        if (__ballot_sync(kFull, live) == 0u) continue;
        double r[6];
        const double det = inv6_spd_half(live ? Bm + j * kBw : nullptr, lane, r);
        __syncwarp();
        if (live) {
            double* bj = Bm + j * kBw;
            gate_pack_half(bj + kGinv, r, det, lane);
            for (int v = lane & 15; v < kNStat; v += 16) bj[v] = v < 27 ? 0.0 : (v < 33 ? INFINITY : -INFINITY);
        }
    }
    __syncthreads();
    PHASE_MARK(3);

    // ---- 2. tiles of 256 points: transform + filter + compaction (Utils.py:379-432), gating + association
    //         (Tracking.py:553-572, Q14), list positions, ring pushes (Tracking.py:691, 338), statistics ------------
    const int uphys = sc.ring_n >= c.ring_size ? sc.ring_head : ring_wrap(sc.ring_head + sc.ring_n, c.ring_size);
    float* udst = const_cast<float*>(uring_frame(a, s, uphys));
    // role of this thread in the statistics: value v (0..5 sums of w -- the x column is a sum of lattice values and
    // therefore exact in any order, which keeps centroid[0] bit-identical to numpy's mean: the pose features sort
    // on x - centroid[0] against zero pads (Utils.py:505-514) --, 6..26 products d_r d_c (r <= c) of d = w - H x,
    // 27..32 minima of w, 33..38 maxima of w), slice q of the list
    const int sv = warp * 10 + lane / 3, sq = lane % 3;
    const bool s_active = lane < 30 && sv < kNStat;
    int s_ra = 0, s_rb = 6;
    if (sv < 6) {
        s_ra = sv;
    } else if (sv < 27) {
        int rem = sv - 6, r = 0;
        while (rem >= 6 - r) { rem -= 6 - r; ++r; }
        s_ra = r; s_rb = r + rem;
    } else {
        s_ra = (sv - 27) % 6;
    }
    int M = 0;
    const int ntiles = (N + kTile - 1) / kTile;
    for (int tile = 0; tile < ntiles; ++tile) {
        float raw[2][kRawCols];
        double w[2][6];
        unsigned km[2];
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
            const int i = tile * kTile + r2 * kStepThreads + tid;
            bool keep = false;
            if (i < N) {
#pragma unroll
                for (int k = 0; k < kRawCols; ++k) raw[r2][k] = stagef[i * kRawCols + k];
                world_from_raw(c, raw[r2][0], raw[r2][1], raw[r2][2], raw[r2][3], w[r2]);
                keep = (w[r2][2] <= c.z_max) && (w[r2][2] > 0.0) && (w[r2][1] > 0.0);      // Utils.py:423-427
            }
            km[r2] = __ballot_sync(kFull, keep);
            if (lane == 0) misc[kCntK + r2 * kStepWarps + warp] = __popc(km[r2]);
        }
        __syncthreads();
        PHASE_MARK(4);
        int pos[2], tile_kept = 0;
        {
            int pre0 = 0, pre1 = 0;
#pragma unroll
            for (int c2 = 0; c2 < kChunks; ++c2) {
                const int v = misc[kCntK + c2];
                tile_kept += v;
                if (c2 < warp) pre0 += v;
                if (c2 < kStepWarps + warp) pre1 += v;
            }
            pos[0] = M + pre0 + __popc(km[0] & ltmask);
            pos[1] = M + pre1 + __popc(km[1] & ltmask);
        }
        // gate: d2 = log|det C| + y' C^-1 y < gate, arg-min over the gated tracks, ties -> lower index
        int grp[2];
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
            grp[r2] = 255;
            if ((km[r2] >> lane) & 1u) {
                double best = INFINITY;
                int bj = -1;
                for (int j = 0; j < T0; ++j) {
                    const double d2 = gate_score(Bm + j * kBw + kGinv, w[r2]);
                    if (d2 < c.gate && d2 < best) { best = d2; bj = j; }
                }
                grp[r2] = bj < 0 ? T0 : bj;
                a.assoc_out[off + pos[r2]] = bj;
            }
        }
        // per chunk and group: count (lane j keeps track j's) and this point's rank inside its group
        int rk[2] = {0, 0};
        {
            int cj[2] = {0, 0};
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
                for (int j = 0; j < T0; ++j) {
                    const unsigned b = __ballot_sync(kFull, grp[r2] == j);
                    if (grp[r2] == j) rk[r2] = __popc(b & ltmask);
                    if (lane == j) cj[r2] = __popc(b);
                }
                const unsigned bu = __ballot_sync(kFull, grp[r2] == T0);
                if (grp[r2] == T0) rk[r2] = __popc(bu & ltmask);
                if (lane < T0) misc[kCntG + (r2 * kStepWarps + warp) * 32 + lane] = cj[r2];
                if (lane == 0) misc[kCntU + r2 * kStepWarps + warp] = __popc(bu);
            }
        }
        __syncthreads();
        PHASE_MARK(5);
        // lane j: points of track j in this tile, where this warp's two chunks start inside its list, the number of
        // points it had before this tile; the tile-local list of track j starts at the exclusive scan over j
        int tot = 0, preg[2] = {0, 0}, base = 0, start;
        {
            if (lane < T0) {
#pragma unroll
                for (int c2 = 0; c2 < kChunks; ++c2) {
                    const int v = misc[kCntG + c2 * 32 + lane];
                    tot += v;
                    if (c2 < warp) preg[0] += v;
                    if (c2 < kStepWarps + warp) preg[1] += v;
                }
                base = misc[kBaseG + (tile & 1) * 32 + lane];
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += up;
            }
            start = incl - tot;
        }
        int ubase = misc[kBaseU + (tile & 1)], utot = 0, preu[2] = {0, 0};
#pragma unroll
        for (int c2 = 0; c2 < kChunks; ++c2) {
            const int v = misc[kCntU + c2];
            utot += v;
            if (c2 < warp) preu[0] += v;
            if (c2 < kStepWarps + warp) preu[1] += v;
        }
        if (warp == 0) {                 // running counts for the next tile (double-buffered: others still read these)
            if (lane < T0) misc[kBaseG + ((tile + 1) & 1) * 32 + lane] = base + tot;
            if (lane == 0) misc[kBaseU + ((tile + 1) & 1)] = ubase + utot;
        }
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
            const int g = grp[r2];
            const int gs = g < T0 ? g : 0;
            const int gb = __shfl_sync(kFull, base, gs), gp0 = __shfl_sync(kFull, preg[r2], gs),
                      gst = __shfl_sync(kFull, start, gs);
            if (g < T0) {
                const int lp = gb + gp0 + rk[r2];                // position in the track's cloud of this frame
                const int tl = gst + gp0 + rk[r2];               // position in the tile's list
#pragma unroll
                for (int k = 0; k < 6; ++k) wl[tl * 6 + k] = w[r2][k];
                if (lp < kFeatPts) {
                    // first 64 associated rows in input order are what format_single_frame can see
                    const TrackRec& t = tr[g];
                    const int phys = t.ring_n >= c.ring_size ? t.ring_head : ring_wrap(t.ring_head + t.ring_n, c.ring_size);
                    float* dst = a.track_ring + (((size_t)s * tcap + t.slot) * kRing + phys) * (kFeatPts * kRawCols) +
                                 (size_t)lp * kRawCols;
#pragma unroll
                    for (int k = 0; k < kRawCols; ++k) dst[k] = raw[r2][k];
                }
            } else if (g == T0) {
                // BatchedData.add_frame(unassigned): the oldest frame is dropped when the ring is full
                float* dst = udst + (size_t)(ubase + preu[r2] + rk[r2]) * kRawCols;
#pragma unroll
                for (int k = 0; k < kRawCols; ++k) dst[k] = raw[r2][k];
            }
        }
        M += tile_kept;
        __syncthreads();
        PHASE_MARK(6);
        // statistics of this tile's lists (PointCluster, Tracking.py:120-136; _get_D 270-290 about H x instead of the
        // centroid: E[d d'] - m m' with d = w - H x, m = centroid - H x below the gate radius, so nothing cancels)
        for (int j = 0; j < T0; ++j) {
            const int nj = __shfl_sync(kFull, tot, j), st = __shfl_sync(kFull, start, j);
            if (nj == 0) continue;
            const double* hxj = Bm + j * kBw + kHx;
            const bool is_mom = sv < 27, is_max = sv >= 33;
            double acc = is_mom ? 0.0 : (is_max ? -INFINITY : INFINITY);
            if (s_active) {
                if (is_mom) {
                    const double ha = s_rb < 6 ? hxj[s_ra] : 0.0, hb = s_rb < 6 ? hxj[s_rb] : 0.0;
                    const int rb = s_rb < 6 ? s_rb : 0;
                    for (int k = sq; k < nj; k += 3) {
                        const double* wp = wl + (st + k) * 6;
                        const double pa = wp[s_ra] - ha;
                        const double pb = s_rb < 6 ? wp[rb] - hb : 1.0;
                        acc += pa * pb;
                    }
                } else {
                    for (int k = sq; k < nj; k += 3) {
                        const double x = wl[(st + k) * 6 + s_ra];
                        acc = is_max ? (x > acc ? x : acc) : (x < acc ? x : acc);
                    }
                }
            }
            // the three slices of a value sit in adjacent lanes; fixed order of combination
            const double a1 = __shfl_down_sync(kFull, acc, 1), a2 = __shfl_down_sync(kFull, acc, 2);
            if (is_mom) {
                acc = (acc + a1) + a2;
            } else if (is_max) {
                acc = a1 > acc ? a1 : acc; acc = a2 > acc ? a2 : acc;
            } else {
                acc = a1 < acc ? a1 : acc; acc = a2 < acc ? a2 : acc;
            }
            if (s_active && sq == 0) {
                double* tv = Bm + j * kBw + sv;
                const double old = *tv;
                *tv = sv < 27 ? old + acc : (sv < 33 ? (acc < old ? acc : old) : (acc > old ? acc : old));
            }
        }
        __syncthreads();
    }
    for (int i = M + tid; i < N; i += kStepThreads) a.assoc_out[off + i] = -2;

    sc.last_M = M;
    sc.dbscan_n = -1;
    if (M == 0) {                        // offline_main.py:55: the frame is skipped entirely (Q23)
        if (tid == 0) {
            sc.last_ran = 0;
            a.scenes[s] = sc;
            a.pose_cnt[s] = 0;
            int4* st4 = reinterpret_cast<int4*>(a.scene_stats + (size_t)s * 8);
            st4[0] = make_int4(N, 0, 0, 0);
            st4[1] = make_int4(0, 0, 0, 0);
        }
        return;
    }
    sc.last_ran = 1;
    const int U = misc[kBaseU + (ntiles & 1)];
    PHASE_MARK(7);

    // ---- 3. per-track association results (Tracking.py:648-653, 314-341) ------------------------------------
    int ring_rows = 0;
    for (int j = 0; j < T0; ++j) {
        TrackRec& t = tr[j];
        const int n = misc[kBaseG + (ntiles & 1) * 32 + j];
        ring_rows += n < kFeatPts ? n : kFeatPts;
        if (n == 0) {
            if (tid == 0) t.lifetime += dt;                      // update_lifetime(dt) (400-407)
            if (tid == 96) Bm[j * kBw + kNest] = t.n_est;
            continue;
        }
        const double* tv = Bm + j * kBw;
        const double dn = (double)n;
        double n_est = t.n_est;
        if (c.enable_est) {                                      // _estimate_point_num (232-244)
            if (dn > n_est) n_est = dn;
            else n_est = (1 - c.a_n) * n_est + c.a_n * dn;
        } else {
            n_est = (double)(n > c.est_pointnum ? n : c.est_pointnum);
        }
        if (tid < 36) {                                          // _estimate_group_disp_matrix (292-297)
            int r = el.i6, q = el.a6;
            if (r > q) { const int tmp = r; r = q; q = tmp; }
            const int p = r * 6 - (r * (r - 1)) / 2 + (q - r);
            const double mr = tv[r] / dn - tv[kHx + r], mq = tv[q] / dn - tv[kHx + q];
            const double dv = tv[6 + p] / dn - mr * mq;
            const double al = dn / n_est;
            t.G[tid] = (1 - al) * t.G[tid] + al * dv;
        } else if (tid >= 64 && tid < 70) {                      // centroid, extrema, _estimate_measurement_spread (246-268)
            const int m = tid - 64;
            const double mnv = tv[27 + m], mxv = tv[33 + m];
            t.centroid[m] = tv[m] / dn;
            t.minv[m] = mnv;
            t.maxv[m] = mxv;
            double spread = mxv - mnv;
            if (n != 1) spread = spread * (double)(n + 1) / (double)(n - 1);
            spread = fmin(2 * c.spread_lim[m], spread);
            spread = fmax(c.spread_lim[m], spread);
            if (spread > t.spread[m]) t.spread[m] = spread;
            else t.spread[m] = (1.0 - c.a_spr) * t.spread[m] + c.a_spr * spread;
        } else if (tid == 96) {
            const int phys = t.ring_n >= c.ring_size ? t.ring_head : ring_wrap(t.ring_head + t.ring_n, c.ring_size);
            if (t.ring_n >= c.ring_size) t.ring_head = ring_wrap(t.ring_head + 1, c.ring_size);
            else t.ring_n += 1;
            t.ring_cnt[phys] = n < kFeatPts ? n : kFeatPts;
            t.lifetime = 0.0;
            t.point_num = n;
            const double c3 = tv[3] / dn, c4 = tv[4] / dn, c5 = tv[5] / dn;
            t.is_static = sqrt(c3 * c3 + c4 * c4 + c5 * c5) < c.vel_thres ? 1 : 0;
            Bm[j * kBw + kNest] = n_est;             // written to the record by the maintenance warp: the
        }                                                        // threads above still read the old value
    }
    {
        if (!sc.ring_ph_gone && sc.ring_n + 1 >= c.ring_size) sc.ring_ph_gone = 1;   // the deque's initial empty frame
        if (sc.ring_n >= c.ring_size) sc.ring_head = ring_wrap(sc.ring_head + 1, c.ring_size);
        else sc.ring_n += 1;
        sc.ring_cnt[uphys] = U;
    }
    __syncthreads();
    PHASE_MARK(8);

    // ---- 4. maintenance (Tracking.py:513-528, Q16): drop timed-out tracks, keep list order -------------------
    if (warp == 0) {
        bool keep = false;
        unsigned slotbit = 0u;
        if (lane < T0) {
            const double lim = tr[lane].is_static ? c.life_sta : c.life_dyn;
            keep = !(tr[lane].lifetime > lim);                   // dropped tracks are INACTIVE: off the list
            tr[lane].n_est = Bm[lane * kBw + kNest];
            if (!keep) slotbit = 1u << tr[lane].slot;
        }
        const unsigned kmask = __ballot_sync(kFull, keep);
        const unsigned freed = __reduce_or_sync(kFull, slotbit);
        if (keep) misc[kOrder + __popc(kmask & ltmask)] = lane;
        if (lane == 0) { misc[kT1] = __popc(kmask); misc[kFreed] = (int)freed; }
    }
    __syncthreads();
    const int T1 = misc[kT1];
    sc.slot_mask &= ~(unsigned)misc[kFreed];
    PHASE_MARK(9);

    // ---- 5. update every surviving track (Tracking.py:598-603, 387-398; Q11-Q13) ----------------------------
    // _get_Rc (299-312): Rm/N + ((N_est - N)/((N_est - 1) N)) * group_disp_est;  S = P[:6,:6] + Rc
    for (int k = 0; k < T1; ++k) {
        if (tid < 36) {
            const int j = misc[kOrder + k];
            const TrackRec& t = tr[j];
            const double Nn = (double)t.point_num;
            const double coef = div_zero_fast(t.n_est - Nn, (t.n_est - 1.0) * Nn);     // N_est == N: zero numerator
            double rm = 0.0;
            if (el.i6 == el.a6) { const double h = t.spread[el.i6] / 2; rm = (h * h) / Nn; }
            const double rc = rm + coef * t.G[tid];
            double* b = Bm + j * kBw;
            b[36 + tid] = rc;
            b[tid] = t.P[el.i6 * 9 + el.a6] + rc;
        }
    }
    __syncthreads();
    for (int k0 = 0; k0 < T1; k0 += 2 * kStepWarps) {
        const int k = k0 + 2 * warp + (lane >> 4);
        const bool live = k < T1;
        if (__ballot_sync(kFull, live) == 0u) continue;
        const int j = live ? misc[kOrder + k] : 0;
        double r[6];
        (void)inv6_spd_half(live ? Bm + j * kBw : nullptr, lane, r);
        __syncwarp();
        const int cl16 = lane & 15;
        if (live && cl16 >= 6 && cl16 < 12) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Bm[j * kBw + i * 6 + (cl16 - 6)] = r[i];
        }
    }
    __syncthreads();
    for (int k = 0; k < T1; ++k) {
        const int j = misc[kOrder + k];
        kf_update_K(tr[j].P, Bm + j * kBw, KKm + j * 108, el);
    }
    __syncthreads();
    for (int k = 0; k < T1; ++k) {
        const int j = misc[kOrder + k];
        TrackRec& t = tr[j];
        kf_update_M1_KR(t.P, KKm + j * 108, Bm + j * kBw + 36, Am + j * 81, KKm + j * 108 + 54, el);
        if (warp == 3) {
            double xv = 0.0;
            if (lane < 9) xv = kf_update_x(t.x, t.centroid, KKm + j * 108, lane);
            __syncwarp();
            if (lane < 9) t.x[lane] = xv;
        }
    }
    __syncthreads();
    for (int k = 0; k < T1; ++k) {
        const int j = misc[kOrder + k];
        TrackRec& t = tr[j];
        kf_update_P(Am + j * 81, KKm + j * 108, KKm + j * 108 + 54, t.P, el);
        if (tid == 96) kf_update_nudge(t.x, t.centroid, t.lifetime == 0.0, c.nudge_thres, c.nudge_gain);
    }
    __syncthreads();
    PHASE_MARK(10);

    // ---- 6. DBSCAN over the fused global ring (Tracking.py:693-700): the grid screen --------------------------
    int Bf = 0;
    int fcnt[kRing], fphys[kRing];
    for (int f = 0; f < kRing; ++f) {
        fphys[f] = ring_wrap(sc.ring_head + f, c.ring_size);
        fcnt[f] = f < sc.ring_n ? sc.ring_cnt[fphys[f]] : 0;
        Bf += fcnt[f];
    }
    const bool run_db = Bf > 0 && T1 < c.tr_max_tracks;
    if (run_db) {
        // The residue of a steady scene cannot hold a core point and is all noise (dbscan.cuh): 99.9 % of the
        // scene-frames.  Whatever the screen lets through is most likely a cluster forming -- a couple of scenes per
        // frame -- and goes to dbscan_big_kernel (512 threads per scene) through the work list.
        bool maybe = Bf >= c.db_min_samples;
        if (maybe) {
            float* Xf = reinterpret_cast<float*>(smem + L.A);
            float* Yf = Xf + 3 * ncap;
            int b0 = 0;
            for (int f = 0; f < kRing; ++f) {
                const float* src = uring_frame(a, s, fphys[f]);
                for (int i = tid; i < fcnt[f]; i += kStepThreads) {
                    double yw, zw;
                    world_yz(c, (double)src[i * kRawCols + 1], (double)src[i * kRawCols + 2], yw, zw);
                    Xf[b0 + i] = src[i * kRawCols + 0];
                    Yf[b0 + i] = (float)yw;
                }
                b0 += fcnt[f];
            }
            __syncthreads();
            maybe = dbscan_grid_may_have_core(c, Xf, Yf, Bf, c.db_min_samples, reinterpret_cast<int*>(Yf + 3 * ncap),
                                              misc + kScan);
        }
        if (maybe) {
            if (tid == 0) a.defer_list[atomicAdd(a.defer_count, 1)] = s;
        } else if (a.labels_out != nullptr) {
            for (int b = tid; b < Bf; b += kStepThreads) a.labels_out[(size_t)s * 3 * ncap + b] = -1;
        }
        sc.dbscan_n = Bf;
    }
    PHASE_MARK(11);

    pdl_launch_dependents();             // late on purpose: dbscan_big's CTAs would otherwise sit on 16 SMs for the whole step
    // ---- 7. write back: track records in list order, scene record, per-scene counters --------------------------
    {
        double* dstbase = reinterpret_cast<double*>(a.tracks + (size_t)s * tcap);
        for (int i = tid; i < T1 * kTrackWords; i += kStepThreads) {
            const int k = i / kTrackWords, wd = i - k * kTrackWords;
            dstbase[i] = reinterpret_cast<const double*>(tr + misc[kOrder + k])[wd];
        }
    }
    if (tid == 0) {
        sc.n_tracks = T1;
        a.scenes[s] = sc;
        a.pose_cnt[s] = T1;
        int4* st4 = reinterpret_cast<int4*>(a.scene_stats + (size_t)s * 8);
        st4[0] = make_int4(N, M, U, run_db ? Bf : 0);
        st4[1] = make_int4(T1, ring_rows, 1, 0);
    }
    PHASE_MARK(12);
    if (a.phase_cycles != nullptr && threadIdx.x == 0) {
        a.phase_cycles[16 + blockIdx.x] = (unsigned long long)(clock64() - kernel_t0);   // last frame's cycles of this scene
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        a.phase_cycles[16 + 2 * gridDim.x + blockIdx.x] = ns;
    }
}

// DBSCAN + spawn for the scenes whose fused cloud passed the step kernel's grid screen (a cluster is forming: a
// couple of scenes per frame): one CTA of 512 threads per scene, new tracks written straight to the scene's list in
// global memory.  The last CTA also folds the step's per-scene counters into the context's totals.
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
            sc.ring_ph_gone = 1;
            if (tid == 0) atomicAdd(&a.counters[5], (unsigned long long)ncl);
        }
        __syncthreads();
        if (tid == 0) { a.scenes[s] = sc; a.pose_cnt[s] = sc.n_tracks; }
        __syncthreads();
        stamp(1);
        if (dbg != nullptr && tid == 0) { atomicAdd(&dbg[2], 1ull); atomicAdd(&dbg[7], (unsigned long long)B); }
    }
    // this step's counters: every scene left eight ints; the grid's last CTA (idle unless 16 clusters form at once)
    // sums them -- one reduction and seven atomics per step instead of seven atomics per scene
    if (blockIdx.x == gridDim.x - 1 && a.scene_stats != nullptr) {
        unsigned long long acc[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int sidx = tid; sidx < a.n_scenes; sidx += kBigThreads) {
            const int4 s0 = reinterpret_cast<const int4*>(a.scene_stats)[2 * sidx];
            const int4 s1 = reinterpret_cast<const int4*>(a.scene_stats)[2 * sidx + 1];
            acc[0] += (unsigned)s1.z; acc[1] += (unsigned)s0.x; acc[2] += (unsigned)s0.y; acc[3] += (unsigned)s0.z;
            acc[4] += (unsigned)s0.w; acc[5] += (unsigned)s1.x; acc[6] += (unsigned)s1.y;
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            unsigned v = (unsigned)acc[k];                   // < 2^32 per thread: at most S / 512 scenes of < 2^16 each
            v = __reduce_add_sync(kFullMask, v);
            if (lane == 0 && v) atomicAdd(&a.counters[k], (unsigned long long)v);
        }
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

int dbscan_big_smem_bytes(int ncap) { return big_w6_offset(ncap) + big_bits_cap(ncap) * 6 * 8; }

cudaError_t launch_step(const StepArgs& a, cudaStream_t stream) {
    const int smem = step_smem_bytes(a.cfg.ncap, a.cfg.tcap);
    static int configured[kMaxDevices] = {0};
    cudaError_t e = ensure_smem_attr(reinterpret_cast<const void*>(step_kernel), configured, smem);
    if (e != cudaSuccess) return e;
    step_kernel<<<a.n_scenes, kStepThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_dbscan_big(const StepArgs& a, cudaStream_t stream) {
    if (a.defer_count == nullptr) return cudaSuccess;
    const int big_smem = dbscan_big_smem_bytes(a.cfg.ncap);
    static int configured[kMaxDevices] = {0};
    cudaError_t e = ensure_smem_attr(reinterpret_cast<const void*>(dbscan_big_kernel), configured, big_smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(dbscan_big_kernel, dim3(16), dim3(kBigThreads), big_smem, stream, dim3(1, 1, 1), a);
}

}  // namespace mmw
