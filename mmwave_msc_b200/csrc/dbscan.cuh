// Block-cooperative exact-neighbourhood DBSCAN with sklearn's canonical labelling
// (Utils.apply_DBscan, Utils.py:250-291; sklearn DBSCAN.fit + dbscan_inner).
//
//   neighbourhood(p) = { q : altered_EuclideanDist(p, q) <= eps }, p included
//   core(p)          = |neighbourhood(p)| >= min_samples
//   clusters         = connected components of the core-core epsilon graph (union-find, smaller index is
//                      the root, so the root is the smallest core index of the component)
//   label(core p)    = rank of its root among all roots in ascending index order
//   label(border p)  = min label over core points within eps, else -1 (noise)
//
// which is exactly what dbscan_inner's ascending outer loop + DFS produces.
#pragma once
#include "mmw_internal.cuh"

namespace mmw {

__device__ __forceinline__ int uf_find(volatile int* par, int i) {
    while (true) {
        const int p = par[i];
        if (p == i) return i;
        i = p;
    }
}

__device__ __forceinline__ void uf_unite(int* par, int a, int b) {
    while (true) {
        a = uf_find(par, a);
        b = uf_find(par, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }   // hook the larger root under the smaller
        const int old = atomicMin(&par[a], b);
        if (old == a) return;
        a = old;                                          // a had already been hooked elsewhere: merge that too
    }
}

// Exclusive block scan of one flag per thread; returns this thread's rank, *total = number of set flags.
// s_scan: blockDim.x/32 + 1 ints of shared memory.  Must be called by every thread of the block.
__device__ __forceinline__ int block_rank(bool flag, int* total, int* s_scan) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_scan[warp] = __popc(m);
    __syncthreads();
    int before = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        const int c = s_scan[w];
        if (w < warp) before += c;
        tot += c;
    }
    __syncthreads();
    *total = tot;
    return before + __popc(m & ((1u << lane) - 1u));
}

// Neighbour predicates ------------------------------------------------------------------------------------
// Exact: float64 world coordinates in shared memory (stage-level entry point mmw_dbscan).
struct NbExact {
    const DevConfig& c;
    const double *X, *Y, *Z;
    double eps;
    __device__ __forceinline__ bool operator()(int b, int q) const {
        return eps_neighbour(c, X[b], Y[b], Z[b], X[q], Y[q], Z[q], eps);
    }
};

// Screened: the fused step keeps the world coordinates of the ring points rounded to fp32 in shared memory and
// evaluates the predicate in fp32 first (FP32 pipe, 4-cycle ops); only pairs whose fp32 distance falls inside a
// guard band around eps -- wider than any fp32 rounding could move it -- are re-evaluated exactly in float64 from
// the raw ring rows.  Decisions are therefore identical to the exact predicate, at a fraction of the FP64 work
// (the FP64/XU pipe is what bounds this kernel, profiles/).
struct NbScreened {
    const DevConfig& c;
    const float *Xf, *Yf, *Zf;        // shared memory, B floats each
    const float* frame[kRing];        // raw rows of the ring frames, oldest first (global)
    int start[kRing + 1];             // first fused index of each frame
    double eps;
    float lo, hi, rw, zw;
    __device__ __forceinline__ void world(int b, double& x, double& y, double& z) const {
        const int f = (b >= start[1] ? 1 : 0) + (b >= start[2] ? 1 : 0);
        const float* r = frame[f] + (size_t)(b - start[f]) * kRawCols;
        x = (double)r[0];
        world_yz(c, (double)r[1], (double)r[2], y, z);
    }
    __device__ __forceinline__ bool operator()(int b, int q) const {
        const float yb = Yf[b], yq = Yf[q];
        const float w = 1.f - 0.5f * (yb + yq) * rw;
        const float dx = Xf[b] - Xf[q], dy = yb - yq, dz = Zf[b] - Zf[q];
        const float d = w * (dx * dx + dy * dy + zw * (dz * dz));
        if (d > hi) return false;
        if (d < lo) return true;
        double x1, y1, z1, x2, y2, z2;                  // inside the guard band (or not finite): decide exactly
        world(b, x1, y1, z1);
        world(q, x2, y2, z2);
        return eps_neighbour(c, x1, y1, z1, x2, y2, z2, eps);
    }
};

// par, cl: B ints each (shared).  On return cl[b] is the label of b.
// Returns the number of clusters (uniform over the block).
// Calls fn(b, q) for every unordered pair q < b < B, the B(B-1)/2 pairs split evenly over the block.
template <class Fn>
__device__ __forceinline__ void for_pairs(int B, Fn fn) {
    const int nt = blockDim.x;
    const int P = B * (B - 1) / 2;
    const int chunk = (P + nt - 1) / nt;
    int p = threadIdx.x * chunk;
    const int pend = min(P, p + chunk);
    if (p >= pend) return;
    // pair index p <-> (b, q), q < b:  p = b(b-1)/2 + q
    int b = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
    while (b * (b - 1) / 2 > p) --b;
    while ((b + 1) * b / 2 <= p) ++b;
    int q = p - b * (b - 1) / 2;
    for (; p < pend; ++p) {
        fn(b, q);
        if (++q == b) { ++b; q = 0; }
    }
}

constexpr int kPairSweepMax = 160;    // above this many fused points the count switches to early-exit rows

template <class Nb>
__device__ inline int dbscan_block(const Nb& nb, int B, int min_samples, int* par, int* cl, int* s_scan) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // 1. neighbour counts (only "count >= min_samples" matters).
    //    Small clouds (the steady-state noise residue, ~90 points): every unordered pair once, spread evenly over
    //    the block, combined with integer shared-memory atomics (order-independent => deterministic).
    //    Large clouds (a person walked in: hundreds of points, mostly one dense blob): one row per work item with
    //    early exit at min_samples -- a blob point is decided after ~40 predicates instead of B -- rows handed out
    //    through a shared-memory queue so the few full-length noise rows do not pile up on one thread.
    if (B <= kPairSweepMax) {
        for (int b = tid; b < B; b += nt) par[b] = 1;                  // a point is its own neighbour
        __syncthreads();
        for_pairs(B, [&](int b, int q) {
            if (nb(b, q)) { atomicAdd(&par[b], 1); atomicAdd(&par[q], 1); }
        });
    } else {
        int* queue = s_scan + 6;
        if (tid == 0) *queue = 0;
        __syncthreads();
        while (true) {
            const int b = atomicAdd(queue, 1);
            if (b >= B) break;
            int cnt = 1;
            for (int q = 0; q < B && cnt < min_samples; ++q)
                if (q != b && nb(b, q)) ++cnt;
            par[b] = cnt;
        }
    }
    __syncthreads();
    int anycore = 0;
    for (int b = tid; b < B; b += nt) {
        const bool core = par[b] >= min_samples;
        par[b] = core ? b : -1;
        anycore |= core ? 1 : 0;
    }
    anycore = __syncthreads_or(anycore);
    if (!anycore) {
        for (int b = tid; b < B; b += nt) cl[b] = -1;
        __syncthreads();
        return 0;
    }
    // 2. connected components over core points (par[x] >= 0 <=> x is core, stable under the unions).  Pairs that
    //    are already in one component are skipped BEFORE the predicate is evaluated: inside a dense blob almost
    //    every pair is.
    for (int b = tid; b < B; b += nt) {
        if (((volatile int*)par)[b] < 0) continue;
        int rb = uf_find(par, b);
        for (int q = 0; q < b; ++q) {
            if (((volatile int*)par)[q] < 0) continue;
            if (uf_find(par, q) == rb) continue;
            if (nb(b, q)) { uf_unite(par, b, q); rb = uf_find(par, b); }
        }
    }
    __syncthreads();
    for (int b = tid; b < B; b += nt) cl[b] = par[b] >= 0 ? uf_find(par, b) : -1;
    __syncthreads();
    // rank of each root among the roots, ascending index
    int ncl = 0;
    for (int base = 0; base < B; base += nt) {
        const int b = base + tid;
        const bool isroot = b < B && cl[b] == b;
        int tot;
        const int r = block_rank(isroot, &tot, s_scan);
        if (isroot) par[b] = ncl + r;          // par[root] now holds the cluster id
        ncl += tot;
    }
    __syncthreads();
    for (int b = tid; b < B; b += nt) {
        const int r = cl[b];
        if (r >= 0) cl[b] = par[r];
    }
    __syncthreads();
    // 3. border points: lowest-numbered cluster with a core point within eps (cl[q] >= 0 <=> q is core here;
    //    cores of clusters that cannot improve the current best are skipped without evaluating the predicate)
    for (int b = tid; b < B; b += nt) {
        if (cl[b] >= 0) continue;
        int best = 0x7fffffff;
        for (int q = 0; q < B; ++q) {
            const int lq = cl[q];
            if (lq >= 0 && lq < best && nb(b, q)) best = lq;
        }
        par[b] = best == 0x7fffffff ? -1 : best;
    }
    __syncthreads();
    for (int b = tid; b < B; b += nt)
        if (cl[b] < 0) cl[b] = par[b];
    __syncthreads();
    return ncl;
}

}  // namespace mmw
