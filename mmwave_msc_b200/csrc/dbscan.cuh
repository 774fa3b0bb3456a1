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

// X, Y, Z: B doubles each (shared).  par, cl: B ints each (shared).  On return cl[b] is the label of b.
// Returns the number of clusters (uniform over the block).
__device__ inline int dbscan_block(const DevConfig& c, const double* X, const double* Y, const double* Z, int B,
                                   double eps, int min_samples, int* par, int* cl, int* s_scan) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // Neighbour counts.  The predicate is symmetric, so each unordered pair is evaluated once and the B(B-1)/2
    // pairs are split evenly over the whole block (integer shared-memory atomics: order-independent, so the
    // counts are deterministic).  Steady state is ~90 noise points = 4005 pairs = 32 per thread instead of one
    // 90-long loop per point.
    for (int b = tid; b < B; b += nt) par[b] = 1;                      // a point is its own neighbour
    __syncthreads();
    {
        const int P = B * (B - 1) / 2;
        const int chunk = (P + nt - 1) / nt;
        int p = tid * chunk;
        const int pend = min(P, p + chunk);
        if (p < pend) {
            // pair index p <-> (b, q), q < b:  p = b(b-1)/2 + q
            int b = (int)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
            while (b * (b - 1) / 2 > p) --b;
            while ((b + 1) * b / 2 <= p) ++b;
            int q = p - b * (b - 1) / 2;
            double x = X[b], y = Y[b], z = Z[b];
            for (; p < pend; ++p) {
                if (eps_neighbour(c, x, y, z, X[q], Y[q], Z[q], eps)) {
                    atomicAdd(&par[b], 1);
                    atomicAdd(&par[q], 1);
                }
                if (++q == b) {
                    ++b; q = 0;
                    if (b < B) { x = X[b]; y = Y[b]; z = Z[b]; }
                }
            }
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
    // connected components over core points
    for (int b = tid; b < B; b += nt) {
        if (((volatile int*)par)[b] < 0) continue;
        const double x = X[b], y = Y[b], z = Z[b];
        for (int q = 0; q < b; ++q) {
            if (((volatile int*)par)[q] < 0) continue;
            if (eps_neighbour(c, x, y, z, X[q], Y[q], Z[q], eps)) uf_unite(par, b, q);
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
    // border points: lowest-numbered cluster with a core point within eps
    for (int b = tid; b < B; b += nt) {
        if (cl[b] >= 0) continue;
        const double x = X[b], y = Y[b], z = Z[b];
        int best = 0x7fffffff;
        for (int q = 0; q < B; ++q) {
            const int lq = cl[q];
            if (lq >= 0 && lq < best && eps_neighbour(c, x, y, z, X[q], Y[q], Z[q], eps)) best = lq;
        }
        par[b] = best == 0x7fffffff ? -1 : best;
    }
    __syncthreads();
    // NB: cl[q] >= 0 meant "core" during the pass above; merge border labels only now.
    for (int b = tid; b < B; b += nt)
        if (cl[b] < 0) cl[b] = par[b];
    __syncthreads();
    return ncl;
}

}  // namespace mmw
