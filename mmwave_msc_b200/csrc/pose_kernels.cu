// Pose stage: TrackBuffer.estimate_posture (Tracking.py:705-734).
//   pose_index_kernel    compacts the tracks of all scenes that ran this frame into pose rows
//   pose_feature_kernel  relative_coordinates + format_single_frame (Utils.py:437-520) per track
//   conv_kernel          Conv(5->16)+ReLU, Conv(16->32)+ReLU, BatchNorm   (train.py:35-47 / 74-85)
//   fc1 / fc2            Dense+ReLU+BatchNorm, Dense(57)                  (train.py:49-57 / 87-96)
// The CUDA-core fp32 kernels in this file are the reference-exact path; the tensor-core (tcgen05) GEMM for the
// dense contraction lives in pose_tc.cu and replaces fc1 when enabled.
#include "mmw_internal.cuh"
#include <cstdlib>

#include "pose.cuh"

namespace mmw {

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pose_index_kernel(SceneRec* scenes, int S, int* pose_total,
                                                          unsigned long long* counters) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < S; base += 1024) {
        const int s = base + tid;
        int v = 0;
        if (s < S && scenes[s].last_ran) v = scenes[s].n_tracks;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int cw = wsum[w];
            if (w < warp) before += cw;
            tot += cw;
        }
        const int r0 = running;
        if (s < S) scenes[s].pose_base = r0 + before + incl - v;
        __syncthreads();
        if (tid == 0) running = r0 + tot;
        __syncthreads();
    }
    if (tid == 0) {
        *pose_total = running;
        atomicAdd(&counters[7], (unsigned long long)running);
    }
}

// ---------------------------------------------------------------------------------------------------
// One warp per (track, ring frame): the 3 frames of a track are independent, so a scene with two tracks keeps six
// warps busy at once instead of two warps walking three frames each (round 1: 21 us of dependent round trips for
// 27 MB of traffic).  Feature rows are relative to the CURRENT centroid for every ring frame (Q21), intensity is
// normalised before padding so pads stay exactly 0, rows are sorted by x with the canonical stable (x, row index)
// order, frames absent from the ring stay zero.  Each lane owns points lane and lane + 32 of the frame: their values
// stay in registers, only the 64 sort keys go through shared memory, and both ranks come out of one pass over them.
constexpr int kFeatWarps = 6;
__global__ void __launch_bounds__(kFeatWarps * 32, 7) pose_feature_kernel(PoseFeatArgs a) {
    __shared__ __align__(16) long long keys[kFeatWarps][kFeatPts];     // order-preserving integer images of the sort keys
    __shared__ __align__(16) float srow[kFeatWarps][kFeatPts * kRawCols];
    __shared__ __align__(16) uint4 spk[kFeatWarps][kFeatPts * 2];
    const int s = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();                          // the tracker kernels have completed
    pdl_launch_dependents();
    const SceneRec sc = a.scenes[s];
    int pose_base = sc.pose_base;
    if (a.dbg & 1) pose_base = 2 * s;
    else if (a.pose_cnt != nullptr) {
        // exclusive scan value of this scene: every CTA adds up the (at most S) counts in front of it; the order of
        // the integer additions does not matter, so the row layout is the same as pose_index_kernel's
        __shared__ int wpart[kFeatWarps];
        int part = 0;
        for (int i = threadIdx.x; i < s; i += kFeatWarps * 32) part += __ldg(a.pose_cnt + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) wpart[warp] = part;
        __syncthreads();
        pose_base = 0;
#pragma unroll
        for (int w = 0; w < kFeatWarps; ++w) pose_base += wpart[w];
        if (threadIdx.x == 0) {
            if (s == a.n_scenes - 1) {
                const int total = pose_base + (sc.last_ran ? sc.n_tracks : 0);
                *a.pose_total = total;
                if (a.rows_hint != nullptr) *a.rows_hint = total;        // zero-copy write to pinned host memory
                atomicAdd(&a.counters[7], (unsigned long long)total);
            }
        }
    }
    if (!sc.last_ran) return;
    const DevConfig& c = a.cfg;
    const int nfr = c.ring_size;
    const int items = sc.n_tracks * nfr;
    for (int item = warp; item < items; item += kFeatWarps) {
        const int k = item / nfr, f = item - k * nfr;
        const TrackRec* t = a.tracks + (size_t)s * c.tcap + k;
        const int row = pose_base + k;
        const int slot = t->slot, rn = t->ring_n, rh = t->ring_head;
        if (f == 0 && lane == 0) {
            a.row_scene[row] = s;
            a.row_track[row] = k;
            a.row_slot[row] = slot;
        }
        float* out = a.feats + ((size_t)row * nfr + f) * (kFeatPts * kRawCols);
        uint4* pk = a.packed ? reinterpret_cast<uint4*>(a.packed + (size_t)row * nfr * kFeatPts * 16) : nullptr;
        if (f >= rn) {
            for (int e = lane; e < kFeatPts * kRawCols; e += 32) out[e] = 0.f;
            if (pk)
                for (int e = lane; e < kFeatPts * 2; e += 32) pk[f * kFeatPts * 2 + e] = make_uint4(0, 0, 0, 0);
            continue;
        }
        const double cx = t->centroid[0], cy = t->centroid[1];
        const int phys = ring_wrap(rh + f, c.ring_size);
        const int cnt = t->ring_cnt[phys];
        const float* src = a.track_ring + (((size_t)s * c.tcap + slot) * kRing + phys) * (kFeatPts * kRawCols);
        float* so = srow[warp];
        {   // the frame's 64 raw rows (1280 contiguous bytes) as 80 float4 loads, not 10 strided scalar loads per lane
            const float4* g4 = reinterpret_cast<const float4*>(src);
            float4* s4 = reinterpret_cast<float4*>(so);
            if (!(a.dbg & 8))
                for (int e = lane; e < kFeatPts * kRawCols / 4; e += 32) s4[e] = __ldg(g4 + e);
        }
        __syncwarp();
        float v[2][kRawCols];
        long long mkey[2];
        // Fast path of the sort (taken when every x of the frame is a multiple of 2^-16 below 128 m -- the sensor's
        // Q-format lattice -- and |cx| < 2^30): there fl(x - cx) is STRICTLY monotone in x (two lattice values differ by
        // >= 2^-16, the rounding moves a difference by < 2^-22), and a pad (key 0.0) sorts like a point at cx.  So the
        // order of (key, index) is the order of the 32-bit word  (2 x 2^16 [pads: the odd or even integer at cx] + 2^25)
        // << 6 | index, all different: one integer compare per pair instead of a 64-bit compare with a tie rule.
        int kx[2];
        bool lattice = fabs(cx) < 1073741824.0;
        {
            const double cs = cx * 65536.0, fl = floor(cs);
            const int kpad = cs >= 8388608.0 ? (1 << 24) + 1 : (cs <= -8388608.0 ? -(1 << 24) - 1
                             : 2 * (int)fl + (cs != fl ? 1 : 0));
            kx[0] = kx[1] = kpad;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            double key = 0.0;
#pragma unroll
            for (int q = 0; q < kRawCols; ++q) v[h][q] = 0.f;
            if (i < cnt) {
                {
                    const float xs = so[i * kRawCols + 0] * 65536.f;
                    lattice = lattice && fabsf(xs) < 8388608.f && xs == rintf(xs);
                    kx[h] = 2 * (int)xs;
                }
                const float x = so[i * kRawCols + 0], y = so[i * kRawCols + 1], z = so[i * kRawCols + 2];
                const float d = so[i * kRawCols + 3], p = so[i * kRawCols + 4];
                double yw, zw;
                world_yz(c, (double)y, (double)z, yw, zw);
                key = __dsub_rn((double)x, cx);
                v[h][0] = (float)key;
                v[h][1] = (float)__dsub_rn(yw, cy);
                v[h][2] = (float)zw;
                v[h][3] = (float)__dmul_rn((double)d, c.doppler_res);
                v[h][4] = (float)__ddiv_rn(__dsub_rn((double)p, c.int_mu), c.int_std);
            }
            const long long b = __double_as_longlong(__dadd_rn(key, 0.0));
            mkey[h] = b ^ ((b >> 63) & 0x7fffffffffffffffLL);
        }
        const bool fast = __all_sync(0xffffffffu, lattice) && !(a.dbg & 16);
        if (!fast) {
            keys[warp][lane] = mkey[0];
            keys[warp][lane + 32] = mkey[1];
        } else {
            unsigned* k32 = reinterpret_cast<unsigned*>(keys[warp]);
            k32[lane] = ((unsigned)(kx[0] + (1 << 25)) << 6) | (unsigned)lane;
            k32[lane + 32] = ((unsigned)(kx[1] + (1 << 25)) << 6) | (unsigned)(lane + 32);
        }
        __syncwarp();
        int rank[2] = {0, 0};
        if (fast) {
            const unsigned c0 = ((unsigned)(kx[0] + (1 << 25)) << 6) | (unsigned)lane;
            const unsigned c1 = ((unsigned)(kx[1] + (1 << 25)) << 6) | (unsigned)(lane + 32);
            const uint4* k4 = reinterpret_cast<const uint4*>(keys[warp]);
#pragma unroll
            for (int q4 = 0; q4 < kFeatPts / 4; ++q4) {
                const uint4 kq = k4[q4];                 // broadcast read
                rank[0] += (kq.x < c0) + (kq.y < c0) + (kq.z < c0) + (kq.w < c0);
                rank[1] += (kq.x < c1) + (kq.y < c1) + (kq.z < c1) + (kq.w < c1);
            }
        } else {
        // Stable rank of (key, index).  Float64 compares (DSETP) issue at a fraction of the integer rate on this part
        // and the 4096 compares per frame were two thirds of the kernel's time (profiles/r02_feat_probe.txt), so the
        // keys are compared as integers: a finite double maps to an int64 with the same order (-0.0 was folded into
        // +0.0 above), and "kq < k  or  (kq == k and q < i)" is "mq < m + (q < i)".
        const long long lt0 = mkey[0], le0 = mkey[0] + 1, lt1 = mkey[1], le1 = mkey[1] + 1;
        const longlong2* k2 = reinterpret_cast<const longlong2*>(keys[warp]);
        const int nq2 = (a.dbg & 4) ? 1 : kFeatPts / 4;
#pragma unroll 8
        for (int q2 = 0; q2 < nq2; ++q2) {               // q = 2 q2, 2 q2 + 1 < 32: below every index of the second point
            const longlong2 kq = k2[q2];                 // broadcast read
            const int q = 2 * q2;
            rank[0] += kq.x < (q < lane ? le0 : lt0) ? 1 : 0;
            rank[0] += kq.y < (q + 1 < lane ? le0 : lt0) ? 1 : 0;
            rank[1] += kq.x < le1 ? 1 : 0;
            rank[1] += kq.y < le1 ? 1 : 0;
        }
#pragma unroll 8
        for (int q2 = 0; q2 < nq2; ++q2) {               // q = 32 + 2 q2 ...: above every index of the first point
            const longlong2 kq = k2[kFeatPts / 4 + q2];
            const int q = 2 * q2;
            rank[0] += kq.x < lt0 ? 1 : 0;
            rank[0] += kq.y < lt0 ? 1 : 0;
            rank[1] += kq.x < (q < lane ? le1 : lt1) ? 1 : 0;
            rank[1] += kq.y < (q + 1 < lane ? le1 : lt1) ? 1 : 0;
        }
        }
        // the sorted frame is put together in shared memory and leaves as whole 16-byte vectors: a row is 20 bytes,
        // written from registers it cost 14 store instructions of 32 scattered sectors each (the L2 request rate, not
        // the bytes, bounded the kernel)
        uint4* sp = spk[warp];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int q = 0; q < kRawCols; ++q) so[rank[h] * kRawCols + q] = v[h][q];
            if (pk) {
                __align__(16) __nv_bfloat16 pv[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) pv[q] = __float2bfloat16_rn(0.f);
#pragma unroll
                for (int q = 0; q < kRawCols; ++q) {
                    const float x = v[h][q];
                    pv[q] = __float2bfloat16_rn(x);
                    pv[8 + q] = __float2bfloat16_rn(x - __bfloat162float(pv[q]));
                }
                // slab order of the tensor-core convs: [row][d = f][w][chunk][h][8]  (pose_tc.cu)
                const int ph = rank[h] >> 3, pw_ = rank[h] & 7;
                sp[(pw_ * 2 + 0) * 8 + ph] = reinterpret_cast<const uint4*>(pv)[0];
                sp[(pw_ * 2 + 1) * 8 + ph] = reinterpret_cast<const uint4*>(pv)[1];
            }
        }
        __syncwarp();
        if (!(a.dbg & 2)) {
            float4* o4 = reinterpret_cast<float4*>(out);
            const float4* s4 = reinterpret_cast<const float4*>(so);
            for (int e = lane; e < kFeatPts * kRawCols / 4; e += 32) o4[e] = s4[e];
            if (pk) {
#pragma unroll
                for (int e = 0; e < 4; ++e) pk[f * kFeatPts * 2 + e * 32 + lane] = sp[e * 32 + lane];
            }
        }
        __syncwarp();                    // the keys are reused by this warp's next item
    }
}

// ---------------------------------------------------------------------------------------------------
// Conv stack, one CTA per pose row.  D = 1 (define_CNN, 8x8x5) or 3 (define_CNN_3D, 3x8x8x5).
// 'same' zero padding, cross-correlation, channels-last; Dropout is identity at inference.
template <int D>
__global__ void __launch_bounds__(256) conv_kernel(ConvArgs a) {
    constexpr int P = D * 64;                 // output positions
    constexpr int TAPS = (D == 3 ? 27 : 9);
    constexpr int A1S = 17;                   // padded row stride of act1 (bank conflicts)
    extern __shared__ __align__(16) float sm[];
    float* w2 = sm;                           // [TAPS][16][32]
    float* w1 = w2 + TAPS * 16 * 32;          // [TAPS][5][16]
    float* in = w1 + TAPS * 5 * 16;           // [P][5]
    float* a1 = in + P * 5;                   // [P][A1S]
    __shared__ int total;
    if (threadIdx.x == 0) total = *a.n_rows;
    for (int i = threadIdx.x; i < TAPS * 16 * 32; i += 256) w2[i] = a.w2[i];
    for (int i = threadIdx.x; i < TAPS * 5 * 16; i += 256) w1[i] = a.w1[i];
    __syncthreads();
    for (int row = blockIdx.x; row < total; row += gridDim.x) {
        const float* feat = a.feats + (size_t)row * P * 5;
        for (int i = threadIdx.x; i < P * 5; i += 256) in[i] = feat[i];
        __syncthreads();
        // conv1: thread = (pos, 8 output channels)
        for (int it = threadIdx.x; it < P * 2; it += 256) {
            const int pos = it >> 1, c0 = (it & 1) * 8;
            const int d = pos / 64, h = (pos / 8) % 8, w = pos % 8;
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = a.b1[c0 + q];
            for (int tap = 0; tap < TAPS; ++tap) {
                const int kd = (D == 3) ? tap / 9 - 1 : 0, kh = (tap / 3) % 3 - 1, kw = tap % 3 - 1;
                const int dd = d + kd, hh = h + kh, ww = w + kw;
                if (dd < 0 || dd >= D || hh < 0 || hh >= 8 || ww < 0 || ww >= 8) continue;
                const float* ip = in + ((dd * 8 + hh) * 8 + ww) * 5;
                const float* wp_ = w1 + tap * 5 * 16 + c0;
#pragma unroll
                for (int ci = 0; ci < 5; ++ci) {
                    const float v = ip[ci];
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[q] = fmaf(v, wp_[ci * 16 + q], acc[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) a1[pos * A1S + c0 + q] = fmaxf(acc[q], 0.f);
        }
        __syncthreads();
        // conv2 + ReLU + BatchNorm: thread = (pos, 8 output channels)
        float* out = a.act2 + (size_t)row * P * 32;
        for (int it = threadIdx.x; it < P * 4; it += 256) {
            const int pos = it % P, c0 = (it / P) * 8;
            const int d = pos / 64, h = (pos / 8) % 8, w = pos % 8;
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = a.b2[c0 + q];
            for (int tap = 0; tap < TAPS; ++tap) {
                const int kd = (D == 3) ? tap / 9 - 1 : 0, kh = (tap / 3) % 3 - 1, kw = tap % 3 - 1;
                const int dd = d + kd, hh = h + kh, ww = w + kw;
                if (dd < 0 || dd >= D || hh < 0 || hh >= 8 || ww < 0 || ww >= 8) continue;
                const float* ip = a1 + ((dd * 8 + hh) * 8 + ww) * A1S;
                const float* wp_ = w2 + tap * 16 * 32 + c0;
#pragma unroll
                for (int ci = 0; ci < 16; ++ci) {
                    const float v = ip[ci];
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[q] = fmaf(v, wp_[ci * 32 + q], acc[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                out[pos * 32 + c0 + q] = fmaf(fmaxf(acc[q], 0.f), a.bn1_scale[c0 + q], a.bn1_shift[c0 + q]);
        }
        __syncthreads();
    }
}

size_t conv_smem_bytes(int D) {
    const int P = D * 64, TAPS = (D == 3 ? 27 : 9);
    return sizeof(float) * (size_t)(TAPS * 16 * 32 + TAPS * 5 * 16 + P * 5 + P * 17);
}

cudaError_t launch_conv(const ConvArgs& a, int D, int grid, cudaStream_t st) {
    const size_t smem = conv_smem_bytes(D);
    cudaError_t e;
    if (D == 3) {
        e = cudaFuncSetAttribute(conv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        conv_kernel<3><<<grid, 256, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        conv_kernel<1><<<grid, 256, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Dense 1 (CUDA-core fp32 path): out[n][H] = BN2(relu(A[n][K] W[K][H] + b)).  64x64 tile, 256 threads, 4x4 each.
__global__ void __launch_bounds__(256) fc1_simt_kernel(FcArgs a) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int n = *a.n_rows;
    const int m0 = blockIdx.y * 64, h0 = blockIdx.x * 64;
    if (m0 >= n) return;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < a.K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i / 16, kk = i % 16;
            As[kk][r] = (m0 + r < n) ? a.A[(size_t)(m0 + r) * a.K + k0 + kk] : 0.f;
        }
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int kk = i / 64, cc = i % 64;
            Bs[kk][cc] = a.W[(size_t)(k0 + kk) * a.H + h0 + cc];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int h = h0 + tx * 4 + j;
            const float v = fmaxf(acc[i][j] + a.bias[h], 0.f);
            a.out[(size_t)m * a.H + h] = fmaf(v, a.bn_scale[h], a.bn_shift[h]);
        }
    }
}

cudaError_t launch_fc1_simt(const FcArgs& a, int max_rows, cudaStream_t st) {
    dim3 grid(a.H / 64, (max_rows + 63) / 64);
    fc1_simt_kernel<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// Dense 2: keypoints[n][57] = act3[n][H] W3[H][57] + b3.  One CTA = 8 rows; thread = (output o, K half); the
// weight element is loaded once and used for the 8 rows staged in shared memory.  Also scatters the result to the
// owning track's keypoint slot (Tracking.py:733-734).
constexpr int kFc2Rows = 8;
__global__ void __launch_bounds__(128) fc2_kernel(Fc2Args a) {
    extern __shared__ float xs[];                 // [kFc2Rows][H]
    __shared__ float part[kFc2Rows][64];
    const int n = *a.n_rows;
    const int o = threadIdx.x & 63, half = threadIdx.x >> 6;
    const int hlen = a.H / 2;
    for (int r0 = blockIdx.x * kFc2Rows; r0 < n; r0 += gridDim.x * kFc2Rows) {
        const int nr = min(kFc2Rows, n - r0);
        for (int i = threadIdx.x; i < kFc2Rows * a.H; i += 128) {
            const int r = i / a.H;
            xs[i] = r < nr ? a.act3[(size_t)(r0 + r) * a.H + (i - r * a.H)] : 0.f;
        }
        __syncthreads();
        float acc[kFc2Rows];
#pragma unroll
        for (int r = 0; r < kFc2Rows; ++r) acc[r] = 0.f;
        if (o < kKp) {
            const float* wp_ = a.W + (size_t)(half * hlen) * kKp + o;
            const float* xp = xs + half * hlen;
#pragma unroll 4
            for (int h = 0; h < hlen; ++h) {
                const float w = __ldg(wp_ + (size_t)h * kKp);
#pragma unroll
                for (int r = 0; r < kFc2Rows; ++r) acc[r] = fmaf(xp[r * a.H + h], w, acc[r]);
            }
        }
        if (half == 1) {
#pragma unroll
            for (int r = 0; r < kFc2Rows; ++r) part[r][o] = acc[r];
        }
        __syncthreads();
        if (half == 0 && o < kKp) {
#pragma unroll
            for (int r = 0; r < kFc2Rows; ++r) {
                if (r >= nr) break;
                const float v = (acc[r] + part[r][o]) + a.bias[o];
                const int row = r0 + r;
                a.out[(size_t)row * kKp + o] = v;
                if (a.keypoints != nullptr) {
                    const int s = a.row_scene[row], slot = a.row_slot[row];
                    a.keypoints[((size_t)s * a.tcap + slot) * kKp + o] = v;
                }
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_fc2(const Fc2Args& a, int grid, cudaStream_t st) {
    const size_t smem = (size_t)kFc2Rows * a.H * sizeof(float);
    static int configured[kMaxDevices] = {0};
    cudaError_t e = ensure_smem_attr(reinterpret_cast<const void*>(fc2_kernel), configured, 64 * 1024);
    if (e != cudaSuccess) return e;
    fc2_kernel<<<grid, 128, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_pose_index(SceneRec* scenes, int S, int* pose_total, unsigned long long* counters,
                              cudaStream_t st) {
    pose_index_kernel<<<1, 1024, 0, st>>>(scenes, S, pose_total, counters);
    return cudaGetLastError();
}

cudaError_t launch_pose_features(const PoseFeatArgs& a, int S, cudaStream_t st) {
    static int configured[kMaxDevices] = {0};   // same shared-memory carve-out as its neighbours in the step
    cudaError_t e = ensure_smem_attr(reinterpret_cast<const void*>(pose_feature_kernel), configured, -1);
    if (e != cudaSuccess) return e;
    const char* env = getenv("MMW_FEAT_DBG");              // read at every launch: a test switches paths in one process
    const int dbg = env ? atoi(env) : 0;
    PoseFeatArgs b = a;
    b.dbg = dbg;
    return launch_pdl(pose_feature_kernel, dim3(S), dim3(kFeatWarps * 32), 0, st, dim3(1, 1, 1), b);
}

}  // namespace mmw
