// Dense layer 1 of the pose network on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   out[m][n] = BN2( relu( sum_k A[m][k] W[k][n] + bias[n] ) ),   A = BN1(relu(conv2)) flattened, K = D*64*32,
//                                                                 N = D*512  (train.py:49-52 / 87-90)
//
// Precision: the network is fp32 in the reference (Keras default) and the joints must hold 1 mm.  A single bf16
// or tf32 pass does not (14 mm / 2 mm worst case), so every operand is split into two bf16
// terms, a = a_hi + a_lo, w = w_hi + w_lo, and three tensor-core products are accumulated in fp32 in TMEM:
//   a_hi*w_hi + a_hi*w_lo + a_lo*w_hi          (the dropped a_lo*w_lo term is ~2^-16 relative)
//
// Kernel shape: one CTA per 128 x 256 output tile; warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM
// allocation), warps 2-5 = epilogue (TMEM -> registers -> bias/ReLU/BatchNorm -> global).  Each pipeline stage
// holds the four operand tiles of one K block {A_hi, A_lo, W_hi, W_lo} (loaded once, used by three MMAs per
// 16-wide K step), so L2->SMEM traffic is 2/3 of what three independent GEMM passes would move.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstring>
#include <string>
#include <vector>

#include "pose_tc.cuh"

namespace mmw {

namespace tc {
constexpr int BM = 128;            // UMMA M
constexpr int BN = 256;            // UMMA N (TMEM columns, fp32)
constexpr int BK = 64;             // K block = 128 bytes of bf16 = one SWIZZLE_128B atom row
constexpr int UK = 16;             // UMMA K for 16-bit inputs
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
constexpr int B_BYTES = BN * BK * 2;                 // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 96 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
}  // namespace tc

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in shared memory, rows of 128 bytes, SWIZZLE_128B (what the TMA wrote):
// 8-row atoms are 1024 bytes apart (SBO); LBO is unused for swizzled K-major tiles; version = 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);           // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                           // leading byte offset
    d |= (uint64_t)((1024 >> 4) & 0x3fff) << 32;      // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                           // version
    d |= (uint64_t)2 << 61;                           // layout type: SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N = 256, M = 128.
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(tc::BN >> 3) << 17) | ((uint32_t)(tc::BM >> 4) << 24);
}

struct TcGemmArgs {
    const int* n_rows;
    const float *bias, *bn_scale, *bn_shift;
    float* out;
    int K, H;
};

__global__ void __launch_bounds__(tc::THREADS, 1)
fc1_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
              const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
              const TcGemmArgs a) {
    using namespace tc;
    const int rows = *a.n_rows;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= rows) return;                                   // uniform per CTA

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = a.K / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                tma_load_2d(st, &map_ah, &full[s], kb * BK, m0);
                tma_load_2d(st + A_BYTES, &map_al, &full[s], kb * BK, m0);
                tma_load_2d(st + 2 * A_BYTES, &map_wh, &full[s], kb * BK, n0);
                tma_load_2d(st + 2 * A_BYTES + B_BYTES, &map_wl, &full[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc();
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t dah = make_desc_sw128(st), dal = make_desc_sw128(st + A_BYTES);
                const uint64_t dwh = make_desc_sw128(st + 2 * A_BYTES), dwl = make_desc_sw128(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / UK; ++kk) {
                    const uint64_t adv = (uint64_t)((kk * UK * 2) >> 4);     // +32 bytes per K step inside the atom
                    umma_bf16(tmem_d, dah + adv, dwh + adv, idesc, (kb | kk) != 0);
                    umma_bf16(tmem_d, dah + adv, dwl + adv, idesc, 1);
                    umma_bf16(tmem_d, dal + adv, dwh + adv, idesc, 1);
                }
                umma_commit(&empty[s]);                       // frees the stage when the MMAs above have read it
            }
            umma_commit(tmem_full);                           // accumulator complete
        }
    } else {
        // ===== epilogue: 4 warps, each owns the TMEM lane quadrant (warp % 4) =====
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + q * 32 + lane;
        float* orow = a.out + (size_t)row * a.H + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (row < rows) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    float* po = reinterpret_cast<float*>(&o);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int n = n0 + c0 + j + t;
                        const float x = fmaxf(__uint_as_float(v[j + t]) + __ldg(a.bias + n), 0.f);
                        po[t] = fmaf(x, __ldg(a.bn_scale + n), __ldg(a.bn_shift + n));
                    }
                    *reinterpret_cast<float4*>(orow + c0 + j) = o;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
    }
}

// fp32 -> (hi, lo) bf16 split of the activation matrix, rows < *n_rows only.
__global__ void split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, const int* n_rows, int K) {
    const size_t total = (size_t)(*n_rows) * K / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        const float f[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            h[t] = __float2bfloat16_rn(f[t]);
            l[t] = __float2bfloat16_rn(f[t] - __bfloat162float(h[t]));
        }
        reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
        reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
    }
}

// ---- host side ---------------------------------------------------------------------------------------
struct TcImpl {
    __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr, *w_hi = nullptr, *w_lo = nullptr;
    CUtensorMap map_ah, map_al, map_wh, map_wl;
};

static std::string g_tc_err;
const char* pose_tc_error() { return g_tc_err.c_str(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2-D bf16 tensor [rows][K] (K contiguous), box = [box_rows][BK], 128-byte swizzle.
static int make_map(CUtensorMap* m, void* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { g_tc_err = "cuTensorMapEncodeTiled not available"; return -1; }
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_tc_err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return -1; }
    return 0;
}

static inline uint16_t f2bf(float f) {            // round-to-nearest-even, like __float2bfloat16_rn
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

int pose_tc_init(PoseTc* t, const float* w_kh, int K, int H, int rows_cap, cudaStream_t) {
    pose_tc_free(t);
    t->K = K; t->H = H; t->rows_cap = rows_cap;
    if (K % tc::BK != 0 || H % tc::BN != 0) { g_tc_err = "K/H not tileable"; return -1; }
    TcImpl* im = new TcImpl();
    t->impl = im;
    const size_t rows_pad = ((size_t)rows_cap + tc::BM - 1) / tc::BM * tc::BM;
    const size_t an = rows_pad * K, wn = (size_t)H * K;
    if (cudaMalloc((void**)&im->a_hi, an * 2) != cudaSuccess || cudaMalloc((void**)&im->a_lo, an * 2) != cudaSuccess ||
        cudaMalloc((void**)&im->w_hi, wn * 2) != cudaSuccess || cudaMalloc((void**)&im->w_lo, wn * 2) != cudaSuccess) {
        g_tc_err = "cudaMalloc failed";
        return -1;
    }
    cudaMemset(im->a_hi, 0, an * 2);
    cudaMemset(im->a_lo, 0, an * 2);
    // W is Keras (in, out) = [K][H]; the B operand wants it K-major per output column: [H][K]
    std::vector<uint16_t> hi(wn), lo(wn);
    for (int k = 0; k < K; ++k)
        for (int h = 0; h < H; ++h) {
            const float w = w_kh[(size_t)k * H + h];
            const uint16_t bh = f2bf(w);
            hi[(size_t)h * K + k] = bh;
            lo[(size_t)h * K + k] = f2bf(w - bf2f(bh));
        }
    cudaMemcpy(im->w_hi, hi.data(), wn * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(im->w_lo, lo.data(), wn * 2, cudaMemcpyHostToDevice);
    if (make_map(&im->map_ah, im->a_hi, rows_pad, K, tc::BM) || make_map(&im->map_al, im->a_lo, rows_pad, K, tc::BM) ||
        make_map(&im->map_wh, im->w_hi, H, K, tc::BN) || make_map(&im->map_wl, im->w_lo, H, K, tc::BN))
        return -1;
    if (cudaFuncSetAttribute(fc1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES) != cudaSuccess) {
        g_tc_err = "cudaFuncSetAttribute(fc1_tc_kernel) failed";
        return -1;
    }
    t->ready = true;
    return 0;
}

int pose_tc_fc1(PoseTc* t, const FcArgs& a, int max_rows, cudaStream_t st, int* n_launches) {
    TcImpl* im = reinterpret_cast<TcImpl*>(t->impl);
    if (!im || !t->ready) { g_tc_err = "tensor-core path not initialised"; return -1; }
    split_bf16_kernel<<<148 * 8, 256, 0, st>>>(a.A, im->a_hi, im->a_lo, a.n_rows, a.K);
    TcGemmArgs g{a.n_rows, a.bias, a.bn_scale, a.bn_shift, a.out, a.K, a.H};
    dim3 grid(a.H / tc::BN, (max_rows + tc::BM - 1) / tc::BM);
    fc1_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(im->map_ah, im->map_al, im->map_wh, im->map_wl, g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_tc_err = cudaGetErrorString(e); return -1; }
    if (n_launches) *n_launches = 2;
    return 0;
}

void pose_tc_free(PoseTc* t) {
    TcImpl* im = reinterpret_cast<TcImpl*>(t->impl);
    if (im) {
        for (void* p : {(void*)im->a_hi, (void*)im->a_lo, (void*)im->w_hi, (void*)im->w_lo})
            if (p) cudaFree(p);
        delete im;
    }
    t->impl = nullptr;
    t->ready = false;
}

}  // namespace mmw
