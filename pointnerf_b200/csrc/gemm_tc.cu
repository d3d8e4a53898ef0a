// Generic fp32-in / fp32-out GEMM on the 5th-generation tensor cores (tcgen05 + TMEM) with the BF16x3 error-compensated split
// of shade_tc.cu -- the GEMM engine of the BACKWARD pass (backward.cu): forward recompute, dX = dZ . W and dW = X^T . dZ.
//
//   C[M x N] = op( sum_k A(m,k) * B(n,k) + bias[n] )        A(m,k) = A[m*a_rs + k*a_ks],  B(n,k) = B[n*b_rs + k*b_ks]
//
// Either stride of an operand may be the unit one: operands whose reduction index is NOT contiguous in memory (W^T buffers,
// the activations of dW = X^T dZ) are transposed on the way into shared memory, so the tensor cores only ever see the K-major
// core-matrix layout that umma.cuh pins (tests/test_gpu_umma.py).  One CTA = one 128 x 128 output tile (UMMA M = 128, N = 128,
// fp32 accumulator in 128 TMEM columns; two CTAs per SM), K blocks of 32 through a 2-stage shared-memory ring:
//   warps 0-3  load the fp32 rows of A (thread = tile row), split into bf16 hi / lo, 16-byte stores into the operand layout; the next
//              K block is already in registers while the current one is converted; afterwards the epilogue (thread = TMEM lane)
//   warps 4-7  the same for the rows of B
//   warp  8    issues 6 tcgen05.mma per K block (A_hi W_hi, A_lo W_hi, A_hi W_lo for the two K=16 steps), one commit per block
// PRECISE mode (forward recompute only): three bf16 parts per operand and the six products of total order <= 2 (a0b0, a1b0, a0b1,
// a2b0, a1b1, a0b2): fp32-level pre-activations, so that the LeakyReLU masks of the backward are those of an fp32 forward (with the
// BF16x3 recompute ~1e-5 of the units land on the other side of zero and each such flip changes a pair's gradient by ~1/256).
// Split-K (dW: the reduction runs over the ~1e5 pair rows of a training batch): blockIdx.z owns a K range and writes its partial
// tile to a workspace; k_splitk_reduce adds the partials in split order (deterministic, no atomics).
#include "common.cuh"
#include "gemm_tc.cuh"
#include "umma.cuh"

namespace pnb {
using namespace umma;

namespace gtc {
constexpr int TM = 128, TN = 128, LOG_NSTAGE = 1, NSTAGE = 1 << LOG_NSTAGE;    // 2 stages (64 KB): two CTAs per SM keep twice the loads in flight
constexpr int BLK = 128 * 64;             // [128 x 32] bf16 block
constexpr int NTHR = 288;                 // warps 0-3: A loaders + epilogue, 4-7: B loaders, 8: issuer
template <int NPART>
struct Smem {
    unsigned char a[NPART][NSTAGE][BLK], b[NPART][NSTAGE][BLK];     // part 0 = hi, 1 = lo (mid), 2 = lo of the 3-part split
    uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_acc;
    uint32_t tmem_base;
};
// 32 K elements of row `r` of an operand into registers.  rs / ks: element strides of the row / reduction index.
__device__ __forceinline__ void fetch_row(const float* __restrict__ P, long rs, long ks, long r, bool row_ok, int k0, int kend, float* v) {
    if (!row_ok || k0 >= kend) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
    } else if (ks == 1 && k0 + 32 <= kend) {
        const float4* src = reinterpret_cast<const float4*>(P + r * rs + k0);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) { const float4 x = __ldg(src + e4); v[4 * e4] = x.x; v[4 * e4 + 1] = x.y; v[4 * e4 + 2] = x.z; v[4 * e4 + 3] = x.w; }
    } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = (k0 + e < kend) ? __ldg(P + r * rs + (long)(k0 + e) * ks) : 0.f;
    }
}
// registers -> NPART bf16 blocks of stage s (tile row t)
template <int NPART>
__device__ __forceinline__ void store_row(float* v, unsigned char (*blk)[NSTAGE][BLK], int s, int t) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t off = tile_offset_bytes<LAYOUT_NONE>(t, 8 * c);
#pragma unroll
        for (int part = 0; part < NPART; ++part) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * c + 2 * i], v[8 * c + 2 * i + 1]);
                v[8 * c + 2 * i] -= __low2float(h); v[8 * c + 2 * i + 1] -= __high2float(h);          // the residual feeds the next part
                w[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(blk[part][s] + off) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}
}  // namespace gtc

template <bool PRECISE>
__global__ void __launch_bounds__(gtc::NTHR, 2) k_gemm_tc(GemmTc g) {
    using namespace gtc;
    constexpr int NPART = PRECISE ? 3 : 2;
    using SmemT = Smem<NPART>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SmemT& sm = *reinterpret_cast<SmemT*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int kbeg = blockIdx.z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
    const int nkb = (kend - kbeg + 31) >> 5;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&sm.bar_full[s], 8); mbar_init(&sm.bar_empty[s], 1); }   // one arrive per loader warp
        mbar_init(&sm.bar_acc, 1);
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc<128>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = sm.tmem_base;

    if (warp == 8) {
        const uint32_t idesc = make_idesc_bf16(128, 128);
        const uint32_t hiw = desc_hi<LAYOUT_NONE>();
        uint32_t a0[NPART], b0[NPART];
#pragma unroll
        for (int i = 0; i < NPART; ++i) { a0[i] = desc_lo<LAYOUT_NONE>(smem_u32(sm.a[i][0])); b0[i] = desc_lo<LAYOUT_NONE>(smem_u32(sm.b[i][0])); }
        constexpr uint32_t KADV = kstep_adv16<LAYOUT_NONE>(), SADV = BLK >> 4;
        bool ok = true;
        for (int kb = 0; kb < nkb && ok; ++kb) {
            const uint32_t s = (uint32_t)kb & (NSTAGE - 1), ph = ((uint32_t)kb >> LOG_NSTAGE) & 1u;
            if (!mbar_wait(&sm.bar_full[s], ph, g.err, 71)) { ok = false; break; }
            tc_fence_after();
            bool first = kb == 0;
#pragma unroll
            for (int i = 0; i < NPART; ++i)
#pragma unroll
                for (int j = 0; j < NPART; ++j) {
                    if (i + j >= NPART) continue;          // products of total order < NPART: 3 (BF16x3) or 6 (3-part split)
                    const uint32_t ad = a0[i] + s * SADV, bd = b0[j] + s * SADV;
                    mma_ss2_w(tacc, ad, hiw, bd, hiw, idesc, first ? 0u : 1u);
                    mma_ss2_w(tacc, ad + KADV, hiw, bd + KADV, hiw, idesc, 1u);
                    first = false;
                }
            mma_commit_w(&sm.bar_empty[s]);
        }
        mma_commit_w(&sm.bar_acc);
    } else {
        // loaders: warps 0-3 own the A rows (thread = tile row), warps 4-7 the B rows; the NEXT K block is fetched into registers
        // before the current one is converted and stored, so a full K block of loads is in flight per thread
        const bool isB = warp >= 4;
        const int t = tid & 127;
        const float* P = isB ? g.B : g.A;
        const long rs = isB ? g.b_rs : g.a_rs, ks = isB ? g.b_ks : g.a_ks;
        const long r = (isB ? n0 : m0) + t;
        const bool row_ok = r < (isB ? g.N : g.M);
        unsigned char (*blk)[NSTAGE][BLK] = isB ? sm.b : sm.a;
        float cur[32], nxt[32];
        fetch_row(P, rs, ks, r, row_ok, kbeg, kend, cur);
        bool ok = true;
        for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t s = (uint32_t)kb & (NSTAGE - 1), ph = ((uint32_t)kb >> LOG_NSTAGE) & 1u;
            if (kb + 1 < nkb) fetch_row(P, rs, ks, r, row_ok, kbeg + 32 * (kb + 1), kend, nxt);
            if (!mbar_wait(&sm.bar_empty[s], ph ^ 1u, g.err, 72)) { ok = false; break; }
            store_row<NPART>(cur, blk, (int)s, t);
            fence_proxy_async();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&sm.bar_full[s]);   // (32 same-address arrives would serialise in the shared-memory atomic unit)
#pragma unroll
            for (int e = 0; e < 32; ++e) cur[e] = nxt[e];
        }
        if (!isB && ok && mbar_wait(&sm.bar_acc, 0u, g.err, 73)) {
            tc_fence_after();
            const long ra = r;
            const bool a_ok = row_ok;
            const uint32_t tl = tacc + ((uint32_t)(warp * 32) << 16);
            float* crow = g.part ? g.part + ((size_t)blockIdx.z * g.M + (size_t)ra) * g.N : g.C + ra * g.ldc;
#pragma unroll 1
            for (int c = 0; c < TN / 16; ++c) {
                const int n = n0 + 16 * c;
                uint32_t v[16];
                tmem_ld16(tl + (uint32_t)(16 * c), v);      // all lanes of the warp (sync.aligned), also those of rows >= M
                tmem_ld_wait();
                if (!a_ok || n >= g.N) continue;
                float y[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) y[e] = __uint_as_float(v[e]);
                if (!g.part) {
                    if (g.bias) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) y[e] += __ldg(g.bias + n + e);
                    }
                    if (g.act) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) y[e] = fmaxf(y[e], 0.01f * y[e]);
                    }
                    if (g.dact && n < g.dact_n) {            // backward of the LeakyReLU below: * (Y > 0 ? 1 : 0.01)
                        const float4* yp = reinterpret_cast<const float4*>(g.dact + ra * g.ldd + n);
#pragma unroll
                        for (int e4 = 0; e4 < 4; ++e4) {
                            const float4 yy = __ldg(yp + e4);
                            y[4 * e4] *= yy.x > 0.f ? 1.0f : 0.01f; y[4 * e4 + 1] *= yy.y > 0.f ? 1.0f : 0.01f;
                            y[4 * e4 + 2] *= yy.z > 0.f ? 1.0f : 0.01f; y[4 * e4 + 3] *= yy.w > 0.f ? 1.0f : 0.01f;
                        }
                    }
                }
                float4* dst = reinterpret_cast<float4*>(crow + n);
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) dst[e4] = make_float4(y[4 * e4], y[4 * e4 + 1], y[4 * e4 + 2], y[4 * e4 + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<128>(sm.tmem_base);
}

// C[m][n] (+)= sum_z part[z][m][n], z ascending
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ part, int splits, int M, int N, float* __restrict__ C, long ldc, int accumulate) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)M * N) return;
    const int m = (int)(i / N), n = (int)(i - (long)m * N);
    float s = accumulate ? C[m * ldc + n] : 0.f;
    for (int z = 0; z < splits; ++z) s += part[(size_t)z * M * N + i];
    C[m * ldc + n] = s;
}

int gemm_tc(const GemmTc& g0, int splits, float* part_ws, size_t part_bytes, int accumulate, cudaStream_t st) {
    GemmTc g = g0;
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return PNB_OK;
    PNB_REQUIRE(g.N % 16 == 0 && g.ldc % 4 == 0, PNB_ERR_INVALID, "gemm_tc: N (%d) must be a multiple of 16 and ldc (%ld) of 4", g.N, g.ldc);
    PNB_REQUIRE((g.a_rs == 1) != (g.a_ks == 1) || g.K == 1 || g.M == 1, PNB_ERR_INVALID, "gemm_tc: one stride of A must be 1");
    PNB_REQUIRE(g.a_ks != 1 || g.a_rs % 4 == 0, PNB_ERR_INVALID, "gemm_tc: leading dimension of A must be a multiple of 4");
    PNB_REQUIRE(g.b_ks != 1 || g.b_rs % 4 == 0, PNB_ERR_INVALID, "gemm_tc: leading dimension of B must be a multiple of 4");
    static int configured[64] = {0};
    int dev = 0;
    PNB_CHECK_CUDA(cudaGetDevice(&dev));
    const bool precise = g.precise != 0;
    const size_t smem = (precise ? sizeof(gtc::Smem<3>) : sizeof(gtc::Smem<2>)) + 128;
    static_assert(sizeof(gtc::Smem<3>) + 128 <= 232448, "precise GEMM tile exceeds the shared memory of an SM");
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(gtc::Smem<2>) + 128)));
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_gemm_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(gtc::Smem<3>) + 128)));
        configured[dev] = 1;
    }
    if (splits <= 1) {
        g.kchunk = g.K; g.part = nullptr;
        dim3 grid((g.N + gtc::TN - 1) / gtc::TN, (g.M + gtc::TM - 1) / gtc::TM, 1);
        PNB_REQUIRE(!accumulate, PNB_ERR_INVALID, "gemm_tc: accumulation needs the split-K path");
        if (precise) k_gemm_tc<true><<<grid, gtc::NTHR, smem, st>>>(g);
        else k_gemm_tc<false><<<grid, gtc::NTHR, smem, st>>>(g);
    } else {
        int kchunk = ((g.K + splits - 1) / splits + 31) / 32 * 32;
        splits = (g.K + kchunk - 1) / kchunk;
        PNB_REQUIRE(part_ws && part_bytes >= (size_t)splits * g.M * g.N * sizeof(float), PNB_ERR_WORKSPACE, "gemm_tc: split-K workspace too small");
        g.kchunk = kchunk; g.part = part_ws;
        dim3 grid((g.N + gtc::TN - 1) / gtc::TN, (g.M + gtc::TM - 1) / gtc::TM, splits);
        if (precise) k_gemm_tc<true><<<grid, gtc::NTHR, smem, st>>>(g);
        else k_gemm_tc<false><<<grid, gtc::NTHR, smem, st>>>(g);
        const long n = (long)g.M * g.N;
        k_splitk_reduce<<<(int)((n + 255) / 256), 256, 0, st>>>(part_ws, splits, g.M, g.N, g.C, g.ldc, accumulate);
    }
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}

}  // namespace pnb
