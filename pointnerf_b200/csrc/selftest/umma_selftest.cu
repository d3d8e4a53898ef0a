// Self-test of the tcgen05 building blocks used by the fused shading kernel: one CTA computes
//   D[128 x N] = A[128 x K] * W[N x K]^T    with the BF16x3 error-compensated split
//   (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32 accumulation in TMEM)
// from fp32 inputs, operands staged in shared memory in the selected K-major layout.  tests/test_gpu_umma.py
// compares D against an fp32/fp64 matmul: this pins descriptors, layouts, TMEM addressing and the split
// accuracy on the real hardware before the big kernel depends on them.
//
// TEST-ONLY library (libpnb200_selftest.so): these entry points are not part of the product ABI (include/pnb200.h); they are
// declared in pnb200_selftest.h next to this file and used by tests/test_gpu_umma.py and tools/umma_*.py.
#include <stdarg.h>

#include "../common.cuh"
#include "../umma.cuh"
#include "pnb200_selftest.h"

namespace pnb {
static thread_local char g_selftest_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_selftest_err, sizeof(g_selftest_err), fmt, ap);
    va_end(ap);
}
}  // namespace pnb
extern "C" const char* pnb_selftest_last_error(void) { return pnb::g_selftest_err; }

namespace pnb {
using namespace umma;

template <int LAYOUT>
__global__ void __launch_bounds__(128, 1) k_umma_selftest(const float* __restrict__ A, const float* __restrict__ W,
                                                          float* __restrict__ D, int K, int N, int* err) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms need 1 KB alignment
    // A_hi | A_lo : nkb blocks of [128 x 32] (8 KB each);  B_hi | B_lo : one block of [N x 32] (N*64 B each)
    const int nkb = (K + BK - 1) / BK;
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + nkb * 8192;
    unsigned char* b_hi = a_lo + nkb * 8192;
    unsigned char* b_lo = b_hi + 256 * 64;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<256>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = tmem_base;

    // stage A (all K blocks)
    for (int i = tid; i < 128 * nkb * BK; i += 128) {
        int r = i / (nkb * BK), k = i - r * (nkb * BK);
        float v = k < K ? A[(size_t)r * K + k] : 0.f;
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        uint32_t off = (uint32_t)(k / BK) * 8192u + tile_offset_bytes<LAYOUT>(r, k % BK);
        *(__nv_bfloat16*)(a_hi + off) = h;
        *(__nv_bfloat16*)(a_lo + off) = l;
    }
    const uint32_t idesc = make_idesc_bf16(128, N);
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
        for (int i = tid; i < N * BK; i += 128) {
            int n = i / BK, k = i - n * BK;
            int kg = kb * BK + k;
            float v = kg < K ? W[(size_t)n * K + kg] : 0.f;
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            uint32_t off = tile_offset_bytes<LAYOUT>(n, k);
            *(__nv_bfloat16*)(b_hi + off) = h;
            *(__nv_bfloat16*)(b_lo + off) = l;
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const int nks = (min(K - kb * BK, BK) + 15) / 16;
            for (int ks = 0; ks < nks; ++ks) {
                uint32_t adv = kstep_advance_bytes<LAYOUT>(ks);
                uint64_t dah = make_smem_desc<LAYOUT>(smem_u32(a_hi + kb * 8192) + adv);
                uint64_t dal = make_smem_desc<LAYOUT>(smem_u32(a_lo + kb * 8192) + adv);
                uint64_t dbh = make_smem_desc<LAYOUT>(smem_u32(b_hi) + adv);
                uint64_t dbl = make_smem_desc<LAYOUT>(smem_u32(b_lo) + adv);
                mma_ss(tacc, dah, dbh, idesc, (kb | ks) ? 1u : 0u);
                mma_ss(tacc, dal, dbh, idesc, 1u);
                mma_ss(tacc, dah, dbl, idesc, 1u);
            }
            mma_commit(&bar);
        }
        // everyone waits for the MMAs of this block before B is overwritten
        if (!mbar_wait(&bar, phase, err, 100 + kb)) break;
        phase ^= 1;
        tc_fence_after();
        __syncthreads();
    }
    // epilogue: warp w reads TMEM lanes 32w..32w+31
    {
        const int row = warp * 32 + (tid & 31);
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tacc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < N) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tacc);
}

// Same product with the A operand in TENSOR MEMORY (tcgen05.mma "TS" form): A_hi at columns 256.., A_lo at 384..
// (two bf16 per 32-bit column, element 2c in the low half), written with tcgen05.st; the last 16 K columns can be
// routed through shared memory instead (ss_tail = 1) to pin mixing TS and SS MMAs on one accumulator.
template <int LAYOUT>
__global__ void __launch_bounds__(128, 1) k_umma_selftest_ts(const float* __restrict__ A, const float* __restrict__ W,
                                                             float* __restrict__ D, int K, int N, int ss_tail, int* err) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* b_hi = smem;
    unsigned char* b_lo = b_hi + 256 * 64;
    unsigned char* t_hi = b_lo + 256 * 64;   // [128 x 32] tail block (only 16 columns used)
    unsigned char* t_lo = t_hi + 8192;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = (K + BK - 1) / BK;
    const int Kts = ss_tail ? K - 16 : K;    // K columns served from TMEM

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = tmem_base;
    const uint32_t tlane = (uint32_t)(warp * 32) << 16;
    // A -> TMEM: thread = row, 16 K elements (8 packed columns) per store
    {
        const int row = warp * 32 + lane;
        for (int k0 = 0; k0 < 256; k0 += 16) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int k = k0 + 2 * i;
                float a = k < Kts ? A[(size_t)row * K + k] : 0.f, b = k + 1 < Kts ? A[(size_t)row * K + k + 1] : 0.f;
                split_bf16x2(a, b, h[i], l[i]);
            }
            tmem_st8(tacc + tlane + 256u + (uint32_t)(k0 >> 1), h);
            tmem_st8(tacc + tlane + 384u + (uint32_t)(k0 >> 1), l);
        }
        tmem_st_wait();
        if (ss_tail) {
            for (int k = 0; k < 32; ++k) {
                int kg = Kts + k;
                float v = (k < 16 && kg < K) ? A[(size_t)row * K + kg] : 0.f;
                __nv_bfloat16 hh, ll;
                split_bf16(v, hh, ll);
                uint32_t off = tile_offset_bytes<LAYOUT>(row, k);
                *(__nv_bfloat16*)(t_hi + off) = hh;
                *(__nv_bfloat16*)(t_lo + off) = ll;
            }
        }
    }
    tc_fence_before();
    const uint32_t idesc = make_idesc_bf16(128, N);
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
        for (int i = tid; i < N * BK; i += 128) {
            int n = i / BK, k = i - n * BK;
            int kg = kb * BK + k;
            float v = kg < K ? W[(size_t)n * K + kg] : 0.f;
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            uint32_t off = tile_offset_bytes<LAYOUT>(n, k);
            *(__nv_bfloat16*)(b_hi + off) = h;
            *(__nv_bfloat16*)(b_lo + off) = l;
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const int nks = (min(K - kb * BK, BK) + 15) / 16;
            for (int ks = 0; ks < nks; ++ks) {
                const int kcol = kb * BK + ks * 16;
                uint32_t adv = kstep_advance_bytes<LAYOUT>(ks);
                uint64_t dbh = make_smem_desc<LAYOUT>(smem_u32(b_hi) + adv);
                uint64_t dbl = make_smem_desc<LAYOUT>(smem_u32(b_lo) + adv);
                uint32_t acc_flag = (kb | ks) ? 1u : 0u;
                if (kcol < Kts) {
                    uint32_t ah = tacc + 256u + (uint32_t)(kcol >> 1), al = tacc + 384u + (uint32_t)(kcol >> 1);
                    mma_ts(tacc, ah, dbh, idesc, acc_flag);
                    mma_ts(tacc, al, dbh, idesc, 1u);
                    mma_ts(tacc, ah, dbl, idesc, 1u);
                } else {
                    uint64_t dah = make_smem_desc<LAYOUT>(smem_u32(t_hi));
                    uint64_t dal = make_smem_desc<LAYOUT>(smem_u32(t_lo));
                    mma_ss(tacc, dah, dbh, idesc, acc_flag);
                    mma_ss(tacc, dal, dbh, idesc, 1u);
                    mma_ss(tacc, dah, dbl, idesc, 1u);
                }
            }
            mma_commit(&bar);
        }
        if (!mbar_wait(&bar, phase, err, 200 + kb)) break;
        phase ^= 1;
        tc_fence_after();
        __syncthreads();
    }
    {
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tacc + tlane + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < N) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tacc);
}

// CTA-pair form (cluster of 2, tcgen05 cta_group::2): D[256 x N] = A[256 x K] * W[N x K]^T.  CTA r stages A rows
// 128r.. (shared memory, mode 0, or tensor memory columns 256../384.., mode 1) and W rows (N/2)r.. ; the rank-0 CTA
// issues the MMAs and commits with a multicast arrive to the barrier of both CTAs.  bench_iters > 0 additionally times
// that many back-to-back MMAs (out[0] issue cycles, out[1] until complete) after D has been written.
template <int LAYOUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    k_umma_selftest2(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D, int K, int N, int mode,
                     int bench_iters, int bench_flags, long long* __restrict__ out, int* err) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nkb = (K + BK - 1) / BK;
    unsigned char* b_hi = smem;                  // [N/2 x 32]
    unsigned char* b_lo = b_hi + 128 * 64;
    unsigned char* a_hi = b_lo + 128 * 64;       // nkb blocks of [128 x 32] (mode 0)
    unsigned char* a_lo = a_hi + nkb * 8192;
    __shared__ uint64_t bar, dummy_bar, spin_bar, bbar[2];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int NH = N >> 1;

    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&dummy_bar, 1); mbar_init(&spin_bar, 1); mbar_init(&bbar[0], 1); mbar_init(&bbar[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc2<512>(&tmem_base);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tacc = tmem_base;
    const uint32_t tlane = (uint32_t)(warp * 32) << 16;
    const int row = warp * 32 + lane, grow = (int)rank * 128 + row;
    if (mode == 0) {
        for (int k = 0; k < nkb * BK; ++k) {
            float v = k < K ? A[(size_t)grow * K + k] : 0.f;
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            uint32_t off = (uint32_t)(k / BK) * 8192u + tile_offset_bytes<LAYOUT>(row, k % BK);
            *(__nv_bfloat16*)(a_hi + off) = h;
            *(__nv_bfloat16*)(a_lo + off) = l;
        }
    } else {
        for (int k0 = 0; k0 < 256; k0 += 16) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int k = k0 + 2 * i;
                float a = k < K ? A[(size_t)grow * K + k] : 0.f, b = k + 1 < K ? A[(size_t)grow * K + k + 1] : 0.f;
                split_bf16x2(a, b, h[i], l[i]);
            }
            tmem_st8(tacc + tlane + 256u + (uint32_t)(k0 >> 1), h);
            tmem_st8(tacc + tlane + 384u + (uint32_t)(k0 >> 1), l);
        }
        tmem_st_wait();
    }
    const uint32_t idesc = make_idesc_bf16(256, N);
    const uint32_t hiw = desc_hi<LAYOUT>();
    constexpr uint32_t KADV = kstep_adv16<LAYOUT>();
    uint32_t phase = 0;
    bool ok = true;
    for (int kb = 0; kb < nkb && ok; ++kb) {
        for (int i = tid; i < NH * BK; i += 128) {
            int n = i / BK, k = i - n * BK;
            int kg = kb * BK + k;
            float v = kg < K ? W[(size_t)((int)rank * NH + n) * K + kg] : 0.f;
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            uint32_t off = tile_offset_bytes<LAYOUT>(n, k);
            *(__nv_bfloat16*)(b_hi + off) = h;
            *(__nv_bfloat16*)(b_lo + off) = l;
        }
        fence_proxy_async();
        tc_fence_before();
        cluster_sync_all();
        if (rank == 0 && warp == 0) {
            tc_fence_after();
            const int nks = (min(K - kb * BK, BK) + 15) / 16;
            for (int ks = 0; ks < nks; ++ks) {
                const uint32_t bh = desc_lo<LAYOUT>(smem_u32(b_hi)) + ks * KADV, bl = desc_lo<LAYOUT>(smem_u32(b_lo)) + ks * KADV;
                const uint32_t acc_flag = (kb | ks) ? 1u : 0u;
                if (mode == 0) {
                    const uint32_t ah = desc_lo<LAYOUT>(smem_u32(a_hi + kb * 8192)) + ks * KADV, al = desc_lo<LAYOUT>(smem_u32(a_lo + kb * 8192)) + ks * KADV;
                    mma2_ss2_w(tacc, ah, hiw, bh, hiw, idesc, acc_flag);
                    mma2_ss2_w(tacc, al, hiw, bh, hiw, idesc, 1u);
                    mma2_ss2_w(tacc, ah, hiw, bl, hiw, idesc, 1u);
                } else {
                    const uint32_t kcol = (uint32_t)(kb * BK + ks * 16);
                    mma2_ts2_w(tacc, tacc + 256u + (kcol >> 1), bh, hiw, idesc, acc_flag);
                    mma2_ts2_w(tacc, tacc + 384u + (kcol >> 1), bh, hiw, idesc, 1u);
                    mma2_ts2_w(tacc, tacc + 256u + (kcol >> 1), bl, hiw, idesc, 1u);
                }
            }
            mma2_commit_w(&bar, 3);
        }
        if (!mbar_wait(&bar, phase, err, 400 + kb)) ok = false;
        phase ^= 1;
        tc_fence_after();
    }
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tacc + tlane + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < N) D[(size_t)grow * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (bench_iters > 0 && ok) {
        long long t0 = 0, t1 = 0;
        if (rank == 0 && warp == 0) {
            const uint32_t bh = desc_lo<LAYOUT>(smem_u32(b_hi)), ah = desc_lo<LAYOUT>(smem_u32(a_hi));
            t0 = clock64();
            const int cevery = bench_flags & 255, ckind = (bench_flags >> 8) & 3;
            const uint32_t rot = (bench_flags >> 16) & 1;      // rotate the operands: no two consecutive MMAs share A or B
            const uint32_t bdelta = desc_lo<LAYOUT>(smem_u32(b_lo)) - bh, adelta = desc_lo<LAYOUT>(smem_u32(a_lo)) - ah;
            for (int i = 0; i < bench_iters; ++i) {
                const uint32_t sel = rot ? (uint32_t)(i & 1) : 0u, ksel = rot ? (uint32_t)((i >> 1) & 1) * KADV : 0u;
                if (mode == 0) mma2_ss2_w(tacc, ah + sel * adelta + ksel, hiw, bh + sel * bdelta + ksel, hiw, idesc, 1u);
                else mma2_ts2_w(tacc, tacc + 256u + sel * 128u + (ksel ? 8u : 0u), bh + sel * bdelta + ksel, hiw, idesc, 1u);
                if (cevery && (i & (cevery - 1)) == cevery - 1) {        // interleaved commits to a barrier nobody waits on (cevery: power of 2)
                    if (ckind == 0) mma2_commit_w(&dummy_bar, 3);
                    else if (ckind == 1) mma2_commit_w(&dummy_bar, 1);
                    else if (ckind == 2) mma2_commit_local_w(&dummy_bar);
                    else mma_commit_w(&dummy_bar);
                }
            }
            mma2_commit_w(&bar, 3);
            t1 = clock64();
        }
        // optional stressors on warps 1-3 of BOTH CTAs while the MMAs run (what else the fused kernel does on the SM)
        if (warp == 1 && (bench_flags & (1 << 13)) && lane == 0) {          // weight-like bulk copies into shared memory
            unsigned char* sink[2] = {a_lo, b_lo};
            uint32_t cnt = 0;
            while (!mbar_test_wait(&bar, phase)) {
                const uint32_t sidx = cnt & 1u;
                if (cnt >= 2 && !mbar_wait(&bbar[sidx], ((cnt >> 1) - 1) & 1u, err, 451)) break;
                mbar_arrive_expect_tx(&bbar[sidx], 8192);
                bulk_g2s(sink[sidx], reinterpret_cast<const unsigned char*>(W) + (size_t)(cnt & 3u) * 8192, 8192, &bbar[sidx]);
                ++cnt;
            }
            for (uint32_t c = (cnt >= 2 ? cnt - 2 : 0); c < cnt; ++c) mbar_wait(&bbar[c & 1u], (c >> 1) & 1u, err, 452);
        } else if ((warp == 2 || (warp == 3 && !(bench_flags & (3 << 14)))) && (bench_flags & (1 << 12))) {   // epilogue-like TMEM traffic
            uint32_t v[16];
            while (!mbar_test_wait(&bar, phase)) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    tmem_ld16(tacc + tlane + 384u + (uint32_t)(c * 16), v);
                    tmem_ld_wait();
                    tmem_st8(tacc + tlane + 384u + (uint32_t)(c * 16), v);
                    tmem_st8(tacc + tlane + 384u + (uint32_t)(c * 16) + 8u, v + 8);
                    tmem_st_wait();
                }
            }
        } else if (warp == 3 && (bench_flags & (1 << 14))) {               // a warp polling an mbarrier with test_wait
            while (!mbar_test_wait(&bar, phase)) { if (mbar_test_wait(&dummy_bar, 1u)) __nanosleep(0); }
        } else if (warp == 3 && (bench_flags & (1 << 15))) {               // ... with try_wait
            while (!mbar_test_wait(&bar, phase)) { if (mbar_try_wait(&spin_bar, 0u)) break; }
        }
        mbar_wait(&bar, phase, err, 450);
        if (rank == 0 && tid == 0) { out[0] = t1 - t0; out[1] = clock64() - t0; }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2<512>(tacc);
}

// Micro-benchmark: cycles per tcgen05.mma (M=128, N=256, K=16, bf16) issued back to back on resident operands.
// mode 0: A and B in shared memory (SS), mode 1: A in tensor memory (TS).  bulk = 1 adds a concurrent stream of
// 16 KB cp.async.bulk copies into a second shared buffer (the weight-streaming traffic of the fused kernel).
template <int LAYOUT>
__global__ void __launch_bounds__(128, 1) k_umma_bench(int mode, int iters, int bulk, const unsigned char* __restrict__ gsrc,
                                                       long long* __restrict__ out, int* err) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a = smem;                 // [128 x 32]
    unsigned char* b = smem + 8192;          // [256 x 32]
    unsigned char* sink = smem + 8192 + 16384;   // 2 x 16 KB landing zone for bulk copies
    __shared__ uint64_t bar, bbar[2], cbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (8192 + 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bbar[0], 1); mbar_init(&bbar[1], 1); mbar_init(&cbar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = tmem_base;
    if (tid == 32 && (bulk & 1)) {       // concurrent weight-like stream
        for (int i = 0; i < iters / 3 + 2; ++i) {
            int s = i & 1;
            if (i >= 2 && !mbar_wait(&bbar[s], ((i >> 1) - 1) & 1, err, 301)) break;
            mbar_arrive_expect_tx(&bbar[s], 16384);
            bulk_g2s(sink + s * 16384, gsrc + (size_t)(i % 64) * 16384, 16384, &bbar[s]);
        }
    }
    __shared__ volatile int stop_flag;
    if (tid == 0) stop_flag = 0;
    __syncthreads();
    if ((bulk & 2) && warp >= 1) {     // concurrent epilogue-like TMEM traffic on columns 384..511 (other lane quadrants)
        const uint32_t tl = ((uint32_t)(warp * 32) << 16) - ((bulk & 4) ? 112u : 0u) - ((bulk & 8) ? 368u : 0u);   // +4: same 128-column block as the A operand; +8: inside the accumulator block
        uint32_t v[16];
        uint32_t acc_x = 0;
        while (!stop_flag) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                tmem_ld16(tacc + tl + 384u + (uint32_t)(c * 16), v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) acc_x += v[e];
                v[0] = acc_x;
                tmem_st8(tacc + tl + 384u + (uint32_t)(c * 16), v);
                tmem_st8(tacc + tl + 384u + (uint32_t)(c * 16) + 8u, v + 8);
                tmem_st_wait();
            }
        }
        if (acc_x == 0x12345678u) out[3] = acc_x;
    }
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 256);
        const uint64_t da = make_smem_desc<LAYOUT>(smem_u32(a)), db = make_smem_desc<LAYOUT>(smem_u32(b));
        long long t0 = clock64();
        const int cevery = (bulk >> 8) & 255;       // interleave a commit (to a barrier nobody waits on) every n MMAs (power of 2)
        const uint32_t rot = (bulk >> 16) & 1;      // rotate operands between two buffers / k-steps
        const uint64_t db2 = make_smem_desc<LAYOUT>(smem_u32(sink)), da2 = make_smem_desc<LAYOUT>(smem_u32(sink + 16384));
        for (int i = 0; i < iters; ++i) {
            const bool alt = rot && (i & 1);
            if (mode == 0) mma_ss(tacc, alt ? da2 : da, alt ? db2 : db, idesc, 1u);
            else mma_ts(tacc, tacc + 256u + (alt ? 128u : 0u), alt ? db2 : db, idesc, 1u);
            if (cevery && (i & (cevery - 1)) == cevery - 1) mma_commit(&cbar);
        }
        mma_commit(&bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0, err, 300);
        long long t2 = clock64();
        out[0] = t1 - t0;      // issue time
        out[1] = t2 - t0;      // until all MMAs completed
        stop_flag = 1;
    }
    __syncthreads();
    if (tid == 32 && (bulk & 1)) { mbar_wait(&bbar[0], 1, nullptr, 0); }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tacc);
}

}  // namespace pnb

using namespace pnb;

// layout: 0 = interleaved (no swizzle), 4 = 64B swizzle.  N multiple of 16 in [16,256], K <= 288.
extern "C" int pnb_umma_selftest(const float* d_A, const float* d_W, float* d_D, int K, int N, int layout, int* d_err,
                                 pnb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PNB_REQUIRE(d_A && d_W && d_D && d_err, PNB_ERR_INVALID, "pnb_umma_selftest: null argument");
    PNB_REQUIRE(K >= 1 && K <= 288 && N >= 16 && N <= 256 && N % 16 == 0, PNB_ERR_INVALID, "pnb_umma_selftest: bad K/N");
    const int nkb = (K + 31) / 32;
    size_t smem = (size_t)nkb * 8192 * 2 + 256 * 64 * 2 + 1024;
    if (layout == umma::LAYOUT_NONE) {
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_selftest<umma::LAYOUT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_umma_selftest<umma::LAYOUT_NONE><<<1, 128, smem, stream>>>(d_A, d_W, d_D, K, N, d_err);
    } else if (layout == umma::LAYOUT_SW64) {
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_selftest<umma::LAYOUT_SW64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_umma_selftest<umma::LAYOUT_SW64><<<1, 128, smem, stream>>>(d_A, d_W, d_D, K, N, d_err);
    } else if (layout == 100 || layout == 101) {   // A operand in tensor memory (TS form); 101: last 16 K columns via SS
        PNB_REQUIRE(K % 16 == 0 && K <= 272 && (layout == 101 || K <= 256) && (layout == 100 || K >= 32), PNB_ERR_INVALID,
                    "pnb_umma_selftest: TS mode needs K %% 16 == 0 and K <= 256 (+16 with the SS tail)");
        size_t smem_ts = 256 * 64 * 2 + 8192 * 2 + 1024;
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_selftest_ts<umma::LAYOUT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ts));
        k_umma_selftest_ts<umma::LAYOUT_NONE><<<1, 128, smem_ts, stream>>>(d_A, d_W, d_D, K, N, layout == 101 ? 1 : 0, d_err);
    } else {
        PNB_REQUIRE(false, PNB_ERR_INVALID, "pnb_umma_selftest: layout %d not supported", layout);
    }
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}

// layout 0 / 4 (interleaved / SW64); mode 0 SS, 1 TS; d_out: int64[2] (issue cycles, total cycles); d_src >= 1 MB.
extern "C" int pnb_umma_bench(int layout, int mode, int iters, int bulk, const void* d_src, long long* d_out, int* d_err,
                              pnb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t smem = 8192 + 16384 + 32768 + 1024;
    if (layout == umma::LAYOUT_SW64) {
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_bench<umma::LAYOUT_SW64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_umma_bench<umma::LAYOUT_SW64><<<1, 128, smem, stream>>>(mode, iters, bulk, (const unsigned char*)d_src, d_out, d_err);
    } else {
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_bench<umma::LAYOUT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_umma_bench<umma::LAYOUT_NONE><<<1, 128, smem, stream>>>(mode, iters, bulk, (const unsigned char*)d_src, d_out, d_err);
    }
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}

// CTA-pair self-test / micro-benchmark (cluster of 2, cta_group::2): d_A [256 x K], d_W [N x K], d_D [256 x N];
// mode 0 = A in shared memory (K <= 288), 1 = A in tensor memory (K <= 256, K % 16 == 0); N % 32 == 0.
// bench_iters > 0: d_out int64[2] = issue cycles / cycles until complete of that many back-to-back MMAs;
// bench_flags: bits 0-7 = interleave a tcgen05.commit every that many MMAs, bits 8-9 = its form (0 multicast to both
// CTAs, 1 multicast mask 1, 2 cta_group::2 without multicast, 3 cta_group::1).
extern "C" int pnb_umma_selftest2(const float* d_A, const float* d_W, float* d_D, int K, int N, int mode, int bench_iters,
                                  int bench_flags, long long* d_out, int* d_err, pnb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PNB_REQUIRE(d_A && d_W && d_D && d_err, PNB_ERR_INVALID, "pnb_umma_selftest2: null argument");
    PNB_REQUIRE(K >= 16 && K <= (mode ? 256 : 288) && (mode == 0 || K % 16 == 0) && N >= 32 && N <= 256 && N % 32 == 0, PNB_ERR_INVALID,
                "pnb_umma_selftest2: bad K/N");
    PNB_REQUIRE(bench_iters == 0 || d_out, PNB_ERR_INVALID, "pnb_umma_selftest2: bench needs d_out");
    const int nkb = (K + 31) / 32;
    size_t smem = (size_t)nkb * 8192 * 2 + 128 * 64 * 2 + 1024;
    PNB_CHECK_CUDA(cudaFuncSetAttribute(k_umma_selftest2<umma::LAYOUT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest2<umma::LAYOUT_NONE><<<2, 128, smem, stream>>>(d_A, d_W, d_D, K, N, mode, bench_iters, bench_flags, d_out, d_err);
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}
