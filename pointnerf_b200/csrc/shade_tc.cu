// Fused per-pair shading on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   gather + weights + positional encoding  ->  block1 (284->256->256)  ->  cat extras  ->  block3 (263->256->256)
//   ->  alpha branch + K-reduction  (reference: /root/reference/models/aggregators/point_aggregators.py:488-628,
//   727-814; gather /root/reference/models/neural_points/neural_points.py:706-717)
//
// One persistent CTA per SM; a tile is 16 valid samples x 8 neighbour slots = 128 (sample,k) pair rows = UMMA M.
// Every layer is D[128x256] = A[128xK] * W[256xK]^T on tcgen05.mma (kind::f16, BF16 operands, FP32 accumulate in
// TMEM).  The reference computes these layers in fp32 (cuBLAS SGEMM, TF32 off) and the parity bar is 1e-4 on the
// rendered radiance, which a single BF16/TF32 pass misses -> error-compensated split:
//   A = A_hi + A_lo, W = W_hi + W_lo (bf16 each);   D = A_hi*W_hi + A_lo*W_hi + A_hi*W_lo    (3 MMAs per k-step)
// Warp roles (320 threads):
//   warps 0-7  workers : build the layer-1 operand (gather, PE, split) and run the epilogues
//                        (tcgen05.ld -> bias -> LeakyReLU -> split -> next layer's A operand in shared memory)
//   warp  8    loader  : streams the pre-packed weight images (16 KB each, already in the UMMA operand layout)
//                        L2 -> shared memory with cp.async.bulk (TMA engine) through a 4-deep mbarrier ring
//   warp  9    issuer  : one elected thread issues tcgen05.mma and commits to mbarriers
// A-operand: 9 K-blocks of [128 x 32] bf16 (hi and lo), interleaved core-matrix layout (see umma.cuh).
#include "common.cuh"
#include "umma.cuh"

namespace pnb {
using namespace umma;

namespace tc {
constexpr int LAYOUT = LAYOUT_NONE;
constexpr int TM = 128;                 // pair rows per tile
constexpr int TSAMP = TM / PNB_MAX_K;   // 16 samples per tile
constexpr int NWORK = 256;              // worker threads
constexpr int NTHR = 320;
constexpr int NSTAGE = 4;
constexpr int IMG = 256 * 64;           // bytes of one weight image ([256 x 32] bf16)
constexpr int ABLK = 128 * 64;          // bytes of one A block ([128 x 32] bf16)
constexpr int NKB_MAX = 9;
__host__ __device__ constexpr int nkb_of(int l) { return (l == 0 || l == 2) ? 9 : 8; }
__host__ __device__ constexpr int img_base(int l) { return l == 0 ? 0 : l == 1 ? 9 : l == 2 ? 17 : 26; }  // in blocks
constexpr int NBLK_TOTAL = 34;
constexpr int IMGS_PER_TILE = 2 * NBLK_TOTAL;
constexpr float LEAKY = 0.01f;

struct Smem {
    unsigned char a_hi[NKB_MAX * ABLK];
    unsigned char a_lo[NKB_MAX * ABLK];
    unsigned char b[NSTAGE][IMG];
    float bias[4][256];
    float wa[256];
    float E[TM][8];
    float wc[TM];
    float alpha_part[2][TM];
    uint32_t samp[TSAMP];
    uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_a_ready, bar_acc_full;
    uint32_t tmem_base;
    int abort;
};
}  // namespace tc

struct ShadeTcParams {
    pnb_query_t q;
    pnb_points_t pts;
    pnb_shade_opts_t o;
    const unsigned char* wimg;   // packed weight images (N = 256 per image)
    const float* bias[4];
    const float* wa;             // alpha_branch.0 weight [256]
    const float* ba;             // alpha_branch.0 bias [1]
    float* hbar;                 // [n_valid][256]
    float* sigma;                // [n_valid]
    int hbar_cap;
    int* err;
    int dbg_no_weights;          // timing experiment only: the loader signals the ring without copying (results are garbage)
    int dbg_flags;               // bit 1: v6 issuer classifies its waits with non-blocking probes (profiling)
    const unsigned char* vcnt;   // v7 row packing: neighbours per PACKED position [n_valid] (vcntp of k_pack_quads)
    const uint32_t* vorder;      // v7: valid-sample index of every packed position [n_valid]
    const uint32_t* quad_first;  // v7: first valid sample of every 32-row quadrant [n_quads + 1]
    const int* pack_cnt;         // v7: [0] = n_quads
    int hbar_fmt;                // 0: hbar[n_valid][256] fp32;  1: bf16 hi/lo A-operand blocks of k_color_tc2 (per 128 samples: 8 K blocks x {hi,lo} x [128x32])
};
__device__ __forceinline__ void prof_add(const ShadeTcParams& p, int slot, long long cyc) {
    if ((p.dbg_flags & 1) && blockIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.err) + 1 + slot, (unsigned long long)cyc);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void rot3t(const float* M, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = x * M[0] + y * M[1] + z * M[2];
    oy = x * M[3] + y * M[4] + z * M[5];
    oz = x * M[6] + y * M[7] + z * M[8];
}
__device__ __forceinline__ void w2pers_t(const pnb_shade_opts_t& o, float px, float py, float pz, float& xp, float& yp, float& zp) {
    float sx = px - o.campos[0], sy = py - o.campos[1], sz = pz - o.campos[2];
    const float* M = o.camrotc2w;
    float xc = sx * M[0] + sy * M[3] + sz * M[6];
    float yc = sx * M[1] + sy * M[4] + sz * M[7];
    float zc = sx * M[2] + sy * M[5] + sz * M[8];
    xp = xc / zc; yp = yc / zc; zp = zc;
}

// sin/cos of x*2^j, j = 0..NF-1: one accurate sincosf + angle doubling (abs error < 1e-6 after 4 doublings; the
// fp32 kernel keeps NF independent sincosf calls).  out[2j] = sin, out[2j+1] = cos  (networks.py:175-190 layout).
template <int NF>
__device__ __forceinline__ void pe_doubling(float x, float* out) {
    float s, c;
    sincosf(x, &s, &c);
    out[0] = s; out[1] = c;
#pragma unroll
    for (int j = 1; j < NF; ++j) {
        float s2 = 2.0f * s * c, c2 = fmaf(c, c, -s * s);
        s = s2; c = c2;
        out[2 * j] = s; out[2 * j + 1] = c;
    }
}

// 8 consecutive K elements (one 16-byte chunk) of row r, block kb, starting at k8 (multiple of 8) -> hi / lo buffers
__device__ __forceinline__ void store_chunk8(tc::Smem& sm, int r, int kb, int k8, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16x2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    uint32_t off = (uint32_t)kb * tc::ABLK + tile_offset_bytes<tc::LAYOUT>(r, k8);
    *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(tc::NTHR, 1) k_shade_tc(ShadeTcParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const pnb_query_t& q = p.q;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], p.hbar_cap);
    const int n_tiles = (n_valid + TSAMP - 1) / TSAMP;
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    // ---------------------------------------------------------------- one-time setup
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&sm.bar_full[s], 1); mbar_init(&sm.bar_empty[s], 1); }
        mbar_init(&sm.bar_a_ready, NWORK);
        mbar_init(&sm.bar_acc_full, 1);
        sm.abort = 0;
        mbar_fence_init();
        if (blockIdx.x == 0 && q.counters[PNB_QC_N_VALID] > p.hbar_cap) atomicExch(p.err, 9);   // capacity exceeded
    }
    if (warp == 9) tmem_alloc<256>(&sm.tmem_base);
    for (int i = tid; i < 4 * 256; i += NTHR) sm.bias[i >> 8][i & 255] = p.bias[i >> 8][i & 255];
    for (int i = tid; i < 256; i += NTHR) sm.wa[i] = p.wa[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = sm.tmem_base;

    if (warp == 8) {
        // ============================================================ loader
        if (lane == 0) {
            const uint32_t total = (uint32_t)my_tiles * IMGS_PER_TILE;
            for (uint32_t n = 0; n < total; ++n) {
                const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1u;
                if (!mbar_wait(&sm.bar_empty[s], ph ^ 1u, p.err, 1)) break;
                mbar_arrive_expect_tx(&sm.bar_full[s], IMG);
                bulk_g2s(sm.b[s], p.wimg + (size_t)(n % IMGS_PER_TILE) * IMG, IMG, &sm.bar_full[s]);
            }
        }
    } else if (warp == 9) {
        // ============================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 256);
            uint32_t n = 0, lyr = 0;
            bool ok = true;
            for (int t = 0; t < my_tiles && ok; ++t) {
                for (int l = 0; l < 4 && ok; ++l, ++lyr) {
                    if (!mbar_wait(&sm.bar_a_ready, lyr & 1u, p.err, 2)) { ok = false; break; }
                    tc_fence_after();
                    const int nkb = nkb_of(l);
                    for (int kb = 0; kb < nkb && ok; ++kb) {
                        const int nks = (l == 2 && kb == 8) ? 1 : 2;   // block3 input: 263 -> 272 columns used
                        {   // W_hi image: A_hi*W_hi + A_lo*W_hi
                            const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1u;
                            if (!mbar_wait(&sm.bar_full[s], ph, p.err, 3)) { ok = false; break; }
                            tc_fence_after();
                            for (int ks = 0; ks < nks; ++ks) {
                                const uint32_t adv = kstep_advance_bytes<LAYOUT>(ks);
                                const uint64_t db = make_smem_desc<LAYOUT>(smem_u32(sm.b[s]) + adv);
                                mma_ss(tacc, make_smem_desc<LAYOUT>(smem_u32(sm.a_hi + kb * ABLK) + adv), db, idesc, (kb | ks) ? 1u : 0u);
                                mma_ss(tacc, make_smem_desc<LAYOUT>(smem_u32(sm.a_lo + kb * ABLK) + adv), db, idesc, 1u);
                            }
                            mma_commit(&sm.bar_empty[s]);
                            ++n;
                        }
                        {   // W_lo image: A_hi*W_lo
                            const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1u;
                            if (!mbar_wait(&sm.bar_full[s], ph, p.err, 4)) { ok = false; break; }
                            tc_fence_after();
                            for (int ks = 0; ks < nks; ++ks) {
                                const uint32_t adv = kstep_advance_bytes<LAYOUT>(ks);
                                mma_ss(tacc, make_smem_desc<LAYOUT>(smem_u32(sm.a_hi + kb * ABLK) + adv),
                                       make_smem_desc<LAYOUT>(smem_u32(sm.b[s]) + adv), idesc, 1u);
                            }
                            mma_commit(&sm.bar_empty[s]);
                            ++n;
                        }
                    }
                    mma_commit(&sm.bar_acc_full);
                }
            }
        }
    } else {
        // ============================================================ workers (warps 0..7)
        const int quad = warp & 3, half = warp >> 2;
        const int erow = quad * 32 + lane;                 // epilogue row = TMEM lane
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        uint32_t lyr = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            // ------------------------------------------------ build the block1 operand: 2 threads per pair row
            {
                const int row = warp * 16 + (lane >> 1), hf = lane & 1;
                const int si = row >> 3, k = row & 7;
                const int vi = tile * TSAMP + si;
                uint32_t s = 0xffffffffu;
                int pidx = -1;
                float lx = 0.f, ly = 0.f, lz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
                if (vi < n_valid) {
                    s = q.valid_list[vi];
                    uint32_t pk = q.samp_ray[s];
                    int r = (int)(pk >> 7), j = (int)(pk & 127u);
                    int d = q.steps[(size_t)r * q.SR + j];
                    float tt = q.t[(size_t)r * q.t_ray_stride + d];
                    vx = q.raydir[3 * r]; vy = q.raydir[3 * r + 1]; vz = q.raydir[3 * r + 2];
                    lx = raypos1(q.campos[0], vx, tt); ly = raypos1(q.campos[1], vy, tt); lz = raypos1(q.campos[2], vz, tt);
                    if (k < q.K) pidx = q.cand_pidx[(size_t)s * q.K + k];
                }
                if ((lane & 15) == 0) sm.samp[si] = s;
                const bool valid = pidx >= 0;
                const int pi = valid ? pidx : 0;
                float ovx, ovy, ovz;
                rot3t(p.o.Rw2c, vx, vy, vz, ovx, ovy, ovz);
                float px = __ldg(&p.pts.xyz[3 * pi]), py = __ldg(&p.pts.xyz[3 * pi + 1]), pz = __ldg(&p.pts.xyz[3 * pi + 2]);
                float dist[6];
                dist[0] = px - lx; dist[1] = py - ly; dist[2] = pz - lz;
                float xpp, ypp, zpp, xsp, ysp, zsp;
                w2pers_t(p.o, px, py, pz, xpp, ypp, zpp);
                w2pers_t(p.o, lx, ly, lz, xsp, ysp, zsp);
                dist[3] = xpp * zpp - xsp * zsp;
                dist[4] = ypp * zpp - ysp * zsp;
                dist[5] = zpp - zsp;
                float nrm = sqrtf(dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2]);
                float w = valid ? 1.0f / fmaxf(nrm, 1e-6f) : 0.f;
                float wsum = hf == 0 ? w : 0.f;   // the 16 lanes of one sample: sum over its 8 rows
                wsum += __shfl_xor_sync(0xffffffffu, wsum, 1);
                wsum += __shfl_xor_sync(0xffffffffu, wsum, 2);
                wsum += __shfl_xor_sync(0xffffffffu, wsum, 4);
                wsum += __shfl_xor_sync(0xffffffffu, wsum, 8);
                w = w / fmaxf(wsum, 1e-8f);
                float cf = __ldg(&p.pts.conf[pi]);
                float cc = fminf(fmaxf(cf, 1e-4f), 1.0f);
                if (hf == 0) sm.wc[row] = valid ? w * cc : 0.f;
                float d0, d1, d2;
                rot3t(p.o.Rw2c, dist[0], dist[1], dist[2], d0, d1, d2);
                dist[0] = d0; dist[1] = d1; dist[2] = d2;
                if (valid) {
                    // raw features hf*16 .. +15 -> columns hf*16.. ; their PE -> columns 32 + 96*hf .. +95
                    const float4* ep = (const float4*)&p.pts.emb[(size_t)pi * PNB_FEAT + hf * 16];
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 v = __ldg(ep + i);
                        f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
                    }
                    store_chunk8(sm, row, 0, hf * 16, f);
                    store_chunk8(sm, row, 0, hf * 16 + 8, f + 8);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {      // 4 features -> 24 PE values -> 3 chunks
                        float pe[24];
#pragma unroll
                        for (int e = 0; e < 4; ++e) pe_doubling<3>(f[g * 4 + e], pe + e * 6);
                        const int col = 32 + 96 * hf + 24 * g;    // multiple of 8
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            int cc0 = col + 8 * c;
                            store_chunk8(sm, row, cc0 >> 5, cc0 & 31, pe + 8 * c);
                        }
                    }
                    // distance PE: index i = d*5 + j -> columns 224 + 2i ; thread hf covers i in [16*hf, 16*hf+16):
                    // hf 0: d = 0,1,2 (+ d=3, j=0) ; hf 1: d = 3 (j>=1), 4, 5 ; 32 values each (4 zeros pad hf 1)
                    {
                        float dp[30];   // 3 distances x 10 values
                        const int dbase = hf * 3;
#pragma unroll
                        for (int e = 0; e < 3; ++e) pe_doubling<5>(dist[dbase + e], dp + 10 * e);
                        float vals[32];
                        if (hf == 0) {
#pragma unroll
                            for (int i = 0; i < 30; ++i) vals[i] = dp[i];
                            float sn, cs;
                            sincosf(dist[3], &sn, &cs);
                            vals[30] = sn; vals[31] = cs;
                        } else {
                            // i = 16..29 -> (d=3, j=1..4), (d=4, j=0..4), (d=5, j=0..4): dp holds d=3,4,5
#pragma unroll
                            for (int i = 0; i < 28; ++i) vals[i] = dp[2 + i];
                            vals[28] = 0.f; vals[29] = 0.f; vals[30] = 0.f; vals[31] = 0.f;
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            int cc0 = 224 + 32 * hf + 8 * c;
                            store_chunk8(sm, row, cc0 >> 5, cc0 & 31, vals + 8 * c);
                        }
                    }
                    if (hf == 1) {
                        float cr = __ldg(&p.pts.color[3 * pi]), cg = __ldg(&p.pts.color[3 * pi + 1]), cb = __ldg(&p.pts.color[3 * pi + 2]);
                        float ddx, ddy, ddz;
                        rot3t(p.o.Rw2c, __ldg(&p.pts.dir[3 * pi]), __ldg(&p.pts.dir[3 * pi + 1]), __ldg(&p.pts.dir[3 * pi + 2]), ddx, ddy, ddz);
                        sm.E[row][0] = cr; sm.E[row][1] = cg; sm.E[row][2] = cb;
                        sm.E[row][3] = ddx - ovx; sm.E[row][4] = ddy - ovy; sm.E[row][5] = ddz - ovz;
                        sm.E[row][6] = ddx * ovx + ddy * ovy + ddz * ovz;
                        sm.E[row][7] = 0.f;
                    }
                } else {
                    float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int c = hf; c < 36; c += 2) store_chunk8(sm, row, c >> 2, (c & 3) * 8, z8);   // 288 columns
                    if (hf == 1)
                        for (int e = 0; e < 8; ++e) sm.E[row][e] = 0.f;
                }
            }
            fence_proxy_async();
            mbar_arrive(&sm.bar_a_ready);

            // ------------------------------------------------ 4 layers: epilogues
            for (int l = 0; l < 4 && ok; ++l, ++lyr) {
                if (!mbar_wait(&sm.bar_acc_full, lyr & 1u, p.err, 5)) { ok = false; break; }
                tc_fence_after();
                float apart = 0.f;
                if (l < 3) {
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        const int c0 = half * 128 + ch * 32;
                        uint32_t v[32];
                        tmem_ld32(tacc + tlane + (uint32_t)c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float x[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float y = __uint_as_float(v[g * 8 + e]) + sm.bias[l][c0 + g * 8 + e];
                                x[e] = fmaxf(y, LEAKY * y);
                            }
                            store_chunk8(sm, erow, c0 >> 5, g * 8, x);
                        }
                    }
                    if (l == 1 && half == 0) {   // block3 extras -> columns 256..271 (zero padded)
                        float e0[8], e1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int e = 0; e < 8; ++e) e0[e] = sm.E[erow][e];
                        store_chunk8(sm, erow, 8, 0, e0);
                        store_chunk8(sm, erow, 8, 8, e1);
                    }
                    tc_fence_before();
                    fence_proxy_async();
                    mbar_arrive(&sm.bar_a_ready);
                } else {
                    // last layer: h stays in registers; alpha branch + weighted K-reduction
                    const float wrow = sm.wc[erow];
                    const int sidx = tile * TSAMP + (erow >> 3);
                    const bool swrite = sidx < n_valid;
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        const int c0 = half * 128 + ch * 32;
                        uint32_t v[32];
                        tmem_ld32(tacc + tlane + (uint32_t)c0, v);
                        tmem_ld_wait();
                        float mine[4];
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            float y = __uint_as_float(v[e]) + sm.bias[3][c0 + e];
                            y = fmaxf(y, LEAKY * y);
                            apart = fmaf(y, sm.wa[c0 + e], apart);
                            float z = y * wrow;
                            z += __shfl_xor_sync(0xffffffffu, z, 1);
                            z += __shfl_xor_sync(0xffffffffu, z, 2);
                            z += __shfl_xor_sync(0xffffffffu, z, 4);
                            if ((e & 7) == (lane & 7)) mine[e >> 3] = z;
                        }
                        if (swrite) {
                            float* dst = p.hbar + (size_t)sidx * 256 + c0 + (lane & 7);
#pragma unroll
                            for (int g = 0; g < 4; ++g) dst[8 * g] = mine[g];
                        }
                    }
                    tc_fence_before();
                    sm.alpha_part[half][erow] = apart;
                    named_bar_sync(1, NWORK);
                    if (half == 0) {
                        float a = sm.alpha_part[0][erow] + sm.alpha_part[1][erow] + __ldg(p.ba) - 1.0f;
                        float sp = a > 20.f ? a : log1pf(expf(a));
                        float z = sp * wrow;
                        z += __shfl_xor_sync(0xffffffffu, z, 1);
                        z += __shfl_xor_sync(0xffffffffu, z, 2);
                        z += __shfl_xor_sync(0xffffffffu, z, 4);
                        if ((lane & 7) == 0 && swrite) p.sigma[sidx] = z;
                    }
                    named_bar_sync(1, NWORK);   // alpha_part / wc / E are reused by the next tile's build
                }
            }
        }
    }
    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<256>(tacc);
}

// =====================================================================================================================
// v3: layers 2-4 read their A operand from TENSOR MEMORY (tcgen05.mma TS form) so the shared-memory operand buffer
// is only needed by layer 1 -> dedicated builder warps construct the NEXT tile's layer-1 operand while the tensor
// core runs layers 2-4 of the current tile.  TMEM: accumulator cols 0..255, A_hi 256..383, A_lo 384..511 (two bf16
// per 32-bit column).  Warp roles (448 threads): 0-7 epilogue (TMEM -> bias/LeakyReLU/split -> TMEM), 8-11 builders,
// 12 loader, 13 issuer.  The 7 block3 extras go through a small [128 x 16] shared-memory operand (one SS k-step).
// Optional in-kernel cycle accounting (block 0 only): err[2 + 2*i], 64-bit counters, see tools/tc_profile.py
// Enabled by dbg_flags bit 0 (tools/tc_profile.py), a kernel parameter: the clock reads and atomics slow block 0 down by
// several per cent, and a persistent kernel is as slow as its slowest CTA.
#define PNB_TIMED_WAIT_L0(slot, expr) [&]() { long long _t0 = clock64(); bool _r = (expr); if ((threadIdx.x & 31) == 0) prof_add(p, slot, clock64() - _t0); return _r; }()
#define PNB_TIMED_WAIT(slot, expr) [&]() { long long _t0 = clock64(); bool _r = (expr); prof_add(p, slot, clock64() - _t0); return _r; }()

namespace tc3 {
constexpr int NEPI_WARPS = 16, NEPI = NEPI_WARPS * 32, NBUILD = 128, NTHR = NEPI + NBUILD + 64;   // 704 threads
constexpr int NSTAGE = 4;
constexpr int XE = 128 * 32;            // bytes of one [128 x 16] bf16 extras operand (SBO = 256)
struct Smem {
    static constexpr int NWC = 2;
    unsigned char a_hi[tc::NKB_MAX * tc::ABLK];
    unsigned char a_lo[tc::NKB_MAX * tc::ABLK];
    unsigned char b[NSTAGE][tc::IMG];
    unsigned char xe_hi[2][XE];
    unsigned char xe_lo[2][XE];
    float wc[2][tc::TM];
    float alpha_part[2][tc::TM];
    uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_a1_ready, bar_a1_free, bar_acc_full, bar_at_ready;
    uint32_t tmem_base;
};
__device__ __forceinline__ uint32_t xe_offset(int r, int k) { return (uint32_t)((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2); }
__device__ __forceinline__ uint64_t xe_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)(128 >> 4) << 16;    // LBO: next 8-column chunk
    d |= (uint64_t)(256 >> 4) << 32;    // SBO: next 8-row group
    d |= (uint64_t)1 << 46;
    return d;                           // layout type 0 (interleaved)
}
}  // namespace tc3

template <class SmemT>
__device__ __forceinline__ void store_chunk8_a1(SmemT& sm, int r, int kb, int k8, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16x2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    uint32_t off = (uint32_t)kb * tc::ABLK + tile_offset_bytes<tc::LAYOUT>(r, k8);
    *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// One pair row of the block1 operand (gather, distance weights, PE by angle doubling, hi/lo split -> shared memory),
// its block3 extras operand and its weight*conf factor.  PART 2: the whole row (v3 / v5: one builder thread per row);
// PART 0 / 1: the two halves of a row for the v6 pipeline (two builder threads per row):
//   part 0 = operand columns 0..151   (raw features, PE of features 0..19)
//   part 1 = operand columns 152..287 (PE of features 20..31, PE of the 6 distances), extras operand, weight*conf
// Sum of v over the lanes st .. st+cnt-1 of the warp (cnt <= 8; every lane calls it, lanes outside any segment pass st = lane,
// cnt = 1).  The order of the additions depends on cnt only, not on where the segment sits in the warp.
__device__ __forceinline__ float seg_scan8(float v, int lane, int st) {          // inclusive scan within the segment
#pragma unroll
    for (int d = 1; d <= 4; d <<= 1) {
        const float tv = __shfl_up_sync(0xffffffffu, v, d);
        if (lane - d >= st) v += tv;
    }
    return v;
}
__device__ __forceinline__ float seg_sum8(float v, int lane, int st, int cnt) {
    return __shfl_sync(0xffffffffu, seg_scan8(v, lane, st), st + cnt - 1);
}

// PACKED = false: row = 8 * (sample in tile) + neighbour slot, 16 samples per tile (rows of empty slots are zero).
// PACKED = true (v7): the caller has packed only the valid (sample, neighbour) pairs into the rows: row `row` is neighbour `pk` of
// valid sample `pvi` (pvi < 0: unused row), and the rows of that sample are lanes pst .. pst+pcnt-1 of this warp.
template <int PART, bool PACKED = false, class SmemT>
__device__ __forceinline__ void build_pair_part(SmemT& sm, const ShadeTcParams& p, int tile, int t, int row, int n_valid, int pvi = -1,
                                                int pk = 0, int pst = 0, int pcnt = 1) {
    using namespace tc;
    constexpr bool P0 = PART != 1, P1 = PART != 0;
    constexpr int G_LO = P0 ? 0 : 5, G_HI = P1 ? 8 : 5;      // feature groups (4 features each) whose PE this part builds
    const pnb_query_t& q = p.q;
    const int k = PACKED ? pk : (row & 7);
    const int vi = PACKED ? pvi : tile * TSAMP + (row >> 3);
    int pidx = -1;
    float lx = 0.f, ly = 0.f, lz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
    if (vi >= 0 && vi < n_valid) {
        uint32_t s = q.valid_list[vi];
        if (P1) {
            uint32_t pk = q.samp_ray[s];
            int r = (int)(pk >> 7), j = (int)(pk & 127u);
            int d = q.steps[(size_t)r * q.SR + j];
            float tt = q.t[(size_t)r * q.t_ray_stride + d];
            vx = q.raydir[3 * r]; vy = q.raydir[3 * r + 1]; vz = q.raydir[3 * r + 2];
            lx = raypos1(q.campos[0], vx, tt); ly = raypos1(q.campos[1], vy, tt); lz = raypos1(q.campos[2], vz, tt);
        }
        if (k < q.K) pidx = q.cand_pidx[(size_t)s * q.K + k];
    }
    const bool valid = pidx >= 0;
    const int pi = valid ? pidx : 0;
    float dist[6];
    float ovx = 0.f, ovy = 0.f, ovz = 0.f;
    if (P1) {
        rot3t(p.o.Rw2c, vx, vy, vz, ovx, ovy, ovz);
        float px = __ldg(&p.pts.xyz[3 * pi]), py = __ldg(&p.pts.xyz[3 * pi + 1]), pz = __ldg(&p.pts.xyz[3 * pi + 2]);
        dist[0] = px - lx; dist[1] = py - ly; dist[2] = pz - lz;
        float xpp, ypp, zpp, xsp, ysp, zsp;
        w2pers_t(p.o, px, py, pz, xpp, ypp, zpp);
        w2pers_t(p.o, lx, ly, lz, xsp, ysp, zsp);
        dist[3] = xpp * zpp - xsp * zsp; dist[4] = ypp * zpp - ysp * zsp; dist[5] = zpp - zsp;
        float nrm = sqrtf(dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2]);
        float w = valid ? 1.0f / fmaxf(nrm, 1e-6f) : 0.f;
        float wsum = w;
        if (PACKED) {
            wsum = seg_sum8(w, row & 31, pst, pcnt);
        } else {                               // 8 consecutive lanes = the 8 rows of one sample
            wsum += __shfl_xor_sync(0xffffffffu, wsum, 1);
            wsum += __shfl_xor_sync(0xffffffffu, wsum, 2);
            wsum += __shfl_xor_sync(0xffffffffu, wsum, 4);
        }
        w = w / fmaxf(wsum, 1e-8f);
        float cf = __ldg(&p.pts.conf[pi]);
        sm.wc[t % SmemT::NWC][row] = valid ? w * fminf(fmaxf(cf, 1e-4f), 1.0f) : 0.f;
        float d0, d1, d2;
        rot3t(p.o.Rw2c, dist[0], dist[1], dist[2], d0, d1, d2);
        dist[0] = d0; dist[1] = d1; dist[2] = d2;
    }
    float ex[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) {
        const float4* ep = (const float4*)&p.pts.emb[(size_t)pi * PNB_FEAT];
#pragma unroll
        for (int g = G_LO; g < G_HI; ++g) {    // 4 features per step
            float4 fv = __ldg(ep + g);
            float f[4] = {fv.x, fv.y, fv.z, fv.w};
            float pe[24];
#pragma unroll
            for (int e = 0; e < 4; ++e) pe_doubling<3>(f[e], pe + e * 6);
            const int col = 32 + 24 * g;
#pragma unroll
            for (int c = 0; c < 3; ++c) store_chunk8_a1(sm, row, (col + 8 * c) >> 5, (col + 8 * c) & 31, pe + 8 * c);
        }
        if (P0) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {      // raw features, 8 per chunk (re-read: L1 hit)
                float4 a = __ldg(ep + 2 * g), b = __ldg(ep + 2 * g + 1);
                float f8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                store_chunk8_a1(sm, row, 0, 8 * g, f8);
            }
        }
        if (P1) {
            float dp[60];
#pragma unroll
            for (int e = 0; e < 6; ++e) pe_doubling<5>(dist[e], dp + 10 * e);
            float z4[8] = {dp[56], dp[57], dp[58], dp[59], 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 7; ++c) store_chunk8_a1(sm, row, (224 + 8 * c) >> 5, (224 + 8 * c) & 31, dp + 8 * c);
            store_chunk8_a1(sm, row, 8, 24, z4);
            float ddx, ddy, ddz;
            rot3t(p.o.Rw2c, __ldg(&p.pts.dir[3 * pi]), __ldg(&p.pts.dir[3 * pi + 1]), __ldg(&p.pts.dir[3 * pi + 2]), ddx, ddy, ddz);
            ex[0] = __ldg(&p.pts.color[3 * pi]); ex[1] = __ldg(&p.pts.color[3 * pi + 1]); ex[2] = __ldg(&p.pts.color[3 * pi + 2]);
            ex[3] = ddx - ovx; ex[4] = ddy - ovy; ex[5] = ddz - ovz;
            ex[6] = ddx * ovx + ddy * ovy + ddz * ovz;
        }
    } else {
        float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        constexpr int C_LO = P0 ? 0 : 19, C_HI = P1 ? 36 : 19;   // 16-byte chunks: column / 8
        for (int c = C_LO; c < C_HI; ++c) store_chunk8_a1(sm, row, c >> 2, (c & 3) * 8, z8);
    }
    if (P1) {   // block3 extras operand [128 x 16]: chunk 0 = extras, chunk 1 = 0
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16x2(ex[2 * i], ex[2 * i + 1], h[i], l[i]);
        uint32_t off = tc3::xe_offset(row, 0);
        *reinterpret_cast<uint4*>(sm.xe_hi[t & 1] + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(sm.xe_lo[t & 1] + off) = make_uint4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<uint4*>(sm.xe_hi[t & 1] + off + 128) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sm.xe_lo[t & 1] + off + 128) = make_uint4(0u, 0u, 0u, 0u);
    }
}
template <class SmemT>
__device__ __forceinline__ void build_pair_row(SmemT& sm, const ShadeTcParams& p, int tile, int t, int row, int n_valid) {
    build_pair_part<2, false>(sm, p, tile, t, row, n_valid);
}

__global__ void __launch_bounds__(tc3::NTHR, 1) k_shade_tc3(ShadeTcParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    tc3::Smem& sm = *reinterpret_cast<tc3::Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const pnb_query_t& q = p.q;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], p.hbar_cap);
    const int n_tiles = (n_valid + TSAMP - 1) / TSAMP;
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < tc3::NSTAGE; ++s) { mbar_init(&sm.bar_full[s], 1); mbar_init(&sm.bar_empty[s], 1); }
        mbar_init(&sm.bar_a1_ready, tc3::NBUILD);
        mbar_init(&sm.bar_a1_free, 1);
        mbar_init(&sm.bar_acc_full, 1);
        mbar_init(&sm.bar_at_ready, tc3::NEPI);
        mbar_fence_init();
        if (blockIdx.x == 0 && q.counters[PNB_QC_N_VALID] > p.hbar_cap) atomicExch(p.err, 9);
    }
    constexpr int W_BUILD = tc3::NEPI_WARPS, W_LOAD = W_BUILD + 4, W_ISSUE = W_LOAD + 1;
    if (warp == W_ISSUE) tmem_alloc<512>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = sm.tmem_base;
    const uint32_t t_ahi = tacc + 256u, t_alo = tacc + 384u;
    const long long _tk0 = clock64();

    if (warp == W_LOAD) {
        // ============================================================ loader
        if (lane == 0) {
            const uint32_t total = (uint32_t)my_tiles * IMGS_PER_TILE;
            for (uint32_t n = 0; n < total; ++n) {
                const uint32_t s = n % tc3::NSTAGE, ph = (n / tc3::NSTAGE) & 1u;
                if (!PNB_TIMED_WAIT(0, mbar_wait(&sm.bar_empty[s], ph ^ 1u, p.err, 11))) break;
                if (p.dbg_no_weights) { mbar_arrive(&sm.bar_full[s]); continue; }
                mbar_arrive_expect_tx(&sm.bar_full[s], IMG);
                bulk_g2s(sm.b[s], p.wimg + (size_t)(n % IMGS_PER_TILE) * IMG, IMG, &sm.bar_full[s]);
            }
        }
    } else if (warp == W_ISSUE) {
        // ============================================================ MMA issuer (lean: ~15 instructions per MMA)
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 256);
            const uint32_t hiw = desc_hi<LAYOUT>(), xe_hiw = (256u >> 4) | (1u << 14);
            const uint32_t b0_lo = desc_lo<LAYOUT>(smem_u32(sm.b[0]));
            const uint32_t ahi_lo = desc_lo<LAYOUT>(smem_u32(sm.a_hi)), alo_lo = desc_lo<LAYOUT>(smem_u32(sm.a_lo));
            const uint32_t xeh_lo0 = desc_lo<LAYOUT_NONE>(smem_u32(sm.xe_hi[0])), xel_lo0 = desc_lo<LAYOUT_NONE>(smem_u32(sm.xe_lo[0]));
            constexpr uint32_t KADV = kstep_adv16<LAYOUT>();
            uint32_t n = 0, n_at = 0;
            bool ok = true;
            for (int t = 0; t < my_tiles && ok; ++t) {
                const uint32_t xeh_lo = xeh_lo0 + (uint32_t)(t & 1) * (tc3::XE >> 4), xel_lo = xel_lo0 + (uint32_t)(t & 1) * (tc3::XE >> 4);
                for (int l = 0; l < 4 && ok; ++l) {
                    if (l == 0) {
                        if (!PNB_TIMED_WAIT(1, mbar_wait(&sm.bar_a1_ready, (uint32_t)t & 1u, p.err, 12))) { ok = false; break; }
                        if (t > 0) { if (!PNB_TIMED_WAIT(2, mbar_wait(&sm.bar_at_ready, n_at & 1u, p.err, 13))) { ok = false; break; } ++n_at; }
                    } else {
                        if (!PNB_TIMED_WAIT(2, mbar_wait(&sm.bar_at_ready, n_at & 1u, p.err, 14))) { ok = false; break; }
                        ++n_at;
                    }
                    tc_fence_after();
                    const int nkb = nkb_of(l);
                    for (int kb = 0; kb < nkb && ok; ++kb) {
                        const bool two = !(l == 2 && kb == 8);             // block3 extras block: one k-step
                        const uint32_t akb_hi = ahi_lo + (uint32_t)kb * (ABLK >> 4), akb_lo = alo_lo + (uint32_t)kb * (ABLK >> 4);
                        const uint32_t tcol = (uint32_t)(kb * 16);
                        {   // ---- W_hi image: A_hi*W_hi + A_lo*W_hi
                            const uint32_t s = n & (tc3::NSTAGE - 1), ph = (n >> 2) & 1u;
                            if (!PNB_TIMED_WAIT(3, mbar_wait(&sm.bar_full[s], ph, p.err, 15))) { ok = false; break; }
                            tc_fence_after();
                            const uint32_t bl = b0_lo + s * (IMG >> 4);
                            if (l == 0) {
                                mma_ss2(tacc, akb_hi, hiw, bl, hiw, idesc, kb ? 1u : 0u);
                                mma_ss2(tacc, akb_lo, hiw, bl, hiw, idesc, 1u);
                                mma_ss2(tacc, akb_hi + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                                mma_ss2(tacc, akb_lo + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                            } else if (kb == 8) {
                                mma_ss2(tacc, xeh_lo, xe_hiw, bl, hiw, idesc, 1u);
                                mma_ss2(tacc, xel_lo, xe_hiw, bl, hiw, idesc, 1u);
                            } else {
                                mma_ts2(tacc, t_ahi + tcol, bl, hiw, idesc, kb ? 1u : 0u);
                                mma_ts2(tacc, t_alo + tcol, bl, hiw, idesc, 1u);
                                mma_ts2(tacc, t_ahi + tcol + 8u, bl + KADV, hiw, idesc, 1u);
                                mma_ts2(tacc, t_alo + tcol + 8u, bl + KADV, hiw, idesc, 1u);
                            }
                            mma_commit(&sm.bar_empty[s]);
                            ++n;
                        }
                        {   // ---- W_lo image: A_hi*W_lo
                            const uint32_t s = n & (tc3::NSTAGE - 1), ph = (n >> 2) & 1u;
                            if (!PNB_TIMED_WAIT(3, mbar_wait(&sm.bar_full[s], ph, p.err, 15))) { ok = false; break; }
                            tc_fence_after();
                            const uint32_t bl = b0_lo + s * (IMG >> 4);
                            if (l == 0) {
                                mma_ss2(tacc, akb_hi, hiw, bl, hiw, idesc, 1u);
                                mma_ss2(tacc, akb_hi + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                            } else if (kb == 8) {
                                mma_ss2(tacc, xeh_lo, xe_hiw, bl, hiw, idesc, 1u);
                            } else {
                                mma_ts2(tacc, t_ahi + tcol, bl, hiw, idesc, 1u);
                                mma_ts2(tacc, t_ahi + tcol + 8u, bl + KADV, hiw, idesc, 1u);
                            }
                            mma_commit(&sm.bar_empty[s]);
                            ++n;
                        }
                        (void)two;
                    }
                    mma_commit(&sm.bar_acc_full);
                    if (l == 0) mma_commit(&sm.bar_a1_free);
                }
            }
        }
    } else if (warp >= W_BUILD) {
        // ============================================================ builders: one thread per pair row
        const int row = (warp - W_BUILD) * 32 + lane;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            if (t > 0 && !(lane == 0 && warp == W_BUILD ? PNB_TIMED_WAIT(4, mbar_wait(&sm.bar_a1_free, (uint32_t)(t - 1) & 1u, p.err, 16)) : mbar_wait(&sm.bar_a1_free, (uint32_t)(t - 1) & 1u, p.err, 16))) { ok = false; break; }
            const long long _tb0 = clock64();
            build_pair_row(sm, p, tile, t, row, n_valid);
            fence_proxy_async();
            mbar_arrive(&sm.bar_a1_ready);
            if (lane == 0 && warp == W_BUILD) prof_add(p, 5, clock64() - _tb0);
        }
    } else {
        // ============================================================ epilogue warps 0..15: 64 columns per thread
        const int quad = warp & 3, part = warp >> 2;
        const int erow = quad * 32 + lane;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        uint32_t n_acc = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            for (int l = 0; l < 4 && ok; ++l, ++n_acc) {
                if (!(tid == 0 ? PNB_TIMED_WAIT(6, mbar_wait(&sm.bar_acc_full, n_acc & 1u, p.err, 17)) : mbar_wait(&sm.bar_acc_full, n_acc & 1u, p.err, 17))) { ok = false; break; }
                const long long _te0 = clock64();
                tc_fence_after();
                if (l < 3) {
                    const float* bias = p.bias[l];
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const int c0 = part * 64 + ch * 16;
                        uint32_t v[16];
                        tmem_ld16(tacc + tlane + (uint32_t)c0, v);
                        tmem_ld_wait();
                        uint32_t hh[8], ll[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c0) + e);
                            float y0 = __uint_as_float(v[2 * e]) + bb.x, y1 = __uint_as_float(v[2 * e + 1]) + bb.y;
                            y0 = fmaxf(y0, LEAKY * y0); y1 = fmaxf(y1, LEAKY * y1);
                            split_bf16x2(y0, y1, hh[e], ll[e]);
                        }
                        const uint32_t colp = (uint32_t)(c0 >> 1);
                        tmem_st8(t_ahi + tlane + colp, hh);
                        tmem_st8(t_alo + tlane + colp, ll);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&sm.bar_at_ready);
                    if (tid == 0) prof_add(p, 7, clock64() - _te0);
                } else {
                    const float wrow = sm.wc[t & 1][erow];
                    const int sidx = tile * TSAMP + (erow >> 3);
                    const bool swrite = sidx < n_valid;
                    const float* bias = p.bias[3];
                    const int j8 = lane & 7;
                    float apart = 0.f;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const int c0 = part * 64 + ch * 16;
                        uint32_t v[16];
                        tmem_ld16(tacc + tlane + (uint32_t)c0, v);
                        tmem_ld_wait();
                        float z[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float y = __uint_as_float(v[e]) + __ldg(bias + c0 + e);
                            y = fmaxf(y, LEAKY * y);
                            apart = fmaf(y, __ldg(p.wa + c0 + e), apart);
                            z[e] = y * wrow;
                        }
                        // reduce-scatter over the 8 rows (lanes) of a sample: 16 -> 8 -> 4 -> 2 columns per lane
                        float r8[8], r4[4], r2[2];
                        const bool b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float send = b4 ? z[i] : z[i + 8], keep = b4 ? z[i + 8] : z[i];
                            r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float send = b2 ? r8[i] : r8[i + 4], keep = b2 ? r8[i + 4] : r8[i];
                            r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            float send = b1 ? r4[i] : r4[i + 2], keep = b1 ? r4[i + 2] : r4[i];
                            r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                        }
                        if (swrite) *reinterpret_cast<float2*>(p.hbar + (size_t)sidx * 256 + c0 + 2 * j8) = make_float2(r2[0], r2[1]);
                    }
                    tc_fence_before();
                    mbar_arrive(&sm.bar_at_ready);          // accumulator drained: the next tile's layer 1 may start
                    if (tid == 0) prof_add(p, 8, clock64() - _te0);
                    if (part < 2) sm.alpha_part[part][erow] = apart;
                    named_bar_sync(1, tc3::NEPI);
                    if (part >= 2) atomicAdd(&sm.alpha_part[part - 2][erow], apart);
                    named_bar_sync(1, tc3::NEPI);
                    if (part == 0) {
                        float a = sm.alpha_part[0][erow] + sm.alpha_part[1][erow] + __ldg(p.ba) - 1.0f;
                        float sp = a > 20.f ? a : log1pf(expf(a));
                        float zz = sp * wrow;
                        zz += __shfl_xor_sync(0xffffffffu, zz, 1);
                        zz += __shfl_xor_sync(0xffffffffu, zz, 2);
                        zz += __shfl_xor_sync(0xffffffffu, zz, 4);
                        if (j8 == 0 && swrite) p.sigma[sidx] = zz;
                    }
                    named_bar_sync(1, tc3::NEPI);
                }
            }
        }
    }
    if (tid == 0) prof_add(p, 9, clock64() - _tk0);
    tc_fence_before();
    __syncthreads();
    if (warp == W_ISSUE) tmem_dealloc<512>(tacc);
}

// ------------------------------------------------------------------------------------------ weight packing
// W^T fp32 [Kpad][256] (rows >= K are zero) -> per K-block: hi image then lo image, each [256 x 32] bf16 in the
// UMMA operand layout.
__global__ void __launch_bounds__(256) k_pack_weights(const float* __restrict__ wt, int Kpad, int nkb, int N, unsigned char* __restrict__ out,
                                                      int ldn = -1, int n_off = 0) {
    if (ldn < 0) ldn = N;
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // (kb, n, k)
    if (i >= nkb * N * BK) return;
    int kb = i / (N * BK), rem = i - kb * N * BK;
    int n = rem / BK, k = rem - n * BK;
    int kg = kb * BK + k;
    float v = kg < Kpad ? wt[(size_t)kg * ldn + n_off + n] : 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    uint32_t off = tile_offset_bytes<tc::LAYOUT>(n, k);
    const size_t img = (size_t)N * 64;
    *(__nv_bfloat16*)(out + (size_t)(2 * kb) * img + off) = h;
    *(__nv_bfloat16*)(out + (size_t)(2 * kb + 1) * img + off) = l;
}

// ------------------------------------------------------------------------------------------ colour branch
// Per valid sample: [hbar(256), PE4(view)(24)] -> 128 -> 128 -> 128 -> 3, sigmoid*1.002-0.001 ; fp32 CUDA cores.
// (reference: point_aggregators.py:631-637, 269-273).  Tile = 64 samples.
namespace cb {
constexpr int TR = 64, XS = 292, KC = 16, NTHREADS = 256;
struct Smem {
    float X[TR * XS];
    float Y[TR * 132];
    float W[2][KC * 128];
};
}  // namespace cb

__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}

// C[64 x 128] = act(A[64 x Kp] * Wt[Kp x 128] + b); thread (ty,tx): rows ty*4..+3, cols tx*4 + 64*j (j=0,1)
__device__ __forceinline__ void gemm64x128(const float* __restrict__ A, int sa, float* __restrict__ C, int sc,
                                           const float* __restrict__ Wt, const float* __restrict__ bias, int Kp,
                                           float (*Wst)[cb::KC * 128]) {
    using namespace cb;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int nchunk = Kp / KC;
    {
        const float4* src = (const float4*)Wt;
#pragma unroll
        for (int i = 0; i < 2; ++i) cp16(&Wst[0][(tid + i * NTHREADS) * 4], src + tid + i * NTHREADS);
        asm volatile("cp.async.commit_group;\n" ::);
    }
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) {
            const float4* src = (const float4*)(Wt + (size_t)(c + 1) * KC * 128);
#pragma unroll
            for (int i = 0; i < 2; ++i) cp16(&Wst[(c + 1) & 1][(tid + i * NTHREADS) * 4], src + tid + i * NTHREADS);
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        const float* Wc = Wst[c & 1];
#pragma unroll
        for (int k4 = 0; k4 < KC; k4 += 4) {
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *(const float4*)&A[(ty * 4 + i) * sa + c * KC + k4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float4 w0 = *(const float4*)&Wc[(k4 + kk) * 128 + tx * 4];
                float4 w1 = *(const float4*)&Wc[(k4 + kk) * 128 + tx * 4 + 64];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                    acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
                    acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
                    acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
                    acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float4 b = *(const float4*)&bias[tx * 4 + 64 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 v;
            v.x = acc[i][j * 4 + 0] + b.x; v.y = acc[i][j * 4 + 1] + b.y;
            v.z = acc[i][j * 4 + 2] + b.z; v.w = acc[i][j * 4 + 3] + b.w;
            v.x = v.x > 0.f ? v.x : tc::LEAKY * v.x; v.y = v.y > 0.f ? v.y : tc::LEAKY * v.y;
            v.z = v.z > 0.f ? v.z : tc::LEAKY * v.z; v.w = v.w > 0.f ? v.w : tc::LEAKY * v.w;
            *(float4*)&C[(ty * 4 + i) * sc + tx * 4 + 64 * j] = v;
        }
    }
    __syncthreads();
}

struct ColorParams {
    pnb_query_t q;
    pnb_shade_opts_t o;
    const float* w[4];   // W^T: [288][128] (rows >= 280 zero), [128][128], [128][128], [128][3]
    const float* b[4];
    const float* hbar;
    const float* sigma;
    int hbar_cap;
    float4* sigma_rgb;
};

__global__ void __launch_bounds__(cb::NTHREADS, 1) k_color_branch(ColorParams p) {
    using namespace cb;
    extern __shared__ __align__(16) unsigned char smem_raw2[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw2);
    const int tid = threadIdx.x;
    const pnb_query_t& q = p.q;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], p.hbar_cap);
    const int n_tiles = (n_valid + TR - 1) / TR;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // inputs: 4 threads per sample row
        {
            const int row = tid >> 2, part = tid & 3;
            const int vi = tile * TR + row;
            float* xr = &sm.X[row * XS];
            if (vi < n_valid) {
                const float4* src = (const float4*)(p.hbar + (size_t)vi * 256 + part * 64);
#pragma unroll
                for (int i = 0; i < 16; ++i) *(float4*)&xr[part * 64 + 4 * i] = __ldg(src + i);
                uint32_t s = q.valid_list[vi];
                int r = (int)(q.samp_ray[s] >> 7);
                float ovx, ovy, ovz;
                rot3t(p.o.Rw2c, q.raydir[3 * r], q.raydir[3 * r + 1], q.raydir[3 * r + 2], ovx, ovy, ovz);
                // PE4(view), ori=True layout: sin block (12) then cos block (12), index d*4 + j ; 3 (d,j) per thread
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    int i = part * 3 + e, dd = i >> 2, jj = i & 3;
                    float sn, cs;
                    sincosf((dd == 0 ? ovx : dd == 1 ? ovy : ovz) * (float)(1 << jj), &sn, &cs);
                    xr[256 + i] = sn;
                    xr[268 + i] = cs;
                }
                if (part == 0) { xr[280] = 0.f; xr[281] = 0.f; xr[282] = 0.f; xr[283] = 0.f; xr[284] = 0.f; xr[285] = 0.f; xr[286] = 0.f; xr[287] = 0.f; }
            } else {
                for (int c = part; c < 288; c += 4) xr[c] = 0.f;
            }
        }
        __syncthreads();
        gemm64x128(sm.X, XS, sm.Y, 132, p.w[0], p.b[0], 288, sm.W);
        gemm64x128(sm.Y, 132, sm.X, XS, p.w[1], p.b[1], 128, sm.W);
        gemm64x128(sm.X, XS, sm.Y, 132, p.w[2], p.b[2], 128, sm.W);
        if (tid < TR * 3) {
            const int row = tid / 3, c = tid - row * 3;
            const int vi = tile * TR + row;
            if (vi < n_valid) {
                float a = __ldg(&p.b[3][c]);
                const float* wl = p.w[3];
                for (int k = 0; k < 128; ++k) a = fmaf(sm.Y[row * 132 + k], __ldg(&wl[k * 3 + c]), a);
                float rgb = 1.0f / (1.0f + expf(-a)) * (1.0f + 2.0f * 0.001f) - 0.001f;
                uint32_t s = q.valid_list[vi];
                float* dst = (float*)&p.sigma_rgb[s];
                dst[1 + c] = rgb;
                if (c == 0) dst[0] = p.sigma[vi];
            }
        }
        __syncthreads();
    }
}

// Last epilogue of a tile, one warp's share: chunks G, G+NG, G+2*NG, ... (NCHUNK of them, 16 accumulator columns each) of the
// layer-4 output:
// +bias, LeakyReLU, partial alpha-branch dot product (returned), weight*conf scaling and the K-reduction over the 8 rows of a
// sample as a warp-shuffle reduce-scatter (lane j8 ends up with columns c0+2*j8, +1 of its sample) -> h-bar.
template <int NG, int NCHUNK>
__device__ __forceinline__ float last_chunks(const ShadeTcParams& p, uint32_t accb, int G, float wrow, int sidx, bool swrite, int lane) {
    using namespace tc;
    const float* bias = p.bias[3];
    const int j8 = lane & 7;
    float apart = 0.f;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
        const int c0 = 16 * (G + NG * i);
        uint32_t v[16];
        tmem_ld16(accb + (uint32_t)c0, v);
        tmem_ld_wait();
        float z[16];
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0) + e4), ww = __ldg(reinterpret_cast<const float4*>(p.wa + c0) + e4);
            const float bq[4] = {bb.x, bb.y, bb.z, bb.w}, wq[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
            for (int e1 = 0; e1 < 4; ++e1) {
                const int e = 4 * e4 + e1;
                float y = __uint_as_float(v[e]) + bq[e1];
                y = fmaxf(y, LEAKY * y);
                apart = fmaf(y, wq[e1], apart);
                z[e] = y * wrow;
            }
        }
        float r8[8], r4[4], r2[2];
        const bool b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
            float send = b4 ? z[ii] : z[ii + 8], keep = b4 ? z[ii + 8] : z[ii];
            r8[ii] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            float send = b2 ? r8[ii] : r8[ii + 4], keep = b2 ? r8[ii + 4] : r8[ii];
            r4[ii] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            float send = b1 ? r4[ii] : r4[ii + 2], keep = b1 ? r4[ii + 2] : r4[ii];
            r2[ii] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
        if (swrite) {
            if (p.hbar_fmt) {       // straight into the colour kernel's operand image (bf16 hi / lo, core-matrix layout)
                uint32_t hh, ll;
                split_bf16x2(r2[0], r2[1], hh, ll);
                const int col = c0 + 2 * j8, srow = sidx & 127;
                unsigned char* dst = reinterpret_cast<unsigned char*>(p.hbar) + ((size_t)(sidx >> 7) * 8 + (col >> 5)) * (2 * 8192) +
                                     tile_offset_bytes<LAYOUT_NONE>(srow, col & 31);
                *reinterpret_cast<uint32_t*>(dst) = hh;
                *reinterpret_cast<uint32_t*>(dst + 8192) = ll;
            } else {
                *reinterpret_cast<float2*>(p.hbar + (size_t)sidx * 256 + c0 + 2 * j8) = make_float2(r2[0], r2[1]);
            }
        }
    }
    return apart;
}

// =====================================================================================================================
// v5: "TMEM role ping-pong", single N=256 pass per layer.  The two 256-column TMEM regions P and Q alternate between
// accumulator and A operand: the epilogue converts the finished accumulator IN PLACE, 16 columns at a time, into the
// packed bf16 operand of the next layer (8 columns hi | 8 columns lo), and signals each 16-column chunk on its own
// mbarrier; the issuer starts the next layer's k-step g (K = 16g..16g+15, accumulating into the OTHER region) as soon
// as chunk g is converted, so the tensor core runs layer l+1 right behind the epilogue of layer l.  The final
// epilogue of a tile runs under layer 1 of the next tile.
//      layer 1: A shared memory, acc Q      layer 2: A = Q, acc P      layer 3: A = P (+extras), acc Q      layer 4: A = Q, acc P
// Warp roles (448 threads): 0-7 epilogue (quadrant = w & 3, chunk group j = w >> 2 handles chunks j, j+2, ..., j+14),
// 8-11 builders, 12 loader, 13 issuer.  Few epilogue warps on purpose: the epilogue only has to stay ahead of the MMAs,
// and every busy warp on the issuer's sub-partition slows the single issuing thread.
namespace tc5 {
constexpr int NEPI_WARPS = 8, NGRP = NEPI_WARPS / 4, NCH = 16 / NGRP;   // chunk groups / chunks per thread
constexpr int NEPI = NEPI_WARPS * 32, NBUILD = 128, NTHR = NEPI + NBUILD + 64;
constexpr int NSTAGE = 4;
struct Smem {
    static constexpr int NWC = 3;          // weight*conf of tile t is read by the last epilogue under layer 1 of tile t+1, while the
                                           // builders already write tile t+2: three buffers make that ordering formal (through bar_drain)
    unsigned char a_hi[tc::NKB_MAX * tc::ABLK];
    unsigned char a_lo[tc::NKB_MAX * tc::ABLK];
    unsigned char b[NSTAGE][tc::IMG];
    unsigned char xe_hi[2][tc3::XE];
    unsigned char xe_lo[2][tc3::XE];
    float wc[NWC][tc::TM];
    float alpha_part[2][tc::TM];
    uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_a1_ready, bar_a1_free, bar_acc_full, bar_drain, bar_kblk[8];   // bar_kblk[kb]: columns 32kb..32kb+31 converted
    uint32_t tmem_base;
};
}  // namespace tc5

__global__ void __launch_bounds__(tc5::NTHR, 1) k_shade_tc5(ShadeTcParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    tc5::Smem& sm = *reinterpret_cast<tc5::Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const pnb_query_t& q = p.q;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], p.hbar_cap);
    const int n_tiles = (n_valid + TSAMP - 1) / TSAMP;
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    constexpr int W_BUILD = tc5::NEPI_WARPS, W_LOAD = W_BUILD + 4, W_ISSUE = W_LOAD + 1;

    if (tid == 0) {
        for (int s = 0; s < tc5::NSTAGE; ++s) { mbar_init(&sm.bar_full[s], 1); mbar_init(&sm.bar_empty[s], 1); }
        mbar_init(&sm.bar_a1_ready, tc5::NBUILD);
        mbar_init(&sm.bar_a1_free, 1);
        mbar_init(&sm.bar_acc_full, 1);
        mbar_init(&sm.bar_drain, tc5::NEPI);
        for (int c = 0; c < 8; ++c) mbar_init(&sm.bar_kblk[c], 32 * 4 * 2);   // 4 quadrant warps x 2 chunks of 16 columns
        mbar_fence_init();
        if (blockIdx.x == 0 && q.counters[PNB_QC_N_VALID] > p.hbar_cap) atomicExch(p.err, 9);
    }
    if (warp == W_ISSUE) tmem_alloc<512>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tP = sm.tmem_base, tQ = sm.tmem_base + 256u;
    const long long _tk0 = clock64();

    if (warp == W_LOAD) {
        // ============================================================ loader
        if (lane == 0) {
            const uint32_t total = (uint32_t)my_tiles * IMGS_PER_TILE;
            for (uint32_t n = 0; n < total; ++n) {
                const uint32_t s = n & (tc5::NSTAGE - 1), ph = (n >> 2) & 1u;
                if (!PNB_TIMED_WAIT(0, mbar_wait(&sm.bar_empty[s], ph ^ 1u, p.err, 41))) break;
                if (p.dbg_no_weights) { mbar_arrive(&sm.bar_full[s]); continue; }
                mbar_arrive_expect_tx(&sm.bar_full[s], IMG);
                bulk_g2s(sm.b[s], p.wimg + (size_t)(n % IMGS_PER_TILE) * IMG, IMG, &sm.bar_full[s]);
            }
        }
    } else if (warp == W_ISSUE) {
        // ============================================================ MMA issuer: the whole warp runs this loop
        // (warp-uniform control flow and operands); one elected lane issues each tcgen05 instruction
        {
            const uint32_t idesc = make_idesc_bf16(128, 256);
            const uint32_t hiw = desc_hi<LAYOUT>(), xe_hiw = (256u >> 4) | (1u << 14);
            const uint32_t b0_lo = desc_lo<LAYOUT>(smem_u32(sm.b[0]));
            const uint32_t ahi_lo = desc_lo<LAYOUT>(smem_u32(sm.a_hi)), alo_lo = desc_lo<LAYOUT>(smem_u32(sm.a_lo));
            const uint32_t xeh_lo0 = desc_lo<LAYOUT_NONE>(smem_u32(sm.xe_hi[0])), xel_lo0 = desc_lo<LAYOUT_NONE>(smem_u32(sm.xe_lo[0]));
            constexpr uint32_t KADV = kstep_adv16<LAYOUT>();
            uint32_t n = 0;            // weight image counter
            uint32_t c_acc = 0;        // completions of bar_acc_full consumed by this thread
            uint32_t c_pack = 0;       // packing rounds (one per layer 1..3 epilogue) consumed on bar_chunk[*]
            bool ok = true;
            for (int t = 0; t < my_tiles && ok; ++t) {
                const uint32_t xeh_lo = xeh_lo0 + (uint32_t)(t & 1) * (tc3::XE >> 4), xel_lo = xel_lo0 + (uint32_t)(t & 1) * (tc3::XE >> 4);
                for (int l = 0; l < 4 && ok; ++l) {
                    const uint32_t acc = (l & 1) ? tP : tQ;           // accumulator of this layer
                    const uint32_t ab = (l & 1) ? tQ : tP;            // packed A operand of this layer (l >= 1)
                    // all MMAs of the previous layer (global order) must be complete before their A region becomes this accumulator
                    if (t > 0 || l > 0) { if (!PNB_TIMED_WAIT_L0(2, mbar_wait(&sm.bar_acc_full, c_acc & 1u, p.err, 42))) { ok = false; break; } ++c_acc; }
                    if (l == 0) { if (!PNB_TIMED_WAIT_L0(1, mbar_wait(&sm.bar_a1_ready, (uint32_t)t & 1u, p.err, 43))) { ok = false; break; } }
                    if (l == 1 && t > 0) { if (!PNB_TIMED_WAIT_L0(2, mbar_wait(&sm.bar_drain, (uint32_t)(t - 1) & 1u, p.err, 44))) { ok = false; break; } }
                    tc_fence_after();
                    const int nkb = nkb_of(l);
                    for (int kb = 0; kb < nkb && ok; ++kb) {
                        const uint32_t s0 = n & (tc5::NSTAGE - 1), ph0 = (n >> 2) & 1u;              // W_hi image
                        const uint32_t s1 = (n + 1) & (tc5::NSTAGE - 1), ph1 = ((n + 1) >> 2) & 1u;   // W_lo image
                        const bool need_chunks = (l >= 1 && kb < 8);  // chunks 2kb, 2kb+1 of the previous layer's output
                        uint64_t* cb0 = need_chunks ? &sm.bar_kblk[kb] : &sm.bar_full[s0];
                        uint64_t* cb1 = &sm.bar_full[s1];
                        const uint32_t cp0 = need_chunks ? (c_pack & 1u) : ph0, cp1 = ph1;
                        // fast path: one overlapped probe of everything this K block needs; slow path: bounded blocking waits
                        if (!mbar_try_wait4(&sm.bar_full[s0], ph0, &sm.bar_full[s1], ph1, cb0, cp0, cb1, cp1)) {
                            if (need_chunks) {
                                if (!PNB_TIMED_WAIT_L0(2, mbar_wait(cb0, cp0, p.err, 45))) { ok = false; break; }
                            }
                            if (!PNB_TIMED_WAIT_L0(3, mbar_wait(&sm.bar_full[s0], ph0, p.err, 46))) { ok = false; break; }
                            if (!PNB_TIMED_WAIT_L0(3, mbar_wait(&sm.bar_full[s1], ph1, p.err, 46))) { ok = false; break; }
                        }
                        tc_fence_after();
                        const uint32_t akb_hi = ahi_lo + (uint32_t)kb * (ABLK >> 4), akb_lo = alo_lo + (uint32_t)kb * (ABLK >> 4);
                        const uint32_t tcol = ab + (uint32_t)(kb * 32);
                        const uint32_t bl = b0_lo + s0 * (IMG >> 4), bl2 = b0_lo + s1 * (IMG >> 4);
                        if (l == 0) {
                            mma_ss2_w(acc, akb_hi, hiw, bl, hiw, idesc, kb ? 1u : 0u);
                            mma_ss2_w(acc, akb_lo, hiw, bl, hiw, idesc, 1u);
                            mma_ss2_w(acc, akb_hi + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                            mma_ss2_w(acc, akb_lo + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                            mma_commit_w(&sm.bar_empty[s0]);
                            mma_ss2_w(acc, akb_hi, hiw, bl2, hiw, idesc, 1u);
                            mma_ss2_w(acc, akb_hi + KADV, hiw, bl2 + KADV, hiw, idesc, 1u);
                            mma_commit_w(&sm.bar_empty[s1]);
                        } else if (kb == 8) {
                            mma_ss2_w(acc, xeh_lo, xe_hiw, bl, hiw, idesc, 1u);
                            mma_ss2_w(acc, xel_lo, xe_hiw, bl, hiw, idesc, 1u);
                            mma_commit_w(&sm.bar_empty[s0]);
                            mma_ss2_w(acc, xeh_lo, xe_hiw, bl2, hiw, idesc, 1u);
                            mma_commit_w(&sm.bar_empty[s1]);
                        } else {
                            mma_ts2_w(acc, tcol, bl, hiw, idesc, kb ? 1u : 0u);
                            mma_ts2_w(acc, tcol + 8u, bl, hiw, idesc, 1u);
                            mma_ts2_w(acc, tcol + 16u, bl + KADV, hiw, idesc, 1u);
                            mma_ts2_w(acc, tcol + 24u, bl + KADV, hiw, idesc, 1u);
                            mma_commit_w(&sm.bar_empty[s0]);
                            mma_ts2_w(acc, tcol, bl2, hiw, idesc, 1u);
                            mma_ts2_w(acc, tcol + 16u, bl2 + KADV, hiw, idesc, 1u);
                            mma_commit_w(&sm.bar_empty[s1]);
                        }
                        n += 2;
                    }
                    if (!ok) break;
                    if (l >= 1) ++c_pack;
                    mma_commit_w(&sm.bar_acc_full);
                    if (l == 0) mma_commit_w(&sm.bar_a1_free);
                }
            }
        }
    } else if (warp >= W_BUILD) {
        // ============================================================ builders: one thread per pair row
        const int row = (warp - W_BUILD) * 32 + lane;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            if (t > 0 && !(lane == 0 && warp == W_BUILD ? PNB_TIMED_WAIT(4, mbar_wait(&sm.bar_a1_free, (uint32_t)(t - 1) & 1u, p.err, 47)) : mbar_wait(&sm.bar_a1_free, (uint32_t)(t - 1) & 1u, p.err, 47))) { ok = false; break; }
            const long long _tb0 = clock64();
            build_pair_row(sm, p, tile, t, row, n_valid);
            fence_proxy_async();
            mbar_arrive(&sm.bar_a1_ready);
            if (lane == 0 && warp == W_BUILD) prof_add(p, 5, clock64() - _tb0);
        }
    } else {
        // ============================================================ epilogue warps
        const int quad = warp & 3, grp = warp >> 2;            // chunk group: chunks grp, grp+NGRP, grp+2*NGRP, ...
        const int erow = quad * 32 + lane;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        uint32_t n_acc = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            for (int l = 0; l < 4 && ok; ++l, ++n_acc) {
                if (!(tid == 0 ? PNB_TIMED_WAIT(6, mbar_wait(&sm.bar_acc_full, n_acc & 1u, p.err, 48)) : mbar_wait(&sm.bar_acc_full, n_acc & 1u, p.err, 48))) { ok = false; break; }
                const long long _te0 = clock64();
                tc_fence_after();
                const uint32_t accb = ((l & 1) ? tP : tQ) + tlane;
                if (l < 3) {
                    const float* bias = p.bias[l];
#pragma unroll
                    for (int i = 0; i < tc5::NCH; ++i) {
                        const int g = grp + tc5::NGRP * i, c0 = 16 * g;
                        uint32_t v[16];
                        tmem_ld16(accb + (uint32_t)c0, v);
                        tmem_ld_wait();
                        uint32_t hh[8], ll[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c0) + e);
                            float y0 = __uint_as_float(v[2 * e]) + bb.x, y1 = __uint_as_float(v[2 * e + 1]) + bb.y;
                            y0 = fmaxf(y0, LEAKY * y0); y1 = fmaxf(y1, LEAKY * y1);
                            split_bf16x2(y0, y1, hh[e], ll[e]);
                        }
                        tmem_st8(accb + (uint32_t)c0, hh);               // in place: 8 columns hi | 8 columns lo
                        tmem_st8(accb + (uint32_t)c0 + 8u, ll);
                        tmem_st_wait();
                        tc_fence_before();
                        mbar_arrive(&sm.bar_kblk[g >> 1]);
                    }
                    if (tid == 0) prof_add(p, 7, clock64() - _te0);
                } else {
                    const float wrow = sm.wc[t % tc5::Smem::NWC][erow];
                    const int sidx = tile * TSAMP + (erow >> 3);
                    const bool swrite = sidx < n_valid;
                    const int j8 = lane & 7;
                    const float apart = last_chunks<tc5::NGRP, tc5::NCH>(p, accb, grp, wrow, sidx, swrite, lane);
                    tc_fence_before();
                    mbar_arrive(&sm.bar_drain);                // accumulator region P drained
                    if (tid == 0) prof_add(p, 8, clock64() - _te0);
                    if (grp < 2) sm.alpha_part[grp][erow] = apart;
                    named_bar_sync(1, tc5::NEPI);
                    if (grp >= 2) atomicAdd(&sm.alpha_part[grp - 2][erow], apart);
                    if (tc5::NGRP > 2) named_bar_sync(1, tc5::NEPI);
                    if (grp == 0) {
                        float a = sm.alpha_part[0][erow] + (tc5::NGRP > 1 ? sm.alpha_part[1][erow] : 0.f) + __ldg(p.ba) - 1.0f;
                        float sp = a > 20.f ? a : log1pf(expf(a));
                        float zz = sp * wrow;
                        zz += __shfl_xor_sync(0xffffffffu, zz, 1);
                        zz += __shfl_xor_sync(0xffffffffu, zz, 2);
                        zz += __shfl_xor_sync(0xffffffffu, zz, 4);
                        if (j8 == 0 && swrite) p.sigma[sidx] = zz;
                    }
                    named_bar_sync(1, tc5::NEPI);
                }
            }
        }
    }
    if (tid == 0) prof_add(p, 9, clock64() - _tk0);
    if (tid == 0 && (p.dbg_flags & 4) && blockIdx.x < 192) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        reinterpret_cast<long long*>(p.err)[32 + blockIdx.x] = ((clock64() - _tk0) & 0xffffffffffffll) | ((long long)smid << 48);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ISSUE) tmem_dealloc<512>(sm.tmem_base);
}

// =====================================================================================================================
// v7: v5 with PACKED ROWS.  v5 gives every valid sample 8 rows (one per neighbour slot) although only 75 % of the slots
// hold a neighbour on the lego frame (P_v / (8 S_v)): a quarter of every MMA multiplies zero rows.  Here a 128-row tile
// is four 32-row quadrants (= the TMEM lane quarter of one warp), and each quadrant holds whole samples packed back to
// back, only their valid neighbours (first-fit packing with a 64-sample look-ahead, k_pack_*: the packed order is the
// permutation `vorder`): 99.3 % of the rows carry a pair on the lego frame.
// The K-reduction over the rows of a sample (1..8 consecutive lanes, never crossing a quadrant) is a segmented
// warp-shuffle scan; its addition order depends on the neighbour count only, so a ray's colour is still independent of
// which rays share the call.  Two builder threads per row and the last epilogue shared between the epilogue and the
// builder warps (as v6); TMEM role ping-pong, chunk hand-off, weight ring and issuer are v5's.
namespace tc7 {
constexpr int NEPI_WARPS = 8, NGRP = NEPI_WARPS / 4, NCH = 16 / NGRP;
constexpr int NEPI = NEPI_WARPS * 32, NBUILD = 256, NTHR = NEPI + NBUILD + 64;     // two builder threads per row (as v6)
constexpr int NSTAGE = 4;
constexpr int PACK_S = 512;             // samples per independently packed super-chunk (its last quadrant may stay partly empty)
constexpr int PACK_WIN = 64;            // look-ahead of the first-fit packing
struct Smem {
    static constexpr int NWC = 2;          // reuse ordered through bar_alpha: the builders run their share of the last epilogue
    unsigned char a_hi[tc::NKB_MAX * tc::ABLK];
    unsigned char a_lo[tc::NKB_MAX * tc::ABLK];
    unsigned char b[NSTAGE][tc::IMG];
    unsigned char xe_hi[2][tc3::XE];
    unsigned char xe_lo[2][tc3::XE];
    float wc[NWC][tc::TM];
    float alpha_part[2][tc::TM];         // builder groups' partial alpha dot products
    float alpha_e[tc::TM];               // sum of the two epilogue groups' partials (two addends onto 0: order-independent)
    uint32_t qhead[NWC][4], qfirst[NWC][4], qtotal[NWC][4];   // per quadrant: bit r = row r starts a sample; first valid-sample index; rows used
    uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_a1_ready, bar_a1_free, bar_acc_full, bar_final, bar_alpha, bar_drain, bar_kblk[8];
    uint32_t tmem_base;
};
}  // namespace tc7

// ---- row packing (runs before k_shade_tc7; all sizes come from device counters)
__global__ void __launch_bounds__(256) k_pack_cnt(pnb_query_t q, int cap, unsigned char* __restrict__ vcnt) {
    const int vi = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], cap);
    if (vi < n_valid) vcnt[vi] = q.samp_nvalid[q.valid_list[vi]];
}
// one thread per super-chunk: first-fit packing of its samples into 32-row quadrants.  Samples are taken in order while they fit;
// a sample that does not fit stays first in line for the next quadrant while up to PACK_WIN later, smaller samples may fill the
// remaining rows (so the order inside a super-chunk becomes a permutation: vorder).  WRITE = false: count the quadrants only.
template <bool WRITE>
__global__ void __launch_bounds__(128) k_pack_quads(pnb_query_t q, int cap, const unsigned char* __restrict__ vcnt,
                                                    uint32_t* __restrict__ sc_quads, uint32_t* __restrict__ quad_first,
                                                    uint32_t* __restrict__ vorder, unsigned char* __restrict__ vcntp) {
    const int sc = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], cap);
    const int i0 = sc * tc7::PACK_S;
    if (i0 >= n_valid) return;
    const int n = min(tc7::PACK_S, n_valid - i0);
    __align__(16) unsigned char c[tc7::PACK_S];   // neighbour counts of this super-chunk; 0 = already placed
    for (int i = 0; i < n; i += 16) {
        const uint4 pk = *reinterpret_cast<const uint4*>(vcnt + i0 + i);      // the buffer is padded to a multiple of 16
        *reinterpret_cast<uint4*>(c + i) = pk;
    }
    uint32_t nq = WRITE ? sc_quads[sc] : 0u;      // WRITE: sc_quads holds the exclusive prefix = first quadrant of this super-chunk
    int pos = 0, emitted = 0;
    while (pos < n) {
        if (WRITE) quad_first[nq] = (uint32_t)(i0 + emitted);
        int rows = 0;
        const int lim = min(n, pos + tc7::PACK_WIN);
        for (int i = pos; i < lim && rows < 32; ++i) {
            const int ci = c[i];
            if (ci != 0 && rows + ci <= 32) {
                if (WRITE) { vorder[i0 + emitted] = (uint32_t)(i0 + i); vcntp[i0 + emitted] = (unsigned char)ci; }
                c[i] = 0;
                rows += ci;
                ++emitted;
            }
        }
        while (pos < n && c[pos] == 0) ++pos;
        ++nq;
    }
    if (!WRITE) sc_quads[sc] = nq;
}
// exclusive scan of the per-super-chunk quadrant counts (one block), total -> pack_cnt[0], sentinel quad_first[n_quads] = n_valid
__global__ void __launch_bounds__(1024) k_pack_scan(pnb_query_t q, int cap, uint32_t* __restrict__ sc_quads, uint32_t* __restrict__ quad_first,
                                                    int* __restrict__ pack_cnt) {
    __shared__ uint32_t part[1024];
    const int n_valid = min(q.counters[PNB_QC_N_VALID], cap);
    const int n_sc = (n_valid + tc7::PACK_S - 1) / tc7::PACK_S;
    const int per = (n_sc + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(b0 + per, n_sc);
    uint32_t sum = 0;
    for (int i = b0; i < b1; ++i) sum += sc_quads[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; ++i) { const uint32_t v = part[i]; part[i] = run; run += v; }
        pack_cnt[0] = (int)run;
        quad_first[run] = (uint32_t)n_valid;
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (int i = b0; i < b1; ++i) { const uint32_t v = sc_quads[i]; sc_quads[i] = run; run += v; }
}

// Last epilogue with packed rows, one warp's share (chunks G, G+NG, ...): +bias, LeakyReLU, partial alpha dot product (returned),
// weight*conf scaling, then the K-reduction over the rows of each sample as a segmented inclusive scan (segments = samples,
// <= 8 lanes, first lane st); the last row of a sample (swrite) holds the sums and writes h-bar.
template <int NG, int NCHUNK>
__device__ __forceinline__ float last_chunks_packed(const ShadeTcParams& p, uint32_t accb, int G, float wrow, int st, bool swrite, int sidx, int lane) {
    using namespace tc;
    const float* bias = p.bias[3];
    float apart = 0.f;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
        const int c0 = 16 * (G + NG * i);
        uint32_t v[16];
        tmem_ld16(accb + (uint32_t)c0, v);
        tmem_ld_wait();
        float z[16];
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0) + e4), ww = __ldg(reinterpret_cast<const float4*>(p.wa + c0) + e4);
            const float bq[4] = {bb.x, bb.y, bb.z, bb.w}, wq[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
            for (int e1 = 0; e1 < 4; ++e1) {
                const int e = 4 * e4 + e1;
                float y = __uint_as_float(v[e]) + bq[e1];
                y = fmaxf(y, LEAKY * y);
                apart = fmaf(y, wq[e1], apart);
                z[e] = y * wrow;
            }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) z[e] = seg_scan8(z[e], lane, st);
        if (swrite) {
            if (p.hbar_fmt) {       // the colour kernel's operand image (bf16 hi / lo, core-matrix layout): two 16-byte rows each
                uint32_t hh[8], ll[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_bf16x2(z[2 * e], z[2 * e + 1], hh[e], ll[e]);
                unsigned char* dst = reinterpret_cast<unsigned char*>(p.hbar) + ((size_t)(sidx >> 7) * 8 + (c0 >> 5)) * (2 * 8192) +
                                     tile_offset_bytes<LAYOUT_NONE>(sidx & 127, c0 & 31);
                *reinterpret_cast<uint4*>(dst) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                *reinterpret_cast<uint4*>(dst + 128) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
                *reinterpret_cast<uint4*>(dst + 8192) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                *reinterpret_cast<uint4*>(dst + 8192 + 128) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
            } else {
                float4* dst = reinterpret_cast<float4*>(p.hbar + (size_t)sidx * 256 + c0);
                dst[0] = make_float4(z[0], z[1], z[2], z[3]);
                dst[1] = make_float4(z[4], z[5], z[6], z[7]);
                dst[2] = make_float4(z[8], z[9], z[10], z[11]);
                dst[3] = make_float4(z[12], z[13], z[14], z[15]);
            }
        }
    }
    return apart;
}
// segment bookkeeping of one quadrant row: first row / index of its sample, whether it is the sample's last row
struct QuadRow { int st, j; bool live, is_end; };
__device__ __forceinline__ QuadRow quad_row(uint32_t head, int tot, int lane) {
    QuadRow r;
    r.live = lane < tot;
    const uint32_t below = head & (0xffffffffu >> (31 - lane));
    r.st = r.live ? 31 - __clz(below) : lane;
    r.j = r.live ? __popc(below) - 1 : 0;
    const uint32_t nxt = lane < 31 ? (head >> (lane + 1)) : 0u;
    r.is_end = r.live && lane == (nxt ? lane + __ffs(nxt) - 1 : tot - 1);
    return r;
}

__global__ void __launch_bounds__(tc7::NTHR, 1) k_shade_tc7(ShadeTcParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    tc7::Smem& sm = *reinterpret_cast<tc7::Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const pnb_query_t& q = p.q;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], p.hbar_cap);
    const int n_quads = p.pack_cnt[0];
    const int n_tiles = (n_quads + 3) >> 2;
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    constexpr int W_BUILD = tc7::NEPI_WARPS, W_LOAD = W_BUILD + tc7::NBUILD / 32, W_ISSUE = W_LOAD + 1;

    if (tid == 0) {
        for (int s = 0; s < tc7::NSTAGE; ++s) { mbar_init(&sm.bar_full[s], 1); mbar_init(&sm.bar_empty[s], 1); }
        mbar_init(&sm.bar_a1_ready, tc7::NBUILD);
        mbar_init(&sm.bar_a1_free, 1);
        mbar_init(&sm.bar_acc_full, 1);
        mbar_init(&sm.bar_final, 1);
        mbar_init(&sm.bar_alpha, tc7::NEPI_WARPS);
        mbar_init(&sm.bar_drain, tc7::NEPI_WARPS + tc7::NBUILD / 32);     // one arrive per warp that reads the layer-4 accumulator
        for (int c = 0; c < 8; ++c) mbar_init(&sm.bar_kblk[c], 32 * 4 * 2);
        mbar_fence_init();
        if (blockIdx.x == 0 && q.counters[PNB_QC_N_VALID] > p.hbar_cap) atomicExch(p.err, 9);
    }
    if (tid < TM) sm.alpha_e[tid] = 0.f;
    if (warp == W_ISSUE) tmem_alloc<512>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tP = sm.tmem_base, tQ = sm.tmem_base + 256u;
    const long long _tk0 = clock64();

    if (warp == W_LOAD) {
        if (lane == 0) {
            const uint32_t total = (uint32_t)my_tiles * IMGS_PER_TILE;
            for (uint32_t n = 0; n < total; ++n) {
                const uint32_t s = n & (tc7::NSTAGE - 1), ph = (n >> 2) & 1u;
                if (!mbar_wait(&sm.bar_empty[s], ph ^ 1u, p.err, 91)) break;
                if (p.dbg_no_weights) { mbar_arrive(&sm.bar_full[s]); continue; }
                mbar_arrive_expect_tx(&sm.bar_full[s], IMG);
                bulk_g2s(sm.b[s], p.wimg + (size_t)(n % IMGS_PER_TILE) * IMG, IMG, &sm.bar_full[s]);
            }
        }
    } else if (warp == W_ISSUE) {
        // ============================================================ MMA issuer (identical to v5)
        const uint32_t idesc = make_idesc_bf16(128, 256);
        const uint32_t hiw = desc_hi<LAYOUT>(), xe_hiw = (256u >> 4) | (1u << 14);
        const uint32_t b0_lo = desc_lo<LAYOUT>(smem_u32(sm.b[0]));
        const uint32_t ahi_lo = desc_lo<LAYOUT>(smem_u32(sm.a_hi)), alo_lo = desc_lo<LAYOUT>(smem_u32(sm.a_lo));
        const uint32_t xeh_lo0 = desc_lo<LAYOUT_NONE>(smem_u32(sm.xe_hi[0])), xel_lo0 = desc_lo<LAYOUT_NONE>(smem_u32(sm.xe_lo[0]));
        constexpr uint32_t KADV = kstep_adv16<LAYOUT>();
        uint32_t n = 0, c_acc = 0, c_pack = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const uint32_t xeh_lo = xeh_lo0 + (uint32_t)(t & 1) * (tc3::XE >> 4), xel_lo = xel_lo0 + (uint32_t)(t & 1) * (tc3::XE >> 4);
            for (int l = 0; l < 4 && ok; ++l) {
                const uint32_t acc = (l & 1) ? tP : tQ;
                const uint32_t ab = (l & 1) ? tQ : tP;
                if (l > 0) { if (!mbar_wait(&sm.bar_acc_full, c_acc & 1u, p.err, 92)) { ok = false; break; } ++c_acc; }
                else if (t > 0) { if (!mbar_wait(&sm.bar_final, (uint32_t)(t - 1) & 1u, p.err, 92)) { ok = false; break; } }
                if (l == 0) { if (!mbar_wait(&sm.bar_a1_ready, (uint32_t)t & 1u, p.err, 93)) { ok = false; break; } }
                if (l == 1 && t > 0) { if (!mbar_wait(&sm.bar_drain, (uint32_t)(t - 1) & 1u, p.err, 94)) { ok = false; break; } }
                tc_fence_after();
                const int nkb = nkb_of(l);
                for (int kb = 0; kb < nkb && ok; ++kb) {
                    const uint32_t s0 = n & (tc7::NSTAGE - 1), ph0 = (n >> 2) & 1u;
                    const uint32_t s1 = (n + 1) & (tc7::NSTAGE - 1), ph1 = ((n + 1) >> 2) & 1u;
                    const bool need_chunks = (l >= 1 && kb < 8);
                    uint64_t* cb0 = need_chunks ? &sm.bar_kblk[kb] : &sm.bar_full[s0];
                    const uint32_t cp0 = need_chunks ? (c_pack & 1u) : ph0;
                    if (!mbar_try_wait4(&sm.bar_full[s0], ph0, &sm.bar_full[s1], ph1, cb0, cp0, &sm.bar_full[s1], ph1)) {
                        if (need_chunks && !mbar_wait(cb0, cp0, p.err, 95)) { ok = false; break; }
                        if (!mbar_wait(&sm.bar_full[s0], ph0, p.err, 96)) { ok = false; break; }
                        if (!mbar_wait(&sm.bar_full[s1], ph1, p.err, 96)) { ok = false; break; }
                    }
                    tc_fence_after();
                    const uint32_t akb_hi = ahi_lo + (uint32_t)kb * (ABLK >> 4), akb_lo = alo_lo + (uint32_t)kb * (ABLK >> 4);
                    const uint32_t tcol = ab + (uint32_t)(kb * 32);
                    const uint32_t bl = b0_lo + s0 * (IMG >> 4), bl2 = b0_lo + s1 * (IMG >> 4);
                    if (l == 0) {
                        mma_ss2_w(acc, akb_hi, hiw, bl, hiw, idesc, kb ? 1u : 0u);
                        mma_ss2_w(acc, akb_lo, hiw, bl, hiw, idesc, 1u);
                        mma_ss2_w(acc, akb_hi + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                        mma_ss2_w(acc, akb_lo + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                        mma_commit_w(&sm.bar_empty[s0]);
                        mma_ss2_w(acc, akb_hi, hiw, bl2, hiw, idesc, 1u);
                        mma_ss2_w(acc, akb_hi + KADV, hiw, bl2 + KADV, hiw, idesc, 1u);
                        mma_commit_w(&sm.bar_empty[s1]);
                    } else if (kb == 8) {
                        mma_ss2_w(acc, xeh_lo, xe_hiw, bl, hiw, idesc, 1u);
                        mma_ss2_w(acc, xel_lo, xe_hiw, bl, hiw, idesc, 1u);
                        mma_commit_w(&sm.bar_empty[s0]);
                        mma_ss2_w(acc, xeh_lo, xe_hiw, bl2, hiw, idesc, 1u);
                        mma_commit_w(&sm.bar_empty[s1]);
                    } else {
                        mma_ts2_w(acc, tcol, bl, hiw, idesc, kb ? 1u : 0u);
                        mma_ts2_w(acc, tcol + 8u, bl, hiw, idesc, 1u);
                        mma_ts2_w(acc, tcol + 16u, bl + KADV, hiw, idesc, 1u);
                        mma_ts2_w(acc, tcol + 24u, bl + KADV, hiw, idesc, 1u);
                        mma_commit_w(&sm.bar_empty[s0]);
                        mma_ts2_w(acc, tcol, bl2, hiw, idesc, 1u);
                        mma_ts2_w(acc, tcol + 16u, bl2 + KADV, hiw, idesc, 1u);
                        mma_commit_w(&sm.bar_empty[s1]);
                    }
                    n += 2;
                }
                if (!ok) break;
                if (l >= 1) ++c_pack;
                mma_commit_w(l < 3 ? &sm.bar_acc_full : &sm.bar_final);      // layers 1-3 -> epilogue warps; layer 4 -> every warp's share of the last epilogue
                if (l == 0) mma_commit_w(&sm.bar_a1_free);
            }
        }
    } else if (warp >= W_BUILD) {
        // ============================================================ builders: two warps per quadrant (operand columns 0..151 | 152..287),
        // lane = row; the same warps run chunk groups 2, 3 of the last epilogue of the previous tile (they are idle under layer 1)
        const int bw = warp - W_BUILD, qw = bw & 3, part = bw >> 2, row = qw * 32 + lane;
        const uint32_t tlane = (uint32_t)(qw * 32) << 16;
        bool ok = true;
        for (int t = 0; t <= my_tiles && ok; ++t) {
            if (t < my_tiles) {
                const int tile = (int)blockIdx.x + t * (int)gridDim.x;
                if (t > 0 && !mbar_wait(&sm.bar_a1_free, (uint32_t)(t - 1) & 1u, p.err, 97)) { ok = false; break; }
                const int qd = tile * 4 + qw;
                uint32_t first = 0, nsamp = 0;
                if (qd < n_quads) { first = p.quad_first[qd]; nsamp = p.quad_first[qd + 1] - first; }
                const int c = lane < (int)nsamp ? (int)p.vcnt[first + lane] : 0;
                int incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
                const uint32_t head = __reduce_or_sync(0xffffffffu, lane < (int)nsamp ? (1u << (incl - c)) : 0u);
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                const QuadRow qr = quad_row(head, total, lane);
                const int cj = __shfl_sync(0xffffffffu, c, qr.j);
                if (lane == 0 && part == 1) { sm.qhead[t & 1][qw] = head; sm.qfirst[t & 1][qw] = first; sm.qtotal[t & 1][qw] = (uint32_t)total; }
                const int pvi = qr.live ? (int)p.vorder[first + qr.j] : -1;
                if (part == 0) build_pair_part<0, true>(sm, p, tile, t, row, n_valid, pvi, lane - qr.st, qr.st, qr.live ? cj : 1);
                else build_pair_part<1, true>(sm, p, tile, t, row, n_valid, pvi, lane - qr.st, qr.st, qr.live ? cj : 1);
                fence_proxy_async();
                mbar_arrive(&sm.bar_a1_ready);
            }
            if (t > 0) {
                const int tf = t - 1;
                if (!mbar_wait(&sm.bar_final, (uint32_t)tf & 1u, p.err, 99)) { ok = false; break; }
                tc_fence_after();
                const QuadRow qr = quad_row(sm.qhead[tf & 1][qw], (int)sm.qtotal[tf & 1][qw], lane);
                const int sidx = qr.live ? (int)p.vorder[sm.qfirst[tf & 1][qw] + qr.j] : 0;
                const bool swrite = qr.is_end && sidx < n_valid;
                const float wrow = sm.wc[tf & 1][row];
                const float apart = last_chunks_packed<4, 4>(p, tP + tlane, 2 + part, wrow, qr.st, swrite, sidx, lane);
                tc_fence_before();
                sm.alpha_part[part][row] = apart;
                named_bar_sync(2, tc7::NBUILD);
                if (!mbar_wait(&sm.bar_alpha, (uint32_t)tf & 1u, p.err, 100)) { ok = false; break; }      // the epilogue warps' partial sums
                if (part == 0) {
                    const float a = (sm.alpha_part[0][row] + sm.alpha_part[1][row]) + sm.alpha_e[row] + __ldg(p.ba) - 1.0f;
                    sm.alpha_e[row] = 0.f;
                    const float sp = a > 20.f ? a : log1pf(expf(a));
                    const float zz = seg_scan8(sp * wrow, lane, qr.st);
                    if (swrite) p.sigma[sidx] = zz;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.bar_drain);                 // after the alpha_e reads
                named_bar_sync(2, tc7::NBUILD);
            }
        }
    } else {
        // ============================================================ epilogue warps
        const int quad = warp & 3, grp = warp >> 2;
        const int erow = quad * 32 + lane;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        uint32_t n_acc = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            for (int l = 0; l < 3 && ok; ++l, ++n_acc) {        // the layer-4 (last) epilogue is shared with the builder warps
                if (!mbar_wait(&sm.bar_acc_full, n_acc & 1u, p.err, 98)) { ok = false; break; }
                tc_fence_after();
                const uint32_t accb = ((l & 1) ? tP : tQ) + tlane;
                {
                    const float* bias = p.bias[l];
#pragma unroll
                    for (int i = 0; i < tc7::NCH; ++i) {
                        const int g = grp + tc7::NGRP * i, c0 = 16 * g;
                        uint32_t v[16];
                        tmem_ld16(accb + (uint32_t)c0, v);
                        tmem_ld_wait();
                        uint32_t hh[8], ll[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c0) + e);
                            float y0 = __uint_as_float(v[2 * e]) + bb.x, y1 = __uint_as_float(v[2 * e + 1]) + bb.y;
                            y0 = fmaxf(y0, LEAKY * y0); y1 = fmaxf(y1, LEAKY * y1);
                            split_bf16x2(y0, y1, hh[e], ll[e]);
                        }
                        tmem_st8(accb + (uint32_t)c0, hh);
                        tmem_st8(accb + (uint32_t)c0 + 8u, ll);
                        tmem_st_wait();
                        tc_fence_before();
                        mbar_arrive(&sm.bar_kblk[g >> 1]);
                    }
                }
            }
            if (!ok) break;
            {   // this warp's share of the LAST epilogue (chunk groups 0, 1; the builder warps take 2, 3)
                if (!mbar_wait(&sm.bar_final, (uint32_t)t & 1u, p.err, 101)) { ok = false; break; }
                tc_fence_after();
                const QuadRow qr = quad_row(sm.qhead[t & 1][quad], (int)sm.qtotal[t & 1][quad], lane);
                const int sidx = qr.live ? (int)p.vorder[sm.qfirst[t & 1][quad] + qr.j] : 0;
                const float apart = last_chunks_packed<4, 4>(p, tP + tlane, grp, sm.wc[t & 1][erow], qr.st, qr.is_end && sidx < n_valid, sidx, lane);
                tc_fence_before();
                atomicAdd(&sm.alpha_e[erow], apart);
                __syncwarp();
                if (lane == 0) { mbar_arrive(&sm.bar_alpha); mbar_arrive(&sm.bar_drain); }
            }
        }
    }
    if (tid == 0 && (p.dbg_flags & 4) && blockIdx.x < 192) {      // per-CTA cycles and SM id; [32 + 192] = number of quadrants
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        reinterpret_cast<long long*>(p.err)[32 + blockIdx.x] = ((clock64() - _tk0) & 0xffffffffffffll) | ((long long)smid << 48);
        if (blockIdx.x == 0) reinterpret_cast<long long*>(p.err)[32 + 192] = n_quads;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ISSUE) tmem_dealloc<512>(sm.tmem_base);
}

// =====================================================================================================================
// v6 (opt-in, pnb_tc_version=6): v5 on a CTA PAIR (cluster of 2, tcgen05 cta_group::2).  Each CTA of the pair owns one 128-row
// tile (its own builders, epilogue warps, TMEM regions P/Q and layer-1 operand buffer); the rank-0 CTA's issuer warp issues
// ONE M=256 MMA for both tiles.  The B operand (weight image [256 x 32]) is split by N between the two CTAs: each loader
// streams only its 8 KB half of every image (half the L2->SMEM weight traffic), a ring stage is a whole K block (hi + lo half
// images, 16 KB) so ONE tcgen05.commit per K block frees it (a commit costs ~140 tensor-pipe cycles), and the tensor cores
// read the other half of B from the peer's shared memory.
// Cross-CTA signalling: commits are multicast to the barrier of both CTAs (ring "empty", accumulator-full, operand-free);
// the peer's builders / epilogue warps arrive remotely on the leader's a1_ready / kblk / drain barriers (default .release.cta
// semantics: an explicit .release.cluster compiles to MEMBAR.ALL.GPU per arrive); the peer's "weight half landed" is
// forwarded to the leader's full barrier (count 2) by a forwarder thread.  Two builder threads per row; the last epilogue
// of a tile is shared by the epilogue and the builder warps (both idle under layer 1 of the next tile).
// Measured (B200, lego frame): a pair alone runs 34.5 k cycles per tile against 38.4 k for v5, but every MMA pulls 4 KB of B
// from the peer SM, and with all 74 pairs active that exchange saturates the intra-GPC SM-to-SM fabric (~20 B/clk/SM): TPCs
// settle at 34.5 k / 38 k / 41.5 k cycles per tile depending on their position in the GPC, and the statically partitioned
// kernel is as slow as its slowest pair -> 3 % slower than v5 end to end.  Kept for the numbers and as the base of a
// future mixed / dynamically scheduled variant.
namespace tc6 {
constexpr int NEPI_WARPS = 8, NGRP = NEPI_WARPS / 4, NCH = 16 / NGRP;
constexpr int NEPI = NEPI_WARPS * 32, NBUILD = 256, NTHR = NEPI + NBUILD + 96;   // two builder threads per row; + loader, issuer, forwarder warps
constexpr int NSTAGE = 4;                // ring stage = one K block: this CTA's half of the W_hi image and of the W_lo image
constexpr int CPK = 1;                   // K blocks per ring commit (a tcgen05.commit costs ~140 tensor-pipe cycles; CPK = 2 frees stages in pairs
                                         // and was measured slower: the 4-stage ring then starves)
constexpr int HIMG = tc::IMG / 2;       // bytes of one half image ([128 x 32] bf16)
struct Smem {
    static constexpr int NWC = 2;          // reuse ordered through bar_alpha (the builders themselves run the last epilogue)
    unsigned char a_hi[tc::NKB_MAX * tc::ABLK];
    unsigned char a_lo[tc::NKB_MAX * tc::ABLK];
    unsigned char b[NSTAGE][2][HIMG];
    unsigned char xe_hi[2][tc3::XE];
    unsigned char xe_lo[2][tc3::XE];
    float wc[2][tc::TM];
    float alpha_part[2][tc::TM];         // builder groups' partial alpha dot products
    float alpha_e[tc::TM];               // sum of the two epilogue groups' partials (atomicAdd of two addends onto 0: order-independent);
                                         // read and re-zeroed by the builders, reuse ordered through bar_drain
    uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_a1_ready, bar_a1_free, bar_acc_full, bar_final, bar_alpha, bar_drain, bar_kblk[8];
    uint32_t tmem_base;
};
}  // namespace tc6

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc6::NTHR, 1) k_shade_tc6(ShadeTcParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    tc6::Smem& sm = *reinterpret_cast<tc6::Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const pnb_query_t& q = p.q;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], p.hbar_cap);
    const int n_tiles = (n_valid + TSAMP - 1) / TSAMP;
    const int n_ptiles = (n_tiles + 1) >> 1;                       // pair tiles: tiles 2i (rank 0) and 2i+1 (rank 1)
    int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
    const bool idle_pair = (p.dbg_flags & 8) && (pair & 1);      // experiment: only every other CTA pair works
    if (p.dbg_flags & 8) { pair >>= 1; npairs = (npairs + 1) >> 1; }
    const int my_tiles = (n_ptiles > pair && !idle_pair) ? (n_ptiles - 1 - pair) / npairs + 1 : 0;
    constexpr int W_BUILD = tc6::NEPI_WARPS, W_LOAD = W_BUILD + tc6::NBUILD / 32, W_ISSUE = W_LOAD + 1, W_FWD = W_ISSUE + 1;
    constexpr uint32_t SMASK = tc6::NSTAGE - 1;

    if (tid == 0) {
        for (int s = 0; s < tc6::NSTAGE; ++s) { mbar_init(&sm.bar_full[s], rank == 0 ? 2 : 1); mbar_init(&sm.bar_empty[s], 1); }
        mbar_init(&sm.bar_a1_ready, 2 * tc6::NBUILD);              // leader's: the builder threads of both CTAs
        mbar_init(&sm.bar_a1_free, 1);
        mbar_init(&sm.bar_acc_full, 1);
        mbar_init(&sm.bar_final, 1);
        mbar_init(&sm.bar_alpha, tc6::NEPI_WARPS);
        mbar_init(&sm.bar_drain, 2 * (tc6::NEPI_WARPS + tc6::NBUILD / 32));   // leader's: every warp of both CTAs that reads the layer-4 accumulator
        for (int c = 0; c < 8; ++c) mbar_init(&sm.bar_kblk[c], 2 * 4 * 2);   // leader's: 2 CTAs x 4 quadrant warps x 2 chunks
        mbar_fence_init();
        if (blockIdx.x == 0 && q.counters[PNB_QC_N_VALID] > p.hbar_cap) atomicExch(p.err, 9);
    }
    if (tid < TM) sm.alpha_e[tid] = 0.f;
    if (warp == W_ISSUE) tmem_alloc2<512>(&sm.tmem_base);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tP = sm.tmem_base, tQ = sm.tmem_base + 256u;
    const long long _tk0 = clock64();

    if (warp == W_LOAD) {
        // ============================================================ loader: this CTA's half of every image
        if (lane == 0) {
            const uint32_t total = (uint32_t)my_tiles * NBLK_TOTAL;            // K blocks
            const unsigned char* src = p.wimg + (size_t)rank * tc6::HIMG;
            for (uint32_t n = 0; n < total; ++n) {
                const uint32_t s = n & SMASK, ph = (n >> 2) & 1u;
                const long long _tl0 = clock64();
                if ((n % tc6::CPK) == 0 && !mbar_wait(&sm.bar_empty[(n / tc6::CPK) % (tc6::NSTAGE / tc6::CPK)], ph ^ 1u, p.err, 61)) break;
                if (blockIdx.x < 2 && (p.dbg_flags & 1)) atomicAdd(reinterpret_cast<unsigned long long*>(p.err) + 1 + (blockIdx.x ? 15 : 0), (unsigned long long)(clock64() - _tl0));
                if (p.dbg_no_weights) { mbar_arrive(&sm.bar_full[s]); continue; }
                mbar_arrive_expect_tx(&sm.bar_full[s], 2 * tc6::HIMG);
                const unsigned char* g = src + (size_t)(n % NBLK_TOTAL) * (2 * IMG);
                bulk_g2s(sm.b[s][0], g, tc6::HIMG, &sm.bar_full[s]);               // W_hi rows 128*rank..
                bulk_g2s(sm.b[s][1], g + IMG, tc6::HIMG, &sm.bar_full[s]);         // W_lo rows 128*rank..
            }
        }
    } else if (warp == W_FWD) {
        // ============================================================ forwarder (peer CTA): "my half landed" -> leader
        if (rank == 1 && lane == 0) {
            const uint32_t total = (uint32_t)my_tiles * NBLK_TOTAL;
            const uint32_t full0 = map_to_cta(&sm.bar_full[0], 0);
            for (uint32_t n = 0; n < total; ++n) {
                const uint32_t s = n & SMASK, ph = (n >> 2) & 1u;
                if (!mbar_wait(&sm.bar_full[s], ph, p.err, 62)) break;
                mbar_arrive_cluster(full0 + 8u * s);
            }
        }
    } else if (warp == W_ISSUE) {
        // ============================================================ MMA issuer: the whole warp of the rank-0 CTA
        if (rank == 0) {
            const uint32_t idesc = make_idesc_bf16(256, 256);
            const uint32_t hiw = desc_hi<LAYOUT>(), xe_hiw = (256u >> 4) | (1u << 14);
            const uint32_t b0_lo = desc_lo<LAYOUT>(smem_u32(sm.b[0]));
            const uint32_t ahi_lo = desc_lo<LAYOUT>(smem_u32(sm.a_hi)), alo_lo = desc_lo<LAYOUT>(smem_u32(sm.a_lo));
            const uint32_t xeh_lo0 = desc_lo<LAYOUT_NONE>(smem_u32(sm.xe_hi[0])), xel_lo0 = desc_lo<LAYOUT_NONE>(smem_u32(sm.xe_lo[0]));
            constexpr uint32_t KADV = kstep_adv16<LAYOUT>();
            uint32_t n = 0;            // weight image counter
            uint32_t c_acc = 0;        // completions of bar_acc_full consumed
            uint32_t c_pack = 0;       // packing rounds consumed on bar_kblk[*]
            bool ok = true;
            for (int t = 0; t < my_tiles && ok; ++t) {
                const uint32_t xeh_lo = xeh_lo0 + (uint32_t)(t & 1) * (tc3::XE >> 4), xel_lo = xel_lo0 + (uint32_t)(t & 1) * (tc3::XE >> 4);
                for (int l = 0; l < 4 && ok; ++l) {
                    const uint32_t acc = (l & 1) ? tP : tQ;
                    const uint32_t ab = (l & 1) ? tQ : tP;
                    if (l > 0) { if (!PNB_TIMED_WAIT_L0(2, mbar_wait(&sm.bar_acc_full, c_acc & 1u, p.err, 63))) { ok = false; break; } ++c_acc; }
                    else if (t > 0) { if (!PNB_TIMED_WAIT_L0(2, mbar_wait(&sm.bar_final, (uint32_t)(t - 1) & 1u, p.err, 63))) { ok = false; break; } }
                    if (l == 0) { if (!PNB_TIMED_WAIT_L0(1, mbar_wait(&sm.bar_a1_ready, (uint32_t)t & 1u, p.err, 64))) { ok = false; break; } }
                    if (l == 1 && t > 0) { if (!PNB_TIMED_WAIT_L0(2, mbar_wait(&sm.bar_drain, (uint32_t)(t - 1) & 1u, p.err, 65))) { ok = false; break; } }
                    tc_fence_after();
                    const int nkb = nkb_of(l);
                    for (int kb = 0; kb < nkb && ok; ++kb) {
                        const uint32_t s0 = n & SMASK, ph0 = (n >> 2) & 1u;                  // ring stage of this K block
                        const bool need_chunks = (l >= 1 && kb < 8);
                        uint64_t* cb0 = need_chunks ? &sm.bar_kblk[kb] : &sm.bar_full[s0];
                        const uint32_t cp0 = need_chunks ? (c_pack & 1u) : ph0;
                        if (p.dbg_flags & 2) {      // diagnosis: classify the wait with non-blocking probes
                            if (need_chunks && !mbar_test_wait(cb0, cp0)) {
                                if (!PNB_TIMED_WAIT_L0(12, mbar_spin_wait(cb0, cp0, p.err, 66))) { ok = false; break; }
                            }
                            if (!mbar_test_wait(&sm.bar_full[s0], ph0)) {
                                const long long _tw = clock64();
                                if (!mbar_spin_wait(&sm.bar_full[s0], ph0, p.err, 67)) { ok = false; break; }
                                if (lane == 0) { prof_add(p, 13, clock64() - _tw); prof_add(p, 14, 1); }
                            }
                        } else
                        if (!mbar_try_wait4(&sm.bar_full[s0], ph0, cb0, cp0, &sm.bar_full[s0], ph0, cb0, cp0)) {
                            if (need_chunks) {
                                if (!PNB_TIMED_WAIT_L0(2, mbar_wait(cb0, cp0, p.err, 66))) { ok = false; break; }
                            }
                            if (!PNB_TIMED_WAIT_L0(3, mbar_wait(&sm.bar_full[s0], ph0, p.err, 67))) { ok = false; break; }
                        }
                        tc_fence_after();
                        const uint32_t akb_hi = ahi_lo + (uint32_t)kb * (ABLK >> 4), akb_lo = alo_lo + (uint32_t)kb * (ABLK >> 4);
                        const uint32_t tcol = ab + (uint32_t)(kb * 32);
                        const uint32_t bl = b0_lo + s0 * (2 * tc6::HIMG >> 4), bl2 = bl + (tc6::HIMG >> 4);
                        if (l == 0) {
                            mma2_ss2_w(acc, akb_hi, hiw, bl, hiw, idesc, kb ? 1u : 0u);
                            mma2_ss2_w(acc, akb_lo, hiw, bl, hiw, idesc, 1u);
                            mma2_ss2_w(acc, akb_hi + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                            mma2_ss2_w(acc, akb_lo + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                            mma2_ss2_w(acc, akb_hi, hiw, bl2, hiw, idesc, 1u);
                            mma2_ss2_w(acc, akb_hi + KADV, hiw, bl2 + KADV, hiw, idesc, 1u);
                        } else if (kb == 8) {
                            mma2_ss2_w(acc, xeh_lo, xe_hiw, bl, hiw, idesc, 1u);
                            mma2_ss2_w(acc, xel_lo, xe_hiw, bl, hiw, idesc, 1u);
                            mma2_ss2_w(acc, xeh_lo, xe_hiw, bl2, hiw, idesc, 1u);
                        } else {
                            mma2_ts2_w(acc, tcol, bl, hiw, idesc, kb ? 1u : 0u);
                            mma2_ts2_w(acc, tcol + 8u, bl, hiw, idesc, 1u);
                            mma2_ts2_w(acc, tcol + 16u, bl + KADV, hiw, idesc, 1u);
                            mma2_ts2_w(acc, tcol + 24u, bl + KADV, hiw, idesc, 1u);
                            mma2_ts2_w(acc, tcol, bl2, hiw, idesc, 1u);
                            mma2_ts2_w(acc, tcol + 16u, bl2 + KADV, hiw, idesc, 1u);
                        }
                        if ((n % tc6::CPK) == tc6::CPK - 1) mma2_commit_w(&sm.bar_empty[(n / tc6::CPK) % (tc6::NSTAGE / tc6::CPK)], 3);   // frees CPK ring stages in both CTAs
                        n += 1;
                    }
                    if (!ok) break;
                    if (l >= 1) ++c_pack;
                    mma2_commit_w(l < 3 ? &sm.bar_acc_full : &sm.bar_final, 3);     // layers 1-3 -> epilogue warps, layer 4 -> builder warps
                    if (l == 0) mma2_commit_w(&sm.bar_a1_free, 3);
                }
            }
        }
    } else if (warp >= W_BUILD) {
        // ============================================================ builders: one thread per pair row of this CTA's tile
        const int bt = (warp - W_BUILD) * 32 + lane, row = bt & 127, part = bt >> 7;     // two threads per row (columns 0..151 | 152..287)
        const uint32_t ready0 = map_to_cta(&sm.bar_a1_ready, 0), drain0 = map_to_cta(&sm.bar_drain, 0);
        // the same warps run the LAST epilogue of a tile (alpha branch, Softplus, weighted K-reduction -> h-bar, sigma): it falls
        // under layer 1 of the next tile, exactly when the builders are idle (the operand buffer is still being read)
        const int quad = warp & 3, grp = (warp - W_BUILD) >> 2;
        const int erow = quad * 32 + lane;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        bool ok = true;
        for (int t = 0; t <= my_tiles && ok; ++t) {
            if (t < my_tiles) {
                const int tile = 2 * (pair + t * npairs) + (int)rank;
                if (t > 0 && !(lane == 0 && warp == W_BUILD ? PNB_TIMED_WAIT(4, mbar_wait(&sm.bar_a1_free, (uint32_t)(t - 1) & 1u, p.err, 68)) : mbar_wait(&sm.bar_a1_free, (uint32_t)(t - 1) & 1u, p.err, 68))) { ok = false; break; }
                const long long _tb0 = clock64();
                if (part == 0) build_pair_part<0, false>(sm, p, tile, t, row, n_valid);
                else build_pair_part<1, false>(sm, p, tile, t, row, n_valid);
                fence_proxy_async();
                if (rank == 0) mbar_arrive(&sm.bar_a1_ready); else mbar_arrive_cluster(ready0);
                if (lane == 0 && warp == W_BUILD) prof_add(p, 5, clock64() - _tb0);
            }
            if (t > 0) {
                const int tf = t - 1, tilef = 2 * (pair + tf * npairs) + (int)rank;
                if (!mbar_wait(&sm.bar_final, (uint32_t)tf & 1u, p.err, 70)) { ok = false; break; }
                const long long _te0 = clock64();
                tc_fence_after();
                const uint32_t accb = tP + tlane;          // layer 4 accumulates into region P
                const float wrow = sm.wc[tf & 1][erow];
                const int sidx = tilef * TSAMP + (erow >> 3);
                const bool swrite = sidx < n_valid;
                const float apart = last_chunks<4, 4>(p, accb, 2 + grp, wrow, sidx, swrite, lane);
                tc_fence_before();
                if (bt == 0) prof_add(p, 8, clock64() - _te0);
                sm.alpha_part[grp][erow] = apart;
                named_bar_sync(2, tc6::NBUILD);
                if (!mbar_wait(&sm.bar_alpha, (uint32_t)tf & 1u, p.err, 71)) { ok = false; break; }      // the epilogue warps' partial sums
                if (grp == 0) {
                    float a = (sm.alpha_part[0][erow] + sm.alpha_part[1][erow]) + sm.alpha_e[erow] + __ldg(p.ba) - 1.0f;
                    sm.alpha_e[erow] = 0.f;
                    float sp = a > 20.f ? a : log1pf(expf(a));
                    float zz = sp * wrow;
                    zz += __shfl_xor_sync(0xffffffffu, zz, 1);
                    zz += __shfl_xor_sync(0xffffffffu, zz, 2);
                    zz += __shfl_xor_sync(0xffffffffu, zz, 4);
                    if ((lane & 7) == 0 && swrite) p.sigma[sidx] = zz;
                }
                __syncwarp();
                if (lane == 0) { if (rank == 0) mbar_arrive(&sm.bar_drain); else mbar_arrive_cluster(drain0); }   // after the alpha_e reads (see Smem)
                named_bar_sync(2, tc6::NBUILD);
            }
        }
    } else {
        // ============================================================ epilogue warps
        const int quad = warp & 3, grp = warp >> 2;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        const uint32_t kblk0 = map_to_cta(&sm.bar_kblk[0], 0), drain0 = map_to_cta(&sm.bar_drain, 0);
        uint32_t n_acc = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            for (int l = 0; l < 3 && ok; ++l, ++n_acc) {        // the layer-4 (last) epilogue runs on the builder warps
                if (!(tid == 0 ? PNB_TIMED_WAIT(6, mbar_wait(&sm.bar_acc_full, n_acc & 1u, p.err, 69)) : mbar_wait(&sm.bar_acc_full, n_acc & 1u, p.err, 69))) { ok = false; break; }
                const long long _te0 = clock64();
                tc_fence_after();
                const uint32_t accb = ((l & 1) ? tP : tQ) + tlane;
                {
                    const float* bias = p.bias[l];
#pragma unroll
                    for (int i = 0; i < tc6::NCH; ++i) {
                        const int g = grp + tc6::NGRP * i, c0 = 16 * g;
                        uint32_t v[16];
                        tmem_ld16(accb + (uint32_t)c0, v);
                        tmem_ld_wait();
                        uint32_t hh[8], ll[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c0) + e);
                            float y0 = __uint_as_float(v[2 * e]) + bb.x, y1 = __uint_as_float(v[2 * e + 1]) + bb.y;
                            y0 = fmaxf(y0, LEAKY * y0); y1 = fmaxf(y1, LEAKY * y1);
                            split_bf16x2(y0, y1, hh[e], ll[e]);
                        }
                        tmem_st8(accb + (uint32_t)c0, hh);
                        tmem_st8(accb + (uint32_t)c0 + 8u, ll);
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (rank == 0) mbar_arrive(&sm.bar_kblk[g >> 1]); else mbar_arrive_cluster(kblk0 + 8u * (uint32_t)(g >> 1)); }
                    }
                    if (tid == 0) prof_add(p, 7, clock64() - _te0);
                }
            }
            if (!ok) break;
            {   // this warp's share of the LAST epilogue (chunk groups 0, 1; the builder warps take 2, 3)
                const int tile = 2 * (pair + t * npairs) + (int)rank;
                if (!mbar_wait(&sm.bar_final, (uint32_t)t & 1u, p.err, 72)) { ok = false; break; }
                tc_fence_after();
                const int erow = quad * 32 + lane;
                const int sidx = tile * TSAMP + (erow >> 3);
                const float apart = last_chunks<4, 4>(p, tP + tlane, grp, sm.wc[t & 1][erow], sidx, sidx < n_valid, lane);
                tc_fence_before();
                atomicAdd(&sm.alpha_e[erow], apart);
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&sm.bar_alpha);
                    if (rank == 0) mbar_arrive(&sm.bar_drain); else mbar_arrive_cluster(drain0);
                }
            }
        }
    }
    if (tid == 0) prof_add(p, 9, clock64() - _tk0);
    if (tid == 0 && (p.dbg_flags & 4) && blockIdx.x < 192) {      // per-CTA cycles and SM id (needs a 512-int err buffer)
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        reinterpret_cast<long long*>(p.err)[32 + blockIdx.x] = ((clock64() - _tk0) & 0xffffffffffffll) | ((long long)smid << 48);
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == W_ISSUE) tmem_dealloc2<512>(sm.tmem_base);
}

// =====================================================================================================================
// Colour branch on the tensor cores: per 128 valid samples  [hbar(256) | PE4(view)(24)] -> 128 -> 128 -> 128 (tcgen05,
// BF16x3) -> 3 (CUDA cores) -> sigmoid*1.002-0.001   (reference: point_aggregators.py:631-637, 269-273).
// Layer 1 reads its operand from shared memory (SS), layers 2-3 from tensor memory (TS).  TMEM: accumulator cols
// 0..127, A_hi 128..191, A_lo 192..255.  320 threads: 8 worker warps (build + epilogues), loader, issuer.
namespace ctc {
constexpr int NTHR = 320, NWORK = 256, NSTAGE = 4;
constexpr int IMG = 128 * 64;                 // [128 x 32] bf16 weight image
constexpr int NBLK = 9 + 4 + 4;               // K blocks of the three layers
constexpr int IMGS_PER_TILE = 2 * NBLK;
__host__ __device__ constexpr int nkb_of(int l) { return l == 0 ? 9 : 4; }
__host__ __device__ constexpr int img_base(int l) { return l == 0 ? 0 : l == 1 ? 9 : 13; }
struct Smem {
    unsigned char a_hi[9 * tc::ABLK];
    unsigned char a_lo[9 * tc::ABLK];
    unsigned char b[NSTAGE][IMG];
    float part[2][128][4];
    uint32_t samp[128];
    uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_a_ready, bar_acc_full;
    uint32_t tmem_base;
};
}  // namespace ctc

struct ColorTcParams {
    pnb_query_t q;
    pnb_shade_opts_t o;
    const unsigned char* wimg;      // packed colour_branch.{0,2,4} images
    const float* bias[3];
    const float* w3t;               // colour_branch.6 W^T [128][3]
    const float* b3;
    const float* hbar;
    const float* sigma;
    int hbar_cap;
    float4* sigma_rgb;
    int* err;
};

__global__ void __launch_bounds__(ctc::NTHR, 1) k_color_tc(ColorTcParams p) {
    using namespace ctc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const pnb_query_t& q = p.q;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], p.hbar_cap);
    const int n_tiles = (n_valid + 127) / 128;
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&sm.bar_full[s], 1); mbar_init(&sm.bar_empty[s], 1); }
        mbar_init(&sm.bar_a_ready, NWORK);
        mbar_init(&sm.bar_acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc<256>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = sm.tmem_base, t_ahi = tacc + 128u, t_alo = tacc + 192u;

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t total = (uint32_t)my_tiles * IMGS_PER_TILE;
            for (uint32_t n = 0; n < total; ++n) {
                const uint32_t s = n % NSTAGE, ph = (n / NSTAGE) & 1u;
                if (!mbar_wait(&sm.bar_empty[s], ph ^ 1u, p.err, 21)) break;
                mbar_arrive_expect_tx(&sm.bar_full[s], IMG);
                bulk_g2s(sm.b[s], p.wimg + (size_t)(n % IMGS_PER_TILE) * IMG, IMG, &sm.bar_full[s]);
            }
        }
    } else if (warp == 9) {
        {   // whole warp, warp-uniform; one elected lane issues
            const uint32_t idesc = make_idesc_bf16(128, 128);
            const uint32_t hiw = desc_hi<tc::LAYOUT>();
            const uint32_t b0_lo = desc_lo<tc::LAYOUT>(smem_u32(sm.b[0]));
            const uint32_t ahi_lo = desc_lo<tc::LAYOUT>(smem_u32(sm.a_hi)), alo_lo = desc_lo<tc::LAYOUT>(smem_u32(sm.a_lo));
            constexpr uint32_t KADV = kstep_adv16<tc::LAYOUT>();
            uint32_t n = 0, lyr = 0;
            bool ok = true;
            for (int t = 0; t < my_tiles && ok; ++t) {
                for (int l = 0; l < 3 && ok; ++l, ++lyr) {
                    if (!mbar_wait(&sm.bar_a_ready, lyr & 1u, p.err, 22)) { ok = false; break; }
                    tc_fence_after();
                    const int nkb = nkb_of(l);
                    for (int kb = 0; kb < nkb && ok; ++kb) {
                        const uint32_t akb_hi = ahi_lo + (uint32_t)kb * (tc::ABLK >> 4), akb_lo = alo_lo + (uint32_t)kb * (tc::ABLK >> 4);
                        const uint32_t tcol = (uint32_t)(kb * 16);
                        {
                            const uint32_t s = n & (NSTAGE - 1), ph = (n >> 2) & 1u;
                            if (!mbar_wait(&sm.bar_full[s], ph, p.err, 23)) { ok = false; break; }
                            tc_fence_after();
                            const uint32_t bl = b0_lo + s * (IMG >> 4);
                            if (l == 0) {
                                mma_ss2_w(tacc, akb_hi, hiw, bl, hiw, idesc, kb ? 1u : 0u);
                                mma_ss2_w(tacc, akb_lo, hiw, bl, hiw, idesc, 1u);
                                mma_ss2_w(tacc, akb_hi + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                                mma_ss2_w(tacc, akb_lo + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                            } else {
                                mma_ts2_w(tacc, t_ahi + tcol, bl, hiw, idesc, kb ? 1u : 0u);
                                mma_ts2_w(tacc, t_alo + tcol, bl, hiw, idesc, 1u);
                                mma_ts2_w(tacc, t_ahi + tcol + 8u, bl + KADV, hiw, idesc, 1u);
                                mma_ts2_w(tacc, t_alo + tcol + 8u, bl + KADV, hiw, idesc, 1u);
                            }
                            mma_commit_w(&sm.bar_empty[s]);
                            ++n;
                        }
                        {
                            const uint32_t s = n & (NSTAGE - 1), ph = (n >> 2) & 1u;
                            if (!mbar_wait(&sm.bar_full[s], ph, p.err, 23)) { ok = false; break; }
                            tc_fence_after();
                            const uint32_t bl = b0_lo + s * (IMG >> 4);
                            if (l == 0) {
                                mma_ss2_w(tacc, akb_hi, hiw, bl, hiw, idesc, 1u);
                                mma_ss2_w(tacc, akb_hi + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                            } else {
                                mma_ts2_w(tacc, t_ahi + tcol, bl, hiw, idesc, 1u);
                                mma_ts2_w(tacc, t_ahi + tcol + 8u, bl + KADV, hiw, idesc, 1u);
                            }
                            mma_commit_w(&sm.bar_empty[s]);
                            ++n;
                        }
                    }
                    mma_commit_w(&sm.bar_acc_full);
                }
            }
        }
    } else {
        const int quad = warp & 3, half = warp >> 2;
        const int erow = quad * 32 + lane;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        uint32_t lyr = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            {   // ---- build: 2 threads per sample row
                const int row = warp * 16 + (lane >> 1), hf = lane & 1;
                const int vi = tile * 128 + row;
                uint32_t s = 0xffffffffu;
                if (vi < n_valid) {
                    s = q.valid_list[vi];
                    const float4* src = (const float4*)(p.hbar + (size_t)vi * 256 + hf * 128);
#pragma unroll 4
                    for (int c = 0; c < 16; ++c) {
                        float4 a = __ldg(src + 2 * c), b = __ldg(src + 2 * c + 1);
                        float f8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                        int col = hf * 128 + 8 * c;
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) split_bf16x2(f8[2 * i], f8[2 * i + 1], h[i], l[i]);
                        uint32_t off = (uint32_t)(col >> 5) * tc::ABLK + tile_offset_bytes<tc::LAYOUT>(row, col & 31);
                        *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                } else {
                    for (int c = 0; c < 16; ++c) {
                        int col = hf * 128 + 8 * c;
                        uint32_t off = (uint32_t)(col >> 5) * tc::ABLK + tile_offset_bytes<tc::LAYOUT>(row, col & 31);
                        *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                if (hf == 0) {
                    sm.samp[row] = s;
                    float pe[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) pe[i] = 0.f;
                    if (s != 0xffffffffu) {
                        int r = (int)(q.samp_ray[s] >> 7);
                        float ov[3];
                        rot3t(p.o.Rw2c, q.raydir[3 * r], q.raydir[3 * r + 1], q.raydir[3 * r + 2], ov[0], ov[1], ov[2]);
#pragma unroll
                        for (int d = 0; d < 3; ++d) {     // ori=True layout: sin block (d*4+j) then cos block
                            float sc[8];
                            pe_doubling<4>(ov[d], sc);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { pe[d * 4 + j] = sc[2 * j]; pe[12 + d * 4 + j] = sc[2 * j + 1]; }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) split_bf16x2(pe[8 * c + 2 * i], pe[8 * c + 2 * i + 1], h[i], l[i]);
                        uint32_t off = 8u * tc::ABLK + tile_offset_bytes<tc::LAYOUT>(row, 8 * c);
                        *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&sm.bar_a_ready);
            for (int l = 0; l < 3 && ok; ++l, ++lyr) {
                if (!mbar_wait(&sm.bar_acc_full, lyr & 1u, p.err, 24)) { ok = false; break; }
                tc_fence_after();
                const float* bias = p.bias[l];
                float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    const int c0 = half * 64 + ch * 32;
                    uint32_t v[32];
                    tmem_ld32(tacc + tlane + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if (l < 2) {
                        uint32_t hh[16], ll[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c0) + e);
                            float y0 = __uint_as_float(v[2 * e]) + bb.x, y1 = __uint_as_float(v[2 * e + 1]) + bb.y;
                            y0 = fmaxf(y0, tc::LEAKY * y0); y1 = fmaxf(y1, tc::LEAKY * y1);
                            split_bf16x2(y0, y1, hh[e], ll[e]);
                        }
                        const uint32_t colp = (uint32_t)(c0 >> 1);
                        tmem_st8(t_ahi + tlane + colp, hh);
                        tmem_st8(t_ahi + tlane + colp + 8u, hh + 8);
                        tmem_st8(t_alo + tlane + colp, ll);
                        tmem_st8(t_alo + tlane + colp + 8u, ll + 8);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            float y = __uint_as_float(v[e]) + __ldg(bias + c0 + e);
                            y = fmaxf(y, tc::LEAKY * y);
                            const float* wr = p.w3t + (c0 + e) * 3;
                            d0 = fmaf(y, __ldg(wr), d0); d1 = fmaf(y, __ldg(wr + 1), d1); d2 = fmaf(y, __ldg(wr + 2), d2);
                        }
                    }
                }
                if (l < 2) {
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&sm.bar_a_ready);
                } else {
                    tc_fence_before();
                    sm.part[half][erow][0] = d0; sm.part[half][erow][1] = d1; sm.part[half][erow][2] = d2;
                    named_bar_sync(1, NWORK);
                    if (half == 0) {
                        uint32_t s = sm.samp[erow];
                        if (s != 0xffffffffu) {
                            float4 o4;
                            o4.x = p.sigma[tile * 128 + erow];
                            float r0 = sm.part[0][erow][0] + sm.part[1][erow][0] + __ldg(p.b3);
                            float r1 = sm.part[0][erow][1] + sm.part[1][erow][1] + __ldg(p.b3 + 1);
                            float r2 = sm.part[0][erow][2] + sm.part[1][erow][2] + __ldg(p.b3 + 2);
                            o4.y = 1.0f / (1.0f + expf(-r0)) * (1.0f + 2.0f * 0.001f) - 0.001f;
                            o4.z = 1.0f / (1.0f + expf(-r1)) * (1.0f + 2.0f * 0.001f) - 0.001f;
                            o4.w = 1.0f / (1.0f + expf(-r2)) * (1.0f + 2.0f * 0.001f) - 0.001f;
                            p.sigma_rgb[s] = o4;
                        }
                    }
                    named_bar_sync(1, NWORK);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<256>(tacc);
}

// =====================================================================================================================
// Colour branch, pipelined (v2).  The pair kernel (v6) writes h-bar already split into bf16 hi / lo and laid out as the
// tcgen05 A-operand blocks of this kernel: per 128 consecutive valid samples 8 K-blocks x {hi, lo} x [128 x 32]
// (interleaved core-matrix layout) = one contiguous 128 KB region.  So the layer-1 operand is ONE bulk copy (TMA engine)
// per tile - no builder warps, no register traffic - issued as soon as the previous tile's layer-1 MMAs have completed,
// and it lands under that tile's layers 2-3.  Same TMEM role ping-pong / in-place accumulator->operand conversion and
// chunk-granular hand-off as the pair kernel; two accumulator sets (tile parity) so a tile's final epilogue
// (128 -> 3 on CUDA cores + sigmoid) runs under the next tile's layer 1.  One tcgen05.commit per K block.
//   layer 1: A smem (8 K-blocks h-bar + 1 K-block PE(view)), acc X      layer 2: A = X, acc Y      layer 3: A = Y, acc X
//   X = 256*(t&1), Y = X + 128 (TMEM columns)
// Warps (480 threads): 0-7 epilogue (quadrant = w & 3, chunk group = w >> 2), 8-11 PE(view) builders (thread = row),
// 12 operand loader, 13 weight loader, 14 issuer (whole warp, warp-uniform).
namespace ctc2 {
constexpr int NEPI_WARPS = 8, NGRP = 2, NCH = 8 / NGRP;     // 8 chunks of 16 accumulator columns per layer
constexpr int NEPI = NEPI_WARPS * 32, NTHR = NEPI + 128 + 96;
constexpr int NSTAGE = 4;
constexpr int BLK = 128 * 64;                // [128 x 32] bf16 block (operand block and weight image)
constexpr int NKB = 9 + 4 + 4;               // K blocks per tile
constexpr int TILE_BYTES = 8 * 2 * BLK;      // h-bar operand of one tile in global memory
struct Smem {
    unsigned char a[8][2][BLK];              // layer-1 operand: K block, hi/lo
    unsigned char pe[2][BLK];                // PE(view) K block, hi/lo
    unsigned char b[NSTAGE][2][BLK];         // weight ring: hi / lo image of one K block
    float part[128][4];
    uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_a_full, bar_a_free, bar_pe_ready, bar_acc_full, bar_drain, bar_kblk[4];
    uint32_t tmem_base;
};
__host__ __device__ constexpr int nkb_of(int l) { return l == 0 ? 9 : 4; }
}  // namespace ctc2

__global__ void __launch_bounds__(ctc2::NTHR, 1) k_color_tc2(ColorTcParams p) {
    using namespace ctc2;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const pnb_query_t& q = p.q;
    const int n_valid = min(q.counters[PNB_QC_N_VALID], p.hbar_cap);
    const int n_tiles = (n_valid + 127) / 128;
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    constexpr int W_PE = NEPI_WARPS, W_LOADA = W_PE + 4, W_LOADW = W_LOADA + 1, W_ISSUE = W_LOADW + 1;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&sm.bar_full[s], 1); mbar_init(&sm.bar_empty[s], 1); }
        mbar_init(&sm.bar_a_full, 1);
        mbar_init(&sm.bar_a_free, 1);
        mbar_init(&sm.bar_pe_ready, 128);
        mbar_init(&sm.bar_acc_full, 1);
        mbar_init(&sm.bar_drain, NEPI_WARPS);
        for (int c = 0; c < 4; ++c) mbar_init(&sm.bar_kblk[c], 4 * 2);      // 4 quadrant warps x 2 chunks of 16 columns
        mbar_fence_init();
    }
    if (warp == W_ISSUE) tmem_alloc<512>(&sm.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = sm.tmem_base;

    if (warp == W_LOADA) {
        // ============================================================ layer-1 operand: one 128 KB region per tile
        if (lane == 0) {
            for (int t = 0; t < my_tiles; ++t) {
                const int tile = (int)blockIdx.x + t * (int)gridDim.x;
                if (t > 0 && !mbar_wait(&sm.bar_a_free, (uint32_t)(t - 1) & 1u, p.err, 81)) break;
                mbar_arrive_expect_tx(&sm.bar_a_full, TILE_BYTES);
                const unsigned char* src = reinterpret_cast<const unsigned char*>(p.hbar) + (size_t)tile * TILE_BYTES;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) bulk_g2s(sm.a[c][0], src + (size_t)c * 2 * BLK, 2 * BLK, &sm.bar_a_full);
            }
        }
    } else if (warp == W_LOADW) {
        // ============================================================ weight ring: hi + lo image of one K block per stage
        if (lane == 0) {
            const uint32_t total = (uint32_t)my_tiles * NKB;
            for (uint32_t n = 0; n < total; ++n) {
                const uint32_t s = n & (NSTAGE - 1), ph = (n >> 2) & 1u;
                if (!mbar_wait(&sm.bar_empty[s], ph ^ 1u, p.err, 82)) break;
                mbar_arrive_expect_tx(&sm.bar_full[s], 2 * BLK);
                bulk_g2s(sm.b[s][0], p.wimg + (size_t)(n % NKB) * 2 * BLK, 2 * BLK, &sm.bar_full[s]);
            }
        }
    } else if (warp == W_ISSUE) {
        // ============================================================ MMA issuer (whole warp, warp-uniform)
        const uint32_t idesc = make_idesc_bf16(128, 128);
        const uint32_t hiw = desc_hi<tc::LAYOUT>();
        const uint32_t b0_lo = desc_lo<tc::LAYOUT>(smem_u32(sm.b[0][0]));
        const uint32_t a0_lo = desc_lo<tc::LAYOUT>(smem_u32(sm.a[0][0])), pe_lo = desc_lo<tc::LAYOUT>(smem_u32(sm.pe[0]));
        constexpr uint32_t KADV = kstep_adv16<tc::LAYOUT>(), BADV = BLK >> 4;
        uint32_t n = 0, c_acc = 0, c_pack = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const uint32_t tX = tbase + 256u * (uint32_t)(t & 1), tY = tX + 128u;
            for (int l = 0; l < 3 && ok; ++l) {
                const uint32_t acc = (l == 1) ? tY : tX, ab = (l == 1) ? tX : tY;
                // WAR on tensor memory: the MMAs of the previous layer read this layer's accumulator region as their A operand and must
                // be complete.  Every completion of bar_acc_full is consumed before the commit of the next one is issued (a parity
                // wait overtaken by two completions would never return): layer 3 of the previous tile is consumed inside layer 1,
                // before its last K block (see below); the previous tile's drain here, where it is complete in steady state.
                if (l == 1 && t > 0) {
                    if (!mbar_wait(&sm.bar_drain, (uint32_t)(t - 1) & 1u, p.err, 86)) { ok = false; break; }
                }
                if (l > 0) { if (!mbar_wait(&sm.bar_acc_full, c_acc & 1u, p.err, 83)) { ok = false; break; } ++c_acc; }
                if (l == 0) {
                    if (!mbar_wait(&sm.bar_a_full, (uint32_t)t & 1u, p.err, 84)) { ok = false; break; }
                    if (!mbar_wait(&sm.bar_pe_ready, (uint32_t)t & 1u, p.err, 85)) { ok = false; break; }
                }
                tc_fence_after();
                const int nkb = nkb_of(l);
                for (int kb = 0; kb < nkb && ok; ++kb, ++n) {
                    const uint32_t s = n & (NSTAGE - 1), ph = (n >> 2) & 1u;
                    const bool need_chunks = l >= 1;
                    if (l == 0 && t > 0 && kb == nkb - 1) {      // layer 3 of the previous tile: complete by now (ring depth < 8 K blocks)
                        if (!mbar_wait(&sm.bar_acc_full, c_acc & 1u, p.err, 83)) { ok = false; break; }
                        ++c_acc;
                    }
                    uint64_t* cb0 = need_chunks ? &sm.bar_kblk[kb] : &sm.bar_full[s];
                    const uint32_t cp0 = need_chunks ? (c_pack & 1u) : ph;
                    if (!mbar_try_wait4(&sm.bar_full[s], ph, cb0, cp0, &sm.bar_full[s], ph, cb0, cp0)) {
                        if (need_chunks && !mbar_wait(cb0, cp0, p.err, 87)) { ok = false; break; }
                        if (!mbar_wait(&sm.bar_full[s], ph, p.err, 88)) { ok = false; break; }
                    }
                    tc_fence_after();
                    const uint32_t bh = b0_lo + s * (2 * BADV), bl = bh + BADV;
                    if (l == 0) {
                        const uint32_t ah = (kb < 8) ? a0_lo + (uint32_t)kb * (2 * BADV) : pe_lo, al = ah + BADV;
                        mma_ss2_w(acc, ah, hiw, bh, hiw, idesc, kb ? 1u : 0u);
                        mma_ss2_w(acc, al, hiw, bh, hiw, idesc, 1u);
                        mma_ss2_w(acc, ah + KADV, hiw, bh + KADV, hiw, idesc, 1u);
                        mma_ss2_w(acc, al + KADV, hiw, bh + KADV, hiw, idesc, 1u);
                        mma_ss2_w(acc, ah, hiw, bl, hiw, idesc, 1u);
                        mma_ss2_w(acc, ah + KADV, hiw, bl + KADV, hiw, idesc, 1u);
                    } else {
                        const uint32_t tcol = ab + (uint32_t)(kb * 32);        // in-place packed operand: 16-column chunk = 8 hi | 8 lo
                        mma_ts2_w(acc, tcol, bh, hiw, idesc, kb ? 1u : 0u);
                        mma_ts2_w(acc, tcol + 8u, bh, hiw, idesc, 1u);
                        mma_ts2_w(acc, tcol + 16u, bh + KADV, hiw, idesc, 1u);
                        mma_ts2_w(acc, tcol + 24u, bh + KADV, hiw, idesc, 1u);
                        mma_ts2_w(acc, tcol, bl, hiw, idesc, 1u);
                        mma_ts2_w(acc, tcol + 16u, bl + KADV, hiw, idesc, 1u);
                    }
                    mma_commit_w(&sm.bar_empty[s]);
                }
                if (!ok) break;
                if (l >= 1) ++c_pack;
                mma_commit_w(&sm.bar_acc_full);
                if (l == 0) mma_commit_w(&sm.bar_a_free);
            }
        }
    } else if (warp >= W_PE) {
        // ============================================================ PE(view) K block of the next tile (thread = sample row)
        const int row = (warp - W_PE) * 32 + lane;
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            if (t > 0 && !mbar_wait(&sm.bar_a_free, (uint32_t)(t - 1) & 1u, p.err, 89)) break;
            const int vi = tile * 128 + row;
            float pe[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) pe[i] = 0.f;
            if (vi < n_valid) {
                const uint32_t s = q.valid_list[vi];
                const int r = (int)(q.samp_ray[s] >> 7);
                float ov[3];
                rot3t(p.o.Rw2c, q.raydir[3 * r], q.raydir[3 * r + 1], q.raydir[3 * r + 2], ov[0], ov[1], ov[2]);
#pragma unroll
                for (int d = 0; d < 3; ++d) {     // ori=True layout: sin block (d*4+j) then cos block
                    float sc[8];
                    pe_doubling<4>(ov[d], sc);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { pe[d * 4 + j] = sc[2 * j]; pe[12 + d * 4 + j] = sc[2 * j + 1]; }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_bf16x2(pe[8 * c + 2 * i], pe[8 * c + 2 * i + 1], h[i], l[i]);
                const uint32_t off = tile_offset_bytes<tc::LAYOUT>(row, 8 * c);
                *reinterpret_cast<uint4*>(sm.pe[0] + off) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(sm.pe[1] + off) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();
            mbar_arrive(&sm.bar_pe_ready);
        }
    } else {
        // ============================================================ epilogue warps
        const int quad = warp & 3, grp = warp >> 2;
        const int erow = quad * 32 + lane;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        uint32_t n_acc = 0;
        bool ok = true;
        for (int t = 0; t < my_tiles && ok; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const uint32_t tX = tbase + 256u * (uint32_t)(t & 1), tY = tX + 128u;
            for (int l = 0; l < 3 && ok; ++l, ++n_acc) {
                if (!mbar_wait(&sm.bar_acc_full, n_acc & 1u, p.err, 90)) { ok = false; break; }
                tc_fence_after();
                const uint32_t accb = ((l == 1) ? tY : tX) + tlane;
                const float* bias = p.bias[l];
                if (l < 2) {
#pragma unroll
                    for (int i = 0; i < NCH; ++i) {
                        const int g = grp + NGRP * i, c0 = 16 * g;
                        uint32_t v[16];
                        tmem_ld16(accb + (uint32_t)c0, v);
                        tmem_ld_wait();
                        uint32_t hh[8], ll[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c0) + e);
                            float y0 = __uint_as_float(v[2 * e]) + bb.x, y1 = __uint_as_float(v[2 * e + 1]) + bb.y;
                            y0 = fmaxf(y0, tc::LEAKY * y0); y1 = fmaxf(y1, tc::LEAKY * y1);
                            split_bf16x2(y0, y1, hh[e], ll[e]);
                        }
                        tmem_st8(accb + (uint32_t)c0, hh);
                        tmem_st8(accb + (uint32_t)c0 + 8u, ll);
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.bar_kblk[g >> 1]);
                    }
                } else {
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
                    for (int i = 0; i < NCH; ++i) {
                        const int c0 = 16 * (grp + NGRP * i);
                        uint32_t v[16];
                        tmem_ld16(accb + (uint32_t)c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float y = __uint_as_float(v[e]) + __ldg(bias + c0 + e);
                            y = fmaxf(y, tc::LEAKY * y);
                            const float* wr = p.w3t + (c0 + e) * 3;
                            d0 = fmaf(y, __ldg(wr), d0); d1 = fmaf(y, __ldg(wr + 1), d1); d2 = fmaf(y, __ldg(wr + 2), d2);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.bar_drain);          // this warp's share of accumulator set X drained
                    if (grp == 1) { sm.part[erow][0] = d0; sm.part[erow][1] = d1; sm.part[erow][2] = d2; }
                    named_bar_sync(1, NEPI);
                    if (grp == 0) {
                        const int vi = tile * 128 + erow;
                        if (vi < n_valid) {
                            const uint32_t s = q.valid_list[vi];
                            float4 o4;
                            o4.x = p.sigma[vi];
                            const float r0 = d0 + sm.part[erow][0] + __ldg(p.b3);
                            const float r1 = d1 + sm.part[erow][1] + __ldg(p.b3 + 1);
                            const float r2 = d2 + sm.part[erow][2] + __ldg(p.b3 + 2);
                            o4.y = 1.0f / (1.0f + expf(-r0)) * (1.0f + 2.0f * 0.001f) - 0.001f;
                            o4.z = 1.0f / (1.0f + expf(-r1)) * (1.0f + 2.0f * 0.001f) - 0.001f;
                            o4.w = 1.0f / (1.0f + expf(-r2)) * (1.0f + 2.0f * 0.001f) - 0.001f;
                            p.sigma_rgb[s] = o4;
                        }
                    }
                    named_bar_sync(1, NEPI);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == W_ISSUE) tmem_dealloc<512>(sm.tmem_base);
}

}  // namespace pnb

using namespace pnb;

static size_t pack_pairs_bytes() { return (size_t)tc::NBLK_TOTAL * 2 * tc::IMG; }
static size_t pack_color_bytes() { return (size_t)ctc::NBLK * 2 * ctc::IMG; }
extern "C" size_t pnb_mlp_pack_bytes(void) { return pack_pairs_bytes() + pack_color_bytes(); }

// Packs block1/block3 weights (pnb_mlp_t W^T buffers, fp32) into tcgen05 operand images.  Call once per weight version.
extern "C" int pnb_mlp_pack(const pnb_mlp_t* mlp, void* d_out, size_t out_bytes, pnb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PNB_REQUIRE(mlp && d_out, PNB_ERR_INVALID, "pnb_mlp_pack: null argument");
    PNB_REQUIRE(out_bytes >= pnb_mlp_pack_bytes(), PNB_ERR_WORKSPACE, "pnb_mlp_pack: buffer too small");
    const int kpad[4] = {288, 256, 272, 256};
    for (int l = 0; l < 4; ++l) {
        int nkb = tc::nkb_of(l);
        int n = nkb * 256 * umma::BK;
        k_pack_weights<<<(n + 255) / 256, 256, 0, stream>>>(mlp->w[l], kpad[l], nkb, 256, (unsigned char*)d_out + (size_t)tc::img_base(l) * 2 * tc::IMG);
    }
    const int ckpad[3] = {288, 128, 128};   // colour_branch.{0,2,4}: W^T [K_pad][128]
    for (int l = 0; l < 3; ++l) {
        int nkb = ctc::nkb_of(l);
        int n = nkb * 128 * umma::BK;
        k_pack_weights<<<(n + 255) / 256, 256, 0, stream>>>(mlp->w[5 + l], ckpad[l], nkb, 128,
                                                            (unsigned char*)d_out + pack_pairs_bytes() + (size_t)ctc::img_base(l) * 2 * ctc::IMG);
    }
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}

static size_t pack_sc_max(int cap) { return (size_t)cap / tc7::PACK_S + 2; }
extern "C" size_t pnb_shade_tc_bytes(int max_valid_samples) {
    const size_t cap = (size_t)max_valid_samples;
    return align_up((cap + 127) / 128 * 128 * 256 * sizeof(float)) + align_up(cap * sizeof(float)) +
           2 * align_up(cap + 16) + align_up(pack_sc_max(max_valid_samples) * 4) + 2 * align_up((cap + 2) * 4) + align_up(16) + 256;   // + v7 packing tables
}

// Tensor-core forward: per-pair MLPs on tcgen05 (BF16x3), colour branch on CUDA cores.  ws: >= pnb_shade_tc_bytes.
// d_err: device int32, set non-zero if the in-kernel pipeline timed out (results invalid).
extern "C" int pnb_shade_forward_tc(const pnb_query_t* q, const pnb_points_t* pts, const pnb_mlp_t* mlp, const void* d_packed,
                                    const pnb_shade_opts_t* opts, float* d_sigma_rgb, void* ws, size_t ws_bytes,
                                    int max_valid_samples, int stage_mask, int* d_err, pnb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PNB_REQUIRE(q && pts && mlp && d_packed && opts && d_sigma_rgb && ws && d_err, PNB_ERR_INVALID, "pnb_shade_forward_tc: null argument");
    PNB_REQUIRE(q->K >= 1 && q->K <= PNB_MAX_K, PNB_ERR_UNSUPPORTED, "pnb_shade_forward_tc: K=%d unsupported", q->K);
    PNB_REQUIRE(ws_bytes >= pnb_shade_tc_bytes(max_valid_samples), PNB_ERR_WORKSPACE, "pnb_shade_forward_tc: workspace too small");
    static int configured[64] = {0}, n_sm_of[64] = {0};      // per device of this process (one process per GPU is the norm)
    int dev = 0;
    PNB_CHECK_CUDA(cudaGetDevice(&dev));
    PNB_REQUIRE(dev >= 0 && dev < 64, PNB_ERR_UNSUPPORTED, "pnb_shade_forward_tc: device ordinal %d", dev);
    // interleaved (non-swizzled) operand layout: 128-byte alignment of the carve-out is sufficient
    constexpr size_t kSmemMax = 232448;   // 227 KB opt-in limit per block on sm_100
    const size_t smem_tc = sizeof(tc::Smem) + 128, smem_tc3 = sizeof(tc3::Smem) + 128, smem_cb = sizeof(cb::Smem),
                 smem_ctc = sizeof(ctc::Smem) + 128, smem_tc5 = sizeof(tc5::Smem) + 128, smem_tc6 = sizeof(tc6::Smem) + 128, smem_ctc2 = sizeof(ctc2::Smem) + 128, smem_tc7 = sizeof(tc7::Smem) + 128;
    static_assert(sizeof(tc5::Smem) + 128 <= kSmemMax, "v5 shared-memory carve-out exceeds the per-block limit");
    static_assert(sizeof(tc6::Smem) + 128 <= kSmemMax, "v6 shared-memory carve-out exceeds the per-block limit");
    static_assert(sizeof(ctc2::Smem) + 128 <= kSmemMax, "colour v2 shared-memory carve-out exceeds the per-block limit");
    static_assert(sizeof(tc7::Smem) + 128 <= kSmemMax, "v7 shared-memory carve-out exceeds the per-block limit");
    static_assert((tc3::NSTAGE & (tc3::NSTAGE - 1)) == 0 && tc3::NSTAGE == 4, "issuer assumes a 4-stage ring");
    static_assert(sizeof(tc::Smem) + 128 <= kSmemMax && sizeof(tc3::Smem) + 128 <= kSmemMax && sizeof(ctc::Smem) + 128 <= kSmemMax &&
                  sizeof(cb::Smem) <= kSmemMax, "shared-memory carve-out exceeds the sm_100 per-block limit");
    if (!configured[dev]) {
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_shade_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_shade_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc3));
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_shade_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc5));
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_shade_tc6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc6));
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_shade_tc7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc7));
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_color_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ctc));
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_color_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ctc2));
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_color_branch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cb));
        PNB_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm_of[dev], cudaDevAttrMultiProcessorCount, dev));
        configured[dev] = 1;
    }
    const int n_sm = n_sm_of[dev];
    Carver c(ws, ws_bytes);
    float* hbar = c.take<float>(((size_t)max_valid_samples + 127) / 128 * 128 * 256);   // whole 128-sample colour tiles
    float* sigma = c.take<float>((size_t)max_valid_samples);
    unsigned char* vcnt = c.take<unsigned char>((size_t)max_valid_samples + 16);
    uint32_t* sc_quads = c.take<uint32_t>(pack_sc_max(max_valid_samples));
    uint32_t* quad_first = c.take<uint32_t>((size_t)max_valid_samples + 2);
    uint32_t* vorder = c.take<uint32_t>((size_t)max_valid_samples + 2);
    unsigned char* vcntp = c.take<unsigned char>((size_t)max_valid_samples + 16);
    int* pack_cnt = c.take<int>(4);
    ShadeTcParams p;
    p.q = *q; p.pts = *pts; p.o = *opts; p.wimg = (const unsigned char*)d_packed;
    for (int l = 0; l < 4; ++l) p.bias[l] = mlp->b[l];
    p.wa = mlp->w[4];
    p.ba = mlp->b[4];
    p.hbar = hbar; p.sigma = sigma; p.hbar_cap = max_valid_samples; p.err = d_err;
    p.dbg_no_weights = (stage_mask & 64) ? 1 : 0;
    p.dbg_flags = (stage_mask >> 8) & 0xff;
    const bool packed_rows = (stage_mask & (1 << 17)) != 0;   // v7: rows packed to the valid pairs
    p.vcnt = vcntp; p.vorder = vorder; p.quad_first = quad_first; p.pack_cnt = pack_cnt;
    const bool color_v2 = (stage_mask & (1 << 16)) != 0;      // pipelined colour kernel fed by operand-format h-bar (written by the v5 / v6 pair kernels)
    PNB_REQUIRE(!color_v2 || (((stage_mask & (128 | 32)) || packed_rows) && (stage_mask & 8)), PNB_ERR_INVALID, "pnb_shade_forward_tc: colour v2 needs the v5 / v6 / v7 pair pipeline");
    p.hbar_fmt = color_v2 ? 1 : 0;
    if (stage_mask & 1) {
        if (packed_rows) {
            const int cap = max_valid_samples, n_sc = (int)pack_sc_max(cap);
            k_pack_cnt<<<(cap + 255) / 256, 256, 0, stream>>>(p.q, cap, vcnt);
            k_pack_quads<false><<<(n_sc + 127) / 128, 128, 0, stream>>>(p.q, cap, vcnt, sc_quads, quad_first, vorder, vcntp);
            k_pack_scan<<<1, 1024, 0, stream>>>(p.q, cap, sc_quads, quad_first, pack_cnt);
            k_pack_quads<true><<<(n_sc + 127) / 128, 128, 0, stream>>>(p.q, cap, vcnt, sc_quads, quad_first, vorder, vcntp);
            k_shade_tc7<<<n_sm, tc7::NTHR, smem_tc7, stream>>>(p);                              // v5 with rows packed to the valid pairs
        } else if (stage_mask & 128) k_shade_tc6<<<n_sm & ~1, tc6::NTHR, smem_tc6, stream>>>(p);     // v5 on CTA pairs (cta_group::2)
        else if (stage_mask & 32) k_shade_tc5<<<n_sm, tc5::NTHR, smem_tc5, stream>>>(p);           // TMEM ping-pong, chunk-pipelined
        else if (stage_mask & 4) k_shade_tc3<<<n_sm, tc3::NTHR, smem_tc3, stream>>>(p);   // TS-form pipeline (A in tensor memory)
        else k_shade_tc<<<n_sm, tc::NTHR, smem_tc, stream>>>(p);
    }
    ColorParams cp;
    cp.q = *q; cp.o = *opts;
    for (int i = 0; i < 4; ++i) { cp.w[i] = mlp->w[5 + i]; cp.b[i] = mlp->b[5 + i]; }
    cp.hbar = hbar; cp.sigma = sigma; cp.hbar_cap = max_valid_samples; cp.sigma_rgb = (float4*)d_sigma_rgb;
    if ((stage_mask & 2) && (stage_mask & 8)) {          // colour branch on tcgen05
        ColorTcParams ct;
        ct.q = *q; ct.o = *opts; ct.wimg = (const unsigned char*)d_packed + pack_pairs_bytes();
        for (int i = 0; i < 3; ++i) ct.bias[i] = mlp->b[5 + i];
        ct.w3t = mlp->w[8]; ct.b3 = mlp->b[8];
        ct.hbar = hbar; ct.sigma = sigma; ct.hbar_cap = max_valid_samples; ct.sigma_rgb = (float4*)d_sigma_rgb; ct.err = d_err;
        if (color_v2) k_color_tc2<<<n_sm, ctc2::NTHR, smem_ctc2, stream>>>(ct);
        else k_color_tc<<<n_sm, ctc::NTHR, smem_ctc, stream>>>(ct);
    } else if (stage_mask & 2) {
        k_color_branch<<<n_sm * 2, cb::NTHREADS, smem_cb, stream>>>(cp);
    }
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}
