// Fused gather + distance weights + positional encoding + MLP chain + K-reduction + colour branch (forward),
// fp32 CUDA-core version ("v1": the exact-fp32 path; the tcgen05 BF16x3 path lives in shade_tc.cu).
//
// Replaces, per valid sample tile, what the reference does with ~40 eager torch kernels and HBM round trips:
//   gather                     /root/reference/models/neural_points/neural_points.py:706-717
//   dists / linear weights     /root/reference/models/aggregators/point_aggregators.py:727-814, :421-429
//   viewmlp (intrp order 2)    /root/reference/models/aggregators/point_aggregators.py:488-644
//   positional_encoding        /root/reference/models/helpers/networks.py:175-190
//   raw2out_density / _color   /root/reference/models/aggregators/point_aggregators.py:262-273
// and, in k_composite, ray distances + alpha compositing + fill_invalid:
//   /root/reference/models/neural_points_volumetric_model.py:271-305, :87-123
//   /root/reference/models/rendering/diff_ray_marching.py:508-554
//
// Tile = 8 valid samples x 8 neighbour slots = 64 (sample,k) pair rows; one warp builds one sample's rows.
// Activations never leave shared memory; weights (W^T, zero padded to multiples of 16 rows) stream from L2
// through a cp.async double buffer.
#include "common.cuh"

namespace pnb {

constexpr int TS = 8;             // samples per tile
constexpr int TR = TS * PNB_MAX_K;  // 64 pair rows
constexpr int XS = 292;           // X row stride (floats): 284 inputs padded to 288, +4 against bank conflicts
constexpr int HS = 284;           // H row stride: holds 280 colour-branch inputs (256 + 24) and 272 block3 inputs
constexpr int KC = 16;            // weight rows per pipeline stage
constexpr int NTHREADS = 256;
constexpr float LEAKY = 0.01f;

struct ShadeParams {
    pnb_query_t q;
    pnb_points_t pts;
    pnb_mlp_t mlp;
    pnb_shade_opts_t o;
    float4* sigma_rgb;
};

struct __align__(16) ShadeSmem {
    float X[TR * XS];
    float H[TR * HS];
    float W[2][KC * 256];
    float E[TR][8];       // block3 extras: colour(3), dir - view(3), <dir,view>(1)
    float wc[TR];         // weight * conf_coefficient per pair (0 for empty slots)
    float alpha[TR];
    float view[TS][28];   // per sample: ori_viewdirs(3) + PE4 (sin 12, cos 12)
    uint32_t samp[TS];    // candidate id of each tile sample (0xffffffff = padding)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : LEAKY * x; }

// C[64 x 256] = act(A[64 x Kp] * Wt[Kp x 256] + b);  A, C in shared memory (row strides sa, sc).
// Thread (ty, tx): rows ty*4..+3, columns tx*4 + 64*j .. +3  (j = 0..3).
__device__ __forceinline__ void gemm64x256(const float* __restrict__ A, int sa, float* __restrict__ C, int sc,
                                           const float* __restrict__ Wt, const float* __restrict__ bias, int Kp,
                                           float (*Wst)[KC * 256], bool act) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    const int nchunk = Kp / KC;
    // stage 0
    {
        const float4* src = (const float4*)Wt;
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async16(&Wst[0][(tid + i * NTHREADS) * 4], src + tid + i * NTHREADS);
        cp_async_commit();
    }
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) {
            const float4* src = (const float4*)(Wt + (size_t)(c + 1) * KC * 256);
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async16(&Wst[(c + 1) & 1][(tid + i * NTHREADS) * 4], src + tid + i * NTHREADS);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
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
                float4 w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = *(const float4*)&Wc[(k4 + kk) * 256 + tx * 4 + 64 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[i][j * 4 + 0] = fmaf(av, w[j].x, acc[i][j * 4 + 0]);
                        acc[i][j * 4 + 1] = fmaf(av, w[j].y, acc[i][j * 4 + 1]);
                        acc[i][j * 4 + 2] = fmaf(av, w[j].z, acc[i][j * 4 + 2]);
                        acc[i][j * 4 + 3] = fmaf(av, w[j].w, acc[i][j * 4 + 3]);
                    }
                }
            }
        }
        __syncthreads();
    }
    // epilogue (A may alias nothing written here: C is the other buffer)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 b = *(const float4*)&bias[tx * 4 + 64 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 v;
            v.x = acc[i][j * 4 + 0] + b.x; v.y = acc[i][j * 4 + 1] + b.y;
            v.z = acc[i][j * 4 + 2] + b.z; v.w = acc[i][j * 4 + 3] + b.w;
            if (act) { v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w); }
            *(float4*)&C[(ty * 4 + i) * sc + tx * 4 + 64 * j] = v;
        }
    }
    __syncthreads();
}

// Small dense layer for the colour branch: out[s][n] = act(sum_k in[s][k] * Wt[k][n] + b[n]),  s < 8, n < Nout.
__device__ __forceinline__ void dense8(const float* __restrict__ in, int sin_, float* __restrict__ out, int sout,
                                       const float* __restrict__ Wt, const float* __restrict__ bias, int Kin, int Nout,
                                       bool act) {
    const int tid = threadIdx.x;
    // 256 threads: n = tid % 128, sample half = tid / 128 (4 samples each)
    const int n = tid & 127, s0 = (tid >> 7) * 4;
    if (n < Nout) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int k = 0; k < Kin; ++k) {
            float w = __ldg(&Wt[(size_t)k * Nout + n]);
            a0 = fmaf(in[(s0 + 0) * sin_ + k], w, a0);
            a1 = fmaf(in[(s0 + 1) * sin_ + k], w, a1);
            a2 = fmaf(in[(s0 + 2) * sin_ + k], w, a2);
            a3 = fmaf(in[(s0 + 3) * sin_ + k], w, a3);
        }
        float b = __ldg(&bias[n]);
        a0 += b; a1 += b; a2 += b; a3 += b;
        if (act) { a0 = lrelu(a0); a1 = lrelu(a1); a2 = lrelu(a2); a3 = lrelu(a3); }
        out[(s0 + 0) * sout + n] = a0; out[(s0 + 1) * sout + n] = a1;
        out[(s0 + 2) * sout + n] = a2; out[(s0 + 3) * sout + n] = a3;
    }
    __syncthreads();
}

__device__ __forceinline__ void rot3(const float* M, float x, float y, float z, float& ox, float& oy, float& oz) {
    // v @ M^T  (row vector times transpose) == M * v
    ox = x * M[0] + y * M[1] + z * M[2];
    oy = x * M[3] + y * M[4] + z * M[5];
    oz = x * M[6] + y * M[7] + z * M[8];
}

__device__ __forceinline__ void w2pers(const pnb_shade_opts_t& o, float px, float py, float pz, float& xp, float& yp,
                                       float& zp) {
    // neural_points.py:604-610 : xyz_c[j] = sum_i shift[i] * camrotc2w[i][j]
    float sx = px - o.campos[0], sy = py - o.campos[1], sz = pz - o.campos[2];
    const float* M = o.camrotc2w;
    float xc = sx * M[0] + sy * M[3] + sz * M[6];
    float yc = sx * M[1] + sy * M[4] + sz * M[7];
    float zc = sx * M[2] + sy * M[5] + sz * M[8];
    xp = xc / zc; yp = yc / zc; zp = zc;
}

__global__ void __launch_bounds__(NTHREADS, 1) k_shade_fwd(ShadeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ShadeSmem& sm = *reinterpret_cast<ShadeSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_valid = p.q.counters[PNB_QC_N_VALID];
    const int n_tiles = (n_valid + TS - 1) / TS;
    const pnb_query_t& q = p.q;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ------------------------------------------------------------------ phase A: inputs (warp = sample)
        {
            const int vi = tile * TS + warp;
            const int k = lane >> 2, part = lane & 3;  // neighbour slot, quarter of the feature vector
            const int row = warp * PNB_MAX_K + k;
            uint32_t s = 0xffffffffu;
            int pidx = -1;
            float lx = 0.f, ly = 0.f, lz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
            if (vi < n_valid) {
                s = q.valid_list[vi];
                uint32_t pk = q.samp_ray[s];
                int r = (int)(pk >> 7), j = (int)(pk & 127u);
                int d = q.steps[(size_t)r * q.SR + j];
                float t = q.t[(size_t)r * q.t_ray_stride + d];
                vx = q.raydir[3 * r]; vy = q.raydir[3 * r + 1]; vz = q.raydir[3 * r + 2];
                lx = raypos1(q.campos[0], vx, t); ly = raypos1(q.campos[1], vy, t); lz = raypos1(q.campos[2], vz, t);
                if (k < q.K) pidx = q.cand_pidx[(size_t)s * q.K + k];
            }
            if (lane == 0) sm.samp[warp] = s;
            const bool valid = pidx >= 0;
            const int pi = valid ? pidx : 0;
            // view direction in the points' canonical frame + its encoding (per sample; lanes 0..11 do the PE)
            float ovx, ovy, ovz;
            rot3(p.o.Rw2c, vx, vy, vz, ovx, ovy, ovz);
            if (lane < 3) sm.view[warp][lane] = lane == 0 ? ovx : lane == 1 ? ovy : ovz;
            if (lane < 12) {  // index d*4 + j  (networks.py:183), ori=True layout: [p, sin(12), cos(12)]
                int dd = lane >> 2, jj = lane & 3;
                float v = (dd == 0 ? ovx : dd == 1 ? ovy : ovz) * (float)(1 << jj);
                float sn, cs;
                sincosf(v, &sn, &cs);
                sm.view[warp][3 + lane] = sn;
                sm.view[warp][15 + lane] = cs;
            }
            // geometry of this pair
            float px = __ldg(&p.pts.xyz[3 * pi]), py = __ldg(&p.pts.xyz[3 * pi + 1]), pz = __ldg(&p.pts.xyz[3 * pi + 2]);
            float dist[6];
            dist[0] = px - lx; dist[1] = py - ly; dist[2] = pz - lz;
            float xpp, ypp, zpp, xsp, ysp, zsp;
            w2pers(p.o, px, py, pz, xpp, ypp, zpp);
            w2pers(p.o, lx, ly, lz, xsp, ysp, zsp);
            dist[3] = xpp * zpp - xsp * zsp;  // point_aggregators.py:782-784
            dist[4] = ypp * zpp - ysp * zsp;
            dist[5] = zpp - zsp;
            float nrm = sqrtf(dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2]);
            float w = valid ? 1.0f / fmaxf(nrm, 1e-6f) : 0.f;       // linear :424, :427
            float wsum = w;                                          // sum over the 8 slots (lanes with part == 0)
            wsum = part == 0 ? wsum : 0.f;
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
            w = w / fmaxf(wsum, 1e-8f);                              // :801-802
            float cf = __ldg(&p.pts.conf[pi]);
            float cc = fminf(fmaxf(cf, 1e-4f), 1.0f);                // gradiant_clamp forward value :722-724
            if (part == 0) sm.wc[row] = valid ? w * cc : 0.f;
            // rotate world offsets into the canonical frame (:526)
            float d0, d1, d2;
            rot3(p.o.Rw2c, dist[0], dist[1], dist[2], d0, d1, d2);
            dist[0] = d0; dist[1] = d1; dist[2] = d2;
            float* xr = &sm.X[row * XS];
            if (valid) {
                // features: this lane's 8 channels + their PE (3 freqs): index (c*3 + j)*2 + {sin,cos}
                const float4* ep = (const float4*)&p.pts.emb[(size_t)pi * PNB_FEAT + part * 8];
                float4 f0 = __ldg(ep), f1 = __ldg(ep + 1);
                float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    int c = part * 8 + e;
                    xr[c] = f[e];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        float sn, cs;
                        sincosf(f[e] * (float)(1 << j), &sn, &cs);
                        xr[32 + (c * 3 + j) * 2] = sn;
                        xr[32 + (c * 3 + j) * 2 + 1] = cs;
                    }
                }
                // distance PE (5 freqs): 30 (d, j) pairs split over the 4 lanes
                for (int i = part; i < 30; i += 4) {
                    int dd = i / 5, jj = i - dd * 5;
                    float sn, cs;
                    sincosf(dist[dd] * (float)(1 << jj), &sn, &cs);
                    xr[224 + i * 2] = sn;
                    xr[224 + i * 2 + 1] = cs;
                }
                if (part == 0) { xr[284] = 0.f; xr[285] = 0.f; xr[286] = 0.f; xr[287] = 0.f; }
                if (part == 1) {  // block3 extras (:560-571)
                    float cr = __ldg(&p.pts.color[3 * pi]), cg = __ldg(&p.pts.color[3 * pi + 1]), cb = __ldg(&p.pts.color[3 * pi + 2]);
                    float ddx, ddy, ddz;
                    rot3(p.o.Rw2c, __ldg(&p.pts.dir[3 * pi]), __ldg(&p.pts.dir[3 * pi + 1]), __ldg(&p.pts.dir[3 * pi + 2]), ddx, ddy, ddz);
                    sm.E[row][0] = cr; sm.E[row][1] = cg; sm.E[row][2] = cb;
                    sm.E[row][3] = ddx - ovx; sm.E[row][4] = ddy - ovy; sm.E[row][5] = ddz - ovz;
                    sm.E[row][6] = ddx * ovx + ddy * ovy + ddz * ovz;
                    sm.E[row][7] = 0.f;
                }
            } else {
                for (int c = part; c < 288; c += 4) xr[c] = 0.f;
                if (part == 1)
                    for (int e = 0; e < 8; ++e) sm.E[row][e] = 0.f;
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ phase B: per-pair MLPs
        gemm64x256(sm.X, XS, sm.H, HS, p.mlp.w[0], p.mlp.b[0], 288, sm.W, true);   // block1.0  284 -> 256
        gemm64x256(sm.H, HS, sm.X, XS, p.mlp.w[1], p.mlp.b[1], 256, sm.W, true);   // block1.2  256 -> 256
        for (int i = tid; i < TR * 16; i += NTHREADS) {                             // cat extras, zero pad to 272
            int row = i >> 4, c = i & 15;
            sm.X[row * XS + 256 + c] = c < 7 ? sm.E[row][c] : 0.f;
        }
        __syncthreads();
        gemm64x256(sm.X, XS, sm.H, HS, p.mlp.w[2], p.mlp.b[2], 272, sm.W, true);   // block3.0  263 -> 256
        gemm64x256(sm.H, HS, sm.X, XS, p.mlp.w[3], p.mlp.b[3], 256, sm.W, true);   // block3.2  256 -> 256
        // alpha branch: 4 lanes per row
        {
            const int row = tid >> 2, part = tid & 3;
            float a = 0.f;
            const float* wa = p.mlp.w[4];
            for (int c = part; c < 256; c += 4) a = fmaf(sm.X[row * XS + c], __ldg(&wa[c]), a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) {
                float x = a + __ldg(&p.mlp.b[4][0]) - 1.0f;           // raw2out_density: Softplus(x - 1), threshold 20
                sm.alpha[row] = x > 20.f ? x : log1pf(expf(x));
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ phase C: K-reduction -> colour input
        for (int i = tid; i < TS * 256; i += NTHREADS) {
            int s = i >> 8, c = i & 255;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < PNB_MAX_K; ++k) acc += sm.X[(s * PNB_MAX_K + k) * XS + c] * sm.wc[s * PNB_MAX_K + k];
            sm.H[s * HS + c] = acc;
        }
        for (int i = tid; i < TS * 24; i += NTHREADS) {
            int s = i / 24, c = i - s * 24;
            sm.H[s * HS + 256 + c] = sm.view[s][3 + c];
        }
        __syncthreads();
        // ------------------------------------------------------------------ phase D: colour branch (8 samples)
        float* B0 = sm.X;            // scratch rows (stride XS)
        float* B1 = sm.X + 8 * XS;
        dense8(sm.H, HS, B0, XS, p.mlp.w[5], p.mlp.b[5], 280, 128, true);
        dense8(B0, XS, B1, XS, p.mlp.w[6], p.mlp.b[6], 128, 128, true);
        dense8(B1, XS, B0, XS, p.mlp.w[7], p.mlp.b[7], 128, 128, true);
        dense8(B0, XS, B1, XS, p.mlp.w[8], p.mlp.b[8], 128, 3, false);
        if (tid < TS) {
            uint32_t s = sm.samp[tid];
            if (s != 0xffffffffu) {
                float sg = 0.f;
                if (p.o.agg_intrp_order == 1) {
                    // agg_intrp_order 1 (point_aggregators.py:573-587): density from the K-aggregated feature (still in sm.H)
                    const float* wa = p.mlp.w[4];
                    float a = 0.f;
                    for (int c = 0; c < 256; ++c) a = fmaf(sm.H[tid * HS + c], __ldg(&wa[c]), a);
                    const float x = a + __ldg(&p.mlp.b[4][0]) - 1.0f;
                    sg = x > 20.f ? x : log1pf(expf(x));
                } else {
#pragma unroll
                    for (int k = 0; k < PNB_MAX_K; ++k) sg += sm.alpha[tid * PNB_MAX_K + k] * sm.wc[tid * PNB_MAX_K + k];
                }
                float4 o4;
                o4.x = sg;
                o4.y = 1.0f / (1.0f + expf(-B1[tid * XS + 0])) * (1.0f + 2.0f * 0.001f) - 0.001f;  // raw2out_color
                o4.z = 1.0f / (1.0f + expf(-B1[tid * XS + 1])) * (1.0f + 2.0f * 0.001f) - 0.001f;
                o4.w = 1.0f / (1.0f + expf(-B1[tid * XS + 2])) * (1.0f + 2.0f * 0.001f) - 0.001f;
                p.sigma_rgb[s] = o4;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ composite
struct CompositeParams {
    pnb_query_t q;
    pnb_shade_opts_t o;
    const float4* sigma_rgb;
    float* ray_color;
    float* opacity;
    float* bg_T;
    int8_t* ray_mask;
};

__global__ void __launch_bounds__(128) k_composite(CompositeParams p) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    const pnb_query_t& q = p.q;
    if (r >= q.R) return;
    const int SR = q.SR;
    const bool hit = q.ray_hit[r] != 0;
    if (p.ray_mask) p.ray_mask[r] = hit ? 1 : 0;
    float* op = p.opacity ? p.opacity + (size_t)r * SR : nullptr;
    if (!hit) {  // fill_invalid: background colour, zero opacity, is_background 1
        p.ray_color[3 * r] = p.o.bg_color[0]; p.ray_color[3 * r + 1] = p.o.bg_color[1]; p.ray_color[3 * r + 2] = p.o.bg_color[2];
        if (p.bg_T) p.bg_T[r] = 1.0f;
        if (op) for (int j = 0; j < SR; ++j) op[j] = 0.f;
        return;
    }
    const int n = q.nsamp[r];
    const uint32_t s0 = q.samp_off[r];
    const float dx = q.raydir[3 * r], dy = q.raydir[3 * r + 1], dz = q.raydir[3 * r + 2];
    const float* tr = q.t + (size_t)r * q.t_ray_stride;
    const float* M = p.o.camrotc2w;
    // camera-space depth of a world position (z of w2pers); unfilled slots hold the world origin
    auto zcam = [&](float x, float y, float z) {
        float sx = x - p.o.campos[0], sy = y - p.o.campos[1], sz = z - p.o.campos[2];
        return sx * M[2] + sy * M[5] + sz * M[8];
    };
    auto zslot = [&](int j) {
        if (j < n) {
            float t = tr[q.steps[(size_t)r * SR + j]];
            return zcam(raypos1(q.campos[0], dx, t), raypos1(q.campos[1], dy, t), raypos1(q.campos[2], dz, t));
        }
        return zcam(0.f, 0.f, 0.f);
    };
    const float vz = p.o.vsize_z;
    float cm = zslot(0);  // cummax so far
    float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f;
    for (int j = 0; j < SR; ++j) {
        float rd;
        if (j + 1 < SR) {
            float zn = zslot(j + 1);
            float cmn = fmaxf(cm, zn);
            rd = cmn - cm;
            cm = cmn;
        } else {
            rd = vz;
        }
        bool m = rd < 1e-8f;
        if (p.o.raydist_mode_unit > 0) m = m || (rd > 2.0f * vz);
        if (m) rd = vz;
        float sg = 0.f, rr = 0.f, gg = 0.f, bb = 0.f;
        bool valid = false;
        if (j < n && q.samp_nvalid[s0 + j] > 0) {
            float4 v = p.sigma_rgb[s0 + j];
            sg = v.x; rr = v.y; gg = v.z; bb = v.w;
            valid = true;
        }
        rd = valid ? rd : 0.f;
        float o1 = 1.0f - expf(-sg * rd);
        float bw = o1 * T;
        cr += rr * bw; cg += gg * bw; cb += bb * bw;
        T = T * (1.0f - o1 + 1e-10f);
        if (op) op[j] = o1;
    }
    p.ray_color[3 * r] = cr + p.o.bg_color[0] * T;
    p.ray_color[3 * r + 1] = cg + p.o.bg_color[1] * T;
    p.ray_color[3 * r + 2] = cb + p.o.bg_color[2] * T;
    if (p.bg_T) p.bg_T[r] = T;
}

// ------------------------------------------------------------------------------------------ auxiliary training outputs
// `weight`, `conf_coefficient`, `blend_weight` of the reference's output dict (neural_points_volumetric_model.py:325-329 with
// point_aggregators.py:421-429, 727-732 and neural_points.py:706-717) in the dense [R', SR, K] layout of the R' hit rays, straight from
// the sample-compacted query: one thread per (hit ray, sample slot).  Unfilled slots follow the reference: position 0, indices -1 ->
// weight 0, and conf_coefficient = clamp(points_conf[0]) (the reference gathers with the index clamped to 0, neural_points.py:707).
struct AuxParams {
    pnb_query_t q;
    const float* xyz;
    const float* conf;
    const long long* rows;     // [n_rows] ray ids of the hit rays, ascending
    int n_rows;
    const float* opacity;      // [R, SR] as written by k_composite
    float* weight;             // [n_rows, SR, K]
    float* cc;                 // [n_rows, SR, K]
    float* blend;              // [n_rows, SR]
    const float* grad_cc;      // backward: [n_rows, SR, K]
    float* grad_conf;          // backward: [N], accumulated into
};

__global__ void __launch_bounds__(128) k_aux_outputs(AuxParams p) {
    const pnb_query_t& q = p.q;
    const int SR = q.SR, K = q.K;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.n_rows * SR) return;
    const int rr = (int)(idx / SR), j = (int)(idx - (long long)rr * SR);
    const int r = (int)p.rows[rr];
    const int n = q.nsamp[r];
    float lx = 0.f, ly = 0.f, lz = 0.f;
    const int32_t* pid = nullptr;
    if (j < n) {
        const float t = q.t[(size_t)r * q.t_ray_stride + q.steps[(size_t)r * SR + j]];
        lx = raypos1(q.campos[0], q.raydir[3 * r], t); ly = raypos1(q.campos[1], q.raydir[3 * r + 1], t); lz = raypos1(q.campos[2], q.raydir[3 * r + 2], t);
        pid = q.cand_pidx + (size_t)(q.samp_off[r] + j) * K;
    }
    float w[PNB_MAX_K], wsum = 0.f;
    const float c_first = fminf(fmaxf(__ldg(p.conf), 1e-4f), 1.0f);
    float* cc = p.cc + (size_t)idx * K;
    for (int k = 0; k < K; ++k) {
        const int pi = pid ? pid[k] : -1;
        w[k] = 0.f;
        float c = c_first;
        if (pi >= 0) {
            const float dx = __ldg(&p.xyz[3 * pi]) - lx, dy = __ldg(&p.xyz[3 * pi + 1]) - ly, dz = __ldg(&p.xyz[3 * pi + 2]) - lz;
            w[k] = 1.0f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-6f);
            c = fminf(fmaxf(__ldg(&p.conf[pi]), 1e-4f), 1.0f);
        }
        wsum += w[k];
        cc[k] = c;
    }
    const float inv = fmaxf(wsum, 1e-8f);
    float* wo = p.weight + (size_t)idx * K;
    for (int k = 0; k < K; ++k) wo[k] = w[k] / inv;
    // blend_weight = opacity * exclusive cumprod(1 - opacity + 1e-10)   (diff_ray_marching.py:536-541)
    const float* op = p.opacity + (size_t)r * SR;
    float T = 1.0f;
    for (int i = 0; i < j; ++i) T *= (1.0f - op[i] + 1e-10f);
    p.blend[idx] = op[j] * T;
}

// d loss / d points_conf through conf_coefficient: the clamp is a straight-through estimator (value clamp(c), gradient 1,
// neural_points.py:713), so every entry adds its gradient to the conf of its (clamped) index; the entries of empty slots all land on
// point 0: summed per warp first, one atomic per warp.
__global__ void __launch_bounds__(128) k_aux_conf_bwd(AuxParams p) {
    const pnb_query_t& q = p.q;
    const int SR = q.SR, K = q.K;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float g0 = 0.f;
    if (idx < (long long)p.n_rows * SR) {
        const int rr = (int)(idx / SR), j = (int)(idx - (long long)rr * SR);
        const int r = (int)p.rows[rr];
        const int32_t* pid = j < q.nsamp[r] ? q.cand_pidx + (size_t)(q.samp_off[r] + j) * K : nullptr;
        const float* g = p.grad_cc + (size_t)idx * K;
        for (int k = 0; k < K; ++k) {
            const int pi = pid ? pid[k] : -1;
            if (pi > 0) atomicAdd(&p.grad_conf[pi], g[k]);
            else g0 += g[k];
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) g0 += __shfl_xor_sync(0xffffffffu, g0, d);
    if ((threadIdx.x & 31) == 0 && g0 != 0.f) atomicAdd(&p.grad_conf[0], g0);
}

}  // namespace pnb

using namespace pnb;

extern "C" size_t pnb_shade_bytes(int cap_samples) {
    (void)cap_samples;
    return 256;
}

extern "C" int pnb_shade_forward(const pnb_query_t* q, const pnb_points_t* pts, const pnb_mlp_t* mlp,
                                 const pnb_shade_opts_t* opts, float* d_sigma_rgb, void* ws, size_t ws_bytes,
                                 pnb_stream_t stream_) {
    (void)ws; (void)ws_bytes;
    cudaStream_t stream = (cudaStream_t)stream_;
    PNB_REQUIRE(q && pts && mlp && opts && d_sigma_rgb, PNB_ERR_INVALID, "pnb_shade_forward: null argument");
    PNB_REQUIRE(q->K >= 1 && q->K <= PNB_MAX_K, PNB_ERR_UNSUPPORTED, "pnb_shade_forward: K=%d unsupported", q->K);
    for (int i = 0; i < 9; ++i)
        PNB_REQUIRE(mlp->w[i] && mlp->b[i], PNB_ERR_INVALID, "pnb_shade_forward: MLP tensor %d is null", i);
    static int smem_set[64] = {0}, n_sm_of[64] = {0};      // per device of this process
    int dev = 0;
    PNB_CHECK_CUDA(cudaGetDevice(&dev));
    PNB_REQUIRE(dev >= 0 && dev < 64, PNB_ERR_UNSUPPORTED, "pnb_shade_forward: device ordinal %d", dev);
    if (!smem_set[dev]) {
        PNB_CHECK_CUDA(cudaFuncSetAttribute(k_shade_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ShadeSmem)));
        PNB_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm_of[dev], cudaDevAttrMultiProcessorCount, dev));
        smem_set[dev] = 1;
    }
    const int n_sm = n_sm_of[dev];
    ShadeParams p;
    p.q = *q; p.pts = *pts; p.mlp = *mlp; p.o = *opts; p.sigma_rgb = (float4*)d_sigma_rgb;
    k_shade_fwd<<<n_sm, NTHREADS, sizeof(ShadeSmem), stream>>>(p);
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}

extern "C" int pnb_composite_forward(const pnb_query_t* q, const pnb_shade_opts_t* opts, const float* d_sigma_rgb,
                                     float* d_ray_color, float* d_opacity, float* d_bg_T, int8_t* d_ray_mask,
                                     pnb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PNB_REQUIRE(q && opts && d_sigma_rgb && d_ray_color, PNB_ERR_INVALID, "pnb_composite_forward: null argument");
    CompositeParams p;
    p.q = *q; p.o = *opts; p.sigma_rgb = (const float4*)d_sigma_rgb;
    p.ray_color = d_ray_color; p.opacity = d_opacity; p.bg_T = d_bg_T; p.ray_mask = d_ray_mask;
    k_composite<<<(q->R + 127) / 128, 128, 0, stream>>>(p);
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}

extern "C" int pnb_aux_outputs(const pnb_query_t* q, const pnb_points_t* pts, const long long* d_rows, int n_rows, const float* d_opacity,
                               float* d_weight, float* d_conf_coefficient, float* d_blend_weight, pnb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PNB_REQUIRE(q && pts && d_opacity && d_weight && d_conf_coefficient && d_blend_weight, PNB_ERR_INVALID, "pnb_aux_outputs: null argument");
    PNB_REQUIRE(q->K >= 1 && q->K <= PNB_MAX_K, PNB_ERR_UNSUPPORTED, "pnb_aux_outputs: K=%d unsupported", q->K);
    if (n_rows <= 0) return PNB_OK;
    PNB_REQUIRE(d_rows, PNB_ERR_INVALID, "pnb_aux_outputs: null row list");
    AuxParams p;
    p.q = *q; p.xyz = pts->xyz; p.conf = pts->conf; p.rows = d_rows; p.n_rows = n_rows; p.opacity = d_opacity;
    p.weight = d_weight; p.cc = d_conf_coefficient; p.blend = d_blend_weight; p.grad_cc = nullptr; p.grad_conf = nullptr;
    const long long n = (long long)n_rows * q->SR;
    k_aux_outputs<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p);
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}

extern "C" int pnb_aux_conf_backward(const pnb_query_t* q, const long long* d_rows, int n_rows, const float* d_grad_conf_coefficient,
                                     float* d_grad_conf, pnb_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PNB_REQUIRE(q && d_grad_conf_coefficient && d_grad_conf, PNB_ERR_INVALID, "pnb_aux_conf_backward: null argument");
    if (n_rows <= 0) return PNB_OK;
    PNB_REQUIRE(d_rows, PNB_ERR_INVALID, "pnb_aux_conf_backward: null row list");
    AuxParams p;
    p.q = *q; p.xyz = nullptr; p.conf = nullptr; p.rows = d_rows; p.n_rows = n_rows; p.opacity = nullptr;
    p.weight = nullptr; p.cc = nullptr; p.blend = nullptr; p.grad_cc = d_grad_conf_coefficient; p.grad_conf = d_grad_conf;
    const long long n = (long long)n_rows * q->SR;
    k_aux_conf_bwd<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p);
    PNB_CHECK_CUDA(cudaGetLastError());
    return PNB_OK;
}
