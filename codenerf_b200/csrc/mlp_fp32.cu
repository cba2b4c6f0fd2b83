// mlp_fp32.cu -- CNB_PRECISION_FP32: the code-conditioned MLP, its backward and the
// render path around it in fp32 on CUDA cores (the "within 1e-4" mode).  Layer by layer
// over row chunks with activations in a caller-provided workspace; one generic strided
// SGEMM with a fused epilogue does forward, dgrad and (split-K) wgrad.
//
// This is the accuracy mode and the on-device cross-check for the tcgen05 path
// (render_sm100.cu); throughput is not its goal.
#include "common.cuh"
#include "mlp_fp32.cuh"

// ---------------------------------------------------------------------------
// C[M,N] = epilogue( sum_k a(m,k) b(k,n) ),  a(m,k) = A[m*sam + k*sak],  b(k,n) = B[k*sbk + n*sbn]
struct GemmArgs {
    const float* A; int64_t sam, sak;
    const float* B; int64_t sbk, sbn;
    int M, N, K;
    int k_split;              // K range per blockIdx.z (multiple of BK); == K when not split
    float* C; int64_t ldc;    // primary output (may be null)
    const float* bias;        // [N] added before relu (may be null)
    int relu;
    int atomic_accumulate;    // C += via atomicAdd (split-K wgrad)
    // secondary output: C2 = primary + rowbias[code(row)][n]   (rowbias != null)
    //               or  C2 = mask[m][n] > 0 ? primary : 0       (mask != null)
    float* C2; int64_t ldc2;
    const float* rowbias; int64_t ld_rowbias;
    int64_t rows_per_code, row_offset; int n_codes;
    const float* mask; int64_t ld_mask;
};

#define BM 64
#define BN 64
#define BK 16

__global__ void __launch_bounds__(256) k_sgemm(GemmArgs g) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.k_split;
    const int kend = min(g.K, kbeg + g.k_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool a_kcontig = (g.sak == 1);
    const bool b_ncontig = (g.sbn == 1);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            int m, k;
            if (a_kcontig) { k = idx & 15; m = idx >> 4; } else { m = idx & 63; k = idx >> 6; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < g.M && gk < kend) ? __ldg(g.A + (int64_t)gm * g.sam + (int64_t)gk * g.sak) : 0.f;
            int n, kb;
            if (b_ncontig) { n = idx & 63; kb = idx >> 6; } else { kb = idx & 15; n = idx >> 4; }
            const int gn = n0 + n, gkb = k0 + kb;
            Bs[kb][n] = (gn < g.N && gkb < kend) ? __ldg(g.B + (int64_t)gkb * g.sbk + (int64_t)gn * g.sbn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        int64_t code = 0;
        if (g.rowbias && g.n_codes > 1) {
            code = (g.row_offset + m) / g.rows_per_code;
            if (code >= g.n_codes) code = g.n_codes - 1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += __ldg(g.bias + n);
            if (g.relu) v = fmaxf(v, 0.f);
            if (g.C) {
                if (g.atomic_accumulate) atomicAdd(g.C + (int64_t)m * g.ldc + n, v);
                else g.C[(int64_t)m * g.ldc + n] = v;
            }
            if (g.C2) {
                float v2;
                if (g.rowbias) v2 = v + __ldg(g.rowbias + code * g.ld_rowbias + n);
                else v2 = (__ldg(g.mask + (int64_t)m * g.ld_mask + n) > 0.f) ? v : 0.f;
                g.C2[(int64_t)m * g.ldc2 + n] = v2;
            }
        }
    }
}

static int launch_gemm(GemmArgs g, int splits, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return CNB_OK;
    if (splits < 1) splits = 1;
    int ks = (g.K + splits - 1) / splits;
    ks = ((ks + BK - 1) / BK) * BK;
    g.k_split = ks;
    splits = (g.K + ks - 1) / ks;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, splits);
    k_sgemm<<<grid, 256, 0, st>>>(g);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

// ---------------------------------------------------------------------------
// Positional encoding -- reference src/model.py:4-7.  Column order: x, sin(2^i x) for
// i<L, cos(2^i x) for i<L.  Writes PE(xyz) [m, ld_x] and PE(viewdir) into the tail
// columns of the concat buffer (model.py:46).  Source: xyz/viewdir arrays (the unfused
// API) or rays + z_vals (fused render: src/utils.py:10-19, 30 evaluated in-kernel).
__global__ void k_encode_inputs(const float* __restrict__ xyz, const float* __restrict__ viewdir, CnbRaySource rs,
                                int64_t row0, int m, int Lx, int Ld,
                                float* __restrict__ pe_x, int ld_x, float* __restrict__ pe_d, int ld_d) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int64_t s = row0 + r;
    float p[3], v[3];
    if (xyz) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { p[k] = __ldg(xyz + s * 3 + k); v[k] = __ldg(viewdir + s * 3 + k); }
    } else {
        const int64_t ray = s / rs.N;
        const int zi = (int)(s - ray * rs.N);
        float o[3];
        cnb_fetch_ray(rs, ray, o, v);
        const int64_t seg = ray / rs.rays_per_segment;
        const float z = __ldg(rs.z_vals + (rs.z_per_segment ? seg * rs.N : 0) + zi);
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = cnb_sample_coord(o[k], v[k], z);
    }
    float* ox = pe_x + (int64_t)r * ld_x;
    float* od = pe_d + (int64_t)r * ld_d;
#pragma unroll
    for (int k = 0; k < 3; ++k) { ox[k] = p[k]; od[k] = v[k]; }
    for (int i = 0; i < Lx; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float a = ldexpf(p[k], i);
            ox[3 + 3 * i + k] = sinf(a);
            ox[3 + 3 * Lx + 3 * i + k] = cosf(a);
        }
    for (int i = 0; i < Ld; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float a = ldexpf(v[k], i);
            od[3 + 3 * i + k] = sinf(a);
            od[3 + 3 * Ld + 3 * i + k] = cosf(a);
        }
}

// ---------------------------------------------------------------------------
// Per-code latent branches z = ReLU(W code + b) -- src/model.py:41, :49.  One warp per output.
__global__ void k_latent_fwd(const float* __restrict__ Wl, const float* __restrict__ bl,
                             const float* __restrict__ codes, int n_codes, int LD, int W,
                             float* __restrict__ z /*[n_codes, ldz]*/, int64_t ldz) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= n_codes * W) return;
    const int o = gw / W, n = gw - o * W;
    float acc = 0.f;
    for (int i = lane; i < LD; i += 32) acc = fmaf(__ldg(Wl + (int64_t)n * LD + i), __ldg(codes + (int64_t)o * LD + i), acc);
    acc = warp_sum_f(acc);
    if (lane == 0) z[(int64_t)o * ldz + n] = fmaxf(acc + __ldg(bl + n), 0.f);
}

// Backward of one latent branch for all codes: dpre = dz * (z > 0);
// dW += dpre (x) code, db += dpre, dcode += W^T dpre.   grid: (W, n_codes), block LD-strided.
__global__ void k_latent_bwd(const float* __restrict__ Wl, const float* __restrict__ codes,
                             const float* __restrict__ z, const float* __restrict__ dz, int64_t ldz,
                             int n_codes, int LD, int W,
                             float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dcodes) {
    const int n = blockIdx.x, o = blockIdx.y;
    const float zz = z[(int64_t)o * ldz + n];
    const float v = zz > 0.f ? dz[(int64_t)o * ldz + n] : 0.f;
    if (v == 0.f) return;
    if (threadIdx.x == 0 && db) atomicAdd(db + n, v);
    for (int i = threadIdx.x; i < LD; i += blockDim.x) {
        if (dW) atomicAdd(dW + (int64_t)n * LD + i, v * __ldg(codes + (int64_t)o * LD + i));
        atomicAdd(dcodes + (int64_t)o * LD + i, v * __ldg(Wl + (int64_t)n * LD + i));
    }
}

// ---------------------------------------------------------------------------
// Narrow heads (sigma: 1 output, rgb.2: 3 outputs) -- src/model.py:45, :52.  One warp per row.
__global__ void k_head_fwd(const float* __restrict__ A, int64_t lda, int K, const float* __restrict__ Wh,
                           const float* __restrict__ bh, int n_out, int m, int softplus,
                           float* __restrict__ pre /*[m,n_out] or null*/, float* __restrict__ out /*[m,n_out]*/) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= m) return;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int i = lane; i < K; i += 32) {
        const float a = A[(int64_t)r * lda + i];
        for (int o = 0; o < n_out; ++o) acc[o] = fmaf(a, __ldg(Wh + (int64_t)o * K + i), acc[o]);
    }
    for (int o = 0; o < n_out; ++o) {
        float v = warp_sum_f(acc[o]);
        if (lane == 0) {
            v += __ldg(bh + o);
            if (pre) pre[(int64_t)r * n_out + o] = v;
            out[(int64_t)r * n_out + o] = softplus ? cnb_softplus(v) : v;
        }
    }
}

// sigma head backward: dpre = dsigma * softplus'(pre) (ATen: z/(z+1), pass-through above 20),
// df[row, :] += dpre * w_sigma.
__global__ void k_sigma_bwd(const float* __restrict__ d_sigma, const float* __restrict__ pre,
                            const float* __restrict__ ws, int m, int W, float* __restrict__ dpre_out,
                            float* __restrict__ df, int64_t ldf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)m * W) return;
    const int r = (int)(i / W), c = (int)(i - (int64_t)r * W);
    const float x = pre[r];
    const float e = expf(x);
    const float dp = x > 20.f ? d_sigma[r] : d_sigma[r] * e / (e + 1.f);
    if (c == 0) dpre_out[r] = dp;
    df[(int64_t)r * ldf + c] += dp * __ldg(ws + c);
}

// Column sums of X[m, n] (+ optional per-code grouping) accumulated with atomics.
//   out[code(row)*ld_out + col] += sum_rows X
__global__ void k_colsum(const float* __restrict__ X, int64_t ldx, int m, int n, int64_t row_offset,
                         int64_t rows_per_code, int n_codes, float* __restrict__ out, int64_t ld_out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int rows_per_block = (m + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
    if (col >= n || r0 >= r1) return;
    int64_t cur = -1; float acc = 0.f;
    for (int r = r0; r < r1; ++r) {
        int64_t code = 0;
        if (n_codes > 1) { code = (row_offset + r) / rows_per_code; if (code >= n_codes) code = n_codes - 1; }
        if (code != cur) {
            if (cur >= 0) atomicAdd(out + cur * ld_out + col, acc);
            cur = code; acc = 0.f;
        }
        acc += X[(int64_t)r * ldx + col];
    }
    if (cur >= 0) atomicAdd(out + cur * ld_out + col, acc);
}

static int launch_colsum(const float* X, int64_t ldx, int m, int n, int64_t row_offset, int64_t rows_per_code,
                         int n_codes, float* out, int64_t ld_out, cudaStream_t st) {
    if (m <= 0 || n <= 0) return CNB_OK;
    dim3 grid((n + 127) / 128, min(64, (m + 255) / 256));
    k_colsum<<<grid, 128, 0, st>>>(X, ldx, m, n, row_offset, rows_per_code, n_codes, out, ld_out);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

// ---------------------------------------------------------------------------
// Composite helpers used by the fused render API in fp32 mode: the standalone
// kernels of rays.cu with per-segment z rows.
extern "C" int cnb_volume_rendering_forward(const float*, const float*, const float*, int64_t, int, int, float*,
                                            float*, float*, cnb_stream_t);
extern "C" int cnb_volume_rendering_backward(const float*, const float*, const float*, int64_t, int, int,
                                             const float*, const float*, float*, float*, cnb_stream_t);

// d_rgb seed of the fused training step: d_rgb = 2 (rgb - target) / (3 rays_per_segment) * scale
// (loss_l2 = mean over the segment, src/trainer.py:75) and per-segment squared error.
__global__ void k_l2_seed(const float* __restrict__ rgb, const float* __restrict__ target, int64_t n_rays,
                          int64_t ray0, int rays_per_segment, float scale, float* __restrict__ d_rgb,
                          float* __restrict__ sq_err /*[n_segments] or null*/) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float e2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float e = rgb[r * 3 + k] - target[r * 3 + k];
        d_rgb[r * 3 + k] = 2.f * e / (3.f * (float)rays_per_segment) * scale;
        e2 += e * e;
    }
    if (sq_err) atomicAdd(sq_err + (ray0 + r) / rays_per_segment, e2);
}

// Launch wrappers shared with the tensor-core path (render_sm100.cu).
int cnb_launch_latent_fwd(const float* Wl, const float* bl, const float* codes, int n_codes, int LD, int W,
                          float* z, int64_t ldz, cudaStream_t st) {
    const int threads = 256;
    const int blocks = (n_codes * W * 32 + threads - 1) / threads;
    k_latent_fwd<<<blocks, threads, 0, st>>>(Wl, bl, codes, n_codes, LD, W, z, ldz);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}
int cnb_launch_latent_bwd(const float* Wl, const float* codes, const float* z, const float* dz, int64_t ldz,
                          int n_codes, int LD, int W, float* dW, float* db, float* dcodes, cudaStream_t st) {
    dim3 grid(W, n_codes);
    k_latent_bwd<<<grid, 128, 0, st>>>(Wl, codes, z, dz, ldz, n_codes, LD, W, dW, db, dcodes);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

// ===========================================================================
// Host orchestration

struct Fp32Buffers {
    float *pe_x, *cat, *spre, *sig, *rgbs, *r1;
    float *y[CNB_MAX_BLOCKS + 1], *u[CNB_MAX_BLOCKS];
    float *h[CNB_MAX_BLOCKS + 1], *ut[CNB_MAX_BLOCKS];
    float *ga, *gb, *graw, *dsig, *drgbs, *dspre;
    float *zs, *zt, *dzs, *dzt;            // [n_codes, blocks*W]
    float *ray_rgb, *ray_depth, *ray_acc, *ray_drgb;   // fused render (per chunk rays)
    int ld_x, ld_cat;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

int64_t cnb_fp32_chunk_rows(int N) {
    int64_t rows = 32768;
    if (N > 0) { int64_t rays = rows / N; if (rays < 1) rays = 1; rows = rays * N; }
    return rows;
}

static size_t carve(const cnb_net_config* c, const CnbLayout& L, int64_t rows, int n_codes, int backward,
                    int64_t chunk_rays, char* base, Fp32Buffers* B) {
    size_t off = 0;
    const int W = c->W;
    auto take = [&](size_t nfloats) -> float* {
        float* p = base ? (float*)(base + off) : nullptr;
        off += align_up(nfloats * sizeof(float));
        return p;
    };
    Fp32Buffers b = {};
    b.ld_x = (L.d_xyz + 3) & ~3;
    b.ld_cat = W + L.d_dir;
    b.pe_x = take((size_t)rows * b.ld_x);
    b.cat = take((size_t)rows * b.ld_cat);
    b.spre = take(rows); b.sig = take(rows); b.rgbs = take((size_t)rows * 3); b.r1 = take((size_t)rows * (W / 2));
    for (int j = 0; j <= c->shape_blocks; ++j) b.y[j] = take((size_t)rows * W);
    for (int j = 0; j < c->shape_blocks; ++j) b.u[j] = take((size_t)rows * W);
    for (int j = 0; j <= c->texture_blocks; ++j) b.h[j] = take((size_t)rows * W);
    for (int j = 0; j < c->texture_blocks; ++j) b.ut[j] = take((size_t)rows * W);
    b.zs = take((size_t)n_codes * c->shape_blocks * W);
    b.zt = take((size_t)n_codes * c->texture_blocks * W);
    if (backward) {
        b.ga = take((size_t)rows * b.ld_cat); b.gb = take((size_t)rows * b.ld_cat); b.graw = take((size_t)rows * W);
        b.dsig = take(rows); b.drgbs = take((size_t)rows * 3); b.dspre = take(rows);
        b.dzs = take((size_t)n_codes * c->shape_blocks * W);
        b.dzt = take((size_t)n_codes * c->texture_blocks * W);
    }
    if (chunk_rays > 0) {
        b.ray_rgb = take((size_t)chunk_rays * 3); b.ray_depth = take(chunk_rays); b.ray_acc = take(chunk_rays);
        b.ray_drgb = take((size_t)chunk_rays * 3);
    }
    if (B) *B = b;
    return off;
}

size_t cnb_fp32_workspace_bytes(const cnb_net_config* c, int64_t S, int N, int n_codes, int backward, int fused) {
    CnbLayout L; cnb_make_layout(c, &L);
    int64_t rows = cnb_fp32_chunk_rows(N);
    if (S < rows) rows = S;
    return carve(c, L, rows, n_codes, backward, fused ? (N > 0 ? rows / N + 1 : 0) : 0, nullptr, nullptr);
}

struct Fp32Ctx {
    const cnb_net_config* c; CnbLayout L; const float* const* P; cudaStream_t st;
    Fp32Buffers B; int n_codes; int64_t rows_per_code;
    const float* w(int i) const { return P[i]; }
    const float* b(int i) const { return P[i + 1]; }
};

static int latent_forward(Fp32Ctx& x, const float* shape_codes, const float* tex_codes) {
    const int W = x.c->W, LD = x.c->latent_dim;
    const int threads = 256;
    const int blocks = (x.n_codes * W * 32 + threads - 1) / threads;
    for (int j = 0; j < x.c->shape_blocks; ++j) {
        k_latent_fwd<<<blocks, threads, 0, x.st>>>(x.w(x.L.i_sl[j]), x.b(x.L.i_sl[j]), shape_codes, x.n_codes, LD, W,
                                                  x.B.zs + (size_t)j * W, (int64_t)x.c->shape_blocks * W);
        CNB_LAUNCH_CHECK();
    }
    for (int j = 0; j < x.c->texture_blocks; ++j) {
        k_latent_fwd<<<blocks, threads, 0, x.st>>>(x.w(x.L.i_tl[j]), x.b(x.L.i_tl[j]), tex_codes, x.n_codes, LD, W,
                                                  x.B.zt + (size_t)j * W, (int64_t)x.c->texture_blocks * W);
        CNB_LAUNCH_CHECK();
    }
    return CNB_OK;
}

// y = act(A W^T + b) (+ second output y + z_next[code]).
static int fwd_layer(Fp32Ctx& x, const float* A, int64_t lda, int K, int wi, int Nout, int relu, float* C, int64_t ldc,
                     float* C2, const float* zrow, int64_t ldz, int m, int64_t row0) {
    GemmArgs g = {};
    g.A = A; g.sam = lda; g.sak = 1;
    g.B = x.w(wi); g.sbk = 1; g.sbn = K;      // b(k,n) = W[n*K + k]
    g.M = m; g.N = Nout; g.K = K;
    g.C = C; g.ldc = ldc; g.bias = x.b(wi); g.relu = relu;
    if (C2) {
        g.C2 = C2; g.ldc2 = ldc; g.rowbias = zrow; g.ld_rowbias = ldz;
        g.rows_per_code = x.rows_per_code; g.row_offset = row0; g.n_codes = x.n_codes;
    }
    return launch_gemm(g, 1, x.st);
}

// Forward of one chunk of m rows starting at global row row0 -- src/model.py:36-53.
static int chunk_forward(Fp32Ctx& x, const float* xyz, const float* viewdir, const CnbRaySource* rs, int64_t row0,
                         int m) {
    const cnb_net_config* c = x.c; const CnbLayout& L = x.L; Fp32Buffers& B = x.B;
    const int W = c->W, SB = c->shape_blocks, TB = c->texture_blocks;
    CnbRaySource src = rs ? *rs : CnbRaySource{};
    k_encode_inputs<<<(m + 127) / 128, 128, 0, x.st>>>(xyz, viewdir, src, row0, m, c->num_xyz_freq, c->num_dir_freq,
                                                     B.pe_x, B.ld_x, B.cat + W, B.ld_cat);
    CNB_LAUNCH_CHECK();
    const int64_t ldzs = (int64_t)SB * W, ldzt = (int64_t)TB * W;
    CNB_TRY(fwd_layer(x, B.pe_x, B.ld_x, L.d_xyz, L.i_enc_xyz, W, 1, B.y[0], W, B.u[0], B.zs, ldzs, m, row0));
    for (int j = 0; j < SB; ++j)
        CNB_TRY(fwd_layer(x, B.u[j], W, W, L.i_s[j], W, 1, B.y[j + 1], W, j + 1 < SB ? B.u[j + 1] : nullptr,
                          B.zs + (size_t)(j + 1) * W, ldzs, m, row0));
    CNB_TRY(fwd_layer(x, B.y[SB], W, W, L.i_enc_shape, W, 0, B.cat, B.ld_cat, nullptr, nullptr, 0, m, row0));
    k_head_fwd<<<(m * 32 + 255) / 256, 256, 0, x.st>>>(B.cat, B.ld_cat, W, x.w(L.i_sigma), x.b(L.i_sigma), 1, m, 1,
                                                     B.spre, B.sig);
    CNB_LAUNCH_CHECK();
    CNB_TRY(fwd_layer(x, B.cat, B.ld_cat, W + L.d_dir, L.i_enc_vd, W, 1, B.h[0], W, B.ut[0], B.zt, ldzt, m, row0));
    for (int j = 0; j < TB; ++j)
        CNB_TRY(fwd_layer(x, B.ut[j], W, W, L.i_t[j], W, 1, B.h[j + 1], W, j + 1 < TB ? B.ut[j + 1] : nullptr,
                          B.zt + (size_t)(j + 1) * W, ldzt, m, row0));
    CNB_TRY(fwd_layer(x, B.h[TB], W, W, L.i_rgb0, W / 2, 1, B.r1, W / 2, nullptr, nullptr, 0, m, row0));
    k_head_fwd<<<(m * 32 + 255) / 256, 256, 0, x.st>>>(B.r1, W / 2, W / 2, x.w(L.i_rgb2), x.b(L.i_rgb2), 3, m, 0,
                                                     nullptr, B.rgbs);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
}

// dW[wi] += dY^T A_in, db += colsum(dY).
static int wgrad(Fp32Ctx& x, float* dP, int64_t w_off, int64_t b_off, const float* dY, int64_t ldy, int Nout,
                 const float* Ain, int64_t lda, int Kin, int m) {
    if (!dP) return CNB_OK;
    GemmArgs g = {};
    g.A = dY; g.sam = 1; g.sak = ldy;          // a(n_out, row) = dY[row*ldy + n_out]
    g.B = Ain; g.sbk = lda; g.sbn = 1;         // b(row, k_in) = Ain[row*lda + k_in]
    g.M = Nout; g.N = Kin; g.K = m;
    g.C = dP + w_off; g.ldc = Kin; g.atomic_accumulate = 1;
    const int tiles = ((Nout + BM - 1) / BM) * ((Kin + BN - 1) / BN);
    int splits = (296 + tiles - 1) / tiles;
    CNB_TRY(launch_gemm(g, splits, x.st));
    return launch_colsum(dY, ldy, m, Nout, 0, 1, 1, dP + b_off, 0, x.st);
}

// raw = dY W (input gradient), masked = raw * (mask > 0).
static int dgrad(Fp32Ctx& x, const float* dY, int64_t ldy, int Nout, int wi, int ldw, int Kin, float* raw,
                 int64_t ldraw, float* masked, int64_t ldmasked, const float* mask, int64_t ldmask, int m) {
    GemmArgs g = {};
    g.A = dY; g.sam = ldy; g.sak = 1;
    g.B = x.w(wi); g.sbk = ldw; g.sbn = 1;     // b(n_out, k_in) = W[n_out*ldw + k_in]
    g.M = m; g.N = Kin; g.K = Nout;
    g.C = raw; g.ldc = ldraw;
    if (masked) { g.C2 = masked; g.ldc2 = ldmasked; g.mask = mask; g.ld_mask = ldmask; }
    return launch_gemm(g, 1, x.st);
}

// Backward of one chunk (forward activations of the same chunk are in the workspace).
// d_sig [m], d_rgbs [m,3] are the seeds.
static int chunk_backward(Fp32Ctx& x, const float* d_sig, const float* d_rgbs, float* dP, int64_t row0, int m) {
    const cnb_net_config* c = x.c; const CnbLayout& L = x.L; Fp32Buffers& B = x.B;
    const int W = c->W, SB = c->shape_blocks, TB = c->texture_blocks, H2 = W / 2;
    const int64_t ldzs = (int64_t)SB * W, ldzt = (int64_t)TB * W;
    float* ga = B.ga; float* gb = B.gb;
    // rgb.2 : dpre(r1) = (d_rgb W2) * (r1 > 0)
    CNB_TRY(wgrad(x, dP, L.rgb2_w, L.rgb2_b, d_rgbs, 3, 3, B.r1, H2, H2, m));
    CNB_TRY(dgrad(x, d_rgbs, 3, 3, L.i_rgb2, H2, H2, nullptr, 0, ga, H2, B.r1, H2, m));
    // rgb.0
    CNB_TRY(wgrad(x, dP, L.rgb0_w, L.rgb0_b, ga, H2, H2, B.h[TB], W, W, m));
    CNB_TRY(dgrad(x, ga, H2, H2, L.i_rgb0, W, W, nullptr, 0, gb, W, B.h[TB], W, m));
    // texture blocks: gb = dpre of texture_layer_{j+1}
    for (int j = TB - 1; j >= 0; --j) {
        CNB_TRY(wgrad(x, dP, L.t_w[j], L.t_b[j], gb, W, W, B.ut[j], W, W, m));
        CNB_TRY(dgrad(x, gb, W, W, L.i_t[j], W, W, B.graw, W, ga, W, B.h[j], W, m));
        CNB_TRY(launch_colsum(B.graw, W, m, W, row0, x.rows_per_code, x.n_codes, B.dzt + (size_t)j * W, ldzt, x.st));
        float* t = ga; ga = gb; gb = t;
    }
    // encoding_viewdir: gb = dpre(h0); raw df into ga (first W input columns only: viewdir needs no grad)
    CNB_TRY(wgrad(x, dP, L.enc_vd_w, L.enc_vd_b, gb, W, W, B.cat, B.ld_cat, W + L.d_dir, m));
    CNB_TRY(dgrad(x, gb, W, W, L.i_enc_vd, W + L.d_dir, W, ga, W, nullptr, 0, nullptr, 0, m));
    // sigma head: df += dpre_sigma * w_sigma
    k_sigma_bwd<<<(unsigned)(((int64_t)m * W + 255) / 256), 256, 0, x.st>>>(d_sig, B.spre, x.w(L.i_sigma), m, W,
                                                                          B.dspre, ga, W);
    CNB_LAUNCH_CHECK();
    CNB_TRY(wgrad(x, dP, L.sigma_w, L.sigma_b, B.dspre, 1, 1, B.cat, B.ld_cat, W, m));
    // encoding_shape (no activation): dY = df = ga
    CNB_TRY(wgrad(x, dP, L.enc_shape_w, L.enc_shape_b, ga, W, W, B.y[SB], W, W, m));
    CNB_TRY(dgrad(x, ga, W, W, L.i_enc_shape, W, W, nullptr, 0, gb, W, B.y[SB], W, m));
    // shape blocks: gb = dpre of shape_layer_{j+1}
    for (int j = SB - 1; j >= 0; --j) {
        CNB_TRY(wgrad(x, dP, L.s_w[j], L.s_b[j], gb, W, W, B.u[j], W, W, m));
        CNB_TRY(dgrad(x, gb, W, W, L.i_s[j], W, W, B.graw, W, ga, W, B.y[j], W, m));
        CNB_TRY(launch_colsum(B.graw, W, m, W, row0, x.rows_per_code, x.n_codes, B.dzs + (size_t)j * W, ldzs, x.st));
        float* t = ga; ga = gb; gb = t;
    }
    // encoding_xyz: gb = dpre(y0)
    CNB_TRY(wgrad(x, dP, L.enc_xyz_w, L.enc_xyz_b, gb, W, W, B.pe_x, B.ld_x, L.d_xyz, m));
    return CNB_OK;
}

static int latent_backward(Fp32Ctx& x, const float* shape_codes, const float* tex_codes, float* dP,
                           float* d_shape, float* d_tex) {
    const int W = x.c->W, LD = x.c->latent_dim;
    CNB_CUDA_TRY(cudaMemsetAsync(d_shape, 0, sizeof(float) * (size_t)x.n_codes * LD, x.st));
    CNB_CUDA_TRY(cudaMemsetAsync(d_tex, 0, sizeof(float) * (size_t)x.n_codes * LD, x.st));
    dim3 grid(W, x.n_codes);
    for (int j = 0; j < x.c->shape_blocks; ++j) {
        k_latent_bwd<<<grid, 128, 0, x.st>>>(x.w(x.L.i_sl[j]), shape_codes, x.B.zs + (size_t)j * W,
                                            x.B.dzs + (size_t)j * W, (int64_t)x.c->shape_blocks * W, x.n_codes, LD, W,
                                            dP ? dP + x.L.sl_w[j] : nullptr, dP ? dP + x.L.sl_b[j] : nullptr, d_shape);
        CNB_LAUNCH_CHECK();
    }
    for (int j = 0; j < x.c->texture_blocks; ++j) {
        k_latent_bwd<<<grid, 128, 0, x.st>>>(x.w(x.L.i_tl[j]), tex_codes, x.B.zt + (size_t)j * W,
                                            x.B.dzt + (size_t)j * W, (int64_t)x.c->texture_blocks * W, x.n_codes, LD, W,
                                            dP ? dP + x.L.tl_w[j] : nullptr, dP ? dP + x.L.tl_b[j] : nullptr, d_tex);
        CNB_LAUNCH_CHECK();
    }
    return CNB_OK;
}

static int make_ctx(Fp32Ctx& x, const cnb_net_config* cfg, const float* const* params, int n_codes,
                    int64_t rows_per_code, int64_t S, int N, int backward, int fused, void* ws, size_t ws_bytes,
                    cudaStream_t st, int64_t* chunk_rows) {
    x.c = cfg; cnb_make_layout(cfg, &x.L); x.P = params; x.st = st; x.n_codes = n_codes;
    x.rows_per_code = rows_per_code > 0 ? rows_per_code : (S > 0 ? S : 1);
    int64_t rows = cnb_fp32_chunk_rows(N);
    if (S < rows) rows = S;
    *chunk_rows = rows;
    const size_t need = carve(cfg, x.L, rows, n_codes, backward, fused ? (N > 0 ? rows / N + 1 : 0) : 0, nullptr, nullptr);
    if (!ws || ws_bytes < need) return CNB_E_WORKSPACE;
    if (((uintptr_t)ws & 255) != 0) return CNB_E_ALIGNMENT;
    carve(cfg, x.L, rows, n_codes, backward, fused ? (N > 0 ? rows / N + 1 : 0) : 0, (char*)ws, &x.B);
    return CNB_OK;
}

// ---- CodeNeRF.forward / backward on arrays (unfused API) --------------------
int cnb_fp32_mlp_forward(const cnb_net_config* cfg, const float* const* params, const float* xyz,
                         const float* viewdir, const float* shape_codes, const float* tex_codes, int n_codes,
                         int64_t samples_per_code, int64_t S, float* sigmas, float* rgbs, void* ws, size_t ws_bytes,
                         cudaStream_t st) {
    Fp32Ctx x; int64_t rows;
    CNB_TRY(make_ctx(x, cfg, params, n_codes, samples_per_code, S, 0, 0, 0, ws, ws_bytes, st, &rows));
    CNB_TRY(latent_forward(x, shape_codes, tex_codes));
    for (int64_t r0 = 0; r0 < S; r0 += rows) {
        const int m = (int)((S - r0 < rows) ? (S - r0) : rows);
        CNB_TRY(chunk_forward(x, xyz, viewdir, nullptr, r0, m));
        CNB_CUDA_TRY(cudaMemcpyAsync(sigmas + r0, x.B.sig, sizeof(float) * m, cudaMemcpyDeviceToDevice, st));
        CNB_CUDA_TRY(cudaMemcpyAsync(rgbs + r0 * 3, x.B.rgbs, sizeof(float) * 3 * m, cudaMemcpyDeviceToDevice, st));
    }
    return CNB_OK;
}

int cnb_fp32_mlp_backward(const cnb_net_config* cfg, const float* const* params, const float* xyz,
                          const float* viewdir, const float* shape_codes, const float* tex_codes, int n_codes,
                          int64_t samples_per_code, int64_t S, const float* d_sigmas, const float* d_rgbs,
                          float* d_params, float* d_shape, float* d_tex, void* ws, size_t ws_bytes, cudaStream_t st) {
    Fp32Ctx x; int64_t rows;
    CNB_TRY(make_ctx(x, cfg, params, n_codes, samples_per_code, S, 0, 1, 0, ws, ws_bytes, st, &rows));
    const int W = cfg->W;
    CNB_TRY(latent_forward(x, shape_codes, tex_codes));
    CNB_CUDA_TRY(cudaMemsetAsync(x.B.dzs, 0, sizeof(float) * (size_t)n_codes * cfg->shape_blocks * W, st));
    CNB_CUDA_TRY(cudaMemsetAsync(x.B.dzt, 0, sizeof(float) * (size_t)n_codes * cfg->texture_blocks * W, st));
    for (int64_t r0 = 0; r0 < S; r0 += rows) {
        const int m = (int)((S - r0 < rows) ? (S - r0) : rows);
        CNB_TRY(chunk_forward(x, xyz, viewdir, nullptr, r0, m));
        CNB_TRY(chunk_backward(x, d_sigmas + r0, d_rgbs + r0 * 3, d_params, r0, m));
    }
    return latent_backward(x, shape_codes, tex_codes, d_params, d_shape, d_tex);
}

// ---- fused render API in fp32 mode ------------------------------------------
// mode 0: forward only; 1: backward from (d_rgb, d_depth); 2: train step (seed from target)
int cnb_fp32_render(const cnb_net_config* cfg, const float* const* params, const cnb_ray_batch* rays, int mode,
                    const float* d_rgb, const float* d_depth, const float* target, float loss_scale,
                    float* rgb, float* depth, float* acc, float* sq_err,
                    float* d_params, float* d_shape, float* d_tex, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int N = rays->n_samples;
    const int64_t S = rays->n_rays * N;
    const int backward = mode != 0;
    Fp32Ctx x; int64_t rows;
    const int64_t rows_per_code = (int64_t)rays->segments_per_code * rays->rays_per_segment * N;
    CNB_TRY(make_ctx(x, cfg, params, rays->n_codes, rays->n_codes > 1 ? rows_per_code : 0, S, N, backward, 1, ws,
                     ws_bytes, st, &rows));
    if (rows % N != 0) rows = (rows / N) * N;
    if (rows <= 0) return CNB_E_UNSUPPORTED;
    const int W = cfg->W;
    CnbRaySource rs = cnb_make_ray_source(rays);
    CNB_TRY(latent_forward(x, rays->shape_codes, rays->texture_codes));
    if (backward) {
        CNB_CUDA_TRY(cudaMemsetAsync(x.B.dzs, 0, sizeof(float) * (size_t)rays->n_codes * cfg->shape_blocks * W, st));
        CNB_CUDA_TRY(cudaMemsetAsync(x.B.dzt, 0, sizeof(float) * (size_t)rays->n_codes * cfg->texture_blocks * W, st));
    }
    if (mode == 2 && sq_err)
        CNB_CUDA_TRY(cudaMemsetAsync(sq_err, 0, sizeof(float) * (size_t)(rays->n_rays / rays->rays_per_segment), st));
    // chunks never straddle a segment when z differs per segment: cut at segment ends
    const int64_t seg_rows = (int64_t)rays->rays_per_segment * N;
    for (int64_t r0 = 0; r0 < S;) {
        int64_t m = (S - r0 < rows) ? (S - r0) : rows;
        if (rays->z_per_segment) {
            const int64_t seg_end = (r0 / seg_rows + 1) * seg_rows;
            if (r0 + m > seg_end) m = seg_end - r0;
        }
        const int64_t ray0 = r0 / N, nr = m / N;
        const float* z = rays->z_vals + (rays->z_per_segment ? (r0 / seg_rows) * N : 0);
        CNB_TRY(chunk_forward(x, nullptr, nullptr, &rs, r0, (int)m));
        float* o_rgb = rgb ? rgb + ray0 * 3 : x.B.ray_rgb;
        float* o_depth = depth ? depth + ray0 : x.B.ray_depth;
        float* o_acc = acc ? acc + ray0 : x.B.ray_acc;
        CNB_TRY(cnb_volume_rendering_forward(x.B.sig, x.B.rgbs, z, nr, N, rays->white_bg, o_rgb, o_depth, o_acc,
                                             (cnb_stream_t)st));
        if (backward) {
            const float* seed_rgb; const float* seed_depth = nullptr;
            if (mode == 2) {
                k_l2_seed<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(o_rgb, target + ray0 * 3, nr, ray0,
                                                                       rays->rays_per_segment, loss_scale,
                                                                       x.B.ray_drgb, sq_err);
                CNB_LAUNCH_CHECK();
                seed_rgb = x.B.ray_drgb;
            } else {
                seed_rgb = d_rgb + ray0 * 3;
                seed_depth = d_depth ? d_depth + ray0 : nullptr;
            }
            CNB_TRY(cnb_volume_rendering_backward(x.B.sig, x.B.rgbs, z, nr, N, rays->white_bg, seed_rgb, seed_depth,
                                                  x.B.dsig, x.B.drgbs, (cnb_stream_t)st));
            CNB_TRY(chunk_backward(x, x.B.dsig, x.B.drgbs, d_params, r0, (int)m));
        }
        r0 += m;
    }
    if (backward) return latent_backward(x, rays->shape_codes, rays->texture_codes, d_params, d_shape, d_tex);
    return CNB_OK;
}
