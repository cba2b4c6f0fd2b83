/*
 * codenerf_oracle.c -- CPU restatement of the CodeNeRF render path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 * The product (codenerf_b200/) never links, imports or falls back to it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so this restatement is pinned against outputs of the
 * reference's own Python, executed in the authoring container by
 * tests/golden/make_golden.py and committed under tests/golden/ (see
 * tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to the reference checkout, yuliangguo/code-nerf).
 *
 * Build: oracle/Makefile.  This file MUST be compiled with -ffp-contract=off:
 * the ray / sample arithmetic is bit-exact only with the explicit roundings
 * written below (fused multiply-adds appear only as explicit fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int shape_blocks;   /* src/model.py:11 */
    int texture_blocks;
    int W;
    int num_xyz_freq;
    int num_dir_freq;
    int latent_dim;
} orc_config;

/* ------------------------------------------------------------------------- */
/* torch.manual_seed(seed); torch.rand(n)  on the CPU default generator.
 * Used by src/utils.py:29.  torch's CPUGeneratorImpl is a 32-bit mt19937 seeded
 * with init_genrand(seed); a float32 sample keeps the low 24 bits of one 32-bit
 * draw and scales by 2^-24 (ATen uniform_real_distribution<float>).            */
typedef struct { uint32_t mt[624]; int idx; } orc_mt;

static void mt_seed(orc_mt* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}
static uint32_t mt_next(orc_mt* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            g->mt[i] = v;
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
void orc_torch_rand(uint32_t seed, int n, float* out) {
    orc_mt g; mt_seed(&g, seed);
    for (int i = 0; i < n; ++i) out[i] = (float)(mt_next(&g) & 0xFFFFFFu) * 0x1p-24f;
}

/* ------------------------------------------------------------------------- */
/* get_rays  -- src/utils.py:10-19.
 * focal_is_f64 != 0 models the trainer's call (focal is an fp64 tensor of shape
 * [1] after DataLoader collation, src/data.py:34 / src/trainer.py:65): the pixel
 * offsets are divided in fp64 and rounded to fp32 by .type_as(c2w) (utils.py:15).
 * Otherwise (python float / 0-dim focal) the division is fp32.
 * rays_d = (p0+p1)+p2 with separately rounded products (torch.sum over a size-3
 * last dim); |rays_d| = sqrt(fma(d2,d2,fma(d1,d1,d0*d0))) (torch.norm, dim=-1). */
void orc_get_rays(int H, int W, double focal, int focal_is_f64, const float* c2w /*4x4 row-major*/,
                  float* rays_o /*[H*W,3]*/, float* viewdirs /*[H*W,3]*/) {
    const float half_w = (float)(W * 0.5), half_h = (float)(H * 0.5);
    const float focal32 = (float)focal;
    for (int r = 0; r < H; ++r) {
        for (int c = 0; c < W; ++c) {
            float fi = (float)c - half_w;          /* utils.py:14  (i - W*.5), fp32 */
            float fj = (float)r - half_h;
            float dx, dy;
            if (focal_is_f64) {
                dx = (float)((double)fi / focal);
                dy = (float)(-((double)fj) / focal);
            } else {
                dx = fi / focal32;
                dy = (-fj) / focal32;
            }
            const float dz = -1.0f;
            float d[3];
            for (int k = 0; k < 3; ++k) {          /* utils.py:15 */
                float p0 = dx * c2w[k * 4 + 0];
                float p1 = dy * c2w[k * 4 + 1];
                float p2 = dz * c2w[k * 4 + 2];
                d[k] = (p0 + p1) + p2;
            }
            float n2 = fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0]));
            float nrm = sqrtf(n2);                 /* utils.py:16 */
            size_t o = ((size_t)r * W + c) * 3;
            for (int k = 0; k < 3; ++k) {
                viewdirs[o + k] = d[k] / nrm;
                rays_o[o + k] = c2w[k * 4 + 3];    /* utils.py:17 */
            }
        }
    }
}

/* torch.linspace(start, end, steps) for float32 on CPU (ATen RangeFactories):
 * step = (end-start)/(steps-1) in fp32; first half start+step*i, second half
 * end-step*(steps-1-i), both fused.                                          */
static void torch_linspace_f32(double start_d, double end_d, int steps, float* out) {
    float start = (float)start_d, end = (float)end_d;
    if (steps == 1) { out[0] = start; return; }
    float step = (end - start) / (float)(steps - 1);
    int half = steps / 2;
    for (int i = 0; i < steps; ++i)
        out[i] = (i < half) ? fmaf(step, (float)i, start) : fmaf(-step, (float)(steps - 1 - i), end);
}

/* z_vals of sample_from_rays -- src/utils.py:24-29.  `rnd` is the torch.rand(N)
 * draw (utils.py:29); pass NULL only with z_fixed.                            */
void orc_z_vals(double near, double far, int N, const float* rnd, int z_fixed, float* z) {
    if (z_fixed) { torch_linspace_f32(near, far, N, z); return; }   /* utils.py:25 */
    double dist = (far - near) / (double)(2 * N);                    /* utils.py:27 */
    torch_linspace_f32(near + dist, far - dist, N, z);               /* utils.py:28 */
    float span = (float)(far - near), den = (float)(2 * N);
    for (int i = 0; i < N; ++i) z[i] = z[i] + (rnd[i] * span) / den; /* utils.py:29 */
}

/* xyz / repeated viewdir of sample_from_rays -- src/utils.py:30-31.
 * Separate multiply and add roundings (no FMA).                               */
void orc_sample_from_rays(const float* ro, const float* vd, const float* z, int R, int N,
                          float* xyz /*[R,N,3]*/, float* vd_rep /*[R,N,3]*/) {
    for (int r = 0; r < R; ++r)
        for (int i = 0; i < N; ++i)
            for (int k = 0; k < 3; ++k) {
                size_t o = ((size_t)r * N + i) * 3 + k;
                float m = vd[r * 3 + k] * z[i];
                xyz[o] = ro[r * 3 + k] + m;
                vd_rep[o] = vd[r * 3 + k];
            }
}

/* PE -- src/model.py:4-7.  Column order: x(3), sin(2^0 x)(3) .. sin(2^(L-1) x)(3),
 * then all cosines.                                                          */
void orc_pe(const float* x /*[n,3]*/, int n, int degree, float* out /*[n,3+6*degree]*/) {
    int d = 3 + 6 * degree;
    for (int s = 0; s < n; ++s) {
        float* o = out + (size_t)s * d;
        for (int k = 0; k < 3; ++k) o[k] = x[s * 3 + k];
        for (int i = 0; i < degree; ++i)
            for (int k = 0; k < 3; ++k) {
                float a = ldexpf(x[s * 3 + k], i);   /* 2.**i * x : exact */
                o[3 + 3 * i + k] = sinf(a);
                o[3 + 3 * degree + 3 * i + k] = cosf(a);
            }
    }
}

/* ------------------------------------------------------------------------- */
/* Parameter layout = CodeNeRF.state_dict() order -- src/model.py:20-34.       */
typedef struct {
    int64_t enc_xyz_w, enc_xyz_b;
    int64_t sl_w[16], sl_b[16], s_w[16], s_b[16];      /* shape_latent_layer_j, shape_layer_j */
    int64_t enc_shape_w, enc_shape_b, sigma_w, sigma_b, enc_vd_w, enc_vd_b;
    int64_t tl_w[16], tl_b[16], t_w[16], t_b[16];
    int64_t rgb0_w, rgb0_b, rgb2_w, rgb2_b;
    int64_t total;
    int d_xyz, d_dir;
} orc_layout;

static void make_layout(const orc_config* c, orc_layout* L) {
    int64_t o = 0; int W = c->W, LD = c->latent_dim;
    L->d_xyz = 3 + 6 * c->num_xyz_freq; L->d_dir = 3 + 6 * c->num_dir_freq;
    L->enc_xyz_w = o; o += (int64_t)W * L->d_xyz; L->enc_xyz_b = o; o += W;
    for (int j = 0; j < c->shape_blocks; ++j) {
        L->sl_w[j] = o; o += (int64_t)W * LD; L->sl_b[j] = o; o += W;
        L->s_w[j] = o; o += (int64_t)W * W; L->s_b[j] = o; o += W;
    }
    L->enc_shape_w = o; o += (int64_t)W * W; L->enc_shape_b = o; o += W;
    L->sigma_w = o; o += W; L->sigma_b = o; o += 1;
    L->enc_vd_w = o; o += (int64_t)W * (W + L->d_dir); L->enc_vd_b = o; o += W;
    for (int j = 0; j < c->texture_blocks; ++j) {
        L->tl_w[j] = o; o += (int64_t)W * LD; L->tl_b[j] = o; o += W;
        L->t_w[j] = o; o += (int64_t)W * W; L->t_b[j] = o; o += W;
    }
    L->rgb0_w = o; o += (int64_t)(W / 2) * W; L->rgb0_b = o; o += W / 2;
    L->rgb2_w = o; o += (int64_t)3 * (W / 2); L->rgb2_b = o; o += 3;
    L->total = o;
}
int64_t orc_param_count(const orc_config* c) {
    if (c->shape_blocks > 16 || c->texture_blocks > 16) return -1;
    orc_layout L; make_layout(c, &L); return L.total;
}

/* y[n] = act(W[n,k] x[k] + b[n])  for one vector (per-code latent layers). */
static void linear_vec(const float* W, const float* b, const float* x, int n, int k, int relu, float* y) {
    for (int i = 0; i < n; ++i) {
        float acc = b ? b[i] : 0.f;
        const float* w = W + (size_t)i * k;
        for (int j = 0; j < k; ++j) acc += w[j] * x[j];
        y[i] = (relu && acc < 0.f) ? 0.f : acc;
    }
}

/* C[m, 0:n] = act(bias + A[m, 0:k] . W[n,k]^T)  for a block of m samples.
 * Wt is W transposed to [k][n] so the inner loop runs over n.                */
static void block_linear(const float* A, int lda, const float* Wt, const float* bias, int m, int n, int k,
                         int relu, float* C, int ldc) {
    for (int r = 0; r < m; ++r) {
        float* c = C + (size_t)r * ldc;
        if (bias) memcpy(c, bias, sizeof(float) * n); else memset(c, 0, sizeof(float) * n);
        const float* a = A + (size_t)r * lda;
        for (int j = 0; j < k; ++j) {
            float av = a[j];
            if (av == 0.f) continue;
            const float* w = Wt + (size_t)j * n;
            for (int i = 0; i < n; ++i) c[i] += av * w[i];
        }
        if (relu) for (int i = 0; i < n; ++i) if (c[i] < 0.f) c[i] = 0.f;
    }
}
static float* transpose_alloc(const float* W, int n, int k) { /* W[n,k] -> Wt[k,n] */
    float* t = (float*)malloc(sizeof(float) * (size_t)n * k);
    for (int i = 0; i < n; ++i) for (int j = 0; j < k; ++j) t[(size_t)j * n + i] = W[(size_t)i * k + j];
    return t;
}

static float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }   /* nn.Softplus(beta=1,threshold=20) */

#define BLK 64

typedef struct {
    float *e0, *vpe, *y[17], *u[17], *f, *cat, *h[17], *ut[17], *r1, *spre;
} blk_acts;

static void alloc_acts(blk_acts* a, const orc_config* c, const orc_layout* L) {
    int W = c->W;
    a->e0 = (float*)malloc(sizeof(float) * BLK * L->d_xyz);
    a->vpe = (float*)malloc(sizeof(float) * BLK * L->d_dir);
    for (int j = 0; j <= c->shape_blocks; ++j) { a->y[j] = (float*)malloc(sizeof(float) * BLK * W); a->u[j] = (float*)malloc(sizeof(float) * BLK * W); }
    a->f = (float*)malloc(sizeof(float) * BLK * W);
    a->cat = (float*)malloc(sizeof(float) * BLK * (W + L->d_dir));
    for (int j = 0; j <= c->texture_blocks; ++j) { a->h[j] = (float*)malloc(sizeof(float) * BLK * W); a->ut[j] = (float*)malloc(sizeof(float) * BLK * W); }
    a->r1 = (float*)malloc(sizeof(float) * BLK * (W / 2));
    a->spre = (float*)malloc(sizeof(float) * BLK);
}
static void free_acts(blk_acts* a, const orc_config* c) {
    free(a->e0); free(a->vpe); free(a->f); free(a->cat); free(a->r1); free(a->spre);
    for (int j = 0; j <= c->shape_blocks; ++j) { free(a->y[j]); free(a->u[j]); }
    for (int j = 0; j <= c->texture_blocks; ++j) { free(a->h[j]); free(a->ut[j]); }
}

typedef struct {
    float *w0t, *st[16], *wet, *wvt, *tt[16], *r0t;
} transposed_w;

static void make_transposed(transposed_w* T, const orc_config* c, const orc_layout* L, const float* P) {
    int W = c->W;
    T->w0t = transpose_alloc(P + L->enc_xyz_w, W, L->d_xyz);
    for (int j = 0; j < c->shape_blocks; ++j) T->st[j] = transpose_alloc(P + L->s_w[j], W, W);
    T->wet = transpose_alloc(P + L->enc_shape_w, W, W);
    T->wvt = transpose_alloc(P + L->enc_vd_w, W, W + L->d_dir);
    for (int j = 0; j < c->texture_blocks; ++j) T->tt[j] = transpose_alloc(P + L->t_w[j], W, W);
    T->r0t = transpose_alloc(P + L->rgb0_w, W / 2, W);
}
static void free_transposed(transposed_w* T, const orc_config* c) {
    free(T->w0t); free(T->wet); free(T->wvt); free(T->r0t);
    for (int j = 0; j < c->shape_blocks; ++j) free(T->st[j]);
    for (int j = 0; j < c->texture_blocks; ++j) free(T->tt[j]);
}

/* Forward of one block of m<=BLK samples -- src/model.py:36-53.
 * zs[j] / zt[j]: the per-code latent branch outputs for this block's code.   */
static void block_forward(const orc_config* c, const orc_layout* L, const float* P, const transposed_w* T,
                          const float* xyz, const float* vdir, int m, float* const* zs, float* const* zt,
                          blk_acts* a, float* sig_out, float* rgb_out) {
    int W = c->W, dv = L->d_dir;
    orc_pe(xyz, m, c->num_xyz_freq, a->e0);                                   /* model.py:37 */
    orc_pe(vdir, m, c->num_dir_freq, a->vpe);                                 /* model.py:38 */
    block_linear(a->e0, L->d_xyz, T->w0t, P + L->enc_xyz_b, m, W, L->d_xyz, 1, a->y[0], W); /* :39 */
    for (int j = 0; j < c->shape_blocks; ++j) {
        for (int r = 0; r < m; ++r) for (int i = 0; i < W; ++i)
            a->u[j][r * W + i] = a->y[j][r * W + i] + zs[j][i];              /* model.py:42 */
        block_linear(a->u[j], W, T->st[j], P + L->s_b[j], m, W, W, 1, a->y[j + 1], W);      /* :43 */
    }
    block_linear(a->y[c->shape_blocks], W, T->wet, P + L->enc_shape_b, m, W, W, 0, a->f, W); /* :44 */
    for (int r = 0; r < m; ++r) {                                             /* model.py:45 */
        float acc = P[L->sigma_b];
        for (int i = 0; i < W; ++i) acc += P[L->sigma_w + i] * a->f[r * W + i];
        a->spre[r] = acc; sig_out[r] = softplus_f(acc);
        memcpy(a->cat + (size_t)r * (W + dv), a->f + (size_t)r * W, sizeof(float) * W);      /* :46 */
        memcpy(a->cat + (size_t)r * (W + dv) + W, a->vpe + (size_t)r * dv, sizeof(float) * dv);
    }
    block_linear(a->cat, W + dv, T->wvt, P + L->enc_vd_b, m, W, W + dv, 1, a->h[0], W);      /* :47 */
    for (int j = 0; j < c->texture_blocks; ++j) {
        for (int r = 0; r < m; ++r) for (int i = 0; i < W; ++i)
            a->ut[j][r * W + i] = a->h[j][r * W + i] + zt[j][i];             /* model.py:50 */
        block_linear(a->ut[j], W, T->tt[j], P + L->t_b[j], m, W, W, 1, a->h[j + 1], W);      /* :51 */
    }
    block_linear(a->h[c->texture_blocks], W, T->r0t, P + L->rgb0_b, m, W / 2, W, 1, a->r1, W / 2); /* :52 */
    for (int r = 0; r < m; ++r) for (int k = 0; k < 3; ++k) {
        float acc = P[L->rgb2_b + k];
        for (int i = 0; i < W / 2; ++i) acc += P[L->rgb2_w + (size_t)k * (W / 2) + i] * a->r1[r * (W / 2) + i];
        rgb_out[r * 3 + k] = acc;
    }
}

static void latent_forward(const orc_config* c, const orc_layout* L, const float* P, const float* shape_codes,
                           const float* tex_codes, int n_codes, float* zs /*[n_codes][SB][W]*/, float* zt) {
    int W = c->W, LD = c->latent_dim;
    for (int o = 0; o < n_codes; ++o) {
        for (int j = 0; j < c->shape_blocks; ++j)                              /* model.py:41 */
            linear_vec(P + L->sl_w[j], P + L->sl_b[j], shape_codes + (size_t)o * LD, W, LD, 1,
                       zs + ((size_t)o * c->shape_blocks + j) * W);
        for (int j = 0; j < c->texture_blocks; ++j)                            /* model.py:49 */
            linear_vec(P + L->tl_w[j], P + L->tl_b[j], tex_codes + (size_t)o * LD, W, LD, 1,
                       zt + ((size_t)o * c->texture_blocks + j) * W);
    }
}

/* CodeNeRF.forward -- src/model.py:36-53.
 * xyz, viewdir: [S,3]; codes: [n_codes, latent]; sample s uses code
 * s / samples_per_code (samples_per_code == 0: every sample uses code 0, the
 * [1,latent] broadcast of src/trainer.py:70).                                */
int orc_mlp_forward(const orc_config* c, const float* P, const float* xyz, const float* viewdir,
                    const float* shape_codes, const float* tex_codes, int n_codes, int64_t samples_per_code,
                    int64_t S, float* sigmas /*[S]*/, float* rgbs /*[S,3]*/) {
    orc_layout L; make_layout(c, &L);
    transposed_w T; make_transposed(&T, c, &L, P);
    int W = c->W;
    float* zs = (float*)malloc(sizeof(float) * (size_t)n_codes * (c->shape_blocks + 1) * W);
    float* zt = (float*)malloc(sizeof(float) * (size_t)n_codes * (c->texture_blocks + 1) * W);
    latent_forward(c, &L, P, shape_codes, tex_codes, n_codes, zs, zt);
    int64_t spc = samples_per_code > 0 ? samples_per_code : S;
#pragma omp parallel
    {
        blk_acts a; alloc_acts(&a, c, &L);
        float* zsp[16]; float* ztp[16];
#pragma omp for schedule(static)
        for (int64_t b0 = 0; b0 < S; b0 += BLK) {
            int64_t end = b0 + BLK < S ? b0 + BLK : S;
            int64_t s = b0;
            while (s < end) {                       /* split the block at code boundaries */
                int64_t code = s / spc; if (code >= n_codes) code = n_codes - 1;
                int64_t lim = (code + 1) * spc; if (lim > end) lim = end;
                for (int j = 0; j < c->shape_blocks; ++j) zsp[j] = zs + ((size_t)code * c->shape_blocks + j) * W;
                for (int j = 0; j < c->texture_blocks; ++j) ztp[j] = zt + ((size_t)code * c->texture_blocks + j) * W;
                block_forward(c, &L, P, &T, xyz + s * 3, viewdir + s * 3, (int)(lim - s), zsp, ztp, &a,
                              sigmas + s, rgbs + s * 3);
                s = lim;
            }
        }
        free_acts(&a, c);
    }
    free(zs); free(zt); free_transposed(&T, c);
    return 0;
}

/* volume_rendering -- src/utils.py:34-47.  acc is weights.sum (utils.py:45),
 * exposed because BASELINE.json's tolerances name it; rgb/depth as returned.  */
void orc_volume_rendering(const float* sigmas /*[B,N]*/, const float* rgbs /*[B,N,3]*/, const float* z /*[N]*/,
                          int64_t B, int N, int white_bg, float* rgb /*[B,3]*/, float* depth /*[B]*/, float* acc /*[B]*/) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, d = 0.f, ws = 0.f;
        for (int i = 0; i < N; ++i) {
            float delta = (i + 1 < N) ? (z[i + 1] - z[i]) : 1e10f;           /* utils.py:35-36 */
            float alpha = 1.f - expf(-sigmas[b * N + i] * delta);             /* utils.py:37 */
            float w = alpha * T;                                              /* utils.py:41 */
            const float* cc = rgbs + ((size_t)b * N + i) * 3;
            cr += w * cc[0]; cg += w * cc[1]; cb += w * cc[2];                /* utils.py:42 */
            d += w * z[i];                                                    /* utils.py:43 */
            ws += w;
            T = T * (1.f - alpha + 1e-10f);                                   /* utils.py:38-40 */
        }
        if (white_bg) { cr = cr + 1.f - ws; cg = cg + 1.f - ws; cb = cb + 1.f - ws; } /* utils.py:44-46 */
        rgb[b * 3 + 0] = cr; rgb[b * 3 + 1] = cg; rgb[b * 3 + 2] = cb; depth[b] = d; if (acc) acc[b] = ws;
    }
}

/* Reverse-mode derivative of orc_volume_rendering (what autograd does for
 * src/utils.py:34-47 under loss.backward(), src/trainer.py:82).
 * Division-free adjoint of the recurrence T_{i+1} = T_i * t_i.               */
void orc_volume_rendering_backward(const float* sigmas, const float* rgbs, const float* z, int64_t B, int N,
                                   int white_bg, const float* d_rgb /*[B,3]*/, const float* d_depth /*[B] or NULL*/,
                                   float* d_sigmas /*[B,N]*/, float* d_rgbs /*[B,N,3]*/) {
#pragma omp parallel
    {
        float* Tbuf = (float*)malloc(sizeof(float) * N);
        float* abuf = (float*)malloc(sizeof(float) * N);
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            float T = 1.f;
            for (int i = 0; i < N; ++i) {
                float delta = (i + 1 < N) ? (z[i + 1] - z[i]) : 1e10f;
                float alpha = 1.f - expf(-sigmas[b * N + i] * delta);
                Tbuf[i] = T; abuf[i] = alpha;
                T = T * (1.f - alpha + 1e-10f);
            }
            float gr = d_rgb[b * 3], gg = d_rgb[b * 3 + 1], gb = d_rgb[b * 3 + 2];
            float gd = d_depth ? d_depth[b] : 0.f;
            float bg = white_bg ? 1.f : 0.f;
            float aT_next = 0.f;                       /* adjoint of T_{i+1} */
            for (int i = N - 1; i >= 0; --i) {
                const float* cc = rgbs + ((size_t)b * N + i) * 3;
                float w = abuf[i] * Tbuf[i];
                float* dc = d_rgbs + ((size_t)b * N + i) * 3;
                dc[0] = w * gr; dc[1] = w * gg; dc[2] = w * gb;
                float g = gr * (cc[0] - bg) + gg * (cc[1] - bg) + gb * (cc[2] - bg) + gd * z[i];
                float t = 1.f - abuf[i] + 1e-10f;
                float a_t = aT_next * Tbuf[i];
                float a_alpha = g * Tbuf[i] - a_t;
                float delta = (i + 1 < N) ? (z[i + 1] - z[i]) : 1e10f;
                d_sigmas[b * N + i] = a_alpha * delta * expf(-sigmas[b * N + i] * delta);
                aT_next = g * abuf[i] + aT_next * t;
            }
        }
        free(Tbuf); free(abuf);
    }
}

/* Backward of CodeNeRF.forward (autograd of src/model.py:36-53): accumulates
 * parameter gradients into dP (flat, state_dict order) and writes code
 * gradients.  Deterministic: per-thread partial sums reduced in thread order. */
int orc_mlp_backward(const orc_config* c, const float* P, const float* xyz, const float* viewdir,
                     const float* shape_codes, const float* tex_codes, int n_codes, int64_t samples_per_code,
                     int64_t S, const float* d_sigmas /*[S]*/, const float* d_rgbs /*[S,3]*/,
                     float* dP /*[param_count], accumulated; may be NULL*/,
                     float* d_shape_codes /*[n_codes,latent], overwritten*/, float* d_tex_codes) {
    orc_layout L; make_layout(c, &L);
    transposed_w T; make_transposed(&T, c, &L, P);
    int W = c->W, LD = c->latent_dim, SB = c->shape_blocks, TB = c->texture_blocks, dv = L.d_dir, H2 = W / 2;
    float* zs = (float*)malloc(sizeof(float) * (size_t)n_codes * (SB + 1) * W);
    float* zt = (float*)malloc(sizeof(float) * (size_t)n_codes * (TB + 1) * W);
    latent_forward(c, &L, P, shape_codes, tex_codes, n_codes, zs, zt);
    int64_t spc = samples_per_code > 0 ? samples_per_code : S;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    float** dPt = (float**)calloc(nthreads, sizeof(float*));
    /* dz accumulators: d(loss)/d(z_j) summed over samples, per thread, per code */
    float** dzs_t = (float**)calloc(nthreads, sizeof(float*));
    float** dzt_t = (float**)calloc(nthreads, sizeof(float*));
#pragma omp parallel num_threads(nthreads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        float* g = (float*)calloc(L.total, sizeof(float)); dPt[tid] = g;
        float* dzs = (float*)calloc((size_t)n_codes * (SB + 1) * W, sizeof(float)); dzs_t[tid] = dzs;
        float* dzt = (float*)calloc((size_t)n_codes * (TB + 1) * W, sizeof(float)); dzt_t[tid] = dzt;
        blk_acts a; alloc_acts(&a, c, &L);
        float* sig = (float*)malloc(sizeof(float) * BLK); float* rgb = (float*)malloc(sizeof(float) * BLK * 3);
        float* da = (float*)malloc(sizeof(float) * (W + dv));   /* running activation gradient (one sample) */
        float* dp = (float*)malloc(sizeof(float) * (W + dv));   /* pre-activation gradient */
        float* zsp[16]; float* ztp[16];
#pragma omp for schedule(static)
        for (int64_t b0 = 0; b0 < S; b0 += BLK) {
            int64_t end = b0 + BLK < S ? b0 + BLK : S;
            int64_t s0 = b0;
            while (s0 < end) {
                int64_t code = s0 / spc; if (code >= n_codes) code = n_codes - 1;
                int64_t lim = (code + 1) * spc; if (lim > end) lim = end;
                int m = (int)(lim - s0);
                for (int j = 0; j < SB; ++j) zsp[j] = zs + ((size_t)code * SB + j) * W;
                for (int j = 0; j < TB; ++j) ztp[j] = zt + ((size_t)code * TB + j) * W;
                block_forward(c, &L, P, &T, xyz + s0 * 3, viewdir + s0 * 3, m, zsp, ztp, &a, sig, rgb);
                for (int r = 0; r < m; ++r) {
                    const float* drgb = d_rgbs + (s0 + r) * 3;
                    /* rgb.2 */
                    const float* r1 = a.r1 + (size_t)r * H2;
                    for (int i = 0; i < H2; ++i) dp[i] = 0.f;
                    for (int k = 0; k < 3; ++k) {
                        g[L.rgb2_b + k] += drgb[k];
                        for (int i = 0; i < H2; ++i) {
                            g[L.rgb2_w + (size_t)k * H2 + i] += drgb[k] * r1[i];
                            dp[i] += P[L.rgb2_w + (size_t)k * H2 + i] * drgb[k];
                        }
                    }
                    for (int i = 0; i < H2; ++i) if (!(r1[i] > 0.f)) dp[i] = 0.f;
                    /* rgb.0 */
                    const float* hT = a.h[TB] + (size_t)r * W;
                    for (int i = 0; i < W; ++i) da[i] = 0.f;
                    for (int n = 0; n < H2; ++n) {
                        float v = dp[n]; if (v == 0.f) continue;
                        g[L.rgb0_b + n] += v;
                        float* gw = g + L.rgb0_w + (size_t)n * W; const float* w = P + L.rgb0_w + (size_t)n * W;
                        for (int i = 0; i < W; ++i) { gw[i] += v * hT[i]; da[i] += v * w[i]; }
                    }
                    /* texture blocks */
                    for (int j = TB - 1; j >= 0; --j) {
                        const float* hj = a.h[j + 1] + (size_t)r * W; const float* ut = a.ut[j] + (size_t)r * W;
                        for (int i = 0; i < W; ++i) dp[i] = hj[i] > 0.f ? da[i] : 0.f;
                        for (int i = 0; i < W; ++i) da[i] = 0.f;
                        for (int n = 0; n < W; ++n) {
                            float v = dp[n]; if (v == 0.f) continue;
                            g[L.t_b[j] + n] += v;
                            float* gw = g + L.t_w[j] + (size_t)n * W; const float* w = P + L.t_w[j] + (size_t)n * W;
                            for (int i = 0; i < W; ++i) { gw[i] += v * ut[i]; da[i] += v * w[i]; }
                        }
                        float* dz = dzt + ((size_t)code * TB + j) * W;
                        for (int i = 0; i < W; ++i) dz[i] += da[i];
                    }
                    /* encoding_viewdir */
                    {
                        const float* h0 = a.h[0] + (size_t)r * W; const float* cat = a.cat + (size_t)r * (W + dv);
                        for (int i = 0; i < W; ++i) dp[i] = h0[i] > 0.f ? da[i] : 0.f;
                        for (int i = 0; i < W + dv; ++i) da[i] = 0.f;
                        for (int n = 0; n < W; ++n) {
                            float v = dp[n]; if (v == 0.f) continue;
                            g[L.enc_vd_b + n] += v;
                            float* gw = g + L.enc_vd_w + (size_t)n * (W + dv); const float* w = P + L.enc_vd_w + (size_t)n * (W + dv);
                            for (int i = 0; i < W + dv; ++i) { gw[i] += v * cat[i]; da[i] += v * w[i]; }
                        }
                    }
                    /* sigma head: softplus backward as ATen (z/(z+1), pass-through above threshold) */
                    {
                        float x = a.spre[r]; float ds = d_sigmas[s0 + r];
                        float zexp = expf(x); float dpre = x > 20.f ? ds : ds * zexp / (zexp + 1.f);
                        g[L.sigma_b] += dpre;
                        const float* f = a.f + (size_t)r * W;
                        for (int i = 0; i < W; ++i) { g[L.sigma_w + i] += dpre * f[i]; da[i] += dpre * P[L.sigma_w + i]; }
                    }
                    /* encoding_shape (no activation) */
                    {
                        const float* ys = a.y[SB] + (size_t)r * W;
                        for (int i = 0; i < W; ++i) dp[i] = da[i];
                        for (int i = 0; i < W; ++i) da[i] = 0.f;
                        for (int n = 0; n < W; ++n) {
                            float v = dp[n]; if (v == 0.f) continue;
                            g[L.enc_shape_b + n] += v;
                            float* gw = g + L.enc_shape_w + (size_t)n * W; const float* w = P + L.enc_shape_w + (size_t)n * W;
                            for (int i = 0; i < W; ++i) { gw[i] += v * ys[i]; da[i] += v * w[i]; }
                        }
                    }
                    /* shape blocks */
                    for (int j = SB - 1; j >= 0; --j) {
                        const float* yj = a.y[j + 1] + (size_t)r * W; const float* u = a.u[j] + (size_t)r * W;
                        for (int i = 0; i < W; ++i) dp[i] = yj[i] > 0.f ? da[i] : 0.f;
                        for (int i = 0; i < W; ++i) da[i] = 0.f;
                        for (int n = 0; n < W; ++n) {
                            float v = dp[n]; if (v == 0.f) continue;
                            g[L.s_b[j] + n] += v;
                            float* gw = g + L.s_w[j] + (size_t)n * W; const float* w = P + L.s_w[j] + (size_t)n * W;
                            for (int i = 0; i < W; ++i) { gw[i] += v * u[i]; da[i] += v * w[i]; }
                        }
                        float* dz = dzs + ((size_t)code * SB + j) * W;
                        for (int i = 0; i < W; ++i) dz[i] += da[i];
                    }
                    /* encoding_xyz (no gradient into xyz: inputs never require grad) */
                    {
                        const float* y0 = a.y[0] + (size_t)r * W; const float* e0 = a.e0 + (size_t)r * L.d_xyz;
                        for (int n = 0; n < W; ++n) {
                            float v = y0[n] > 0.f ? da[n] : 0.f; if (v == 0.f) continue;
                            g[L.enc_xyz_b + n] += v;
                            float* gw = g + L.enc_xyz_w + (size_t)n * L.d_xyz;
                            for (int i = 0; i < L.d_xyz; ++i) gw[i] += v * e0[i];
                        }
                    }
                }
                s0 = lim;
            }
        }
        free_acts(&a, c); free(sig); free(rgb); free(da); free(dp);
    }
    /* ordered reduction over threads, then the per-code latent layers */
    float* G = (float*)calloc(L.total, sizeof(float));
    float* dzs = (float*)calloc((size_t)n_codes * (SB + 1) * W, sizeof(float));
    float* dzt = (float*)calloc((size_t)n_codes * (TB + 1) * W, sizeof(float));
    for (int t = 0; t < nthreads; ++t) {
        if (!dPt[t]) continue;
        for (int64_t i = 0; i < L.total; ++i) G[i] += dPt[t][i];
        for (size_t i = 0; i < (size_t)n_codes * SB * W; ++i) dzs[i] += dzs_t[t][i];
        for (size_t i = 0; i < (size_t)n_codes * TB * W; ++i) dzt[i] += dzt_t[t][i];
        free(dPt[t]); free(dzs_t[t]); free(dzt_t[t]);
    }
    memset(d_shape_codes, 0, sizeof(float) * (size_t)n_codes * LD);
    memset(d_tex_codes, 0, sizeof(float) * (size_t)n_codes * LD);
    for (int o = 0; o < n_codes; ++o) {
        for (int j = 0; j < SB; ++j) {                 /* shape_latent_layer_j: Linear + ReLU on the code */
            const float* zj = zs + ((size_t)o * SB + j) * W; const float* dz = dzs + ((size_t)o * SB + j) * W;
            const float* code = shape_codes + (size_t)o * LD;
            for (int n = 0; n < W; ++n) {
                float v = zj[n] > 0.f ? dz[n] : 0.f; if (v == 0.f) continue;
                G[L.sl_b[j] + n] += v;
                for (int i = 0; i < LD; ++i) {
                    G[L.sl_w[j] + (size_t)n * LD + i] += v * code[i];
                    d_shape_codes[(size_t)o * LD + i] += v * P[L.sl_w[j] + (size_t)n * LD + i];
                }
            }
        }
        for (int j = 0; j < TB; ++j) {
            const float* zj = zt + ((size_t)o * TB + j) * W; const float* dz = dzt + ((size_t)o * TB + j) * W;
            const float* code = tex_codes + (size_t)o * LD;
            for (int n = 0; n < W; ++n) {
                float v = zj[n] > 0.f ? dz[n] : 0.f; if (v == 0.f) continue;
                G[L.tl_b[j] + n] += v;
                for (int i = 0; i < LD; ++i) {
                    G[L.tl_w[j] + (size_t)n * LD + i] += v * code[i];
                    d_tex_codes[(size_t)o * LD + i] += v * P[L.tl_w[j] + (size_t)n * LD + i];
                }
            }
        }
    }
    if (dP) for (int64_t i = 0; i < L.total; ++i) dP[i] += G[i];
    free(G); free(dzs); free(dzt); free(dPt); free(dzs_t); free(dzt_t);
    free(zs); free(zt); free_transposed(&T, c);
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm: torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline is allowed all host cores. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
