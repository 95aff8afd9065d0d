/*
 * gs_oracle.c -- CPU ORACLE for the 3D-Gaussian-splatting rasterizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pf3plat_b200/ or
 * diff_gaussian_rasterization/ may import, link or execute this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the algorithm restated here lives in a third-party pip
 * dependency of the reference that is NOT present under /root/reference:
 *   git+https://github.com/dcharatan/diff-gaussian-rasterization-modified
 *   (unpinned; /root/reference/requirements.txt:2), a fork of
 *   graphdeco-inria/diff-gaussian-rasterization.
 * The reference's own contact with it is the import and three call sites in
 *   /root/reference/src/model/decoder/cuda_splatting.py:5-8, 99-124, 192-217.
 * The reference ships no tests or golden vectors for this path (SURVEY.md
 * section 4), so this file restates the published algorithm of the upstream
 * project and is pinned by (1) hand-derived
 * known-answer cases, (2) fp64 finite differences against its own analytic
 * backward and (3) an independent vectorised PyTorch autograd restatement
 * (oracle/torch_oracle.py).  See tests/test_oracle_*.py.
 *
 * Conventions (cuda_splatting.py:85-87 corroborates the transposes): the
 * 4x4 matrices arrive transposed, so the flat buffer m[] is column-major:
 * standard M[row][col] == m[col*4 + row].
 *
 * Build: `make -C oracle` produces libgs_oracle_f32.so (real=float) and
 * libgs_oracle_f64.so (real=double, -DGSO_DOUBLE) with the same entry points.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef GSO_DOUBLE
typedef double real;
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#define R_FMA fma
#define R_FABS fabs
#else
typedef float real;
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#define R_FMA fmaf
#define R_FABS fabsf
#endif

#define RL(x) ((real)(x))
#define TILE 16

/* constants of the upstream algorithm, named so Appendix C's unknowns are
 * visible */
typedef struct gso_params {
    int32_t P;            /* number of Gaussians */
    int32_t M;            /* SH coefficients per channel (stride), 0 if no SH */
    int32_t sh_degree;    /* active degree requested by the caller */
    int32_t H, W;
    int32_t prefiltered;
    int32_t sh_eval_max_degree; /* upstream evaluates at most degree 3 */
    int32_t pad_;
    real tanfovx, tanfovy;
    real scale_modifier;
    real near_cull_z;     /* 0.2 upstream */
    real dilation;        /* 0.3 upstream */
    real guard_band;      /* 1.3 upstream */
    real bg[3];
    real view[16];
    real proj[16];
    real campos[3];
} gso_params;

typedef struct gso_handle {
    gso_params p;
    /* borrowed copies of inputs needed by backward */
    real *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *cov3D_in;
    /* geometry state */
    real *depths;        /* P */
    real *xy;            /* P*2 */
    real *conic_opacity; /* P*4 */
    real *rgb;           /* P*3 (SH result or copy of colors_precomp) */
    real *cov3D;         /* P*6 (precomp copy or computed) */
    uint8_t *clamped;    /* P*3 */
    int32_t *radii;      /* P */
    int32_t *tiles_touched; /* P */
    int32_t *rect;       /* P*4: minx,miny,maxx,maxy */
    int32_t *rect_outer, *rect_inner; /* P*4 each: fragility bookkeeping */
    uint8_t *geom_fragile; /* P */
    /* binning state */
    int64_t D;
    uint32_t *point_list; /* D */
    uint64_t *keys;       /* D (sorted) */
    int64_t *ranges;      /* tiles*2 */
    /* image state */
    real *final_T;       /* H*W */
    int32_t *n_contrib;  /* H*W */
    uint8_t *px_fragile; /* H*W */
    int64_t vis;
    int64_t n_pairs_eval; /* (pixel,Gaussian) pairs evaluated in forward */
    int has_depth;
} gso_handle;

static const real SH_C0 = RL(0.28209479177387814);
static const real SH_C1 = RL(0.4886025119029199);
static const real SH_C2[5] = {RL(1.0925484305920792), RL(-1.0925484305920792),
                              RL(0.31539156525252005), RL(-1.0925484305920792),
                              RL(0.5462742152960396)};
static const real SH_C3[7] = {RL(-0.5900435899266435), RL(2.890611442640554),
                              RL(-0.4570457994644658), RL(0.3731763325901154),
                              RL(-0.4570457994644658), RL(1.445305721320277),
                              RL(-0.5900435899266435)};

/* transformPoint4x3 / 4x4 (Appendix A "Conventions").  Written as an explicit
 * fma chain so that the CUDA kernels, which use the same chain, produce
 * bit-identical camera-space depth (the sort key). */
static inline void xform4x3(const real *m, const real *p, real *o) {
    o[0] = R_FMA(m[0], p[0], R_FMA(m[4], p[1], R_FMA(m[8], p[2], m[12])));
    o[1] = R_FMA(m[1], p[0], R_FMA(m[5], p[1], R_FMA(m[9], p[2], m[13])));
    o[2] = R_FMA(m[2], p[0], R_FMA(m[6], p[1], R_FMA(m[10], p[2], m[14])));
}
static inline void xform4x4(const real *m, const real *p, real *o) {
    o[0] = R_FMA(m[0], p[0], R_FMA(m[4], p[1], R_FMA(m[8], p[2], m[12])));
    o[1] = R_FMA(m[1], p[0], R_FMA(m[5], p[1], R_FMA(m[9], p[2], m[13])));
    o[2] = R_FMA(m[2], p[0], R_FMA(m[6], p[1], R_FMA(m[10], p[2], m[14])));
    o[3] = R_FMA(m[3], p[0], R_FMA(m[7], p[1], R_FMA(m[11], p[2], m[15])));
}
static inline real ndc2pix(real v, int S) { return ((v + RL(1.0)) * (real)S - RL(1.0)) * RL(0.5); }

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline real rmin(real a, real b) { return a < b ? a : b; }
static inline real rmax(real a, real b) { return a > b ? a : b; }

/* "near an integer" helper for fragility flags */
static inline int near_int(real v, real rel) {
    real r = (real)floor((double)v + 0.5);
    return R_FABS(v - r) <= rel * rmax(RL(1.0), R_FABS(v));
}

/* Sigma = R S^2 R^T from (scale_modifier*scales, quaternion (r,x,y,z) NOT
 * normalised) -- Appendix A "cov3D". */
static void cov3d_from_scale_rot(const real *s, real mod, const real *q, real *c6) {
    real r = q[0], x = q[1], y = q[2], z = q[3];
    real R[3][3] = {{RL(1.0) - RL(2.0) * (y * y + z * z), RL(2.0) * (x * y - r * z), RL(2.0) * (x * z + r * y)},
                    {RL(2.0) * (x * y + r * z), RL(1.0) - RL(2.0) * (x * x + z * z), RL(2.0) * (y * z - r * x)},
                    {RL(2.0) * (x * z - r * y), RL(2.0) * (y * z + r * x), RL(1.0) - RL(2.0) * (x * x + y * y)}};
    real s2[3] = {mod * s[0] * mod * s[0], mod * s[1] * mod * s[1], mod * s[2] * mod * s[2]};
    real S[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            real a = 0;
            for (int k = 0; k < 3; k++) a += R[i][k] * s2[k] * R[j][k];
            S[i][j] = a;
        }
    c6[0] = S[0][0]; c6[1] = S[0][1]; c6[2] = S[0][2];
    c6[3] = S[1][1]; c6[4] = S[1][2]; c6[5] = S[2][2];
}

/* rows m0,m1 of M = J * Rview (2x3), plus clamp masks -- Appendix A "cov2D" */
typedef struct { real m0[3], m1[3]; real t[3]; real fx, fy; int xmask, ymask; } proj_jac;

static void build_jac(const gso_params *p, const real *mean, proj_jac *o) {
    real t[3];
    xform4x3(p->view, mean, t);
    real limx = p->guard_band * p->tanfovx, limy = p->guard_band * p->tanfovy;
    real txtz = t[0] / t[2], tytz = t[1] / t[2];
    o->xmask = !(txtz < -limx || txtz > limx);
    o->ymask = !(tytz < -limy || tytz > limy);
    t[0] = rmin(limx, rmax(-limx, txtz)) * t[2];
    t[1] = rmin(limy, rmax(-limy, tytz)) * t[2];
    real fx = (real)p->W / (RL(2.0) * p->tanfovx), fy = (real)p->H / (RL(2.0) * p->tanfovy);
    real J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
    real J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
    /* Rview[i][j] = view[j*4+i] */
    for (int j = 0; j < 3; j++) {
        o->m0[j] = J00 * p->view[j * 4 + 0] + J02 * p->view[j * 4 + 2];
        o->m1[j] = J11 * p->view[j * 4 + 1] + J12 * p->view[j * 4 + 2];
    }
    o->t[0] = t[0]; o->t[1] = t[1]; o->t[2] = t[2];
    o->fx = fx; o->fy = fy;
}

static inline void sym6_mul(const real *c, const real *v, real *o) {
    o[0] = c[0] * v[0] + c[1] * v[1] + c[2] * v[2];
    o[1] = c[1] * v[0] + c[3] * v[1] + c[4] * v[2];
    o[2] = c[2] * v[0] + c[4] * v[1] + c[5] * v[2];
}

static void cov2d(const gso_params *p, const real *mean, const real *c6, real *abc) {
    proj_jac j;
    build_jac(p, mean, &j);
    real s0[3], s1[3];
    sym6_mul(c6, j.m0, s0);
    sym6_mul(c6, j.m1, s1);
    abc[0] = j.m0[0] * s0[0] + j.m0[1] * s0[1] + j.m0[2] * s0[2] + p->dilation;
    abc[1] = j.m0[0] * s1[0] + j.m0[1] * s1[1] + j.m0[2] * s1[2];
    abc[2] = j.m1[0] * s1[0] + j.m1[1] * s1[1] + j.m1[2] * s1[2] + p->dilation;
}

/* SH -> RGB (Appendix A "Colour").  deg is min(sh_degree, sh_eval_max_degree). */
static void sh_to_rgb(int deg, int M, const real *mean, const real *campos, const real *sh,
                      real *rgb, uint8_t *clamped) {
    real d[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
    real inv = RL(1.0) / R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    real x = d[0] * inv, y = d[1] * inv, z = d[2] * inv;
    (void)M;
    for (int c = 0; c < 3; c++) {
#define SH(k) sh[(k) * 3 + c]
        real r = SH_C0 * SH(0);
        if (deg > 0) {
            r = r - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
            if (deg > 1) {
                real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) +
                    SH_C2[2] * (RL(2.0) * zz - xx - yy) * SH(6) + SH_C2[3] * xz * SH(7) +
                    SH_C2[4] * (xx - yy) * SH(8);
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (RL(3.0) * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
                        SH_C3[2] * y * (RL(4.0) * zz - xx - yy) * SH(11) +
                        SH_C3[3] * z * (RL(2.0) * zz - RL(3.0) * xx - RL(3.0) * yy) * SH(12) +
                        SH_C3[4] * x * (RL(4.0) * zz - xx - yy) * SH(13) +
                        SH_C3[5] * z * (xx - yy) * SH(14) + SH_C3[6] * x * (xx - RL(3.0) * yy) * SH(15);
                }
            }
        }
#undef SH
        r += RL(0.5);
        clamped[c] = (uint8_t)(r < 0);
        rgb[c] = rmax(r, RL(0.0));
    }
}

static real *dup_real(const real *src, size_t n) {
    if (!src) return NULL;
    real *d = (real *)malloc(sizeof(real) * (n ? n : 1));
    const size_t chunk = (size_t)1 << 20; /* copied in parallel: the SH block alone is 150 MB in the C2 workload */
    const int64_t nchunks = (int64_t)((n + chunk - 1) / chunk);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nchunks; c++) {
        size_t lo = (size_t)c * chunk, len = n - lo < chunk ? n - lo : chunk;
        memcpy(d + lo, src + lo, sizeof(real) * len);
    }
    return d;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

void gso_free(gso_handle *h) {
    if (!h) return;
    free(h->means3D); free(h->shs); free(h->colors_precomp); free(h->opacities);
    free(h->scales); free(h->rotations); free(h->cov3D_in);
    free(h->depths); free(h->xy); free(h->conic_opacity); free(h->rgb); free(h->cov3D);
    free(h->clamped); free(h->radii); free(h->tiles_touched); free(h->rect); free(h->rect_outer); free(h->rect_inner); free(h->geom_fragile);
    free(h->point_list); free(h->keys); free(h->ranges);
    free(h->final_T); free(h->n_contrib); free(h->px_fragile);
    free(h);
}

#ifdef _OPENMP
#include <omp.h>
#endif
/* Number of OpenMP threads the oracle uses from now on (torchrun exports OMP_NUM_THREADS=1 to its children; the
 * CPU-baseline legs of bench.py ask for all host cores explicitly).  Returns the value in effect. */
int gso_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* Fragility bookkeeping only: two consecutive contributors of a pixel whose fp32 depths lie within this many ulps are
 * flagged as an undecided order (default 4).  0 switches the flag off -- for checking an implementation whose depth
 * keys are bit-identical to this file's (same fma chain), in regimes where exact ties are the rule (the
 * fake-orthographic camera of cuda_splatting.py:154-165 puts every depth near 1.2e3, where fp32 resolves 1.2e-4). */
static float g_depth_tie_ulps = 4.0f;
void gso_set_depth_tie_ulps(float u) { g_depth_tie_ulps = u; }

int gso_real_size(void) { return (int)sizeof(real); }
int gso_params_size(void) { return (int)sizeof(gso_params); }

/*
 * Forward: Appendix A "Preprocess", "Binning", "Composite forward".
 * out_color: (3,H,W); out_depth: optional (H,W) = sum_i alpha_i T_i z_i (the
 * quantity render_depth_cuda obtains by rendering depth as colour with bg=0,
 * cuda_splatting.py:226-269); radii: (P).
 * frag_rel: relative width of the "fragile" band around each discontinuous
 * decision (0 disables flagging).
 */
gso_handle *gso_forward(const gso_params *pp, const real *means3D, const real *shs,
                        const real *colors_precomp, const real *opacities, const real *scales,
                        const real *rotations, const real *cov3D_precomp, real *out_color,
                        real *out_depth, int32_t *radii_out, double frag_rel_d) {
    const real frag_rel = (real)frag_rel_d;       /* band around the alpha / transmittance thresholds */
    const real geo_rel = (real)(frag_rel_d * 0.1); /* band around ceil()/int()/cull decisions of the geometry */
    gso_handle *h = (gso_handle *)calloc(1, sizeof(gso_handle));
    h->p = *pp;
    const gso_params *p = &h->p;
    const int P = p->P, H = p->H, W = p->W;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const int ntiles = gx * gy;
    h->means3D = dup_real(means3D, (size_t)P * 3);
    h->shs = dup_real(shs, (size_t)P * 3 * (size_t)p->M);
    h->colors_precomp = dup_real(colors_precomp, (size_t)P * 3);
    h->opacities = dup_real(opacities, (size_t)P);
    h->scales = dup_real(scales, (size_t)P * 3);
    h->rotations = dup_real(rotations, (size_t)P * 4);
    h->cov3D_in = dup_real(cov3D_precomp, (size_t)P * 6);
    h->depths = (real *)calloc((size_t)P + 1, sizeof(real));
    h->xy = (real *)calloc((size_t)P * 2 + 1, sizeof(real));
    h->conic_opacity = (real *)calloc((size_t)P * 4 + 1, sizeof(real));
    h->rgb = (real *)calloc((size_t)P * 3 + 1, sizeof(real));
    h->cov3D = (real *)calloc((size_t)P * 6 + 1, sizeof(real));
    h->clamped = (uint8_t *)calloc((size_t)P * 3 + 1, 1);
    h->radii = (int32_t *)calloc((size_t)P + 1, sizeof(int32_t));
    h->tiles_touched = (int32_t *)calloc((size_t)P + 1, sizeof(int32_t));
    h->rect = (int32_t *)calloc((size_t)P * 4 + 1, sizeof(int32_t));
    h->rect_outer = (int32_t *)calloc((size_t)P * 4 + 1, sizeof(int32_t));
    h->rect_inner = (int32_t *)calloc((size_t)P * 4 + 1, sizeof(int32_t));
    h->geom_fragile = (uint8_t *)calloc((size_t)P + 1, 1);
    h->has_depth = out_depth != NULL;
    int deg = p->sh_degree < p->sh_eval_max_degree ? p->sh_degree : p->sh_eval_max_degree;

    /* ---- preprocess ---- */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        const real *m = means3D + 3 * (size_t)i;
        real pv[3];
        xform4x3(p->view, m, pv);
        if (geo_rel > 0 && R_FABS(pv[2] - p->near_cull_z) <= geo_rel * p->near_cull_z) h->geom_fragile[i] = 2;
        if (pv[2] <= p->near_cull_z) continue;
        real ph[4];
        xform4x4(p->proj, m, ph);
        real pw = RL(1.0) / (ph[3] + RL(0.0000001));
        real pproj[2] = {ph[0] * pw, ph[1] * pw};
        real *c6 = h->cov3D + 6 * (size_t)i;
        if (cov3D_precomp) memcpy(c6, cov3D_precomp + 6 * (size_t)i, sizeof(real) * 6);
        else cov3d_from_scale_rot(scales + 3 * (size_t)i, p->scale_modifier, rotations + 4 * (size_t)i, c6);
        real abc[3];
        cov2d(p, m, c6, abc);
        real det = abc[0] * abc[2] - abc[1] * abc[1];
        if (det == 0) continue;
        real det_inv = RL(1.0) / det;
        real conic[3] = {abc[2] * det_inv, -abc[1] * det_inv, abc[0] * det_inv};
        real mid = RL(0.5) * (abc[0] + abc[2]);
        real sq = R_SQRT(rmax(RL(0.1), mid * mid - det));
        real l1 = mid + sq, l2 = mid - sq;
        real r3 = RL(3.0) * R_SQRT(rmax(l1, l2));
        real rad = R_CEIL(r3);
        int radfrag = geo_rel > 0 && near_int(r3, geo_rel);
        if (radfrag) h->geom_fragile[i] |= 1;
        real px = ndc2pix(pproj[0], W), py = ndc2pix(pproj[1], H);
        real q[4] = {(px - rad) / (real)TILE, (py - rad) / (real)TILE,
                     (px + rad + (real)(TILE - 1)) / (real)TILE, (py + rad + (real)(TILE - 1)) / (real)TILE};
        int rminx = imin(gx, imax(0, (int)q[0])), rminy = imin(gy, imax(0, (int)q[1]));
        int rmaxx = imin(gx, imax(0, (int)q[2])), rmaxy = imin(gy, imax(0, (int)q[3]));
        {
            int32_t *rc = h->rect + 4 * (size_t)i;
            rc[0] = rminx; rc[1] = rminy; rc[2] = rmaxx; rc[3] = rmaxy;
        }
        if (geo_rel > 0) {
            /* outer / inner rect under a +-geo_rel perturbation (and radius +-1 if the ceil is undecided):
             * tiles in outer \ inner are the ones a 1-ulp-different implementation may or may not touch */
            real rlo = radfrag ? rad - 1 : rad, rhi = radfrag ? rad + 1 : rad;
            real e = geo_rel * RL(16.0);
            int32_t *ro = h->rect_outer + 4 * (size_t)i, *ri = h->rect_inner + 4 * (size_t)i;
            ro[0] = imin(gx, imax(0, (int)((px - rhi) / (real)TILE - e)));
            ro[1] = imin(gy, imax(0, (int)((py - rhi) / (real)TILE - e)));
            ro[2] = imin(gx, imax(0, (int)((px + rhi + (real)(TILE - 1)) / (real)TILE + e)));
            ro[3] = imin(gy, imax(0, (int)((py + rhi + (real)(TILE - 1)) / (real)TILE + e)));
            ri[0] = imin(gx, imax(0, (int)((px - rlo) / (real)TILE + e)));
            ri[1] = imin(gy, imax(0, (int)((py - rlo) / (real)TILE + e)));
            ri[2] = imin(gx, imax(0, (int)((px + rlo + (real)(TILE - 1)) / (real)TILE - e)));
            ri[3] = imin(gy, imax(0, (int)((py + rlo + (real)(TILE - 1)) / (real)TILE - e)));
            if (memcmp(ro, ri, 4 * sizeof(int32_t)) != 0) h->geom_fragile[i] |= 4;
        }
        h->xy[2 * (size_t)i] = px;
        h->xy[2 * (size_t)i + 1] = py;
        {
            real *co = h->conic_opacity + 4 * (size_t)i;
            co[0] = conic[0]; co[1] = conic[1]; co[2] = conic[2]; co[3] = opacities[i];
        }
        if ((rmaxx - rminx) * (rmaxy - rminy) == 0) continue; /* xy/conic above are only valid where radii>0 */
        if (colors_precomp) {
            memcpy(h->rgb + 3 * (size_t)i, colors_precomp + 3 * (size_t)i, sizeof(real) * 3);
        } else {
            sh_to_rgb(deg, p->M, m, p->campos, shs + (size_t)i * 3 * (size_t)p->M, h->rgb + 3 * (size_t)i,
                      h->clamped + 3 * (size_t)i);
        }
        h->depths[i] = pv[2];
        h->radii[i] = (int32_t)rad;
        h->tiles_touched[i] = (rmaxy - rminy) * (rmaxx - rminx);
    }
    if (radii_out) memcpy(radii_out, h->radii, sizeof(int32_t) * (size_t)P);

    /* ---- binning: duplicate with keys, order by (tile, depth bits, index), tile ranges ----
     * Upstream: ONE stable radix sort of (tile << 32 | depth bits) over all instances, emitted in index order.  The same
     * order is produced here in a form that uses every host thread (bench.py times this code as the CPU baseline): a
     * counting sort by tile over index-ordered chunks of Gaussians, then each tile's list sorted on its own by
     * (depth bits, index) -- a total order, so the result does not depend on the number of threads. */
    int64_t D = 0, vis = 0;
    int nchunks = 1;
#ifdef _OPENMP
    nchunks = omp_get_max_threads();
#endif
    if (nchunks > P) nchunks = P > 0 ? P : 1;
    int64_t *cnt = (int64_t *)calloc((size_t)nchunks * (size_t)ntiles + 1, sizeof(int64_t));
#pragma omp parallel for schedule(static, 1)
    for (int c = 0; c < nchunks; c++) {
        int64_t *mine = cnt + (size_t)c * ntiles;
        int lo = (int)((int64_t)P * c / nchunks), hi = (int)((int64_t)P * (c + 1) / nchunks);
        for (int i = lo; i < hi; i++) {
            if (h->radii[i] <= 0) continue;
            const int32_t *rc = h->rect + 4 * (size_t)i;
            for (int y = rc[1]; y < rc[3]; y++)
                for (int x = rc[0]; x < rc[2]; x++) mine[y * gx + x]++;
        }
    }
    for (int i = 0; i < P; i++) vis += h->radii[i] > 0;
    h->ranges = (int64_t *)calloc((size_t)ntiles * 2, sizeof(int64_t));
    for (int t = 0; t < ntiles; t++) { /* cnt[c][t] becomes chunk c's first slot in tile t */
        int64_t start = D;
        for (int c = 0; c < nchunks; c++) {
            int64_t n = cnt[(size_t)c * ntiles + t];
            cnt[(size_t)c * ntiles + t] = D;
            D += n;
        }
        if (D > start) {
            h->ranges[2 * t] = start;
            h->ranges[2 * t + 1] = D;
        }
    }
    h->D = D;
    h->vis = vis;
    h->keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(D + 1));
    h->point_list = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(D + 1));
    uint64_t *dk = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(D + 1)); /* (depth bits << 32 | index) per instance */
#pragma omp parallel for schedule(static, 1)
    for (int c = 0; c < nchunks; c++) {
        int64_t *mine = cnt + (size_t)c * ntiles;
        int lo = (int)((int64_t)P * c / nchunks), hi = (int)((int64_t)P * (c + 1) / nchunks);
        for (int i = lo; i < hi; i++) {
            if (h->radii[i] <= 0) continue;
            const int32_t *rc = h->rect + 4 * (size_t)i;
            float df = (float)h->depths[i]; /* key always uses the fp32 bit pattern */
            uint32_t dbits;
            memcpy(&dbits, &df, 4);
            for (int y = rc[1]; y < rc[3]; y++)
                for (int x = rc[0]; x < rc[2]; x++) dk[mine[y * gx + x]++] = ((uint64_t)dbits << 32) | (uint32_t)i;
        }
    }
    free(cnt);
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < ntiles; t++) {
        int64_t r0 = h->ranges[2 * t], r1 = h->ranges[2 * t + 1];
        if (r1 - r0 > 1) qsort(dk + r0, (size_t)(r1 - r0), sizeof(uint64_t), cmp_u64); /* keys are distinct */
        for (int64_t k = r0; k < r1; k++) {
            h->keys[k] = ((uint64_t)(uint32_t)t << 32) | (dk[k] >> 32);
            h->point_list[k] = (uint32_t)dk[k];
        }
    }
    free(dk);

    /* ---- composite forward ---- */
    h->final_T = (real *)calloc((size_t)H * W, sizeof(real));
    h->n_contrib = (int32_t *)calloc((size_t)H * W, sizeof(int32_t));
    h->px_fragile = (uint8_t *)calloc((size_t)H * W, 1);
    /* (tile, Gaussian) pairs whose membership is undecided (outer \ inner rect of a geometry-fragile
     * Gaussian): a pixel is fragile only if that Gaussian would be visible there (alpha >= 1/255) */
    int all_fragile = 0;
    int64_t n_unc = 0, cap_unc = 1024;
    int32_t *unc = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)cap_unc);
    for (int i = 0; i < P; i++) {
        if (!h->geom_fragile[i]) continue;
        if (h->geom_fragile[i] & 2) { all_fragile = 1; continue; } /* near-cull undecided: whole footprint in doubt */
        const int32_t *ro = h->rect_outer + 4 * (size_t)i, *ri = h->rect_inner + 4 * (size_t)i;
        for (int y = ro[1]; y < ro[3]; y++)
            for (int x = ro[0]; x < ro[2]; x++)
                if (!(x >= ri[0] && x < ri[2] && y >= ri[1] && y < ri[3])) {
                    if (n_unc == cap_unc) { cap_unc *= 2; unc = (int32_t *)realloc(unc, sizeof(int32_t) * 2 * (size_t)cap_unc); }
                    unc[2 * n_unc] = y * gx + x;
                    unc[2 * n_unc + 1] = i;
                    n_unc++;
                }
    }
    int64_t n_pairs = 0;
    const real inv255 = RL(1.0) / RL(255.0);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : n_pairs)
    for (int t = 0; t < ntiles; t++) {
        int tx = t % gx, ty = t / gx;
        int64_t r0 = h->ranges[2 * t], r1 = h->ranges[2 * t + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                real T = RL(1.0), C[3] = {0, 0, 0}, Dacc = 0;
                int32_t contributor = 0, last = 0;
                int frag = all_fragile;
                for (int64_t u = 0; u < n_unc && !frag; u++) {
                    if (unc[2 * u] != t) continue;
                    int id = unc[2 * u + 1];
                    real dx = h->xy[2 * (size_t)id] - (real)x, dy = h->xy[2 * (size_t)id + 1] - (real)y;
                    const real *co = h->conic_opacity + 4 * (size_t)id;
                    real power = RL(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power <= 0 && co[3] * R_EXP(power) >= inv255 * (RL(1.0) - frag_rel)) frag = 1;
                }
                float last_depth = -1.0f;
                for (int64_t k = r0; k < r1; k++) {
                    contributor++;
                    uint32_t id = h->point_list[k];
                    real dx = h->xy[2 * (size_t)id] - (real)x, dy = h->xy[2 * (size_t)id + 1] - (real)y;
                    const real *co = h->conic_opacity + 4 * (size_t)id;
                    real power = RL(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    n_pairs++;
                    if (power > 0) continue;
                    real alpha = rmin(RL(0.99), co[3] * R_EXP(power));
                    /* Fragility band around alpha = 1/255: frag_rel, widened by the fp32 forward-error bound of the
                     * quadratic form itself.  For a needle-shaped footprint evaluated far along its major axis the three
                     * terms are hundreds of units each and cancel to a power of about -5: ANY fp32 evaluation (upstream's
                     * order, this file's, a kernel's pre-scaled form) carries an absolute error of a few eps * sum|terms|
                     * in the power, i.e. that much RELATIVE error in alpha -- beyond 1e-4 once sum|terms| > ~100.
                     * Second fp32 limit: the projected centre itself.  ndc2Pix evaluates ((v + 1) * S - 1) / 2 with
                     * (v + 1) * S up to 2S, so a pixel coordinate is only defined to about eps * S (6e-5 px at S = 512) --
                     * contracting the expression into an fma or not already moves it by that much -- and a sharp
                     * footprint (conic up to 1 / dilation = 3.3) turns that into |grad power| * 6e-5 ~ 2e-4 of alpha.
                     * (Found on C4: one of 262 144 pixels, a Gaussian at alpha = 0.99985 / 255, worth 1e-3 of colour.) */
                    if (frag_rel > 0) {
                        real cond = RL(0.5) * (R_FABS(co[0]) * dx * dx + R_FABS(co[2]) * dy * dy) + R_FABS(co[1] * dx * dy);
                        real pos_err = RL(1.1920929e-7) * (real)(W > H ? W : H);
                        real slope = R_FABS(co[0] * dx + co[1] * dy) + R_FABS(co[2] * dy + co[1] * dx);
                        real band = frag_rel + RL(16.0) * RL(1.1920929e-7) * cond + pos_err * slope;
                        if (R_FABS(alpha - inv255) <= band * inv255) frag = 1;
                    }
                    if (alpha < inv255) continue;
                    real test_T = T * (RL(1.0) - alpha);
                    if (frag_rel > 0 && R_FABS(test_T - RL(0.0001)) <= frag_rel * RL(0.0001)) frag = 1;
                    if (test_T < RL(0.0001)) break; /* done; this Gaussian is NOT added */
                    if (frag_rel > 0) {
                        float dcur = (float)h->depths[id];
                        if (last_depth > 0 && g_depth_tie_ulps > 0 && fabsf(dcur - last_depth) <= g_depth_tie_ulps * 1.1920929e-7f * dcur) frag = 1;
                        last_depth = dcur;
                    }
                    const real *col = h->rgb + 3 * (size_t)id;
                    real w = alpha * T;
                    C[0] += col[0] * w; C[1] += col[1] * w; C[2] += col[2] * w;
                    Dacc += h->depths[id] * w;
                    T = test_T;
                    last = contributor;
                }
                size_t pix = (size_t)y * W + x;
                h->final_T[pix] = T;
                h->n_contrib[pix] = last;
                h->px_fragile[pix] = (uint8_t)frag;
                for (int c = 0; c < 3; c++) out_color[(size_t)c * H * W + pix] = C[c] + T * p->bg[c];
                if (out_depth) out_depth[pix] = Dacc;
            }
    }
    free(unc);
    h->n_pairs_eval = n_pairs;
    return h;
}

/* ---- accessors (for intermediate-level parity tests and roofline byte counts) ---- */
int64_t gso_num_rendered(const gso_handle *h) { return h->D; }
int64_t gso_num_visible(const gso_handle *h) { return h->vis; }
int64_t gso_num_pairs(const gso_handle *h) { return h->n_pairs_eval; }
const real *gso_depths(const gso_handle *h) { return h->depths; }
const real *gso_xy(const gso_handle *h) { return h->xy; }
const real *gso_conic_opacity(const gso_handle *h) { return h->conic_opacity; }
const real *gso_rgb(const gso_handle *h) { return h->rgb; }
const real *gso_final_T(const gso_handle *h) { return h->final_T; }
const int32_t *gso_n_contrib(const gso_handle *h) { return h->n_contrib; }
const int32_t *gso_tiles_touched(const gso_handle *h) { return h->tiles_touched; }
const uint32_t *gso_point_list(const gso_handle *h) { return h->point_list; }
const int64_t *gso_ranges(const gso_handle *h) { return h->ranges; }
const uint8_t *gso_px_fragile(const gso_handle *h) { return h->px_fragile; }
const uint8_t *gso_geom_fragile(const gso_handle *h) { return h->geom_fragile; }
const uint8_t *gso_clamped(const gso_handle *h) { return h->clamped; }

/*
 * Backward: Appendix A "Composite backward" + "Preprocess backward".
 * dL_dpix: (3,H,W); dL_ddepthpix: optional (H,W), gradient w.r.t. out_depth.
 * Outputs (any may be NULL): dL_dmeans3D (P,3), dL_dmeans2D (P,3; NDC-scaled
 * screen-space gradient, z=0), dL_dsh (P,M,3), dL_dcolors (P,3),
 * dL_dopacity (P), dL_dscales (P,3), dL_drot (P,4), dL_dcov3D (P,6).
 */
void gso_backward(const gso_handle *h, const real *dL_dpix, const real *dL_ddepthpix,
                  real *dL_dmeans3D, real *dL_dmeans2D, real *dL_dsh, real *dL_dcolors_out,
                  real *dL_dopacity_out, real *dL_dscales, real *dL_drot, real *dL_dcov3D_out) {
    const gso_params *p = &h->p;
    const int P = p->P, H = p->H, W = p->W;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const int ntiles = gx * gy;
    real *g_mean2D = (real *)calloc((size_t)P * 2 + 1, sizeof(real));
    real *g_conic = (real *)calloc((size_t)P * 3 + 1, sizeof(real)); /* x, y(stored once), z */
    real *g_opac = (real *)calloc((size_t)P + 1, sizeof(real));
    real *g_col = (real *)calloc((size_t)P * 3 + 1, sizeof(real));
    real *g_z = (real *)calloc((size_t)P + 1, sizeof(real)); /* dL/d(depth_i) from the depth channel */
    const real inv255 = RL(1.0) / RL(255.0);
    const real ddelx_dx = RL(0.5) * (real)W, ddely_dy = RL(0.5) * (real)H;

    /* serial over tiles: accumulation order is then deterministic */
    for (int t = 0; t < ntiles; t++) {
        int tx = t % gx, ty = t / gx;
        int64_t r0 = h->ranges[2 * t];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                size_t pix = (size_t)y * W + x;
                const real T_final = h->final_T[pix];
                real T = T_final;
                int32_t last = h->n_contrib[pix];
                real accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
                real accum_d = 0, last_d = 0;
                real dLp[3] = {dL_dpix[pix], dL_dpix[(size_t)H * W + pix], dL_dpix[2 * (size_t)H * W + pix]};
                real dLd = dL_ddepthpix ? dL_ddepthpix[pix] : 0;
                real bg_dot = p->bg[0] * dLp[0] + p->bg[1] * dLp[1] + p->bg[2] * dLp[2];
                for (int64_t k = r0 + last - 1; k >= r0; k--) {
                    uint32_t id = h->point_list[k];
                    real dx = h->xy[2 * (size_t)id] - (real)x, dy = h->xy[2 * (size_t)id + 1] - (real)y;
                    const real *co = h->conic_opacity + 4 * (size_t)id;
                    real power = RL(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0) continue;
                    real G = R_EXP(power);
                    real alpha = rmin(RL(0.99), co[3] * G);
                    if (alpha < inv255) continue;
                    T = T / (RL(1.0) - alpha);
                    real w = alpha * T;
                    real dL_dalpha = 0;
                    const real *col = h->rgb + 3 * (size_t)id;
                    for (int c = 0; c < 3; c++) {
                        accum[c] = last_alpha * last_color[c] + (RL(1.0) - last_alpha) * accum[c];
                        last_color[c] = col[c];
                        dL_dalpha += (col[c] - accum[c]) * dLp[c];
                        g_col[3 * (size_t)id + c] += w * dLp[c];
                    }
                    if (dL_ddepthpix) {
                        accum_d = last_alpha * last_d + (RL(1.0) - last_alpha) * accum_d;
                        last_d = h->depths[id];
                        dL_dalpha += (last_d - accum_d) * dLd;
                        g_z[id] += w * dLd;
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (RL(1.0) - alpha)) * bg_dot;
                    real dL_dG = co[3] * dL_dalpha; /* straight through the 0.99 clamp */
                    real gdx = G * dx, gdy = G * dy;
                    real dG_ddelx = -gdx * co[0] - gdy * co[1];
                    real dG_ddely = -gdy * co[2] - gdx * co[1];
                    g_mean2D[2 * (size_t)id] += dL_dG * dG_ddelx * ddelx_dx;
                    g_mean2D[2 * (size_t)id + 1] += dL_dG * dG_ddely * ddely_dy;
                    g_conic[3 * (size_t)id] += RL(-0.5) * gdx * dx * dL_dG;
                    g_conic[3 * (size_t)id + 1] += RL(-0.5) * gdx * dy * dL_dG;
                    g_conic[3 * (size_t)id + 2] += RL(-0.5) * gdy * dy * dL_dG;
                    g_opac[id] += G * dL_dalpha;
                }
            }
    }

    int deg = p->sh_degree < p->sh_eval_max_degree ? p->sh_degree : p->sh_eval_max_degree;
    if (dL_dmeans3D) memset(dL_dmeans3D, 0, sizeof(real) * (size_t)P * 3);
    if (dL_dmeans2D) memset(dL_dmeans2D, 0, sizeof(real) * (size_t)P * 3);
    if (dL_dsh) memset(dL_dsh, 0, sizeof(real) * (size_t)P * 3 * (size_t)p->M);
    if (dL_dcolors_out) memset(dL_dcolors_out, 0, sizeof(real) * (size_t)P * 3);
    if (dL_dopacity_out) memset(dL_dopacity_out, 0, sizeof(real) * (size_t)P);
    if (dL_dscales) memset(dL_dscales, 0, sizeof(real) * (size_t)P * 3);
    if (dL_drot) memset(dL_drot, 0, sizeof(real) * (size_t)P * 4);
    if (dL_dcov3D_out) memset(dL_dcov3D_out, 0, sizeof(real) * (size_t)P * 6);

#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (h->radii[i] <= 0) continue;
        const real *m = h->means3D + 3 * (size_t)i;
        const real *c6 = h->cov3D + 6 * (size_t)i;
        real gmean[3] = {0, 0, 0};
        real gcov[6] = {0, 0, 0, 0, 0, 0};
        /* -- conic -> cov2D -> cov3D and mean (through J) -- */
        {
            proj_jac j;
            build_jac(p, m, &j);
            real s0[3], s1[3];
            sym6_mul(c6, j.m0, s0);
            sym6_mul(c6, j.m1, s1);
            real a = j.m0[0] * s0[0] + j.m0[1] * s0[1] + j.m0[2] * s0[2] + p->dilation;
            real b = j.m0[0] * s1[0] + j.m0[1] * s1[1] + j.m0[2] * s1[2];
            real c = j.m1[0] * s1[0] + j.m1[1] * s1[1] + j.m1[2] * s1[2] + p->dilation;
            real denom = a * c - b * b;
            real d2inv = RL(1.0) / (denom * denom + RL(0.0000001));
            real gcx = g_conic[3 * (size_t)i], gcy = g_conic[3 * (size_t)i + 1], gcz = g_conic[3 * (size_t)i + 2];
            real dL_da = 0, dL_db = 0, dL_dc = 0;
            if (d2inv != 0) {
                dL_da = d2inv * (-c * c * gcx + RL(2.0) * b * c * gcy + (denom - a * c) * gcz);
                dL_dc = d2inv * (-a * a * gcz + RL(2.0) * a * b * gcy + (denom - a * c) * gcx);
                dL_db = d2inv * RL(2.0) * (b * c * gcx - (denom + RL(2.0) * b * b) * gcy + a * b * gcz);
                const real *m0 = j.m0, *m1 = j.m1;
                gcov[0] = m0[0] * m0[0] * dL_da + m0[0] * m1[0] * dL_db + m1[0] * m1[0] * dL_dc;
                gcov[3] = m0[1] * m0[1] * dL_da + m0[1] * m1[1] * dL_db + m1[1] * m1[1] * dL_dc;
                gcov[5] = m0[2] * m0[2] * dL_da + m0[2] * m1[2] * dL_db + m1[2] * m1[2] * dL_dc;
                gcov[1] = RL(2.0) * m0[0] * m0[1] * dL_da + (m0[0] * m1[1] + m0[1] * m1[0]) * dL_db + RL(2.0) * m1[0] * m1[1] * dL_dc;
                gcov[2] = RL(2.0) * m0[0] * m0[2] * dL_da + (m0[0] * m1[2] + m0[2] * m1[0]) * dL_db + RL(2.0) * m1[0] * m1[2] * dL_dc;
                gcov[4] = RL(2.0) * m0[2] * m0[1] * dL_da + (m0[1] * m1[2] + m0[2] * m1[1]) * dL_db + RL(2.0) * m1[1] * m1[2] * dL_dc;
            }
            /* dL/dM rows */
            real gm0[3], gm1[3];
            for (int k = 0; k < 3; k++) {
                gm0[k] = RL(2.0) * dL_da * s0[k] + dL_db * s1[k];
                gm1[k] = RL(2.0) * dL_dc * s1[k] + dL_db * s0[k];
            }
            /* dL/dJ_kl = sum_j dL/dM_kj * Rview[l][j];  Rview[l][j] = view[j*4+l] */
            real dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
            for (int jj = 0; jj < 3; jj++) {
                dJ00 += gm0[jj] * p->view[jj * 4 + 0];
                dJ02 += gm0[jj] * p->view[jj * 4 + 2];
                dJ11 += gm1[jj] * p->view[jj * 4 + 1];
                dJ12 += gm1[jj] * p->view[jj * 4 + 2];
            }
            real tz = RL(1.0) / j.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
            real dtx = (real)j.xmask * -j.fx * tz2 * dJ02;
            real dty = (real)j.ymask * -j.fy * tz2 * dJ12;
            real dtz = -j.fx * tz2 * dJ00 - j.fy * tz2 * dJ11 + (RL(2.0) * j.fx * j.t[0]) * tz3 * dJ02 +
                       (RL(2.0) * j.fy * j.t[1]) * tz3 * dJ12;
            /* dL/dmean = Rview^T dL/dt ; Rview^T[jj][l] = view[jj*4+l] */
            for (int jj = 0; jj < 3; jj++)
                gmean[jj] += p->view[jj * 4 + 0] * dtx + p->view[jj * 4 + 1] * dty + p->view[jj * 4 + 2] * dtz;
        }
        /* -- mean2D (NDC units) -> mean3D through the perspective divide -- */
        {
            real mh[4];
            xform4x4(p->proj, m, mh);
            real mw = RL(1.0) / (mh[3] + RL(0.0000001));
            real mul1 = mh[0] * mw * mw, mul2 = mh[1] * mw * mw;
            real gxn = g_mean2D[2 * (size_t)i], gyn = g_mean2D[2 * (size_t)i + 1];
            const real *pr = p->proj;
            gmean[0] += (pr[0] * mw - pr[3] * mul1) * gxn + (pr[1] * mw - pr[3] * mul2) * gyn;
            gmean[1] += (pr[4] * mw - pr[7] * mul1) * gxn + (pr[5] * mw - pr[7] * mul2) * gyn;
            gmean[2] += (pr[8] * mw - pr[11] * mul1) * gxn + (pr[9] * mw - pr[11] * mul2) * gyn;
        }
        /* -- fused depth channel: z_i = (view * mean).z -- */
        if (dL_ddepthpix) {
            gmean[0] += p->view[2] * g_z[i];
            gmean[1] += p->view[6] * g_z[i];
            gmean[2] += p->view[10] * g_z[i];
        }
        /* -- colour -> SH coefficients and mean (view direction) -- */
        if (h->shs) {
            const real *sh = h->shs + (size_t)i * 3 * (size_t)p->M;
            real d[3] = {m[0] - p->campos[0], m[1] - p->campos[1], m[2] - p->campos[2]};
            real len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            real inv = RL(1.0) / R_SQRT(len2);
            real x = d[0] * inv, y = d[1] * inv, z = d[2] * inv;
            real gdir[3] = {0, 0, 0};
            for (int c = 0; c < 3; c++) {
                real g = h->clamped[3 * (size_t)i + c] ? RL(0.0) : g_col[3 * (size_t)i + c];
#define SH(k) sh[(k) * 3 + c]
#define GSH(k) dL_dsh[((size_t)i * (size_t)p->M + (k)) * 3 + c]
                real dx_ = 0, dy_ = 0, dz_ = 0;
                if (dL_dsh) GSH(0) = SH_C0 * g;
                if (deg > 0) {
                    if (dL_dsh) { GSH(1) = -SH_C1 * y * g; GSH(2) = SH_C1 * z * g; GSH(3) = -SH_C1 * x * g; }
                    dx_ = -SH_C1 * SH(3); dy_ = -SH_C1 * SH(1); dz_ = SH_C1 * SH(2);
                    if (deg > 1) {
                        real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        if (dL_dsh) {
                            GSH(4) = SH_C2[0] * xy * g; GSH(5) = SH_C2[1] * yz * g;
                            GSH(6) = SH_C2[2] * (RL(2.0) * zz - xx - yy) * g;
                            GSH(7) = SH_C2[3] * xz * g; GSH(8) = SH_C2[4] * (xx - yy) * g;
                        }
                        dx_ += SH_C2[0] * y * SH(4) + SH_C2[2] * RL(2.0) * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * RL(2.0) * x * SH(8);
                        dy_ += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * RL(2.0) * -y * SH(6) + SH_C2[4] * RL(2.0) * -y * SH(8);
                        dz_ += SH_C2[1] * y * SH(5) + SH_C2[2] * RL(4.0) * z * SH(6) + SH_C2[3] * x * SH(7);
                        if (deg > 2) {
                            if (dL_dsh) {
                                GSH(9) = SH_C3[0] * y * (RL(3.0) * xx - yy) * g;
                                GSH(10) = SH_C3[1] * xy * z * g;
                                GSH(11) = SH_C3[2] * y * (RL(4.0) * zz - xx - yy) * g;
                                GSH(12) = SH_C3[3] * z * (RL(2.0) * zz - RL(3.0) * xx - RL(3.0) * yy) * g;
                                GSH(13) = SH_C3[4] * x * (RL(4.0) * zz - xx - yy) * g;
                                GSH(14) = SH_C3[5] * z * (xx - yy) * g;
                                GSH(15) = SH_C3[6] * x * (xx - RL(3.0) * yy) * g;
                            }
                            dx_ += SH_C3[0] * SH(9) * RL(6.0) * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * RL(-2.0) * xy +
                                   SH_C3[3] * SH(12) * RL(-6.0) * xz + SH_C3[4] * SH(13) * (RL(-3.0) * xx + RL(4.0) * zz - yy) +
                                   SH_C3[5] * SH(14) * RL(2.0) * xz + SH_C3[6] * SH(15) * RL(3.0) * (xx - yy);
                            dy_ += SH_C3[0] * SH(9) * RL(3.0) * (xx - yy) + SH_C3[1] * SH(10) * xz +
                                   SH_C3[2] * SH(11) * (RL(-3.0) * yy + RL(4.0) * zz - xx) + SH_C3[3] * SH(12) * RL(-6.0) * yz +
                                   SH_C3[4] * SH(13) * RL(-2.0) * xy + SH_C3[5] * SH(14) * RL(-2.0) * yz +
                                   SH_C3[6] * SH(15) * RL(-6.0) * xy;
                            dz_ += SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * RL(8.0) * yz +
                                   SH_C3[3] * SH(12) * RL(3.0) * (RL(2.0) * zz - xx - yy) + SH_C3[4] * SH(13) * RL(8.0) * xz +
                                   SH_C3[5] * SH(14) * (xx - yy);
                        }
                    }
                }
#undef SH
#undef GSH
                gdir[0] += dx_ * g; gdir[1] += dy_ * g; gdir[2] += dz_ * g;
            }
            /* through dir = d/|d| */
            real dot = d[0] * gdir[0] + d[1] * gdir[1] + d[2] * gdir[2];
            real inv3 = inv * inv * inv;
            for (int k = 0; k < 3; k++) gmean[k] += (gdir[k] * len2 - d[k] * dot) * inv3;
        } else if (dL_dcolors_out) {
            for (int c = 0; c < 3; c++) dL_dcolors_out[3 * (size_t)i + c] = g_col[3 * (size_t)i + c];
        }
        /* -- cov3D -> scales / rotations -- */
        if (h->scales) {
            const real *s = h->scales + 3 * (size_t)i;
            const real *q = h->rotations + 4 * (size_t)i;
            real mod = p->scale_modifier;
            real r = q[0], x = q[1], y = q[2], z = q[3];
            real R[3][3] = {{RL(1.0) - RL(2.0) * (y * y + z * z), RL(2.0) * (x * y - r * z), RL(2.0) * (x * z + r * y)},
                            {RL(2.0) * (x * y + r * z), RL(1.0) - RL(2.0) * (x * x + z * z), RL(2.0) * (y * z - r * x)},
                            {RL(2.0) * (x * z - r * y), RL(2.0) * (y * z + r * x), RL(1.0) - RL(2.0) * (x * x + y * y)}};
            /* full symmetric dL/dSigma */
            real G[3][3] = {{gcov[0], RL(0.5) * gcov[1], RL(0.5) * gcov[2]},
                            {RL(0.5) * gcov[1], gcov[3], RL(0.5) * gcov[4]},
                            {RL(0.5) * gcov[2], RL(0.5) * gcov[4], gcov[5]}};
            /* Sigma = R diag(v) R^T, v_k = (mod*s_k)^2.
             * dL/dv_k = (R^T G R)_kk ; dL/dR = 2 G R diag(v) */
            real sv[3] = {mod * s[0], mod * s[1], mod * s[2]};
            real GR[3][3];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) GR[a][b] = G[a][0] * R[0][b] + G[a][1] * R[1][b] + G[a][2] * R[2][b];
            real dR[3][3];
            for (int k = 0; k < 3; k++) {
                real dv = R[0][k] * GR[0][k] + R[1][k] * GR[1][k] + R[2][k] * GR[2][k];
                if (dL_dscales) dL_dscales[3 * (size_t)i + k] = dv * RL(2.0) * sv[k] * mod;
                for (int a = 0; a < 3; a++) dR[a][k] = RL(2.0) * GR[a][k] * sv[k] * sv[k];
            }
            if (dL_drot) {
                real *gq = dL_drot + 4 * (size_t)i;
                gq[0] = RL(2.0) * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
                gq[1] = RL(2.0) * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - RL(2.0) * x * dR[1][1] - r * dR[1][2] +
                                   z * dR[2][0] + r * dR[2][1] - RL(2.0) * x * dR[2][2]);
                gq[2] = RL(2.0) * (RL(-2.0) * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] -
                                   r * dR[2][0] + z * dR[2][1] - RL(2.0) * y * dR[2][2]);
                gq[3] = RL(2.0) * (RL(-2.0) * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - RL(2.0) * z * dR[1][1] +
                                   y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
            }
        } else if (dL_dcov3D_out) {
            memcpy(dL_dcov3D_out + 6 * (size_t)i, gcov, sizeof(real) * 6);
        }
        if (dL_dmeans3D) memcpy(dL_dmeans3D + 3 * (size_t)i, gmean, sizeof(real) * 3);
        if (dL_dmeans2D) {
            dL_dmeans2D[3 * (size_t)i] = g_mean2D[2 * (size_t)i];
            dL_dmeans2D[3 * (size_t)i + 1] = g_mean2D[2 * (size_t)i + 1];
        }
        if (dL_dopacity_out) dL_dopacity_out[i] = g_opac[i];
    }
    free(g_mean2D); free(g_conic); free(g_opac); free(g_col); free(g_z);
}
