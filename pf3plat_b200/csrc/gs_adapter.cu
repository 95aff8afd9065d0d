// gs_adapter.cu -- fused Gaussian adapter: the per-Gaussian arithmetic of
// /root/reference/src/model/encoder/common/gaussian_adapter.py:48-98 (GaussianAdapter.forward) in ONE kernel per
// direction, writing the harmonics directly in the (d_sh, xyz) layout the rasterizer reads (the relayout +
// .contiguous() copy of /root/reference/src/model/decoder/cuda_splatting.py:75 disappears).
//
// Per Gaussian g of camera v (reference line in brackets):
//   scales    = (smin + (smax - smin) * sigmoid(raw[0:3])) * depth * multiplier[v]                  [:62-70]
//   rotations = raw[3:7] / (|raw[3:7]| + eps)                                                       [:73]
//   sh        = raw[7:].view(3, d_sh) * sh_mask                                                     [:76-77]
//   cov       = C (Rq S S^T Rq^T) C^T, Rq = quaternion_to_matrix(rotations) (xyzw), C = c2w[:3,:3]  [:80-82]
//   means     = c2w[:3,3] + C normalize(Kinv (x, y, 1)) * depth                                     [:86-87]
//   harmonics = D_v sh  (block-diagonal Wigner-D of C, computed by the caller; NULL = identity)     [:92]
// C is detached in cov and harmonics (:81), not in means.  The tiny per-camera quantities (Kinv, multiplier, D)
// stay with the caller, which keeps them differentiable with ordinary tensor ops; the backward kernel returns the
// per-camera sums dL/dKinv, dL/dc2w, dL/dmultiplier.
//
// B200 design: one thread per Gaussian, 128 per CTA, one camera per blockIdx.y.  The CTA's contiguous block of raw
// rows (128 x 328 B) arrives by one bulk TMA copy; the harmonics block (128 x 300 B) is assembled in shared memory and
// leaves by one bulk TMA store, so both big streams are fully coalesced; the backward mirrors it (gradient of the
// harmonics in by TMA, gradient of the raw rows out by TMA).  HBM-bound: 328 + 12 B in, 300 + 76 B out per Gaussian.
#include "gs_common.cuh"

namespace {

struct AdapterArgs {
    int V, R, d_sh, C;  // C = 7 + 3 * d_sh
    float smin, smax, eps;
    const float *c2w, *kinv, *mult, *sh_rot, *sh_mask;
    const float *coords, *depths, *raw;
};
struct AdapterOut {
    float *means, *cov, *harm, *scales, *rot;
};
struct AdapterOutGrads {
    const float *means, *cov, *harm, *scales, *rot;
};
struct AdapterInGrads {
    float *coords, *depths, *raw, *c2w, *kinv, *mult;
};

constexpr int AD_THREADS = 128;
constexpr int AD_MAX_DSH = 25;
constexpr float AD_QUAT_EPS = 1e-8f;  // quaternion_to_matrix's own eps (/root/reference/src/model/encoder/common/gaussians.py:10)

struct AdHdr {
    float kinv[9], rot[9], t[3], mult;
    float mask[AD_MAX_DSH];
    float red[AD_THREADS / 32][24];
    uint64_t bar;
};
constexpr int AD_HDR_BYTES = (sizeof(AdHdr) + 127) / 128 * 128;

__device__ __forceinline__ void load_camera(const AdapterArgs &a, int v, AdHdr *h) {
    const int tid = threadIdx.x;
    if (tid < 9) {
        h->kinv[tid] = a.kinv[(size_t)v * 9 + tid];
        h->rot[tid] = a.c2w[(size_t)v * 16 + (tid / 3) * 4 + tid % 3];
    } else if (tid < 12) {
        h->t[tid - 9] = a.c2w[(size_t)v * 16 + (tid - 9) * 4 + 3];
    } else if (tid == 12) {
        h->mult = a.mult[v];
    }
    if (tid >= 32 && tid < 32 + a.d_sh) h->mask[tid - 32] = a.sh_mask[tid - 32];
}

// The geometry part of one Gaussian (everything but the harmonics); kept by the backward for its chain rule.
struct AdGeom {
    float dc[3], inv_n;   // camera-space unit direction, 1 / |Kinv (x,y,1)|
    float dw[3];          // world-space direction
    float sg[3], s0[3], s[3];
    float q[4], nq, A;    // normalised quaternion (i,j,k,r), |raw quaternion|, two_s
    float Rq[9];
};

__device__ __forceinline__ void adapter_geometry(const AdHdr &h, float x, float y, float z, const float *sraw, const float *qraw,
                                                 float smin, float smax, float eps, AdGeom &g) {
    float u[3];
#pragma unroll
    for (int i = 0; i < 3; i++) u[i] = h.kinv[i * 3] * x + h.kinv[i * 3 + 1] * y + h.kinv[i * 3 + 2];
    g.inv_n = 1.0f / sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
#pragma unroll
    for (int i = 0; i < 3; i++) g.dc[i] = u[i] * g.inv_n;
#pragma unroll
    for (int i = 0; i < 3; i++) g.dw[i] = h.rot[i * 3] * g.dc[0] + h.rot[i * 3 + 1] * g.dc[1] + h.rot[i * 3 + 2] * g.dc[2];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        g.sg[k] = 1.0f / (1.0f + expf(-sraw[k]));
        g.s0[k] = smin + (smax - smin) * g.sg[k];
        g.s[k] = g.s0[k] * z * h.mult;
    }
    g.nq = sqrtf(qraw[0] * qraw[0] + qraw[1] * qraw[1] + qraw[2] * qraw[2] + qraw[3] * qraw[3]);
    const float inv = 1.0f / (g.nq + eps);
#pragma unroll
    for (int k = 0; k < 4; k++) g.q[k] = qraw[k] * inv;
    const float i = g.q[0], j = g.q[1], k = g.q[2], r = g.q[3];
    g.A = 2.0f / (i * i + j * j + k * k + r * r + AD_QUAT_EPS);
    const float A = g.A;
    g.Rq[0] = 1.0f - A * (j * j + k * k); g.Rq[1] = A * (i * j - k * r);        g.Rq[2] = A * (i * k + j * r);
    g.Rq[3] = A * (i * j + k * r);        g.Rq[4] = 1.0f - A * (i * i + k * k); g.Rq[5] = A * (j * k - i * r);
    g.Rq[6] = A * (i * k - j * r);        g.Rq[7] = A * (j * k + i * r);        g.Rq[8] = 1.0f - A * (i * i + j * j);
}

// N = C Rq diag(s): the world-space covariance is N N^T
__device__ __forceinline__ void adapter_factor(const AdHdr &h, const AdGeom &g, float *N) {
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) M[i * 3 + k] = g.Rq[i * 3 + k] * g.s[k];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++)
            N[i * 3 + k] = h.rot[i * 3] * M[k] + h.rot[i * 3 + 1] * M[3 + k] + h.rot[i * 3 + 2] * M[6 + k];
}

__device__ __forceinline__ int isqrt_small(int n) {
    int r = 0;
    while ((r + 1) * (r + 1) <= n) r++;
    return r;
}

__global__ void __launch_bounds__(AD_THREADS) k_adapter_fwd(const AdapterArgs a, const AdapterOut o) {
    extern __shared__ __align__(128) unsigned char ad_smem[];
    AdHdr *h = reinterpret_cast<AdHdr *>(ad_smem);
    float *raw_s = reinterpret_cast<float *>(ad_smem + AD_HDR_BYTES);  // [AD_THREADS][C]
    float *harm_s = raw_s + (size_t)AD_THREADS * a.C;                   // [AD_THREADS][3 * d_sh]

    const int v = blockIdx.y, g0 = blockIdx.x * AD_THREADS, tid = threadIdx.x;
    const int n = min(AD_THREADS, a.R - g0);
    const size_t first = (size_t)v * a.R + g0;
    const bool active = tid < n;
    const int HS = 3 * a.d_sh;

    const float *src = a.raw + first * a.C;
    const uint32_t in_bytes = (uint32_t)n * a.C * 4u;
    const bool bulk_in = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((in_bytes & 15u) == 0);
    if (bulk_in) {
        if (tid == 0) {
            mbar_init(&h->bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&h->bar, in_bytes);
            tma_load_1d(raw_s, src, in_bytes, &h->bar);
        }
    } else {
        for (uint32_t k = tid; k < (uint32_t)n * a.C; k += AD_THREADS) raw_s[k] = src[k];
    }
    load_camera(a, v, h);
    float x = 0.f, y = 0.f, z = 0.f;
    if (active) {
        x = a.coords[(first + tid) * 2];
        y = a.coords[(first + tid) * 2 + 1];
        z = a.depths[first + tid];
    }
    __syncthreads();
    if (bulk_in) mbar_wait(&h->bar, 0);

    if (active) {
        const float *row = raw_s + (size_t)tid * a.C;
        AdGeom g;
        adapter_geometry(*h, x, y, z, row, row + 3, a.smin, a.smax, a.eps, g);
        float N[9];
        adapter_factor(*h, g, N);
        const size_t gi = first + tid;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            o.means[gi * 3 + i] = h->t[i] + g.dw[i] * z;
            o.scales[gi * 3 + i] = g.s[i];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) o.rot[gi * 4 + k] = g.q[k];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                o.cov[gi * 9 + i * 3 + j] = N[i * 3] * N[j * 3] + N[i * 3 + 1] * N[j * 3 + 1] + N[i * 3 + 2] * N[j * 3 + 2];
        // harmonics: out[d][c] = sum_j D[d][j] * raw[c][j] * mask[j], D block-diagonal by degree
        float *out = harm_s + (size_t)tid * HS;
        const float *sh = row + 7;
        if (a.sh_rot) {
            const float *D = a.sh_rot + (size_t)v * a.d_sh * a.d_sh;  // same address across the CTA: L1 broadcast
            const int degs = isqrt_small(a.d_sh);
            for (int l = 0; l < degs; l++) {
                const int b = l * l, w = 2 * l + 1;
                for (int i = 0; i < w; i++) {
                    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
                    for (int j = 0; j < w; j++) {
                        const float d = D[(b + i) * a.d_sh + b + j] * h->mask[b + j];
                        acc0 = fmaf(d, sh[b + j], acc0);
                        acc1 = fmaf(d, sh[a.d_sh + b + j], acc1);
                        acc2 = fmaf(d, sh[2 * a.d_sh + b + j], acc2);
                    }
                    out[(b + i) * 3] = acc0;
                    out[(b + i) * 3 + 1] = acc1;
                    out[(b + i) * 3 + 2] = acc2;
                }
            }
        } else {
            for (int d = 0; d < a.d_sh; d++) {
                const float m = h->mask[d];
                out[d * 3] = sh[d] * m;
                out[d * 3 + 1] = sh[a.d_sh + d] * m;
                out[d * 3 + 2] = sh[2 * a.d_sh + d] * m;
            }
        }
    }
    float *dst = o.harm + first * HS;
    const uint32_t out_bytes = (uint32_t)n * HS * 4u;
    const bool bulk_out = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((out_bytes & 15u) == 0);
    if (bulk_out) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tma_store_1d(dst, harm_s, out_bytes);
            tma_store_commit_wait();
        }
    } else {
        __syncthreads();
        for (uint32_t k = tid; k < (uint32_t)n * HS; k += AD_THREADS) dst[k] = harm_s[k];
    }
}

__global__ void __launch_bounds__(AD_THREADS)
k_adapter_bwd(const AdapterArgs a, const AdapterOutGrads go, const AdapterInGrads gi_) {
    extern __shared__ __align__(128) unsigned char ad_smem[];
    AdHdr *h = reinterpret_cast<AdHdr *>(ad_smem);
    float *draw_s = reinterpret_cast<float *>(ad_smem + AD_HDR_BYTES);  // [AD_THREADS][C]   gradient of the raw rows
    float *gh_s = draw_s + (size_t)AD_THREADS * a.C;                    // [AD_THREADS][3 * d_sh] incoming dL/dharmonics

    const int v = blockIdx.y, g0 = blockIdx.x * AD_THREADS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(AD_THREADS, a.R - g0);
    const size_t first = (size_t)v * a.R + g0;
    const bool active = tid < n;
    const int HS = 3 * a.d_sh;

    bool bulk_in = false;
    if (go.harm) {
        const float *src = go.harm + first * HS;
        const uint32_t in_bytes = (uint32_t)n * HS * 4u;
        bulk_in = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((in_bytes & 15u) == 0);
        if (bulk_in) {
            if (tid == 0) {
                mbar_init(&h->bar, 1);
                mbar_fence_init();
                mbar_expect_tx(&h->bar, in_bytes);
                tma_load_1d(gh_s, src, in_bytes, &h->bar);
            }
        } else {
            for (uint32_t k = tid; k < (uint32_t)n * HS; k += AD_THREADS) gh_s[k] = src[k];
        }
    }
    load_camera(a, v, h);
    float x = 0.f, y = 0.f, z = 0.f, sraw[3] = {0, 0, 0}, qraw[4] = {0, 0, 0, 1};
    const size_t gi = first + tid;
    if (active) {
        x = a.coords[gi * 2];
        y = a.coords[gi * 2 + 1];
        z = a.depths[gi];
        const float *row = a.raw + gi * a.C;
#pragma unroll
        for (int k = 0; k < 3; k++) sraw[k] = row[k];
#pragma unroll
        for (int k = 0; k < 4; k++) qraw[k] = row[3 + k];
    }
    __syncthreads();
    if (bulk_in) mbar_wait(&h->bar, 0);

    // per-camera sums: [0..8] dKinv, [9..17] dC (through the means only), [18..20] dt, [21] dmultiplier
    float cam[22];
#pragma unroll
    for (int k = 0; k < 22; k++) cam[k] = 0.f;

    if (active) {
        AdGeom g;
        adapter_geometry(*h, x, y, z, sraw, qraw, a.smin, a.smax, a.eps, g);
        float *drow = draw_s + (size_t)tid * a.C;
        float dz = 0.f;
        // ---- means = t + dw * z ----
        float gm[3] = {0, 0, 0};
        if (go.means) {
#pragma unroll
            for (int i = 0; i < 3; i++) gm[i] = go.means[gi * 3 + i];
        }
        float ddc[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 3; i++) {
            cam[18 + i] = gm[i];
            const float ddw = gm[i] * z;
            dz = fmaf(gm[i], g.dw[i], dz);
#pragma unroll
            for (int j = 0; j < 3; j++) {
                cam[9 + i * 3 + j] = ddw * g.dc[j];
                ddc[j] = fmaf(h->rot[i * 3 + j], ddw, ddc[j]);
            }
        }
        const float dot = ddc[0] * g.dc[0] + ddc[1] * g.dc[1] + ddc[2] * g.dc[2];
        float du[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            du[i] = (ddc[i] - g.dc[i] * dot) * g.inv_n;
            cam[i * 3] = du[i] * x;
            cam[i * 3 + 1] = du[i] * y;
            cam[i * 3 + 2] = du[i];
        }
        if (gi_.coords) {
            gi_.coords[gi * 2] = h->kinv[0] * du[0] + h->kinv[3] * du[1] + h->kinv[6] * du[2];
            gi_.coords[gi * 2 + 1] = h->kinv[1] * du[0] + h->kinv[4] * du[1] + h->kinv[7] * du[2];
        }
        // ---- covariance = N N^T, N = C Rq diag(s) ----
        float ds[3] = {0, 0, 0}, dRq[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (go.cov) {
            float N[9], G[9], dN[9];
            adapter_factor(*h, g, N);
#pragma unroll
            for (int k = 0; k < 9; k++) G[k] = go.cov[gi * 9 + k];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    dN[i * 3 + k] = (G[i * 3] + G[i]) * N[k] + (G[i * 3 + 1] + G[3 + i]) * N[3 + k] +
                                    (G[i * 3 + 2] + G[6 + i]) * N[6 + k];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float dM = h->rot[i] * dN[k] + h->rot[3 + i] * dN[3 + k] + h->rot[6 + i] * dN[6 + k];  // C^T dN
                    ds[k] = fmaf(dM, g.Rq[i * 3 + k], ds[k]);
                    dRq[i * 3 + k] = dM * g.s[k];
                }
        }
        // ---- scales ----
        if (go.scales) {
#pragma unroll
            for (int k = 0; k < 3; k++) ds[k] += go.scales[gi * 3 + k];
        }
        float dmult = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dz = fmaf(ds[k], g.s0[k] * h->mult, dz);
            dmult = fmaf(ds[k], g.s0[k] * z, dmult);
            drow[k] = ds[k] * z * h->mult * (a.smax - a.smin) * g.sg[k] * (1.0f - g.sg[k]);
        }
        cam[21] = dmult;
        // ---- quaternion ----
        {
            const float i = g.q[0], j = g.q[1], k = g.q[2], r = g.q[3], A = g.A;
            const float *d = dRq;
            const float dA = -(j * j + k * k) * d[0] + (i * j - k * r) * d[1] + (i * k + j * r) * d[2] + (i * j + k * r) * d[3] -
                             (i * i + k * k) * d[4] + (j * k - i * r) * d[5] + (i * k - j * r) * d[6] + (j * k + i * r) * d[7] -
                             (i * i + j * j) * d[8];
            float dq[4];
            dq[0] = A * (j * d[1] + k * d[2] + j * d[3] - 2.f * i * d[4] - r * d[5] + k * d[6] + r * d[7] - 2.f * i * d[8]);
            dq[1] = A * (-2.f * j * d[0] + i * d[1] + r * d[2] + i * d[3] + k * d[5] - r * d[6] + k * d[7] - 2.f * j * d[8]);
            dq[2] = A * (-2.f * k * d[0] - r * d[1] + i * d[2] + r * d[3] - 2.f * k * d[4] + j * d[5] + i * d[6] + j * d[7]);
            dq[3] = A * (-k * d[1] + j * d[2] + k * d[3] - i * d[5] - j * d[6] + i * d[7]);
            const float back = -A * A * dA;  // dA/dq_m = -A^2 q_m
#pragma unroll
            for (int m = 0; m < 4; m++) dq[m] = fmaf(back, g.q[m], dq[m]);
            if (go.rot) {
#pragma unroll
                for (int m = 0; m < 4; m++) dq[m] += go.rot[gi * 4 + m];
            }
            // q = qraw / (|qraw| + eps)
            const float inv = 1.0f / (g.nq + a.eps);
            const float dotq = dq[0] * qraw[0] + dq[1] * qraw[1] + dq[2] * qraw[2] + dq[3] * qraw[3];
            const float radial = dotq * inv * inv / g.nq;
#pragma unroll
            for (int m = 0; m < 4; m++) drow[3 + m] = dq[m] * inv - qraw[m] * radial;
        }
        if (gi_.depths) gi_.depths[gi] = dz;
        // ---- harmonics: d raw[c][j] = mask[j] * sum_i D[i][j] g[i][c] ----
        float *dsh = drow + 7;
        if (!go.harm) {
            for (int k = 0; k < HS; k++) dsh[k] = 0.f;
        } else {
            const float *gh = gh_s + (size_t)tid * HS;
            if (a.sh_rot) {
                const float *D = a.sh_rot + (size_t)v * a.d_sh * a.d_sh;
                const int degs = isqrt_small(a.d_sh);
                for (int l = 0; l < degs; l++) {
                    const int b = l * l, w = 2 * l + 1;
                    for (int j = 0; j < w; j++) {
                        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
                        for (int i = 0; i < w; i++) {
                            const float d = D[(b + i) * a.d_sh + b + j];
                            acc0 = fmaf(d, gh[(b + i) * 3], acc0);
                            acc1 = fmaf(d, gh[(b + i) * 3 + 1], acc1);
                            acc2 = fmaf(d, gh[(b + i) * 3 + 2], acc2);
                        }
                        const float m = h->mask[b + j];
                        dsh[b + j] = acc0 * m;
                        dsh[a.d_sh + b + j] = acc1 * m;
                        dsh[2 * a.d_sh + b + j] = acc2 * m;
                    }
                }
            } else {
                for (int d = 0; d < a.d_sh; d++) {
                    const float m = h->mask[d];
                    dsh[d] = gh[d * 3] * m;
                    dsh[a.d_sh + d] = gh[d * 3 + 1] * m;
                    dsh[2 * a.d_sh + d] = gh[d * 3 + 2] * m;
                }
            }
        }
    }
    // ---- per-camera sums: warp reduce, then one atomic per value and CTA ----
#pragma unroll
    for (int k = 0; k < 22; k++) {
        float s = cam[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) h->red[warp][k] = s;
    }
    float *dst = gi_.raw ? gi_.raw + first * a.C : nullptr;
    const uint32_t out_bytes = (uint32_t)n * a.C * 4u;
    const bool bulk_out = dst && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((out_bytes & 15u) == 0);
    if (bulk_out) fence_proxy_async_smem();
    __syncthreads();
    if (tid < 22) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < AD_THREADS / 32; w++) s += h->red[w][tid];
        if (tid < 9) {
            if (gi_.kinv) atomicAdd(&gi_.kinv[(size_t)v * 9 + tid], s);
        } else if (tid < 18) {
            const int e = tid - 9;
            if (gi_.c2w) atomicAdd(&gi_.c2w[(size_t)v * 16 + (e / 3) * 4 + e % 3], s);
        } else if (tid < 21) {
            if (gi_.c2w) atomicAdd(&gi_.c2w[(size_t)v * 16 + (tid - 18) * 4 + 3], s);
        } else {
            if (gi_.mult) atomicAdd(&gi_.mult[v], s);
        }
    }
    if (dst) {
        if (bulk_out) {
            if (tid == 32) {
                tma_store_1d(dst, draw_s, out_bytes);
                tma_store_commit_wait();
            }
        } else {
            for (uint32_t k = tid; k < (uint32_t)n * a.C; k += AD_THREADS) dst[k] = draw_s[k];
        }
    }
}

constexpr size_t adapter_smem_bytes(int d_sh) { return AD_HDR_BYTES + (size_t)AD_THREADS * ((7 + 3 * d_sh) + 3 * d_sh) * 4; }

}  // namespace

static int adapter_args(const GsAdapterConfig *cfg, const GsAdapterInputs *in, AdapterArgs &a) {
    if (!cfg || !in) return gs_set_error(GS_ERR_INVALID, "null adapter config/inputs");
    if (cfg->V < 0 || cfg->R < 0) return gs_set_error(GS_ERR_INVALID, "negative adapter sizes");
    const int d = cfg->d_sh;
    if (d != 1 && d != 4 && d != 9 && d != 16 && d != 25)
        return gs_set_error(GS_ERR_INVALID, "d_sh must be (sh_degree + 1)^2 with sh_degree <= 4");
    if ((size_t)cfg->V * (size_t)cfg->R > 0) {
        if (!cfg->c2w || !cfg->kinv || !cfg->multiplier || !cfg->sh_mask)
            return gs_set_error(GS_ERR_INVALID, "adapter camera arrays missing");
        if (!in->coordinates || !in->depths || !in->raw_gaussians)
            return gs_set_error(GS_ERR_INVALID, "adapter inputs missing");
    }
    if (cfg->V > 65535) return gs_set_error(GS_ERR_INVALID, "more than 65535 cameras in one adapter call");
    a.V = cfg->V; a.R = cfg->R; a.d_sh = d; a.C = 7 + 3 * d;
    a.smin = cfg->scale_min; a.smax = cfg->scale_max; a.eps = cfg->eps;
    a.c2w = cfg->c2w; a.kinv = cfg->kinv; a.mult = cfg->multiplier; a.sh_rot = cfg->sh_rotation; a.sh_mask = cfg->sh_mask;
    a.coords = in->coordinates; a.depths = in->depths; a.raw = in->raw_gaussians;
    return GS_OK;
}

extern "C" GS_API int gs_adapter_forward(const GsAdapterConfig *cfg, const GsAdapterInputs *in, const GsAdapterOutputs *out,
                                         void *stream) {
    AdapterArgs a;
    int rc = adapter_args(cfg, in, a);
    if (rc != GS_OK) return rc;
    if (!out) return gs_set_error(GS_ERR_INVALID, "null adapter outputs");
    if ((size_t)a.V * a.R == 0) return GS_OK;
    if (!out->means || !out->covariances || !out->harmonics || !out->scales || !out->rotations)
        return gs_set_error(GS_ERR_INVALID, "adapter outputs missing");
    const AdapterOut o{out->means, out->covariances, out->harmonics, out->scales, out->rotations};
    const size_t smem = adapter_smem_bytes(a.d_sh);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GS_CUDA_OK(cudaFuncSetAttribute(k_adapter_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((a.R + AD_THREADS - 1) / AD_THREADS, a.V);
    k_adapter_fwd<<<grid, AD_THREADS, smem, st>>>(a, o);
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}

extern "C" GS_API int gs_adapter_backward(const GsAdapterConfig *cfg, const GsAdapterInputs *in, const GsAdapterOutGrads *gout,
                                          const GsAdapterInGrads *gin, void *stream) {
    AdapterArgs a;
    int rc = adapter_args(cfg, in, a);
    if (rc != GS_OK) return rc;
    if (!gout || !gin) return gs_set_error(GS_ERR_INVALID, "null adapter gradients");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (gin->c2w && a.V) GS_CUDA_OK(cudaMemsetAsync(gin->c2w, 0, (size_t)a.V * 16 * 4, st));
    if (gin->kinv && a.V) GS_CUDA_OK(cudaMemsetAsync(gin->kinv, 0, (size_t)a.V * 9 * 4, st));
    if (gin->multiplier && a.V) GS_CUDA_OK(cudaMemsetAsync(gin->multiplier, 0, (size_t)a.V * 4, st));
    if ((size_t)a.V * a.R == 0) return GS_OK;
    const AdapterOutGrads go{gout->means, gout->covariances, gout->harmonics, gout->scales, gout->rotations};
    const AdapterInGrads gi{gin->coordinates, gin->depths, gin->raw_gaussians, gin->c2w, gin->kinv, gin->multiplier};
    const size_t smem = adapter_smem_bytes(a.d_sh);
    GS_CUDA_OK(cudaFuncSetAttribute(k_adapter_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((a.R + AD_THREADS - 1) / AD_THREADS, a.V);
    k_adapter_bwd<<<grid, AD_THREADS, smem, st>>>(a, go, gi);
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}
