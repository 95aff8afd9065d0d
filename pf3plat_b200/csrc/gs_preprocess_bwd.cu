// gs_preprocess_bwd.cu -- stage 2 of the backward: per-(view, Gaussian) accumulators -> input gradients.
//
// Semantics: SURVEY.md Appendix A "Preprocess backward" (upstream computeCov2DCUDA + preprocessCUDA backward):
// conic -> cov2D (with 1/(det^2+1e-7)) -> cov3D(6) and mean through the clamped projection Jacobian; mean2D
// (NDC-scaled) -> mean through the perspective divide; colour -> SH coefficients and mean through the view
// direction, honouring the clamp flags; cov3D -> scales/rotations when those were the inputs.
//
// B200 design: one thread per Gaussian looping over the views of its scene and
// summing in registers / shared memory, so scene-level gradients are written ONCE, coalesced, with no atomics
// and no zero-fill pass (every output element is written, zeros included).  The SH block is staged in by one
// bulk TMA load and its gradient block leaves by one bulk TMA store.
#include "gs_common.cuh"

namespace {

constexpr int PB_THREADS = 128;
// Resident CTAs per SM the register allocation is bounded for.  The kernel is latency-bound; unbounded it takes 150
// registers (3 CTAs/SM, 18 % occupancy).  Measured on C2: 3 CTAs 0.240 ms, 4 CTAs (128 registers, 36 B of spills)
// 0.220 ms, 5 CTAs (96 registers, 304 B of spills inside the view loop) 0.339 ms.
constexpr int PB_MIN_CTAS = 4;

struct PbSmem {
    ViewCam cams[GS_CAM_CHUNK];
    uint64_t bar;
};
constexpr int PB_SMEM_HDR = (sizeof(PbSmem) + 127) / 128 * 128;

constexpr int PB_SH_EVAL = 16;  // coefficients per channel the evaluator can touch (bands 0..3)

// dL/dsh_k += basis_k(dir) * g  and  dL/ddir, for one colour channel.  sh: stride-3 view of this channel's
// coefficients in shared memory; gsh: stride-3 view of this channel's accumulators (REGISTERS: all indices are
// compile-time constants after inlining).
template <int GS = 3 /* stride of gsh */>
__device__ __forceinline__ void sh_backward_channel(int deg, const float *sh, float *gsh, float g, float x, float y,
                                                    float z, float &ddx, float &ddy, float &ddz) {
    gsh[0 * GS] += GS_SH_C0 * g;
    if (deg < 1) return;
    gsh[1 * GS] += -GS_SH_C1 * y * g;
    gsh[2 * GS] += GS_SH_C1 * z * g;
    gsh[3 * GS] += -GS_SH_C1 * x * g;
    float dx_ = -GS_SH_C1 * sh[3 * 3], dy_ = -GS_SH_C1 * sh[1 * 3], dz_ = GS_SH_C1 * sh[2 * 3];
    if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        gsh[4 * GS] += GS_SH_C2_0 * xy * g;
        gsh[5 * GS] += GS_SH_C2_1 * yz * g;
        gsh[6 * GS] += GS_SH_C2_2 * (2.0f * zz - xx - yy) * g;
        gsh[7 * GS] += GS_SH_C2_3 * xz * g;
        gsh[8 * GS] += GS_SH_C2_4 * (xx - yy) * g;
        dx_ += GS_SH_C2_0 * y * sh[4 * 3] + GS_SH_C2_2 * 2.0f * -x * sh[6 * 3] + GS_SH_C2_3 * z * sh[7 * 3] +
               GS_SH_C2_4 * 2.0f * x * sh[8 * 3];
        dy_ += GS_SH_C2_0 * x * sh[4 * 3] + GS_SH_C2_1 * z * sh[5 * 3] + GS_SH_C2_2 * 2.0f * -y * sh[6 * 3] +
               GS_SH_C2_4 * 2.0f * -y * sh[8 * 3];
        dz_ += GS_SH_C2_1 * y * sh[5 * 3] + GS_SH_C2_2 * 4.0f * z * sh[6 * 3] + GS_SH_C2_3 * x * sh[7 * 3];
        if (deg > 2) {
            gsh[9 * GS] += GS_SH_C3_0 * y * (3.0f * xx - yy) * g;
            gsh[10 * GS] += GS_SH_C3_1 * xy * z * g;
            gsh[11 * GS] += GS_SH_C3_2 * y * (4.0f * zz - xx - yy) * g;
            gsh[12 * GS] += GS_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * g;
            gsh[13 * GS] += GS_SH_C3_4 * x * (4.0f * zz - xx - yy) * g;
            gsh[14 * GS] += GS_SH_C3_5 * z * (xx - yy) * g;
            gsh[15 * GS] += GS_SH_C3_6 * x * (xx - 3.0f * yy) * g;
            dx_ += GS_SH_C3_0 * sh[9 * 3] * 6.0f * xy + GS_SH_C3_1 * sh[10 * 3] * yz +
                   GS_SH_C3_2 * sh[11 * 3] * -2.0f * xy + GS_SH_C3_3 * sh[12 * 3] * -6.0f * xz +
                   GS_SH_C3_4 * sh[13 * 3] * (-3.0f * xx + 4.0f * zz - yy) + GS_SH_C3_5 * sh[14 * 3] * 2.0f * xz +
                   GS_SH_C3_6 * sh[15 * 3] * 3.0f * (xx - yy);
            dy_ += GS_SH_C3_0 * sh[9 * 3] * 3.0f * (xx - yy) + GS_SH_C3_1 * sh[10 * 3] * xz +
                   GS_SH_C3_2 * sh[11 * 3] * (-3.0f * yy + 4.0f * zz - xx) + GS_SH_C3_3 * sh[12 * 3] * -6.0f * yz +
                   GS_SH_C3_4 * sh[13 * 3] * -2.0f * xy + GS_SH_C3_5 * sh[14 * 3] * -2.0f * yz +
                   GS_SH_C3_6 * sh[15 * 3] * -6.0f * xy;
            dz_ += GS_SH_C3_1 * sh[10 * 3] * xy + GS_SH_C3_2 * sh[11 * 3] * 8.0f * yz +
                   GS_SH_C3_3 * sh[12 * 3] * 3.0f * (2.0f * zz - xx - yy) + GS_SH_C3_4 * sh[13 * 3] * 8.0f * xz +
                   GS_SH_C3_5 * sh[14 * 3] * (xx - yy);
        }
    }
    ddx += dx_ * g;
    ddy += dy_ * g;
    ddz += dz_ * g;
}

template <bool HAS_SH, int MINB>
__global__ void __launch_bounds__(PB_THREADS, MINB)
k_preprocess_bwd(const DevCfg c, const DevInputs in, const uint8_t *__restrict__ meta, const float *__restrict__ acc,
                    const GsInGrads g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PbSmem *sm = reinterpret_cast<PbSmem *>(smem_raw);
    float *sh_s = reinterpret_cast<float *>(smem_raw + PB_SMEM_HDR);

    const int scene = blockIdx.y;
    const int g0 = blockIdx.x * PB_THREADS;
    const int n = min(PB_THREADS, c.P - g0);
    const int tid = threadIdx.x;
    const int i = g0 + tid;
    const bool active = tid < n;
    const size_t sg = (size_t)scene * c.P + i;
    const uint32_t sh_floats = (uint32_t)n * c.M * 3u;

    bool bulk = false;
    if (HAS_SH) {
        const float *src = in.shs + ((size_t)scene * c.P + g0) * c.M * 3;
        const float *dst = g.dL_dshs ? g.dL_dshs + ((size_t)scene * c.P + g0) * c.M * 3 : nullptr;
        bulk = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && (((sh_floats * 4u) & 15u) == 0) &&
               ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
        if (bulk) {
            if (tid == 0) {
                mbar_init(&sm->bar, 1);
                mbar_fence_init();
                mbar_expect_tx(&sm->bar, sh_floats * 4u);
                tma_load_1d(sh_s, src, sh_floats * 4u, &sm->bar);
            }
        } else {
            for (uint32_t k = tid; k < sh_floats; k += PB_THREADS) sh_s[k] = src[k];
        }
    }
    float gsh[PB_SH_EVAL * 3];
#pragma unroll
    for (int k = 0; k < PB_SH_EVAL * 3; k++) gsh[k] = 0.f;

    float3 mean = make_float3(0, 0, 0);
    float c6[6] = {0, 0, 0, 0, 0, 0};
    float sc[3] = {0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (active) {
        mean = make_float3(in.means3D[sg * 3 + 0], in.means3D[sg * 3 + 1], in.means3D[sg * 3 + 2]);
        if (in.cov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = in.cov3D[sg * 6 + k];
        } else {
#pragma unroll
            for (int k = 0; k < 3; k++) sc[k] = in.scales[sg * 3 + k];
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = in.rotations[sg * 4 + k];
            cov3d_from_scale_rot(sc, c.scale_modifier, q, c6);
        }
    }
    if (HAS_SH) {
        __syncthreads();
        if (bulk) mbar_wait(&sm->bar, 0);
    }

    float gmean[3] = {0, 0, 0}, gcov[6] = {0, 0, 0, 0, 0, 0}, gopac = 0.f;

    for (int vi = 0; vi < c.VPS; vi++) {
        const int v = scene * c.VPS + vi;
        if (vi % GS_CAM_CHUNK == 0) {
            __syncthreads();
            load_view_cams(c, v, min(GS_CAM_CHUNK, c.VPS - vi), sm->cams);
            __syncthreads();
        }
        if (!active) continue;
        const ViewCam &cam = sm->cams[vi % GS_CAM_CHUNK];
        const size_t o = (size_t)v * c.P + i;
        const uint32_t mb = meta[o];
        float m2d[2] = {0.f, 0.f}, gcol[3] = {0.f, 0.f, 0.f};
        if (mb & GS_META_VISIBLE) {
            const float *a = acc + o * GS_ACC_STRIDE;
            gcol[0] = a[0]; gcol[1] = a[1]; gcol[2] = a[2];
            m2d[0] = a[3]; m2d[1] = a[4];
            const float gcx = a[5], gcy = a[6], gcz = a[7];
            gopac += a[8];
            const float gz = a[9];
            const float s = cam.scale, s2 = s * s;
            const float3 m = make_float3(mean.x * s, mean.y * s, mean.z * s);
            const float cv[6] = {c6[0] * s2, c6[1] * s2, c6[2] * s2, c6[3] * s2, c6[4] * s2, c6[5] * s2};
            float gm[3] = {0.f, 0.f, 0.f};  // dL/dm (view-scaled mean)
            // ---- conic -> cov2D -> cov3D, mean (through J) ----
            {
                ProjJac j;
                build_jac(cam, c, m, j);
                float s0[3], s1[3];
                sym6_mul(cv, j.m0, s0);
                sym6_mul(cv, j.m1, s1);
                const float aa = j.m0[0] * s0[0] + j.m0[1] * s0[1] + j.m0[2] * s0[2] + c.dilation;
                const float bb = j.m0[0] * s1[0] + j.m0[1] * s1[1] + j.m0[2] * s1[2];
                const float cc = j.m1[0] * s1[0] + j.m1[1] * s1[1] + j.m1[2] * s1[2] + c.dilation;
                const float denom = aa * cc - bb * bb;
                const float d2inv = 1.0f / (denom * denom + 0.0000001f);
                float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
                if (d2inv != 0.0f) {
                    dL_da = d2inv * (-cc * cc * gcx + 2.0f * bb * cc * gcy + (denom - aa * cc) * gcz);
                    dL_dc = d2inv * (-aa * aa * gcz + 2.0f * aa * bb * gcy + (denom - aa * cc) * gcx);
                    dL_db = d2inv * 2.0f * (bb * cc * gcx - (denom + 2.0f * bb * bb) * gcy + aa * bb * gcz);
                    const float *m0 = j.m0, *m1 = j.m1;
                    gcov[0] += s2 * (m0[0] * m0[0] * dL_da + m0[0] * m1[0] * dL_db + m1[0] * m1[0] * dL_dc);
                    gcov[3] += s2 * (m0[1] * m0[1] * dL_da + m0[1] * m1[1] * dL_db + m1[1] * m1[1] * dL_dc);
                    gcov[5] += s2 * (m0[2] * m0[2] * dL_da + m0[2] * m1[2] * dL_db + m1[2] * m1[2] * dL_dc);
                    gcov[1] += s2 * (2.0f * m0[0] * m0[1] * dL_da + (m0[0] * m1[1] + m0[1] * m1[0]) * dL_db +
                                     2.0f * m1[0] * m1[1] * dL_dc);
                    gcov[2] += s2 * (2.0f * m0[0] * m0[2] * dL_da + (m0[0] * m1[2] + m0[2] * m1[0]) * dL_db +
                                     2.0f * m1[0] * m1[2] * dL_dc);
                    gcov[4] += s2 * (2.0f * m0[2] * m0[1] * dL_da + (m0[1] * m1[2] + m0[2] * m1[1]) * dL_db +
                                     2.0f * m1[1] * m1[2] * dL_dc);
                }
                float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float gm0 = 2.0f * dL_da * s0[k] + dL_db * s1[k];
                    const float gm1 = 2.0f * dL_dc * s1[k] + dL_db * s0[k];
                    dJ00 += gm0 * cam.view[k * 4 + 0];
                    dJ02 += gm0 * cam.view[k * 4 + 2];
                    dJ11 += gm1 * cam.view[k * 4 + 1];
                    dJ12 += gm1 * cam.view[k * 4 + 2];
                }
                const float tz = 1.0f / j.tz, tz2 = tz * tz, tz3 = tz2 * tz;
                const float dtx = j.xin ? -j.fx * tz2 * dJ02 : 0.0f;
                const float dty = j.yin ? -j.fy * tz2 * dJ12 : 0.0f;
                const float dtz = -j.fx * tz2 * dJ00 - j.fy * tz2 * dJ11 + (2.0f * j.fx * j.tx) * tz3 * dJ02 +
                                  (2.0f * j.fy * j.ty) * tz3 * dJ12;
#pragma unroll
                for (int k = 0; k < 3; k++)
                    gm[k] += cam.view[k * 4 + 0] * dtx + cam.view[k * 4 + 1] * dty + cam.view[k * 4 + 2] * dtz;
            }
            // ---- mean2D (NDC units) -> mean through the perspective divide ----
            {
                const float4 mh = xform4x4(cam.proj, m);
                const float mw = 1.0f / (mh.w + 0.0000001f);
                const float mul1 = mh.x * mw * mw, mul2 = mh.y * mw * mw;
                const float *pr = cam.proj;
                gm[0] += (pr[0] * mw - pr[3] * mul1) * m2d[0] + (pr[1] * mw - pr[3] * mul2) * m2d[1];
                gm[1] += (pr[4] * mw - pr[7] * mul1) * m2d[0] + (pr[5] * mw - pr[7] * mul2) * m2d[1];
                gm[2] += (pr[8] * mw - pr[11] * mul1) * m2d[0] + (pr[9] * mw - pr[11] * mul2) * m2d[1];
            }
            // ---- fused depth channel: z = (view * m).z ----
            gm[0] += cam.view[2] * gz;
            gm[1] += cam.view[6] * gz;
            gm[2] += cam.view[10] * gz;
            // ---- colour -> SH coefficients and mean (view direction) ----
            if (HAS_SH) {
                const float d[3] = {m.x - cam.campos[0], m.y - cam.campos[1], m.z - cam.campos[2]};
                const float len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                const float inv = 1.0f / sqrtf(len2);
                const float x = d[0] * inv, y = d[1] * inv, z = d[2] * inv;
                float ddx = 0.f, ddy = 0.f, ddz = 0.f;
                const float *sh = sh_s + (size_t)tid * c.M * 3;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    const float gch = (mb & (1u << ch)) ? 0.0f : gcol[ch];
                    sh_backward_channel(c.deg, sh + ch, gsh + ch, gch, x, y, z, ddx, ddy, ddz);
                }
                const float dot = d[0] * ddx + d[1] * ddy + d[2] * ddz;
                const float inv3 = inv * inv * inv;
                gm[0] += (ddx * len2 - d[0] * dot) * inv3;
                gm[1] += (ddy * len2 - d[1] * dot) * inv3;
                gm[2] += (ddz * len2 - d[2] * dot) * inv3;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) gmean[k] += s * gm[k];
        }
        if (g.dL_dmeans2D) {
            g.dL_dmeans2D[o * 3 + 0] = m2d[0];
            g.dL_dmeans2D[o * 3 + 1] = m2d[1];
            g.dL_dmeans2D[o * 3 + 2] = 0.0f;
        }
        if (g.dL_dcolors) {
            g.dL_dcolors[o * 3 + 0] = gcol[0];
            g.dL_dcolors[o * 3 + 1] = gcol[1];
            g.dL_dcolors[o * 3 + 2] = gcol[2];
        }
    }

    if (active) {
        if (g.dL_dmeans3D) {
#pragma unroll
            for (int k = 0; k < 3; k++) g.dL_dmeans3D[sg * 3 + k] = gmean[k];
        }
        if (g.dL_dopacities) g.dL_dopacities[sg] = gopac;
        if (in.cov3D) {
            if (g.dL_dcov3D) {
#pragma unroll
                for (int k = 0; k < 6; k++) g.dL_dcov3D[sg * 6 + k] = gcov[k];
            }
        } else {
            // Sigma = R diag(v) R^T, v_k = (mod*s_k)^2 : dL/dv_k = (R^T G R)_kk ; dL/dR = 2 G R diag(v)
            float R[3][3];
            quat_to_R(q, R);
            const float G[3][3] = {{gcov[0], 0.5f * gcov[1], 0.5f * gcov[2]},
                                   {0.5f * gcov[1], gcov[3], 0.5f * gcov[4]},
                                   {0.5f * gcov[2], 0.5f * gcov[4], gcov[5]}};
            const float mod = c.scale_modifier;
            float dR[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float sv = mod * sc[k];
                float GRk[3];
#pragma unroll
                for (int a = 0; a < 3; a++) GRk[a] = G[a][0] * R[0][k] + G[a][1] * R[1][k] + G[a][2] * R[2][k];
                const float dv = R[0][k] * GRk[0] + R[1][k] * GRk[1] + R[2][k] * GRk[2];
                if (g.dL_dscales) g.dL_dscales[sg * 3 + k] = dv * 2.0f * sv * mod;
#pragma unroll
                for (int a = 0; a < 3; a++) dR[a][k] = 2.0f * GRk[a] * sv * sv;
            }
            if (g.dL_drotations) {
                const float r = q[0], x = q[1], y = q[2], z = q[3];
                float *gq = g.dL_drotations + sg * 4;
                gq[0] = 2.0f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
                gq[1] = 2.0f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.0f * x * dR[1][1] - r * dR[1][2] +
                                z * dR[2][0] + r * dR[2][1] - 2.0f * x * dR[2][2]);
                gq[2] = 2.0f * (-2.0f * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] -
                                r * dR[2][0] + z * dR[2][1] - 2.0f * y * dR[2][2]);
                gq[3] = 2.0f * (-2.0f * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - 2.0f * z * dR[1][1] +
                                y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
            }
        }
    }

    if (HAS_SH && g.dL_dshs) {
        // the staged input block is dead now: overwrite it with the gradient block (zeros beyond the evaluated
        // bands) and send it out with one bulk store
        float *dst = g.dL_dshs + ((size_t)scene * c.P + g0) * c.M * 3;
        __syncthreads();  // every thread has finished reading sh_s
        if (active) {
            float *row = sh_s + (size_t)tid * c.M * 3;
            const int used = min(c.M, PB_SH_EVAL) * 3;
#pragma unroll
            for (int k = 0; k < PB_SH_EVAL * 3; k++)
                if (k < used) row[k] = gsh[k];
            for (int k = used; k < c.M * 3; k++) row[k] = 0.f;
        }
        if (bulk) {
            fence_proxy_async_smem();  // make the generic-proxy smem writes visible to the TMA engine
            __syncthreads();
            if (tid == 0) {
                tma_store_1d(dst, sh_s, sh_floats * 4u);
                tma_store_commit_wait();
            }
        } else {
            __syncthreads();
            for (uint32_t k = tid; k < sh_floats; k += PB_THREADS) dst[k] = sh_s[k];
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// Round-2 experiment (GS_TUNE_PBWD_2PHASE; measured SLOWER, kept for the A/B): two phases per thread instead of one
// ---------------------------------------------------------------------------------------------------------
// v1 carried the 48 SH-gradient accumulators through the whole geometry chain: 150 registers unbounded, 128 bounded,
// 18-23 % occupancy, issue slots 56 % busy (latency-bound).  The SH gradient is linear in (basis(dir_v) x dL/dcolour_v)
// and touches the rest of the backward only through dL/ddir, so the loop over the views is run four times with a small
// live state each time:
//   phase G   conic -> cov2D -> cov3D / mean, mean2D -> mean, depth -> mean            (no SH state at all)
//   phase S_c for each colour channel c: 16 accumulators; the channel's 16 staged INPUT coefficients are dead once its
//             loop over the views is done and are overwritten in place by its gradients
// Only the 16 coefficients per channel the evaluator can touch are staged (192 instead of 300 bytes per thread).  The
// gradient rows leave by ONE bulk TMA store when M <= 16, else by coalesced stores that append the zero bands.
// Result on C2: 80 registers, 35 % occupancy (single pass: 128 registers, 23 %), issue slots 69 % busy (56 %) -- but 213 M
// warp-instructions instead of 137 M (the per-view camera / direction / basis work is redone per phase): 0.271 ms
// against 0.220 ms.  The single-pass kernel stays the default.
constexpr int PB2_MIN_CTAS = 6;

template <bool HAS_SH>
__global__ void __launch_bounds__(PB_THREADS, PB2_MIN_CTAS)
k_preprocess_bwd_2phase(const DevCfg c, const DevInputs in, const uint8_t *__restrict__ meta, const float *__restrict__ acc,
                 const GsInGrads g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PbSmem *sm = reinterpret_cast<PbSmem *>(smem_raw);
    float *sh_s = reinterpret_cast<float *>(smem_raw + PB_SMEM_HDR);

    const int scene = blockIdx.y;
    const int g0 = blockIdx.x * PB_THREADS;
    const int n = min(PB_THREADS, c.P - g0);
    const int tid = threadIdx.x;
    const int i = g0 + tid;
    const bool active = tid < n;
    const size_t sg = (size_t)scene * c.P + i;
    const int MS = c.M < PB_SH_EVAL ? c.M : PB_SH_EVAL;   // coefficients staged per Gaussian
    const int RS = c.M == MS ? MS * 3 : MS * 3 + 1;       // floats per staged row: compacted rows get an odd stride (no
                                                          // 16-way bank conflicts between the threads of a warp)
    const uint32_t stage_floats = (uint32_t)n * MS * 3u;

    bool bulk = false;
    if (HAS_SH) {
        const float *src = in.shs + ((size_t)scene * c.P + g0) * c.M * 3;
        const float *dst = g.dL_dshs ? g.dL_dshs + ((size_t)scene * c.P + g0) * c.M * 3 : nullptr;
        bulk = c.M == MS && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && (((stage_floats * 4u) & 15u) == 0) &&
               ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
        if (bulk) {
            if (tid == 0) {
                mbar_init(&sm->bar, 1);
                mbar_fence_init();
                mbar_expect_tx(&sm->bar, stage_floats * 4u);
                tma_load_1d(sh_s, src, stage_floats * 4u, &sm->bar);
            }
        } else {
            const uint32_t row_f = (uint32_t)c.M * 3u, take = (uint32_t)MS * 3u;
            for (uint32_t e = tid; e < stage_floats; e += PB_THREADS) {   // asynchronous 4-byte gathers (LDGSTS)
                const uint32_t row = e / take, col = e - row * take;
                cp_async4(sh_s + row * RS + col, src + (size_t)row * row_f + col);
            }
            cp_async_commit();
        }
    }

    float3 mean = make_float3(0, 0, 0);
    float c6[6] = {0, 0, 0, 0, 0, 0};
    float sc[3] = {0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (active) {
        mean = make_float3(in.means3D[sg * 3 + 0], in.means3D[sg * 3 + 1], in.means3D[sg * 3 + 2]);
        if (in.cov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = in.cov3D[sg * 6 + k];
        } else {
#pragma unroll
            for (int k = 0; k < 3; k++) sc[k] = in.scales[sg * 3 + k];
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = in.rotations[sg * 4 + k];
            cov3d_from_scale_rot(sc, c.scale_modifier, q, c6);
        }
    }
    if (HAS_SH) {
        cp_async_wait<0>();
        __syncthreads();
        if (bulk) mbar_wait(&sm->bar, 0);
    }

    float gmean[3] = {0, 0, 0};

    // ================= phase G: geometry =================
    {
        float gcov[6] = {0, 0, 0, 0, 0, 0}, gopac = 0.f;
        for (int vi = 0; vi < c.VPS; vi++) {
            const int v = scene * c.VPS + vi;
            if (vi % GS_CAM_CHUNK == 0) {
                __syncthreads();
                load_view_cams(c, v, min(GS_CAM_CHUNK, c.VPS - vi), sm->cams);
                __syncthreads();
            }
            if (!active) continue;
            const ViewCam &cam = sm->cams[vi % GS_CAM_CHUNK];
            const size_t o = (size_t)v * c.P + i;
            const uint32_t mb = meta[o];
            float m2d[2] = {0.f, 0.f}, gcol[3] = {0.f, 0.f, 0.f};
            if (mb & GS_META_VISIBLE) {
                const float *a = acc + o * GS_ACC_STRIDE;
                gcol[0] = a[0]; gcol[1] = a[1]; gcol[2] = a[2];
                m2d[0] = a[3]; m2d[1] = a[4];
                const float gcx = a[5], gcy = a[6], gcz = a[7];
                gopac += a[8];
                const float gz = a[9];
                const float s = cam.scale, s2 = s * s;
                const float3 m = make_float3(mean.x * s, mean.y * s, mean.z * s);
                const float cv[6] = {c6[0] * s2, c6[1] * s2, c6[2] * s2, c6[3] * s2, c6[4] * s2, c6[5] * s2};
                float gm[3] = {0.f, 0.f, 0.f};  // dL/dm (view-scaled mean)
                // ---- conic -> cov2D -> cov3D, mean (through J) ----
                {
                    ProjJac j;
                    build_jac(cam, c, m, j);
                    float s0[3], s1[3];
                    sym6_mul(cv, j.m0, s0);
                    sym6_mul(cv, j.m1, s1);
                    const float aa = j.m0[0] * s0[0] + j.m0[1] * s0[1] + j.m0[2] * s0[2] + c.dilation;
                    const float bb = j.m0[0] * s1[0] + j.m0[1] * s1[1] + j.m0[2] * s1[2];
                    const float cc = j.m1[0] * s1[0] + j.m1[1] * s1[1] + j.m1[2] * s1[2] + c.dilation;
                    const float denom = aa * cc - bb * bb;
                    const float d2inv = 1.0f / (denom * denom + 0.0000001f);
                    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
                    if (d2inv != 0.0f) {
                        dL_da = d2inv * (-cc * cc * gcx + 2.0f * bb * cc * gcy + (denom - aa * cc) * gcz);
                        dL_dc = d2inv * (-aa * aa * gcz + 2.0f * aa * bb * gcy + (denom - aa * cc) * gcx);
                        dL_db = d2inv * 2.0f * (bb * cc * gcx - (denom + 2.0f * bb * bb) * gcy + aa * bb * gcz);
                        const float *m0 = j.m0, *m1 = j.m1;
                        gcov[0] += s2 * (m0[0] * m0[0] * dL_da + m0[0] * m1[0] * dL_db + m1[0] * m1[0] * dL_dc);
                        gcov[3] += s2 * (m0[1] * m0[1] * dL_da + m0[1] * m1[1] * dL_db + m1[1] * m1[1] * dL_dc);
                        gcov[5] += s2 * (m0[2] * m0[2] * dL_da + m0[2] * m1[2] * dL_db + m1[2] * m1[2] * dL_dc);
                        gcov[1] += s2 * (2.0f * m0[0] * m0[1] * dL_da + (m0[0] * m1[1] + m0[1] * m1[0]) * dL_db +
                                         2.0f * m1[0] * m1[1] * dL_dc);
                        gcov[2] += s2 * (2.0f * m0[0] * m0[2] * dL_da + (m0[0] * m1[2] + m0[2] * m1[0]) * dL_db +
                                         2.0f * m1[0] * m1[2] * dL_dc);
                        gcov[4] += s2 * (2.0f * m0[2] * m0[1] * dL_da + (m0[1] * m1[2] + m0[2] * m1[1]) * dL_db +
                                         2.0f * m1[1] * m1[2] * dL_dc);
                    }
                    float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const float gm0 = 2.0f * dL_da * s0[k] + dL_db * s1[k];
                        const float gm1 = 2.0f * dL_dc * s1[k] + dL_db * s0[k];
                        dJ00 += gm0 * cam.view[k * 4 + 0];
                        dJ02 += gm0 * cam.view[k * 4 + 2];
                        dJ11 += gm1 * cam.view[k * 4 + 1];
                        dJ12 += gm1 * cam.view[k * 4 + 2];
                    }
                    const float tz = 1.0f / j.tz, tz2 = tz * tz, tz3 = tz2 * tz;
                    const float dtx = j.xin ? -j.fx * tz2 * dJ02 : 0.0f;
                    const float dty = j.yin ? -j.fy * tz2 * dJ12 : 0.0f;
                    const float dtz = -j.fx * tz2 * dJ00 - j.fy * tz2 * dJ11 + (2.0f * j.fx * j.tx) * tz3 * dJ02 +
                                      (2.0f * j.fy * j.ty) * tz3 * dJ12;
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        gm[k] += cam.view[k * 4 + 0] * dtx + cam.view[k * 4 + 1] * dty + cam.view[k * 4 + 2] * dtz;
                }
                // ---- mean2D (NDC units) -> mean through the perspective divide ----
                {
                    const float4 mh = xform4x4(cam.proj, m);
                    const float mw = 1.0f / (mh.w + 0.0000001f);
                    const float mul1 = mh.x * mw * mw, mul2 = mh.y * mw * mw;
                    const float *pr = cam.proj;
                    gm[0] += (pr[0] * mw - pr[3] * mul1) * m2d[0] + (pr[1] * mw - pr[3] * mul2) * m2d[1];
                    gm[1] += (pr[4] * mw - pr[7] * mul1) * m2d[0] + (pr[5] * mw - pr[7] * mul2) * m2d[1];
                    gm[2] += (pr[8] * mw - pr[11] * mul1) * m2d[0] + (pr[9] * mw - pr[11] * mul2) * m2d[1];
                }
                // ---- fused depth channel: z = (view * m).z ----
                gm[0] += cam.view[2] * gz;
                gm[1] += cam.view[6] * gz;
                gm[2] += cam.view[10] * gz;
#pragma unroll
                for (int k = 0; k < 3; k++) gmean[k] += s * gm[k];
            }
            if (g.dL_dmeans2D) {
                g.dL_dmeans2D[o * 3 + 0] = m2d[0];
                g.dL_dmeans2D[o * 3 + 1] = m2d[1];
                g.dL_dmeans2D[o * 3 + 2] = 0.0f;
            }
            if (g.dL_dcolors) {
                g.dL_dcolors[o * 3 + 0] = gcol[0];
                g.dL_dcolors[o * 3 + 1] = gcol[1];
                g.dL_dcolors[o * 3 + 2] = gcol[2];
            }
        }
        if (active) {
            if (g.dL_dopacities) g.dL_dopacities[sg] = gopac;
            if (in.cov3D) {
                if (g.dL_dcov3D) {
#pragma unroll
                    for (int k = 0; k < 6; k++) g.dL_dcov3D[sg * 6 + k] = gcov[k];
                }
            } else {
                // Sigma = R diag(v) R^T, v_k = (mod*s_k)^2 : dL/dv_k = (R^T G R)_kk ; dL/dR = 2 G R diag(v)
                float R[3][3];
                quat_to_R(q, R);
                const float G[3][3] = {{gcov[0], 0.5f * gcov[1], 0.5f * gcov[2]},
                                       {0.5f * gcov[1], gcov[3], 0.5f * gcov[4]},
                                       {0.5f * gcov[2], 0.5f * gcov[4], gcov[5]}};
                const float mod = c.scale_modifier;
                float dR[3][3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float sv = mod * sc[k];
                    float GRk[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) GRk[a] = G[a][0] * R[0][k] + G[a][1] * R[1][k] + G[a][2] * R[2][k];
                    const float dv = R[0][k] * GRk[0] + R[1][k] * GRk[1] + R[2][k] * GRk[2];
                    if (g.dL_dscales) g.dL_dscales[sg * 3 + k] = dv * 2.0f * sv * mod;
#pragma unroll
                    for (int a = 0; a < 3; a++) dR[a][k] = 2.0f * GRk[a] * sv * sv;
                }
                if (g.dL_drotations) {
                    const float r = q[0], x = q[1], y = q[2], z = q[3];
                    float *gq = g.dL_drotations + sg * 4;
                    gq[0] = 2.0f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
                    gq[1] = 2.0f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.0f * x * dR[1][1] - r * dR[1][2] +
                                    z * dR[2][0] + r * dR[2][1] - 2.0f * x * dR[2][2]);
                    gq[2] = 2.0f * (-2.0f * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] -
                                    r * dR[2][0] + z * dR[2][1] - 2.0f * y * dR[2][2]);
                    gq[3] = 2.0f * (-2.0f * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - 2.0f * z * dR[1][1] +
                                    y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
                }
            }
        }
    }

    // ================= phases S_0..S_2: colour -> SH coefficients and mean (view direction), one channel at a time ====
    if (HAS_SH) {
        float *row = sh_s + (size_t)tid * RS;
#pragma unroll 1
        for (int ch = 0; ch < 3; ch++) {
            float gsh[PB_SH_EVAL];
#pragma unroll
            for (int k = 0; k < PB_SH_EVAL; k++) gsh[k] = 0.f;
            for (int vi = 0; vi < c.VPS; vi++) {
                const int v = scene * c.VPS + vi;
                if (c.VPS > GS_CAM_CHUNK && vi % GS_CAM_CHUNK == 0) {  // (a single chunk is still staged from phase G)
                    __syncthreads();
                    load_view_cams(c, v, min(GS_CAM_CHUNK, c.VPS - vi), sm->cams);
                    __syncthreads();
                }
                if (!active) continue;
                const ViewCam &cam = sm->cams[vi % GS_CAM_CHUNK];
                const size_t o = (size_t)v * c.P + i;
                const uint32_t mb = meta[o];
                if (!(mb & GS_META_VISIBLE) || (mb & (1u << ch))) continue;   // invisible, or this channel was clamped at 0
                const float gch = acc[o * GS_ACC_STRIDE + ch];
                const float s = cam.scale;
                const float d[3] = {mean.x * s - cam.campos[0], mean.y * s - cam.campos[1], mean.z * s - cam.campos[2]};
                const float len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                const float inv = 1.0f / sqrtf(len2);
                float ddx = 0.f, ddy = 0.f, ddz = 0.f;
                sh_backward_channel<1>(c.deg, row + ch, gsh, gch, d[0] * inv, d[1] * inv, d[2] * inv, ddx, ddy, ddz);
                const float dot = d[0] * ddx + d[1] * ddy + d[2] * ddz;
                const float inv3 = inv * inv * inv;
                gmean[0] += s * (ddx * len2 - d[0] * dot) * inv3;
                gmean[1] += s * (ddy * len2 - d[1] * dot) * inv3;
                gmean[2] += s * (ddz * len2 - d[2] * dot) * inv3;
            }
            // this channel's staged inputs are dead: its gradients take their place
            if (active) {
#pragma unroll
                for (int k = 0; k < PB_SH_EVAL; k++)
                    if (k < MS) row[k * 3 + ch] = gsh[k];
            }
        }
    }
    if (active && g.dL_dmeans3D) {
#pragma unroll
        for (int k = 0; k < 3; k++) g.dL_dmeans3D[sg * 3 + k] = gmean[k];
    }

    if (HAS_SH && g.dL_dshs) {
        float *dst = g.dL_dshs + ((size_t)scene * c.P + g0) * c.M * 3;
        if (bulk) {
            fence_proxy_async_smem();  // make the generic-proxy smem writes visible to the TMA engine
            __syncthreads();
            if (tid == 0) {
                tma_store_1d(dst, sh_s, stage_floats * 4u);
                tma_store_commit_wait();
            }
        } else {
            __syncthreads();
            // rows of M*3 floats: the staged MS*3 gradients, then zeros for the bands the evaluator never reads
            const uint32_t row_f = (uint32_t)c.M * 3u, take = (uint32_t)MS * 3u, total = (uint32_t)n * row_f;
            for (uint32_t e = tid; e < total; e += PB_THREADS) {
                const uint32_t r = e / row_f, col = e - r * row_f;
                dst[e] = col < take ? sh_s[r * RS + col] : 0.f;
            }
        }
    }
}

}  // namespace

int launch_preprocess_bwd(const DevCfg &c, const DevInputs &in, const GsSaved &s, const float *grad_acc,
                          const GsInGrads &g, cudaStream_t st, int variant) {
    if (c.P == 0) return GS_OK;
    dim3 grid((c.P + PB_THREADS - 1) / PB_THREADS, c.S);
    if (variant != 1) {  // single pass over the views (default)
        if (in.shs) {
            size_t smem = PB_SMEM_HDR + (size_t)PB_THREADS * c.M * 12;
            GS_CUDA_OK(cudaFuncSetAttribute(k_preprocess_bwd<true, PB_MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_preprocess_bwd<true, PB_MIN_CTAS><<<grid, PB_THREADS, smem, st>>>(c, in, s.meta, grad_acc, g);
        } else {
            k_preprocess_bwd<false, PB_MIN_CTAS><<<grid, PB_THREADS, PB_SMEM_HDR, st>>>(c, in, s.meta, grad_acc, g);
        }
    } else if (in.shs) {
        size_t smem = PB_SMEM_HDR + (size_t)PB_THREADS * (c.M <= PB_SH_EVAL ? c.M * 3 : PB_SH_EVAL * 3 + 1) * 4;
        GS_CUDA_OK(cudaFuncSetAttribute(k_preprocess_bwd_2phase<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_preprocess_bwd_2phase<true><<<grid, PB_THREADS, smem, st>>>(c, in, s.meta, grad_acc, g);
    } else {
        k_preprocess_bwd_2phase<false><<<grid, PB_THREADS, PB_SMEM_HDR, st>>>(c, in, s.meta, grad_acc, g);
    }
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}
