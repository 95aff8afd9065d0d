// gs_preprocess.cu -- stage 1 of the forward: per (view, Gaussian) projection, culling, 2-D conic, SH colour.
//
// Semantics: SURVEY.md Appendix A "Preprocess" (the external extension's preprocessCUDA; call site
// /root/reference/src/model/decoder/cuda_splatting.py:117-124).  B200 design:
//  * one thread per Gaussian, LOOPING over the views of its scene, so the scene-level inputs (mean, covariance
//    and the 300-byte SH block) are read from HBM once per call instead of once per view;
//  * the coefficients the evaluator reads (bands 0..3: 16 of PF3plat's 25) are staged into shared memory -- one 1-D
//    bulk TMA copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier when M <= 16, a 4-byte cp.async gather into
//    odd-stride rows otherwise -- fully coalesced, no register staging;
//  * k_sh_colour (below): the colour words alone, pulled out of PINNED HOST memory, for gs_render_host's split pipeline,
//    with k_preprocess running geometry-only (COLOUR = 0) next to it;
//  * outputs are three float4 SoA planes written with 16-byte stores;
//  * tile binning keeps, of upstream's 3-sigma square, only the tiles the alpha >= 1/255 ellipse really reaches
//    (exact box test), so tile lists only hold Gaussians that can contribute (identical pixels, -27 % instances).
#include "gs_common.cuh"

namespace {

constexpr int PRE_THREADS = 128;
constexpr int EMIT_BATCH = 4;  // cursor atomics in flight per thread

struct PreSmem {
    ViewCam cams[GS_CAM_CHUNK];
    float strata[GS_CAM_CHUNK][BIN_SUB];  // depth-stratum boundaries of the staged views (PreEmit::strata)
    uint64_t bar;
};
constexpr int PRE_SMEM_HDR = (sizeof(PreSmem) + 127) / 128 * 128;

// SH colour of one Gaussian seen from direction `dir` (unit vector), bands 0..deg (Appendix A "SH -> RGB": upstream's
// left-to-right sum, +0.5, clamp at 0 with the clamp recorded per channel).  Every operation is an explicit
// round-to-nearest intrinsic -- no contraction or re-association left to the compiler -- because TWO kernels inline this
// function (k_preprocess and k_sh_colour, the split pipeline of gs_render_host) and must produce the same bits.
__device__ __forceinline__ float3 eval_sh(int deg, const float *sh /* [M][3] for this Gaussian */, float3 dir,
                                          uint32_t &clamped) {
    const float x = dir.x, y = dir.y, z = dir.z;
    float v[3];
    // band by band (basis values of one band live at a time), every channel's sum in upstream's left-to-right order
#pragma unroll
    for (int c = 0; c < 3; c++) v[c] = __fmul_rn(GS_SH_C0, sh[c]);
    if (deg > 0) {   // (uniform branches: the degree is a launch constant)
        const float b1 = __fmul_rn(-GS_SH_C1, y), b2 = __fmul_rn(GS_SH_C1, z), b3 = __fmul_rn(-GS_SH_C1, x);
#pragma unroll
        for (int c = 0; c < 3; c++)
            v[c] = __fmaf_rn(b3, sh[3 * 3 + c], __fmaf_rn(b2, sh[2 * 3 + c], __fmaf_rn(b1, sh[1 * 3 + c], v[c])));
        if (deg > 1) {
            const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
            const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
            {
                const float b4 = __fmul_rn(GS_SH_C2_0, xy), b5 = __fmul_rn(GS_SH_C2_1, yz);
                const float b6 = __fmul_rn(GS_SH_C2_2, __fsub_rn(__fsub_rn(__fmul_rn(2.0f, zz), xx), yy));
                const float b7 = __fmul_rn(GS_SH_C2_3, xz), b8 = __fmul_rn(GS_SH_C2_4, __fsub_rn(xx, yy));
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float t = __fmaf_rn(b5, sh[5 * 3 + c], __fmaf_rn(b4, sh[4 * 3 + c], v[c]));
                    t = __fmaf_rn(b7, sh[7 * 3 + c], __fmaf_rn(b6, sh[6 * 3 + c], t));
                    v[c] = __fmaf_rn(b8, sh[8 * 3 + c], t);
                }
            }
            if (deg > 2) {
                const float zz4 = __fsub_rn(__fsub_rn(__fmul_rn(4.0f, zz), xx), yy);
                const float b9 = __fmul_rn(__fmul_rn(GS_SH_C3_0, y), __fsub_rn(__fmul_rn(3.0f, xx), yy));
                const float b10 = __fmul_rn(__fmul_rn(GS_SH_C3_1, xy), z);
                const float b11 = __fmul_rn(__fmul_rn(GS_SH_C3_2, y), zz4);
                const float b12 = __fmul_rn(__fmul_rn(GS_SH_C3_3, z), __fsub_rn(__fsub_rn(__fmul_rn(2.0f, zz), __fmul_rn(3.0f, xx)), __fmul_rn(3.0f, yy)));
                const float b13 = __fmul_rn(__fmul_rn(GS_SH_C3_4, x), zz4);
                const float b14 = __fmul_rn(__fmul_rn(GS_SH_C3_5, z), __fsub_rn(xx, yy));
                const float b15 = __fmul_rn(__fmul_rn(GS_SH_C3_6, x), __fsub_rn(xx, __fmul_rn(3.0f, yy)));
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float t = __fmaf_rn(b10, sh[10 * 3 + c], __fmaf_rn(b9, sh[9 * 3 + c], v[c]));
                    t = __fmaf_rn(b12, sh[12 * 3 + c], __fmaf_rn(b11, sh[11 * 3 + c], t));
                    t = __fmaf_rn(b14, sh[14 * 3 + c], __fmaf_rn(b13, sh[13 * 3 + c], t));
                    v[c] = __fmaf_rn(b15, sh[15 * 3 + c], t);
                }
            }
        }
    }
    float r[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float w = __fadd_rn(v[c], 0.5f);
        if (w < 0.0f) clamped |= (1u << c);
        r[c] = fmaxf(w, 0.0f);
    }
    return make_float3(r[0], r[1], r[2]);
}

// unit vector from the camera to the (scaled) mean; pinned like eval_sh
__device__ __forceinline__ float3 sh_view_dir(float3 m, const float *campos) {
    const float dx = __fsub_rn(m.x, campos[0]), dy = __fsub_rn(m.y, campos[1]), dz = __fsub_rn(m.z, campos[2]);
    const float inv = __frcp_rn(__fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)))));
    return make_float3(__fmul_rn(dx, inv), __fmul_rn(dy, inv), __fmul_rn(dz, inv));
}

// Everything preprocess derives for one (view, Gaussian).
struct Splat {
    float4 r0, r1, r2;
    int32_t radius;
    uint32_t meta;
    ushort4 rect;
};

// Appendix A "Preprocess" for one Gaussian in one view.  COLOUR = 2: `sh` points at this Gaussian's [M][3] block in
// shared memory; 1: `rgb_in` is its precomputed colour; 0: geometry only (the colour words are written by k_sh_colour).
template <int COLOUR>
__device__ __forceinline__ void project_splat(const DevCfg &c, const ViewCam &cam, float3 mean, const float *c6,
                                              float opac, const float *sh, const float *rgb_in, Splat &out) {
    out.radius = 0;
    out.meta = 0;
    out.rect = make_ushort4(0, 0, 0, 0);
    const float s = cam.scale, s2 = s * s;
    const float3 m = make_float3(mean.x * s, mean.y * s, mean.z * s);
    const float3 pv = xform4x3(cam.view, m);
    if (!(pv.z > c.near_cull_z)) return;  // in_frustum
    const float4 ph = xform4x4(cam.proj, m);
    const float pw = 1.0f / (ph.w + 0.0000001f);
    const float cv[6] = {c6[0] * s2, c6[1] * s2, c6[2] * s2, c6[3] * s2, c6[4] * s2, c6[5] * s2};
    ProjJac j;
    build_jac(cam, c, m, j);
    float s0[3], s1[3];
    sym6_mul(cv, j.m0, s0);
    sym6_mul(cv, j.m1, s1);
    const float a = j.m0[0] * s0[0] + j.m0[1] * s0[1] + j.m0[2] * s0[2] + c.dilation;
    const float b = j.m0[0] * s1[0] + j.m0[1] * s1[1] + j.m0[2] * s1[2];
    const float cc = j.m1[0] * s1[0] + j.m1[1] * s1[1] + j.m1[2] * s1[2] + c.dilation;
    const float det = a * cc - b * b;
    if (det == 0.0f) return;
    const float det_inv = 1.0f / det;
    const float A = cc * det_inv, B = -b * det_inv, C = a * det_inv;
    const float mid = 0.5f * (a + cc);
    const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float rad = ceilf(3.0f * sqrtf(fmaxf(mid + sq, mid - sq)));
    const float px = ndc2pix(ph.x * pw, c.W), py = ndc2pix(ph.y * pw, c.H);
    // upstream's 3-sigma square in tiles (getRect)
    const int rminx = min(c.gx, max(0, (int)((px - rad) / (float)GS_TILE)));
    const int rminy = min(c.gy, max(0, (int)((py - rad) / (float)GS_TILE)));
    const int rmaxx = min(c.gx, max(0, (int)((px + rad + (float)(GS_TILE - 1)) / (float)GS_TILE)));
    const int rmaxy = min(c.gy, max(0, (int)((py + rad + (float)(GS_TILE - 1)) / (float)GS_TILE)));
    if ((rmaxx - rminx) * (rmaxy - rminy) == 0) return;
    out.radius = (int32_t)rad;
    out.meta = GS_META_VISIBLE;
    float3 rgb;
    if (COLOUR == 2) rgb = eval_sh(c.deg, sh, sh_view_dir(m, cam.campos), out.meta);
    else if (COLOUR == 1) rgb = make_float3(rgb_in[0], rgb_in[1], rgb_in[2]);
    else rgb = make_float3(0.f, 0.f, 0.f);   // COLOUR == 0: k_sh_colour fills the three colour words of the record
    // Tight binning.  A pixel can only receive this Gaussian if alpha = o*G >= 1/255, i.e. log2 G >= -log2(255 o).
    // The candidate tiles are upstream's 3-sigma square; k_preprocess / k_emit_buckets keep a candidate only if the
    // alpha >= 1/255 ellipse really reaches it (gs_box_reaches, exact).  Tiles dropped this way cannot change any
    // pixel.  The threshold carries a margin for the fp32 rounding of log2 G in the compositor (it grows with the
    // reach of the footprint).
    const float reach = rad + (float)GS_TILE;
    const float margin = 1e-3f + 4e-6f * reach * reach;              // in units of tau = 2 ln(255 o)
    const float tau = 2.0f * logf(255.0f * opac) + margin;
    const float hA = (-0.5f * GS_LOG2E) * A, nB = -GS_LOG2E * B, hC = (-0.5f * GS_LOG2E) * C;
    const float reach2 = -(0.5f * GS_LOG2E) * tau;                     // same threshold in log2-G units
    if (tau > 0.0f)  // else opacity < 1/255: visible (radius reported) but it can never contribute
        out.rect = make_ushort4((unsigned short)rminx, (unsigned short)rminy, (unsigned short)rmaxx, (unsigned short)rmaxy);
    out.r0 = make_float4(px, py, hA, nB);
    out.r1 = make_float4(hC, opac, rgb.x, rgb.y);
    out.r2 = make_float4(rgb.z, pv.z, reach2, 0.0f);
}

// MINB: resident CTAs per SM the register allocation is bounded for.  Unbounded the kernel takes 96 registers (5 CTAs/SM,
// 28 % occupancy; ncu: issue slots 50 % busy, latency-bound).  Measured on C2 with the 16-coefficient staging: 96
// registers 0.222 ms; bounded to 64 (8 CTAs/SM, 44 % occupancy, 116 bytes of spills = +15 % instructions) 0.275 ms.
template <int COLOUR, int MINB>
__global__ void __launch_bounds__(PRE_THREADS, MINB)
k_preprocess(const DevCfg c, const DevInputs in, float4 *__restrict__ rec0, float4 *__restrict__ rec1,
             float4 *__restrict__ rec2, uint8_t *__restrict__ meta, int32_t *__restrict__ radii,
             ushort4 *__restrict__ rects, const PreEmit emit, const int g_begin, const int g_end, const int sh_raw16) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PreSmem *sm = reinterpret_cast<PreSmem *>(smem_raw);
    float *sh_s = reinterpret_cast<float *>(smem_raw + PRE_SMEM_HDR);

    const int scene = blockIdx.y;
    const int g0 = g_begin + blockIdx.x * PRE_THREADS;  // this launch covers Gaussians [g_begin, g_end) of every scene
    const int n = min(PRE_THREADS, g_end - g0);
    const int tid = threadIdx.x;
    const int i = g0 + tid;
    const bool active = tid < n;
    const size_t sg = (size_t)scene * c.P + i;  // scene-level index

    // ---- stage this CTA's SH block ----
    // Only the bands the evaluator can touch are wanted: MS = min(M, 16) coefficients per Gaussian (PF3plat hands over
    // M = 25 of which 16..24 are never read).
    //  * M <= 16: the block is contiguous and arrives by ONE 1-D bulk TMA copy (cp.async.bulk, SASS UBLKCP) completing on
    //    an mbarrier.
    //  * M > 16 (default): the wanted 192 bytes of every row are gathered with 4-byte cp.async into rows compacted to an
    //    ODD stride (49 floats): with 48, thread t's row starts at bank 16 t mod 32 and all 32 lanes of a warp hit two
    //    banks (measured: 90 M bank conflicts, preprocess 0.24 -> 0.47 ms).  Rows are only 4-byte aligned (300 = 18 * 16 +
    //    12), so neither a bulk copy per row nor a 2-D tensor map is possible.
    //  * M > 16, sh_raw16 (gs_render_host's ZERO-COPY feed; GS_TUNE_PRE_SH_RAW16 for the A/B): rows keep their own stride of
    //    3 M floats in shared memory and only the 16-byte pieces that hold a wanted coefficient are copied (16-byte
    //    cp.async, SASS LDGSTS.128).  Built for a source in PINNED HOST memory, which is pulled over PCIe at the rate of
    //    16-byte loads (scripts/probes/pcie_pull_probe.cu: 4-byte loads 2.96 ms, 16-byte loads 2.53 ms, the copy engine
    //    2.70 ms for the whole rows).  On device-resident SH it loses to the compacting gather (C2: 0.246 vs 0.220 ms --
    //    38 instead of 25 KB of shared memory per CTA leave the L1 28 KB at five CTAs per SM).
    const int MS = c.M < 16 ? c.M : 16;
    const int RS = (c.M == MS || sh_raw16) ? c.M * 3 : 49;   // floats per staged row (3 M = 75 for PF3plat: 11 t mod 32, conflict-free)
    constexpr bool HAS_SH = COLOUR == 2;
    bool bulk = false;
    if (HAS_SH) {
        const float *src = in.shs + ((size_t)scene * c.P + g0) * c.M * 3;
        if (c.M == MS) {
            const uint32_t bytes = (uint32_t)n * c.M * 12u;
            bulk = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((bytes & 15u) == 0);
            if (bulk) {
                if (tid == 0) {
                    mbar_init(&sm->bar, 1);
                    mbar_fence_init();
                    mbar_expect_tx(&sm->bar, bytes);
                    tma_load_1d(sh_s, src, bytes, &sm->bar);
                }
            } else {  // unaligned slice or ragged tail: plain coalesced copy
                for (uint32_t k = tid; k < (uint32_t)n * c.M * 3u; k += PRE_THREADS) sh_s[k] = src[k];
            }
        } else if (sh_raw16) {
            const uint32_t row_f = (uint32_t)c.M * 3u, want_f = (uint32_t)MS * 3u, nfl = (uint32_t)n * row_f, nq = nfl >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            for (uint32_t q = tid; q < nq; q += PRE_THREADS) {
                const uint32_t f0 = q << 2, r = f0 / row_f, c0 = f0 - r * row_f;
                // floats [c0, c0 + 4) of row r (running over into the next row's first coefficients when c0 + 3 >= row_f)
                if (c0 < want_f || c0 + 3u >= row_f) cp_async16(sh_s + f0, src4 + q);
            }
            for (uint32_t f = (nq << 2) + tid; f < nfl; f += PRE_THREADS)   // ragged end of the block (n * row_f % 4 floats)
                if (f % row_f < want_f) cp_async4(sh_s + f, src + f);
            cp_async_commit();
        } else {
            // asynchronous 4-byte copies (cp.async, SASS LDGSTS): all 48 per thread are in flight together while the
            // scalar inputs load -- a plain load/store loop here cost 0.25 ms on C2 (serialised latencies)
            const uint32_t row_f = (uint32_t)c.M * 3u;
            for (uint32_t e = tid; e < (uint32_t)n * 48u; e += PRE_THREADS) {
                const uint32_t row = e / 48u, col = e - row * 48u;
                cp_async4(sh_s + row * 49u + col, src + (size_t)row * row_f + col);
            }
            cp_async_commit();
        }
    }
    // ---- scene-level inputs (issued while the bulk copy is in flight) ----
    float3 mean = make_float3(0, 0, 0);
    float c6[6] = {0, 0, 0, 0, 0, 0};
    float opac = 0.f;
    if (active) {
        mean = make_float3(in.means3D[sg * 3 + 0], in.means3D[sg * 3 + 1], in.means3D[sg * 3 + 2]);
        opac = in.opacities[sg];
        if (in.cov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = in.cov3D[sg * 6 + k];
        } else {
            const float s[3] = {in.scales[sg * 3 + 0], in.scales[sg * 3 + 1], in.scales[sg * 3 + 2]};
            const float q[4] = {in.rotations[sg * 4 + 0], in.rotations[sg * 4 + 1], in.rotations[sg * 4 + 2],
                                in.rotations[sg * 4 + 3]};
            cov3d_from_scale_rot(s, c.scale_modifier, q, c6);
        }
    }
    if (HAS_SH) {
        cp_async_wait<0>();  // this thread's share of the strided copy (no-op on the other paths)
        __syncthreads();     // barrier init (bulk) or the copies (fallbacks) visible to everyone
        if (bulk) mbar_wait(&sm->bar, 0);
    }

    // appends whose cursor value is still on its way back from L2 (see the emission below)
    bool pend_ok[EMIT_BATCH];
    uint32_t pend_slot[EMIT_BATCH], pend_pos[EMIT_BATCH];
    uint64_t pend_key = 0;
#pragma unroll
    for (int k = 0; k < EMIT_BATCH; k++) {
        pend_ok[k] = false;
        pend_slot[k] = pend_pos[k] = 0;
    }
    auto flush_pending = [&]() {
#pragma unroll
        for (int k = 0; k < EMIT_BATCH; k++) {
            if (pend_ok[k] && pend_pos[k] < emit.sub_cap)
                emit.bucket[(size_t)pend_slot[k] * emit.sub_cap + pend_pos[k]] = pend_key;
            pend_ok[k] = false;
        }
    };

    for (int v0 = 0; v0 < c.VPS; v0 += GS_CAM_CHUNK) {
        const int nv = min(GS_CAM_CHUNK, c.VPS - v0);
        __syncthreads();
        load_view_cams(c, scene * c.VPS + v0, nv, sm->cams);
        if (emit.strata && !emit.strata_per_tile)
            for (int t = tid; t < nv * BIN_SUB; t += PRE_THREADS)
                sm->strata[t / BIN_SUB][t % BIN_SUB] = emit.strata[(size_t)(scene * c.VPS + v0) * BIN_SUB + t];
        __syncthreads();
        if (!active) continue;
        for (int vi = 0; vi < nv; vi++) {
            const int v = scene * c.VPS + v0 + vi;
            const size_t o = (size_t)v * c.P + i;
            Splat sp;
            project_splat<COLOUR>(c, sm->cams[vi], mean, c6, opac, sh_s + (size_t)tid * RS,
                                  COLOUR == 1 ? in.colors_precomp + o * 3 : nullptr, sp);
            if (sp.radius > 0) {
                rec0[o] = sp.r0;
                if (COLOUR != 0) {
                    rec1[o] = sp.r1;
                    rec2[o] = sp.r2;
                } else {   // the colour words (rec1.zw, rec2.x) belong to k_sh_colour, which may be running right now
                    *reinterpret_cast<float2 *>(&rec1[o]) = make_float2(sp.r1.x, sp.r1.y);
                    reinterpret_cast<float *>(&rec2[o])[1] = sp.r2.y;
                    reinterpret_cast<float2 *>(&rec2[o])[1] = make_float2(sp.r2.z, sp.r2.w);
                }
            }
            radii[o] = sp.radius;
            rects[o] = sp.rect;
            // Hand this Gaussian to every (view, tile) list it joins.  Exact path: count it (RED.ADD, no return;
            // k_emit_buckets appends it later, once the offsets are known).  Speculative path: append it now to the
            // tile's fixed-capacity sub-bucket; an entry beyond the capacity is dropped and the overflow is detected
            // from the cursor by k_spec_check (the call is then redone on the exact path).  Sub-bucket
            // i % BIN_SUB, one counter per 32-byte sector -- see gs_binning.cu.  (With very many (view, tile) buckets the
            // fused appends of all views thrash L2; gs_forward then passes counters = NULL and emits per view instead.)
            if (emit.counters) {
                const int w = sp.rect.z - sp.rect.x, nt = w * (sp.rect.w - sp.rect.y);
                const uint32_t tbase = (uint32_t)v * (uint32_t)c.ntiles;
                uint32_t sub = (uint32_t)i & (BIN_SUB - 1);
                if (emit.strata && !emit.strata_per_tile) {
                    sub = 0;
#pragma unroll
                    for (int q = 0; q < BIN_SUB - 1; q++) sub += sp.r2.y >= sm->strata[vi][q] ? 1u : 0u;
                }
                if (emit.bucket) {
                    // The cursor atomics return the slot to write, an L2 round trip each (42 % of this kernel's stall
                    // samples when the store followed its atomic directly).  Candidates go EMIT_BATCH at a time,
                    // all their atomics in flight together, and the dependent stores are deferred until the next batch
                    // is about to be issued -- normally one whole view of arithmetic later.
                    // candidates in row-major order of the rectangle; (cx, cy) walks it without the integer division
                    // that t / w cost per candidate (the emission loop was 36 % of this kernel's instructions)
                    int cx = sp.rect.x, cy = sp.rect.y;
                    for (int t0 = 0; t0 < nt; t0 += EMIT_BATCH) {
                        bool ok[EMIT_BATCH];
                        uint32_t slot[EMIT_BATCH];
#pragma unroll
                        for (int k = 0; k < EMIT_BATCH; k++) {
                            ok[k] = cy < (int)sp.rect.w && gs_tile_reached(sp.r0, sp.r1, sp.r2, cx, cy);
                            const uint32_t tile = tbase + (uint32_t)(cy * c.gx + cx);
                            // per-tile boundaries (clouds whose tiles each see a narrow depth range): the candidate
                            // tile's own eight floats decide the stratum
                            const uint32_t subk = (emit.strata_per_tile && ok[k]) ? gs_tile_stratum(emit.strata, tile, sp.r2.y) : sub;
                            slot[k] = tile * BIN_SUB + subk;
                            if (++cx == (int)sp.rect.z) {
                                cx = sp.rect.x;
                                cy++;
                            }
                        }
                        flush_pending();
                        pend_key = ((uint64_t)__float_as_uint(sp.r2.y) << 32) | (uint32_t)i;
#pragma unroll
                        for (int k = 0; k < EMIT_BATCH; k++) {
                            pend_ok[k] = ok[k];
                            pend_slot[k] = slot[k];
                            if (ok[k]) pend_pos[k] = atomicAdd(&emit.counters[(size_t)slot[k] * BIN_PAD], 1u);
                        }
                    }
                } else {
                    for (int ty = sp.rect.y; ty < (int)sp.rect.w; ty++)
                        for (int tx = sp.rect.x; tx < (int)sp.rect.z; tx++)
                            if (gs_tile_reached(sp.r0, sp.r1, sp.r2, tx, ty))
                                atomicAdd(&emit.counters[(size_t)((tbase + (uint32_t)(ty * c.gx + tx)) * BIN_SUB + sub) * BIN_PAD], 1u);
                }
            }
            meta[o] = (uint8_t)sp.meta;
        }
    }
    flush_pending();
}

// ---------------------------------------------------------------------------------------------------------
// k_sh_colour: the colour words of the splat records, computed apart from the geometry (gs_render_host's split pipeline)
// ---------------------------------------------------------------------------------------------------------
// With host buffers the SH block is 88 % of the bytes that cross PCIe, and geometry + binning + the tile sort need none
// of it.  gs_render_host therefore runs k_preprocess in geometry-only mode (COLOUR = 0) followed by the tile sort on the
// launch stream, while THIS kernel, on a second stream, pulls the SH rows straight out of the caller's pinned buffer
// (16-byte pieces at the rows' own stride, as k_preprocess's sh_raw16 mode) and writes rec1.zw / rec2.x; the compositor
// waits for both.  Link-bound by design: a persistent grid of three CTAs per SM keeps ~11 MB of host reads in flight (see
// launch_sh_colour for what that does to the geometry kernel on the same SMs, and why it is still the fastest).
// Same eval_sh / sh_view_dir as k_preprocess (pinned arithmetic): the split pipeline's images are bit-identical.
constexpr int SHC_THREADS = 128;
constexpr int SHC_VIEWS = 64;   // cameras staged at a time

__global__ void __launch_bounds__(SHC_THREADS, 3)
k_sh_colour(const DevCfg c, const float *__restrict__ means3D, const float *__restrict__ shs, float4 *__restrict__ rec1,
            float4 *__restrict__ rec2, uint8_t *__restrict__ clamp_out /* [V*P] or NULL */, const int blocks_per_scene) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *cam_s = reinterpret_cast<float4 *>(smem_raw);                    // (campos xyz, scale) per staged view
    float *sh_s = reinterpret_cast<float *>(smem_raw + SHC_VIEWS * sizeof(float4));
    const int tid = threadIdx.x;
    const uint32_t row_f = (uint32_t)c.M * 3u, want_f = (uint32_t)(c.M < 16 ? c.M : 16) * 3u;
    const int total = blocks_per_scene * c.S;
    for (int blk = blockIdx.x; blk < total; blk += gridDim.x) {
        const int scene = blk / blocks_per_scene, g0 = (blk - scene * blocks_per_scene) * SHC_THREADS;
        const int n = min(SHC_THREADS, c.P - g0);
        const bool active = tid < n;
        const size_t sg = (size_t)scene * c.P + g0 + tid;
        __syncthreads();   // the previous block's rows and cameras have been read by everybody
        {
            const float *src = shs + ((size_t)scene * c.P + g0) * row_f;
            const uint32_t nfl = (uint32_t)n * row_f, nq = nfl >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            for (uint32_t q = tid; q < nq; q += SHC_THREADS) {
                const uint32_t f0 = q << 2, r = f0 / row_f, c0 = f0 - r * row_f;
                if (c0 < want_f || c0 + 3u >= row_f) cp_async16(sh_s + f0, src4 + q);
            }
            for (uint32_t f = (nq << 2) + tid; f < nfl; f += SHC_THREADS)
                if (f % row_f < want_f) cp_async4(sh_s + f, src + f);
            cp_async_commit();
        }
        float3 mean = make_float3(0.f, 0.f, 0.f);
        if (active) mean = make_float3(means3D[sg * 3 + 0], means3D[sg * 3 + 1], means3D[sg * 3 + 2]);
        for (int v0 = 0; v0 < c.VPS; v0 += SHC_VIEWS) {
            const int nv = min(SHC_VIEWS, c.VPS - v0);
            if (v0) __syncthreads();
            for (int t = tid; t < nv; t += SHC_THREADS) {
                const int v = scene * c.VPS + v0 + t;
                cam_s[t] = make_float4(c.campos[v * 3 + 0], c.campos[v * 3 + 1], c.campos[v * 3 + 2], c.view_scale ? c.view_scale[v] : 1.0f);
            }
            if (v0 == 0) cp_async_wait<0>();
            __syncthreads();
            if (!active) continue;
            const float *sh = sh_s + (size_t)tid * row_f;
            for (int vi = 0; vi < nv; vi++) {
                const float4 cam = cam_s[vi];
                const size_t o = (size_t)(scene * c.VPS + v0 + vi) * c.P + g0 + tid;
                const float3 m = make_float3(mean.x * cam.w, mean.y * cam.w, mean.z * cam.w);
                const float cp[3] = {cam.x, cam.y, cam.z};
                uint32_t clamped = 0;
                const float3 rgb = eval_sh(c.deg, sh, sh_view_dir(m, cp), clamped);
                reinterpret_cast<float2 *>(&rec1[o])[1] = make_float2(rgb.x, rgb.y);
                reinterpret_cast<float *>(&rec2[o])[0] = rgb.z;
                if (clamp_out) clamp_out[o] = (uint8_t)clamped;
            }
        }
    }
}

__global__ void k_mark_visible(const DevCfg c, const float *__restrict__ means3D, uint8_t *__restrict__ present) {
    __shared__ ViewCam cam;
    const int v = blockIdx.y;
    load_view_cams(c, v, 1, &cam);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.P) return;
    const size_t sg = (size_t)(v / c.VPS) * c.P + i;
    const float s = cam.scale;
    const float3 m = make_float3(means3D[sg * 3] * s, means3D[sg * 3 + 1] * s, means3D[sg * 3 + 2] * s);
    present[(size_t)v * c.P + i] = xform4x3(cam.view, m).z > c.near_cull_z;
}

}  // namespace

int launch_preprocess(const DevCfg &c, const DevInputs &in, float4 *rec0, float4 *rec1, float4 *rec2, uint8_t *meta,
                      int32_t *radii, ushort4 *rects, const PreEmit &emit, cudaStream_t st, int g_begin, int g_end,
                      int variant) {
    if (g_end < 0) g_end = c.P;
    if (g_end <= g_begin) return GS_OK;
    const bool more_ctas = (variant & 1) != 0;
    dim3 grid((g_end - g_begin + PRE_THREADS - 1) / PRE_THREADS, c.S);
    // M > 16, on request: 16-byte pieces at the rows' own stride -- if every CTA's block starts 16-byte aligned (CTAs cover
    // 128 rows = 1536 M bytes, so the first one decides -- per scene); else the compacting 4-byte gather
    int sh_raw16 = 0;
    if (in.shs && c.M > 16 && (variant & 2) && !(variant & 4)) {
        const uintptr_t first = reinterpret_cast<uintptr_t>(in.shs) + (size_t)g_begin * c.M * 12;
        sh_raw16 = (first & 15u) == 0 && (c.S == 1 || ((size_t)c.P * c.M * 12) % 16 == 0);
    }
    const size_t smem = PRE_SMEM_HDR + ((in.shs && !(variant & 4)) ? (size_t)PRE_THREADS * ((c.M <= 16 || sh_raw16) ? c.M * 3 : 49) * 4 : 0);
    if (smem > (size_t)227 * 1024) return gs_set_error(GS_ERR_INVALID, "too many SH coefficients per Gaussian for the staging buffer");
    auto launch = [&](auto kern) -> int {
        GS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, PRE_THREADS, smem, st>>>(c, in, rec0, rec1, rec2, meta, radii, rects, emit, g_begin, g_end, sh_raw16);
        return GS_OK;
    };
    int rc;
    if (variant & 4) rc = more_ctas ? launch(k_preprocess<0, 6>) : launch(k_preprocess<0, 5>);   // geometry only
    else if (in.shs) rc = more_ctas ? launch(k_preprocess<2, 6>) : launch(k_preprocess<2, 5>);
    else rc = more_ctas ? launch(k_preprocess<1, 6>) : launch(k_preprocess<1, 5>);
    if (rc != GS_OK) return rc;
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}

bool sh_colour_supported(const DevCfg &c, const DevInputs &in) {
    return in.shs && c.M > 16 && (reinterpret_cast<uintptr_t>(in.shs) & 15u) == 0 &&
           (c.S == 1 || ((size_t)c.P * c.M * 12) % 16 == 0) &&
           SHC_VIEWS * sizeof(float4) + (size_t)SHC_THREADS * c.M * 12 <= (size_t)72 * 1024;
}

int launch_sh_colour(const DevCfg &c, const DevInputs &in, float4 *rec1, float4 *rec2, uint8_t *clamp_out, cudaStream_t st) {
    if (c.P == 0 || c.V == 0) return GS_OK;
    if (!sh_colour_supported(c, in)) return gs_set_error(GS_ERR_INVALID, "k_sh_colour: unsupported SH layout");
    static int sms_of_device[64] = {};
    int dev = 0;
    GS_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && sms_of_device[dev] == 0)
        GS_CUDA_OK(cudaDeviceGetAttribute(&sms_of_device[dev], cudaDevAttrMultiProcessorCount, dev));
    const int sms = (dev >= 0 && dev < 64 && sms_of_device[dev] > 0) ? sms_of_device[dev] : 148;
    const int bps = (c.P + SHC_THREADS - 1) / SHC_THREADS;
    const long long total = (long long)bps * c.S;
    // Three CTAs per SM, each issuing its block's pieces all at once: the link wants its reads deep.  What that does to the
    // kernels sharing the SMs was the surprise -- an SM with host reads in flight slows all its other memory instructions in
    // proportion (C2, inside gs_render_host; colour kernel / geometry / sort / whole call in ms; alone: geometry 0.19, sort
    // 0.107; fused k_preprocess pulling the block instead: 3.64):
    //   3 CTAs, whole block at once (this)   2.63 / 2.63 / 0.108 after it / 3.57-3.58   geometry ends WITH the pull, sort exposed
    //   3 CTAs, 1 piece per thread and group 2.73 / 0.82 / 0.59 / 3.66                   (two groups in flight: cp.async.wait_group 1)
    //   2 CTAs, 2 pieces                     2.84 / 1.02 / 0.67 / 3.61-3.62
    //   1 CTA,  1 piece                      2.87 / 0.33 / 0.20 / 3.84
    //   48 CTAs in all, whole block          2.76 / 0.33 / 0.25 / 3.65                   few SMs pulling: the interference is per SM
    //   per-row bulk copies (TMA), 2 CTAs    2.68 / 2.04 / 0.63 / 3.53-3.66              same starvation: not the LSU queue
    //   few reads until geometry + sort are done (a flag polled per block), then whole blocks: 2.80 / 0.82 / 0.60 / 3.64
    // Hiding geometry + sort costs the pull more than the 0.3 ms they take; the pull at full rate wins.
    const int per_sm = 3;
    const int grid = (int)(total < (long long)sms * per_sm ? total : (long long)sms * per_sm);
    const size_t smem = SHC_VIEWS * sizeof(float4) + (size_t)SHC_THREADS * c.M * 12;
    GS_CUDA_OK(cudaFuncSetAttribute(k_sh_colour, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sh_colour<<<grid, SHC_THREADS, smem, st>>>(c, in.means3D, in.shs, rec1, rec2, clamp_out, bps);
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}

int launch_mark_visible(const DevCfg &c, const float *means3D, uint8_t *present, cudaStream_t st) {
    if (c.P == 0) return GS_OK;
    dim3 grid((c.P + 255) / 256, c.V);
    k_mark_visible<<<grid, 256, 0, st>>>(c, means3D, present);
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}
