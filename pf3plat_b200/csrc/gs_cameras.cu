// gs_cameras.cu -- the camera glue of the render boundary for all views of a call in ONE tiny kernel.
//
// Restates what render_cuda does before its per-view loop, /root/reference/src/model/decoder/cuda_splatting.py:64-87,
// with the helpers it calls: get_fov (/root/reference/src/geometry/projection.py:233-247) and get_projection_matrix
// (cuda_splatting.py:17-44).  The reference spends ~50 small tensor ops here (two matrix inversions through cuSOLVER,
// four host-built constants copied to the device, per-view `.item()` syncs); at PF3plat's call sizes that host-side
// time is comparable to the rasterizer itself.  One thread per view, arithmetic in fp64, results rounded to fp32.
#include "gs_common.cuh"

namespace {

__device__ void inverse3(const double *m, double *o) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double inv = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
    o[0] = c00 * inv; o[1] = (m[2] * m[7] - m[1] * m[8]) * inv; o[2] = (m[1] * m[5] - m[2] * m[4]) * inv;
    o[3] = c01 * inv; o[4] = (m[0] * m[8] - m[2] * m[6]) * inv; o[5] = (m[2] * m[3] - m[0] * m[5]) * inv;
    o[6] = c02 * inv; o[7] = (m[1] * m[6] - m[0] * m[7]) * inv; o[8] = (m[0] * m[4] - m[1] * m[3]) * inv;
}

// general 4x4 inverse (the reference calls extrinsics.inverse(), not a rigid-body shortcut)
__device__ void inverse4(const double *m, double *o) {
    const double s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[6] - m[4] * m[2], s2 = m[0] * m[7] - m[4] * m[3];
    const double s3 = m[1] * m[6] - m[5] * m[2], s4 = m[1] * m[7] - m[5] * m[3], s5 = m[2] * m[7] - m[6] * m[3];
    const double c5 = m[10] * m[15] - m[14] * m[11], c4 = m[9] * m[15] - m[13] * m[11], c3 = m[9] * m[14] - m[13] * m[10];
    const double c2 = m[8] * m[15] - m[12] * m[11], c1 = m[8] * m[14] - m[12] * m[10], c0 = m[8] * m[13] - m[12] * m[9];
    const double inv = 1.0 / (s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
    o[0] = (m[5] * c5 - m[6] * c4 + m[7] * c3) * inv;
    o[1] = (-m[1] * c5 + m[2] * c4 - m[3] * c3) * inv;
    o[2] = (m[13] * s5 - m[14] * s4 + m[15] * s3) * inv;
    o[3] = (-m[9] * s5 + m[10] * s4 - m[11] * s3) * inv;
    o[4] = (-m[4] * c5 + m[6] * c2 - m[7] * c1) * inv;
    o[5] = (m[0] * c5 - m[2] * c2 + m[3] * c1) * inv;
    o[6] = (-m[12] * s5 + m[14] * s2 - m[15] * s1) * inv;
    o[7] = (m[8] * s5 - m[10] * s2 + m[11] * s1) * inv;
    o[8] = (m[4] * c4 - m[5] * c2 + m[7] * c0) * inv;
    o[9] = (-m[0] * c4 + m[1] * c2 - m[3] * c0) * inv;
    o[10] = (m[12] * s4 - m[13] * s2 + m[15] * s0) * inv;
    o[11] = (-m[8] * s4 + m[9] * s2 - m[11] * s0) * inv;
    o[12] = (-m[4] * c3 + m[5] * c1 - m[6] * c0) * inv;
    o[13] = (m[0] * c3 - m[1] * c1 + m[2] * c0) * inv;
    o[14] = (-m[12] * s3 + m[13] * s1 - m[14] * s0) * inv;
    o[15] = (m[8] * s3 - m[9] * s1 + m[10] * s0) * inv;
}

__device__ double ray_dot(const double *kinv, double ax, double ay, double bx, double by) {
    double a[3], b[3];
    for (int i = 0; i < 3; i++) {
        a[i] = kinv[i * 3] * ax + kinv[i * 3 + 1] * ay + kinv[i * 3 + 2];
        b[i] = kinv[i * 3] * bx + kinv[i * 3 + 1] * by + kinv[i * 3 + 2];
    }
    const double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), nb = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / (na * nb);
}

__global__ void k_view_batch(int B, int scale_invariant, const float *__restrict__ extrinsics,
                             const float *__restrict__ intrinsics, const float *__restrict__ near_,
                             const float *__restrict__ far_, float *__restrict__ viewmatrix, float *__restrict__ projmatrix,
                             float *__restrict__ campos, float *__restrict__ tanfov, float *__restrict__ scale_out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    double e[16], k[9], kinv[9], w2c[16];
    for (int i = 0; i < 16; i++) e[i] = extrinsics[(size_t)v * 16 + i];
    for (int i = 0; i < 9; i++) k[i] = intrinsics[(size_t)v * 9 + i];
    double nr = near_[v], fr = far_[v], s = 1.0;
    if (scale_invariant) {                       // cuda_splatting.py:64-71: the scene is rescaled by 1/near
        s = 1.0 / nr;
        e[3] *= s; e[7] *= s; e[11] *= s;
        nr *= s;
        fr *= s;
    }
    inverse3(k, kinv);
    // get_fov: angle between the rays through the mid-points of opposite image edges
    const double fov_x = acos(ray_dot(kinv, 0.0, 0.5, 1.0, 0.5)), fov_y = acos(ray_dot(kinv, 0.5, 0.0, 0.5, 1.0));
    const double tx = tan(0.5 * fov_x), ty = tan(0.5 * fov_y);
    // get_projection_matrix (row-major p), then both matrices transposed as the rasterizer expects them
    const double top = ty * nr, right = tx * nr;
    double p[16] = {0};
    p[0] = 2.0 * nr / (2.0 * right);
    p[5] = 2.0 * nr / (2.0 * top);
    p[14] = 1.0;
    p[10] = fr / (fr - nr);
    p[11] = -(fr * nr) / (fr - nr);
    inverse4(e, w2c);
    float *vm = viewmatrix + (size_t)v * 16, *pm = projmatrix + (size_t)v * 16;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            vm[i * 4 + j] = (float)w2c[j * 4 + i];  // view^T
            double acc = 0.0;                      // full = view^T @ proj^T  ->  full[i][j] = sum_k w2c[k][i] * p[j][k]
            for (int q = 0; q < 4; q++) acc += w2c[q * 4 + i] * p[j * 4 + q];
            pm[i * 4 + j] = (float)acc;
        }
    campos[(size_t)v * 3 + 0] = (float)e[3];
    campos[(size_t)v * 3 + 1] = (float)e[7];
    campos[(size_t)v * 3 + 2] = (float)e[11];
    tanfov[(size_t)v * 2 + 0] = (float)tx;
    tanfov[(size_t)v * 2 + 1] = (float)ty;
    scale_out[v] = (float)s;
}

}  // namespace

extern "C" GS_API int gs_view_batch(int32_t B, int32_t scale_invariant, const float *extrinsics, const float *intrinsics,
                                    const float *near_, const float *far_, float *viewmatrix, float *projmatrix, float *campos,
                                    float *tanfov, float *view_scale, void *stream) {
    if (B < 0) return gs_set_error(GS_ERR_INVALID, "negative view count");
    if (B == 0) return GS_OK;
    if (!extrinsics || !intrinsics || !near_ || !far_ || !viewmatrix || !projmatrix || !campos || !tanfov || !view_scale)
        return gs_set_error(GS_ERR_INVALID, "gs_view_batch: null argument");
    k_view_batch<<<(B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(B, scale_invariant, extrinsics, intrinsics, near_,
                                                                               far_, viewmatrix, projmatrix, campos, tanfov,
                                                                               view_scale);
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}
