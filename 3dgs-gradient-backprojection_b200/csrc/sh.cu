// View-dependent colours from spherical-harmonics coefficients: the `sh_degree=3` branch of
// `rasterization` that renders the RGB image handed to the 2-D encoder (backproject.py:88-100,
// segment.py:197-208; gsplat-1.4.0 spherical_harmonics + `clamp_min(colors + 0.5, 0)`).
//   dir = normalise(mean - camera position);  colour = max(sum_k Y_k(dir) * coeff[k] + 0.5, 0)
// One thread per (Gaussian, channel): HBM-bound, reads (degree+1)^2 coefficients x 3 channels per Gaussian.
// The basis is the real-SH polynomial table of the 3DGS code base (degree <= 4); oracle/gsplat_oracle.py::sh_basis
// derives the same functions from associated Legendre polynomials, so the two are independent.
#include "common.cuh"

namespace gwbp {

__global__ void __launch_bounds__(256) sh_colors_kernel(int64_t n, int degree, const float *__restrict__ means,
                                                        const float *__restrict__ coeffs, int64_t sN, int64_t sK,
                                                        int64_t sC, float cx, float cy, float cz,
                                                        float *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3 * n) return;
    const int64_t g = idx / 3;
    const int c = (int)(idx - 3 * g);
    float x = means[3 * g] - cx, y = means[3 * g + 1] - cy, z = means[3 * g + 2] - cz;
    const float inorm = rsqrtf(fmaxf(x * x + y * y + z * z, 1e-30f));
    x *= inorm; y *= inorm; z *= inorm;
    const float *sh = coeffs + g * sN + c * sC;
#define SH(k) __ldg(sh + (k) * sK)
    float r = 0.2820947917738781f * SH(0);
    if (degree >= 1) {
        r += 0.48860251190292f * (-y * SH(1) + z * SH(2) - x * SH(3));
        if (degree >= 2) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            r += 1.092548430592079f * xy * SH(4) - 1.092548430592079f * yz * SH(5) +
                 0.3153915652525201f * (2.0f * zz - xx - yy) * SH(6) - 1.092548430592079f * xz * SH(7) +
                 0.5462742152960395f * (xx - yy) * SH(8);
            if (degree >= 3) {
                r += -0.5900435899266435f * y * (3.0f * xx - yy) * SH(9) + 2.890611442640554f * xy * z * SH(10) -
                     0.4570457994644658f * y * (4.0f * zz - xx - yy) * SH(11) +
                     0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SH(12) -
                     0.4570457994644658f * x * (4.0f * zz - xx - yy) * SH(13) +
                     1.445305721320277f * z * (xx - yy) * SH(14) - 0.5900435899266435f * x * (xx - 3.0f * yy) * SH(15);
                if (degree >= 4) {
                    r += 2.5033429417967046f * xy * (xx - yy) * SH(16) - 1.7701307697799304f * yz * (3.0f * xx - yy) * SH(17) +
                         0.9461746957575601f * xy * (7.0f * zz - 1.0f) * SH(18) -
                         0.6690465435572892f * yz * (7.0f * zz - 3.0f) * SH(19) +
                         0.10578554691520431f * (zz * (35.0f * zz - 30.0f) + 3.0f) * SH(20) -
                         0.6690465435572892f * xz * (7.0f * zz - 3.0f) * SH(21) +
                         0.47308734787878004f * (xx - yy) * (7.0f * zz - 1.0f) * SH(22) -
                         1.7701307697799304f * xz * (xx - 3.0f * yy) * SH(23) +
                         0.6258357354491761f * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy)) * SH(24);
                }
            }
        }
    }
#undef SH
    out[idx] = fmaxf(r + 0.5f, 0.0f);
}

int launch_sh_colors(int64_t n, int degree, const float *means, const float *coeffs, int64_t sN, int64_t sK, int64_t sC,
                     const float *cam_pos_host, float *out, cudaStream_t st) {
    if (n == 0) return 0;
    const int64_t work = 3 * n;
    sh_colors_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(n, degree, means, coeffs, sN, sK, sC, cam_pos_host[0],
                                                                   cam_pos_host[1], cam_pos_host[2], out);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
