// Sky statistics after the hot path (reference utils/sky.py; used at L1_to_L2/gen_cal_image.py:639-651):
//   rip_medfit_solve    the pixel-independent part of sky.medfit (:137-175): Legendre basis at the centres of the N x N
//                       regions, normal equations over the regions with a finite median, their solution, and the
//                       Legendre polynomials on the pixel grid that rip_medfit_eval_dev contracts with the coefficients
//   rip_bin_masked_dev  binkxk(np.where(~mask, slope, nan), k) (:20-43 as called at gen_cal_image.py:641)
//   rip_gauss_hist_dev  the smoothed histogram of sky.smooth_mode (:78-83): sum_i exp(-0.5 ((z_j - a_i) / w)^2), NaNs skipped
// Everything here is float64 host arithmetic or a plain reduction kernel; sums are reduced in a fixed order (per-CTA
// partials, then one CTA), so results are reproducible from run to run.
#include <cmath>
#include <vector>

#include "rip_rt.h"

namespace rip {

// P_0 .. P_order at z with scipy.special.legendre_p's recurrence and rounding order:
//   P_k = ((2k-1)/k * z) * P_{k-1} + (-(k-1)/k) * P_{k-2}      (verified bit-identical for k <= 6: tests/test_sky_host.py)
static void legendre_all(int order, double z, double* p) {
    p[0] = 1.0;
    if (order >= 1) p[1] = z;
    for (int k = 2; k <= order; ++k) {
        const double fac0 = -(double)(k - 1) / (double)k, fac1 = (double)(2 * k - 1) / (double)k;
        const double a = fac1 * z;
        const double t1 = a * p[k - 1], t0 = fac0 * p[k - 2];
        p[k] = t1 + t0;
    }
}

// np.linspace(start, stop, num)[i] as NumPy computes it: i * step + start with step = (stop - start) / (num - 1), last = stop
static double linspace_at(double start, double stop, int num, int i) {
    if (num == 1) return start;
    if (i == num - 1) return stop;
    const double step = (stop - start) / (double)(num - 1);
    const double t = (double)i * step;
    return t + start;
}

// dense solve A x = b, n <= 45, partial pivoting (the LU of LAPACK's dgesv in its unblocked form)
static bool lu_solve(int n, std::vector<double>& A, std::vector<double>& b) {
    std::vector<int> piv(n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = std::fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(A[i * n + k]) > best) { best = std::fabs(A[i * n + k]); p = i; }
        if (!(best > 0.0)) return false;
        if (p != k) {
            for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]);
            std::swap(b[k], b[p]);
        }
        const double r = 1.0 / A[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            const double l = A[i * n + k] * r;
            A[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
            b[i] -= l * b[k];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * b[j];
        b[i] = s / A[i * n + i];
    }
    return true;
}

__global__ void bin_masked_kernel(const float* __restrict__ arr, const uint8_t* __restrict__ mask, int ny, int nx, int k,
                                  float* __restrict__ out) {
    const int nxo = nx / k, nyo = ny / k;
    const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y;
    if (xo >= nxo || yo >= nyo) return;
    // np.mean of the k x k block in float32 pairwise order is not reproduced bit for bit (the value only seeds a
    // histogram mode); float32 accumulation row by row
    float s = 0.0f;
    bool bad = false;
    for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) {
            const long p = (long)(yo * k + dy) * nx + (xo * k + dx);
            if (mask && mask[p]) bad = true;
            s += arr[p];
        }
    out[(long)yo * nxo + xo] = bad ? __int_as_float(0x7fc00000) : s / (float)(k * k);
}

constexpr int GH_MAXZ = 32, GH_THREADS = 256;
// partial[b][j] = sum over the elements of CTA b of exp(-0.5 ((z_j - a) / w)^2)
__global__ void __launch_bounds__(GH_THREADS) gauss_hist_partial_kernel(const float* __restrict__ arr, long count, int nz, double inv_w,
                                                                        const double* __restrict__ z, double* __restrict__ partial) {
    __shared__ double zs[GH_MAXZ];
    __shared__ double red[GH_THREADS / 32][GH_MAXZ];
    if (threadIdx.x < nz) zs[threadIdx.x] = z[threadIdx.x];
    __syncthreads();
    double acc[GH_MAXZ];
#pragma unroll
    for (int j = 0; j < GH_MAXZ; ++j) acc[j] = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
        const float a = arr[i];
        if (a != a) continue;
        const double ad = (double)a;
#pragma unroll
        for (int j = 0; j < GH_MAXZ; ++j)
            if (j < nz) {
                const double t = (zs[j] - ad) * inv_w;
                acc[j] += exp(-0.5 * (t * t));
            }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < GH_MAXZ; ++j) {
        double v = acc[j];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < nz) {
        double v = 0.0;
        for (int w = 0; w < GH_THREADS / 32; ++w) v += red[w][threadIdx.x];
        partial[(long)blockIdx.x * GH_MAXZ + threadIdx.x] = v;
    }
}
__global__ void gauss_hist_final_kernel(const double* __restrict__ partial, int nblk, int nz, double* __restrict__ out) {
    const int j = threadIdx.x;
    if (j >= nz) return;
    double v = 0.0;
    for (int b = 0; b < nblk; ++b) v += partial[(long)b * GH_MAXZ + j];
    out[j] = v;
}

}  // namespace rip

using namespace rip;

extern "C" int rip_medfit_solve(int ny, int nx, int N, int order, const float* meds, double* coef, double* LPX, double* LPY) {
    RIP_API_BEGIN
    RIP_REQUIRE(meds && coef && ny > 0 && nx > 0, "rip_medfit_solve: null argument");
    RIP_REQUIRE(order >= 0 && order <= 8, "rip_medfit_solve: order %d outside 0..8", order);
    RIP_REQUIRE(N >= 1 && nx / N >= 1 && ny / N >= 1, "rip_medfit_solve: %d regions do not fit a %d x %d array", N, ny, nx);
    const int kx = nx / N, ky = ny / N, px = (nx % N) / 2, py = (ny % N) / 2;
    const int nc = (order + 1) * (order + 2) / 2;
    // region centres in [-1, 1): u = 2 (px - 0.5 + kx c_i) / nx - 1 with c = np.linspace(0.5, N - 0.5, N)   (utils/sky.py:141-142)
    std::vector<double> PU((size_t)N * (order + 1)), PV((size_t)N * (order + 1));
    for (int i = 0; i < N; ++i) {
        const double c = linspace_at(0.5, (double)N - 0.5, N, i);
        const double tu = (double)kx * c, tv = (double)ky * c;
        const double su = ((double)px - 0.5) + tu, sv = ((double)py - 0.5) + tv;
        const double u = (2.0 * su) / (double)nx - 1.0, v = (2.0 * sv) / (double)ny - 1.0;
        legendre_all(order, u, &PU[(size_t)i * (order + 1)]);
        legendre_all(order, v, &PV[(size_t)i * (order + 1)]);
    }
    std::vector<double> A((size_t)nc * nc, 0.0), b(nc, 0.0), basis(nc);
    for (int ipix = 0; ipix < N; ++ipix)      // x index of the region (outer loop of the reference, :158-162)
        for (int jpix = 0; jpix < N; ++jpix) {  // y index
            const float m = meds[jpix * N + ipix];
            if (m != m) continue;
            int k = 0;
            for (int i = 0; i <= order; ++i)
                for (int j = 0; j <= order - i; ++j) basis[k++] = PU[(size_t)ipix * (order + 1) + i] * PV[(size_t)jpix * (order + 1) + j];
            for (int r = 0; r < nc; ++r) {
                for (int c2 = 0; c2 < nc; ++c2) A[(size_t)r * nc + c2] += basis[r] * basis[c2];
                b[r] += (double)m * basis[r];
            }
        }
    RIP_REQUIRE(lu_solve(nc, A, b), "rip_medfit_solve: singular normal equations (too few regions with a finite median)");
    for (int k = 0; k < nc; ++k) coef[k] = b[k];
    // Legendre polynomials on the pixel grid, u_x = np.linspace(-1, 1 - 2/nx, nx)   (:167-175)
    if (LPX) {
        std::vector<double> p(order + 1);
        for (int x = 0; x < nx; ++x) {
            legendre_all(order, linspace_at(-1.0, 1.0 - 2.0 / (double)nx, nx, x), p.data());
            for (int i = 0; i <= order; ++i) LPX[(size_t)i * nx + x] = p[i];
        }
    }
    if (LPY) {
        std::vector<double> p(order + 1);
        for (int y = 0; y < ny; ++y) {
            legendre_all(order, linspace_at(-1.0, 1.0 - 2.0 / (double)ny, ny, y), p.data());
            for (int j = 0; j <= order; ++j) LPY[(size_t)j * ny + y] = p[j];
        }
    }
    RIP_API_END
}

extern "C" int rip_bin_masked_dev(int device, const float* d_arr, const uint8_t* d_mask, int ny, int nx, int k, float* d_out,
                                  void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_arr && d_out && k >= 1 && ny >= k && nx >= k, "rip_bin_masked_dev: bad argument");
    use_device(device);
    dim3 grid((nx / k + 127) / 128, ny / k);
    RIP_LAUNCH(bin_masked_kernel, grid, 128, 0, (cudaStream_t)stream, d_arr, d_mask, ny, nx, k, d_out);
    RIP_API_END
}

extern "C" int rip_gauss_hist_dev(int device, const float* d_arr, long count, const double* z, int nz, double width, double* sums,
                                  void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_arr && z && sums && count >= 1 && nz >= 1 && nz <= GH_MAXZ, "rip_gauss_hist_dev: bad argument (nz <= %d)", GH_MAXZ);
    RIP_REQUIRE(width > 0.0, "rip_gauss_hist_dev: width must be positive");
    use_device(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (int)std::min<long>(592, (count + GH_THREADS - 1) / GH_THREADS);
    DevBuf<double> dz(GH_MAXZ), part((size_t)nblk * GH_MAXZ), out(GH_MAXZ);
    RIP_CUDA(cudaMemcpyAsync(dz.p, z, nz * sizeof(double), cudaMemcpyHostToDevice, st));
    RIP_LAUNCH(gauss_hist_partial_kernel, nblk, GH_THREADS, 0, st, d_arr, count, nz, 1.0 / width, (const double*)dz.p, part.p);
    RIP_LAUNCH(gauss_hist_final_kernel, 1, GH_MAXZ, 0, st, (const double*)part.p, nblk, nz, out.p);
    RIP_CUDA(cudaMemcpyAsync(sums, out.p, nz * sizeof(double), cudaMemcpyDeviceToHost, st));
    RIP_CUDA(cudaStreamSynchronize(st));
    RIP_API_END
}
