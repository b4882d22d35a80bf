// Pixel solid angle from a FITS zenithal (TAN / STG) WCS with SIP distortion, on the device.
// Replaces the host computation of AreaFactor in calibrateimage (reference L1_to_L2/gen_cal_image.py:618-622 ->
// utils/coordutils.py:17-82): instead of uploading a 67 MB plane per exposure the caller uploads the ~1.7 kB of WCS
// coefficients and the plane is produced where it is consumed.  Same construction as the reference: world coordinates
// on the (N+2)^2 grid of pixel positions -1 .. N, equal-area reprojection about the pole of the image's hemisphere,
// central differences, |det J|.  float64 throughout; one CTA evaluates a (TX+2) x (TY+2) patch of the grid into
// shared memory and differentiates it (the WCS is evaluated 1.27x per pixel, nothing touches HBM but the output).
#include "rip_rt.h"

namespace rip {

constexpr int AREA_MAXSIP = 9;
constexpr int AREA_NHEAD = 11;
constexpr int AREA_NW = AREA_NHEAD + 2 * (AREA_MAXSIP + 1) * (AREA_MAXSIP + 1);
constexpr int AREA_TX = 64, AREA_TY = 8;

struct AreaWcs {
    double w[AREA_NW];  // coordutils.FitsWCS.pack(): crpix(2) crval(2) cd(4) lonpole proj order | A[10][10] | B[10][10]
};

// (u, v) of the reference's equal-area plane for 0-based pixel position (x, y); `flip` = the reference's dec[0] > 0
__host__ __device__ inline void area_uv(const AreaWcs& W, double x, double y, bool flip, double& uo, double& vo, double* dec_out = nullptr) {
    const double deg = 3.14159265358979323846 / 180.0;
    const double* h = W.w;
    const double u = x + 1.0 - h[0], v = y + 1.0 - h[1];
    const int order = (int)h[10];
    double f = 0.0, g = 0.0;
    if (order > 0) {
        const double* A = W.w + AREA_NHEAD;
        const double* B = A + (AREA_MAXSIP + 1) * (AREA_MAXSIP + 1);
        for (int p = order; p >= 0; --p) {
            double ca = 0.0, cb = 0.0;
            for (int q = order - p; q >= 0; --q) {
                ca = ca * v + A[p * (AREA_MAXSIP + 1) + q];
                cb = cb * v + B[p * (AREA_MAXSIP + 1) + q];
            }
            f = f * u + ca;
            g = g * u + cb;
        }
    }
    const double uu = u + f, vv = v + g;
    const double xi = h[4] * uu + h[5] * vv, eta = h[6] * uu + h[7] * vv;
    const double r = hypot(xi, eta) * deg;
    const double phi = atan2(xi, -eta);
    const double theta = (h[9] == 0.0) ? atan2(1.0, r) : (3.14159265358979323846 / 2.0 - 2.0 * atan(r / 2.0));
    const double dp = h[3] * deg, dphi = phi - h[8] * deg;
    double st, ct, sdp, cdp, sph, cph;
    sincos(theta, &st, &ct);
    sincos(dp, &sdp, &cdp);
    sincos(dphi, &sph, &cph);
    double sd = st * sdp + ct * cdp * cph;
    sd = sd < -1.0 ? -1.0 : (sd > 1.0 ? 1.0 : sd);
    const double dec = asin(sd);
    const double ra = h[2] * deg + atan2(-ct * sph, st * cdp - ct * sdp * cph);
    if (dec_out) *dec_out = dec;
    const double th = flip ? (3.14159265358979323846 / 2.0 - dec) : (3.14159265358979323846 / 2.0 + dec);
    const double rho = 2.0 * sin(th / 2.0);
    double sr, cr;
    sincos(ra, &sr, &cr);
    uo = rho * cr;
    vo = rho * sr;
}

template <typename TO>
__global__ void __launch_bounds__(AREA_TX* AREA_TY) pixel_area_kernel(const AreaWcs W, int N, int flip, double inv_omega, TO* __restrict__ out) {
    __shared__ double su[AREA_TY + 2][AREA_TX + 2], sv[AREA_TY + 2][AREA_TX + 2];
    const int x0 = blockIdx.x * AREA_TX, y0 = blockIdx.y * AREA_TY;
    const int t = threadIdx.y * AREA_TX + threadIdx.x;
    // grid index k <-> pixel position k - 1 (np.linspace(-1, N, N + 2))
    for (int i = t; i < (AREA_TX + 2) * (AREA_TY + 2); i += AREA_TX * AREA_TY) {
        const int ly = i / (AREA_TX + 2), lx = i % (AREA_TX + 2);
        double u, v;
        area_uv(W, (double)(x0 + lx - 1), (double)(y0 + ly - 1), flip != 0, u, v);
        su[ly][lx] = u;
        sv[ly][lx] = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= N || y >= N) return;
    const int lx = threadIdx.x + 1, ly = threadIdx.y + 1;
    const double J11 = (su[ly][lx + 1] - su[ly][lx - 1]) / 2.0, J12 = (su[ly + 1][lx] - su[ly - 1][lx]) / 2.0;
    const double J21 = (sv[ly][lx + 1] - sv[ly][lx - 1]) / 2.0, J22 = (sv[ly + 1][lx] - sv[ly - 1][lx]) / 2.0;
    const double a = fabs(__dsub_rn(__dmul_rn(J11, J22), __dmul_rn(J21, J12)));
    out[(size_t)y * N + x] = (TO)(a * inv_omega);
}

// out: device plane [N,N] of Area * inv_omega (pass inv_omega = 1 / pars.Omega_ideal for AreaFactor, 1 for steradians)
void launch_pixel_area(const double* wcs, int nwcs, int N, double inv_omega, void* d_out, int out_dtype, cudaStream_t st) {
    RIP_REQUIRE(wcs && nwcs == AREA_NW, "pixel area: the packed WCS must hold %d doubles (got %d)", AREA_NW, nwcs);
    RIP_REQUIRE(N >= 1 && N <= 16384, "pixel area: N=%d out of range", N);
    RIP_REQUIRE(out_dtype == RIP_F32 || out_dtype == RIP_F64, "pixel area: output dtype must be f32 or f64");
    AreaWcs W;
    memcpy(W.w, wcs, sizeof W.w);
    RIP_REQUIRE(W.w[10] >= 0 && W.w[10] <= AREA_MAXSIP, "pixel area: SIP order out of range");
    double u, v, dec0;
    area_uv(W, -1.0, -1.0, false, u, v, &dec0);  // the reference picks the hemisphere of the first grid point
    const int flip = dec0 > 0.0 ? 1 : 0;
    dim3 grid((N + AREA_TX - 1) / AREA_TX, (N + AREA_TY - 1) / AREA_TY), block(AREA_TX, AREA_TY);
    if (out_dtype == RIP_F64) RIP_LAUNCH(pixel_area_kernel<double>, grid, block, 0, st, W, N, flip, inv_omega, (double*)d_out);
    else RIP_LAUNCH(pixel_area_kernel<float>, grid, block, 0, st, W, N, flip, inv_omega, (float*)d_out);
}

}  // namespace rip

using namespace rip;

extern "C" int rip_pixel_area_dev(int device, const double* wcs, int nwcs, int N, double inv_omega, void* d_out, int out_dtype,
                                  void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_out, "rip_pixel_area_dev: null output");
    use_device(device);
    launch_pixel_area(wcs, nwcs, N, inv_omega, d_out, out_dtype, (cudaStream_t)stream);
    RIP_API_END
}

extern "C" int rip_pixel_area_host(int device, const double* wcs, int nwcs, int N, double inv_omega, void* out, int out_dtype) {
    RIP_API_BEGIN
    RIP_REQUIRE(out, "rip_pixel_area_host: null output");
    use_device(device);
    DevRaw d;
    const size_t bytes = (size_t)N * N * dtype_size(out_dtype);
    d.alloc(bytes);
    launch_pixel_area(wcs, nwcs, N, inv_omega, d.p, out_dtype, 0);
    RIP_CUDA(cudaMemcpy(out, d.p, bytes, cudaMemcpyDeviceToHost));
    RIP_API_END
}
