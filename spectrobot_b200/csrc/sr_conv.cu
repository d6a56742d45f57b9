// sr_conv.cu -- instrument convolution of hi-res spectra to low-res channels on the device.
//
// Reference: SpectralIntensity.hires_to_lowres -> SpectralObject.convolve_to_grid_from_irregular
// (spect_classes.py:1180-1191, 883-918) with gaussian() (:1926-1934) and conv_single() (:1162-1164):
// for every channel (centre f, width w) the hi-res points with f - n_sigma*w <= x <= f + n_sigma*w
// are multiplied by the normalised Gaussian 1/(w sqrt(2 pi)) exp(-((x-f)/w)^2/2) and integrated
// with the trapezoid rule over the (possibly irregular) hi-res grid; a channel that sees no hi-res
// point is 0 (:903-905).  Doing this on the device turns a LOS from 1.2e6 doubles into ~1e2 before
// it crosses PCIe.
#include <cmath>
#include "sr_common.h"

namespace {

constexpr int CONV_NT = 256;
constexpr int CONV_SPB = 4;   // spectra per CTA: the Gaussian weights are computed once for them

__global__ void __launch_bounds__(CONV_NT) k_convolve_lowres(
    const double* __restrict__ x, long n_pts, const double* __restrict__ spec, int n_spec,
    const double* __restrict__ centre, const double* __restrict__ width, int n_chan,
    double n_sigma, double* __restrict__ out) {
    const int c = blockIdx.x, s0 = blockIdx.y * CONV_SPB;
    const double f = centre[c], w = width[c];
    const double lo = f - n_sigma * w, hi = f + n_sigma * w;
    // first index with x >= lo, last index with x <= hi (grid ascending)
    long a = 0, b = n_pts;
    while (a < b) { const long m = (a + b) >> 1; if (x[m] < lo) a = m + 1; else b = m; }
    const long i0 = a;
    b = n_pts;
    while (a < b) { const long m = (a + b) >> 1; if (x[m] <= hi) a = m + 1; else b = m; }
    const long i1 = a - 1;
    const double fac = 1.0 / (w * sqrt(2.0 * M_PI));
    double acc[CONV_SPB];
#pragma unroll
    for (int q = 0; q < CONV_SPB; q++) acc[q] = 0.0;
    // trapezoid segments i .. i+1, i in [i0, i1)
    for (long i = i0 + threadIdx.x; i < i1; i += CONV_NT) {
        const double xa = x[i], xb = x[i + 1];
        const double ta = (xa - f) / w, tb = (xb - f) / w;
        const double ga = fac * exp(-0.5 * (ta * ta)), gb = fac * exp(-0.5 * (tb * tb));
        const double hd = (xb - xa) * 0.5;
#pragma unroll
        for (int q = 0; q < CONV_SPB; q++) {
            if (s0 + q < n_spec) {
                const double* __restrict__ y = spec + (size_t)(s0 + q) * n_pts;
                acc[q] = fma(hd, fma(y[i], ga, y[i + 1] * gb), acc[q]);
            }
        }
    }
    __shared__ double red[CONV_SPB][CONV_NT / 32];
#pragma unroll
    for (int q = 0; q < CONV_SPB; q++) {
        double v = acc[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < CONV_SPB && s0 + threadIdx.x < n_spec) {
        double v = 0.0;
        for (int k = 0; k < CONV_NT / 32; k++) v += red[threadIdx.x][k];
        out[(size_t)(s0 + threadIdx.x) * n_chan + c] = v;
    }
}

}  // namespace

extern "C" {

int sr_convolve_lowres_dev(const double* grid_dev, long n_pts, const double* spec_dev, int n_spec,
                           const double* centre_dev, const double* width_dev, int n_chan,
                           double n_sigma, double* out_dev, void* stream) {
    if (!grid_dev || !spec_dev || !centre_dev || !width_dev || !out_dev || n_pts < 1 ||
        n_spec < 1 || n_chan < 1 || !(n_sigma > 0.0))
        return sr::fail(SR_ERR_ARG, "sr_convolve_lowres_dev: bad argument");
    dim3 grid((unsigned)n_chan, (unsigned)((n_spec + CONV_SPB - 1) / CONV_SPB));
    SR_LAUNCH(k_convolve_lowres, grid, CONV_NT, 0, (cudaStream_t)stream, grid_dev, n_pts, spec_dev,
              n_spec, centre_dev, width_dev, n_chan, n_sigma, out_dev);
    return SR_OK;
}

int sr_convolve_lowres_host(const double* grid, long n_pts, const double* spec, int n_spec,
                            const double* centre, const double* width, int n_chan,
                            double n_sigma, double* out) {
    if (!grid || !spec || !centre || !width || !out || n_pts < 1 || n_spec < 1 || n_chan < 1)
        return sr::fail(SR_ERR_ARG, "sr_convolve_lowres_host: bad argument");
    for (int c = 0; c < n_chan; c++)
        if (!(width[c] > 0.0)) return sr::fail(SR_ERR_ARG, "channel %d: width %g", c, width[c]);
    sr::DevBuf<double> dx, dy, dc, dw, dout;
    SR_CUDA(dx.upload(grid, (size_t)n_pts));
    SR_CUDA(dy.upload(spec, (size_t)n_pts * n_spec));
    SR_CUDA(dc.upload(centre, n_chan));
    SR_CUDA(dw.upload(width, n_chan));
    SR_CUDA(dout.alloc((size_t)n_spec * n_chan));
    int rc = sr_convolve_lowres_dev(dx.p, n_pts, dy.p, n_spec, dc.p, dw.p, n_chan, n_sigma,
                                    dout.p, nullptr);
    if (rc) return rc;
    SR_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * n_spec * n_chan, cudaMemcpyDeviceToHost));
    return SR_OK;
}

}  // extern "C"
