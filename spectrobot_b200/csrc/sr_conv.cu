// sr_conv.cu -- instrument convolution of hi-res spectra to low-res channels on the device.
//
// Reference: SpectralIntensity.hires_to_lowres -> SpectralObject.convolve_to_grid_from_irregular
// (spect_classes.py:1180-1191, 883-918) with gaussian() (:1926-1934) and conv_single() (:1162-1164):
// for every channel (centre f, width w) the hi-res points with f - n_sigma*w <= x <= f + n_sigma*w
// are multiplied by the normalised Gaussian 1/(w sqrt(2 pi)) exp(-((x-f)/w)^2/2) and integrated
// with the trapezoid rule over the (possibly irregular) hi-res grid; a channel that sees no hi-res
// point is 0 (:903-905).  Doing this on the device turns a LOS from 1.2e6 doubles into ~1e2 before
// it crosses PCIe.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "sr_common.h"

namespace {

constexpr int CONV_NT = 256;
constexpr int CONV_SLAB = 512;   // spectra per launch pair (bounds the partial-sum workspace)
// CONV_SPB spectra per CTA (the Gaussian weights are computed once for them) x CONV_PB trapezoid
// segments per CTA are template parameters; (8, 1024) measured best on B200 (SR_CONV_CFG)

// per channel: segments [i0, i1) of the trapezoid rule = hi-res points i0..i1 inside the window
// units == SR_CHAN_NM_FROM_CM1: the grid is in cm-1 and the channels in nm.  The reference converts
// the hi-res spectrum first (convert_grid_to('nm'), spect_classes.py:771-778: grid -> 1e7/grid,
// spectrum -> spectrum*grid^2*1e-7, both reversed so that the grid ascends) and convolves on the
// wavelength axis.  Here the axis is mirrored instead of reversed - abscissa u = -1e7/x ascends with
// x, the channel sits at -centre - which leaves the Gaussian (even in u - f) and every trapezoid
// (|du| and the pair of ordinates) unchanged.
__device__ __forceinline__ double conv_abscissa(double x, int units) {
    return units == SR_CHAN_NM_FROM_CM1 ? -1.0e7 / x : x;
}

__global__ void k_convolve_windows(const double* __restrict__ x, long n_pts,
                                   const double* __restrict__ centre,
                                   const double* __restrict__ width, int n_chan, double n_sigma,
                                   int units, long* __restrict__ win) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chan) return;
    const double f = units == SR_CHAN_NM_FROM_CM1 ? -centre[c] : centre[c], w = width[c];
    const double lo = f - n_sigma * w, hi = f + n_sigma * w;
    long a = 0, b = n_pts;   // first index with u >= lo, last index with u <= hi (u ascending)
    while (a < b) { const long m = (a + b) >> 1; if (conv_abscissa(x[m], units) < lo) a = m + 1; else b = m; }
    win[2 * c] = a;
    b = n_pts;
    while (a < b) { const long m = (a + b) >> 1; if (conv_abscissa(x[m], units) <= hi) a = m + 1; else b = m; }
    win[2 * c + 1] = a - 1;
}

// One CTA = CONV_PB consecutive trapezoid segments x CONV_SPB spectra, staged ONCE in shared
// memory (every hi-res value is read from HBM exactly once; the overlapping channel windows
// re-read it from shared memory), the Gaussian weight of a (channel, point) is computed once per
// CTA, and the per-block sums go to partial[channel][block][spectrum]; k_convolve_reduce adds the
// blocks in a fixed order, so the result does not depend on scheduling (no atomics).
//   sum_i hd_i (y_i g_i + y_{i+1} g_{i+1})  =  sum_p y_p g_p (hd_{p-1} [p > s0] + hd_p [p < s1])
template <int CONV_SPB, int CONV_PB>
__global__ void __launch_bounds__(CONV_NT) k_convolve_lowres(
    const double* __restrict__ x, long n_pts, const double* __restrict__ spec, int n_spec,
    const double* __restrict__ centre, const double* __restrict__ width, int n_chan,
    const long* __restrict__ win, int n_blocks, int units, double* __restrict__ partial) {
    const long bstart = (long)blockIdx.x * CONV_PB;
    const long bend = min(bstart + CONV_PB, n_pts - 1);
    const int np = (int)(bend - bstart) + 1;          // points bstart .. bend
    const int sp0 = blockIdx.y * CONV_SPB;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    extern __shared__ __align__(16) double csm[];
    double* xs = csm;                                  // [CONV_PB + 1]
    double* ys = csm + CONV_PB + 1;                    // [CONV_SPB][CONV_PB + 1]
    __shared__ double red[CONV_SPB][CONV_NT / 32];
    // the channels whose window reaches into this block, ascending (one test per thread and a
    // ballot compaction instead of every thread scanning all channels); more than CONV_NT
    // channels fall back to the scan
    __shared__ int clist[CONV_NT];
    __shared__ int wcount[CONV_NT / 32];
    const bool listed = n_chan <= CONV_NT;
    int n_list = n_chan;
    if (listed) {
        const int t = threadIdx.x;
        const bool hit = t < n_chan && min(win[2 * t + 1], bend) > max(win[2 * t], bstart);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) wcount[wid] = __popc(m);
        __syncthreads();
        int base = 0;
        n_list = 0;
#pragma unroll
        for (int w = 0; w < CONV_NT / 32; w++) {
            base += w < wid ? wcount[w] : 0;
            n_list += wcount[w];
        }
        if (hit) clist[base + __popc(m & ((1u << lane) - 1u))] = t;
        if (n_list == 0) return;                       // uniform over the CTA
    } else {
        bool any = false;
        for (int c = 0; c < n_chan; c++)
            any = any || min(win[2 * c + 1], bend) > max(win[2 * c], bstart);
        if (!any) return;                              // uniform over the CTA
    }
    // asynchronous global -> shared copies (LDGSTS): every thread has all its 8-byte copies in
    // flight at once instead of one load per dependent shared-memory store
    auto cp8 = [](double* dst, const double* src) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    };
    for (int i = threadIdx.x; i < np; i += CONV_NT) cp8(xs + i, x + bstart + i);
#pragma unroll
    for (int q = 0; q < CONV_SPB; q++) {
        if (sp0 + q < n_spec) {
            const double* __restrict__ y = spec + (size_t)(sp0 + q) * n_pts + bstart;
            for (int i = threadIdx.x; i < np; i += CONV_NT) cp8(ys + q * (CONV_PB + 1) + i, y + i);
        } else {
            for (int i = threadIdx.x; i < np; i += CONV_NT) ys[q * (CONV_PB + 1) + i] = 0.0;
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (units == SR_CHAN_NM_FROM_CM1) {   // spectrum per nm on the mirrored wavelength axis
        for (int i = threadIdx.x; i < np; i += CONV_NT) {
            const double xv = xs[i], jac = xv * xv * 1.e-7;
#pragma unroll
            for (int q = 0; q < CONV_SPB; q++) ys[q * (CONV_PB + 1) + i] *= jac;
            xs[i] = -1.0e7 / xv;
        }
        __syncthreads();
    }
    for (int ci = 0; ci < n_list; ci++) {              // (clist is published by the barrier above)
        const int c = listed ? clist[ci] : ci;
        const long s0l = max(win[2 * c], bstart), s1l = min(win[2 * c + 1], bend);
        if (s1l <= s0l) continue;                      // uniform over the CTA
        const int s0 = (int)(s0l - bstart), s1 = (int)(s1l - bstart);
        const double f = units == SR_CHAN_NM_FROM_CM1 ? -centre[c] : centre[c], w = width[c];
        const double fac = 1.0 / (w * sqrt(2.0 * M_PI));
        double acc[CONV_SPB];
#pragma unroll
        for (int q = 0; q < CONV_SPB; q++) acc[q] = 0.0;
        for (int p = s0 + threadIdx.x; p <= s1; p += CONV_NT) {
            const double xp = xs[p];
            double hd = 0.0;
            if (p > s0) hd += (xp - xs[p - 1]) * 0.5;
            if (p < s1) hd += (xs[p + 1] - xp) * 0.5;
            const double t = (xp - f) / w;
            const double g = fac * exp(-0.5 * (t * t)) * hd;
#pragma unroll
            for (int q = 0; q < CONV_SPB; q++) acc[q] = fma(ys[q * (CONV_PB + 1) + p], g, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < CONV_SPB; q++) {
            double v = acc[q];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
            if (lane == 0) red[q][wid] = v;
        }
        __syncthreads();
        if (threadIdx.x < CONV_SPB && sp0 + threadIdx.x < n_spec) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < CONV_NT / 32; k++) v += red[threadIdx.x][k];
            partial[((size_t)c * n_blocks + blockIdx.x) * n_spec + sp0 + threadIdx.x] = v;
        }
        __syncthreads();
    }
}

__global__ void k_convolve_reduce(const double* __restrict__ partial, const long* __restrict__ win,
                                  int n_spec, int n_chan, int n_blocks, long n_pts, int CONV_PB,
                                  double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (s >= n_spec) return;
    const long i0 = win[2 * c], i1 = min(win[2 * c + 1], n_pts - 1);
    double v = 0.0;
    if (i1 > i0) {   // blocks holding the segments [i0, i1)
        const int b0 = (int)(i0 / CONV_PB), b1 = (int)((i1 - 1) / CONV_PB);
        for (int b = b0; b <= b1; b++) v += partial[((size_t)c * n_blocks + b) * n_spec + s];
    }
    out[(size_t)s * n_chan + c] = v;
}

}  // namespace

extern "C" {

int sr_convolve_lowres_dev(const double* grid_dev, long n_pts, const double* spec_dev, int n_spec,
                           const double* centre_dev, const double* width_dev, int n_chan,
                           double n_sigma, double* out_dev, void* stream) {
    sr_channels ch{n_chan, centre_dev, width_dev, n_sigma, SR_CHAN_SAME_UNITS};
    return sr_convolve_channels_dev(grid_dev, n_pts, spec_dev, n_spec, &ch, out_dev, stream);
}

int sr_convolve_channels_dev(const double* grid_dev, long n_pts, const double* spec_dev, int n_spec,
                             const sr_channels* ch, double* out_dev, void* stream) {
    if (!ch) return sr::fail(SR_ERR_ARG, "sr_convolve_channels_dev: bad argument");
    const double* centre_dev = ch->centre_dev;
    const double* width_dev = ch->width_dev;
    const int n_chan = ch->n_chan, units = ch->units;
    const double n_sigma = ch->n_sigma;
    if (!grid_dev || !spec_dev || !centre_dev || !width_dev || !out_dev || n_pts < 1 ||
        n_spec < 1 || n_chan < 1 || !(n_sigma > 0.0) ||
        (units != SR_CHAN_SAME_UNITS && units != SR_CHAN_NM_FROM_CM1))
        return sr::fail(SR_ERR_ARG, "sr_convolve_channels_dev: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    static const int cfg = getenv("SR_CONV_CFG") ? atoi(getenv("SR_CONV_CFG")) : 0;   // tuning aid
    const int spb = (cfg == 1 || cfg == 3) ? 4 : 8;
    const int pb = cfg == 2 ? 512 : (cfg == 3 ? 2048 : 1024);
    const int n_blocks = (int)std::max<long>(1, (n_pts - 1 + pb - 1) / pb);
    const int slab = std::min(n_spec, CONV_SLAB);
    // stream-ordered scratch: channel windows + partial sums [n_chan][n_blocks][slab]
    long* win = nullptr;
    double* partial = nullptr;
    sr::pool_keep();
    SR_CUDA(cudaMallocAsync(&win, sizeof(long) * 2 * n_chan, st));
    cudaError_t e = cudaMallocAsync(&partial, sizeof(double) * (size_t)n_chan * n_blocks * slab, st);
    if (e != cudaSuccess) {
        cudaFreeAsync(win, st);
        return sr::fail(SR_ERR_CUDA, "sr_convolve_lowres_dev: %s", cudaGetErrorString(e));
    }
    auto launch = [&](auto kern, int ns, const double* sp, dim3 grid) -> int {
        const size_t smem = (size_t)(spb + 1) * (pb + 1) * sizeof(double);
        SR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SR_LAUNCH(kern, grid, CONV_NT, smem, st, grid_dev, n_pts, sp, ns, centre_dev, width_dev, n_chan,
                  win, n_blocks, units, partial);
        return SR_OK;
    };
    auto body = [&]() -> int {
        SR_LAUNCH(k_convolve_windows, (n_chan + 63) / 64, 64, 0, st, grid_dev, n_pts, centre_dev,
                  width_dev, n_chan, n_sigma, units, win);
        for (int s0 = 0; s0 < n_spec; s0 += slab) {
            const int ns = std::min(slab, n_spec - s0);
            dim3 grid((unsigned)n_blocks, (unsigned)((ns + spb - 1) / spb));
            const double* sp = spec_dev + (size_t)s0 * n_pts;
            int rc;
            if (cfg == 1) rc = launch(k_convolve_lowres<4, 1024>, ns, sp, grid);
            else if (cfg == 2) rc = launch(k_convolve_lowres<8, 512>, ns, sp, grid);
            else if (cfg == 3) rc = launch(k_convolve_lowres<4, 2048>, ns, sp, grid);
            else rc = launch(k_convolve_lowres<8, 1024>, ns, sp, grid);
            if (rc) return rc;
            SR_LAUNCH(k_convolve_reduce, dim3((unsigned)((ns + 127) / 128), (unsigned)n_chan), 128, 0,
                      st, partial, win, ns, n_chan, n_blocks, n_pts, pb, out_dev + (size_t)s0 * n_chan);
        }
        return SR_OK;
    };
    const int rc = body();
    cudaFreeAsync(partial, st);
    cudaFreeAsync(win, st);
    return rc;
}

int sr_convolve_lowres_host(const double* grid, long n_pts, const double* spec, int n_spec,
                            const double* centre, const double* width, int n_chan,
                            double n_sigma, double* out) {
    return sr_convolve_channels_host(grid, n_pts, spec, n_spec, centre, width, n_chan, n_sigma,
                                     SR_CHAN_SAME_UNITS, out);
}

int sr_convolve_channels_host(const double* grid, long n_pts, const double* spec, int n_spec,
                              const double* centre, const double* width, int n_chan,
                              double n_sigma, int units, double* out) {
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
    sr_channels ch{n_chan, dc.p, dw.p, n_sigma, units};
    int rc = sr_convolve_channels_dev(dx.p, n_pts, dy.p, n_spec, &ch, dout.p, nullptr);
    if (rc) return rc;
    SR_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * n_spec * n_chan, cudaMemcpyDeviceToHost));
    return SR_OK;
}

}  // extern "C"
