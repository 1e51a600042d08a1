// Batched Prophet predict for sm_100a.  Replaces the per-model body of
// forecast_time_series_udf (reference src/jobs/prophet_scorer.py:35-102): fbprophet 0.5
// Prophet.predict = predict_trend (piecewise_linear / piecewise_logistic) +
// predict_seasonal_components + yhat = trend*(1+multiplicative)+additive, then the
// scorer's int truncation and floor clamp (prophet_scorer.py:73-84), and -- when
// uncertainty_samples > 0 -- predict_uncertainty (sample_posterior_predictive ->
// sample_model -> sample_predictive_trend, percentiles over the draws).
//
// One CTA per (model, tile of future points).  Output is the HBM-bound part:
// 8 B timestamp in, 8..28 B per forecast point out.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/prophet_b200.h"

namespace pb200 {

struct PredictArgs {
    const double* params;
    const double* tchange;
    const int* meta_i32;
    const long long* meta_i64;
    const double* meta_f64;
    const long long* future_ds;
    const double* floor;
    const double* cap;
    int n_models, horizon;
    int smax, kmax, pstride;
    int growth, mult;
    double* yhat;
    double* trend;           // optional (may be null)
    int* yhat_int;
};

constexpr double PI_FL = 3.141592653589793;

// sin / cos of a Fourier argument.  With t in days since 1970 the arguments are 1e4 ... 1e6 radians (daily harmonics of a
// 2022 date: 2 pi 4 19000 = 4.8e5), beyond the 1.05e5 up to which the library's sincos reduces with its short
// Cody-Waite path: every daily and most weekly terms took the Payne-Hanek slow path, and 14 of those per point made
// predict_kernel ~7 % of its HBM roofline.  Reduce here instead: 2 pi in three parts, the first two
// of 32 significant bits, so that k C1 and k C2 are exact for k < 2^21 and r = x - k 2 pi is good to 4.4e-16 absolute
// (checked in exact rational arithmetic over +-8e6); then the library call sees |r| <= pi.  The argument x itself stays
// the ROUNDED double numpy forms -- the reduction is of that value, not of the mathematical angle.
__device__ __forceinline__ void sincos_reduced(const double x, double* s, double* c) {
    const double k = rint(x * 0.15915494309189535);
    if (fabs(k) < 2097152.0) {
        double r = fma(-k, 6.2831853069365025, x);             // 0x1.921fb544p+2 : exact
        r = fma(-k, 2.4308402025215864e-10, r);                // 0x1.0b4611a6p-32
        r = fma(-k, 8.089064995183803e-21, r);                 // 0x1.3198a2e037073p-67
        sincos(r, s, c);
    } else {
        sincos(x, s, c);
    }
}

// Prophet.fourier_series evaluated exactly as numpy does: fun(2.0 * (i + 1) * np.pi * t / period)
__device__ __forceinline__ double seas_dot(const double tau, const double period, const int order, const double* beta) {
    double acc = 0.0;
    for (int i = 0; i < order; ++i) {
        const double arg = (2.0 * (double)(i + 1)) * PI_FL * tau / period;
        double s, c;
        sincos_reduced(arg, &s, &c);
        acc = fma(s, beta[2 * i], acc);
        acc = fma(c, beta[2 * i + 1], acc);
    }
    return acc;
}

// shared per-model state used by both the deterministic and the MC kernels
struct ModelSm {
    double k, m, sigma, y_scale, floor, cap_s, t_scale, lam;
    long long start;
    int S, mask, status, K;
    double delta[32];
    double tc[32];
    double gamma[32];
    double beta[40];
};

__device__ __forceinline__ void load_model(ModelSm& ms, const PredictArgs& a, const int model, const int tid, const int nt) {
    const int* mi = a.meta_i32 + (size_t)model * 8;
    const double* pr = a.params + (size_t)model * a.pstride;
    if (tid == 0) {
        ms.S = mi[1];
        ms.mask = mi[3];
        ms.status = mi[4];
        ms.start = a.meta_i64[(size_t)model * 2];
        ms.t_scale = (double)a.meta_i64[(size_t)model * 2 + 1];
        ms.y_scale = a.meta_f64[(size_t)model * 4];
        const bool logi = a.growth == PB200_GROWTH_LOGISTIC;
        const double fl = logi ? a.floor[model] : 0.0;
        ms.floor = fl;
        ms.cap_s = logi ? (a.cap[model] - fl) / ms.y_scale : 0.0;
        ms.k = pr[0];
        ms.m = pr[1];
        ms.sigma = pr[2];
        int K = 0;
        if (ms.mask & 1) K += 20;
        if (ms.mask & 2) K += 6;
        if (ms.mask & 4) K += 8;
        ms.K = K;
    }
    for (int s = tid; s < 32; s += nt) {
        ms.delta[s] = s < a.smax ? pr[3 + s] : 0.0;
        ms.tc[s] = s < a.smax ? a.tchange[(size_t)model * a.smax + s] : 0.0;
    }
    for (int q = tid; q < 40; q += nt) ms.beta[q] = q < a.kmax ? pr[3 + a.smax + q] : 0.0;
    __syncthreads();
    if (tid == 0) {
        const int S = ms.S;
        // Prophet.piecewise_logistic gammas / piecewise_linear gammas; lam = mean|delta| + 1e-8
        double acc = 0.0, kc = ms.k, ad = 0.0;
        for (int s = 0; s < S; ++s) {
            const double kn = kc + ms.delta[s];
            double g;
            if (a.growth == PB200_GROWTH_LOGISTIC) {
                g = (ms.tc[s] - ms.m - acc) * (1.0 - kc / kn);
                acc += g;
            } else {
                g = -ms.tc[s] * ms.delta[s];
            }
            ms.gamma[s] = g;
            kc = kn;
            ad += fabs(ms.delta[s]);
        }
        ms.lam = ad / (double)S + 1e-8;
    }
    __syncthreads();
}

__device__ __forceinline__ double seasonal_term(const ModelSm& ms, const long long d) {
    const double tau = (1e-9 * (double)d) / 86400.0;
    double acc = 0.0;
    int col = 0;
    if (ms.mask & 1) { acc += seas_dot(tau, 365.25, 10, ms.beta + col); col += 20; }
    if (ms.mask & 2) { acc += seas_dot(tau, 7.0, 3, ms.beta + col); col += 6; }
    if (ms.mask & 4) { acc += seas_dot(tau, 1.0, 4, ms.beta + col); col += 8; }
    return acc;
}

__global__ void __launch_bounds__(256) predict_kernel(const PredictArgs a) {
    __shared__ ModelSm ms;
    const int model = blockIdx.x;
    const int tid = threadIdx.x;
    load_model(ms, a, model, tid, blockDim.x);
    const bool ok = ms.status >= 0;
    const int S = ms.S;
    for (int h = blockIdx.y * blockDim.x + tid; h < a.horizon; h += gridDim.y * blockDim.x) {
        const size_t o = (size_t)model * a.horizon + h;
        if (!ok) {
            a.yhat[o] = NAN;
            if (a.trend) a.trend[o] = NAN;
            a.yhat_int[o] = INT32_MIN;
            continue;
        }
        const long long d = a.future_ds[o];
        const double t = (double)(d - ms.start) / ms.t_scale;
        double kt = ms.k, mt = ms.m;
        for (int s = 0; s < S; ++s) {
            if (t >= ms.tc[s]) {
                kt += ms.delta[s];
                mt += ms.gamma[s];
            }
        }
        double tr;
        if (a.growth == PB200_GROWTH_LOGISTIC) tr = ms.cap_s / (1.0 + exp(-kt * (t - mt)));
        else tr = kt * t + mt;
        tr = tr * ms.y_scale + ms.floor;
        const double sd = ms.K > 0 ? seasonal_term(ms, d) : 0.0;
        const double yh = a.mult ? tr * (1.0 + sd) : tr + sd * ms.y_scale;
        a.yhat[o] = yh;
        if (a.trend) a.trend[o] = tr;
        // prophet_scorer.py:73 astype(int) truncates toward zero; :76-84 values < floor -> floor
        double yt = trunc(yh);
        const double fcfg = a.floor[model];
        if (yt < fcfg) yt = fcfg;
        yt = fmin(fmax(yt, -2147483648.0), 2147483647.0);
        a.yhat_int[o] = (int)yt;
    }
}

__global__ void make_future_kernel(const long long* last_ds, long long n_models, int horizon, long long freq_ns,
                                   long long* out) {
    const long long n = n_models * (long long)horizon;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long mdl = i / horizon;
        const int j = (int)(i - mdl * horizon);
        out[i] = last_ds[mdl] + (long long)(j + 1) * freq_ns;
    }
}

}  // namespace pb200
