/*
 * prophet_oracle.c -- plain-C CPU restatement of the batched-Prophet hot path.
 * TEST / BASELINE INFRASTRUCTURE ONLY: linked by nothing in time_series_spark_b200/.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
 *
 * PARITY UNPINNED (same caveat as prophet_oracle.py): fbprophet 0.5 / PyStan 2.19.1.1
 * (reference environment.yml:12-13) are not installable here; this file restates their
 * published algorithm and is cross-checked against the independent numpy restatement
 * (tests/test_oracle_c.py), not against fbprophet itself.
 *
 * Follows, function by function:
 *   po_prepare      Prophet.fit -> setup_dataframe / initialize_scales / set_auto_seasonalities /
 *                   fourier_series / set_changepoints / {linear,logistic}_growth_init
 *                   (reached from reference src/jobs/prophet_modeler.py:65-66)
 *   po_eval         python/stan/unix/prophet.stan, model block, negated (propto, no Jacobian),
 *                   with a hand-written reverse-mode gradient.  Unlike Stan's dense A*delta it
 *                   uses per-segment sums (A is a step matrix) -- the same O(T*K) algorithm the
 *                   GPU kernel uses, so the CPU baseline is not handicapped.
 *   po_lbfgs        stan/optimization/bfgs.hpp BFGSMinimizer::step + lbfgs_update.hpp +
 *                   bfgs_linesearch.hpp (WolfeLineSearch, WolfLSZoom, CubicInterp)
 *   po_newton       stan/services/optimize/newton.hpp + stan/optimization/newton.hpp (newton_step,
 *                   make_negative_definite_and_solve) + stan/model/grad_hess_log_prob.hpp -- what
 *                   fbprophet 0.5's fit() falls back to when PyStan raises on a line-search failure
 *   po_fit_batch    the per-group loop Spark runs (one group per task, prophet_modeler.py:139-141),
 *                   here an OpenMP loop over series on the host cores.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PO_MAXP 96
#define PO_MAXS 32
#define PO_HIST 5
#define NS_DAY (86400LL * 1000000000LL)

typedef struct {
    int growth;          /* 0 linear, 1 logistic */
    int multiplicative;
    int n_changepoints;
    double changepoint_range, tau, seas_prior;
    int yearly, weekly, daily;      /* -1 auto, 0 off, 1 on */
    int max_iter;
    double init_alpha, tol_obj, tol_rel_obj, tol_grad, tol_rel_grad, tol_param;
    int algorithm;       /* 0 L-BFGS, Newton retry on line-search failure (fbprophet 0.5 fit()); 1 L-BFGS only; 2 Newton only */
    int reserved;
} po_opts;

typedef struct {
    int T, S, K, ncp, logistic, mult, mask;
    double *t, *y, *X;               /* X is T x K row-major */
    int bidx[PO_MAXS + 1];           /* first point index of segment j+1 */
    double tc[PO_MAXS];
    double cap_s, tau, seas_prior;
    double y_scale, floor;
    long long start, span;
} po_prep;

static void po_default(po_opts* o) {
    o->growth = 1; o->multiplicative = 1; o->n_changepoints = 25; o->changepoint_range = 0.8;
    o->tau = 0.05; o->seas_prior = 10.0; o->yearly = o->weekly = o->daily = -1; o->max_iter = 10000;
    o->init_alpha = 1e-3; o->tol_obj = 1e-12; o->tol_rel_obj = 1e4; o->tol_grad = 1e-8;
    o->tol_rel_grad = 1e7; o->tol_param = 1e-8; o->algorithm = 0; o->reserved = 0;
}

void po_default_options(po_opts* o) { po_default(o); }

/* returns 0 ok, <0 status as in include/prophet_b200.h */
static int po_prepare(const long long* ds, const double* yraw, int T, double floor_, double cap, const po_opts* o,
                      po_prep* p, double* theta0) {
    memset(p, 0, sizeof *p);
    if (T < 2) return -3;
    p->T = T; p->logistic = o->growth == 1; p->mult = o->multiplicative != 0;
    const double fl = p->logistic ? floor_ : 0.0;
    double amax = 0.0, ymin = INFINITY, ymax = -INFINITY;
    long long mindt = INT64_MAX;
    for (int i = 0; i < T; ++i) {
        const double v = yraw[i];
        if (!isfinite(v)) return -5;
        if (fabs(v - fl) > amax) amax = fabs(v - fl);
        if (v < ymin) ymin = v;
        if (v > ymax) ymax = v;
        if (i > 0) {
            const long long dt = ds[i] - ds[i - 1];
            if (dt < 0) return -5;
            if (dt != 0 && dt < mindt) mindt = dt;
        }
    }
    const long long start = ds[0], last = ds[T - 1], span = last - start;
    if (span <= 0) return -5;
    if (p->logistic && !(cap > fl)) return -4;
    double y_scale = amax == 0.0 ? 1.0 : amax;
    p->y_scale = y_scale; p->floor = fl; p->start = start; p->span = span;
    p->cap_s = p->logistic ? (cap - fl) / y_scale : 0.0;
    p->tau = o->tau; p->seas_prior = o->seas_prior;
    const int has_dt = mindt != INT64_MAX;
    const int ydis = span < 730 * NS_DAY;
    const int wdis = (span < 14 * NS_DAY) || (has_dt && mindt >= 7 * NS_DAY);
    const int ddis = (span < 2 * NS_DAY) || (has_dt && mindt >= NS_DAY);
    int mask = 0;
    if (o->yearly < 0 ? !ydis : o->yearly > 0) mask |= 1;
    if (o->weekly < 0 ? !wdis : o->weekly > 0) mask |= 2;
    if (o->daily < 0 ? !ddis : o->daily > 0) mask |= 4;
    p->mask = mask;
    const int K = ((mask & 1) ? 20 : 0) + ((mask & 2) ? 6 : 0) + ((mask & 4) ? 8 : 0);
    p->K = K > 0 ? K : 1;
    p->t = (double*)malloc(sizeof(double) * T);
    p->y = (double*)malloc(sizeof(double) * T);
    p->X = (double*)calloc((size_t)T * p->K, sizeof(double));
    const double dspan = (double)span;
    for (int i = 0; i < T; ++i) {
        p->t[i] = (double)(ds[i] - start) / dspan;
        p->y[i] = (yraw[i] - fl) / y_scale;
        if (K > 0) {
            const double tau_d = (1e-9 * (double)ds[i]) / 86400.0;
            int col = 0;
            const double periods[3] = {365.25, 7.0, 1.0};
            const int orders[3] = {10, 3, 4};
            for (int q = 0; q < 3; ++q) {
                if (!(mask & (1 << q))) continue;
                for (int h = 0; h < orders[q]; ++h) {
                    const double arg = 2.0 * (h + 1) * 3.141592653589793 * tau_d / periods[q];
                    p->X[(size_t)i * p->K + col++] = sin(arg);
                    p->X[(size_t)i * p->K + col++] = cos(arg);
                }
            }
        }
    }
    /* changepoints */
    int hist = (int)floor((double)T * o->changepoint_range);
    int ncp = o->n_changepoints;
    if (ncp + 1 > hist) ncp = hist - 1;
    if (ncp < 0) ncp = 0;
    p->ncp = ncp;
    if (ncp > 0) {
        p->S = ncp;
        const double step = (double)(hist - 1) / (double)ncp;
        for (int s = 0; s < ncp; ++s) {
            int idx = s == ncp - 1 ? hist - 1 : (int)rint((double)(s + 1) * step);
            p->tc[s] = p->t[idx];
            int b = idx;
            while (b > 0 && p->t[b - 1] >= p->tc[s]) --b;
            p->bidx[s] = b;
        }
    } else {
        p->S = 1; p->tc[0] = 0.0; p->bidx[0] = 0;
    }
    /* initial point */
    int i1 = T - 1;
    while (i1 > 0 && ds[i1 - 1] == last) --i1;
    const int P = p->S + p->K + 3;
    for (int q = 0; q < P; ++q) theta0[q] = 0.0;
    const double y0 = p->y[0], y1 = p->y[i1], Tsp = p->t[i1] - p->t[0];
    if (p->logistic) {
        const double C0 = p->cap_s;
        const double yy0 = fmax(0.01 * C0, fmin(0.99 * C0, y0)), yy1 = fmax(0.01 * C0, fmin(0.99 * C0, y1));
        double r0 = C0 / yy0; const double r1 = C0 / yy1;
        if (fabs(r0 - r1) <= 0.01) r0 = 1.05 * r0;
        const double L0 = log(r0 - 1.0), L1 = log(r1 - 1.0);
        theta0[1] = L0 * Tsp / (L0 - L1);
        theta0[0] = (L0 - L1) / Tsp;
    } else {
        theta0[0] = (y1 - y0) / Tsp;
        theta0[1] = y0 - theta0[0] * p->t[0];
    }
    if (!p->logistic && ymin == ymax) return 50;
    return 0;
}

static void po_free(po_prep* p) { free(p->t); free(p->y); free(p->X); }

/* objective + gradient; returns 0 ok, nonzero = Stan ModelAdaptor error */
static int po_eval(const po_prep* p, const double* th, double* f_out, double* g) {
    const int S = p->S, K = p->K, T = p->T, Kreal = p->mask ? K : 0;
    const double k = th[0], m = th[1], u = th[2 + S];
    const double* delta = th + 2;
    const double* beta = th + 3 + S;
    double kc[PO_MAXS + 1], mc[PO_MAXS + 1], rho[PO_MAXS], U[PO_MAXS + 1], V[PO_MAXS + 1];
    for (int q = 0; q < S + K + 3; ++q) if (!isfinite(th[q])) return 1;
    const double sigma = exp(u);
    if (!(sigma > 0.0) || !isfinite(sigma)) return 1;
    double cum = 0.0;
    kc[0] = k;
    for (int s = 0; s < S; ++s) { cum += delta[s]; kc[s + 1] = k + cum; }
    mc[0] = m;
    if (p->logistic) {
        for (int s = 0; s < S; ++s) {
            rho[s] = kc[s] / kc[s + 1];
            const double gam = (p->tc[s] - mc[s]) * (1.0 - rho[s]);
            mc[s + 1] = mc[s] + gam;
        }
    } else {
        double c2 = 0.0;
        for (int s = 0; s < S; ++s) { c2 += -p->tc[s] * delta[s]; mc[s + 1] = m + c2; }
    }
    double gb[PO_MAXP];
    for (int q = 0; q < K; ++q) gb[q] = 0.0;
    for (int j = 0; j <= S; ++j) U[j] = V[j] = 0.0;
    double ss = 0.0;
    int j = 0;
    for (int i = 0; i < T; ++i) {
        while (j < S && i >= p->bidx[j]) ++j;
        const double* x = p->X + (size_t)i * K;
        double dot = 0.0;
        for (int q = 0; q < Kreal; ++q) dot += x[q] * beta[q];
        const double t = p->t[i], tm = t - mc[j];
        double gtr, sig = 0.0;
        if (p->logistic) { sig = 1.0 / (1.0 + exp(-(kc[j] * tm))); gtr = p->cap_s * sig; }
        else gtr = kc[j] * t + mc[j];
        const double opm = p->mult ? 1.0 + dot : 1.0;
        const double yhat = p->mult ? gtr * opm : gtr + dot;
        if (!isfinite(yhat)) return 1;
        const double r = p->y[i] - yhat;
        ss += r * r;
        const double cb = p->mult ? r * gtr : r;
        for (int q = 0; q < Kreal; ++q) gb[q] += cb * x[q];
        const double qv = r * opm;
        if (p->logistic) { const double dz = qv * gtr * (1.0 - sig); U[j] += dz * tm; V[j] += dz; }
        else { U[j] += qv * t; V[j] += qv; }
    }
    const double inv_s2 = 1.0 / (sigma * sigma), scale = -inv_s2;
    double gk, gm, gd[PO_MAXS];
    if (p->logistic) {
        double kbar[PO_MAXS + 1], gmc[PO_MAXS + 1], rbar[PO_MAXS];
        for (int q = 0; q <= S; ++q) { kbar[q] = scale * U[q]; gmc[q] = scale * (-kc[q]) * V[q]; }
        double abar = gmc[S];
        for (int s = S - 1; s >= 0; --s) { rbar[s] = abar * (mc[s] - p->tc[s]); abar = gmc[s] + rho[s] * abar; }
        for (int s = 0; s < S; ++s) { kbar[s] += rbar[s] / kc[s + 1]; kbar[s + 1] += -(rbar[s] * rho[s]) / kc[s + 1]; }
        double tot = 0.0;
        for (int q = S; q >= 0; --q) { tot += kbar[q]; if (q >= 1) gd[q - 1] = tot; }
        gk = tot + k / 25.0;
        gm = abar + m / 25.0;
    } else {
        double totU = 0.0, totV = 0.0;
        for (int q = 0; q <= S; ++q) { totU += U[q]; totV += V[q]; }
        double su = 0.0, sv = 0.0;
        for (int s = S - 1; s >= 0; --s) { su += U[s + 1]; sv += V[s + 1]; gd[s] = scale * (su - p->tc[s] * sv); }
        gk = scale * totU + k / 25.0;
        gm = scale * totV + m / 25.0;
    }
    double ad = 0.0, pb = 0.0;
    for (int s = 0; s < S; ++s) {
        const double d = delta[s];
        ad += fabs(d);
        g[2 + s] = gd[s] + (d > 0 ? 1.0 : (d < 0 ? -1.0 : 0.0)) / p->tau;
    }
    const double isg = Kreal ? 1.0 / (p->seas_prior * p->seas_prior) : 1.0;
    for (int q = 0; q < K; ++q) {
        g[3 + S + q] = scale * gb[q] + beta[q] * isg;
        pb += 0.5 * beta[q] * beta[q] * isg;
    }
    g[0] = gk; g[1] = gm;
    g[2 + S] = -ss * inv_s2 + (double)T + 4.0 * sigma * sigma;
    const double f = 0.5 * ss * inv_s2 + (double)T * u + k * k / 50.0 + m * m / 50.0 + ad / p->tau + 2.0 * sigma * sigma + pb;
    *f_out = f;
    for (int q = 0; q < S + K + 3; ++q) if (!isfinite(g[q])) return 3;
    return isfinite(f) ? 0 : 2;
}

static double vdot(const double* a, const double* b, int n) { double s = 0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s; }

static double cubic_interp(double df0, double x1, double f1, double df1, double loX, double hiX) {
    const double c3 = (-12 * f1 + 6 * x1 * (df0 + df1)) / (x1 * x1 * x1);
    const double c2 = -(4 * df0 + 2 * df1) / x1 + 6 * f1 / (x1 * x1);
    const double c1 = df0;
    const double t_s = sqrt(c2 * c2 - 2.0 * c1 * c3);
    const double s1 = -(c2 + t_s) / c3, s2 = -(c2 - t_s) / c3;
    double minF = loX * (loX * (loX * c3 / 3.0 + c2) / 2.0 + c1), minX = loX;
    double tmpF = hiX * (hiX * (hiX * c3 / 3.0 + c2) / 2.0 + c1);
    if (tmpF < minF) { minF = tmpF; minX = hiX; }
    if (loX < s1 && s1 < hiX) { tmpF = s1 * (s1 * (s1 * c3 / 3.0 + c2) / 2.0 + c1); if (tmpF < minF) { minF = tmpF; minX = s1; } }
    if (loX < s2 && s2 < hiX) { tmpF = s2 * (s2 * (s2 * c3 / 3.0 + c2) / 2.0 + c1); if (tmpF < minF) { minF = tmpF; minX = s2; } }
    return minX;
}

/* returns Stan termination code; x holds the result */
static int po_lbfgs(const po_prep* p, const po_opts* o, double* x, double* f_final, int* iters_out, int* nevals_out,
                    double* trace, int trace_cap) {
    const int P = p->S + p->K + 3;
    const double eps = 2.220446049250313e-16;
    const double c1 = 1e-4, c2 = 0.9, minAlpha = 1e-12;
    const int maxLSIts = 20, maxLSRestarts = 10;
    double g[PO_MAXP], pk[PO_MAXP], xt[PO_MAXP], gt[PO_MAXP], pp[PO_MAXP], gprev[PO_MAXP], xprev[PO_MAXP];
    double HY[PO_HIST][PO_MAXP], HS[PO_HIST][PO_MAXP], hrho[PO_HIST], halpha[PO_HIST];
    int nev = 0, it = 0, hn = 0, hhead = 0;
    double fk, ft = 0, fk_1 = 0, alphak_1 = 0, alpha = 0;
    int err = po_eval(p, x, &fk, g); ++nev;
    *iters_out = 0; *nevals_out = nev; *f_final = fk;
    if (err) return -2;
    for (int q = 0; q < P; ++q) pk[q] = -g[q];
    for (;;) {
        ++it;
        int resetB = it == 1 ? 1 : 0;
        for (;;) {
            if (resetB) for (int q = 0; q < P; ++q) pk[q] = -g[q];
            const double dfp = vdot(g, pk, P);
            if (it > 1 && resetB != 2) alpha = fmin(1.0, 1.01 * cubic_interp(vdot(gprev, pp, P), alphak_1, fk - fk_1, dfp, minAlpha, 1.0));
            else alpha = o->init_alpha;
            int ret = 0;
            {
                const double c1dfp = c1 * dfp, c2dfp = c2 * dfp;
                double alpha0 = minAlpha, prevF = fk, prevDFp = dfp;
                int nits = 0, lsR = 0, zoom = 0;
                double alo = 0, aloF = 0, aloD = 0, ahi = 0, ahiF = 0, ahiD = 0;
                for (;;) {
                    if (nits >= maxLSIts) { ret = 1; break; }
                    for (int q = 0; q < P; ++q) xt[q] = x[q] + alpha * pk[q];
                    err = po_eval(p, xt, &ft, gt); ++nev;
                    if (err) { if (lsR >= maxLSRestarts) { ret = 1; break; } alpha = 0.5 * (alpha0 + alpha); ++lsR; continue; }
                    lsR = 0;
                    const double nd = vdot(gt, pk, P);
                    if (ft > fk + alpha * c1dfp || (ft >= prevF && nits > 0)) { zoom = 1; alo = alpha0; aloF = prevF; aloD = prevDFp; ahi = alpha; ahiF = ft; ahiD = nd; break; }
                    if (fabs(nd) <= -c2dfp) { ret = 0; break; }
                    if (nd >= 0) { zoom = 1; alo = alpha; aloF = ft; aloD = nd; ahi = alpha0; ahiF = prevF; ahiD = prevDFp; break; }
                    alpha0 = alpha; prevF = ft; prevDFp = nd; alpha *= 10.0; ++nits;
                }
                if (zoom) {
                    int itNum = 0; ret = 0;
                    for (;;) {
                        ++itNum;
                        if (fabs(alo - ahi) < 1e-16) { ret = 1; break; }
                        { /* [guard, not in Stan] bracket = adjacent doubles: upstream would spin forever */
                            const double mid = 0.5 * (alo + ahi);
                            if (mid == alo || mid == ahi) { ret = 1; break; }
                        }
                        if (itNum % 5 == 0) alpha = 0.5 * (alo + ahi);
                        else {
                            const double d1 = aloD + ahiD - 3 * (aloF - ahiF) / (alo - ahi);
                            double d2 = sqrt(d1 * d1 - aloD * ahiD);
                            if (ahi < alo) d2 = -d2;
                            alpha = ahi - (ahi - alo) * (ahiD + d2 - d1) / (ahiD - aloD + 2 * d2);
                            const double lo = fmin(alo, ahi), hi = fmax(alo, ahi);
                            if (!isfinite(alpha) || alpha < lo + 0.01 * fabs(alo - ahi) || alpha > hi - 0.01 * fabs(alo - ahi)) alpha = 0.5 * (alo + ahi);
                        }
                        int giveup = 0;
                        for (;;) {
                            for (int q = 0; q < P; ++q) xt[q] = x[q] + alpha * pk[q];
                            err = po_eval(p, xt, &ft, gt); ++nev;
                            if (!err) break;
                            alpha = 0.5 * (alpha + fmin(alo, ahi));
                            if (fabs(fmin(alo, ahi) - alpha) < 1e-16) { giveup = 1; break; }
                        }
                        if (giveup) { ret = 1; break; }
                        const double nd = vdot(gt, pk, P);
                        if (ft > (fk + alpha * c1dfp) || ft >= aloF) { ahi = alpha; ahiF = ft; ahiD = nd; }
                        else {
                            if (fabs(nd) <= -c2dfp) break;
                            if (nd * (ahi - alo) >= 0) { ahi = alo; ahiF = aloF; ahiD = aloD; }
                            alo = alpha; aloF = ft; aloD = nd;
                        }
                    }
                }
            }
            if (ret) {
                if (resetB) { *iters_out = it; *nevals_out = nev; *f_final = fk; return -1; }
                resetB = 2;
                continue;
            }
            break;
        }
        /* accept */
        memcpy(xprev, x, sizeof(double) * P); memcpy(gprev, g, sizeof(double) * P); memcpy(pp, pk, sizeof(double) * P);
        memcpy(x, xt, sizeof(double) * P); memcpy(g, gt, sizeof(double) * P);
        fk_1 = fk; fk = ft;
        if (trace && it <= trace_cap) {     /* row it-1: iteration, f_k, alpha_k, evaluations so far */
            double* tr = trace + (size_t)(it - 1) * 4;
            tr[0] = (double)it; tr[1] = fk; tr[2] = alpha; tr[3] = (double)nev;
        }
        if (resetB) { hn = 0; hhead = 0; }
        int slot;
        if (hn < PO_HIST) { slot = (hhead + hn) % PO_HIST; ++hn; } else { slot = hhead; hhead = (hhead + 1) % PO_HIST; }
        double sy = 0, yy = 0, ssn = 0, gg = 0;
        for (int q = 0; q < P; ++q) {
            const double sv = x[q] - xprev[q], yv = g[q] - gprev[q];
            HS[slot][q] = sv; HY[slot][q] = yv;
            sy += sv * yv; yy += yv * yv; ssn += sv * sv; gg += g[q] * g[q];
        }
        if (resetB) { const double B0 = yy / sy; for (int q = 0; q < P; ++q) pp[q] /= B0; alphak_1 = alpha * B0; }
        else alphak_1 = alpha;
        const double gammak = sy / yy;
        hrho[slot] = 1.0 / sy;
        for (int q = 0; q < P; ++q) pk[q] = -g[q];
        for (int h = hn - 1; h >= 0; --h) {
            const int sl = (hhead + h) % PO_HIST;
            const double al = hrho[sl] * vdot(HS[sl], pk, P);
            for (int q = 0; q < P; ++q) pk[q] -= al * HY[sl][q];
            halpha[sl] = al;
        }
        for (int q = 0; q < P; ++q) pk[q] *= gammak;
        for (int h = 0; h < hn; ++h) {
            const int sl = (hhead + h) % PO_HIST;
            const double be = hrho[sl] * vdot(HY[sl], pk, P);
            const double cf = halpha[sl] - be;
            for (int q = 0; q < P; ++q) pk[q] += cf * HS[sl][q];
        }
        const double df = fabs(fk_1 - fk), gp = vdot(g, pk, P);
        int st = 0;
        if (df < o->tol_obj) st = 20;
        else if (df < o->tol_rel_obj * eps * fmax(fabs(fk_1), fmax(fabs(fk), 1.0))) st = 21;
        else if (sqrt(gg) < o->tol_grad) st = 30;
        else if (fabs(gp) < o->tol_rel_grad * eps * fmax(fabs(fk), 1.0)) st = 31;
        else if (sqrt(ssn) < o->tol_param) st = 10;
        else if (it >= o->max_iter) st = 40;
        if (st) { *iters_out = it; *nevals_out = nev; *f_final = fk; return st; }
    }
}


/* ---- Newton fallback (stan::optimization::newton_step and what it calls) ---- */

/* eigen-decomposition of the symmetric n x n matrix A (row-major, destroyed) by cyclic Jacobi rotations:
 * on return the diagonal of A holds the eigenvalues and the COLUMNS of V the eigenvectors */
static void po_jacobi(double* A, double* V, int n) {
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) { diag += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j]; }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s_ = t * c;
                for (int k = 0; k < n; ++k) {      /* columns p, q of A and V */
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s_ * akq; A[k * n + q] = s_ * akp + c * akq;
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s_ * vkq; V[k * n + q] = s_ * vkp + c * vkq;
                }
                for (int k = 0; k < n; ++k) {      /* rows p, q of A */
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s_ * aqk; A[q * n + k] = s_ * apk + c * aqk;
                }
            }
    }
}

/* returns a Stan-style status: 60 = model from the Newton run, -1 = an evaluation inside grad_hess_log_prob failed
 * (Stan throws; PyStan raises RuntimeError a second time and the reference UDF drops the series) */
static int po_newton(const po_prep* p, const po_opts* o, double* x, double* f_final, int* iters_out, int* nevals_out) {
    const int P = p->S + p->K + 3;
    const double eps = 1e-3, pert[4] = {-2 * eps, -eps, eps, 2 * eps}, coef[4] = {1.0 / 12.0, -2.0 / 3.0, 2.0 / 3.0, -1.0 / 12.0};
    const double half_inv_eps = 0.5 / eps;   /* see the note in prophet_oracle.py::_grad_hess on Stan's half_epsilon */
    double* H = (double*)malloc(sizeof(double) * P * P * 2);
    double* V = H + P * P;
    double g[PO_MAXP], gi[PO_MAXP], xp[PO_MAXP], u[PO_MAXP], w[PO_MAXP], xn[PO_MAXP];
    int nev = 0, it = 0, st = 60;
    double f, ftmp;
    int err = po_eval(p, x, &f, g); ++nev;
    if (err) f = INFINITY;
    for (it = 1; it <= o->max_iter; ++it) {
        double f0;
        err = po_eval(p, x, &f0, g); ++nev;
        if (err) { st = -1; break; }
        for (int q = 0; q < P * P; ++q) H[q] = 0.0;
        memcpy(xp, x, sizeof(double) * P);
        for (int d = 0; d < P && st == 60; ++d) {
            for (int i = 0; i < 4; ++i) {
                xp[d] = x[d] + pert[i];
                err = po_eval(p, xp, &ftmp, gi); ++nev;
                if (err) { st = -1; break; }
                for (int dd = 0; dd < P; ++dd) { const double inc = half_inv_eps * coef[i] * gi[dd]; H[d * P + dd] += inc; H[dd * P + d] += inc; }
            }
            xp[d] = x[d];
        }
        if (st != 60) break;
        po_jacobi(H, V, P);
        for (int j = 0; j < P; ++j) { double s_ = 0.0; for (int k = 0; k < P; ++k) s_ += V[k * P + j] * g[k]; w[j] = s_ / fabs(H[j * P + j]); }
        for (int k = 0; k < P; ++k) { double s_ = 0.0; for (int j = 0; j < P; ++j) s_ += V[k * P + j] * w[j]; u[k] = s_; }
        double step = 2.0, f1 = INFINITY;
        int moved = 0;
        for (;;) {
            step *= 0.5;
            if (step < 1e-50) break;
            for (int q = 0; q < P; ++q) xn[q] = x[q] - step * u[q];
            err = po_eval(p, xn, &f1, gi); ++nev;
            if (err || !(f1 <= f0)) continue;
            moved = 1;
            break;
        }
        const double last = f;
        if (moved) { memcpy(x, xn, sizeof(double) * P); f = f1; } else f = f0;
        if (it > 1 && fabs(f - last) < 1e-8) break;
    }
    if (it > o->max_iter) it = o->max_iter;
    free(H);
    *f_final = f; *iters_out = it; *nevals_out = nev;
    return st;
}

/* objective + gradient at a caller-supplied theta (for cross-checks); returns err */
int po_objective(const long long* ds, const double* y, int T, double floor_, double cap, const po_opts* o,
                 const double* theta, double* f, double* g, int* S_out, int* K_out) {
    po_prep p; double th0[PO_MAXP];
    int st = po_prepare(ds, y, T, floor_, cap, o, &p, th0);
    if (st < 0) return st;
    *S_out = p.S; *K_out = p.K;
    int err = po_eval(&p, theta, f, g);
    po_free(&p);
    return err;
}

/*
 * Fits n series.  theta_out rows (stride pstride) hold Stan's unconstrained optimum
 * k, m, delta[S], log sigma, beta[K]; info rows: status, iters, n_evals, S, K.
 */
int po_fit_batch_trace(const long long* ds, const double* y, const long long* offsets, int n, double floor_, double cap_multiplier,
                       const po_opts* o, double* theta_out, int pstride, double* f_out, int* info, int nthreads,
                       double* trace, int trace_cap) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; ++i) {
        const long long a = offsets[i];
        const int T = (int)(offsets[i + 1] - a);
        double ymax = -INFINITY;
        for (int q = 0; q < T; ++q) if (y[a + q] > ymax) ymax = y[a + q];
        po_prep p; double th[PO_MAXP];
        int st = po_prepare(ds + a, y + a, T, floor_, ymax * cap_multiplier, o, &p, th);
        int iters = 0, nev = 0; double f = NAN;
        double th0[PO_MAXP];
        memcpy(th0, th, sizeof th0);
        if (st == 0 && o->algorithm != 2) st = po_lbfgs(&p, o, th, &f, &iters, &nev, trace ? trace + (size_t)i * trace_cap * 4 : NULL, trace_cap);
        if ((st == -1 && o->algorithm == 0) || (st == 0 && o->algorithm == 2)) {
            /* fbprophet 0.5 fit(): except RuntimeError -> optimizing(init=stan_init, algorithm='Newton') */
            int it2 = 0, nev2 = 0;
            memcpy(th, th0, sizeof th0);
            st = po_newton(&p, o, th, &f, &it2, &nev2);
            iters += it2; nev += nev2;
        }
        if (st >= 0 || st == -1 || st == -2) {
            for (int q = 0; q < p.S + p.K + 3 && q < pstride; ++q) theta_out[(size_t)i * pstride + q] = th[q];
            info[i * 5 + 3] = p.S; info[i * 5 + 4] = p.K;
            po_free(&p);
        }
        info[i * 5 + 0] = st; info[i * 5 + 1] = iters; info[i * 5 + 2] = nev;
        f_out[i] = f;
    }
    return 0;
}

int po_fit_batch(const long long* ds, const double* y, const long long* offsets, int n, double floor_, double cap_multiplier,
                 const po_opts* o, double* theta_out, int pstride, double* f_out, int* info, int nthreads) {
    return po_fit_batch_trace(ds, y, offsets, n, floor_, cap_multiplier, o, theta_out, pstride, f_out, info, nthreads, NULL, 0);
}
