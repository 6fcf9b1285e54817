/* hs_profile.h -- arrival times for NON-constant rate profiles,
 * shared by the kernels, the oracle and the ctypes twins exactly like hs_sampler.h.
 *
 * Restates, operation for operation (all IEEE binary64, no contraction):
 *   Profile.get_rate            happysimulator/load/profile.py:37-110
 *                               (ConstantRateProfile, LinearRampProfile, SpikeProfile)
 *   ArrivalTimeProvider.next_arrival_time, general path
 *                               happysimulator/load/arrival_time_provider.py:84-144
 *   integrate_adaptive_simpson  happysimulator/numerics/integration.py:10-90
 *   brentq                      happysimulator/numerics/root_finding.py:27-152
 * The recursion of the adaptive Simpson rule is unrolled onto an explicit stack
 * that visits the intervals in the reference's order (left subtree, right
 * subtree, then left + right), so every partial sum is rounded identically.
 */
#ifndef HS_PROFILE_H
#define HS_PROFILE_H

#include "hs_sampler.h"
#include "../../include/hs_b200.h"      /* hs_profile_desc, HS_PROF_* */

/* a Source whose provider raised RuntimeError (rate zero indefinitely / no bracket / no
 * convergence, arrival_time_provider.py:123-144): it stops ticking and -- unlike a
 * "time travel" tick -- no SourceEvent object (hence no sort index) is created */
#define HS_T_EXHAUSTED (INT64_MAX - 1)

#if defined(__CUDACC__)
#define HS_PROF_FN __host__ __device__ __noinline__
#else
#define HS_PROF_FN static
#endif

HS_HD double hs_fabs(double x) { return HS_BITS2D(HS_D2BITS(x) & 0x7fffffffffffffffULL); }

/* rate_fn(t) = profile.get_rate(Instant.from_seconds(t)); get_rate reads time.to_seconds() */
HS_HD double hs_profile_rate(const hs_profile_desc *P, double t_seconds)
{
    const double t = hs_ns_to_seconds(hs_seconds_to_ns(t_seconds));
    if (P->kind == HS_PROF_LINEAR_RAMP) {           /* profile.py:65-74 */
        if (t <= 0.0) return P->p[1];
        if (t >= P->p[0]) return P->p[2];
        const double fraction = HS_DIV(t, P->p[0]);
        return HS_ADD(P->p[1], HS_MUL(fraction, HS_SUB(P->p[2], P->p[1])));
    }
    if (P->kind == HS_PROF_STEP) {                  /* a user-defined step function, e.g. examples/queuing/m_m_1_queue.py:137-169 */
        const double *tab = (const double *)(uintptr_t)HS_D2BITS(P->p[2]);
        const int n = (int)P->p[1];
        int lo = 0, hi = n;                         /* number of breakpoints <= t (they ascend) */
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (t < tab[mid]) hi = mid; else lo = mid + 1; }
        return tab[n + lo];
    }
    if (P->kind == HS_PROF_SPIKE) {                 /* profile.py:98-110 */
        if (t < P->p[2]) return P->p[0];
        if (t < HS_ADD(P->p[2], P->p[3])) return P->p[1];
        return P->p[0];
    }
    return P->p[0];
}

HS_HD double hs_simpson(double fa, double fm, double fb, double h)
{ return HS_MUL(HS_DIV(h, 3.0), HS_ADD(HS_ADD(fa, HS_MUL(4.0, fm)), fb)); }

typedef struct hs_simpson_frame {
    double a, b, fa, fb, s_whole, tol, m, fm, s_left, s_right, left_result;
    int32_t depth, state;                           /* state: 0 fresh, 1 left pending, 2 right pending */
} hs_simpson_frame;

#define HS_SIMPSON_MAX_DEPTH 50

/* integrate_adaptive_simpson(rate_fn, a, b, tol=1e-10)[0] for a <= b */
HS_PROF_FN double hs_integrate_rate(const hs_profile_desc *P, double a, double b)
{
    if (a == b) return 0.0;
    hs_simpson_frame st[HS_SIMPSON_MAX_DEPTH + 2];
    int sp = 0;
    {
        const double fa = hs_profile_rate(P, a), fb = hs_profile_rate(P, b);
        const double m = HS_DIV(HS_ADD(a, b), 2.0);
        const double fm = hs_profile_rate(P, m);
        const double h = HS_DIV(HS_SUB(b, a), 2.0);
        st[0].a = a; st[0].b = b; st[0].fa = fa; st[0].fb = fb;
        st[0].s_whole = hs_simpson(fa, fm, fb, h); st[0].tol = 1e-10; st[0].depth = 0; st[0].state = 0;
    }
    double ret = 0.0;
    int have_ret = 0;
    while (sp >= 0) {
        hs_simpson_frame *f = &st[sp];
        if (have_ret) {
            if (f->state == 1) {                     /* left subtree done: descend right */
                f->left_result = ret; f->state = 2; have_ret = 0;
                hs_simpson_frame *c = &st[++sp];
                c->a = f->m; c->b = f->b; c->fa = f->fm; c->fb = f->fb; c->s_whole = f->s_right;
                c->tol = HS_DIV(f->tol, 2.0); c->depth = f->depth + 1; c->state = 0;
            } else {                                 /* both done: left + right */
                ret = HS_ADD(f->left_result, ret);
                --sp;
            }
            continue;
        }
        /* state 0: evaluate this interval (integration.py:57-75) */
        const double m = HS_DIV(HS_ADD(f->a, f->b), 2.0);
        const double h = HS_DIV(HS_SUB(f->b, f->a), 2.0);
        const double fm = hs_profile_rate(P, m);
        const double lm = HS_DIV(HS_ADD(f->a, m), 2.0);
        const double rm = HS_DIV(HS_ADD(m, f->b), 2.0);
        const double flm = hs_profile_rate(P, lm);
        const double frm = hs_profile_rate(P, rm);
        const double h2 = HS_DIV(h, 2.0);
        const double s_left = hs_simpson(f->fa, flm, fm, h2);
        const double s_right = hs_simpson(fm, frm, f->fb, h2);
        const double s_combined = HS_ADD(s_left, s_right);
        const double err = HS_DIV(HS_SUB(s_combined, f->s_whole), 15.0);
        if (f->depth >= HS_SIMPSON_MAX_DEPTH || hs_fabs(err) < f->tol) {
            ret = HS_ADD(s_combined, err);           /* Richardson extrapolation */
            have_ret = 1;
            --sp;
            continue;
        }
        f->m = m; f->fm = fm; f->s_left = s_left; f->s_right = s_right; f->state = 1;
        hs_simpson_frame *c = &st[++sp];
        c->a = f->a; c->b = m; c->fa = f->fa; c->fb = fm; c->s_whole = s_left;
        c->tol = HS_DIV(f->tol, 2.0); c->depth = f->depth + 1; c->state = 0;
    }
    return ret;
}

/* objective_func(t) = area(t_start, t) - target_area (arrival_time_provider.py:88-92) */
HS_HD double hs_arrival_objective(const hs_profile_desc *P, double t_start, double t, double target)
{ return HS_SUB(hs_integrate_rate(P, t_start, t), target); }

/* brentq(objective, a, b) with the reference's defaults; returns 0 if not converged */
HS_PROF_FN int hs_brentq_arrival(const hs_profile_desc *P, double t_start, double target,
                                 double a, double b, double *root)
{
    const double xtol = 1e-12, rtol = 4.0 * 2.220446049250313e-16;
    double fa = hs_arrival_objective(P, t_start, a, target);
    double fb = hs_arrival_objective(P, t_start, b, target);
    /* fa * fb > 0 raises ValueError in the reference; the bracket search rules it out */
    if (hs_fabs(fa) < hs_fabs(fb)) { double t = a; a = b; b = t; t = fa; fa = fb; fb = t; }
    double c = a, fc = fa, d = HS_SUB(b, a), e = d;
    for (int it = 0; it < 100; ++it) {
        const double tol = HS_ADD(HS_MUL(HS_MUL(2.0, rtol), hs_fabs(b)), xtol);
        const double m = HS_DIV(HS_SUB(c, b), 2.0);
        if (hs_fabs(m) <= tol || fb == 0.0) { *root = b; return 1; }
        if (hs_fabs(e) >= tol && hs_fabs(fa) > hs_fabs(fb)) {
            const double s = HS_DIV(fb, fa);
            double p, q;
            if (a == c) {
                p = HS_MUL(HS_MUL(2.0, m), s);
                q = HS_SUB(1.0, s);
            } else {
                q = HS_DIV(fa, fc);
                const double r = HS_DIV(fb, fc);
                /* p = s * (2.0 * m * q * (q - r) - (b - a) * (r - 1.0)) */
                p = HS_MUL(s, HS_SUB(HS_MUL(HS_MUL(HS_MUL(2.0, m), q), HS_SUB(q, r)),
                                     HS_MUL(HS_SUB(b, a), HS_SUB(r, 1.0))));
                q = HS_MUL(HS_MUL(HS_SUB(q, 1.0), HS_SUB(r, 1.0)), HS_SUB(s, 1.0));
            }
            if (p > 0.0) q = -q; else p = -p;
            /* min(3.0 * m * q - abs(tol * q), abs(e * q)) -- Python's min keeps the first on ties */
            const double x1 = HS_SUB(HS_MUL(HS_MUL(3.0, m), q), hs_fabs(HS_MUL(tol, q)));
            const double x2 = hs_fabs(HS_MUL(e, q));
            const double mn = (x2 < x1) ? x2 : x1;
            if (HS_MUL(2.0, p) < mn) { e = d; d = HS_DIV(p, q); }
            else { d = m; e = m; }
        } else { d = m; e = m; }
        a = b; fa = fb;
        if (hs_fabs(d) > tol) b = HS_ADD(b, d);
        else if (m > 0.0) b = HS_ADD(b, tol);
        else b = HS_SUB(b, tol);
        fb = hs_arrival_objective(P, t_start, b, target);
        if (HS_MUL(fb, fc) > 0.0) { c = a; fc = fa; d = HS_SUB(b, a); e = d; }
        else if (hs_fabs(fc) < hs_fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    }
    *root = b;
    return 0;
}

/* ArrivalTimeProvider.next_arrival_time for a non-constant profile
 * (arrival_time_provider.py:84-144): returns the new current_time in ns, or
 * HS_T_EXHAUSTED where the reference raises RuntimeError. */
HS_PROF_FN int64_t hs_next_arrival_profile_ns(const hs_profile_desc *P, int64_t cur_ns, double target)
{
    const double t_start = hs_ns_to_seconds(cur_ns);
    const double current_rate = hs_profile_rate(P, t_start);
    double t_high;
    if (current_rate > 0.0) {
        double est = HS_MUL(HS_DIV(target, current_rate), 2.0);
        /* max(_MIN_INTER_ARRIVAL_S, min(est, _MAX_EXPLORATION_TIME_S)) with Python's tie rules */
        const double mn = (3600.0 < est) ? 3600.0 : est;
        est = (mn > 1e-9) ? mn : 1e-9;
        t_high = HS_ADD(t_start, est);
    } else {
        t_high = HS_ADD(t_start, 0.1);
    }
    const double t_low = t_start;
    int found = 0;
    for (int i = 0; i < 50; ++i) {
        const double val = hs_arrival_objective(P, t_start, t_high, target);
        if (val > 0.0) { found = 1; break; }
        const double span = HS_SUB(t_high, t_low);
        const double step = (span > 1e-6) ? span : 1e-6;        /* max(_MIN_STEP_SIZE, t_high - t_low) */
        t_high = HS_ADD(t_high, HS_MUL(step, 2.0));
    }
    if (!found) return HS_T_EXHAUSTED;
    double root;
    if (!hs_brentq_arrival(P, t_start, target, t_low, t_high, &root)) return HS_T_EXHAUSTED;
    return hs_seconds_to_ns(root);
}

#endif /* HS_PROFILE_H */
