/* tiktak_oracle.cpp -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for libtiktak_cuda.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * The reference (/root/reference, Fortran 90, ifort-only) cannot be compiled in this image (no
 * Fortran compiler; SHARE='DENYRW' is Intel-specific), so this is a line-faithful C++ restatement of
 * the single-instance (seqNo = 1) run, serial, values kept in memory.  Each function cites the
 * reference lines it follows.  PARITY PINNING: the reference ships no tests, golden vectors or sample
 * outputs, so the oracle is pinned by the known-answer vectors derived from the
 * reference's formulas (tests/test_oracle_kat.py) and by an independent numpy
 * restatement of the small pieces (tests/golden/make_golden.py).  With respect to outputs of the
 * reference binary itself: "parity unpinned".
 *
 * Build: make -C oracle   (g++ -O2 -ffp-contract=off: no FMA contraction, like the reference build)
 *
 * Floating-point conventions restated from the reference:
 *   - all stage arithmetic is unfused and evaluated left to right as written;
 *   - Fortran SUM / DOT_PRODUCT are taken in index order;
 *   - MINLOC / MAXLOC return the first occurrence;
 *   - REAL() is single precision (nrtype.f90:5): w11 and itratio go through float.
 * math_mode 0 uses libm cos (what the reference's intrinsic COS is); math_mode 1 uses the
 * deterministic tk_cos of include/tiktak_math.h so that GPU and CPU trajectories can be compared
 * bit for bit (the two modes agree to <= 2 ulp per cos; tests check objective values to 1e-12).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../include/tiktak_cuda.h" /* tiktak_config layout only */
#include "../include/tiktak_math.h" /* tk_cos for math_mode 1, tk_* bit casts */
#include "../include/tiktak_rng.h"  /* the documented substitute for RANDOM_NUMBER (lottery) */
#include "../tables/sobol_bf1111.inc"

namespace {

/* ------------------------------------------------------------------------------------------
 * Sobol generator: insobl (utilities.f90:883-1043) and i4_sobol (utilities.f90:1045-1116)
 * ------------------------------------------------------------------------------------------ */
struct Sobol {
    int s = 0, maxcol = 0, scount = 0;
    double recipd = 0.0;
    std::vector<int> lastq;            /* 1-based [1..s] */
    std::vector<std::vector<int>> v;   /* v[i][j], 1-based, i<=s, j<=maxcol */
    int insobl(int dim_num, int atmost) {
        const int maxdim = 1111, maxdeg = 13, maxbit = 30;
        s = dim_num;
        if (dim_num < 1 || maxdim < dim_num || atmost <= 0 || (1 << maxbit) <= atmost) return -1; /* :921-955 */
        int i = atmost; /* :959-964 number of bits in atmost */
        maxcol = 0;
        do { maxcol++; i = i / 2; } while (i > 0);
        v.assign(s + 1, std::vector<int>(maxcol + 1, 0));
        for (int j = 1; j <= maxcol; j++) v[1][j] = 1; /* :977 */
        for (i = 2; i <= s; i++) {                      /* :981-1024 */
            int j = tk_sobol_poly[i - 2], m = 0;
            for (;;) { j = j / 2; if (j > 0) m++; else break; }
            bool includ[maxdeg + 1];
            j = tk_sobol_poly[i - 2];
            for (int k = m; k >= 1; k--) { includ[k] = (j % 2 == 1); j = j / 2; }
            for (int jj = 1; jj <= m && jj <= maxcol; jj++) v[i][jj] = tk_sobol_vinit[i - 2][jj - 1];
            for (int jj = m + 1; jj <= maxcol; jj++) {
                int newv = v[i][jj - m], l = 1;
                for (int k = 1; k <= m; k++) {
                    l = 2 * l;
                    if (includ[k]) newv = newv ^ (l * v[i][jj - k]);
                }
                v[i][jj] = newv;
            }
        }
        int l = 1; /* :1027-1031 multiply columns by powers of 2 */
        for (int j = maxcol - 1; j >= 1; j--) {
            l = 2 * l;
            for (int ii = 1; ii <= s; ii++) v[ii][j] = v[ii][j] * l;
        }
        recipd = 0.5 / (double)l; /* :1035 */
        scount = 0;
        lastq.assign(s + 1, 0);
        return 0;
    }
    /* quirk = 1: literal `i = floor(real(i/2))` with single-precision REAL (utilities.f90:1090);
     * quirk = 0: exact integer halving (identical for Scount <= 33,554,433). */
    int next(double* quasi, int quirk) {
        int i = scount, l = 1;
        while (i % 2 == 1) {
            l = l + 1;
            if (quirk) i = (int)floorf((float)(i / 2));
            else i = i / 2;
        }
        if (maxcol < l) return -1; /* :1095-1102 */
        for (int d = 1; d <= s; d++) {
            lastq[d] = lastq[d] ^ v[d][l];
            quasi[d - 1] = (double)lastq[d] * recipd;
        }
        scount++;
        return 0;
    }
};

/* ------------------------------------------------------------------------------------------
 * indexx (utilities.f90:1334-1424): NR index quicksort, NN=15, NSTACK=50.  1-based inside.
 * ------------------------------------------------------------------------------------------ */
int indexx(int n, const double* arr0, int* index0) {
    const int NN = 15, NSTACK = 50;
    const double* arr = arr0 - 1; /* arr[1..n] */
    int* index = index0 - 1;
    int istack[NSTACK + 1];
    for (int i = 1; i <= n; i++) index[i] = i;
    int jstack = 0, l = 1, r = n;
    auto icomp_xchg = [&](int& i, int& j) { if (arr[j] < arr[i]) { int s = i; i = j; j = s; } };
    for (;;) {
        if (r - l < NN) {
            for (int j = l + 1; j <= r; j++) {
                int indext = index[j];
                double a = arr[indext];
                int i;
                for (i = j - 1; i >= l; i--) {
                    if (arr[index[i]] <= a) break;
                    index[i + 1] = index[i];
                }
                index[i + 1] = indext;
            }
            if (jstack == 0) return 0;
            r = istack[jstack];
            l = istack[jstack - 1];
            jstack -= 2;
        } else {
            int k = (l + r) / 2;
            std::swap(index[k], index[l + 1]);
            icomp_xchg(index[l], index[r]);
            icomp_xchg(index[l + 1], index[r]);
            icomp_xchg(index[l], index[l + 1]);
            int i = l + 1, j = r;
            int indext = index[l + 1];
            double a = arr[indext];
            for (;;) {
                do { i++; } while (!(arr[index[i]] >= a));
                do { j--; } while (!(arr[index[j]] <= a));
                if (j < i) break;
                std::swap(index[i], index[j]);
            }
            index[l + 1] = index[j];
            index[j] = indext;
            jstack += 2;
            if (jstack > NSTACK) return -1; /* 'indexx: NSTACK too small' -> stop 0 */
            if (r - i + 1 >= j - l) {
                istack[jstack] = r;
                istack[jstack - 1] = i;
                r = j - 1;
            } else {
                istack[jstack] = j - 1;
                istack[jstack - 1] = l;
                l = i;
            }
        }
    }
}

/* value as re-read from a '200f40.20' record (independent of tk_f4020_roundtrip: real text) */
double text4020(double x) {
    if (!std::isfinite(x)) return x;
    char buf[512];
    snprintf(buf, sizeof buf, "%40.20f", x);
    return strtod(buf, nullptr);
}

#include "smm_oracle.inc"

struct Row { /* one searchResults.dat record (TiktakGlobalSearch.f90:594) */
    int seq, k, imp;
    double fval;
    std::vector<double> x;
};

struct Oracle {
    tiktak_config c;
    std::vector<double> range, init, bound; /* column-major copies */
    int math_mode = 0;
    double penalty0 = 0, penaltyc = 0, max_penalty = 0; /* objective.f90:35-36 */
    const double myeps = 2.220446049250313e-16;         /* EPSILON(1.0_DP), objective.f90:24 */
    std::vector<double> rsq;                            /* 1/sqrt(i) for the Griewank plugin */
    long nfev = 0;
    /* pre-test state */
    int64_t kcut = 0, nlegit = 0;
    std::vector<double> sob_f;   /* fval of points 1..kcut */
    std::vector<double> sob_x;   /* (kcut, nx) row-major */
    /* chooseSobol output */
    std::vector<double> x_starts; /* (K, nx+1) row-major: fval, x */
    std::vector<int> start_idx;
    int n_start_rows = 0;         /* rows actually filled (reference leaves the rest uninitialised) */
    bool tie_straddle = false;
    /* local stage */
    std::vector<Row> rows;
    long best_tie_events = 0;
    SmmOracle smm;

    double COS(double a) const { return math_mode ? tk_cos(a) : cos(a); }
    const double* lo_r() const { return range.data(); }
    const double* hi_r() const { return range.data() + c.nx; }
    const double* bl() const { return bound.data(); }
    const double* bu() const { return bound.data() + c.nx; }

    /* getPenalty (objective.f90:51-74) */
    double getPenalty(const double* theta) const {
        double penalty = 0.0;
        for (int i = 0; i < c.nx; i++) {
            double a = std::max(0.0, bl()[i] - theta[i]) / std::max(0.1, fabs(bl()[i]));
            double b = std::max(0.0, theta[i] - bu()[i]) / std::max(0.1, fabs(bu()[i]));
            double temp = penaltyc * (a * a) + penaltyc * (b * b);
            penalty = penalty + temp;
        }
        if (penalty > myeps) return std::min(max_penalty, penalty0 + penalty);
        return 0.0;
    }

    /* the scalar v_err that the template replicates over the nm moments (objective.f90:96-117),
     * for each synthetic objective.  Plain unfused arithmetic, index order. */
    double model_value(const double* x) const {
        const int n = c.nx;
        switch (c.objective_id) {
            case TIKTAK_OBJ_TEMPLATE: { /* objective.f90:98-109, literally */
                double val1 = (x[0] * x[0]) / 200.0;
                for (int ii = 2; ii <= n; ii++) val1 = val1 + (x[ii - 1] * x[ii - 1]) / 200.0;
                double val2 = COS(x[0]);
                for (int ii = 2; ii <= n; ii++) val2 = val2 * COS(x[ii - 1] / sqrt(ii + 0.0));
                double v = val1 - val2 + 1.0;
                return v + 1.0;
            }
            case TIKTAK_OBJ_GRIEWANK: {
                double s = 0.0, p = 1.0;
                for (int i = 0; i < n; i++) {
                    s = s + (x[i] * x[i]) * 2.5e-4;
                    p = p * COS(x[i] * rsq[i]);
                }
                double v = s - p + 1.0;
                return v + 1.0;
            }
            case TIKTAK_OBJ_RASTRIGIN: {
                double s = 10.0 * (double)n;
                for (int i = 0; i < n; i++) s = s + (x[i] * x[i] - 10.0 * COS(6.283185307179586 * x[i]));
                return s + 1.0;
            }
            case TIKTAK_OBJ_ROSENBROCK: {
                double s = 0.0;
                for (int i = 0; i + 1 < n; i++) {
                    double t = x[i + 1] - x[i] * x[i];
                    double u = 1.0 - x[i];
                    s = s + (100.0 * (t * t) + u * u);
                }
                return s + 1.0;
            }
        }
        return NAN;
    }

    /* dfovec (objective.f90:76-123) */
    void dfovec(const double* x, double* F) const {
        double v_err = getPenalty(x);
        if (v_err > myeps) {
            double f = v_err / sqrt((double)c.nm);
            for (int m = 0; m < c.nm; m++) F[m] = f;
        } else if (c.objective_id == TIKTAK_OBJ_SMM) {
            smm.residuals(x, F);
        } else if (c.objective_id == 100) { /* TEST ONLY (not in the library): Rosenbrock as residuals, nm = 2 (nx - 1) [+ 1: a constant
                                              * residual 0.5, so that F_min = 0.25 and the F <= 1e-12 exit (:831) cannot fire] */
            for (int i = 0; i + 1 < c.nx; i++) { F[2 * i] = 10.0 * (x[i + 1] - x[i] * x[i]); F[2 * i + 1] = 1.0 - x[i]; }
            for (int m = 2 * (c.nx - 1); m < c.nm; m++) F[m] = 0.5;
        } else if (c.objective_id == 101) { /* TEST ONLY: linear residuals A x - b with a fixed dense A (nm rows) */
            for (int m = 0; m < c.nm; m++) {
                double r = -(0.5 + 0.25 * cos(1.0 + m));
                for (int j = 0; j < c.nx; j++) r = r + sin(1.3 * (m + 1) + 0.7 * (j + 1) + 0.1 * (m + 1) * (j + 1)) * x[j];
                F[m] = r;
            }
        } else {
            double v = model_value(x);
            for (int m = 0; m < c.nm; m++) F[m] = v;
        }
    }
    /* objFun (objective.f90:125-142) */
    double objFun(const double* theta) {
        nfev++;
        std::vector<double> F(c.nm);
        dfovec(theta, F.data());
        double dot = 0.0;
        for (int m = 0; m < c.nm; m++) dot = dot + F[m] * F[m];
        return dot + getPenalty(theta);
    }

    /* setupSobol (TiktakGlobalSearch.f90:825-838) + solveAtSobolPoints (:405-471) */
    int pretest() {
        Sobol sb;
        if (c.qr_ndraw <= 0) { kcut = 0; nlegit = 0; return 0; }
        if (sb.insobl(c.nx, c.qr_ndraw)) return -1;
        std::vector<double> q(c.nx), x(c.nx);
        sob_f.clear(); sob_x.clear();
        int64_t whichPoint = 1, legit = 0; /* counters lastSobol.dat / legitSobol.dat */
        while (whichPoint <= c.qr_ndraw && legit < c.legitimate) { /* :429 */
            if (sb.next(q.data(), c.sobol_quirk)) return -2;
            for (int i = 0; i < c.nx; i++) x[i] = lo_r()[i] + q[i] * (hi_r()[i] - lo_r()[i]); /* :834 */
            double fval = objFun(x.data());                                                   /* :437 */
            if (fval < c.fvalmax) legit++;                                                     /* :438-439 */
            sob_f.push_back(fval);
            sob_x.insert(sob_x.end(), x.begin(), x.end());
            whichPoint++;
        }
        kcut = whichPoint - 1;
        nlegit = legit;
        return 0;
    }

    /* chooseSobol (TiktakGlobalSearch.f90:840-899).  mode 0: the reference's indexx (unstable
     * among equal keys); mode 1: ascending fval then ascending Sobol index (what the GPU does). */
    int choose(int mode) {
        const int K = c.maxpoints_plus1, nx = c.nx;
        x_starts.assign((size_t)K * (nx + 1), NAN);
        start_idx.assign(K, 0);
        x_starts[0] = objFun(init.data()); /* :853 */
        for (int i = 0; i < nx; i++) x_starts[1 + i] = init[i];
        n_start_rows = 1;
        tie_straddle = false;
        if (c.qr_ndraw > 0 && kcut > 0 && K > 1) {
            const int numrows = (int)kcut;
            /* the rows as re-read from sobolFnVal.dat (:860) */
            std::vector<double> fv(numrows);
            for (int r = 0; r < numrows; r++) fv[r] = c.text_roundtrip ? text4020(sob_f[r]) : sob_f[r];
            std::vector<int> idx(numrows);
            if (mode == 0) {
                if (indexx(numrows, fv.data(), idx.data())) return -1; /* :882 */
            } else {
                for (int r = 0; r < numrows; r++) idx[r] = r + 1;
                std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return fv[a - 1] < fv[b - 1]; });
            }
            auto put = [&](int j /*1-based row*/, int src /*1-based sobol row*/) {
                double* row = &x_starts[(size_t)(j - 1) * (nx + 1)];
                row[0] = fv[src - 1];
                for (int i = 0; i < nx; i++) {
                    double xv = sob_x[(size_t)(src - 1) * nx + i];
                    row[1 + i] = c.text_roundtrip ? text4020(xv) : xv;
                }
                start_idx[j - 1] = src;
            };
            put(2, idx[0]); /* :884 */
            int j = 3;
            for (int i = 2; i <= numrows; i++) { /* :886-892 */
                if (j > K) break;                /* (the reference tests after the copy; same effect) */
                if (idx[i - 1] != idx[i - 2]) { put(j, idx[i - 1]); j++; }
            }
            n_start_rows = std::min(j - 1, K);
            /* does a run of equal keys cross the selection boundary? then the SET depends on tie order */
            const int M = K - 1;
            if (numrows > M && M >= 1) {
                std::vector<double> s(fv);
                std::nth_element(s.begin(), s.begin() + (M - 1), s.end());
                double kth = s[M - 1];
                long below = 0, equal = 0;
                for (double f : fv) { if (f < kth) below++; else if (f == kth) equal++; }
                if (below + equal > M && equal > 1) tie_straddle = true;
            }
        }
        return 0;
    }

    /* getSortedResults (:804-823) + getBestLine (:781-802) on a given set of rows */
    int best_row(const std::vector<Row>& rs, int upto) {
        if (upto <= 0) return -1;
        std::vector<double> fv(upto);
        for (int i = 0; i < upto; i++) fv[i] = rs[i].fval;
        std::vector<int> idx(upto);
        indexx(upto, fv.data(), idx.data());
        int b = idx[0] - 1;
        for (int i = 0; i < upto; i++)
            if (i != b && rs[i].fval == rs[b].fval) { best_tie_events++; break; }
        return b;
    }

    /* simplex_coordinates2 (simplex.f90:346-438); x is (n, n+1) column-major */
    static void simplex_coordinates2(int n, std::vector<double>& x) {
        x.assign((size_t)n * (n + 1), 0.0);
        auto X = [&](int i, int j) -> double& { return x[(size_t)(j - 1) * n + (i - 1)]; };
        for (int j = 1; j <= n; j++) X(j, j) = 1.0;
        double a = (1.0 - sqrt(1.0 + (double)n)) / (double)n;
        for (int i = 1; i <= n; i++) X(i, n + 1) = a;
        std::vector<double> cvec(n + 1);
        for (int i = 1; i <= n; i++) {
            double sum = 0.0;
            for (int j = 1; j <= n + 1; j++) sum = sum + X(i, j);
            cvec[i] = sum / (double)(n + 1);
        }
        for (int j = 1; j <= n + 1; j++)
            for (int i = 1; i <= n; i++) X(i, j) = X(i, j) - cvec[i];
        double ss = 0.0;
        for (int i = 1; i <= n; i++) ss = ss + X(i, 1) * X(i, 1);
        double s = sqrt(ss);
        for (int j = 1; j <= n + 1; j++)
            for (int i = 1; i <= n; i++) X(i, j) = X(i, j) / s;
    }

    int amoeba_trips = 0, amoeba_shrinks = 0; /* loop trips and shrink steps of the last amoeba call */
    /* amoeba + amotry (minimize.f90:7-135).  p is (ndim+1, ndim) ROW-major here (p[i][j]). */
    void amoeba(std::vector<double>& p, std::vector<double>& y, double ftol, int& iter, int ITMAX) {
        const double TINY = (double)1.0e-10f; /* REAL(dp), PARAMETER :: TINY=1.0e-10 -- SP literal (:28) */
        const int ndim = c.nx, np = ndim + 1;
        auto P = [&](int i, int j) -> double& { return p[(size_t)i * ndim + j]; };
        std::vector<double> psum(ndim), ptry(ndim);
        auto iminloc = [&]() { int b = 0; for (int i = 1; i < np; i++) if (y[i] < y[b]) b = i; return b; };
        auto imaxloc = [&]() { int b = 0; for (int i = 1; i < np; i++) if (y[i] > y[b]) b = i; return b; };
        auto sum_psum = [&]() {
            for (int j = 0; j < ndim; j++) { double s = 0.0; for (int i = 0; i < np; i++) s = s + P(i, j); psum[j] = s; }
        };
        int ihi = 0;
        auto amotry = [&](double fac) {
            double fac1 = (1.0 - fac) / ndim;
            double fac2 = fac1 - fac;
            for (int j = 0; j < ndim; j++) ptry[j] = psum[j] * fac1 - P(ihi, j) * fac2;
            double ytry = objFun(ptry.data());
            if (ytry < y[ihi]) {
                y[ihi] = ytry;
                for (int j = 0; j < ndim; j++) { psum[j] = psum[j] - P(ihi, j) + ptry[j]; P(ihi, j) = ptry[j]; }
            }
            return ytry;
        };
        iter = 0;
        amoeba_trips = 0; amoeba_shrinks = 0;
        sum_psum();
        for (;;) {
            int ilo = iminloc();
            ihi = imaxloc();
            double ytmp = y[ihi];
            y[ihi] = y[ilo];
            int inhi = imaxloc();
            y[ihi] = ytmp;
            double rtol = 2.0 * fabs(y[ihi] - y[ilo]) / (fabs(y[ihi]) + fabs(y[ilo]) + TINY);
            if (rtol < ftol || iter >= ITMAX) { /* :85-96 */
                std::swap(y[0], y[ilo]);
                for (int j = 0; j < ndim; j++) std::swap(P(0, j), P(ilo, j));
                return;
            }
            amoeba_trips++; /* one pass of the loop: the virtual clock of the asynchronous-instances run (local_async) */
            double ytry = amotry(-1.0);
            iter++;
            if (ytry <= y[ilo]) {
                ytry = amotry(2.0);
                iter++;
            } else if (ytry >= y[inhi]) {
                double ysave = y[ihi];
                ytry = amotry(0.5);
                iter++;
                if (ytry >= ysave) {
                    for (int i = 0; i < np; i++)
                        for (int j = 0; j < ndim; j++) P(i, j) = 0.5 * (P(i, j) + P(ilo, j));
                    /* NB: row ilo is its own fixed point, so updating rows in order is the same as
                     * the array expression with spread() (:107) */
                    for (int i = 0; i < np; i++)
                        if (i != ilo) y[i] = objFun(&P(i, 0));
                    iter += ndim;
                    amoeba_shrinks++;
                    sum_psum();
                }
            }
        }
    }

    /* per-coordinate width that scales the initial simplex: p_range (:623) or, in shrink mode 1, the
     * README.md:36 rule (bounds of the best local minima so far) -- see DESIGN.md "nm_shrink_mode". */
    void simplex_width(int k, const std::vector<Row>& snapshot, int count, std::vector<double>& w) {
        const int nx = c.nx;
        w.resize(nx);
        for (int i = 0; i < nx; i++) w[i] = hi_r()[i] - lo_r()[i];
        if (c.nm_shrink_mode != 1) return;
        float itratio = (float)k / (float)c.maxpoints_plus1;
        if (!(itratio > 0.5f)) return;
        /* rows that are local-search outcomes (k >= 1) */
        std::vector<int> cand;
        for (int i = 0; i < count; i++) if (snapshot[i].k >= 1) cand.push_back(i);
        if ((int)cand.size() < 2) return;
        std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
            if (snapshot[a].fval != snapshot[b].fval) return snapshot[a].fval < snapshot[b].fval;
            return snapshot[a].k < snapshot[b].k; });
        int T = std::min<int>((int)cand.size(), 10);
        for (int i = 0; i < nx; i++) {
            double mn = snapshot[cand[0]].x[i], mx = mn;
            for (int t = 1; t < T; t++) { double v = snapshot[cand[t]].x[i]; mn = std::min(mn, v); mx = std::max(mx, v); }
            double ww = mx - mn;
            if (ww > 0.0) w[i] = ww;
        }
    }

    /* runAmoeba (TiktakGlobalSearch.f90:602-633) */
    void runAmoeba(std::vector<double>& pt, const std::vector<double>& width, int& iters) {
        const int nx = c.nx, K = c.maxpoints_plus1;
        std::vector<double> x;
        simplex_coordinates2(nx, x);
        std::vector<double> xT((size_t)(nx + 1) * nx), fnVals(nx + 1), temp(nx);
        const double scale = pow((double)K, 1.0 / nx); /* DBLE(p_maxpoints)**(1.0_dp/p_nx) */
        for (int ia = 1; ia <= nx + 1; ia++) {
            for (int i = 0; i < nx; i++) {
                double t = pt[i] + x[(size_t)(ia - 1) * nx + i] * width[i] / scale; /* :623 */
                t = std::max(std::min(t, bu()[i]), bl()[i]);                          /* :624 */
                temp[i] = t;
            }
            fnVals[ia - 1] = objFun(temp.data());
            for (int i = 0; i < nx; i++) xT[(size_t)(ia - 1) * nx + i] = temp[i];
        }
        amoeba(xT, fnVals, c.nm_ftol, iters, c.nm_itmax); /* the 5th positional is `iter`; ITMAX stays 5000 */
        int ia = 0;
        for (int i = 1; i <= nx; i++) if (fnVals[i] < fnVals[ia]) ia = i; /* minloc, first occurrence */
        for (int i = 0; i < nx; i++) pt[i] = xT[(size_t)ia * nx + i];
    }

    void write_row(int seq, int k, int imp, double fval, const double* x) {
        Row r;
        r.seq = seq; r.k = k; r.imp = imp;
        r.fval = c.text_roundtrip ? text4020(fval) : fval;
        r.x.resize(c.nx);
        for (int i = 0; i < c.nx; i++) r.x[i] = c.text_roundtrip ? text4020(x[i]) : x[i];
        rows.push_back(r);
    }

    /* EST_dfpmin with dograd and lnsrch (minimize.f90:2572-2816): BFGS with central-difference gradients.
     * Sums (DOT_PRODUCT, MATMUL) in index order.  Where lnsrch returns early on slope >= 0 (:2760-2763) the
     * reference leaves x and f unset; here they stay at xold / fold. */
    int dfpmin(std::vector<double>& p, double& fret, int itmax, double gtol) {
        const int nn = c.nx;
        const double stpmx = 100.0, eps = 2.220446049250313e-16, tolx = 4.0 * eps, xinc = 1.0e-5;
        auto func = [&](const std::vector<double>& x) { return objFun(x.data()); };
        auto dograd = [&](const std::vector<double>& x) {
            std::vector<double> dh(nn), yy(nn), grad(nn);
            const double tolera = 10.0 * xinc;
            for (int i = 0; i < nn; i++) dh[i] = (fabs(x[i]) >= tolera) ? x[i] * xinc : xinc;
            for (int i = 0; i < nn; i++) {
                for (int j = 0; j < nn; j++) yy[j] = x[j] - (i == j ? dh[i] : 0.0);
                double tempf1 = func(yy);
                for (int j = 0; j < nn; j++) yy[j] = x[j] + (i == j ? dh[i] : 0.0);
                double tempf2 = func(yy);
                grad[i] = (tempf2 - tempf1) / (2.0 * dh[i]);
            }
            return grad;
        };
        auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) { double s = 0.0; for (int i = 0; i < nn; i++) s = s + a[i] * b[i]; return s; };
        auto lnsrch = [&](const std::vector<double>& xold, double fold, const std::vector<double>& g, std::vector<double>& pdir,
                          std::vector<double>& x, double& f, double stpmax) {
            const double alf = 1.0e-4, tolx2 = eps;
            double a, alam, alam2 = 0, alamin, b, disc, f2 = 0, pabs, rhs1, rhs2, slope, tmplam;
            pabs = sqrt(dot(pdir, pdir));
            if (pabs > stpmax) for (int i = 0; i < nn; i++) pdir[i] = pdir[i] * stpmax / pabs;
            slope = dot(g, pdir);
            x = xold; f = fold;
            if (slope >= 0.0) return;
            double mx = 0.0;
            for (int i = 0; i < nn; i++) mx = std::max(mx, fabs(pdir[i]) / std::max(fabs(xold[i]), 1.0));
            alamin = tolx2 / mx;
            alam = 1.0;
            for (;;) {
                for (int i = 0; i < nn; i++) x[i] = xold[i] + alam * pdir[i];
                f = func(x);
                if (alam < alamin) { x = xold; return; }
                else if (f <= fold + alf * alam * slope) return;
                else {
                    if (alam == 1.0) tmplam = -slope / (2.0 * (f - fold - slope));
                    else {
                        rhs1 = f - fold - alam * slope;
                        rhs2 = f2 - fold - alam2 * slope;
                        a = (rhs1 / (alam * alam) - rhs2 / (alam2 * alam2)) / (alam - alam2);
                        b = (-alam2 * rhs1 / (alam * alam) + alam * rhs2 / (alam2 * alam2)) / (alam - alam2);
                        if (a == 0.0) tmplam = -slope / (2.0 * b);
                        else {
                            disc = b * b - 3.0 * a * slope;
                            if (disc < 0.0) tmplam = 0.5 * alam;
                            else if (b <= 0.0) tmplam = (-b + sqrt(disc)) / (3.0 * a);
                            else tmplam = -slope / (b + sqrt(disc));
                        }
                        if (tmplam > 0.5 * alam) tmplam = 0.5 * alam;
                    }
                }
                alam2 = alam;
                f2 = f;
                alam = std::max(tmplam, 0.1 * alam);
            }
        };
        double fp = func(p);
        std::vector<double> g = dograd(p), dg(nn), hdg(nn), pnew(nn), xi(nn), hessin((size_t)nn * nn, 0.0);
        for (int i = 0; i < nn; i++) hessin[(size_t)i * nn + i] = 1.0;
        for (int i = 0; i < nn; i++) xi[i] = -g[i];
        const double stpmax = stpmx * std::max(sqrt(dot(p, p)), (double)nn);
        fret = fp;
        for (int its = 1; its <= itmax; its++) {
            lnsrch(p, fp, g, xi, pnew, fret, stpmax);
            fp = fret;
            for (int i = 0; i < nn; i++) xi[i] = pnew[i] - p[i];
            p = pnew;
            double t = 0.0;
            for (int i = 0; i < nn; i++) t = std::max(t, fabs(xi[i]) / std::max(fabs(p[i]), 1.0));
            if (t < tolx) return its;
            dg = g;
            g = dograd(p);
            const double den = std::max(fret, 1.0);
            t = 0.0;
            for (int i = 0; i < nn; i++) t = std::max(t, fabs(g[i]) * std::max(fabs(p[i]), 1.0) / den);
            if (t < gtol) return its;
            for (int i = 0; i < nn; i++) dg[i] = g[i] - dg[i];
            for (int i = 0; i < nn; i++) { double sm = 0.0; for (int j = 0; j < nn; j++) sm = sm + hessin[(size_t)i * nn + j] * dg[j]; hdg[i] = sm; }
            double fac = dot(dg, xi), fae = dot(dg, hdg), sumdg = dot(dg, dg), sumxi = dot(xi, xi);
            if (fac > sqrt(eps * sumdg * sumxi)) {
                fac = 1.0 / fac;
                const double fad = 1.0 / fae;
                for (int i = 0; i < nn; i++) dg[i] = fac * xi[i] - fad * hdg[i];
                for (int i = 0; i < nn; i++)
                    for (int j = 0; j < nn; j++)
                        hessin[(size_t)i * nn + j] = hessin[(size_t)i * nn + j] + fac * (xi[i] * xi[j]) - fad * (hdg[i] * hdg[j]) + fae * (dg[i] * dg[j]);
            }
            for (int i = 0; i < nn; i++) { double sm = 0.0; for (int j = 0; j < nn; j++) sm = sm + hessin[(size_t)i * nn + j] * g[j]; xi[i] = -sm; }
        }
        return itmax;
    }

    /* getLotteryPoint (TiktakGlobalSearch.f90:722-766): among the best min(count, lotteryPoints) result rows, pick
     * rank i with probability proportional to 1/i.  Deviations from the literal code, all documented in DESIGN.md:
     * rows are ordered by (fval, k) instead of indexx's unstable order; RANDOM_NUMBER is replaced by
     * tk_uniform01(seed 123456, restart k); `y = sortedPoints(i,4:)` (a size mismatch) is read as
     * (i,5:), the parameters; if rounding leaves u >= the last cumulative probability the last row is taken. */
    int lottery_row(const std::vector<Row>& snap, int count, int k) {
        std::vector<int> ord(count);
        for (int i = 0; i < count; i++) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
            if (snap[a].fval != snap[b].fval) return snap[a].fval < snap[b].fval;
            return snap[a].k < snap[b].k; });
        const int numPoints = std::min(count, (int)c.lottery_points);
        const long long sumVal = ((long long)numPoints * (numPoints + 1)) / 2;
        std::vector<double> prob(numPoints);
        for (int i = 1; i <= numPoints; i++) prob[i - 1] = (double)sumVal / (double)i;
        double sum2 = 0.0;
        for (int i = 0; i < numPoints; i++) sum2 = sum2 + prob[i]; /* the remaining entries of probOfPoint are 0 */
        for (int i = 0; i < numPoints; i++) prob[i] = prob[i] / sum2;
        for (int i = 1; i < numPoints; i++) prob[i] = prob[i - 1] + prob[i];
        const double u = tk_uniform01(TK_LOTTERY_SEED, TK_LOTTERY_STREAM, (uint64_t)k);
        int sel = numPoints - 1;
        for (int i = 0; i < numPoints; i++) if (u < prob[i]) { sel = i; break; }
        lottery_pick.push_back(snap[ord[sel]].k);
        return ord[sel];
    }
    std::vector<int> lottery_pick; /* lotPoint of every blended restart, for tests */

    /* LocalMinimizations (:473-544) + getModifiedParam (:683-720) + completeSearch (:546-600),
     * restarts 1..K, `round_size` restarts share the best-so-far of the round's start. */
    int local(int alg, int round_size, double* starts_out, double* results_out, int32_t* evals_out, int k_first = 1);
    int async_clock = TIKTAK_ASYNC_CLOCK_TRIPS; /* Nelder-Mead: the tick of local_async's virtual clock (tiktak_cuda_local_async_clock) */
    int local_async(int alg, int instances, double* starts_out, double* results_out, int32_t* evals_out, int32_t* fin_out, int32_t* inst_out);
    int missing_search(int alg, int k, double* start_out, double* row_out, int32_t* evals_out);
    int dfnls(std::vector<double>& x, double rhobeg, double rhoend, int maxfun, int& nf, int& nrescue);
    int dfnls_force_first = 0, dfnls_force_every = 0; /* test hook (orc_set_dfnls_force_rescue) */
    long nrescue_total = 0, rescue_evals_total = 0;    /* RESCUE_H calls of all dfnls runs of this oracle, evaluations made inside them */
};

/* SolveMissingSearch (TiktakGlobalSearch.f90:1104-1146) for one missing restart k: getModifiedParam is called with whichPoint =
 * p_maxpoints (:1121), so the blend weight is that of the LAST restart number; completeSearch(alg, k) follows (:1126) */
int Oracle::missing_search(int alg, int k, double* start_out, double* row_out, int32_t* evals_out) {
    const int K = c.maxpoints_plus1, nx = c.nx;
    if (k < 1 || k > K || rows.empty()) return -1;
    const int count = (int)rows.size();
    const std::vector<Row> snapshot(rows.begin(), rows.end());
    std::vector<double> evalParam(nx);
    const double* evalPoint = &x_starts[(size_t)(k - 1) * (nx + 1) + 1];
    if (count <= 1) {
        for (int i = 0; i < nx; i++) evalParam[i] = evalPoint[i];
    } else {
        const std::vector<double>& base = (c.search_type == 1) ? snapshot[lottery_row(snapshot, count, k)].x : snapshot[best_row(snapshot, count)].x;
        float wf = sqrtf((float)K / (float)K);
        wf = std::min(std::max(0.10f, wf), 0.95f);
        const double w11 = (double)wf;
        for (int i = 0; i < nx; i++) evalParam[i] = w11 * base[i] + (1.0 - w11) * evalPoint[i];
    }
    if (start_out) for (int i = 0; i < nx; i++) start_out[i] = evalParam[i];
    const long nfev0 = nfev;
    if (alg == TIKTAK_ALG_AMOEBA) {
        std::vector<double> width;
        simplex_width(k, snapshot, count, width);
        int iters = 0;
        runAmoeba(evalParam, width, iters);
    } else if (alg == TIKTAK_ALG_DFNLS) {
        double itratio = (double)((float)k / (float)K);
        double mn = bu()[0] - bl()[0];
        for (int i = 1; i < nx; i++) mn = std::min(mn, bu()[i] - bl()[i]);
        double rhobeg, rhoend;
        if (itratio > 0.5) { rhobeg = (mn / 2.5) / (4 * itratio); rhoend = 1.0e-3 / (4 * itratio); }
        else { rhobeg = mn / 2.50; rhoend = 1.0e-3; }
        int nf = 0, nresc = 0;
        dfnls(evalParam, rhobeg, rhoend, c.maxeval, nf, nresc);
    } else return -2;
    const double fn_val = objFun(evalParam.data());
    if (evals_out) *evals_out = (int32_t)(nfev - nfev0);
    const int b = best_row(rows, (int)rows.size());
    int imp = rows[b].imp;
    if (fn_val < rows[b].fval) imp++;
    write_row(1, k, imp, fn_val, evalParam.data());
    if (row_out) {
        const Row& r = rows.back();
        row_out[0] = r.seq; row_out[1] = r.k; row_out[2] = r.imp; row_out[3] = r.fval;
        for (int i = 0; i < nx; i++) row_out[4 + i] = r.x[i];
    }
    return 0;
}

/* k_first > 1: continue an earlier run (option 3, genericParams.f90:133-251): `rows` already holds the k = 0 record and
 * the finished restarts 1..k_first-1 (orc_set_rows), as re-read from searchResults.dat */
int Oracle::local(int alg, int round_size, double* starts_out, double* results_out, int32_t* evals_out, int k_first) {
    const int K = c.maxpoints_plus1, nx = c.nx;
    if ((int)x_starts.size() != K * (nx + 1)) return -1;
    if (round_size < 1) round_size = 1;
    const int seqNo = 1;
    lottery_pick.clear();
    best_tie_events = 0;
    if (k_first <= 1) {
        k_first = 1;
        rows.clear();
        write_row(seqNo, 0, 0, x_starts[0], &x_starts[1]); /* :496-501, whichPoint == 1 */
    } else if ((int)rows.size() != k_first) return -3;
    for (int k0 = k_first; k0 <= K; k0 += round_size) {
        const int k1 = std::min(K, k0 + round_size - 1);
        const int count = (int)rows.size(); /* rows visible to this round */
        const std::vector<Row> snapshot(rows.begin(), rows.end());
        int bsnap = best_row(snapshot, count);
        for (int k = k0; k <= k1; k++) {
            std::vector<double> evalParam(nx);
            const double* evalPoint = &x_starts[(size_t)(k - 1) * (nx + 1) + 1];
            if (count <= 1) { /* :701-705 */
                for (int i = 0; i < nx; i++) evalParam[i] = evalPoint[i];
            } else {
                const std::vector<double>& base = (c.search_type == 1) ? snapshot[lottery_row(snapshot, count, k)].x : snapshot[bsnap].x;
                float wf = sqrtf((float)k / (float)K); /* :718, single precision */
                wf = std::min(std::max(0.10f, wf), 0.95f);
                double w11 = (double)wf;
                for (int i = 0; i < nx; i++) evalParam[i] = w11 * base[i] + (1.0 - w11) * evalPoint[i]; /* :719 */
            }
            if (starts_out) for (int i = 0; i < nx; i++) starts_out[(size_t)i * K + (k - 1)] = evalParam[i];
            long nfev0 = nfev;
            if (alg == TIKTAK_ALG_AMOEBA) {
                std::vector<double> width;
                simplex_width(k, snapshot, count, width);
                int iters = 0;
                runAmoeba(evalParam, width, iters);
            } else if (alg == TIKTAK_ALG_DFNLS) {
                /* :564 itratio is REAL(DP) but the quotient REAL(k)/REAL(K) is single precision */
                double itratio = (double)((float)k / (float)K);
                double mn = bu()[0] - bl()[0];
                for (int i = 1; i < nx; i++) mn = std::min(mn, bu()[i] - bl()[i]);
                double rhobeg, rhoend;
                if (itratio > 0.5) { /* :565-571 */
                    rhobeg = (mn / 2.5) / (4 * itratio);
                    rhoend = 1.0e-3 / (4 * itratio);
                } else {
                    rhobeg = mn / 2.50;
                    rhoend = 1.0e-3;
                }
                int nf = 0, nresc = 0;
                dfnls(evalParam, rhobeg, rhoend, c.maxeval, nf, nresc);
            } else if (alg == TIKTAK_ALG_DFPMIN) { /* CASE (2), :575-576 */
                double fret = 0;
                dfpmin(evalParam, fret, 1000, 1.0e-6);
            } else return -2;
            double fn_val = objFun(evalParam.data()); /* :582 */
            if (evals_out) evals_out[k - 1] = (int32_t)(nfev - nfev0);
            int b = best_row(rows, (int)rows.size()); /* getBestLine :585 */
            int imp = rows[b].imp;
            if (fn_val < rows[b].fval) imp++;
            write_row(seqNo, k, imp, fn_val, evalParam.data()); /* :594 */
        }
    }
    if (results_out) {
        for (int k = 1; k <= K; k++) {
            const Row& r = rows[k]; /* rows[0] is the k=0 record */
            results_out[(size_t)0 * K + (k - 1)] = r.seq;
            results_out[(size_t)1 * K + (k - 1)] = r.k;
            results_out[(size_t)2 * K + (k - 1)] = r.imp;
            results_out[(size_t)3 * K + (k - 1)] = r.fval;
            for (int i = 0; i < nx; i++) results_out[(size_t)(4 + i) * K + (k - 1)] = r.x[i];
        }
    }
    return 0;
}

/* The reference's multi-process mode as a sequential event simulation (restated on the device by nm_quad_async_kernel):
 * P instances; an instance that finishes a restart appends its row and takes the next restart number; the new start is blended
 * toward the best row of the file as it is then (completeSearch :546-600, getModifiedParam :683-720).  "Then" is a virtual time:
 * a restart costs cb + c0 + trips + shrinks * c0 ticks (cb = TIKTAK_ASYNC_PICKUP_TICKS: taking the next restart; trips = passes of
 * amoeba's loop; c0 = ceil((nx+1)/4): what the initial simplex and a shrink step cost four warps).  Restart numbers go out in the order of the finish times, ties by instance number; a start sees the
 * rows with finish time <= its start time; the best row is the lowest value, ties by the order of the file (finish time, instance, k).
 * Nelder-Mead, search_type 0, nm_shrink_mode 0. */
int Oracle::local_async(int alg, int instances, double* starts_out, double* results_out, int32_t* evals_out, int32_t* fin_out, int32_t* inst_out) {
    const int K = c.maxpoints_plus1, nx = c.nx;
    if ((int)x_starts.size() != K * (nx + 1)) return -1;
    if (c.search_type != 0 || (alg == TIKTAK_ALG_AMOEBA && c.nm_shrink_mode != 0)) return -2;
    if (alg != TIKTAK_ALG_AMOEBA && alg != TIKTAK_ALG_DFNLS) return -2;
    const int P = std::max(1, std::min(instances, K));
    const int c0 = (nx + 1 + 3) / 4;
    const int seqNo = 1;
    rows.clear();
    write_row(seqNo, 0, 0, x_starts[0], &x_starts[1]); /* :496-501: the k = 0 record */
    std::vector<int> fin(K + 1, -1), inst(K + 1, -1);
    fin[0] = 0;
    std::vector<Row> byk(K + 1);
    byk[0] = rows[0];
    std::vector<long> free_at(P, 0); /* when instance i is free again */
    auto before = [&](int a, int b) { /* row a before row b: lower value, then the order of the file */
        if (byk[a].fval != byk[b].fval) return byk[a].fval < byk[b].fval;
        if (fin[a] != fin[b]) return fin[a] < fin[b];
        if (inst[a] != inst[b]) return inst[a] < inst[b];
        return a < b;
    };
    for (int k = 1; k <= K; k++) {
        int i = 0;
        for (int q = 1; q < P; q++) if (free_at[q] < free_at[i]) i = q; /* earliest free instance, lowest number among equals */
        const int T = (int)free_at[i];
        int count = 0, b = -1;
        for (int j = 0; j < k; j++) {
            if (fin[j] < 0 || fin[j] > T) continue;
            if (!(byk[j].fval == byk[j].fval)) continue;
            count++;
            if (b < 0 || before(j, b)) b = j;
        }
        std::vector<double> evalParam(nx);
        const double* evalPoint = &x_starts[(size_t)(k - 1) * (nx + 1) + 1];
        if (count <= 1) {
            for (int d = 0; d < nx; d++) evalParam[d] = evalPoint[d];
        } else {
            float wf = sqrtf((float)k / (float)K);
            wf = std::min(std::max(0.10f, wf), 0.95f);
            const double w11 = (double)wf;
            for (int d = 0; d < nx; d++) evalParam[d] = w11 * byk[b].x[d] + (1.0 - w11) * evalPoint[d];
        }
        if (starts_out) for (int d = 0; d < nx; d++) starts_out[(size_t)d * K + (k - 1)] = evalParam[d];
        const long nfev0 = nfev;
        int steps;
        if (alg == TIKTAK_ALG_AMOEBA) {
            std::vector<double> width;
            simplex_width(k, rows, 1, width);
            int iters = 0;
            runAmoeba(evalParam, width, iters);
            steps = TIKTAK_ASYNC_PICKUP_TICKS + c0 + amoeba_trips + amoeba_shrinks * c0; /* TIKTAK_ASYNC_CLOCK_TRIPS; EVALS: below */
        } else { /* DFNLS: the tick is one objective evaluation (for the SMM objective: one pass over the panel) */
            const double itratio = (double)((float)k / (float)K);
            double mn = bu()[0] - bl()[0];
            for (int d = 1; d < nx; d++) mn = std::min(mn, bu()[d] - bl()[d]);
            double rhobeg, rhoend;
            if (itratio > 0.5) { rhobeg = (mn / 2.5) / (4 * itratio); rhoend = 1.0e-3 / (4 * itratio); }
            else { rhobeg = mn / 2.50; rhoend = 1.0e-3; }
            int nf = 0, nresc = 0;
            dfnls(evalParam, rhobeg, rhoend, c.maxeval, nf, nresc);
            steps = 0;
        }
        const double fn_val = objFun(evalParam.data()); /* :582 */
        if (alg == TIKTAK_ALG_DFNLS) steps = TIKTAK_ASYNC_PICKUP_TICKS + (int)(nfev - nfev0);
        else if (async_clock == TIKTAK_ASYNC_CLOCK_EVALS) steps = TIKTAK_ASYNC_PICKUP_TICKS_NM_EVALS + (int)(nfev - nfev0);
        if (evals_out) evals_out[k - 1] = (int32_t)(nfev - nfev0);
        Row r;
        r.seq = seqNo; r.k = k; r.imp = 0;
        r.fval = c.text_roundtrip ? text4020(fn_val) : fn_val;
        r.x.resize(nx);
        for (int d = 0; d < nx; d++) r.x[d] = c.text_roundtrip ? text4020(evalParam[d]) : evalParam[d];
        byk[k] = r;
        fin[k] = T + steps;
        inst[k] = i;
        free_at[i] = fin[k];
    }
    /* the file: rows in the order of the finish events; the #improvements column as completeSearch :585-589 computes it there */
    std::vector<int> order(K);
    for (int k = 1; k <= K; k++) order[k - 1] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return fin[a] != fin[b] ? fin[a] < fin[b] : (inst[a] != inst[b] ? inst[a] < inst[b] : a < b); });
    double best_f = rows[0].fval;
    int best_imp = 0;
    for (int k : order) {
        Row& r = byk[k];
        if (!(r.fval == r.fval)) continue;
        int imp = best_imp;
        if (r.fval < best_f) { imp = best_imp + 1; best_f = r.fval; best_imp = imp; }
        r.imp = imp;
        rows.push_back(r);
    }
    if (results_out) {
        for (int k = 1; k <= K; k++) {
            const Row& r = byk[k];
            results_out[(size_t)0 * K + (k - 1)] = r.seq;
            results_out[(size_t)1 * K + (k - 1)] = r.k;
            results_out[(size_t)2 * K + (k - 1)] = r.imp;
            results_out[(size_t)3 * K + (k - 1)] = r.fval;
            for (int d = 0; d < nx; d++) results_out[(size_t)(4 + d) * K + (k - 1)] = r.x[d];
        }
    }
    if (fin_out) for (int k = 1; k <= K; k++) fin_out[k - 1] = fin[k];
    if (inst_out) for (int k = 1; k <= K; k++) inst_out[k - 1] = inst[k];
    return 0;
}

#include "dfnls_oracle.inc"

} /* namespace */

/* ------------------------------------------------------------------------------------------
 * C API (ctypes).  Mirrors include/tiktak_cuda.h entry for entry where that makes sense.
 * ------------------------------------------------------------------------------------------ */
extern "C" {

void* orc_create(const tiktak_config* cfg, int math_mode) {
    if (!cfg || cfg->nx < 1 || cfg->nx > 1111 || cfg->nm < 1) return nullptr;
    Oracle* o = new Oracle();
    o->c = *cfg;
    const int nx = cfg->nx;
    o->range.assign(cfg->range, cfg->range + 2 * nx);
    o->init.assign(cfg->init, cfg->init + nx);
    o->bound.assign(cfg->bound, cfg->bound + 2 * nx);
    o->c.range = o->range.data(); o->c.init = o->init.data(); o->c.bound = o->bound.data();
    o->math_mode = math_mode;
    o->penalty0 = cfg->fvalmax; /* objective.f90:35-36 */
    o->penaltyc = 10 * cfg->fvalmax;
    o->max_penalty = 1000.0 * o->penalty0;
    o->rsq.resize(nx);
    for (int i = 0; i < nx; i++) o->rsq[i] = 1.0 / sqrt((double)(i + 1));
    if (cfg->objective_id == TIKTAK_OBJ_SMM) {
        const tiktak_smm_params* sp = (const tiktak_smm_params*)cfg->user;
        if (!sp || nx != 7 || cfg->nm != 30 * sp->T) { delete o; return nullptr; }
        o->smm.init(sp);
    }
    return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }

/* (count, nx) column-major, points first..first+count-1 (generated from point 1: the reference's
 * generator is sequential) */
int orc_sobol_points(void* h, int64_t first, int64_t count, double* out, int quirk, int atmost) {
    Oracle* o = (Oracle*)h;
    Sobol sb;
    if (sb.insobl(o->c.nx, atmost > 0 ? atmost : (int)std::max<int64_t>(first + count - 1, 1))) return -1;
    std::vector<double> q(o->c.nx);
    for (int64_t n = 1; n < first + count; n++) {
        if (sb.next(q.data(), quirk)) return -2;
        if (n >= first)
            for (int i = 0; i < o->c.nx; i++)
                out[(size_t)i * count + (n - first)] = o->lo_r()[i] + q[i] * (o->hi_r()[i] - o->lo_r()[i]);
    }
    return 0;
}
/* raw unit-cube points + the l sequence, for the KATs */
int orc_sobol_unit(int nx, int atmost, int64_t count, double* out, int quirk) {
    Sobol sb;
    if (sb.insobl(nx, atmost)) return -1;
    std::vector<double> q(nx);
    for (int64_t n = 0; n < count; n++) {
        if (sb.next(q.data(), quirk)) return -2;
        for (int i = 0; i < nx; i++) out[(size_t)n * nx + i] = q[i];
    }
    return 0;
}
/* rightmost-zero column l for a given Scount under both rules (finding 2 of SURVEY.md) */
int orc_sobol_l(int scount, int quirk) {
    int i = scount, l = 1;
    while (i % 2 == 1) { l++; if (quirk) i = (int)floorf((float)(i / 2)); else i = i / 2; }
    return l;
}
int orc_sobol_maxcol(int nx, int atmost) { Sobol sb; if (sb.insobl(nx, atmost)) return -1; return sb.maxcol; }
/* direction numbers v(i, j) scaled to 32 bits (v * 2^(32-maxcol)), row-major (nx, 32): lets the
 * tests compare the library's table without going through points */
int orc_sobol_dirnums(int nx, int atmost, uint32_t* out) {
    Sobol sb;
    if (sb.insobl(nx, atmost)) return -1;
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= 32; j++)
            out[(size_t)(i - 1) * 32 + (j - 1)] = j <= sb.maxcol ? ((uint32_t)sb.v[i][j]) << (32 - sb.maxcol) : 0u;
    return 0;
}

int orc_objfun(void* h, const double* x, int64_t n, double* fval) { /* x (n, nx) column-major */
    Oracle* o = (Oracle*)h;
    std::vector<double> t(o->c.nx);
    for (int64_t r = 0; r < n; r++) {
        for (int i = 0; i < o->c.nx; i++) t[i] = x[(size_t)i * n + r];
        fval[r] = o->objFun(t.data());
    }
    return 0;
}
int orc_dfovec(void* h, const double* x, double* F) { ((Oracle*)h)->dfovec(x, F); return 0; }
double orc_penalty(void* h, const double* x) { return ((Oracle*)h)->getPenalty(x); }

int orc_pretest(void* h, int64_t* n_eval, int64_t* n_legit) {
    Oracle* o = (Oracle*)h;
    int rc = o->pretest();
    if (n_eval) *n_eval = o->kcut;
    if (n_legit) *n_legit = o->nlegit;
    return rc;
}
int orc_pretest_values(void* h, int64_t first, int64_t count, double* fval) {
    Oracle* o = (Oracle*)h;
    for (int64_t i = 0; i < count; i++) {
        int64_t n = first + i;
        fval[i] = (n >= 1 && n <= o->kcut) ? o->sob_f[n - 1] : NAN;
    }
    return 0;
}
/* x_starts (K, nx+1) column-major; idx[K]; returns rows filled; *tie = 1 when equal keys straddle the cut */
int orc_choose(void* h, int mode, double* x_starts, int32_t* idx, int* tie) {
    Oracle* o = (Oracle*)h;
    if (o->choose(mode)) return -1;
    const int K = o->c.maxpoints_plus1, nx = o->c.nx;
    for (int k = 0; k < K; k++) {
        for (int j = 0; j <= nx; j++) x_starts[(size_t)j * K + k] = o->x_starts[(size_t)k * (nx + 1) + j];
        if (idx) idx[k] = o->start_idx[k];
    }
    if (tie) *tie = o->tie_straddle ? 1 : 0;
    return o->n_start_rows;
}
/* overwrite x_starts (e.g. with the GPU's, to test the local stage from identical starts) */
int orc_set_starts(void* h, const double* x_starts) {
    Oracle* o = (Oracle*)h;
    const int K = o->c.maxpoints_plus1, nx = o->c.nx;
    o->x_starts.resize((size_t)K * (nx + 1));
    for (int k = 0; k < K; k++)
        for (int j = 0; j <= nx; j++) o->x_starts[(size_t)k * (nx + 1) + j] = x_starts[(size_t)j * K + k];
    o->n_start_rows = K;
    return 0;
}
/* starts (K, nx), results (K, nx+4) column-major, evals[K] */
int orc_local(void* h, int alg, int round_size, double* starts, double* results, int32_t* evals) {
    return ((Oracle*)h)->local(alg, round_size, starts, results, evals);
}
/* resume: rows_in = the n finished restarts k = 1..n, (n, nx+4) column-major = seq, k, #improvements, fval, x as
 * re-read from searchResults.dat; the k = 0 record is rebuilt from x_starts; then restarts n+1..K are run */
int orc_local_resume(void* h, int alg, int round_size, const double* rows_in, int n, double* starts, double* results, int32_t* evals) {
    Oracle* o = (Oracle*)h;
    const int nx = o->c.nx;
    o->rows.clear();
    o->write_row(1, 0, 0, o->x_starts[0], &o->x_starts[1]);
    for (int r = 0; r < n; r++) {
        Row q;
        q.seq = (int)rows_in[(size_t)0 * n + r]; q.k = (int)rows_in[(size_t)1 * n + r]; q.imp = (int)rows_in[(size_t)2 * n + r];
        q.fval = rows_in[(size_t)3 * n + r];
        q.x.resize(nx);
        for (int i = 0; i < nx; i++) q.x[i] = rows_in[(size_t)(4 + i) * n + r];
        if (q.k != r + 1) return -4;
        o->rows.push_back(q);
    }
    return o->local(alg, round_size, starts, results, evals, n + 1);
}
int orc_best(void* h, double* fbest, double* xbest, int32_t* k) {
    Oracle* o = (Oracle*)h;
    if (o->rows.empty()) return -1;
    int b = o->best_row(o->rows, (int)o->rows.size());
    if (fbest) *fbest = o->rows[b].fval;
    if (xbest) for (int i = 0; i < o->c.nx; i++) xbest[i] = o->rows[b].x[i];
    if (k) *k = o->rows[b].k;
    return 0;
}
/* lastSearch (TiktakGlobalSearch.f90:635-681): EST_dfpmin from the best row, itmax 1000, gtol 1e-6 */
int orc_last_search(void* h, double* fval, double* x, int32_t* iters) {
    Oracle* o = (Oracle*)h;
    if (o->rows.empty()) return -1;
    int b = o->best_row(o->rows, (int)o->rows.size());
    std::vector<double> p = o->rows[b].x;
    double fret = o->rows[b].fval;
    int it = o->dfpmin(p, fret, 1000, 1.0e-6);
    if (fval) *fval = fret;
    if (x) for (int i = 0; i < o->c.nx; i++) x[i] = p[i];
    if (iters) *iters = it;
    return 0;
}
long orc_nfev(void* h) { return ((Oracle*)h)->nfev; }
long orc_best_tie_events(void* h) { return ((Oracle*)h)->best_tie_events; }

/* one Nelder-Mead restart from `start` exactly as runAmoeba does it for restart k (width = p_range) */
void orc_set_async_clock(void* h, int clock) { ((Oracle*)h)->async_clock = clock; }
int orc_local_async(void* h, int alg, int instances, double* starts, double* results, int32_t* evals, int32_t* fin, int32_t* inst) {
    return ((Oracle*)h)->local_async(alg, instances, starts, results, evals, fin, inst);
}
int orc_run_amoeba(void* h, double* pt, int32_t* iters) {
    Oracle* o = (Oracle*)h;
    std::vector<double> p(pt, pt + o->c.nx), w(o->c.nx);
    for (int i = 0; i < o->c.nx; i++) w[i] = o->hi_r()[i] - o->lo_r()[i];
    int it = 0;
    o->runAmoeba(p, w, it);
    for (int i = 0; i < o->c.nx; i++) pt[i] = p[i];
    if (iters) *iters = it;
    return 0;
}
int orc_simplex_coordinates2(int n, double* x) { /* (n, n+1) column-major */
    std::vector<double> v;
    Oracle::simplex_coordinates2(n, v);
    memcpy(x, v.data(), v.size() * sizeof(double));
    return 0;
}
int orc_indexx(int n, const double* arr, int32_t* index) {
    std::vector<int> idx(n);
    int rc = indexx(n, arr, idx.data());
    for (int i = 0; i < n; i++) index[i] = idx[i];
    return rc;
}
double orc_text4020(double x) { return text4020(x); }

/* solveAtSobolPoints WITH the reference's file protocol (TiktakGlobalSearch.f90:426-463, stateControl.f90:83-148,
 * 187-201): per point one getNextNumber on lastSobol.dat (read cycle + replace cycle), one writeToLog append, one
 * getNextNumber / getNumber on legitSobol.dat and one append to sobolFnVal.dat -- five to six open/close cycles
 * around an objective evaluation.  Only used to time what a user of the reference's process-per-core scheme pays
 * (bench.py cpu_baseline "faithful"); `dir` must exist and be private to the caller.  Returns points evaluated. */
static long file_get_number(const std::string& f) { long y = 0; FILE* h = fopen(f.c_str(), "r"); if (h) { if (fscanf(h, "%ld", &y) != 1) y = 0; fclose(h); } return y; }
static long file_get_next_number(const std::string& f) {
    const long y = file_get_number(f) + 1;
    FILE* h = fopen(f.c_str(), "w");
    if (h) { fprintf(h, " %11ld\n", y); fclose(h); }
    return y;
}
long orc_pretest_faithful(void* hh, const char* dir, long npoints) {
    Oracle* o = (Oracle*)hh;
    const int nx = o->c.nx;
    const std::string d(dir);
    const std::string lastSobol = d + "/lastSobol.dat", legit = d + "/legitSobol.dat", fnval = d + "/sobolFnVal.dat", logfile = d + "/logFile.txt";
    for (const std::string& f : {lastSobol, legit}) { FILE* h = fopen(f.c_str(), "w"); if (!h) return -1; fprintf(h, " %11d\n", 0); fclose(h); }
    for (const std::string& f : {fnval, logfile}) { FILE* h = fopen(f.c_str(), "w"); if (!h) return -1; fclose(h); }
    Sobol sb;
    if (sb.insobl(nx, o->c.qr_ndraw)) return -1;
    std::vector<double> q(nx), x(nx);
    long whichPoint = file_get_next_number(lastSobol), legitSobol = file_get_number(legit), done = 0;
    while (whichPoint <= npoints && legitSobol < o->c.legitimate) {
        { FILE* h = fopen(logfile.c_str(), "a"); if (h) { fprintf(h, " Sequence# %5d is solving sobol point %6ld\n", 1, whichPoint); fclose(h); } }
        if (sb.next(q.data(), 0)) return -2;
        for (int i = 0; i < nx; i++) x[i] = o->lo_r()[i] + q[i] * (o->hi_r()[i] - o->lo_r()[i]);
        const double fval = o->objFun(x.data());
        legitSobol = (fval < o->c.fvalmax) ? file_get_next_number(legit) : file_get_number(legit);
        { FILE* h = fopen(fnval.c_str(), "a"); if (h) { fprintf(h, "%8ld%40.20f", whichPoint, fval); for (int i = 0; i < nx; i++) fprintf(h, "%40.20f", x[i]); fputc('\n', h); fclose(h); } }
        done++;
        whichPoint = file_get_next_number(lastSobol);
    }
    return done;
}

/* the lottery's random numbers (include/tiktak_rng.h) */
void orc_philox4x32_10(uint32_t* c, uint32_t k0, uint32_t k1) { tk_philox4x32_10(c, k0, k1); }
double orc_uniform01(uint32_t seed, uint32_t stream, uint64_t counter) { return tk_uniform01(seed, stream, counter); }
float orc_w11(int k, int K) { float wf = sqrtf((float)k / (float)K); return std::min(std::max(0.10f, wf), 0.95f); }
/* test hook: every later DFNLS run of this oracle treats the denominator tests (minimize.f90:746, :782) as failed when NF reaches
 * `first` (and every `every` evaluations after it), i.e. calls RESCUE_H there; first <= 0 switches it off.  Returns the number of
 * RESCUE_H calls so far. */
long orc_set_dfnls_force_rescue(void* h, int first, int every) {
    Oracle* o = (Oracle*)h;
    o->dfnls_force_first = first; o->dfnls_force_every = every;
    return o->nrescue_total;
}
long orc_dfnls_rescue_evals(void* h) { return ((Oracle*)h)->rescue_evals_total; }
/* one DFNLS run (bobyqa_h) from x with explicit radii */
int orc_dfnls(void* h, double* x, double rhobeg, double rhoend, int maxfun, int32_t* nf, int32_t* nrescue) {
    Oracle* o = (Oracle*)h;
    std::vector<double> p(x, x + o->c.nx);
    int n1 = 0, n2 = 0;
    int rc = o->dfnls(p, rhobeg, rhoend, maxfun, n1, n2);
    for (int i = 0; i < o->c.nx; i++) x[i] = p[i];
    if (nf) *nf = n1;
    if (nrescue) *nrescue = n2;
    return rc;
}
/* the rows of an interrupted run (k = 0 record first, then the finished restarts in file order; (n, nx+4) column-major) become
 * the oracle's searchResults; orc_missing_search then runs one missing restart (states 8-9) and returns its start, its
 * searchResults row (seq, k, #improvements, fval, x) and its evaluation count */
int orc_set_rows(void* h, const double* rows_in, int n) {
    Oracle* o = (Oracle*)h;
    const int nx = o->c.nx;
    o->rows.clear();
    for (int r = 0; r < n; r++) {
        std::vector<double> x(nx);
        for (int i = 0; i < nx; i++) x[i] = rows_in[(size_t)(4 + i) * n + r];
        o->write_row((int)rows_in[(size_t)0 * n + r], (int)rows_in[(size_t)1 * n + r], (int)rows_in[(size_t)2 * n + r], rows_in[(size_t)3 * n + r], x.data());
    }
    return 0;
}
int orc_missing_search(void* h, int alg, int k, double* start_out, double* row_out, int32_t* evals_out) {
    return ((Oracle*)h)->missing_search(alg, k, start_out, row_out, evals_out);
}
/* objFun at the Sobol points first..first+count-1 without generating the points before them: the generator's state
 * after n calls is lastq = XOR of the columns v(:,j) over the set bits j of the Gray code n ^ (n >> 1) (SURVEY 8a3;
 * checked against the sequential generator by tests/test_oracle_kat.py), Scount = n.  Lets P processes share a
 * large pre-test (C3: 1e8 points) by index range.  Returns the number of values below fvalmax, or < 0. */
long orc_objfun_range(void* h, int64_t first, int64_t count, double* fval) {
    Oracle* o = (Oracle*)h;
    Sobol sb;
    if (first < 1 || count < 0 || sb.insobl(o->c.nx, o->c.qr_ndraw)) return -1;
    const int64_t n0 = first - 1;
    const int64_t g = n0 ^ (n0 >> 1);
    for (int j = 1; j <= sb.maxcol; j++)
        if ((g >> (j - 1)) & 1)
            for (int d = 1; d <= sb.s; d++) sb.lastq[d] ^= sb.v[d][j];
    sb.scount = (int)n0;
    std::vector<double> q(o->c.nx), x(o->c.nx);
    long legit = 0;
    for (int64_t r = 0; r < count; r++) {
        if (sb.next(q.data(), 0)) return -2;
        for (int i = 0; i < o->c.nx; i++) x[i] = o->lo_r()[i] + q[i] * (o->hi_r()[i] - o->lo_r()[i]); /* :834 */
        fval[r] = o->objFun(x.data());
        if (fval[r] < o->c.fvalmax) legit++;
    }
    return legit;
}
/* the Sobol points themselves for such a range: (count, nx) column-major, same seek */
int orc_points_range(void* h, int64_t first, int64_t count, double* out) {
    Oracle* o = (Oracle*)h;
    Sobol sb;
    if (first < 1 || count < 0 || sb.insobl(o->c.nx, o->c.qr_ndraw)) return -1;
    const int64_t n0 = first - 1;
    const int64_t g = n0 ^ (n0 >> 1);
    for (int j = 1; j <= sb.maxcol; j++)
        if ((g >> (j - 1)) & 1)
            for (int d = 1; d <= sb.s; d++) sb.lastq[d] ^= sb.v[d][j];
    sb.scount = (int)n0;
    std::vector<double> q(o->c.nx);
    for (int64_t r = 0; r < count; r++) {
        if (sb.next(q.data(), 0)) return -2;
        for (int i = 0; i < o->c.nx; i++) out[(size_t)i * count + r] = o->lo_r()[i] + q[i] * (o->hi_r()[i] - o->lo_r()[i]);
    }
    return 0;
}
/* completeSearch (TiktakGlobalSearch.f90:546-600) for restart k from a GIVEN start (evalParam): the local search with the
 * radii / simplex of restart k, then the re-evaluation :582.  x: in = start, out = result; returns objective evaluations. */
int orc_complete_search(void* h, int alg, int k, double* x, double* fn_val) {
    Oracle* o = (Oracle*)h;
    const int nx = o->c.nx, K = o->c.maxpoints_plus1;
    std::vector<double> p(x, x + nx);
    const long nfev0 = o->nfev;
    if (alg == TIKTAK_ALG_AMOEBA) {
        std::vector<double> w(nx);
        for (int i = 0; i < nx; i++) w[i] = o->hi_r()[i] - o->lo_r()[i];
        int it = 0;
        o->runAmoeba(p, w, it);
    } else if (alg == TIKTAK_ALG_DFNLS) {
        double itratio = (double)((float)k / (float)K); /* :564 */
        double mn = o->bu()[0] - o->bl()[0];
        for (int i = 1; i < nx; i++) mn = std::min(mn, o->bu()[i] - o->bl()[i]);
        double rhobeg, rhoend;
        if (itratio > 0.5) { rhobeg = (mn / 2.5) / (4 * itratio); rhoend = 1.0e-3 / (4 * itratio); } /* :565-571 */
        else { rhobeg = mn / 2.50; rhoend = 1.0e-3; }
        int nf = 0, nresc = 0;
        o->dfnls(p, rhobeg, rhoend, o->c.maxeval, nf, nresc);
    } else return -2;
    const double f = o->objFun(p.data()); /* :582 */
    for (int i = 0; i < nx; i++) x[i] = p[i];
    if (fn_val) *fn_val = f;
    return (int)(o->nfev - nfev0);
}
/* one DFNLS run that records the model state at every entry of label 60 (minimize.f90:525) for the invariant checks of
 * tests/test_oracle_pins.py.  Each snapshot = [KOPT, NF, XBASE(n), XOPT(n), XPT(npt,n), FVAL_V(mv,npt), GQV(mv,n), HQV(mv,nh),
 * PQV(mv,npt), BMAT(ndim,n), ZMAT(npt,nptm)], arrays column-major as in the Fortran.  Returns the number of snapshots. */
int orc_dfnls_trace(void* h, double* x, double rhobeg, double rhoend, int maxfun, int max_snap, double* out, int32_t* nf) {
    Oracle* o = (Oracle*)h;
    Dfnls d;
    d.force_first = o->dfnls_force_first; d.force_every = o->dfnls_force_every > 0 ? o->dfnls_force_every : 1 << 30;
    d.dfovec = [o](const double* xx, double* F) { o->nfev++; o->dfovec(xx, F); };
    d.getPenalty = [o](const double* xx) { return o->getPenalty(xx); };
    int nsnap = 0;
    double* w = out;
    d.trace = [&](const Dfnls& s, int KOPT, int NF) {
        if (nsnap >= max_snap) return;
        *w++ = KOPT; *w++ = NF;
        auto put = [&](const std::vector<double>& v, size_t n) { for (size_t i = 0; i < n; i++) *w++ = v[i]; };
        put(s.XBASE, s.N); put(s.XOPT, s.N); put(s.XPT, (size_t)s.NPT * s.N); put(s.FVAL_V, (size_t)s.mv * s.NPT);
        put(s.GQV, (size_t)s.mv * s.N); put(s.HQV, (size_t)s.mv * s.NH); put(s.PQV, (size_t)s.mv * s.NPT);
        put(s.BMAT, (size_t)s.NDIM * s.N); put(s.ZMAT, (size_t)s.NPT * s.NPTM);
        nsnap++;
    };
    const int n = o->c.nx, npt = 2 * n + 1;
    std::vector<double> p(x, x + n);
    int n1 = 0;
    d.BOBYQA_H(n, npt, p.data(), o->bl(), o->bu(), rhobeg, rhoend, maxfun, o->c.nm, n1);
    for (int i = 0; i < n; i++) x[i] = p[i];
    if (nf) *nf = n1;
    o->nrescue_total += d.nrescue; o->rescue_evals_total += d.rescue_evals;
    if (getenv("TK_DFNLS_DEBUG")) fprintf(stderr, "dfnls exit_line %d nf %d\n", d.exit_line, n1);
    return nsnap;
}
/* runtime self-check that this build does not contract a*b+c (the reference build has no FMA) */
int orc_unfused_selfcheck(void) {
    volatile double a = 1.0 + ldexp(1.0, -30), b = 1.0 - ldexp(1.0, -30), cc = -1.0;
    double unf = a * b + cc;      /* a*b rounds to 1 - 2^-60 -> 1.0?  exact a*b = 1 - 2^-60; rounds to 1.0 */
    double fus = fma(a, b, cc);   /* exact: -2^-60 */
    return (unf == 0.0 && fus != 0.0) ? 0 : -1;
}

} /* extern "C" */
