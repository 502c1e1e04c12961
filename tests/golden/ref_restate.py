"""Independent pure-Python restatements of the reference's local searches, written from the Fortran
(/root/reference/minimize.f90, TiktakGlobalSearch.f90, objective.f90) and sharing no text with the C++ oracle
or the CUDA kernels.  TEST INFRASTRUCTURE: they pin the oracle (tests/test_oracle_pins.py) value by value.

Python floats are IEEE doubles and every operation below is rounded separately (no FMA), `math.cos` is the
same glibc libm the oracle's MATH_LIBM mode calls, so agreement is expected to the bit.

    amoeba / amotry       minimize.f90:7-135
    runAmoeba             TiktakGlobalSearch.f90:602-633
    simplex_coordinates2  simplex.f90:346-438
    EST_dfpmin, dograd, lnsrch   minimize.f90:2572-2816
    getPenalty, dfovec, objFun   objective.f90:51-142 (template + the three synthetic functions of BASELINE.json)
"""
import math

import numpy as np

EPS = float(np.finfo(np.float64).eps)
TINY = float(np.float32(1.0e-10))  # `TINY=1.0e-10` is a single-precision literal (minimize.f90:28)


# ------------------------------------------------------------------------------------------- objective
class Objective:
    """objFun of objective.f90 for one configuration; kind in {'template', 'griewank', 'rastrigin', 'rosenbrock'}"""

    def __init__(self, kind, nm, bl, bu, fvalmax):
        self.kind, self.nm, self.bl, self.bu = kind, int(nm), [float(v) for v in bl], [float(v) for v in bu]
        self.penalty0 = float(fvalmax)  # obj_initialize :35-36
        self.penaltyc = 10 * float(fvalmax)
        self.max_penalty = 1000.0 * self.penalty0
        self.nfev = 0

    def get_penalty(self, theta):  # :51-74
        pen = 0.0
        for i in range(len(theta)):
            a = max(0.0, self.bl[i] - theta[i]) / max(0.1, abs(self.bl[i]))
            b = max(0.0, theta[i] - self.bu[i]) / max(0.1, abs(self.bu[i]))
            temp = self.penaltyc * (a * a) + self.penaltyc * (b * b)
            pen = pen + temp
        if pen > EPS:
            return min(self.max_penalty, self.penalty0 + pen)
        return 0.0

    def v_err(self, x):
        n = len(x)
        if self.kind == "template":  # :98-112
            val1 = (x[0] * x[0]) / 200.0
            for ii in range(2, n + 1):
                val1 = val1 + (x[ii - 1] * x[ii - 1]) / 200.0
            val2 = math.cos(x[0])
            for ii in range(2, n + 1):
                val2 = val2 * math.cos(x[ii - 1] / math.sqrt(ii + 0.0))
            return (val1 - val2 + 1.0) + 1.0
        if self.kind == "griewank":
            s, p = 0.0, 1.0
            for i in range(n):
                s = s + (x[i] * x[i]) * 2.5e-4
                p = p * math.cos(x[i] * (1.0 / math.sqrt(float(i + 1))))
            return (s - p + 1.0) + 1.0
        if self.kind == "rastrigin":
            s = 10.0 * float(n)
            for i in range(n):
                s = s + (x[i] * x[i] - 10.0 * math.cos(6.283185307179586 * x[i]))
            return s + 1.0
        if self.kind == "rosenbrock":
            s = 0.0
            for i in range(n - 1):
                t = x[i + 1] - x[i] * x[i]
                u = 1.0 - x[i]
                s = s + (100.0 * (t * t) + u * u)
            return s + 1.0
        raise KeyError(self.kind)

    def dfovec(self, x):  # :76-123
        pen = self.get_penalty(x)
        if pen > EPS:
            return [pen / math.sqrt(float(self.nm))] * self.nm
        return [self.v_err(x)] * self.nm

    def __call__(self, theta):  # objFun :125-142
        self.nfev += 1
        F = self.dfovec(theta)
        dot = 0.0
        for v in F:
            dot = dot + v * v
        return dot + self.get_penalty(theta)


# ------------------------------------------------------------------------------------------- Nelder-Mead
def simplex_coordinates2(n):
    x = [[0.0] * (n + 1) for _ in range(n)]  # x[i][j], i coordinate, j vertex
    for j in range(n):
        x[j][j] = 1.0
    a = (1.0 - math.sqrt(1.0 + float(n))) / float(n)
    for i in range(n):
        x[i][n] = a
    c = []
    for i in range(n):
        s = 0.0
        for j in range(n + 1):
            s = s + x[i][j]
        c.append(s / float(n + 1))
    for j in range(n + 1):
        for i in range(n):
            x[i][j] = x[i][j] - c[i]
    s = 0.0
    for i in range(n):
        s = s + x[i][0] ** 2
    s = math.sqrt(s)
    for j in range(n + 1):
        for i in range(n):
            x[i][j] = x[i][j] / s
    return x


def _iminloc(y):
    b = 0
    for i in range(1, len(y)):
        if y[i] < y[b]:
            b = i
    return b


def _imaxloc(y):
    b = 0
    for i in range(1, len(y)):
        if y[i] > y[b]:
            b = i
    return b


def amoeba(p, y, ftol, func, itmax=5000, stats=None):
    """minimize.f90:40-134; p: list of ndim+1 vertices (lists), y their values; returns iter (function evaluations).
    stats (optional dict) counts the branches taken: reflect / expand / contract / shrink."""
    ndim = len(p[0])
    it = 0
    psum = []
    for j in range(ndim):
        s = 0.0
        for i in range(ndim + 1):
            s = s + p[i][j]
        psum.append(s)
    state = {"ihi": 0}

    def amotry(fac):
        ihi = state["ihi"]
        fac1 = (1.0 - fac) / ndim
        fac2 = fac1 - fac
        ptry = [psum[j] * fac1 - p[ihi][j] * fac2 for j in range(ndim)]
        ytry = func(ptry)
        if ytry < y[ihi]:
            y[ihi] = ytry
            for j in range(ndim):
                psum[j] = psum[j] - p[ihi][j] + ptry[j]
            p[ihi] = ptry
        return ytry

    while True:
        ilo = _iminloc(y)
        ihi = _imaxloc(y)
        ytmp = y[ihi]
        y[ihi] = y[ilo]
        inhi = _imaxloc(y)
        y[ihi] = ytmp
        state["ihi"] = ihi
        rtol = 2.0 * abs(y[ihi] - y[ilo]) / (abs(y[ihi]) + abs(y[ilo]) + TINY)
        if rtol < ftol or it >= itmax:
            y[0], y[ilo] = y[ilo], y[0]
            p[0], p[ilo] = p[ilo], p[0]
            return it
        ytry = amotry(-1.0)
        it += 1
        if stats is not None:
            stats["reflect"] = stats.get("reflect", 0) + 1
        if ytry <= y[ilo]:
            ytry = amotry(2.0)
            it += 1
            if stats is not None:
                stats["expand"] = stats.get("expand", 0) + 1
        elif ytry >= y[inhi]:
            ysave = y[ihi]
            ytry = amotry(0.5)
            it += 1
            if stats is not None:
                stats["contract"] = stats.get("contract", 0) + 1
            if ytry >= ysave:
                if stats is not None:
                    stats["shrink"] = stats.get("shrink", 0) + 1
                plo = list(p[ilo])
                for i in range(ndim + 1):
                    p[i] = [0.5 * (p[i][j] + plo[j]) for j in range(ndim)]
                for i in range(ndim + 1):
                    if i != ilo:
                        y[i] = func(p[i])
                it += ndim
                for j in range(ndim):
                    s = 0.0
                    for i in range(ndim + 1):
                        s = s + p[i][j]
                    psum[j] = s


def run_amoeba(pt, func, rng_lo, rng_hi, bl, bu, p_maxpoints, tolf=1.0e-4, itmax=5000, stats=None):
    """TiktakGlobalSearch.f90:602-633; returns (amoeba_pt, iter)."""
    n = len(pt)
    x = simplex_coordinates2(n)
    scale = float(p_maxpoints) ** (1.0 / n)
    xT, fn = [], []
    for ia in range(n + 1):
        temp = [pt[i] + x[i][ia] * (rng_hi[i] - rng_lo[i]) / scale for i in range(n)]
        temp = [max(min(temp[i], bu[i]), bl[i]) for i in range(n)]
        fn.append(func(temp))
        xT.append(temp)
    it = amoeba(xT, fn, tolf, func, itmax, stats)
    ia = _iminloc(fn)
    return list(xT[ia]), it


# ------------------------------------------------------------------------------------------- BFGS polish
def _dot(a, b):
    s = 0.0
    for u, v in zip(a, b):
        s = s + u * v
    return s


def dograd(x, xinc, func):
    n = len(x)
    tolera = 10.0 * xinc
    dh = [(x[i] * xinc) if abs(x[i]) >= tolera else xinc for i in range(n)]
    grad = []
    for i in range(n):
        yy = [x[j] - (dh[i] if i == j else 0.0) for j in range(n)]
        f1 = func(yy)
        yy = [x[j] + (dh[i] if i == j else 0.0) for j in range(n)]
        f2 = func(yy)
        grad.append((f2 - f1) / (2.0 * dh[i]))
    return grad


def lnsrch(func, xold, fold, g, pdir, stpmax):
    """minimize.f90:2719-2814; returns (x, f, pdir).  When slope >= 0 the reference returns with x and f undefined; like the
    oracle and the library this keeps the old point and value."""
    n = len(xold)
    alf, tolx = 1.0e-4, EPS
    pabs = math.sqrt(_dot(pdir, pdir))
    if pabs > stpmax:
        pdir = [pdir[i] * stpmax / pabs for i in range(n)]
    slope = _dot(g, pdir)
    if slope >= 0.0:
        return list(xold), fold, pdir
    alamin = tolx / max(abs(pdir[i]) / max(abs(xold[i]), 1.0) for i in range(n))
    alam, alam2, f2 = 1.0, 0.0, 0.0
    while True:
        x = [xold[i] + alam * pdir[i] for i in range(n)]
        f = func(x)
        if alam < alamin:
            return list(xold), f, pdir
        if f <= fold + alf * alam * slope:
            return x, f, pdir
        if alam == 1.0:
            tmplam = -slope / (2.0 * (f - fold - slope))
        else:
            rhs1 = f - fold - alam * slope
            rhs2 = f2 - fold - alam2 * slope
            a = (rhs1 / alam ** 2 - rhs2 / alam2 ** 2) / (alam - alam2)
            b = (-alam2 * rhs1 / alam ** 2 + alam * rhs2 / alam2 ** 2) / (alam - alam2)
            if a == 0.0:
                tmplam = -slope / (2.0 * b)
            else:
                disc = b * b - 3.0 * a * slope
                if disc < 0.0:
                    tmplam = 0.5 * alam
                elif b <= 0.0:
                    tmplam = (-b + math.sqrt(disc)) / (3.0 * a)
                else:
                    tmplam = -slope / (b + math.sqrt(disc))
            if tmplam > 0.5 * alam:
                tmplam = 0.5 * alam
        alam2, f2 = alam, f
        alam = max(tmplam, 0.1 * alam)


def est_dfpmin(p, func, itmax=1000, gtol=1.0e-6):
    """minimize.f90:2572-2665; returns (p, fret, its)."""
    stpmx, tolx, xinc = 100.0, 4.0 * EPS, 1.0e-5
    nn = len(p)
    p = list(p)
    fp = func(p)
    g = dograd(p, xinc, func)
    hessin = [[1.0 if i == j else 0.0 for j in range(nn)] for i in range(nn)]
    xi = [-v for v in g]
    stpmax = stpmx * max(math.sqrt(_dot(p, p)), float(nn))
    fret = fp
    for its in range(1, itmax + 1):
        pnew, fret, xi = lnsrch(func, p, fp, g, xi, stpmax)
        fp = fret
        xi = [pnew[i] - p[i] for i in range(nn)]
        p = pnew
        if max(abs(xi[i]) / max(abs(p[i]), 1.0) for i in range(nn)) < tolx:
            return p, fret, its
        dg = list(g)
        g = dograd(p, xinc, func)
        den = max(fret, 1.0)
        if max(abs(g[i]) * max(abs(p[i]), 1.0) / den for i in range(nn)) < gtol:
            return p, fret, its
        dg = [g[i] - dg[i] for i in range(nn)]
        hdg = [_dot(hessin[i], dg) for i in range(nn)]
        fac = _dot(dg, xi)
        fae = _dot(dg, hdg)
        sumdg = _dot(dg, dg)
        sumxi = _dot(xi, xi)
        if fac > math.sqrt(EPS * sumdg * sumxi):
            fac = 1.0 / fac
            fad = 1.0 / fae
            dg = [fac * xi[i] - fad * hdg[i] for i in range(nn)]
            for i in range(nn):
                for j in range(nn):
                    hessin[i][j] = hessin[i][j] + fac * (xi[i] * xi[j]) - fad * (hdg[i] * hdg[j]) + fae * (dg[i] * dg[j])
        xi = [-_dot(hessin[i], g) for i in range(nn)]
    return p, fret, itmax
