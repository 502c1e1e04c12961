#!/usr/bin/env python3
"""Generates tests/golden/golden_small.json from an INDEPENDENT pure-Python restatement of the small
pieces of the reference (so the C++ oracle is pinned by something that shares no code with it):

  * Sobol: direction numbers by the Bratley-Fox recurrence (utilities.f90:981-1024) from the table
    file, points by the Gray-code closed form of i4_sobol (utilities.f90:1086-1106)
  * template objective + getPenalty (objective.f90:51-142) with Python's libm cos, sequential sums
  * theta-blend weights (TiktakGlobalSearch.f90:718) with numpy float32
  * simplex_coordinates2 (simplex.f90:346-438)

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def tables():
    txt = open(os.path.join(ROOT, "tables", "sobol_bf1111.inc")).read()
    poly = [int(v) for v in re.search(r"tk_sobol_poly\[1110\] = \{(.*?)\};", txt, re.S).group(1).replace("\n", "").split(",") if v.strip()]
    rows = re.findall(r"\{([0-9,]+)\},", txt.split("tk_sobol_vinit")[1])
    vinit = [[int(v) for v in r.split(",")] for r in rows]
    return poly, vinit


def dirnums(nx, nbits=30):
    poly, vinit = tables()
    m = []
    for d in range(nx):
        if d == 0:
            m.append([1] * nbits)
            continue
        a = poly[d - 1]
        deg = a.bit_length() - 1
        mm = list(vinit[d - 1][:deg])
        for j in range(deg, nbits):
            nv = mm[j - deg]
            for k in range(1, deg + 1):
                if (a >> (deg - k)) & 1:
                    nv ^= mm[j - k] << k
            mm.append(nv)
        m.append(mm)
    return m


def sobol_points(nx, n, nbits=30):
    m = dirnums(nx, nbits)
    pts = np.zeros((n, nx))
    for p in range(1, n + 1):
        g = p ^ (p >> 1)
        for d in range(nx):
            q = 0
            for j in range(nbits):
                if (g >> j) & 1:
                    q ^= m[d][j] << (nbits - 1 - j)
            pts[p - 1, d] = q / float(1 << nbits)
    return pts


def penalty(theta, bl, bu, fvalmax):
    pc = 10 * fvalmax
    pen = 0.0
    for t, l, u in zip(theta, bl, bu):
        a = max(0.0, l - t) / max(0.1, abs(l))
        b = max(0.0, t - u) / max(0.1, abs(u))
        pen = pen + (pc * (a * a) + pc * (b * b))
    if pen > np.finfo(float).eps:
        return min(1000.0 * fvalmax, fvalmax + pen)
    return 0.0


def objfun_template(theta, nm, bl, bu, fvalmax):
    pen = penalty(theta, bl, bu, fvalmax)
    if pen > np.finfo(float).eps:
        f = pen / math.sqrt(float(nm))
    else:
        val1 = (theta[0] * theta[0]) / 200.0
        for i in range(1, len(theta)):
            val1 = val1 + (theta[i] * theta[i]) / 200.0
        val2 = math.cos(theta[0])
        for i in range(1, len(theta)):
            val2 = val2 * math.cos(theta[i] / math.sqrt(i + 1 + 0.0))
        f = (val1 - val2 + 1.0) + 1.0
    dot = 0.0
    for _ in range(nm):
        dot = dot + f * f
    return dot + pen


def simplex2(n):
    x = np.zeros((n, n + 1))
    for j in range(n):
        x[j, j] = 1.0
    x[:, n] = (1.0 - math.sqrt(1.0 + n)) / n
    c = np.array([sum(x[i, j] for j in range(n + 1)) / (n + 1) for i in range(n)])
    x = x - c[:, None]
    s = math.sqrt(sum(x[i, 0] ** 2 for i in range(n)))
    return x / s


def main():
    out = {}
    out["sobol_unit_nx4_first16"] = sobol_points(4, 16).tolist()
    out["sobol_unit_nx10_first200"] = sobol_points(10, 200).tolist()
    out["sobol_unit_nx50_sample"] = sobol_points(50, 40)[[0, 1, 7, 38, 39]].tolist()
    pts = -1.0 + sobol_points(4, 10) * 2.0
    bl, bu = [-3.0] * 4, [3.0] * 4
    out["c1_points"] = pts.tolist()
    out["c1_fvals"] = [objfun_template(list(p), 8, bl, bu, 50.0) for p in pts]
    out["c1_finit"] = objfun_template([0.811530031, -0.185093302, 0.299819619, 0.196328895], 8, bl, bu, 50.0)
    out["c1_penalty_case"] = objfun_template([3.5, 0.0, 0.0, 0.0], 8, bl, bu, 50.0)
    out["w11_K4"] = [float(np.minimum(np.maximum(np.float32(0.10), np.sqrt(np.float32(k) / np.float32(4))), np.float32(0.95))) for k in range(1, 5)]
    out["w11_K1001"] = [float(np.minimum(np.maximum(np.float32(0.10), np.sqrt(np.float32(k) / np.float32(1001))), np.float32(0.95))) for k in (1, 2, 10, 11, 500, 903, 904, 1001)]
    out["simplex4"] = simplex2(4).tolist()
    out["simplex10"] = simplex2(10).tolist()
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden_small.json"), "w"), indent=0)
    print("wrote golden_small.json")


if __name__ == "__main__":
    main()
