"""config.txt parser -- host-side mirror of parseConfig (genericParams.f90:265-450).

Same positional grammar: lines starting with '!' are comments; 9 scalars (nx, nmom, maxeval,
qr_ndraw, legitimate, fvalmax, maxpoints, selectType, lotteryPoints), then nx range rows, nx initial
guesses, nx bound rows.  Each value line is read list-directed: the first token(s) count, the rest
of the line (e.g. a trailing '! comment') is ignored.  Defaults for -1 follow :307-338, and
p_maxpoints = maxpoints + 1 (:369).
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class TikTakConfig:
    nx: int
    nm: int
    maxeval: int
    qr_ndraw: int
    legitimate: int
    fvalmax: float
    maxpoints: int  # raw config value; p_maxpoints = maxpoints + 1
    search_type: int
    lottery_points: int
    range: np.ndarray  # (nx, 2)
    init: np.ndarray  # (nx,)
    bound: np.ndarray  # (nx, 2)
    # knobs outside the reference's config.txt
    objective_id: int = 0
    nm_itmax: int = 5000  # effective cap of amoeba (minimize.f90:27-35); p_itmax_amoeba=1000 is dead
    nm_ftol: float = 1.0e-4  # p_tolf_amoeba (genericParams.f90:25)
    nm_shrink_mode: int = 0
    sobol_quirk: int = 0
    text_roundtrip: int = 1
    round_size: int = 1
    extra: dict = field(default_factory=dict)

    @property
    def p_maxpoints(self):
        return self.maxpoints + 1

    @property
    def p_ninterppt(self):
        return 2 * self.nx + 1  # genericParams.f90:311


def apply_defaults(nx, maxeval, qr_ndraw, legitimate, fvalmax, maxpoints, lottery_points):
    """genericParams.f90:307-338."""
    if maxeval < 0:
        maxeval = 40 * (nx + 1)
    if qr_ndraw < 0:
        qr_ndraw = 500
    if legitimate < 0:
        legitimate = qr_ndraw + 10000
    if fvalmax < 0.0:
        fvalmax = 100000000.0
    if qr_ndraw == 0:
        maxpoints = 0
    elif maxpoints < 0:
        maxpoints = min(200, qr_ndraw)
    elif maxpoints > qr_ndraw:
        raise ValueError("<genericParams:parseConfig> Error in config file. maxpoints > # sobol points")
    if lottery_points < 0:
        lottery_points = maxpoints
    return maxeval, qr_ndraw, legitimate, fvalmax, maxpoints, lottery_points


def _tokens(line, n):
    """list-directed read of n numbers from one record (comma or blank separated, Fortran D exponents)."""
    body = line.split("!")[0].replace(",", " ").split()
    if len(body) < n:
        raise ValueError(f"config: expected {n} value(s) in line {line!r}")
    return [float(t.replace("D", "E").replace("d", "e")) for t in body[:n]]


def parse_config(path, **overrides):
    with open(path, "r", encoding="latin-1") as f:
        lines = [ln.rstrip("\n") for ln in f]
    pos = 0

    def next_data_line(what):
        nonlocal pos
        while pos < len(lines):
            ln = lines[pos]
            pos += 1
            if ln[:1] == "!":
                continue
            return ln
        raise ValueError(f"<genericParams:parseConfig> Error. {path} Unable to read {what}")

    # the reference only skips '!' lines before the FIRST record of each group (:283-303, :378-395)
    scal = [_tokens(next_data_line("first line of parameters"), 1)[0]]
    for _ in range(8):
        if pos >= len(lines):
            raise ValueError(f"<genericParams:parseConfig> Error. {path} Unable to read first line of parameters")
        scal.append(_tokens(lines[pos], 1)[0])
        pos += 1
    nx, nm, maxeval, qr_ndraw, legit = (int(v) for v in scal[:5])
    fvalmax = float(scal[5])
    maxpoints, stype, lottery = (int(v) for v in scal[6:9])
    maxeval, qr_ndraw, legit, fvalmax, maxpoints, lottery = apply_defaults(
        nx, maxeval, qr_ndraw, legit, fvalmax, maxpoints, lottery)

    def group(ncols, what):
        nonlocal pos
        rows = [_tokens(next_data_line(what), ncols)]
        for _ in range(nx - 1):
            if pos >= len(lines):
                raise ValueError(f"<genericParams:parseConfig> Error. {path} Unable to read {what}")
            rows.append(_tokens(lines[pos], ncols))
            pos += 1
        return np.asarray(rows, dtype=np.float64)

    rng = group(2, "range of values for all parameters")
    init = group(1, "initial guesses for all parameters")[:, 0]
    bnd = group(2, "bounds for all parameters")
    cfg = TikTakConfig(nx=nx, nm=nm, maxeval=maxeval, qr_ndraw=qr_ndraw, legitimate=legit, fvalmax=fvalmax,
                       maxpoints=maxpoints, search_type=stype, lottery_points=lottery, range=rng, init=init, bound=bnd)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def make_config(nx, nm, qr_ndraw, maxpoints, lo, hi, init=None, bound=None, legitimate=-1, fvalmax=-1.0,
                maxeval=-1, search_type=0, lottery_points=-1, **kw):
    """Build a config programmatically with the same default rules as the file parser."""
    maxeval, qr_ndraw, legitimate, fvalmax, maxpoints, lottery_points = apply_defaults(
        nx, maxeval, qr_ndraw, legitimate, fvalmax, maxpoints, lottery_points)
    lo = np.broadcast_to(np.asarray(lo, dtype=np.float64), (nx,))
    hi = np.broadcast_to(np.asarray(hi, dtype=np.float64), (nx,))
    rng = np.stack([lo, hi], axis=1)
    bnd = rng.copy() if bound is None else np.asarray(bound, dtype=np.float64).reshape(nx, 2)
    ini = (0.5 * (lo + hi) if init is None else np.broadcast_to(np.asarray(init, dtype=np.float64), (nx,))).copy()
    return TikTakConfig(nx=nx, nm=nm, maxeval=maxeval, qr_ndraw=qr_ndraw, legitimate=legitimate, fvalmax=fvalmax,
                        maxpoints=maxpoints, search_type=search_type, lottery_points=lottery_points, range=rng,
                        init=ini, bound=bnd, **kw)


# the five configurations BASELINE.json names (SURVEY.md section 8d / BASELINE.md section 2)
def baseline_config(name, **kw):
    from ._abi import OBJ_GRIEWANK, OBJ_RASTRIGIN, OBJ_ROSENBROCK, OBJ_TEMPLATE

    if name == "C1":
        return make_config(4, 8, 20, 3, -1.0, 1.0, init=[0.811530031, -0.185093302, 0.299819619, 0.196328895],
                           bound=[[-3.0, 3.0]] * 4, legitimate=10, fvalmax=50.0, lottery_points=10,
                           objective_id=OBJ_TEMPLATE, **kw)
    if name == "C2":
        return make_config(10, 1, 1_000_000, 1000, -400.0, 600.0, init=[100.0] * 10, objective_id=OBJ_GRIEWANK, **kw)
    if name == "C3":
        return make_config(20, 1, 100_000_000, 10_000, -4.12, 5.12, init=[1.0] * 20, objective_id=OBJ_RASTRIGIN, **kw)
    if name == "C4":
        return make_config(50, 1, 10_000_000, 20_000, -5.0, 10.0, init=[2.5] * 50, objective_id=OBJ_ROSENBROCK, **kw)
    if name == "C5":
        return smm_config(**kw)
    raise KeyError(name)


#: true parameters and box of the synthetic SMM problem (theta = rho, sig_alpha, sig_eps, sig_eta1, sig_eta2, p_mix, mu_eta1)
SMM_THETA_STAR = [0.90, 0.45, 0.30, 0.20, 0.60, 0.70, 0.05]
SMM_BOUNDS = [[0.50, 0.99], [0.05, 1.50], [0.05, 1.50], [0.05, 1.50], [0.05, 1.50], [0.05, 0.95], [-0.50, 0.50]]


def smm_config(np_persons=131072, T=40, qr_ndraw=200_000, legitimate=100_000, maxpoints=1000, fvalmax=50.0, **kw):
    """BASELINE.json configs[4]: synthetic SMM least squares (30 T moments, 7 parameters), DFNLS local stage."""
    from ._abi import OBJ_SMM

    b = np.asarray(SMM_BOUNDS, dtype=np.float64)
    extra = {"smm": {"np": np_persons, "T": T, "seed": 314159265, "theta_star": SMM_THETA_STAR}}
    return make_config(7, 30 * T, qr_ndraw, maxpoints, b[:, 0], b[:, 1], init=0.5 * (b[:, 0] + b[:, 1]), legitimate=legitimate,
                       fvalmax=fvalmax, objective_id=OBJ_SMM, extra=extra, **kw)
