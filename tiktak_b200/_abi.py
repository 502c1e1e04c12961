"""ctypes mirror of include/tiktak_cuda.h (struct layout + constants).  No compute here."""
import ctypes as C

import numpy as np

OBJ_TEMPLATE, OBJ_GRIEWANK, OBJ_RASTRIGIN, OBJ_ROSENBROCK, OBJ_SMM = 0, 1, 2, 3, 4
ALG_DFNLS, ALG_AMOEBA, ALG_DFPMIN = 0, 1, 2
SOBOL_BLOCK = 16384

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class TiktakConfig(C.Structure):
    """struct tiktak_config (include/tiktak_cuda.h); field order must match exactly."""

    _fields_ = [
        ("nx", C.c_int32),
        ("nm", C.c_int32),
        ("maxeval", C.c_int32),
        ("qr_ndraw", C.c_int32),
        ("legitimate", C.c_int32),
        ("maxpoints_plus1", C.c_int32),
        ("search_type", C.c_int32),
        ("lottery_points", C.c_int32),
        ("fvalmax", C.c_double),
        ("range", c_double_p),
        ("init", c_double_p),
        ("bound", c_double_p),
        ("objective_id", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("nm_itmax", C.c_int32),
        ("nm_shrink_mode", C.c_int32),
        ("sobol_quirk", C.c_int32),
        ("text_roundtrip", C.c_int32),
        ("round_size", C.c_int32),
        ("nm_ftol", C.c_double),
        ("user", C.c_void_p),
    ]


def as_f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def dptr(a):
    return a.ctypes.data_as(c_double_p)


class TiktakSmmParams(C.Structure):
    """struct tiktak_smm_params (include/tiktak_cuda.h)"""

    _fields_ = [("np", C.c_int32), ("T", C.c_int32), ("seed", C.c_int32), ("reserved", C.c_int32), ("theta_star", C.c_double * 7)]


def smm_user(cfg):
    """ctypes object for tiktak_config.user from cfg.extra['smm'] (or None)"""
    d = getattr(cfg, "extra", {}).get("smm") if getattr(cfg, "extra", None) else None
    if not d:
        return None
    return TiktakSmmParams(np=int(d["np"]), T=int(d["T"]), seed=int(d.get("seed", 314159265)), reserved=0,
                           theta_star=(C.c_double * 7)(*[float(v) for v in d["theta_star"]]))
