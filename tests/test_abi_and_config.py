"""CPU tests: the C ABI library loads and exports every declared symbol; struct layout; config parser."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from tiktak_b200 import TikTakError, TikTakSearch, baseline_config, lib, parse_config
from tiktak_b200._abi import TiktakConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "tiktak_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tiktak_cuda_\w+)\s*\(", hdr)))


def test_library_exports_every_header_symbol(built):
    L = C.CDLL(lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tiktak_cuda.h but not exported"
    assert sorted(lib.SIGNATURES) == names, "tiktak_b200/lib.py binds a different set than the header declares"


def test_fortran_interface_binds_exported_symbols(built):
    """fortran/tiktak_cuda_mod.f90 (the ISO_C_BINDING side of the boundary) and INTEGRATION.md name only functions that
    the header declares and the library exports, with the header's argument counts; the BIND(C) struct has the
    header's field order"""
    f90 = open(os.path.join(ROOT, "fortran", "tiktak_cuda_mod.f90")).read()
    assert f90.split("MODULE tiktak_cuda_mod", 1)[1].strip() in open(os.path.join(ROOT, "INTEGRATION.md")).read().replace("\r", ""), \
        "INTEGRATION.md shows a different module than fortran/tiktak_cuda_mod.f90"
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "tiktak_cuda.h")).read(), flags=re.S)
    L = C.CDLL(lib.LIB_PATH)
    binds = re.findall(r"(?:FUNCTION|SUBROUTINE)\s+(\w+)\s*\(([^)]*)\)\s*&?\s*BIND\(C, NAME='(\w+)'\)", f90, flags=re.S)
    assert len(binds) >= 14
    for fname, fargs, cname in binds:
        assert fname == cname
        assert hasattr(L, cname), f"{cname} bound in Fortran but not exported"
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % cname, hdr, flags=re.S)
        assert m, f"{cname} not declared in the header"
        assert len([a for a in m.group(1).split(",") if a.strip() not in ("", "void")]) == len([a for a in fargs.split(",") if a.strip()]), cname
    # field order of TYPE, BIND(C) :: tiktak_config against the ctypes mirror (itself checked against the C struct below)
    body = f90.split("TYPE, BIND(C) :: tiktak_config", 1)[1].split("END TYPE", 1)[0]
    fields = []
    for line in body.splitlines():
        line = line.split("!")[0]
        if "::" in line:
            fields += [v.strip() for v in line.split("::", 1)[1].split(",")]
    assert fields == [f[0] for f in TiktakConfig._fields_]


def test_struct_layout_matches_c(built):
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "tiktak_cuda.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(tiktak_config), offsetof(tiktak_config,fvalmax), offsetof(tiktak_config,range), offsetof(tiktak_config,objective_id), offsetof(tiktak_config,nm_ftol), offsetof(tiktak_config,user));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    exp = [C.sizeof(TiktakConfig), TiktakConfig.fvalmax.offset, TiktakConfig.range.offset, TiktakConfig.objective_id.offset,
           TiktakConfig.nm_ftol.offset, TiktakConfig.user.offset]
    assert [int(v) for v in out] == exp


def test_no_cpu_fallback(built):
    """without a GPU the product path must fail loudly (never route through the oracle)"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(TikTakError) as e:
        TikTakSearch(baseline_config("C1"))
    assert e.value.code == -4
    src = open(os.path.join(ROOT, "tiktak_b200", "search.py")).read() + open(os.path.join(ROOT, "tiktak_b200", "lib.py")).read()
    assert "oracle" not in src.replace("the oracle", "")


def test_config_defaults_entry_point(built):
    L = lib.load()
    me, nd, lg, mp, lp = (C.c_int32(v) for v in (-1, -1, -1, -1, -1))
    fv = C.c_double(-1.0)
    assert L.tiktak_cuda_config_defaults(4, C.byref(me), C.byref(nd), C.byref(lg), C.byref(fv), C.byref(mp), C.byref(lp)) == 0
    assert (me.value, nd.value, lg.value, fv.value, mp.value, lp.value) == (200, 500, 10500, 1e8, 200, 200)
    nd, mp = C.c_int32(10), C.c_int32(11)
    assert L.tiktak_cuda_config_defaults(4, C.byref(me), C.byref(nd), C.byref(lg), C.byref(fv), C.byref(mp), C.byref(lp)) == -1
    assert b"maxpoints" in L.tiktak_cuda_last_error()


def test_parse_shipped_config():
    cfg = parse_config(os.path.join(ROOT, "tests", "golden", "config_c1.txt"))
    assert (cfg.nx, cfg.nm, cfg.maxeval, cfg.qr_ndraw, cfg.legitimate, cfg.fvalmax, cfg.maxpoints, cfg.search_type,
            cfg.lottery_points) == (4, 8, 40 * 5, 20, 10, 50.0, 3, 0, 10)
    assert cfg.p_maxpoints == 4 and cfg.p_ninterppt == 9
    assert np.array_equal(cfg.range, [[-1, 1]] * 4) and np.array_equal(cfg.bound, [[-3, 3]] * 4)
    assert np.array_equal(cfg.init, [0.811530031, -0.185093302, 0.299819619, 0.196328895])
    ref = baseline_config("C1")
    for k in ("nx", "nm", "maxeval", "qr_ndraw", "legitimate", "fvalmax", "maxpoints", "lottery_points"):
        assert getattr(cfg, k) == getattr(ref, k)


def test_parse_config_defaults_and_errors(tmp_path):
    body = "! c\n3\n2\n-1\n-1\n-1\n-1.0\n-1\n0\n-1\n!\n0 1\n0,1\n0 1\n!\n0.5\n0.5\n0.5\n!\n0 1\n0 1\n0 1\n"
    p = tmp_path / "c.txt"; p.write_text(body)
    cfg = parse_config(str(p))
    assert (cfg.maxeval, cfg.qr_ndraw, cfg.legitimate, cfg.fvalmax, cfg.maxpoints, cfg.lottery_points) == (160, 500, 10500, 1e8, 200, 200)
    p.write_text(body.replace("\n-1\n0\n-1\n!", "\n600\n0\n-1\n!", 1))
    with pytest.raises(ValueError):
        parse_config(str(p))
    p.write_text("3\n2\n")
    with pytest.raises(ValueError):
        parse_config(str(p))


def test_driver_exit_broadcast(built, tmp_path):
    """option -1 = exitState (stateControl.f90:203-216): state.dat = -1 and poisoned work counters, so that every instance
    sharing the directory -- CPU instances of the reference included -- stops at its next getNextNumber / getState"""
    import subprocess

    exe = os.path.join(ROOT, "driver", "TiktakGlobalSearch")
    assert os.path.exists(exe)
    subprocess.check_call([exe, "-1"], cwd=tmp_path)
    assert int(open(tmp_path / "state.dat").read().split()[0]) == -1
    for f in ("lastSobol.dat", "lastParam.dat", "seqNo.dat"):
        assert int(open(tmp_path / f).read().split()[0]) == -10000000


def test_fortran_patch_applies_to_the_reference(tmp_path):
    """fortran/TiktakGlobalSearch.patch is a unified diff against the reference's driver: it must apply cleanly, and the patched
    file must call the library where the stage bodies were (the reference tree only exists in the build container)"""
    import shutil
    import subprocess

    ref = "/root/reference/TiktakGlobalSearch.f90"
    if not os.path.exists(ref) or shutil.which("patch") is None:
        pytest.skip("reference tree or patch(1) not available here")
    shutil.copy(ref, tmp_path / "TiktakGlobalSearch.f90")
    patch = os.path.join(ROOT, "fortran", "TiktakGlobalSearch.patch")
    subprocess.check_call(["patch", "-p1", "--dry-run", "-i", patch], cwd=tmp_path, stdout=subprocess.DEVNULL)
    subprocess.check_call(["patch", "-p1", "-i", patch], cwd=tmp_path, stdout=subprocess.DEVNULL)
    out = open(tmp_path / "TiktakGlobalSearch.f90", encoding="latin-1").read()
    for call in ("tiktak_cuda_create", "tiktak_cuda_sobol_points", "tiktak_cuda_pretest(", "tiktak_cuda_choose", "tiktak_cuda_local_launch",
                 "tiktak_cuda_local_fetch", "tiktak_cuda_last_search"):
        assert call in out, call
    f90 = open(os.path.join(ROOT, "fortran", "tiktak_cuda_mod.f90")).read()
    for name in re.findall(r"(tiktak_cuda_\w+)\(", out):
        assert f"NAME='{name}'" in f90, name  # every call of the patched driver is bound by the interface module


def test_run_time_compiled_sobol_kernel_sources_build():
    """rtc.cu: the sources of the fixed-nx Sobol kernel that travel inside the library compile with NVRTC for (objective, nx)
    pairs that have no compiled-in instantiation -- no device needed for that (libnvrtc cross-compiles for sm_100a)"""
    import ctypes as C

    from tiktak_b200 import lib

    L = lib.load()
    L.tiktak_cuda_debug_rtc_compile.restype = C.c_long
    L.tiktak_cuda_debug_rtc_compile.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    log = C.create_string_buffer(8000)
    if L.tiktak_cuda_debug_rtc_compile(1, 7, 8, 1, 0, log, 8000) == -1:
        import pytest
        pytest.skip("libnvrtc is not installed here")
    for obj, nx, P, inb, G in ((1, 7, 8, 1, 0), (0, 12, 4, 1, 4), (2, 33, 16, 0, 0), (3, 64, 32, 1, 0), (1, 15, 8, 1, 5), (2, 1, 4, 1, 0)):
        size = L.tiktak_cuda_debug_rtc_compile(obj, nx, P, inb, G, log, 8000)
        assert size > 10000, (obj, nx, P, inb, G, size, log.value.decode()[:2000])
