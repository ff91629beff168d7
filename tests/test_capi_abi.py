"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU,
exports every symbol include/cimba_b200.h declares, keeps the reference's struct
layouts, and fails loudly (never falls back) when no CUDA device is present."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_functions():
    text = (ROOT / "include" / "cimba_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(cimba_b200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_and_library_agree(cb):
    from cimba_b200 import _lib
    names = declared_functions()
    assert len(names) >= 16
    raw = C.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/cimba_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "python binding and header drifted apart"


def test_struct_layouts_match_the_header(cb):
    from cimba_b200 import _lib
    # struct cmb_datasummary is 8 x 8 bytes (reference include/cmb_datasummary.h:42-51)
    assert C.sizeof(_lib.DataSummaryStruct) == 64
    assert _lib.DataSummaryStruct.count.offset == 8 and _lib.DataSummaryStruct.m1.offset == 32
    # the ctypes mirrors against the C compiler's view of include/cimba_b200.h: size and every field offset
    import subprocess, tempfile
    structs = {"cimba_b200_device_job": _lib.DeviceJob, "cimba_b200_experiment": _lib.Experiment,
               "cimba_b200_awacs_terrain": _lib.AwacsTerrain, "cimba_b200_wtdsummary": _lib.WtdSummaryStruct}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include/cimba_b200.h"}"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    with tempfile.TemporaryDirectory() as td:
        src, exe = Path(td) / "layout.c", Path(td) / "layout"
        src.write_text("\n".join(lines))
        subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True, capture_output=True)
        seen = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(seen[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(seen[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)
    # the reference's struct trial (benchmark/MM1_multi.c:39-45) is a prefix of TRIAL_DTYPE
    f = cb.TRIAL_DTYPE.fields
    assert [f[k][1] for k in ("arr_mean", "srv_mean", "obj_cnt", "sum_wait", "avg_wait")] == [0, 8, 16, 24, 32]


def test_version_and_seed_derivation(cb):
    assert cb.__version__ == "0.1.0"
    # cmb_random_fmix64 known answers
    assert cb.fmix64(0x34F05C64D7AD598F, 0) == 0xA9668314774003F8
    assert cb.fmix64(0x34F05C64D7AD598F, 1) == 0x3A4431CFB7782955


def test_no_cpu_fallback_without_a_device(cb):
    """On a box without CUDA every compute entry point must fail with ENODEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this checks the GPU-less behaviour")
    from cimba_b200 import _lib
    assert _lib.lib.cimba_b200_device_count() == 0
    exp = np.zeros(4, dtype=cb.TRIAL_DTYPE)
    exp["arr_mean"], exp["srv_mean"] = 1.1, 1.0
    with pytest.raises(cb.CimbaError) as e:
        cb.cimba_run_experiment(exp, num_objects=10, master_seed=1)
    assert e.value.code == _lib.ENODEVICE
    assert not exp["obj_cnt"].any()                     # nothing was computed anywhere
    with pytest.raises(ValueError):
        cb.launch_trials(torch.ones(2, dtype=torch.float64), torch.ones(2, dtype=torch.float64),
                         num_objects=1, master_seed=1)


def test_argument_validation_mirrors_reference_asserts(cb):
    # cimba_run_experiment asserts array != NULL, num_trials > 0, struct size > 0 (src/cimba.c:155-158)
    with pytest.raises(ValueError):
        cb.cimba_run_experiment(np.zeros(0, dtype=cb.TRIAL_DTYPE), num_objects=1, master_seed=1)
    with pytest.raises(TypeError):
        cb.cimba_run_experiment(np.zeros(3), num_objects=1, master_seed=1)
    with pytest.raises(TypeError):
        cb.cimba_run_experiment(np.zeros(3, dtype=[("x", "<f8")]), num_objects=1, master_seed=1)
    from cimba_b200 import _lib
    job = _lib.DeviceJob(num_trials=0)
    assert _lib.lib.cimba_b200_launch(C.byref(job), None) == _lib.EINVAL
    assert b"num_trials" in _lib.lib.cimba_b200_last_error()


def test_product_never_touches_the_oracle():
    """No file of the shipped package may reference oracle/ (parity would be void)."""
    for p in list((ROOT / "cimba_b200").rglob("*.py")) + list((ROOT / "cimba_b200" / "csrc").glob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".h"}:
            text = p.read_text()
            code = "\n".join(l for l in text.splitlines()
                             if not l.lstrip().startswith(("//", "/*", "*", "#", '"""')))
            assert "oracle_libs" not in code and "liboracle" not in code and "cimba_port" not in code, p


def test_awacs_entry_points_validate_before_touching_a_device(cb):
    """cimba_b200_awacs_set_terrain rejects an unusable struct terrain (tutorial/tut_5_1.c:96-108) with EINVAL, a usable
    one needs a device, MODEL_AWACS refuses the host-buffer path, and its workspace is 44 KB per trial."""
    import torch
    from cimba_b200 import _lib
    assert C.sizeof(_lib.AwacsTerrain) == 8 + 2 * 4 + 6 * 4
    bad = _lib.AwacsTerrain(map=None, cols=10, rows=10, x_scale=1.0, y_scale=1.0, x_min=-1.0, x_max=1.0, y_min=-1.0, y_max=1.0)
    assert _lib.lib.cimba_b200_awacs_set_terrain(C.byref(bad)) == -1          # CIMBA_B200_EINVAL
    assert b"terrain" in _lib.lib.cimba_b200_last_error()
    flat = _lib.AwacsTerrain(map=8, cols=10, rows=10, x_scale=1.0, y_scale=1.0, x_min=1.0, x_max=1.0, y_min=-1.0, y_max=1.0)
    assert _lib.lib.cimba_b200_awacs_set_terrain(C.byref(flat)) == -1
    assert _lib.lib.cimba_b200_awacs_upload_terrain(C.byref(bad)) == -1
    job = _lib.DeviceJob(model=cb.MODEL_AWACS, num_trials=3)
    assert _lib.lib.cimba_b200_workspace_bytes(C.byref(job)) == 3 * 1024 * 44
    if not torch.cuda.is_available():
        ok = _lib.AwacsTerrain(map=8, cols=10, rows=10, x_scale=1.0, y_scale=1.0, x_min=-1.0, x_max=1.0, y_min=-1.0, y_max=1.0)
        assert _lib.lib.cimba_b200_awacs_set_terrain(C.byref(ok)) == -2       # CIMBA_B200_ENODEVICE
        host = np.zeros(100, dtype=np.float32)
        ok.map = host.ctypes.data
        assert _lib.lib.cimba_b200_awacs_upload_terrain(C.byref(ok)) == -2
        exp = np.zeros(2, dtype=cb.TRIAL_DTYPE)
        with pytest.raises(cb.CimbaError):
            cb.cimba_run_experiment(exp, model=cb.MODEL_AWACS, num_objects=60, master_seed=1)


def test_thread_hooks_can_be_set_without_a_device(cb):
    """cimba_set_thread_hooks / cimba_thread_context (include/cimba.h:148-195): outside a worker thread there is no
    context, and setting or clearing the hooks needs no GPU."""
    from cimba_b200 import _lib
    calls = []
    init = _lib.THREAD_INIT_FUNC(lambda usrarg, tid: calls.append(tid) or 0)
    _lib.lib.cimba_b200_set_thread_hooks(C.cast(init, C.c_void_p), None, None)
    assert _lib.lib.cimba_b200_thread_context() is None
    _lib.lib.cimba_b200_set_thread_hooks(None, None, None)
    assert calls == []


def test_the_header_is_plain_c_and_the_stub_of_integration_md_links(cb, tmp_path):
    """include/cimba_b200.h compiled as C11 by gcc with warnings as errors, linked against the library, run."""
    import subprocess
    from cimba_b200 import _lib
    exe = tmp_path / "c_binding_stub"
    lib_dir = Path(_lib.LIB_PATH).parent
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests/c_binding_stub.c"),
                    "-o", str(exe), f"-L{lib_dir}", "-lcimba_b200", f"-Wl,-rpath,{lib_dir}"], check=True, capture_output=True)
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "fmix a9668314774003f8" in run.stdout and "mean 2.0" in run.stdout and "N        5  Mean    2.000" in run.stdout
