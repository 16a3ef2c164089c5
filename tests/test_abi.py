"""CPU: the C-ABI shared library loads and exports every symbol that include/ubm.h declares; without a CUDA
device the library refuses to create a handle (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ubm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ubm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from unbiasedmcmc_b200 import _abi
    lib = _abi.load_lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/ubm.h but not exported by libubm.so: {missing}"


def test_abi_struct_sizes_match_header():
    """ctypes mirrors must have the layout the C compiler gives the header's structs."""
    import subprocess
    import tempfile
    from unbiasedmcmc_b200 import _abi
    names = ["ubm_draws", "ubm_model_normal_mixture", "ubm_model_mvn", "ubm_bins", "ubm_estimator_out",
             "ubm_model_ising_pt", "ubm_model_pg_logistic", "ubm_model_varsel"]
    hdr = open(os.path.join(ROOT, "include", "ubm.h")).read()
    names = [n for n in names if re.search(r"}\s*" + n + r"\s*;", hdr)]
    prog = '#include <stdio.h>\n#include "ubm.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        cfile = os.path.join(td, "s.c")
        open(cfile, "w").write(prog)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert C.sizeof(getattr(_abi, n)) == int(sz), n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from unbiasedmcmc_b200 import _abi
    from unbiasedmcmc_b200.engine import Engine, UbmError
    with pytest.raises(UbmError) as e:
        Engine(0)
    assert e.value.status == _abi.UBM_ERR_CUDA
