"""CPU: libpedoni_cuda.so loads without a GPU and exports every symbol include/pedoni_cuda.h declares;
the ctypes mirror of PedoniConfig has the C layout; compute entry points fail loudly without a device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from pedoni_b200 import _capi

HEADER = Path(__file__).resolve().parent.parent / "include" / "pedoni_cuda.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(pedoni_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _capi.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pedoni_cuda.h but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_capi.SIGNATURES) == names
    assert lib.pedoni_abi_version() == _capi.PEDONI_ABI_VERSION


def test_rust_extern_block_declares_the_same_functions():
    """ffi/pedoni-cuda-sys cannot be compiled here (no Rust toolchain); at least its extern block names
    exactly the functions of the header, and its PedoniConfig lists the header's fields in order."""
    root = HEADER.parent.parent
    rs = (root / "ffi" / "pedoni-cuda-sys" / "src" / "lib.rs").read_text()
    assert sorted(set(re.findall(r"pub fn (pedoni_[a-z_0-9]+)\s*\(", rs))) == declared_symbols()
    header = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    c_fields = re.findall(r"\b([a-z_0-9]+)\s*;", re.search(r"typedef struct PedoniConfig \{(.*?)\} PedoniConfig;", header, re.S).group(1))
    rs_fields = re.findall(r"pub ([a-z_0-9]+):", re.search(r"pub struct PedoniConfig \{(.*?)\n\}", rs, re.S).group(1))
    assert rs_fields == c_fields


def test_config_struct_layout_matches_the_header(tmp_path):
    """Compile the real header with gcc (as C: the boundary is a C ABI) and compare every field offset."""
    import subprocess
    fields = [n for n, _ in _capi.PedoniConfig._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pedoni_cuda.h"\nint main(void){\n'
                   'printf("%zu\\n", sizeof(PedoniConfig));\n' +
                   "".join(f'printf("%zu\\n", offsetof(PedoniConfig, {f}));\n' for f in fields) +
                   'printf("%zu\\n", sizeof(PedoniKernelTimes));\n'
                   'printf("%zu\\n", sizeof(PedoniObservables));\nprintf("%zu\\n", sizeof(PedoniSpawnGroup));\n'
                   'printf("%zu\\n", offsetof(PedoniObservables, arrived));\nreturn 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(HEADER.parent), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == C.sizeof(_capi.PedoniConfig)
    assert out[1:-4] == [getattr(_capi.PedoniConfig, f).offset for f in fields]
    assert out[-4] == C.sizeof(_capi.PedoniKernelTimes)
    assert out[-3] == C.sizeof(_capi.PedoniObservables) and out[-2] == C.sizeof(_capi.PedoniSpawnGroup)
    assert out[-1] == _capi.PedoniObservables.arrived.offset


def test_slab_rows_is_host_only():
    lib = _capi.load()
    r0, r1 = C.c_int32(), C.c_int32()
    assert lib.pedoni_slab_rows(2260, 8, 7, C.byref(r0), C.byref(r1)) == 0
    assert (r0.value, r1.value) == (1978, 2260)  # 2260 = 8 * 282 + 4: the first four slabs own 283 rows
    assert lib.pedoni_slab_rows(10, 0, 0, C.byref(r0), C.byref(r1)) == _capi.PEDONI_ERR_INVALID


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    import numpy as np
    import helpers
    from pedoni_b200 import PedoniError, SimulatorOptions, SocialForceModelCuda
    sc = helpers.scenario_of((10.0, 10.0), waypoints=[(1, 1, 1, 9, 1.0)])
    field = helpers.oracle_field(sc)
    with pytest.raises(PedoniError) as e:
        SocialForceModelCuda(SimulatorOptions(), sc, field)
    assert e.value.code == _capi.PEDONI_ERR_CUDA and "no CPU fallback" in str(e.value)
