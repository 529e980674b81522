"""CPU suite, part 2: the C-ABI library loads without a GPU, exports every symbol that
include/pdmpc_b200.h declares, the ctypes mirrors have the C compiler's struct layout,
and device-less calls fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pdmpc_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdmpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from pdmpc_b200 import capi
    lib = capi.load_library()
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pdmpc_b200.h but not exported"
    assert set(names) == set(capi.EXPORTED_SYMBOLS)
    assert lib.pdmpc_abi_version() == 2


def test_ctypes_structs_match_the_c_layout(tmp_path, built):
    from pdmpc_b200 import capi
    structs = {"pdmpc_mpa_desc": capi.MpaDesc, "pdmpc_batch_in": capi.BatchIn,
               "pdmpc_batch_out": capi.BatchOut, "pdmpc_stats": capi.Stats,
               "pdmpc_timestep_deps": capi.TimestepDepsC, "pdmpc_mcts_params": capi.MctsParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "inc.c"
    src.write_text(f'#include "{HEADER}"\nint main(void){{return PDMPC_OK;}}\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-c", "-o", str(tmp_path / "inc.o"), str(src)],
                   check=True)


def test_no_cpu_fallback_without_a_device(built):
    """On a box without CUDA pdmpc_create must fail with PDMPC_ERR_CUDA and say so; on a GPU
    box this test only checks the NULL-handle paths."""
    from pdmpc_b200 import capi
    lib = capi.load_library()
    assert lib.pdmpc_create(0, None) == capi.PDMPC_ERR_BAD_INPUT
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(capi.PdmpcError) as e:
            capi.Planner(0)
        assert e.value.code == capi.PDMPC_ERR_CUDA and "no CPU fallback" in str(e.value)
    assert lib.pdmpc_run_staged(None) == capi.PDMPC_ERR_BAD_INPUT
    assert lib.pdmpc_destroy(None) == capi.PDMPC_OK


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under p-dmpc_b200/ (or the MEX shim) may use it."""
    pkg = os.path.join(ROOT, "p-dmpc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".m")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', txt), f
                assert not re.search(r"\boracle_(plan|pq|interx|intersect|sincos)\w*\s*\(", txt), f


def test_pipeline_chunk_schedule_host_side():
    """pdmpc_pipeline_bounds (host-only): the chunks of the copy/search pipeline cover the batch exactly, in order, in
    at most 16 chunks; the default is ~60 000 searches per chunk with a first chunk of half the size."""
    from pdmpc_b200 import capi
    import numpy as np
    for n in (4, 1000, 16384, 16385, 20000, 47999, 56000, 168000, 179200, 358400, 716800, 2867200, 10_000_000):
        b = capi.pipeline_bounds(n)
        sizes = np.diff(b)
        assert b[0] == 0 and b[-1] == n and (sizes > 0).all() and 2 <= sizes.size <= 12, (n, b)
        assert sizes.size == min(12, max(2, n // 60000)), (n, sizes)
        if sizes.size >= 3:
            assert abs(int(sizes[0]) * 2 - int(sizes[1])) <= 2 and int(sizes[1:].max()) - int(sizes[1:].min()) <= 2, (n, sizes)
        for c in (2, 3, 5, 16):
            if n >= 2 * c:
                e = capi.pipeline_bounds(n, c)
                assert e.size == c + 1 and e[0] == 0 and e[-1] == n and (np.diff(e) > 0).all()
    assert capi.pipeline_bounds(358400).tolist() == [0, 39822, 119466, 199111, 278755, 358400]
    with pytest.raises(capi.PdmpcError):
        capi.pipeline_bounds(100, 1)
    with pytest.raises(capi.PdmpcError):
        capi.pipeline_bounds(100, 17)
