"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/g4c.h
declares, and the ctypes mirrors have the C struct sizes (compiled with gcc from the header)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built_lib():
    from graphs4cfd_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "g4c.h")).read()
    return re.findall(r"^G4C_API [\w\s\*]+?\b(g4c_\w+)\(", text, flags=re.M)


def test_library_exports_every_declared_symbol(built_lib):
    names = declared_symbols()
    assert len(names) >= 13
    handle = ctypes.CDLL(built_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    assert set(names) == set(built_lib.EXPORTS), "ctypes table and header disagree"
    assert built_lib.lib().g4c_version() == 100


def test_ctypes_struct_sizes_match_header(built_lib, tmp_path):
    pairs = {"G4cMlp": built_lib.Mlp, "G4cSeg": built_lib.Seg, "G4cRowMlpDesc": built_lib.RowMlpDesc,
             "G4cMpDesc": built_lib.MpDesc, "G4cSegReduceDesc": built_lib.SegReduceDesc,
             "G4cProjectDesc": built_lib.ProjectDesc, "G4cEdgeToNodeDesc": built_lib.EdgeToNodeDesc,
             "G4cInterpDesc": built_lib.InterpDesc, "G4cStepUpdateDesc": built_lib.StepUpdateDesc,
             "G4cHaloDesc": built_lib.HaloDesc, "G4cEdgeDesc": built_lib.EdgeDesc, "G4cRowTcDesc": built_lib.RowTcDesc,
             "G4cKnnDesc": built_lib.KnnDesc, "G4cHaloPutDesc": built_lib.HaloPutDesc}
    src = tmp_path / "sz.c"
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in pairs)
    src.write_text(f'#include <stdio.h>\n#include "g4c.h"\nint main(void){{{body}return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        name, size = line.split()
        assert ctypes.sizeof(pairs[name]) == int(size), name


def test_ctypes_field_offsets_match_header(built_lib, tmp_path):
    """Every field of every descriptor sits at the offset the C compiler gives it (same names on both sides)."""
    pairs = {"G4cMlp": built_lib.Mlp, "G4cSeg": built_lib.Seg, "G4cRowMlpDesc": built_lib.RowMlpDesc,
             "G4cMpDesc": built_lib.MpDesc, "G4cEdgeDesc": built_lib.EdgeDesc, "G4cRowTcDesc": built_lib.RowTcDesc,
             "G4cSegReduceDesc": built_lib.SegReduceDesc, "G4cProjectDesc": built_lib.ProjectDesc,
             "G4cEdgeToNodeDesc": built_lib.EdgeToNodeDesc, "G4cInterpDesc": built_lib.InterpDesc,
             "G4cStepUpdateDesc": built_lib.StepUpdateDesc, "G4cHaloDesc": built_lib.HaloDesc, "G4cKnnDesc": built_lib.KnnDesc, "G4cHaloPutDesc": built_lib.HaloPutDesc}
    lines = []
    for cname, cls in pairs.items():
        for fname, *_ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "off.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "g4c.h"\nint main(void){' + "".join(lines) + "return 0;}\n")
    exe = tmp_path / "off"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    n = 0
    for line in out.strip().splitlines():
        cname, fname, off = line.split()
        assert getattr(pairs[cname], fname).offset == int(off), (cname, fname)
        n += 1
    assert n == len(lines)


def test_host_plan_helper_matches_python_loop(built_lib):
    import numpy as np
    rng = np.random.default_rng(0)
    n, k = 500, 5
    senders = rng.integers(0, n, size=(n, k))
    want = np.ones(n, dtype=bool)
    for i in range(n):
        if want[i]:
            want[senders[i]] = False
    got = built_lib.host_guillard(senders, n)
    assert (got == want).all()


def test_bad_descriptor_is_reported_not_crashed(built_lib):
    d = built_lib.MpDesc()
    d.hidden, d.n_targets, d.n_edges, d.fixed_k = 128, 10, 0, 0
    rc = built_lib.lib().g4c_mp_fwd(ctypes.byref(d), None)
    assert rc != 0 and b"g4c_mp_fwd" in built_lib.lib().g4c_last_error()
