"""The C-ABI library must load and export every symbol include/psim.h declares; without a CUDA
device it must fail loudly (no CPU fallback). No compute calls here — those are -m gpu tests."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "psim.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psim_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("psim_create", "psim_destroy", "psim_last_error", "psim_set_chromosomes", "psim_set_pedigree",
                 "psim_set_haplostructs", "psim_simulate", "psim_get_haplostructs", "psim_get_meiotic_configs",
                 "psim_set_map", "psim_set_allele_codes", "psim_emit", "psim_emit_all_dose"):
        assert must in names


def test_library_exports_every_declared_symbol(psim_lib):
    import pedigreesim_b200._lib as lib
    for name in declared_functions():
        assert hasattr(psim_lib, name), name
    assert sorted(lib.EXPORTS) == declared_functions()
    assert psim_lib.psim_abi_version() == 2


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "psim.h"\nint main(void){ psim_config c; psim_emit_targets t; (void)c; (void)t; return PSIM_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def test_struct_layouts_match_the_header(tmp_path):
    import pedigreesim_b200._lib as lib
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "psim.h"\nint main(void){ printf("%zu %zu %zu %zu\\n", sizeof(psim_config), '
                   'sizeof(psim_emit_targets), sizeof(psim_stats), sizeof(psim_device_hs)); return 0; }\n')
    exe = tmp_path / "s"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(lib.Config), C.sizeof(lib.EmitTargets), C.sizeof(lib.Stats), C.sizeof(lib.DeviceHS)]


def test_no_cpu_fallback_without_a_device(psim_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import pedigreesim_b200
    with pytest.raises(pedigreesim_b200.PsimError) as e:
        pedigreesim_b200.PsimContext(2)
    assert e.value.code == 2  # PSIM_ECUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "pedigreesim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in text and "_oracle" not in text and "psim_oracle" not in text, f
