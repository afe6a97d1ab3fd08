"""CPU: the C-ABI library loads, exports every symbol include/mifgpu.h declares, and fails loudly (no CPU
fallback) when there is no CUDA device.  No compute call needs a GPU here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mifgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mifgpu_[a-z_0-9]+)\s*\(", text)) - {"mifgpu_face_callback"})


def test_header_symbols_are_exported(mif):
    lib = mif.lib()
    names = declared_symbols()
    assert len(names) >= 16
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mifgpu.h but not exported by libmifgpu.so"
    assert sorted(mif.EXPORTED_SYMBOLS) == names


def test_abi_version(mif):
    assert mif.lib().mifgpu_abi_version() == 1


def test_struct_layout_matches_header(mif):
    # mifgpu_params: 3 x u64, 8 x f64, u32, 3 x i32, 3 x i32, i32 -> 120 bytes with natural alignment
    assert ctypes.sizeof(mif.Params) == 120
    assert mif.Params.Re.offset == 72 and mif.Params.num_time_steps.offset == 88 and mif.Params.device.offset == 116


def test_invalid_parameters_are_rejected(mif):
    with pytest.raises(mif.MifGpuError, match="-1"):
        mif.Context(1, 8, 8, 1, 1, 1, 0, 0, 0, 1.0, 1.0, 1)  # fewer than 2 points
    with pytest.raises(mif.MifGpuError, match="-1"):
        mif.Context(8, 8, 8, 1, 1, 1, 0, 0, 0, 1.0, 1.0, 0)  # zero time steps
    with pytest.raises(mif.MifGpuError, match="-1"):
        mif.Context(8, 8, 8, 1, 1, 1, 0, 0, 0, 1.0, 1.0, 1, Py=2, Pz=2, rank=7)  # rank outside the grid


def test_no_silent_cpu_fallback(mif):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(mif.MifGpuError, match="no CPU fallback"):
        mif.Context(8, 8, 8, 1, 1, 1, 0, 0, 0, 1.0, 1.0, 1)


def test_product_does_not_touch_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "mpi-incompressible-fluid_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or name == "Makefile":
                text = open(os.path.join(dirpath, name), errors="ignore").read()
                for line in text.splitlines():
                    code = line.split("//")[0].split("#")[0] if not name.endswith(".py") else line.split("#")[0]
                    assert "mif_oracle" not in code and "oracle/" not in code.replace("oracle/fft_cpu.h mo_r2r_exec", ""), \
                        (name, line)
