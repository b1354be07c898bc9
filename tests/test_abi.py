"""CPU: the C-ABI library loads without a GPU and exports every symbol include/omni_avsr.h declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "omni_avsr.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(omni_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from omni_avsr_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/omni_avsr.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_abi_version_and_no_gpu_error_path():
    from omni_avsr_b200 import _lib
    assert _lib.lib.omni_abi_version() == 1
    import torch
    if not torch.cuda.is_available():
        assert _lib.lib.omni_device_cc() < 0          # reports an error code instead of crashing


def test_cpu_tensors_are_refused():
    import pytest
    import torch
    from omni_avsr_b200 import ops
    from omni_avsr_b200._lib import OmniKernelError
    with pytest.raises(OmniKernelError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(OmniKernelError):
        ops.matryoshka_compress(torch.zeros(1, 8, 8, dtype=torch.bfloat16), 8, 2)


def test_product_does_not_import_oracle():
    pkg = ROOT / "omni_avsr_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
