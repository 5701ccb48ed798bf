"""The C-ABI library builds, loads and exports every symbol include/remora_b200.h declares.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

from remora_b200 import _native, build_native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_loads():
    path = build_native.build()
    assert os.path.isfile(path)
    lib = _native.load_library()
    assert lib.rb200_version() == 1


def test_every_declared_symbol_is_exported():
    header = open(os.path.join(ROOT, "include", "remora_b200.h")).read()
    declared = set(re.findall(r"\b(rb200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_model_desc_layout_matches_header():
    # 8 int32 + 12 conv descs (4 int32 + 2 int64 = 32 B) + int32 (+pad) + 6 int64 + 2 int32 + 2 int64
    assert ctypes.sizeof(_native.ConvDesc) == 32
    assert ctypes.sizeof(_native.ModelDesc) == 8 * 4 + 12 * 32 + 8 + 6 * 8 + 8 + 16


def test_sass_is_sm100a_with_tma():
    """The shipped cubin targets sm_100a and the encoder uses the bulk-copy (TMA) unit."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout
    sass = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True,
                          text=True).stdout
    funcs = sass.split("Function : ")
    enc = [f for f in funcs if "encode_dense_tma_kernel" in f.split("\n", 1)[0]]
    assert enc and "UBLKCP" in enc[0]  # cp.async.bulk shared->global


def test_sass_has_blackwell_tensor_core_and_tma_paths():
    """Evidence that the shipped cubin uses the sm_100a units the design claims: tcgen05.mma
    (SASS UTC*MMA), tcgen05.ld (LDTM), bulk TMA copies (UBLKCP) and packed fp32 FMA (FFMA2)."""
    import subprocess
    import pytest
    out = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    funcs = {f.split("\n", 1)[0]: f for f in out.stdout.split("Function : ")[1:]}
    k2tc = next(v for k, v in funcs.items() if "k2tc_kernel" in k)
    assert "UTCHMMA" in k2tc and "LDTM" in k2tc and "UBLKCP" in k2tc
    k1 = next(v for k, v in funcs.items() if "k1_front_kernel" in k)
    assert "FFMA2" in k1 and "UBLKCP" in k1
    k3 = next(v for k, v in funcs.items() if "k3_lstm_kernel" in k)
    assert "FFMA2" in k3 and "SHFL" in k3


def test_refinement_kernel_has_no_fused_multiply_add():
    """Bit-exact parity of the banded DP needs every product and sum rounded separately, like the
    reference's generated C: the kernel's SASS must not contain a single fused multiply-add, and its
    chain must be the FADD + FMNMX pair the design describes."""
    import subprocess
    import pytest
    out = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    funcs = {f.split("\n", 1)[0]: f for f in out.stdout.split("Function : ")[1:]}
    kernels = [v for k, v in funcs.items() if "refine_dp_kernel" in k]
    assert kernels
    for sass in kernels:
        assert "FFMA" not in sass and "FMNMX" in sass and "FADD" in sass and "LDS.128" in sass
