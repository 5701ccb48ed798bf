"""Build recipe for the oracle's native pieces.  TEST INFRASTRUCTURE.

  * oracle/liboracle.so          — our C restatement (oracle/oracle_encode.c), always built.
  * oracle/_ref/encoded_kmers*.so, oracle/_ref/refine_signal_map_core*.so — the REFERENCE's own
    Cython encoder and banded-DP refinement core, compiled from the sources where they lie
    (/root/reference/src/remora/{encoded_kmers,refine_signal_map_core}.pyx) with cython + gcc.  Only built
    when /root/reference is present (this container); the GPU box uses the prebuilt file that
    travels with the snapshot.  No reference source is copied into the repo: the generated .c
    goes to a temp dir, only the .so lands in oracle/_ref/ (git-ignored).
"""
import os
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/src/remora/encoded_kmers.pyx"
REF_REFINE_PYX = "/root/reference/src/remora/refine_signal_map_core.pyx"


def build_liboracle(force=False):
    srcs = [os.path.join(HERE, name) for name in ("oracle_encode.c", "oracle_refine.c")]
    out = os.path.join(HERE, "liboracle.so")
    if not force and os.path.isfile(out) and all(os.path.getmtime(out) >= os.path.getmtime(s)
                                                 for s in srcs):
        return out
    # -ffp-contract=off: the refinement DP must round every product and sum like the reference's
    # generated C does (no fused multiply-add)
    subprocess.run(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fPIC", "-shared", "-o", out]
                   + srcs + ["-lm"], check=True)
    return out


def _ref_ext_path(name):
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(HERE, "_ref", name + ext)


def _build_ref_ext(name, pyx, force=False):
    out = _ref_ext_path(name)
    if os.path.isfile(out) and not force:
        return out
    if not os.path.isfile(pyx):
        return None
    os.makedirs(os.path.dirname(out), exist_ok=True)
    import numpy as np
    with tempfile.TemporaryDirectory() as tmp:
        c_file = os.path.join(tmp, name + ".c")
        subprocess.run([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        inc = sysconfig.get_paths()["include"]
        subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-fwrapv",
                        "-I", inc, "-I", np.get_include(), c_file, "-o", out], check=True,
                       stderr=subprocess.DEVNULL)
    return out


def ref_encoder_path():
    return _ref_ext_path("encoded_kmers")


def build_ref_encoder(force=False):
    """Returns the path of oracle/_ref/encoded_kmers*.so, or None when it cannot be built."""
    return _build_ref_ext("encoded_kmers", REF_PYX, force)


def ref_refine_core_path():
    return _ref_ext_path("refine_signal_map_core")


def build_ref_refine_core(force=False):
    """The reference's banded-DP Cython module (refine_signal_map_core.pyx) into oracle/_ref.  It
    imports `remora.RemoraError` and `remora.constants` at load time: load it through
    load_ref_refine_core(), which provides a two-symbol stand-in package when the reference is absent
    (GPU box)."""
    return _build_ref_ext("refine_signal_map_core", REF_REFINE_PYX, force)


def load_ref_refine_core():
    """Imports oracle/_ref/refine_signal_map_core*.so (None when it was never built).  The module's
    only imports from its package are an exception class and two algorithm-name constants; when the
    reference package is not importable they are supplied by a stand-in module."""
    import importlib.util
    import types
    path = ref_refine_core_path()
    if not os.path.isfile(path):
        return None
    stand_in = "remora" not in sys.modules
    if stand_in:
        pkg = types.ModuleType("remora")
        pkg.RemoraError = type("RemoraError", (Exception,), {})
        consts = types.ModuleType("remora.constants")
        consts.REFINE_ALGO_VIT_NAME = "Viterbi"
        consts.REFINE_ALGO_DWELL_PEN_NAME = "dwell_penalty"
        pkg.constants = consts
        pkg.__path__ = []
        sys.modules["remora"] = pkg
        sys.modules["remora.constants"] = consts
    spec = importlib.util.spec_from_file_location("remora.refine_signal_map_core", path)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    finally:
        if stand_in:  # the extension has bound what it needs; do not shadow a later real import
            sys.modules.pop("remora", None)
            sys.modules.pop("remora.constants", None)
    return mod


if __name__ == "__main__":
    print(build_liboracle(force=True))
    print(build_ref_encoder(force=True))
    print(build_ref_refine_core(force=True))
