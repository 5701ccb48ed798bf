"""Build recipe for the oracle's native pieces.  TEST INFRASTRUCTURE.

  * oracle/liboracle.so          — our C restatement (oracle/oracle_encode.c), always built.
  * oracle/_ref/encoded_kmers*.so — the REFERENCE's own Cython encoder, compiled from the source
    where it lies (/root/reference/src/remora/encoded_kmers.pyx) with cython + gcc.  Only built
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


def build_liboracle(force=False):
    src = os.path.join(HERE, "oracle_encode.c")
    out = os.path.join(HERE, "liboracle.so")
    if not force and os.path.isfile(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", out, src], check=True)
    return out


def ref_encoder_path():
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(HERE, "_ref", "encoded_kmers" + ext)


def build_ref_encoder(force=False):
    """Returns the path of oracle/_ref/encoded_kmers*.so, or None when it cannot be built."""
    out = ref_encoder_path()
    if os.path.isfile(out) and not force:
        return out
    if not os.path.isfile(REF_PYX):
        return None
    os.makedirs(os.path.dirname(out), exist_ok=True)
    import numpy as np
    with tempfile.TemporaryDirectory() as tmp:
        c_file = os.path.join(tmp, "encoded_kmers.c")
        subprocess.run([sys.executable, "-m", "cython", "-3", REF_PYX, "-o", c_file], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        inc = sysconfig.get_paths()["include"]
        subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-fwrapv",
                        "-I", inc, "-I", np.get_include(), c_file, "-o", out], check=True,
                       stderr=subprocess.DEVNULL)
    return out


if __name__ == "__main__":
    print(build_liboracle(force=True))
    print(build_ref_encoder(force=True))
