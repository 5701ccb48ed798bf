"""Import the UNMODIFIED reference (nanoporetech/remora @ /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (remora_b200/) may import this
module.  It exists so that (a) tests/golden/make_golden.py can generate golden vectors by
running the reference's own code, and (b) the `-m "not gpu"` tests can cross-check the
oracle restatement against the reference whenever /root/reference is present (it is NOT
present on the GPU box, where every such test skips).

What it does:
  * copies /root/reference/{src,models,setup.py,setup.cfg,README.rst} to a scratch dir
    under /tmp (the reference tree is read-only) and runs the reference's own
    `setup.py build_ext --inplace` there (3 Cython extensions),
  * inserts permissive stub modules for the third-party packages that are absent from
    this image and are never touched by the hot path (pysam, pod5, polars, plotnine,
    parasail, thop),
  * returns the imported `remora` package.
No reference source is copied into this repository.
"""
import importlib
import os
import shutil
import subprocess
import sys
import types

REFERENCE_ROOT = os.environ.get("REMORA_REFERENCE_ROOT", "/root/reference")
SCRATCH = os.environ.get("REMORA_REF_SCRATCH", "/tmp/remora_ref_scratch")
_ABSENT = ("pysam", "pod5", "polars", "plotnine", "parasail", "thop")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "remora", "encoded_kmers.pyx"))


class _Permissive(types.ModuleType):
    """Stub module: any attribute is another permissive stub; calling it raises."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Permissive(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        raise RuntimeError(f"stubbed third-party symbol {self.__name__} was called")


def _install_stubs():
    for name in _ABSENT:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Permissive(name)


def _build_scratch():
    marker = os.path.join(SCRATCH, ".built")
    if os.path.isfile(marker):
        return
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    os.makedirs(SCRATCH)
    for item in ("src", "models", "setup.py", "setup.cfg", "README.rst", "pyproject.toml"):
        src = os.path.join(REFERENCE_ROOT, item)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(SCRATCH, item), symlinks=False,
                            ignore=shutil.ignore_patterns("trained_models"))
        elif os.path.isfile(src):
            shutil.copy(src, os.path.join(SCRATCH, item))
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=SCRATCH,
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    open(marker, "w").close()


def import_reference():
    """Returns the reference `remora` package (built in scratch, third-party stubs in place)."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    _build_scratch()
    _install_stubs()
    src = os.path.join(SCRATCH, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import remora  # noqa: F401  (the reference package)
    return remora


def reference_model_path(arch="ConvLSTM_w_ref"):
    return os.path.join(SCRATCH, "models", f"{arch}.py")
