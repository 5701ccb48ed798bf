import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def forward_cases():
    return np.load(os.path.join(GOLDEN, "forward_cases.npz"))


@pytest.fixture(scope="session")
def encode_cases():
    return np.load(os.path.join(GOLDEN, "encode_cases.npz"))


@pytest.fixture(scope="session")
def read_cases():
    meta = json.load(open(os.path.join(GOLDEN, "read_cases.json")))
    return meta, np.load(os.path.join(GOLDEN, "read_cases.npz"))


def unpack_encode_case(cases, cid):
    shape = tuple(int(x) for x in cases[f"c{cid}_shape"])
    n = int(np.prod(shape))
    want = np.unpackbits(cases[f"c{cid}_bits"])[:n].reshape(shape).astype(np.float32)
    return cases[f"c{cid}_seqs"], cases[f"c{cid}_maps"], cases[f"c{cid}_lens"], want


_STATE_CACHE = {}


def load_golden_model(name):
    """(state_dict, derived metadata) of a committed TorchScript fixture, on CPU."""
    if name not in _STATE_CACHE:
        from remora_b200 import model_util
        sd, md = model_util._raw_load_torchscript(os.path.join(GOLDEN, name + ".pt"))
        model_util.add_derived_metadata(md)
        _STATE_CACHE[name] = (sd, md)
    return _STATE_CACHE[name]
