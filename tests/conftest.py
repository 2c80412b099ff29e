import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name), "r", encoding="utf-8") as f:
        return json.load(f)


def unhex(x):
    return float.fromhex(x)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def dense_small():
    z = np.load(os.path.join(GOLDEN, "dense_small.npz"))
    return z["x"], z["q"], load_golden("dense_small.json")


@pytest.fixture(scope="session")
def e2e_data():
    z = np.load(os.path.join(GOLDEN, "e2e_embeddings.npz"))
    table = {str(t): v for t, v in zip(z["qtexts"], z["qemb"])}
    return load_golden("e2e_retrieve.json"), z["emb"], table
