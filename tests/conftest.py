import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with gzip.open(os.path.join(GOLD, name), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


ALL9 = {'basepair': ['cWW', 'cSS', 'cHH', 'cHS', 'cHW', 'cSH', 'cSW', 'cWH', 'cWS', 'tSS', 'tHH', 'tHS', 'tHW', 'tSH',
                     'tSW', 'tWH', 'tWS', 'tWW', 'cWB', 'cBW'],
        'basepair_detail': [], 'stacking': [], 'sO': [], 'backbone': [], 'coplanar': [], 'sugar_ribose': [],
        'covalent': [], 'near': []}
