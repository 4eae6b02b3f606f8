import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DATA = os.path.join(ROOT, "oracle", "_ref", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def dict_paths(name):
    files = {"snort": ["snort.dict"], "et": ["et.dict"], "merged": ["snort.dict", "et.dict"]}[name]
    paths = [os.path.join(DATA, f) for f in files]
    for p in paths:
        if not os.path.exists(p):
            pytest.fail(f"{p} missing: run `make -C oracle` in the build container (copies the reference's data files)")
    return paths


@pytest.fixture(scope="session")
def oracle_merged():
    from oracle_lib import Oracle
    o = Oracle()
    for p in dict_paths("merged"):
        o.add_dict_file(p)
    o.compile()
    return o


@pytest.fixture(scope="session")
def dict_merged():
    import patternmatching_b200 as pm
    d = pm.Dictionary()
    for p in dict_paths("merged"):
        d.add_file(p)
    d.compile()
    return d


@pytest.fixture(scope="session")
def engine_merged(dict_merged):
    import patternmatching_b200 as pm
    return pm.Engine(dict_merged, device=0)


TINY_DICT = b"abcdefg\ncdefg\nefg\nafg\nfg\nhe\nshe\nhis\nhers\n|41 42|CD\n|41 |\n|4|\nABCDABDABCXYZ\n"
TINY_STREAM = b"ushers abcdefg xafg ABCD ABCDABDABCXYZ"
