import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the compiled reference in oracle/_ref")


@pytest.fixture(scope="session", autouse=True)
def built_artifacts():
    """A fresh checkout has no binaries (they are git-ignored): build the product library, the host programs
    and the checker once per session.  Building is not a fallback: without nvcc/gcc the tests fail loudly."""
    import subprocess

    import smart_b200

    if not os.path.exists(smart_b200.library_path()):
        smart_b200.build_library()
    if not os.path.exists(os.path.join(ROOT, "bin", "algos", "bf")):
        subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "smart_b200", "host")], check=True, capture_output=True)
    if not os.path.exists(os.path.join(ROOT, "oracle", "libsmart_oracle.so")):
        subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "oracle"), "oracle", "ref"], check=True, capture_output=True)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib.Oracle()
