import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The C restatement of the reference algorithm (oracle/nufi_oracle.c) -- the checker, never the product."""
    from oracle.oracle_py import Oracle, build

    if not os.path.exists(os.path.join(ROOT, "oracle", "build", "liboracle.so")):
        build(ref=os.path.isdir("/root/reference"))
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The real reference headers compiled in place (oracle/_ref); skipped when not built."""
    from oracle.oracle_py import Reference

    if not Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return Reference()
