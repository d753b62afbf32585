import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G:
        def __getattr__(self, name):
            d = np.load(os.path.join(GOLDEN, name + ".npz"))
            setattr(self, name, d)
            return d
    return G()


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The compiled reference, or None where oracle/_ref has not been built."""
    from oracle.oracle import Reference, reference_available
    return Reference() if reference_available() else None


@pytest.fixture(scope="session")
def sb():
    import scrappie_b200
    if not os.path.exists(scrappie_b200.LIB_PATH):
        scrappie_b200.build_library()
    scrappie_b200.lib()
    return scrappie_b200


@pytest.fixture(scope="session")
def engine(sb):
    eng = sb.Engine(0)
    yield eng
    eng.close()


def bundled_signal(golden, i):
    """float32 pA signal of bundled read i, scaled as src/fast5_interface.c:196-202."""
    import numpy as np
    reads = golden.reads
    sig = reads["r%d_signal" % i]
    dig, off, rng = [np.float32(v) for v in reads["r%d_meta" % i]]
    return ((sig.astype(np.float32) + off) * np.float32(rng / dig)).astype(np.float32)
