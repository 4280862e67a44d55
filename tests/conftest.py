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
    import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def keys(oracle):
    """SecretKey + ComputeKey at DEFAULT_128 from the harness seed."""
    return oracle.Keys()


@pytest.fixture(scope="session")
def client(oracle, keys):
    return oracle.Client(keys)


@pytest.fixture(scope="session")
def small_keys(oracle):
    return oracle.Keys(oracle.small_params(256, 32), seed=0x5EED)


@pytest.fixture(scope="session")
def evaluation(keys):
    """The product path: spf_b200.Evaluation over libspf_b200.so on cuda:0."""
    import spf_b200

    ev = spf_b200.Evaluation(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft)
    yield ev
    ev.close()
