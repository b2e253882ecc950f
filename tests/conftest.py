import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lte-gnu-radio-code_b200", "python")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npy"))
    g = {k: load(k) for k in ("tx_bits", "tx_online", "rx_offline", "chan_est_tim", "tx_data_0")}
    g["tx_bits"] = np.unpackbits(g["tx_bits"], axis=1)[:, :21600].astype(np.int64)
    return g
