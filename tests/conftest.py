import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref/libfpohm_ref.so), the strongest checker we have."""
    from oracle import ref_oracle
    if not ref_oracle.available():
        pytest.skip("oracle/_ref/libfpohm_ref.so not built (needs /root/reference at build time)")
    return ref_oracle


@pytest.fixture(scope="session")
def fp():
    import fpohm_b200
    return fpohm_b200


@pytest.fixture(scope="session")
def ctx(fp):
    if fp.device_count() == 0:
        pytest.skip("no CUDA device")
    c = fp.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def port():
    """CPU restatement oracle (oracle/port); built on demand."""
    import subprocess
    from oracle import port_oracle
    if not port_oracle.available():
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle" / "port")], check=True)
    return port_oracle
