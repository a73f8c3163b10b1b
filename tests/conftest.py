import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.oracle()


@pytest.fixture(scope="session")
def reference():
    import oracle_lib
    if not oracle_lib.have_reference():
        pytest.skip("oracle/_ref/libzzref.so not built (needs /root/reference)")
    return oracle_lib.reference()


class Golden:
    """tests/golden/vectors.npz (made by tools/make_golden.py from the unmodified reference)."""

    def __init__(self):
        self.z = np.load(ROOT / "tests" / "golden" / "vectors.npz")
        self.cases = sorted({k.split("/")[0] for k in self.z.files if k.endswith("/input")})

    def input(self, case) -> bytes:
        return self.z[f"{case}/input"].tobytes()

    def stream(self, case, level):
        return self.z[f"{case}/L{level}/stream"].tobytes(), bool(self.z[f"{case}/L{level}/stream_ok"][0])

    def chunks(self, case, level):
        blob = self.z[f"{case}/L{level}/chunks"].tobytes()
        sizes = self.z[f"{case}/L{level}/sizes"].tolist()
        well = self.z[f"{case}/L{level}/wellformed"].tolist()
        out, pos = [], 0
        for s in sizes:
            out.append(blob[pos: pos + s]); pos += s
        return out, well

    def __getitem__(self, key):
        return self.z[key]


@pytest.fixture(scope="session")
def golden():
    return Golden()


def has_cuda() -> bool:
    try:
        import zzflate_b200
        return zzflate_b200.device_count() > 0
    except Exception:
        return False
