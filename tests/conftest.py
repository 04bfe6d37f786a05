import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "multimodal-auv_b200", ROOT / "oracle", ROOT):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) libmauv_b200.so; nvcc cross-compiles without a GPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mauv_build", ROOT / "multimodal-auv_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()
