import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    return json.loads((GOLDEN / "golden.json").read_text())


@pytest.fixture(scope="session")
def ctx():
    """One libflairb200 context for the whole GPU session (fails loudly without a GPU)."""
    import flair1_b200._native as nat
    c = nat.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def trained_3_15():
    """Briefly trained synthetic 3-band / 15-class checkpoint + the oracle model holding it."""
    from oracle import synth
    from oracle.unet_smp033 import Unet
    sd = synth.cached_checkpoint(3, 15)
    m = Unet(3, 15)
    m.load_state_dict(sd, strict=True)
    m.eval()
    return sd, m
