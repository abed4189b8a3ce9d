import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "genie-tts_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

FIXTURE_ROOT = os.environ.get("GENIE_FIXTURE_ROOT", "/tmp/genie_b200_fixtures")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu)")


def fixture_dir(version: str, seed: int) -> str:
    """Random-init model directory in the converter's layout, written once per box."""
    from fixture_models import write_fixture
    d = os.path.join(FIXTURE_ROOT, f"{version}_seed{seed}")
    marker = os.path.join(d, ".complete")
    if not os.path.exists(marker):
        write_fixture(d, version, seed)
        open(marker, "w").close()
    return d


@pytest.fixture(scope="session")
def v2_dir():
    return fixture_dir("v2", 0)


@pytest.fixture(scope="session")
def v2pp_dir():
    return fixture_dir("v2ProPlus", 1)
