import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


# ---- parity numbers, not just gates: every GPU test may record what it measured; the session writes them to
# gpurun_out/parity_numbers.json (copied to profiles/rNN_parity_numbers.json for the judge)
_PARITY = {}


@pytest.fixture
def record(request):
    def _rec(**kv):
        _PARITY.setdefault(request.node.nodeid, {}).update({k: (float(v) if hasattr(v, "__float__") else v) for k, v in kv.items()})
    return _rec


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "parity_numbers.json")
    old = {}
    if os.path.exists(path):
        try:
            old = json.load(open(path))
        except Exception:
            old = {}
    old.update(_PARITY)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
