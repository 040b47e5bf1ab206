"""On the GPU box the driver selects `-m gpu`, which would leave out the tests that pin the CHECKER the parity tests
rely on. This one runs them there too: the oracle restatement against the committed fixtures (tests/test_golden.py)
and against the compiled reference libraries that travelled with the snapshot (tests/test_oracle_vs_ref.py; the cases
that need /root/reference itself skip themselves)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_is_pinned_on_this_machine_too():
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_golden.py"), os.path.join(ROOT, "tests", "test_oracle_vs_ref.py"),
                        os.path.join(ROOT, "tests", "test_tables_and_format.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout
