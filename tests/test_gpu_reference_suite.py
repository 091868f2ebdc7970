"""The reference's OWN test files, unmodified, against the drop-in on the B200.

tools/make_dropin_testdir.py (run by __graft_entry__.build() in the build container) copies feng/ddc/testing/*.py and
feng/pytest.ini from the reference checkout into the git-ignored baseline/_ref/feng/ and puts INTEGRATION.md's option-A shims
+ the regenerated coefficient files next to them; this test runs the reference's documented invocation there:

    cd feng/ddc/testing && PYTHONPATH=../src pytest            (feng/ddc/testing/test_ddc.py:2-3,14)
"""
import json
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "baseline", "_ref", "feng", "ddc", "testing")


@pytest.mark.skipif(not os.path.isdir(SUITE), reason="baseline/_ref/feng not materialised (tools/make_dropin_testdir.py)")
def test_reference_test_suite_passes_unmodified_on_the_drop_in():
    with open(os.path.join(ROOT, "baseline", "_ref", "feng", "MANIFEST.json")) as f:
        n_files = len(json.load(f)["reference_tests_sha256"])
    assert n_files == 2      # test_cwg.py, test_ddc.py
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(["../src", ROOT, env.get("PYTHONPATH", "")])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-q", "test_cwg.py", "test_ddc.py"], cwd=SUITE, env=env,
                       capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    mm = re.search(r"(\d+) passed", r.stdout)
    assert mm and int(mm.group(1)) == 6 and "failed" not in r.stdout, tail      # the reference's six known-answer tests
