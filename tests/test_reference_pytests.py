"""The reference's own Python test-suite (/root/reference/tests/python_test.py, kept verbatim under
tests/golden/reference_tests) run UNCHANGED against the drop-in module mappy-rs_b200/mappy_rs: construction from the
reference's test.mmi, k / w / n_seq / seq_names / seq(), map() of the pinned read (target 0..400), map_batch over
list / tuple / iterator / generator inputs, 100 000-read batches with and without back-off, and every error message
the reference pins."""
import filecmp
import os
import subprocess
import sys

import pytest

from conftest import GOLDEN, PKG

REF_TESTS = os.path.join(GOLDEN, "reference_tests")


def test_reference_test_file_is_the_reference_copy():
    """If the reference tree is present (build container), the golden copy must be byte-identical to it."""
    src = "/root/reference/tests/python_test.py"
    if os.path.exists(src):
        assert filecmp.cmp(src, os.path.join(REF_TESTS, "tests", "python_test.py"), shallow=False)
        for f in ("test.fa", "test.mmi"):
            assert filecmp.cmp(os.path.join("/root/reference/resources/test", f), os.path.join(REF_TESTS, "resources", "test", f), shallow=False)


@pytest.mark.gpu
def test_reference_python_tests_pass_unchanged():
    env = dict(os.environ)
    env["PYTHONPATH"] = PKG + os.pathsep + env.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", REF_TESTS, "--confcutdir", REF_TESTS,
                        os.path.join(REF_TESTS, "tests", "python_test.py")], env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout, tail
    assert "19 passed" in r.stdout, tail   # 7 property/map tests + 2 x 100 000-read batches + 4 parametrised map_batch + 6 failure tests
