"""The CUDA path on the repeat-rich golden case (tests/golden/pe100_repeats: an index with planted exact repeats, so
REPEAT-type reads, several occurrences per SA interval, the random pick among them, mapQ 0 and pairing over many positions
all occur).  The case was added after the round's last GPU run - on the CPU the oracle and the host-instantiated device
functions already match it (tests/test_golden.py, tests/test_emul_golden.py) - so its GPU check sits in a file that sorts
last: whatever it finds cannot hide the results of the other GPU tests."""
import pytest

import make_golden_path  # noqa: F401  (puts tests/golden on sys.path)
import make_golden
from test_golden import _cuda_case


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(make_golden.GPU_LATE))
def test_cuda_path_against_repeat_golden(name, tmp_path):
    _cuda_case(name, tmp_path)
