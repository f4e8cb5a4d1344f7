"""pytest plugin: runs the chain-level `-m gpu` tests on a machine without a GPU by putting tests/fake_sweeper.py in
GpuSweeper's place (both arms of every comparison are then the oracle: what is checked is the Python side -- the
GpuSweeper branch of api.runMCMC, the tests' own code -- not the device).

    PYTHONPATH=tests python -m pytest -p fake_gpu_plugin tests/test_gpu_chain.py tests/test_zz_gpu_late.py -m gpu \
        -k "chain" -q

Never loaded by the normal test runs (a plugin has to be named with -p)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import jwas_b200  # noqa: E402
from fake_sweeper import FakeSweeper  # noqa: E402

jwas_b200.api.GpuSweeper = FakeSweeper
jwas_b200.GpuSweeper = FakeSweeper
