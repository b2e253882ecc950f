"""Every kernel on small ragged cases: odd payload alignment, frames cut off by the end of the buffer,
windows before/after the buffer, all modulations.  (compute-sanitizer is closed on the GPU pool, so the
bounds are exercised here and results compared between differently aligned copies of the same data.)"""
import os
import runpy

import pytest

pytestmark = pytest.mark.gpu
pytest.importorskip("torch")


def test_edge_case_driver_runs_clean():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, "tools", "sanitize_cases.py"), run_name="__main__")
