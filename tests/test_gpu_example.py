"""The notebook-shaped example (examples/flow_over_cyl.py) runs end to end on the GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flow_over_cylinder_example(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "flow_over_cyl.py"), "--scale", "2", "--steps", "3000",
                        "--vtk", str(tmp_path / "cyl")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MCUPS" in r.stdout and os.path.getsize(tmp_path / "cyl.vtk") > 10000
