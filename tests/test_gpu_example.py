"""The notebook-shaped examples (flow over a cylinder, lid-driven cavity, porous flow) run end to end on the GPU."""
import os
import re
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


def test_ldc_cavity_example():
    """tests/ldcFVDBM.ipynb shape: 20 000 of the notebook's 500 001 steps; the centre-lines are already within 0.1 U_lid
    of ldc_Re100.mat (the full run is tests/test_gpu_ldc.py)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "ldc_cavity.py"), "--steps", "20000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "us/step" in r.stdout and "RMS centre-line error" in r.stdout
    e_u, e_v = map(float, re.search(r"u\(y\) ([0-9.]+), v\(x\) ([0-9.]+)", r.stdout).groups())
    assert e_u < 0.1 and e_v < 0.1, r.stdout


def test_porous_flow_example(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "porous_flow.py"), "--scale", "1.5", "--steps", "2000",
                        "--vtk", str(tmp_path / "porous")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MCUPS" in r.stdout and "non-finite values: 0" in r.stdout and os.path.getsize(tmp_path / "porous.vtk") > 10000
