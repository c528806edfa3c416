"""examples/bars_learning.py: a script written against the reference's import names, run on this engine."""
import os
import re
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "examples", "bars_learning.py")


def test_example_reaches_the_engine_and_fails_loudly_without_a_gpu(tmp_path):
    """Everything up to the first EM step is host code (alias installation, model construction, data generation,
    standard_init, handlers): it must work anywhere, and the first step must refuse to run without a CUDA device."""
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test below")
    out = subprocess.run([sys.executable, SCRIPT, "bsc", "200"], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
    assert "Bars test BSC" in out.stdout and os.path.isdir(str(tmp_path / "output"))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bsc", "mca", "dsc"])
def test_example_recovers_the_bars(tmp_path, kind):
    """examples/bars_learning.py end to end on the device (flow of examples/barstests/bars-learning.py:23-88): run on a
    B200 in round 2, 15 s for the three models."""
    out = subprocess.run([sys.executable, SCRIPT, kind, "1000"], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    mae = float(re.search(r"generating bars: ([0-9.]+)", out.stdout).group(1))
    assert "Done" in out.stdout and mae < 1.0


REF_EXAMPLES = "/root/reference/examples"


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("script", ["anneal-example.py", "datalog-example.py"])
def test_reference_cpu_examples_print_the_same_on_this_package(tmp_path, script):
    """The reference's two examples that need no model run UNMODIFIED on `install_as_prosper()`; their output is compared
    line by line with the output of the reference itself (imported through the oracle's harness: fake mpi4py / tables)."""
    ours = ("import sys, runpy; sys.path.insert(0, %r); import prosper_b200; prosper_b200.install_as_prosper(); "
            "sys.argv = [%r]; runpy.run_path(%r, run_name='__main__')" % (ROOT, script, os.path.join(REF_EXAMPLES, script)))
    ref = ("import sys, runpy; sys.path.insert(0, %r); from oracle import ref_harness; ref_harness.load(); "
           "sys.argv = [%r]; runpy.run_path(%r, run_name='__main__')" % (ROOT, script, os.path.join(REF_EXAMPLES, script)))
    outs = []
    for code in (ours, ref):
        r = subprocess.run([sys.executable, "-W", "ignore", "-c", code], cwd=str(tmp_path), capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert len(outs[0].splitlines()) > 50 and outs[0] == outs[1]
