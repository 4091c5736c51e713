"""oracle/workload.py rebuilds the headline workload without the product (bench.py --impl reference must not load
libikb200.so); it has to be the SAME workload the GPU arm solves."""
import subprocess
import sys
import os

import numpy as np

from ik_b200 import workloads as W
from oracle import workload as OW
from tests.common import make_workload, oracle_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_standalone_workload_is_bit_identical_to_the_product_side_one():
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    q0, tg, _ = make_workload(pb, om, 300, seed=12345, standing=W.CASSIE_STANDING, b0=65536)
    opb, q0s, tgs = OW.cassie_feet_pelvis(300, 12345, 65536)
    assert np.array_equal(q0, q0s) and np.array_equal(tg, tgs)
    assert opb.rows == 12 and opb.target_size == 36


def test_reference_arm_does_not_load_the_product_library():
    code = ("import sys; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','1','--batch','512'];"
            "import runpy; runpy.run_path(%r, run_name='__main__')" % os.path.join(ROOT, "bench.py"))
    probe = ("import sys\ntry:\n    exec(%r)\nexcept SystemExit:\n    pass\n"
             "maps=open('/proc/self/maps').read()\nassert 'libikb200' not in maps, 'reference arm mapped the product library'\n"
             "assert 'ik_b200' not in sys.modules\nprint('CLEAN')\n" % code)
    r = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "CLEAN" in r.stdout, r.stdout + r.stderr
    assert '"impl": "reference"' in r.stdout
