"""Diagnostics (not a test): replay one golden scenario on the CUDA engine with stage timings."""
import json
import os
import sys

os.environ["VCE_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from golden_check import check_scenario  # noqa: E402
from rtg_tools_b200 import vcfeval  # noqa: E402

name = sys.argv[1]
sc = json.load(open(os.path.join(ROOT, "tests", "golden", "vcfeval_fixtures.json")))["scenarios"][name]
print(check_scenario(vcfeval.Engine(0), sc))
