"""Measured denominators of the roofline on this GPU (ifk_debug_fp32_peak / ifk_debug_latencies).
    python tools/hw_microbench.py   -> one JSON line
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import _native  # noqa: E402

if __name__ == "__main__":
    print(json.dumps(_native.hw_microbench()))
