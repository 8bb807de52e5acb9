"""The per-step weight preparation of a workload alone (for an ncu launch list of prepare_kernel / wave_pack_kernel).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prep_only.py [--workload glow_imagenet32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from inverse_flow_b200.stack import InvConvStack  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="glow_imagenet32", choices=sorted(WORKLOADS))
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    stages, batch, _ = WORKLOADS[args.workload]
    stack = InvConvStack(stages, batch, groups=1)
    for _ in range(args.reps):
        for st in stack.stages:
            stack.prepare_stage(st)
        torch.cuda.synchronize()
    # warm, back to back, CUDA events per stage
    for st in stack.stages:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            stack.prepare_stage(st)
        e1.record()
        e1.synchronize()
        print("stage C=%d: %.1f us per prepare (prepare_kernel + pack, 20 back to back)" % (st.C, e0.elapsed_time(e1) / 20 * 1e3))


if __name__ == "__main__":
    main()
