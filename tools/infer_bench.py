#!/usr/bin/env python
"""bench.py's inference sub-metric on its own (padded vs length-bucketed passages and queries). GPU only."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

args = types.SimpleNamespace(nccl_ctas=0, no_cpu=True)
print(json.dumps(bench.bench_inference(bench.Ctx(args)), indent=1))
