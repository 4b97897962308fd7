"""Reads one bench.py JSON line on stdin and prints: ms/step, dominant-kernel ms, roofline fraction, forward ms, backward ms.

    python bench.py --workload c2 --no-cpu-baseline | tail -1 | python benchmarks/c2_summary.py
"""
import json,sys
d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms"],4), round(d["roofline"]["frac"],3), round(d["roofline"]["fwd_ms"],4), round(d["roofline"]["bwd_ms"],4))
