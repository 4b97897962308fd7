import json,sys
d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms"],4), round(d["roofline"]["frac"],3), round(d["roofline"]["fwd_ms"],4), round(d["roofline"]["bwd_ms"],4))
