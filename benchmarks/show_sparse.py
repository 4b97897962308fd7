import json,sys
d=json.load(open(sys.argv[1]))
c=d.get('configs',{}).get('c1',d)
print(c.get("ms_per_step"), c.get("gpu_launches_per_step"), c.get("eager_ms_per_step"))
for k,v in sorted(c.get("kernels",{}).items(), key=lambda kv:-kv[1]["ms_per_step"])[:14]: print(k, v)
