"""Pretty-print the parts of a bench.py JSON line that matter: python scripts/show_bench.py file.json"""
import json
import sys
d = json.load(open(sys.argv[1]))
print(json.dumps({k: v for k, v in d.items() if k not in ("config", "roofline", "kernel_profile_ms_per_step")})[:1800])
print(json.dumps(d["config"].get("also_measured"), indent=1)[:5000])
print(json.dumps(d.get("roofline"), indent=1)[:7000])
print(json.dumps(d.get("kernel_profile_ms_per_step"), indent=1))
