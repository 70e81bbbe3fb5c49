"""stdin: a bench.py JSON line -> one short line (value, single lane, compose / front kernel times)"""
import json, sys
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("c2 value", round(l["value"]), "single", round(l["single_lane"]["value"]), "compose_us", round(l["roofline"]["kernel_ms"] * 1e3, 1),
      "front_us", round(list(l["roofline"]["other_kernels_ms"].values())[0] * 1e3, 1), "e2e", round(l["e2e"]["value"]))
