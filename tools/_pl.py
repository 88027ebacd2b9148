import json
import sys
d = json.loads(sys.stdin.readline())
print(round(d["value"]), round(d["ms_per_step"], 4), round(d["final_loss"], 4), d["gpu_launches"] / d["steps"])
