import sys, json
sys.path.insert(0, "/root/repo/end-to-end-self-supervised-slam_b200"); sys.path.insert(0, "/root/repo")
import torch
from benchmarks import c2_bench
print(json.dumps(c2_bench.run("cuda:0"), indent=1))
