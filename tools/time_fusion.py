"""Time PointFusion over the 60-frame C3 sequence (run twice: default and E2E_FUSION_SEQUENCE=loop)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200"))
import torch  # noqa: E402
from benchmarks import fusion_bench  # noqa: E402

r = fusion_bench.run(torch.device("cuda", 0), repeats=5)
print(json.dumps({k: r[k] for k in ("value", "ms_per_sequence", "final_map_points", "gpu_launches")} | {"frac": r["roofline"]["frac"],
                  "mode": os.environ.get("E2E_FUSION_SEQUENCE", "coop")}))
print(json.dumps(r.get("batched")))
