"""Fixed workload for ncu: PointFusion over the first N frames of the synthetic room sequence (config C3)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200.slam import PointFusion, RGBDImages
from e2e_slam_b200.synthetic import room_sequence
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # B identical sequences in one cooperative launch
dev = torch.device("cuda:0")
depth, rgb, K, poses = room_sequence(frames, 480, 640, device=dev)
rgbd = RGBDImages(rgb.unsqueeze(0).repeat(batch, 1, 1, 1, 1), depth.unsqueeze(0).unsqueeze(-1).repeat(batch, 1, 1, 1, 1),
                  K.view(1, 1, 4, 4).repeat(batch, 1, 1, 1), poses.unsqueeze(0).repeat(batch, 1, 1, 1))
slam = PointFusion(odom="gt", dist_th=0.05, angle_th=20, sigma=0.6, device=dev)
with torch.no_grad():
    for _ in range(2):
        pc, _ = slam(rgbd)
torch.cuda.synchronize()
print("ok", int(pc._maps[0].n_dev.item()))
