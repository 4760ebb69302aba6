import sys, torch
sys.path.insert(0, ".")
from deepi2p_b200 import point_ops
B, M, N, K = 64, 64, 16384, 64
g = torch.Generator(device="cuda").manual_seed(0)
pts = (torch.rand((B, 3, N), device="cuda", generator=g) * 20).contiguous()
nodes = (torch.rand((B, 3, M), device="cuda", generator=g) * 20).contiguous()
for _ in range(3):
    point_ops.ball_query_xyz_forward(pts, nodes, 2.0, K)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = point_ops.ball_query_xyz_forward(pts, nodes, 2.0, K)
e1.record(); e1.synchronize()
print("ball_query_xyz config-3 shape: %.1f us per call (grid build + query)" % (e0.elapsed_time(e1) * 100))
