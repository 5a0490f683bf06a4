import os, sys, torch
sys.path.insert(0, os.getcwd())
import gwbp
n, d = 5_800_000, 512
feats = torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=1)
text = torch.from_numpy(gwbp.scene.make_text_queries(3, d, 0)).cuda()
for _ in range(2): gwbp.get_mask3d(feats, text, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): gwbp.get_mask3d(feats, text, 1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("mask3d ms", ms, "GB/s", n * d * 4 / ms / 1e6)
num = torch.randn(n, d, device="cuda"); den = torch.rand(n, device="cuda") + 0.1
out = torch.empty_like(num)
for _ in range(2): gwbp.finalize(num, den, out)
torch.cuda.synchronize()
e0.record()
for _ in range(5): gwbp.finalize(num, den, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("finalize ms", ms, "GB/s (r+w)", 2 * n * d * 4 / ms / 1e6)
