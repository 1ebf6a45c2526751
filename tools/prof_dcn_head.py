"""Kernel-level breakdown of the deformable head (CenterHead dcn_head='fold_z') at the config-5 shape:
[B, 256, 16, 64, 160] -> forward + backward, torch.profiler CUDA-kernel table."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtpose_b200 import det3d_compat as D  # noqa: E402

B, C = int(os.environ.get("B", 4)), int(os.environ.get("C", 256))
torch.manual_seed(0)
head = D.CenterHead(in_channels=C, tasks=[dict(num_class=1, class_names=["Pelvis"])], dataset="cruw_pose", weight=0.7,
                    code_weights=[1.0] * 45, common_heads={"reg": (45, 2)}, share_conv_channel=C, dcn_head="fold_z").cuda()
x = torch.randn(B, C, 16, 64, 160, device="cuda", requires_grad=True)


def step():
    for p in head.parameters():
        p.grad = None
    x.grad = None
    preds, _ = head(x)
    (preds[0]["hm"].sum() + preds[0]["reg"].sum()).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
print("head fwd+bwd: %.1f ms per step (B=%d, C=%d)" % (e0.elapsed_time(e1) / 3, B, C))
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
