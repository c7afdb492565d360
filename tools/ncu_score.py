"""ncu target: score_topk (K11/K12) on 9,472 users x 1 M items x d = 768."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops
I, d, U = 1_000_000, 768, 9472
table = (torch.randn(I, d, device="cuda") * d ** -0.5).to(torch.bfloat16)
users = torch.randn(U, d, device="cuda").to(torch.bfloat16)
hist = torch.randint(1, I, (U, 20), device="cuda", dtype=torch.int32)
for _ in range(2):
    ops.score_topk(users, table, id_base=0, history=hist, k=10)
torch.cuda.synchronize()
print("done")
