"""One launch of the bf16 similarity + top-100 kernel at the BASELINE.json configs[3] per-GPU shard shape on 8 GPUs
(32 768 query rows x 262 144 database rows x 32 768-D, database 17.2 GB in bf16) for `ncu --set full`."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
sys.path.insert(0, ROOT)
import bench
from pyvisim_b200 import retrieval
dev = torch.device("cuda", 0)
db = bench.vlad_like_device(262144, 32768, dev, torch.Generator(device=dev).manual_seed(4321), torch.bfloat16)
s, i = retrieval.cosine_topk(db[:32768], db, 100)
torch.cuda.synchronize()
print("ok", float(s[0, 0]), int(i[0, 0]))
