import sys, time, os
sys.path.insert(0, "/root/repo")
import torch, bang_b200
from bang_b200 import builder, synth
dev = torch.device("cuda", 0)
n = int(float(sys.argv[1]))
base, c = synth.make_clustered(n, 128, "uint8", device=dev)
med = builder.find_medoid(base)
for passes in (2,):
    torch.cuda.synchronize(); t0 = time.time()
    nb = builder.build_vamana_gpu(base, med, L=64, passes=passes, device_out=True)
    torch.cuda.synchronize(); print("n", n, "passes", passes, "build s", round(time.time() - t0, 2), "deg", float((nb >= 0).sum(1).float().mean()))
