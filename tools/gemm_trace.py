import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import ops
dev=torch.device('cuda',0)
rng=np.random.default_rng(0)
a=torch.as_tensor(rng.standard_normal((2245,256)).astype(np.float32)).to(dev)
w=torch.as_tensor((rng.standard_normal((256,768))/16).astype(np.float32)).to(dev)
for i in range(3):
    out=ops.dense(a,w)
torch.cuda.synchronize()
print("done", float(out.abs().max()))
