"""Whole 1080p solves under every combination of iteration kernels (default, no temporal blocking,
no cluster-resident kernel): prints whether flows and iteration counts agree bitwise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import optical_flow_1_b200 as pkg
P=16; nx,ny=1920,1080
I0, I1 = pkg.synth.make_batch_torch(P, nx, ny, seed=1234, device="cuda")
outs=[]
for env in [{}, {"TVL1_NO_TB":"1"}, {"TVL1_NO_TB":"1","TVL1_NO_RESIDENT":"1"}, {"TVL1_NO_RESIDENT":"1"}]:
    for k in ("TVL1_NO_TB","TVL1_NO_RESIDENT"): os.environ.pop(k,None)
    os.environ.update(env)
    g = pkg.TVL1(0, max_batch=P)
    u1, u2 = torch.empty_like(I0), torch.empty_like(I0)
    it, er = g.solve_batch_device(I0.data_ptr(), I1.data_ptr(), u1.data_ptr(), u2.data_ptr(), P, nx, ny, want_iters=True)
    outs.append((u1,u2,it,er)); g.close()
    a=outs[0]
    print(env, "iters same:", np.array_equal(a[2],it), "max|du|", float((a[0]-u1).abs().max()), float((a[1]-u2).abs().max()),
          "nnz", int((a[0]!=u1).sum()), "errs max rel", float(np.abs(a[3]-er).max()/np.abs(er).max()))
    if not np.array_equal(a[2],it):
        print(np.argwhere(a[2]!=it)[:10], a[2][a[2]!=it][:10], it[a[2]!=it][:10])
