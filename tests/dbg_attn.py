import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sst_b200
from sst_b200 import lib as L
DEV='cuda'
B,H,Lq,dh=3,4,21,96
D=H*dh
torch.manual_seed(0)
qkv=torch.randn(B*Lq,3*D,device=DEV)*0.7
q_lens=torch.tensor([21,9,14],device=DEV,dtype=torch.int32)
o=torch.empty(B*Lq,D,device=DEV); lse=torch.empty(B*H*Lq,device=DEV)
d=L.attn_desc(L.F32,B,H,Lq,Lq,dh,3*D,3*D,3*D,D,True,True,0,1/math.sqrt(dh),0.0,0)
L.attn_fwd(d,qkv,qkv[:,D:],qkv[:,2*D:],None,q_lens,q_lens,o,lse)
torch.cuda.synchronize()
print('lse',lse.view(B,H,Lq)[2,3])
v=qkv[:,2*D:].view(B,Lq,H,dh)
print('o last row h3',o.view(B,Lq,H,dh)[2,20,3,:6]); print('mean v',v[2,:,3,:6].mean(0)); print('sum v', v[2,:,3,:6].sum(0))
print('o b1 row 10 h0',o.view(B,Lq,H,dh)[1,10,0,:6]); print('mean v',v[1,:,0,:6].mean(0))
