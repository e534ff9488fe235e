import sys, os
sys.path[:0]=['/root/repo','/root/repo/gp-vae_b200','/root/repo/oracle','/root/repo/tests']
import torch, gp_kl_oracle as orc
from gpu_util import run_cuda
dev=torch.device('cuda:0')
torch.set_printoptions(precision=4, linewidth=200)
for grid in (True, False):
    case=orc.synthetic_batch(1,2,8,1,ragged=False,seed=108,grid=grid)
    fwd,_=run_cuda(case,dev,tier='warp',grad_ell_p=False)
    fg,_=run_cuda(case,dev,tier='generic',grad_ell_p=False)
    print('grid',grid,'\n warp',fwd['z'].t().cpu(),'\n gen ',fg['z'].t().cpu(),'\n mean',case['mean'].t())
