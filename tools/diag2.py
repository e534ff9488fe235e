import sys, os
sys.path[:0]=['/root/repo','/root/repo/gp-vae_b200','/root/repo/oracle','/root/repo/tests']
import torch, gp_kl_oracle as orc
from conftest import rel_err
from gpu_util import compare, run_cuda, run_oracle
dev=torch.device('cuda:0')
for (B,D,T,S,r) in [(2,3,8,1,False),(4,5,7,1,True)]:
    case=orc.synthetic_batch(B,D,T,S,ragged=r,seed=100+T)
    e=compare(case,dev,kernel='rbf',S=S,tier='warp',grad_ell_p=False); print(T,e)
    fwd,bwd=run_cuda(case,dev,tier='warp',grad_ell_p=False)
    fg,bg=run_cuda(case,dev,tier='generic',grad_ell_p=False)
    print(' kl warp',fwd['kl_pairs'][:6].cpu().numpy(),'\n kl gen ',fg['kl_pairs'][:6].cpu().numpy())
    print(' z diff',(fwd['z']-fg['z']).abs().max().item(),' gmean diff',(bwd['g_mean']-bg['g_mean']).abs().max().item())
    print(' gq warp',bwd['g_ell_q'].cpu().numpy(),' gen',bg['g_ell_q'].cpu().numpy())
    print(' logdets diff',(fwd['logdets']-fg['logdets']).abs().max(0).values.cpu().numpy())
    print(' times',case['times'][0])
