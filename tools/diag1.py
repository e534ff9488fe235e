import sys, os
sys.path[:0]=['/root/repo','/root/repo/gp-vae_b200','/root/repo/oracle','/root/repo/tests']
import torch, gp_kl_oracle as orc
from conftest import load_golden, rel_err
from gpu_util import compare, run_cuda
dev=torch.device('cuda:0')
g=load_golden('g4_v1_fixed_prior_grad')
fwd,bwd=run_cuda(g,dev,S=1,noise=g['noise'],grad_ell_p=True)
print('g4', rel_err(fwd['kl_pairs'],g['kl_pairs']), rel_err(fwd['z'],g['z']), rel_err(bwd['g_mean'],g['g_mean']), rel_err(bwd['g_ell_q'],g['g_ell_q']), rel_err(bwd['g_ell_p'],g['g_ell_p']))
print(bwd['g_ell_q'].cpu(), g['g_ell_q'], bwd['g_ell_p'].cpu(), g['g_ell_p'])
out,grads=orc.gp_prior_kl_grads(g['mean'],g['times'],g['lengths'],g['ell_q'],g['ell_p'],g['eps'],g['g_z'])
print('oracle vs golden', rel_err(grads['ell_q'],g['g_ell_q']), rel_err(grads['ell_p'],g['g_ell_p']))
for (B,D,T,S,r,post) in [(3,3,31,1,True,'gp'),(2,2,64,1,True,'gp'),(2,7,20,1,False,'diag')]:
    case=orc.synthetic_batch(B,D,T,S,ragged=r,seed=(100 if post=='gp' else 200)+T,posterior=post)
    print(T,post,compare(case,dev,kernel='rbf',S=S,posterior=post,grad_ell_p=True))
    print(' ell_q',case['ell_q'],'lengths',case['lengths'])
