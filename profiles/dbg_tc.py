import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests'))
import oracle
from art_tts_b200 import monotonic_align, _lib
from conftest import rect_mask
cuda=torch.device('cuda:0')
B,F,T_x,T_y=int(sys.argv[1]) if len(sys.argv)>1 else 1024,80,190,872
rng=np.random.default_rng(7)
x_len=rng.integers(60,T_x+1,B).astype(np.int32)
y_len=np.minimum(870,4*x_len+rng.integers(0,100,B)).astype(np.int32)
gen=torch.Generator(device=cuda).manual_seed(3)
mu_x=torch.randn(B,F,T_x,device=cuda,generator=gen); y=torch.randn(B,F,T_y,device=cuda,generator=gen)
tx,ty=torch.from_numpy(x_len).to(cuda),torch.from_numpy(y_len).to(cuda)
path,dur,score,lp=monotonic_align.maximum_path_from_prior(mu_x,None,y,tx,ty,return_score=True,return_log_prior=True,flags=32)
p2,d2,s2=monotonic_align.maximum_path_from_prior(mu_x,None,y,tx,ty,return_score=True,flags=32)
torch.cuda.synchronize()
print('tap vs no-tap dur equal:', torch.equal(dur,d2), 'mismatching utts', (dur!=d2).any(1).nonzero().flatten()[:20].tolist())
bad=[]
pick=list(range(0,B,max(1,B//64)))
lpn=lp[pick].cpu().numpy()
want=oracle.maximum_path(lpn, rect_mask(x_len[pick],y_len[pick],T_x,T_y), n_threads=8)
got=path[pick].cpu().numpy(); got2=p2[pick].cpu().numpy()
for i,u in enumerate(pick):
    a=np.array_equal(got[i],want[i]); b=np.array_equal(got2[i],want[i])
    if not (a and b): bad.append((u,a,b,int(x_len[u]),int(y_len[u]), int((got2[i]!=want[i]).sum())))
print('bad (utt, tap ok, notap ok, tx, ty, ncells):', bad[:30], 'of', len(pick))
