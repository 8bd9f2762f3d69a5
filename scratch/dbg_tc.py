import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import test_rgbnet_tc_gpu as T
from plenvdb_b200 import synth
scene = synth.make_scene(96, "dense"); net = synth.rgbnet_init()
rays = synth.ray_batch(2048, H=200, W=200, K=synth.intrinsics(200, 200), seed=777)
a,*_=T._run(scene,net,rays,False); b,*_=T._run(scene,net,rays,True)
ga,gb=a.net_grad.cpu().numpy(), b.net_grad.cpu().numpy()
segs=[('w0',0,128*39,(128,39)),('b0',4992,5120,(128,)),('w1',5120,5120+16384,(128,128)),('b1',21504,21632,(128,)),('w2',21632,22016,(3,128)),('b2',22016,22019,(3,))]
for n,lo,hi,sh in segs:
    x,y=ga[lo:hi].reshape(sh),gb[lo:hi].reshape(sh)
    print(n,'max|ref|=%.3e max|err|=%.3e'%(np.abs(x).max(),np.abs(x-y).max()), 'nz tc', (y!=0).mean())
    if n in('w0','w1'):
        print('  row0 ref',x[0,:6],'\n  row0 tc ',y[0,:6]); print('  col err profile', np.abs(x-y).max(0)[:48].round(7))
        print('  row err profile', np.abs(x-y).max(1)[:16].round(7))
M=a.counters()['M_keep']
ka,kb=a.k0.grad.cpu().numpy(),b.k0.grad.cpu().numpy()
print('k0 grad max ref %.3e err %.3e'%(np.abs(ka).max(),np.abs(ka-kb).max()))
# dh1/dh0 check vs fp32 math recomputed in torch
h1=b.t['k_h1'][:M]; h0=b.t['k_h0'][:M]; g=b.t['k_rgb'][:M]
W=[torch.from_numpy(w).cuda() for w in synth.unpack_net(net)]
dh1=(g@W[4])*(h1>0); dh0=(dh1@W[2])*(h0>0)
print('dh1 err',float((b.t['k_dh1'][:M]-dh1).abs().max()), float(dh1.abs().max()))
print('dh0 err',float((b.t['k_dh0'][:M]-dh0).abs().max()), float(dh0.abs().max()))
x=b.t['k_x'][:M]; print('x[:,39] max', float(x[:,39].abs().max()), 'x vs feat', float((x[:,:12]-b.t['k_feat'][:M]).abs().max()))
dW1=dh1.t()@h0; print('dW1 torch vs tc', float((dW1-torch.from_numpy(gb[5120:5120+16384]).cuda().reshape(128,128)).abs().max()), float(dW1.abs().max()))
