import sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from importlib import import_module
import spiral_b200 as sp
ops = import_module(sp.__name__ + ".models.conv_ops")
def ref_conv(x, w):
    return F.conv2d(F.pad(x.permute(0,3,1,2),(1,2,1,2)), w.permute(3,2,0,1), stride=2).permute(0,2,3,1)
def rel(a,b): return float((a.double()-b).abs().max()/b.abs().max())
for (B,H,Cin,Cout) in [(4,32,64,128),(9,16,128,256),(40,4,512,1024),(70,2,1024,2048)]:
    g=torch.Generator().manual_seed(1)
    x=torch.randn((B,H,H,Cin),generator=g,dtype=torch.float64); w=torch.randn((5,5,Cin,Cout),generator=g,dtype=torch.float64)/np.sqrt(25*Cin)
    dy=torch.randn((B,H//2,H//2,Cout),generator=g,dtype=torch.float64)
    xr,wr=x.clone().requires_grad_(),w.clone().requires_grad_()
    y=ref_conv(xr,wr); dx,dw=torch.autograd.grad((y*dy).sum(),(xr,wr))
    for prec in (0,1,3):
        xc,wc,dyc=x.float().cuda(),w.float().cuda(),dy.float().cuda()
        print((B,H,Cin,Cout),"prec",prec,"fwd %.2e dgrad %.2e wgrad %.2e"%(rel(ops.conv_forward(xc,wc,prec).cpu(),y.detach()),rel(ops.conv_dgrad(dyc,wc,prec).cpu(),dx),rel(ops.conv_wgrad(xc,dyc,prec).cpu(),dw)))
