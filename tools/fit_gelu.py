"""Fit of the sigmoid-form exact-erf GELU used by csrc/common.cuh gelu_erf (run offline; prints the coefficients)."""
import numpy as np
from scipy.special import erfc
from scipy.optimize import least_squares
R=7.0
x=np.linspace(-R,R,40001); x=x[np.abs(x)>1e-9]
phi=0.5*erfc(-x/np.sqrt(2)); tgt=x*phi
def model(c,x,dt=np.float64):
    x=x.astype(dt); u=x*x
    p=np.full_like(x,dt(c[-1]))
    for ck in c[-2::-1]: p=p*u+dt(ck)
    z=x*p
    return x/(dt(1)+np.exp2(z))   # coefficients carry -log2(e)
deg=4
c=np.array([1.59565627214167,0.0729375808032982,-0.0002497224494760908,-6.116218759930543e-05,2.238175024077743e-06])*(-np.log2(np.e))
wts=np.ones_like(x)
best=None
for it in range(200):
    r=least_squares(lambda c:(model(c,x)-tgt)*wts,c,xtol=1e-15,ftol=1e-15,gtol=1e-15)
    c=r.x
    e=np.abs(model(c,x)-tgt)
    if best is None or e.max()<best[0]: best=(e.max(),c.copy())
    wts=wts*(1+ e/e.max()); wts/=wts.mean()
emax,c=best
print('fit max err',emax)
xx=np.linspace(-40,40,800001)
ref=xx*0.5*erfc(-xx/np.sqrt(2))
e64=np.abs(model(c,xx)-ref)
with np.errstate(over='ignore'):
    e32=np.abs(model(c,xx,np.float32).astype(np.float64)-ref)
print('max err f64',e64.max(),'at',xx[e64.argmax()],' f32',e32.max(),'at',xx[e32.argmax()])
rel=e32/np.maximum(np.abs(ref),1e-30)
m=(np.abs(ref)>6e-5)
print('max rel err where |gelu|>6e-5:',rel[m].max(),'at',xx[m][rel[m].argmax()])
# min of p(u) for monotonic saturation
u=np.linspace(0,1e4,1000001); p=np.polyval(c[::-1],u); print('max p (should be <0):',p.max())
print([float(np.float32(v)) for v in c])
print(['%.9e'%v for v in c])
