import numpy as np
from scipy.special import erf
X=4.0
def fit(deg, X=4.0):
    x = np.linspace(1e-3, X, 4000); t=x*x
    e = erf(x/np.sqrt(2)); u=np.arctanh(np.clip(e,-1+1e-16,1-1e-16)); q=u/x
    w = 0.5*x*x/np.cosh(u)**2
    V = np.vander(t, deg+1, increasing=True)
    c,*_ = np.linalg.lstsq(V*w[:,None], q*w, rcond=None)
    return c
def h(a): return a.astype(np.float16)
def f(a): return a.astype(np.float32)
def fma16(a,b,c): return h(f(a)*f(b)+f(c))
def gelu_h(xh, c, X, tanh_noise=0):
    xc = np.clip(xh, np.float16(-X), np.float16(X))
    th = h(f(xc)*f(xc))
    ch=[np.float16(v) for v in c]
    p = np.full_like(th, ch[-1])
    for k in range(len(ch)-2,-1,-1): p = fma16(p, th, np.full_like(th, ch[k]))
    uh = h(f(xc)*f(p))
    tn = np.tanh(uh.astype(np.float64))
    tn = h(tn*(1+tanh_noise*(np.random.default_rng(0).uniform(-1,1,tn.shape))))
    hx = h(f(xh)*np.float32(0.5))
    return fma16(hx, tn, hx)
xs = h(np.random.default_rng(1).normal(0,1.2,2000000).astype(np.float32))
x64 = xs.astype(np.float64)
exact = 0.5*x64*(1+erf(x64/np.sqrt(2)))
ideal = h(exact).astype(np.float64)
print('ideal fp16 out rounding: max', np.abs(ideal-exact).max(), 'rms', np.sqrt(((ideal-exact)**2).mean()))
for deg in (2,3):
  for Xc in (4.0,5.0):
    c=fit(deg,Xc)
    for noise in (0, 2**-11):
        g = gelu_h(xs,c,Xc,noise).astype(np.float64)
        err=np.abs(g-exact)
        print(deg,Xc,noise,'max',err.max(),'at',float(xs[err.argmax()]),'rms',np.sqrt((err**2).mean()))
# current f32 formulation then round
def gelu_cur(x):
    x=f(x); t=x*x
    p=np.float32(5.393212099136235e-09)
    for cc in (-3.9339354884759814e-07,1.1591534530452918e-05,-0.0001604528952157125,-9.176623279927298e-05,0.10483267903327942,2.3022100925445557):
        p=p*t+np.float32(cc)
    e=np.exp2((-x*p).astype(np.float64))
    return x/(1+e)
g=h(gelu_cur(xs)).astype(np.float64); err=np.abs(g-exact); print('current','max',err.max(),'rms',np.sqrt((err**2).mean()))
