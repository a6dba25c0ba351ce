"""Weighted-minimax fit of q(u) = log2(erfc(u/sqrt 2)) used by gelu_erf_fast (csrc/sepconv_common.cuh):
GELU(z) = max(z,0) - |0.5 z 2^q(|z|)|.  Prints, per degree, the float32-evaluated max abs error against scipy and the
coefficients.  Development aid (CPU): python tools/fit_gelu.py"""
import numpy as np
from scipy.special import erfc, erf
def target(u):  # log2(erfc(u/sqrt2))
    return np.log2(erfc(u/np.sqrt(2.0)))
U=8.0
def fit(N, iters=30):
    # weighted minimax via iteratively reweighted LSQ (Lawson)
    u=np.linspace(0,U,20001)
    t=target(u)
    E=erfc(u/np.sqrt(2))
    wt=(0.5*u*E*np.log(2))+1e-9   # d(GELU)/dq
    lw=np.ones_like(u)
    V=np.vander(u,N+1,increasing=True)
    for it in range(iters):
        w=wt*np.sqrt(lw)
        c,_,_,_=np.linalg.lstsq(V*w[:,None], t*w, rcond=None)
        err=np.abs((V@c-t)*wt)
        lw=lw*(err/err.max()+1e-3); lw/=lw.sum()
    return c
def gelu_ref(z): return 0.5*z*(1+erf(z/np.sqrt(2)))
z=np.linspace(-9,9,400001)
for N in (5,6,7,8):
    c=fit(N)
    c32=c.astype(np.float32)
    a=np.abs(z).astype(np.float32)
    q=np.full_like(a,c32[-1])
    for k in range(N-1,-1,-1): q=(q*a+c32[k]).astype(np.float32)
    E=np.exp2(q.astype(np.float64)).astype(np.float32)
    hz=(np.float32(0.5)*z.astype(np.float32))
    t=(hz*E).astype(np.float32)
    g=(np.maximum(z.astype(np.float32),0)-np.abs(t)).astype(np.float32)
    err=np.abs(g.astype(np.float64)-gelu_ref(z))
    print(N, 'max abs err', err.max(), 'at z=', z[err.argmax()], 'lead coef', c[-1], 'c0', c[0])
    print('   coefs', ', '.join(f'{x:.9e}' for x in c))
