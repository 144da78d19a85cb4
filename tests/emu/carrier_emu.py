"""dev check: spur-free dynamic range of the c1024 warp kernel arithmetic, stepped on the CPU"""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from oracle import ref as R
emu=C.CDLL('/root/repo/tests/emu/libemu.so')
emu.emu_w1024.argtypes=[C.c_int,C.c_void_p,C.c_void_p,C.c_longlong]
r=R.ref()
N=1024
def carrier(k,m):
    amp=1.0 if m%3==0 else 1.1
    freq = k/N if k<N/2 else (k-N)/N
    dphi=2*np.pi*freq
    if dphi<0: dphi+=2*np.pi
    phi0=(m%4)*0.125*np.pi
    ph=np.empty(N); phi=phi0
    for j in range(N):
        ph[j]=phi; phi+=dphi
        if phi>=np.pi: phi-=2*np.pi
    return np.stack([amp*np.cos(ph).astype(np.float32), amp*np.sin(ph).astype(np.float32)],-1).ravel().astype(np.float32)
def dyn(y,k):
    y=y.astype(np.float64); p=y[0::2]**2+y[1::2]**2
    return 10*np.log10(p[k])-10*np.log10(np.delete(p,k).max())
worst_e=1e9; worst_r=1e9
for m,k in enumerate(range(0,N,N//16)):
    x=carrier(k,m); o=np.zeros_like(x)
    emu.emu_w1024(0,x.ctypes.data,o.ctypes.data,1)
    w=r.transform(N,1,x,0,True)
    de,dr=dyn(o,k),dyn(w,k)
    worst_e=min(worst_e,de); worst_r=min(worst_r,dr)
    print(k,round(de,1),round(dr,1))
print("worst emu",worst_e,"worst ref",worst_r)
