"""dev check: CPU-stepped generic kernel logic vs the reference library (not collected by pytest)"""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from oracle import ref as R
r=R.ref()
emu=C.CDLL('/root/repo/tests/emu/libemu.so')
emu.emu_generic.argtypes=[C.c_int]*6+[C.c_void_p,C.c_void_p]+[C.c_longlong]*4+[C.c_int]
L_C_ORD,L_C_Z,L_R_TIME,L_R_ORD,L_R_Z=range(5)
S_C_ORD,S_C_Z,S_R_TIME,S_R_ORD,S_R_Z=range(5)
rng=np.random.default_rng(1)
def run(prec,N,tr,d,lm,sm,x):
    dt=np.float32 if prec==0 else np.float64
    n=N if tr==0 else 2*N
    x=np.ascontiguousarray(x,dtype=dt); o=np.zeros(n,dt)
    rc=emu.emu_generic(prec,N,tr,d,lm,sm,x.ctypes.data,o.ctypes.data,1,n,n,-1,n)
    assert rc==0,rc
    return o
worst=0
for prec,dt,tol in ((0,np.float32,2e-6),(1,np.float64,1e-14)):
  for N in [16,32,48,64,80,96,128,160,192,240,256,288,320,384,480,512,576,640,800,864,960,1024,2048,2592,4000,4096,12000]:
    for tr in (0,1):
        if not r.lib.pffft_is_valid_size(N,tr): continue
        n=N if tr==0 else 2*N
        x=(rng.random(n)*2-1).astype(dt)
        # forward ordered
        want=r.transform(N,tr,x,0,True,dt)
        got=run(prec,N,tr,0,L_R_TIME if tr==0 else L_C_ORD,S_R_ORD if tr==0 else S_C_ORD,x)
        e1=R.relmax(got,want)
        # forward unordered
        wantz=r.transform(N,tr,x,0,False,dt)
        gotz=run(prec,N,tr,0,L_R_TIME if tr==0 else L_C_ORD,S_R_Z if tr==0 else S_C_Z,x)
        e2=R.relmax(gotz,wantz)
        # backward ordered from reference spectrum
        wb=r.transform(N,tr,want,1,True,dt)
        gb=run(prec,N,tr,1,L_R_ORD if tr==0 else L_C_ORD,S_R_TIME if tr==0 else S_C_ORD,want)
        e3=R.relmax(gb,wb)
        # backward from z-domain
        wbz=r.transform(N,tr,wantz,1,False,dt)
        gbz=run(prec,N,tr,1,L_R_Z if tr==0 else L_C_Z,S_R_TIME if tr==0 else S_C_ORD,wantz)
        e4=R.relmax(gbz,wbz)
        m=max(e1,e2,e3,e4); worst=max(worst,m/tol)
        flag='' if m<tol else '  <<<<<< FAIL'
        print(f"{dt.__name__} N={N} tr={tr} fwd={e1:.2e} fwdz={e2:.2e} bwd={e3:.2e} bwdz={e4:.2e}{flag}")
print("worst/tol",worst)
