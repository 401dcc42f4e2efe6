"""NumPy prototype / check of the two-stage leaf-angle iteration used by lidf_kernel (exact steps, then a
Taylor model of the map around the hand-over iterate): prints the deviation from the step-by-step
iteration of the oracle and the iteration counts of both stages for several (TAU, degree) choices."""
import numpy as np, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / 'oracle'))
import spart_oracle as so
def lidf_two_stage(a,b,tau=2.5e-4,deg=5):
    a=np.asarray(a,float); b=np.asarray(b,float); n=a.size
    rd=np.pi/180
    thetas=[10*i for i in range(1,9)]+[82,84,86,88]
    F=np.zeros((n,14)); itA=np.zeros(n); itB=np.zeros(n)
    for i,th in enumerate(thetas,1):
        t2=np.full(n,2*rd*th); x=t2.copy(); y=np.zeros(n)
        stage=np.zeros(n,int)   # 0=A running, 1=moved to B, 2=done in A
        # stage A
        while (stage==0).any():
            act=stage==0
            s,c=np.sin(x),np.cos(x)
            ynew=s*(a+b*c)
            dx=0.5*(ynew-x+t2)
            yp=a*c+b*(2*c*c-1)
            xn=x+dx
            done=~(np.abs(dx)>1e-8)
            sw=(np.abs(dx)<=tau*(0.5*(1-yp)))&~done
            y=np.where(act,ynew,y); x=np.where(act,xn,x)
            itA+=act
            stage=np.where(act&done,2,np.where(act&sw,1,stage))
        # stage B for stage==1
        inB=stage==1
        s,c=np.sin(x),np.cos(x); s2=2*s*c; c2=2*c*c-1
        y0=a*s+0.5*b*s2
        y1=a*c+b*c2
        y2=(-a*s-2*b*s2)/2
        y3=(-a*c-4*b*c2)/6
        y4=(a*s+8*b*s2)/24
        y5=(a*c+16*b*c2)/120
        k0=t2-x
        u=np.zeros(n); act=inB.copy(); yt=y0.copy()
        while act.any():
            p=y0+u*(y1+u*(y2+u*(y3+u*(y4+u*y5)))) if deg==5 else y0+u*(y1+u*(y2+u*(y3+u*y4)))
            du=0.5*(p-u+k0)
            yt=np.where(act,p,yt)
            u=np.where(act,u+du,u)
            itB+=act
            act&=np.abs(du)>1e-8
        yfin=np.where(inB,yt,y)
        F[:,i]=(2*yfin+t2)/np.pi
    F[:,13]=1
    return np.diff(F,axis=1), itA, itB
rng=np.random.default_rng(0)
for name,(lo,hi) in (('bench',(-0.5,0.5)),('full',(-1,1))):
    n=100000
    a=rng.uniform(lo,hi,n); b=rng.uniform(lo,hi,n)
    k=np.abs(a)+np.abs(b)<=1; a=a[k]; b=b[k]
    ref=so.leafangles(a,b)
    for tau,deg in ((2.5e-4,5),(1e-3,5),(4e-3,5),(1e-3,4)):
        l,ia,ib=lidf_two_stage(a,b,tau,deg)
        e=np.abs(l-ref)
        print(name,tau,deg,'max abs err %.2e'%e.max(),'n>1e-12:',(e.max(1)>1e-12).sum(),'mean itA %.1f itB %.1f'%(ia.mean(),ib.mean()))
print('---- sweep higher degree')
import math
def lidf_two_stage_deg(a,b,tau,deg):
    a=np.asarray(a,float); b=np.asarray(b,float); n=a.size
    rd=np.pi/180
    thetas=[10*i for i in range(1,9)]+[82,84,86,88]
    F=np.zeros((n,14)); itA=np.zeros(n); itB=np.zeros(n)
    for i,th in enumerate(thetas,1):
        t2=np.full(n,2*rd*th); x=t2.copy(); y=np.zeros(n)
        stage=np.zeros(n,int)
        while (stage==0).any():
            act=stage==0
            s,c=np.sin(x),np.cos(x)
            ynew=s*(a+b*c); dx=0.5*(ynew-x+t2); yp=a*c+b*(2*c*c-1); xn=x+dx
            done=~(np.abs(dx)>1e-8)
            sw=(np.abs(dx)<=tau*(0.5*(1-yp)))&~done
            y=np.where(act,ynew,y); x=np.where(act,xn,x); itA+=act
            stage=np.where(act&done,2,np.where(act&sw,1,stage))
        inB=stage==1
        s,c=np.sin(x),np.cos(x); s2=2*s*c; c2=2*c*c-1
        # derivatives of y = a sin x + b/2 sin 2x: y^(k) = a sin^(k) x + b 2^(k-1) sin^(k)(2x)
        cyc1=[s,c,-s,-c]; cyc2=[s2,c2,-s2,-c2]
        co=[(a*cyc1[k%4]+b*(2.0**(k-1))*cyc2[k%4])/math.factorial(k) for k in range(deg+1)]
        k0=t2-x
        u=np.zeros(n); act=inB.copy(); yt=co[0].copy()
        while act.any():
            p=co[deg]
            for k in range(deg-1,-1,-1): p=p*u+co[k]
            du=0.5*(p-u+k0)
            yt=np.where(act,p,yt); u=np.where(act,u+du,u); itB+=act
            act&=np.abs(du)>1e-8
        yfin=np.where(inB,yt,y)
        F[:,i]=(2*yfin+t2)/np.pi
    F[:,13]=1
    return np.diff(F,axis=1), itA, itB
n=60000
a=rng.uniform(-1,1,n); b=rng.uniform(-1,1,n); k=np.abs(a)+np.abs(b)<=1; a=a[k]; b=b[k]
ref=so.leafangles(a,b)
m=(np.abs(a)<=0.5)&(np.abs(b)<=0.5)
for tau,deg in ((4e-3,5),(1.6e-2,5),(1.6e-2,6),(1.6e-2,7),(5e-2,7),(5e-2,8),(5e-2,9),(0.15,9),(0.15,11),(0.4,13)):
    l,ia,ib=lidf_two_stage_deg(a,b,tau,deg)
    e=np.abs(l-ref)
    cost=ia[m].mean()*65+ib[m].mean()*(deg+12)+12*(60+4*deg)
    print(tau,deg,'max err %.1e'%e.max(),'n>1e-12:',(e.max(1)>1e-12).sum(),'bench itA %.1f itB %.1f'%(ia[m].mean(),ib[m].mean()),'est instr/sample %.0f'%cost)
