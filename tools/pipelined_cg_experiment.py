"""dev experiment (CPU, scipy) on the real K of a tiled device: textbook PCG, Chronopoulos-Gear CG and pipelined CG
(Ghysels-Vanroose) with M^-1 = D^-1 + W E^-1 W^T — iteration counts under the library's restart scheme and the level at
which the TRUE residual of each recurrence stagnates (7e-11 against 9e-13 on tiled_100k).
    python tools/pipelined_cg_experiment.py [tiled_100k]"""
import sys, time
sys.path.insert(0, __import__('os').getcwd())
import numpy as np, scipy.sparse as sp, scipy.sparse.csgraph as csg, bench
from oracle import oracle as O
from devicekmc_b200.host import VACANCY
name=sys.argv[1] if len(sys.argv)>1 else 'tiled_100k'
el,x,y,z,lat,nc,p=bench.workload(name); el=bench.substoichiometric(el,p)
N=len(x)
nb,nn=O.neighbor_list(x,y,z,lat,p.pbc,p.nn_dist,method=1)
q=O.update_charge(nb,el,p.metals,np.zeros(N,np.int32))
cs=O.csr_structure(nb,nc,nc)
val,rhs=O.assemble_K(nb,nc,nc,el,q,p.metals,p.high_G,p.low_G,10.0,cs['row_ptr'],cs['col'])
m=len(rhs)
A=sp.csr_matrix((val,cs['col'],cs['row_ptr']),shape=(m,m))
d=A.diagonal(); dinv=1/d
# clusters: uncharged vacancies adjacent
unch=((el==VACANCY)&(q==0))[nc:N-nc]
idx=np.nonzero(unch)[0]
sub=A[idx][:,idx]; sub=sub-sp.diags(sub.diagonal())
ncomp,lab=csg.connected_components(sub!=0,directed=False)
sizes=np.bincount(lab)
keep=sizes[lab]>1
rows=idx[keep]; labs=lab[keep]
u,inv=np.unique(labs,return_inverse=True)
W=sp.csr_matrix((np.ones(len(rows)),(rows,inv)),shape=(m,len(u)))
E=(W.T@A@W).diagonal()   # clusters are not adjacent to each other -> diagonal
print('m',m,'clusters',len(u),'rows',len(rows))
def Minv(v): return v*dinv + W@((W.T@v)/E)
xs,_=O.solve(cs['row_ptr'],cs['col'],val,rhs,tol=1e-15,refine=4)
def err(xx): return np.abs(xx-xs).max()/np.abs(xs).max()
def pcg(x0,b,tol,maxit=5000):
    xx=x0.copy(); r=b-A@xx; z=Minv(r); pp=z.copy(); rz=r@z; bb=b@Minv(b); it=0
    while rz>tol*tol*bb and it<maxit:
        Ap=A@pp; a=rz/(pp@Ap); xx+=a*pp; r-=a*Ap; z=Minv(r); rzn=r@z; pp=z+(rzn/rz)*pp; rz=rzn; it+=1
    return xx,it
def cgcg(x0,b,tol,maxit=5000):   # Chronopoulos-Gear (current kernel)
    xx=x0.copy(); r=b-A@xx; u=Minv(r); w=A@u; g=r@u; dl=w@u; bb=b@Minv(b); a=g/dl; beta=0; pp=np.zeros(m); s=np.zeros(m); it=0
    while g>tol*tol*bb and it<maxit:
        pp=u+beta*pp; s=w+beta*s; xx+=a*pp; r-=a*s; u=Minv(r); w=A@u; gn=r@u; dl=w@u; beta=gn/g; a=gn/(dl-beta*gn/a); g=gn; it+=1
    return xx,it
def pipe(x0,b,tol,maxit=5000):   # Ghysels-Vanroose pipelined PCG
    xx=x0.copy(); r=b-A@xx; u=Minv(r); w=A@u; bb=b@Minv(b)
    z=np.zeros(m); qv=np.zeros(m); s=np.zeros(m); pp=np.zeros(m); it=0; g_old=1; a_old=1
    while True:
        g=r@u; dl=w@u
        if g<=tol*tol*bb or it>=maxit: break
        mm=Minv(w); n=A@mm
        if it>0: beta=g/g_old; a=g/(dl-beta*g/a_old)
        else: beta=0; a=g/dl
        z=n+beta*z; qv=mm+beta*qv; s=w+beta*s; pp=u+beta*pp
        xx+=a*pp; r-=a*s; u-=a*qv; w-=a*z
        g_old=g; a_old=a; it+=1
    return xx,it
def refined(solver,tol1=None):
    # mimic solve_refined: first solve from warm start, then restarts on true residual
    x0=np.zeros(m); tot=0
    xx,it=solver(x0,rhs,1e-9); tot+=it; out=[(it,err(xx))]
    for k in range(4):
        res=rhs-A@xx   # (double, not double-double)
        e,it=solver(np.zeros(m),res,1e-6); tot+=it; xx=xx+e; out.append((it,err(xx)))
        if err(xx)<1e-13: break
    return tot,out
for nm,sv in (('pcg',pcg),('cg-cg',cgcg),('pipelined',pipe)):
    t=time.time(); print(nm, refined(sv), '%.1fs'%(time.time()-t))
for nm,sv in (('pcg',pcg),('cg-cg',cgcg),('pipelined',pipe)):
    xx,it=sv(np.zeros(m),rhs,1e-13); print(nm,'single solve tol 1e-13: its',it,'err %.2e'%err(xx))
print('--- stagnation of the recurrence residual')
for tol in (1e-9,1e-10,1e-11,1e-12):
    for nm,sv in (('cg-cg',cgcg),('pipelined',pipe)):
        xx,it=sv(np.zeros(m),rhs,tol,maxit=1500)
        rt=rhs-A@xx; print(nm,'tol',tol,'its',it,'true rel res (M-norm) %.2e'%np.sqrt((rt@Minv(rt))/(rhs@Minv(rhs))))
