"""dev experiment (CPU, scipy): the floating electrode lines of the 10 nm crossbar under the reference GPU branch's contact size
(144 Dirichlet sites per side) as coarse unknowns of the preconditioner, beside the small uncharged-vacancy clusters the
library deflates today.  DESIGN.md 8-5 quotes its output (714 -> 347 textbook PCG iterations at 110 k sites).
    python tools/island_coarse_space_experiment.py"""
import sys, time
sys.path.insert(0, __import__('os').getcwd())
import numpy as np, scipy.sparse as sp, scipy.sparse.csgraph as csg, bench
from oracle import oracle as O
from devicekmc_b200.host import VACANCY
from devicekmc_b200 import structures as S
el,x,y,z,lat,nc_cpu,vd=S.load_structure("crossbar_10nm_5pitch")
d=np.load('devicekmc_b200/data/crossbar_10nm_5pitch.npz'); nc=int(d['n_first_layer'])   # the reference GPU branch's contact size: 144
from devicekmc_b200.host import KMCParameters
p=KMCParameters(lattice=tuple(lat),num_atoms_contact=nc,num_atoms_first_layer=nc)
el=bench.substoichiometric(el,p)
N=len(x)
nb,nn=O.neighbor_list(x,y,z,lat,p.pbc,p.nn_dist,method=1)
q=O.update_charge(nb,el,p.metals,np.zeros(N,np.int32))
cs=O.csr_structure(nb,nc,nc)
val,rhs=O.assemble_K(nb,nc,nc,el,q,p.metals,p.high_G,p.low_G,15.0,cs['row_ptr'],cs['col'])
m=len(rhs)
A=sp.csr_matrix((val,cs['col'],cs['row_ptr']),shape=(m,m))
dinv=1/A.diagonal()
# strongly coupled components of the interior: off-diagonal == -high_G
C=A.copy(); C.setdiag(0); C.eliminate_zeros()
strong=C.multiply(C< -0.5)
ncomp,lab=csg.connected_components(strong!=0,directed=False)
sizes=np.bincount(lab)
# a component is grounded if one of its rows has a high_G link to a contact: diag contribution... detect via rhs != 0 with |rhs| ~ Vd/2*high_G
grounded=np.zeros(ncomp,bool)
hi_contact=np.abs(rhs)>0.5*7.5*0.9   # a high_G link to a contact at +-7.5 V
grounded[np.unique(lab[hi_contact])]=True
comps=[c for c in range(ncomp) if sizes[c]>1 and not grounded[c]]
print('interior rows',m,'floating strongly coupled components',len(comps),'sizes',sorted(sizes[comps])[-6:])
def make_W(cc):
    rows=np.concatenate([np.nonzero(lab==c)[0] for c in cc]) if cc else np.zeros(0,int)
    cols=np.concatenate([np.full((lab==c).sum(),k) for k,c in enumerate(cc)]) if cc else np.zeros(0,int)
    return sp.csr_matrix((np.ones(len(rows)),(rows,cols)),shape=(m,len(cc)))
def pcg(W,tol=1e-12,maxit=6000):
    if W.shape[1]:
        E=(W.T@A@W).tocsc(); Es=sp.linalg.splu(E)
        Minv=lambda v: dinv*v+W@Es.solve(W.T@v)
    else: Minv=lambda v: dinv*v
    xx=np.zeros(m); r=rhs.copy(); zv=Minv(r); pp=zv.copy(); rz=r@zv; bb=rhs@Minv(rhs); it=0
    while rz>tol*tol*bb and it<maxit:
        Ap=A@pp; a=rz/(pp@Ap); xx+=a*pp; r-=a*Ap; zv=Minv(r); rzn=r@zv; pp=zv+(rzn/rz)*pp; rz=rzn; it+=1
    rt=rhs-A@xx
    return it, np.sqrt((rt@Minv(rt))/bb)
import scipy.sparse.linalg
small=[c for c in comps if sizes[c]<100]
t=time.time(); print('Jacobi + small clusters only (what runs today):', pcg(make_W(small)), '%.0fs'%(time.time()-t))
t=time.time(); print('Jacobi + small clusters + floating metal islands:', pcg(make_W(comps)), '%.0fs'%(time.time()-t))
