import numpy as np, sys
sys.path.insert(0,'.')
import qkan_implementation_b200 as Q
from qkan_implementation_b200.fable import fable
errs=[]
cheb=Q.ChebyshevStep(8)
for seed in range(40):
    x=np.random.default_rng(seed).uniform(-1,1,4)
    A=cheb.create_dilated_chebyshev(x,1)
    circ,alpha=fable(A,0)
    blk=circ.block().real*alpha*4
    errs.append(np.linalg.norm(blk-A)/np.linalg.norm(A))
print("cheb block-encoding rel err: max %.3e median %.3e"%(max(errs),np.median(errs)))
