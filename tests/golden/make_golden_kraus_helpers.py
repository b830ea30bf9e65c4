"""Golden values for pytdscf_b200/kraus.py from the unmodified reference (pytdscf/kraus.py:17-124, :434-470); build container only."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle.reference_loader import load_reference
load_reference()
from pytdscf.kraus import lindblad_to_kraus as ref_l2k, trace_kraus_dim as ref_tr
import pytdscf_b200.kraus as k
rng=np.random.default_rng(5)
out={}
# (real jump operators only: the reference's own consistency assertion fails for complex L, its order="F" reshapes conjugate the map)
for name,d,nl,dt in (("spin1",3,2,0.3),("qubit_real",2,3,0.7),("d4",4,1,0.05)):
    Ls=[rng.standard_normal((d,d))*0.4+0j for _ in range(nl)]
    if name=="spin1": Ls=[np.diag([1.0,0.0,-1.0])*0.5, np.diag([1.0,1.0],1)*0.3]
    Bref=np.asarray(ref_l2k([L.copy() for L in Ls], dt))
    B=k.lindblad_to_kraus(Ls, dt)
    G1=sum(np.kron(b,b.conj()) for b in Bref); G2=sum(np.kron(b,b.conj()) for b in B)
    print(name, Bref.shape, B.shape, np.abs(G1-G2).max(), np.abs(sum(b.conj().T@b for b in B)-np.eye(d)).max())
    out[name+"_L"]=np.stack(Ls); out[name+"_dt"]=np.array(dt); out[name+"_map"]=G1; out[name+"_k"]=np.array(Bref.shape[0])
r=rng.standard_normal((5,6,6)); print(np.abs(ref_tr(r,3)-k.trace_kraus_dim(r,3)).max(), np.abs(ref_tr(r[0],2)-k.trace_kraus_dim(r[0],2)).max())
out["tr_in"]=r; out["tr_out3"]=ref_tr(r,3); out["tr_out2"]=ref_tr(r[0],2)
np.savez_compressed('/root/repo/tests/golden/kraus_helpers.npz', **out)
