import numpy as np, torch, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import mulit_view_object_detection_b200 as m, oracle
from helpers import to_dev
for (B,X,Y,Z,C) in [(1,4,4,8,64),(1,4,4,8,128),(1,4,4,8,256)]:
    rng=np.random.default_rng(1); F=C
    W=(rng.standard_normal((3,3,3,C+F,4*F))*np.sqrt(2.0/(27*(C+F)+4*F))).astype(np.float32)
    b=rng.normal(0,0.1,4*F).astype(np.float32)
    x0=rng.standard_normal((B,X,Y,Z,C)).astype(np.float32); x1=rng.standard_normal((B,X,Y,Z,C)).astype(np.float32)
    dW,db,dx0,dx1=to_dev(W,b,x0,x1)
    cell=m.ConvLSTMTensorCore(dW,db,1.0)
    z=np.zeros((B,X,Y,Z,F),np.float32)
    h1,c1=cell.step(dx0,None,None,relu_in=True)
    oh1,oc1=oracle.convlstm_cell_step(np.maximum(x0,0),z,z,W,b)
    hf,cf=m.convlstm_step(dx0,None,None,dW,db,relu_in=True)
    e=np.abs(c1.cpu().numpy()-oc1); ef=np.abs(cf.cpu().numpy()-oc1)
    print(C,'step1 K=',27*C,'tc max',e.max(),'mean',e.mean(),'signed mean',(np.abs(c1.cpu().numpy())-np.abs(oc1)).mean(),' fp32kernel max',ef.max(),'mean',ef.mean())
    h2,c2=cell.step(dx1,h1,c1)
    oh2,oc2=oracle.convlstm_cell_step(x1,oc1,oh1,W,b)
    e=np.abs(c2.cpu().numpy()-oc2)
    print(C,'step2 K=',27*2*C,'tc max',e.max(),'mean',e.mean(),'signed mean',(np.abs(c2.cpu().numpy())-np.abs(oc2)).mean())
