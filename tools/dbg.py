import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import engine_lib as el, oracle_lib as ol
from assistedmanipulation_b200 import abi
K,T,nu=1024,100,2
h=abi.make_config(abi.SYSTEM_TOY,abi.OBJECTIVE_TOY,K,1.0,smoothing=None)
o=ol.Oracle(ol.load(),h,abi.default_toy_objective()); e=el.Engine(h,abi.default_toy_objective())
rng=np.random.default_rng(11)
for u in range(3):
    eps=rng.standard_normal((K+2,T,nu))
    o.update(np.zeros(4),0.05*u,None,eps); e.update(np.zeros(4),0.05*u,None,eps)
    a=e.read(abi.READ_NOISE,(K+2)*T*nu).reshape(K+2,T,nu); b=o.read(abi.READ_NOISE,(K+2)*T*nu).reshape(K+2,T,nu)
    bad=np.argwhere(a!=b)
    print('update',u,'nbad',len(bad), bad[:10].tolist(), 'rows', np.unique(bad[:,0])[:10] if len(bad) else None)
    if len(bad):
        i=tuple(bad[0]); print(a[i],b[i],eps[i])
