import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import engine_lib as el
from assistedmanipulation_b200 import abi
def run(K, n):
    h=abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_TRACK_POINT,K,0.64,precision=abi.FP64,dynamics_mode=abi.DYNAMICS_FUSED)
    e=el.Engine(h,abi.default_track_point()); x0=abi.huddled_state(); ds=[]
    for u in range(n):
        assert e.update(x0,0.05*u,None,seed=5)==0
        ds.append(e.device_seconds())
    e.close(); return np.median(ds[5:])*1e6
print(sys.argv[1] if len(sys.argv)>1 else '', 'K4096 %.1f us  K131072 %.1f us' % (run(4096,60), run(131072,14)), flush=True)
