import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import engine_lib as el
from assistedmanipulation_b200 import abi
def run(K, prec, n=14):
    h=abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_TRACK_POINT,K,0.64,precision=prec,dynamics_mode=abi.DYNAMICS_FUSED)
    e=el.Engine(h,abi.default_track_point()); x0=abi.huddled_state(); ds=[]
    for u in range(n):
        assert e.update(x0,0.05*u,None,seed=5)==0
        ds.append(e.device_seconds())
    e.close(); return np.median(ds[4:])*1e6
for prec,name in ((abi.FP64,'f64'),(abi.FP32,'f32')):
    print(sys.argv[1], name, ' '.join('K%d=%.0f' % (K, run(K, prec)) for K in (4096, 8192, 16384, 32768, 65536, 131072)), flush=True)
