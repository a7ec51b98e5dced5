import sys, time, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import engine_lib as el
from assistedmanipulation_b200 import abi
import ctypes as C
def run(label, K, prec, n=40):
    h=abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_TRACK_POINT,K,0.64,precision=prec,dynamics_mode=abi.DYNAMICS_FUSED)
    e=el.Engine(h,abi.default_track_point()); x0=abi.huddled_state()
    e.lib.mppi_b200_set_profiling(e.h,1)
    st=np.zeros((n,8)); ds=[]
    for u in range(n):
        assert e.update(x0,0.05*u,None,seed=5)==0
        e.lib.mppi_b200_stage_seconds(e.h, st[u].ctypes.data_as(C.POINTER(C.c_double)), 8); ds.append(e.device_seconds())
    m=np.median(st[5:],axis=0)*1e6
    print(label, 'K',K,'device us %.1f'%(np.median(ds[5:])*1e6), ' '.join('%s=%.1f'%(a,b) for a,b in zip(abi.STAGES,m)), flush=True)
    e.close()
run('f64',4096,abi.FP64); run('f32',4096,abi.FP32); run('f64',131072,abi.FP64,12); run('f32',131072,abi.FP32,12)
