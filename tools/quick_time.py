import sys, time, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import engine_lib as el
from assistedmanipulation_b200 import abi
import cases
def run(label, system, obj, params, K, hor, prec, mode, x0, wrench=None, n=30):
    h=abi.make_config(system,obj,K,hor,precision=prec,dynamics_mode=mode)
    e=el.Engine(h,params)
    ts=[];ds=[]
    for u in range(n):
        t0=time.perf_counter(); rc=e.update(x0,0.05*u,wrench,seed=5); t1=time.perf_counter()
        assert rc==0,e.error()
        ts.append(t1-t0); ds.append(e.device_seconds())
    ts=np.array(ts[5:]);ds=np.array(ds[5:])
    T=e.query(abi.QUERY_STEP_COUNT)
    print(f"{label}: K={K} T={T} wall p50 {np.median(ts)*1e6:.1f}us p99 {np.percentile(ts,99)*1e6:.1f}us device p50 {np.median(ds)*1e6:.1f}us -> {(K+2)*T/np.median(ds)/1e6:.1f} M rollout-steps/s", flush=True)
    e.close()
x=abi.huddled_state()
run('toy f64',abi.SYSTEM_TOY,abi.OBJECTIVE_TOY,abi.default_toy_objective(),1024,1.0,abi.FP64,0,np.zeros(4))
for mode in (1,0):
    run(f'cfg2 TP f64 mode{mode}',abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_TRACK_POINT,abi.default_track_point(),4096,0.64,abi.FP64,mode,x)
    run(f'cfg2 TP f32 mode{mode}',abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_TRACK_POINT,abi.default_track_point(),4096,0.64,abi.FP32,mode,x)
run('cfg3 AM f32 fused',abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_ASSISTED_MANIPULATION,cases.assisted_params(True,1),16384,1.28,abi.FP32,1,abi.huddled_state(10.0),cases.constant_wrench(128))
run('cfg3 AM f64 fused',abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_ASSISTED_MANIPULATION,cases.assisted_params(True,1),16384,1.28,abi.FP64,1,abi.huddled_state(10.0),cases.constant_wrench(128),n=12)
run('TP f32 K=131072',abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_TRACK_POINT,abi.default_track_point(),131072,0.64,abi.FP32,1,x,n=10)
run('TP f64 K=131072',abi.SYSTEM_FRANKA_RIDGEBACK,abi.OBJECTIVE_TRACK_POINT,abi.default_track_point(),131072,0.64,abi.FP64,1,x,n=8)
