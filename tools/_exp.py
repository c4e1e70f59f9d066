import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'tools')
from tune_fused import *
device.init(device_id=0); ctx=device.current_context()
for v in (0,2,1):
    for thr,ppt,bps in ((128,1,0),(256,1,0),(128,2,0),(256,2,0),(128,1,6),(128,1,8),(64,1,0),(128,4,0)):
        try:
            ms,mean=time_config(ctx,10_000_000,501,'f32',[0.95],thr,bps,ppt,variant=v)
            print('variant',v,thr,ppt,bps,round(ms,3),'ms', '%.3e'%(1e7*500/ms*1e3), mean)
        except Exception as e: print('skip',v,thr,ppt,bps,e)
