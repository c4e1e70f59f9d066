import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'tools')
from tune_fused import *
device.init(device_id=0); ctx=device.current_context()
for v in (1,4,5):
    for thr,ppt in ((256,1),(128,1),(256,2)):
        ms,mean=time_config(ctx,10_000_000,501,'f32',[0.95],thr,0,ppt,variant=v)
        print('variant',v,thr,ppt,round(ms,3),'ms', mean)
