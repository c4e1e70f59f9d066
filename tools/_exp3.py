import sys, ctypes, math; sys.path.insert(0,'/root/repo')
from cocos_b200 import _lib, device
device.init(device_id=0); ctx=device.current_context(); L=_lib.lib()
HP=_lib.HestonParams(1.0, math.log(1.0319), 6.21, 0.019, 0.61, -0.7, 0.0, 0.101**2)
strikes64=[0.70+0.60*i/63 for i in range(64)]
for name,sfx,R,N,ks in (("f64_1",'f64',12_500_000,253,[0.95]),("f64_64",'f64',12_500_000,253,strikes64),("f32_1",'f32',10_000_000,501,[0.95])):
    k_arr,nK=_lib.strikes_array(ks); fn=getattr(L,f"cb_heston_price_launch_{sfx}")
    for thr,ppt in ((256,1),(128,1),(256,2),(128,2)):
        ctx.set_launch(thr,0,ppt)
        best=1e9
        for rep in range(4):
            ctx.timer_start(); _lib.check(fn(ctx.handle, ctypes.byref(HP), R, N, k_arr, nK, 0, 9, 0, (R+1)//2, None,None,None,None)); ms=ctx.timer_stop()
            if rep: best=min(best,ms)
        print(name,thr,ppt,round(best,3),'ms','%.3e'%(R*(N-1)/best*1e3))
