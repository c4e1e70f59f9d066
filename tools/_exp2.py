import sys, ctypes; sys.path.insert(0,'/root/repo')
from cocos_b200 import _lib, device
device.init(device_id=0); ctx=device.current_context()
names={0:'ffma',9:'lop3',7:'philox',11:'imad.wide',12:'imad.hi+lo',13:'imad.wide+4ffma',14:'imad.hi',15:'imad.lo',16:'imad.wide+4lop3',17:'imad.hi+lo+4ffma'}
for k,n in names.items():
    ops, ms = ctypes.c_double(0), ctypes.c_float(0)
    _lib.check(_lib.lib().cb_peak_measure(ctx.handle, k, 3, ctypes.byref(ops), ctypes.byref(ms)))
    per = 148*4*1.95e9/ (ops.value/32)   # SMSP cycles per warp-op
    print(f"{n:20s} {ops.value:.3e} ops/s  {ms.value:.3f} ms  -> {per:.2f} SMSP-cycles per warp-level op group")
