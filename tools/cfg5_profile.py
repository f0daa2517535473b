import re,sys
src=open('/root/repo/tools/configs_bench.py').read()
i=src.index('# ---- config 5')
head=src[:src.index('# ---- config 3')]
exec(compile(head,'head','exec'))
body=src[i:src.index('res = run5(); res = run5()')]
exec(compile(body,'body','exec'))
import time
run5(); run5()
t0=time.perf_counter()
for _ in range(10):
    for i in range(200): pool5.reset(i)
ctx.synchronize()
print("200 resets: %.3f ms"%((time.perf_counter()-t0)/10*1e3))
t0=time.perf_counter()
for _ in range(10): ctx.decode_tb_batch(pool5, tbs5, 10)
print("decode_tb_batch only (cb_crc set -> all blocks skipped): %.3f ms"%((time.perf_counter()-t0)/10*1e3))
# C-level: prebuild descriptors once
arr=(pkg.TbDesc*200)()
keep=[]
for i,d in enumerate(tbs5):
    e=np.ascontiguousarray(d["e_bits"],dtype=np.int16); out=np.zeros(d["tbs"]//8+8,np.uint8); keep+= [e,out]
    arr[i]=pkg.TbDesc(d["tbs"],d["qm"],d["rv"],e.shape[0],d["softbuffer"],e.ctypes.data,out.ctypes.data,0,0.0)
L=pkg.lib()
def c_only():
    for i in range(200): L.srslte_b200_harq_reset(ctx._h, pool5._p, i)
    L.srslte_b200_decode_tb_batch(ctx._h, pool5._p, arr, 200, 10)
c_only(); c_only()
t0=time.perf_counter()
for _ in range(20): c_only()
print("C calls only (200 resets + decode): %.3f ms"%((time.perf_counter()-t0)/20*1e3))
t0=time.perf_counter()
for _ in range(20):
    for i in range(200): L.srslte_b200_harq_reset(ctx._h, pool5._p, i)
ctx.synchronize()
print("C resets only: %.3f ms"%((time.perf_counter()-t0)/20*1e3))
ctx.enable_timing(True)
for _ in range(5): c_only()
ctx.synchronize()
for k,name in ((0,'W16'),(1,'W8'),(2,'gen'),(3,'layout'),(4,'rm')):
    ms,n=ctx.kernel_time(k); print(name, "%.3f ms per call over %d launches"%(ms/5, n))
