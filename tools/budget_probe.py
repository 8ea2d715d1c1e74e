import os, sys, json, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg.load()
st = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for L in (26, 28, 30):
    m=1<<L
    src=torch.randint(0,2**62,(m,),dtype=torch.int64,device="cuda"); dst=torch.empty_like(src)
    for mb in (0, 40000):
        plan=lib.plan(L, twist_table_max_mb=mb)
        f=timeit(lambda: plan.forward(dst.data_ptr(),src.data_ptr(),st)); i=timeit(lambda: plan.inverse(dst.data_ptr(),src.data_ptr(),st))
        per=[]
        for p in range(3):
            per.append((round(timeit(lambda: plan.run_pass(p,0,dst.data_ptr(),src.data_ptr(),st))*1e3), round(timeit(lambda: plan.run_pass(p,1,dst.data_ptr(),src.data_ptr(),st))*1e3)))
        print(json.dumps({"L":L,"budget_mb":mb,"fwd_ms":f,"inv_ms":i,"per_pass_us(fwd,inv)":per,"free_gb":torch.cuda.mem_get_info()[0]/2**30}),flush=True)
        plan.close()
    del src,dst
