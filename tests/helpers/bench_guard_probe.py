"""Executes the tail of bench.run_b200 (everything after the timed region) with stand-ins for the GPU work and an e2e
section that never returns; the guard must print the result line with e2e marked unavailable and exit NON-ZERO (3): a stall is a failure."""
import types, textwrap, time, os, sys
os.environ["B200VTON_E2E_TIMEOUT"] = "1"
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
src=open(os.path.join(ROOT,'bench.py')).read()
start=src.index("    ms_per_step = total_ms / args.steps\n")
end=src.index("def main():")
seg=textwrap.dedent(src[start:end])
import importlib.util
spec=importlib.util.spec_from_file_location("bench_mod",os.path.join(ROOT,"bench.py"))
mod=importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
ns=dict(vars(mod))
class Clk:
    def summary(self): return {"sm_mhz":1,"sm_max_mhz":2,"reasons":[],"samples":1}
def slow_pipeline(*a):
    time.sleep(30)
_args=types.SimpleNamespace(steps=2,warmup=3,no_e2e=False,no_cpu_baseline=True,no_eager_baseline=True,config=2,height=None,width=None,
                            denoise_steps=None,batch=None,requests=None,shared_garment=False)
ns.update(dict(total_ms=2000.0, args=_args, cfg=mod.resolve_config(_args), n_requests=2, Bg=2, T=30, groups=[[0,1]], mine=range(2),
   HEIGHT_=1024, WIDTH_=768, kv_gb=9.4, world=1, B=2, rank=0,
   h=128,w=96, device="cpu", unet=None, unet_enc=None, log=lambda m: print("LOG:",m, file=sys.stderr), clocks=Clk(), per_step_ms=[1000.0,1000.0],
   launches_per_denoise_step=934, eager_launches=7744, bcast_ms=0.0, barrier=lambda: None, den=types.SimpleNamespace(window=30), make_pipeline=slow_pipeline,
   time_dominant_kernel=lambda d,b: dict(kernel="k",ms=0.07,n=20,flops=8e10,tflops=1100.0)))
from idm_vton_b200.engine import SDXL_GARMENT, SDXL_TRYON
ns.update(SDXL_GARMENT=SDXL_GARMENT, SDXL_TRYON=SDXL_TRYON)
exec(compile(seg,"bench_tail","exec"), ns)
print("SHOULD NOT REACH")
