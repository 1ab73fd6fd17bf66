import importlib, os, sys
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import ctypes as C
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
import bench_workloads as bw
lm = nsagp._lib; L = lm.lib()
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); lm.check(L.nsagp_set_device(lr))
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if os.environ.get("DBG_LEGACY") == "1":
    lm.check(L.nsagp_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
T, itts = 200000, 3
rng = np.random.default_rng(7)
hyp = nsagp.synth.speech_hypers(16, 3, rng)
y = bw._tiled_signal(nsagp, hyp, "exp", "matern52", T, 7000)
mdl, tabs = bw._model(nsagp, hyp, "exp", "matern52", 16, 3, False)
damp = np.linspace(0.01, 0.1, itts)
plan = nsagp.Plan(lm.KIND_FULL, [mdl], [(bw._mom(nsagp), np.log([hyp.w_lik]), hyp.W)], 0.75, damp, itts, y[None, :], lm.MODE_PREDICT)
dc = nsagp.chunked.DeviceComm.connect_torch(plan)
for par in (None, (16, 20000)):
    if par: plan.set_adf_parallel(*par)
    nsagp.chunked.run_chunked_device(plan, dc)
    print("rank", rank, "par", par, "nlZ", [repr(float(v)) for v in plan.fetch(0, ("nlZ",))["nlZ"]], "mismatch", plan.adf_mismatch(),
          "adf ms", plan.timings()["adf"], flush=True)
dist.barrier()
dist.destroy_process_group()
