#!/usr/bin/env python
"""Cycle stamps of one steady-state tile of the fused field backward (debug build of the library):

    PS_LIB_SUFFIX=_dbg PS_NVCC_DEFS=-DPS_PHASE_CLOCKS python -m presight_b200.build
    PS_LIB_SUFFIX=_dbg python tools/phase_clocks.py > gpurun_out/phase_clocks.txt

Thread 0 of CTA 0 writes (code, clock64) pairs: 9 tile start, 8 inputs staged, 0 epilogue done (before the barrier),
1 barrier passed, 2 GEMM group issued, 5 about to wait, 3 GEMM group complete, 4 weight-gradient group complete,
7 tile done."""
import ctypes as C, os, sys
os.environ.setdefault("PS_LIB_SUFFIX", "_dbg")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from presight_b200 import _lib, synthetic
from presight_b200.cameras.rays import RayBundle
from presight_b200.model import VIDEO_ID

lib = _lib.load()
dev = torch.device("cuda", 0)
bufs = {}
for name in ("ps_debug_phase_buf_bwd2", "ps_debug_phase_buf_bwd", "ps_debug_phase_buf_fwd"):
    if hasattr(lib, name):
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p]
        bufs[name] = torch.zeros(1024, dtype=torch.int64, device=dev)
        assert fn(bufs[name].data_ptr()) == 0
cfg = bench.build_config("c2", "b200")
torch.manual_seed(42)
rays = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
host = synthetic.make_rays(rays, seed=42)
model = bench.build_model("c2", cfg, host, dev).train()
params = [p for p in model.parameters() if p.requires_grad]
keys = ("origins", "directions", "camera_indices", "video_ids", "rgb", "features", "sky")
b = {k: host[k].to(dev) for k in keys}
for _ in range(3):
    for p in params:
        p.grad = None
    rb = RayBundle(origins=b["origins"], directions=b["directions"], camera_indices=b["camera_indices"],
                   metadata={VIDEO_ID: b["video_ids"]})
    model.proposal_sampler._step = 0
    loss = bench.step_loss(model, model(rb), b)
    loss.backward()
torch.cuda.synchronize()
names = {9: "tile start", 8: "inputs staged", 0: "epilogue done", 1: "barrier passed", 2: "group issued",
         5: "about to wait", 3: "group complete", 4: "wgrad group complete", 7: "tile done"}
for name, buf in bufs.items():
    v = buf.cpu().tolist()
    roles = [v[i:i + 256] for i in range(0, len(v), 256)] if name.endswith("bwd2") else [v]
    t0 = min([r[1] for r in roles if r[1]] or [0])
    for k, rv in enumerate(roles):
        print(f"== {name}" + (f" role {k} (0 group R, 1 group S, 2 issuing warp 8, 3 issuing warp 9)" if len(roles) > 1 else ""))
        prev = None
        for i in range(0, len(rv), 2):
            code, clk = rv[i], rv[i + 1]
            if clk == 0:
                break
            print(f"{clk - t0:8d}  +{(clk - prev) if prev is not None else 0:6d}  {code} {names.get(code, '?')}")
            prev = clk
