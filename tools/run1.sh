mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/run1_gpu.txt 2>&1
nproc >> gpurun_out/run1_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/run1_gpu.txt
python -m pytest tests -m gpu -x -q -k "fp32 or empty or properties" -s > gpurun_out/run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run1_pytest.log
tail -30 gpurun_out/run1_pytest.log
python - <<'PY' > gpurun_out/run1_time.log 2>&1
import sys, time, torch
sys.path.insert(0,'tests')
from cases import RENDER_CASES
from helpers import make_renderer, torch_pyramid
from scenerf_b200 import synth
cfg = synth.config_A()
import numpy as np
r = make_renderer(cfg, "fp32")
x = {k: torch.randn(c,h,w, device='cuda')*0.5 for k,(c,h,w) in zip(synth.SCALE_KEYS, synth.pyramid_shapes(cfg.sphere_W,cfg.sphere_H))}
pix = torch.from_numpy(synth.random_pixels(1, 1024, cfg.img_W, cfg.img_H)).cuda()
K,T = torch.from_numpy(cfg.K), torch.from_numpy(cfg.T)
for i in range(2): r.render_rays_batch(K,T,x,sampled_pixels=pix, outputs="minimal")
torch.cuda.synchronize(); t=time.time()
for i in range(3): r.render_rays_batch(K,T,x,sampled_pixels=pix, outputs="minimal")
torch.cuda.synchronize(); dt=(time.time()-t)/3
print("config A fp32 SIMT: %.1f ms/call -> %.0f rays/s, launches %d" % (dt*1e3, 1024/dt, r.last_launches))
PY
cat gpurun_out/run1_time.log
