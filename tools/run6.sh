mkdir -p gpurun_out
export SRF_TC_PROF=1
for cg in 1 2; do for skip in 0 1; do
  if [ "$cg$skip" = "21" ]; then continue; fi
  echo "== prof cg=$cg skip=$skip"
  SRF_TC_CTA_GROUP=$cg timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rays 60000 --skip-zero-chunks $skip 2>&1 >/dev/null | grep "srf tc prof" | tail -2
done; done
unset SRF_TC_PROF
echo "== pair+skip failure hunt"
timeout 300 python - <<'PY' 2>&1 | tail -15
import sys, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bench
from scenerf_b200 import synth, _lib
from scenerf_b200.renderer import B200Renderer
cfg, pix, desc = bench.workload("B")
pm, pg = synth.make_model_params(cfg)
to_t = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
r = B200Renderer(bench.hp_from_cfg(cfg), to_t(pm), to_t(pg), device="cuda:0", precision="fp16", rng="philox", skip_zero_chunks=True)
x = {k: torch.randn((c,h,w), device="cuda")*0.5 for k,(c,h,w) in zip(synth.SCALE_KEYS, synth.pyramid_shapes(cfg.sphere_W,cfg.sphere_H))}
K,T = torch.from_numpy(cfg.K), torch.from_numpy(cfg.T)
for n in (512, 2048, 8192, 40000):
    try:
        out = r.render_rays_batch(K,T,x,sampled_pixels=torch.from_numpy(pix[:n]), outputs="minimal")
        torch.cuda.synchronize()
        print("n=%d ok depth mean %.3f" % (n, out["depth"].mean().item()))
    except Exception as e:
        print("n=%d FAILED: %s" % (n, str(e).splitlines()[0]), "watchdog flag 0x%x" % (_lib.load().srf_debug_watchdog_flag() & 0xffffffff))
        break
PY
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rays 60000 > gpurun_out/run6_ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/run6_ncu_bench.log
echo "== ncu full capture of the main-pass kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:point_mlp_tc -s 1 -c 1 -o gpurun_out/r1_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rays 60000 > gpurun_out/run6_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/run6_ncu_full.log; ls -la gpurun_out/*.ncu-rep
