mkdir -p gpurun_out
echo "== full default bench"
timeout 900 python bench.py > gpurun_out/r1_bench_default.json 2> gpurun_out/r1_bench_default.err; echo "rc=$?"; cut -c1-600 gpurun_out/r1_bench_default.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err; echo "rc=$?"; cut -c1-400 gpurun_out/r1_bench_reference.json
echo "== ncu launch list (full config B, 1 timed step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r1_ncu_launch_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full capture (main pass, full config B)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:point_mlp_tc -s 1 -c 1 -o gpurun_out/r1_prof_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r1_ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r1_prof_final.ncu-rep
