mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 | cut -c1-200
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | cut -c1-300
echo "== default bench"; timeout 900 python bench.py > gpurun_out/r1f_bench_default.json 2> gpurun_out/r1f_bench_default.err; echo "rc=$?"; cut -c1-300 gpurun_out/r1f_bench_default.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1f_bench_reference.json 2> gpurun_out/r1f_bench_reference.err; echo "rc=$?"; cut -c1-200 gpurun_out/r1f_bench_reference.json
for w in A C; do echo "== workload $w"; timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-variants > gpurun_out/r1f_bench_$w.json 2> gpurun_out/r1f_bench_$w.err; echo "rc=$?"; cut -c1-260 gpurun_out/r1f_bench_$w.json; done
echo "== train fp32"; timeout 300 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/r1f_train_fp32.json 2> gpurun_out/r1f_train_fp32.err; echo "rc=$?"; cut -c1-420 gpurun_out/r1f_train_fp32.json
echo "== train tf32"; timeout 300 python bench.py --workload train --train-matmul tf32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_train_tf32.json 2> gpurun_out/r1f_train_tf32.err; echo "rc=$?"; cut -c1-420 gpurun_out/r1f_train_tf32.json
