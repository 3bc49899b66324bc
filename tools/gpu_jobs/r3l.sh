out=gpurun_out/r3l; mkdir -p $out
timeout 600 python bench.py > $out/r2_bench_1gpu.json 2> $out/bench_1gpu.err
timeout 300 python bench.py --images 128 --no-cpu > $out/r2_bench_128img_1gpu.json 2> $out/bench_128.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/r2_bench_reference_arm.json 2> $out/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
N="ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ --csv"
timeout 600 $N -c 150 --log-file $out/r2_launches_ncu_128img.csv python bench.py --images 128 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $out/ncu2.log 2>&1
