out=gpurun_out/r2h; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 300 python tools/bench_configs.py > $out/configs.json 2> $out/configs.err
for n in 1024 128; do
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e > $out/bench_$n.json 2> $out/bench_$n.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 150 --csv --log-file $out/launches_128.csv python bench.py --images 128 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $out/ncu_128.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 150 --csv --log-file $out/launches_1024.csv python bench.py --images 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $out/ncu_1024.log 2>&1
