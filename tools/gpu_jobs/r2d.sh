set -x
out=gpurun_out/r2d; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=30 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
for n in 1024 128; do
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e > $out/bench_$n.json 2> $out/bench_$n.err
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e --no-pdl > $out/bench_${n}_nopdl.json 2> $out/bench_${n}_nopdl.err
done
timeout 300 python bench.py --images 128 --steps 5 --warmup 3 --no-cpu --no-e2e --no-graph > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_128.csv python bench.py --images 128 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $out/ncu_128.log 2>&1
tail -5 $out/pytest.log
