set -x
out=gpurun_out/r2e; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
for n in 1024 128; do
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e > $out/bench_$n.json 2> $out/bench_$n.err
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e --streams 1 > $out/bench_${n}_s1.json 2> $out/bench_${n}_s1.err
done
timeout 300 python tools/bench_configs.py > $out/configs.json 2> $out/configs.err
tail -5 $out/pytest.log
