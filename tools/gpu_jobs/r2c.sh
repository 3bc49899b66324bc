set -x
out=gpurun_out/r2c; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=30 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
for n in 1024 128; do
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e > $out/bench_$n.json 2> $out/bench_$n.err
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e --step-graph > $out/bench_${n}_sg.json 2> $out/bench_${n}_sg.err
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e --step-graph --pdl > $out/bench_${n}_sg_pdl.json 2> $out/bench_${n}_sg_pdl.err
done
tail -5 $out/pytest.log
