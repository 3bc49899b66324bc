set -x
out=gpurun_out/r2b; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > $out/bench_1024.json 2> $out/bench_1024.err
timeout 300 python bench.py --images 128 --steps 20 --warmup 3 --no-cpu --no-e2e > $out/bench_128.json 2> $out/bench_128.err
tail -5 $out/pytest.log
