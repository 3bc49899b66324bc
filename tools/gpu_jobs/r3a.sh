out=gpurun_out/r3a; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python bench.py --no-cpu --no-e2e > $out/bench.json 2> $out/bench.err
timeout 300 python bench.py --no-cpu --no-e2e --images 128 > $out/bench128.json 2> $out/bench128.err
timeout 300 python tools/bench_configs.py > $out/r2_configs_1gpu.jsonl 2> $out/configs.err
