out=gpurun_out/r2n; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
/usr/bin/time -v timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
tail -3 $out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 python tools/bench_configs.py config5 > $out/c5.json 2> $out/c5.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
