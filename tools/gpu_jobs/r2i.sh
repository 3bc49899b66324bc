out=gpurun_out/r2i; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 -k "config5 or band_sharded or full_size or golden or grid or large_block" > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 300 python tools/bench_configs.py > $out/configs.json 2> $out/configs.err
