out=gpurun_out/r2o; mkdir -p $out
SECONDS=0
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench default took $SECONDS s" >> $out/bench_default.err
tail -3 $out/bench_default.err
timeout 300 python tools/bench_configs.py config5 > $out/c5.json 2> $out/c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 80 --csv --log-file $out/launches_c5.csv python tools/bench_configs.py config5 > $out/ncu_c5.log 2>&1
timeout 300 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 -k "config5 or band_sharded or long_stream or framing or malformed" > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
