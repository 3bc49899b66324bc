out=gpurun_out/r2q; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python tools/bench_configs.py > $out/configs.json 2> $out/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 80 --csv --log-file $out/launches_c5.csv python tools/bench_configs.py config5 > $out/ncu_c5.log 2>&1
