out=gpurun_out/r2k; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 300 python tools/bench_configs.py > $out/configs.json 2> $out/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 60 --csv --log-file $out/launches_c3.csv python tools/bench_configs.py config3 > $out/ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jb_fwd_large -s 2 -c 1 -o $out/fwd_large python tools/bench_configs.py config3 > $out/ncu_full_c3.log 2>&1
