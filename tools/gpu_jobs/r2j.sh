out=gpurun_out/r2j; mkdir -p $out
timeout 300 python tools/bench_configs.py config3 > $out/c3.json 2> $out/c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 60 --csv --log-file $out/launches_c3.csv python tools/bench_configs.py config3 > $out/ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jb_fwd_large -s 2 -c 1 -o $out/fwd_large python tools/bench_configs.py config3 > $out/ncu_full_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jb_inv_large -s 2 -c 1 -o $out/inv_large python tools/bench_configs.py config3 > $out/ncu_full_c3i.log 2>&1
ls -la $out
