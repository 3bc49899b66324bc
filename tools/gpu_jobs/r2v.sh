out=gpurun_out/r2v; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python tools/bench_configs.py > $out/r2_configs_1gpu.jsonl 2> $out/configs.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jb_inv_large -s 2 -c 1 -o $out/inv_large_c3 python tools/bench_configs.py config3 > $out/ncu8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jb_fwd_large -s 2 -c 1 -o $out/fwd_large_c3 python tools/bench_configs.py config3 > $out/ncu2.log 2>&1
