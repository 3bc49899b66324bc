out=gpurun_out/r3k; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python bench.py --no-cpu --no-e2e > $out/bench.json 2> $out/bench.err
timeout 300 python tools/bench_configs.py > $out/r2_configs_1gpu.jsonl 2> $out/configs.err
B="python bench.py --images 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph"
F="ncu --set full --clock-control none --import-source on"
timeout 900 $F -k regex:jb_fwd_fast -s 4 -c 1 -o $out/fwd_fast_1024 $B > $out/ncu5.log 2>&1
timeout 900 $F -k regex:jb_inv_fast -s 4 -c 1 -o $out/inv_fast_1024 $B > $out/ncu6.log 2>&1
timeout 900 $F -k regex:jb_fwd_fast -s 2 -c 1 -o $out/fwd_fast_dft_c5 python tools/bench_configs.py config5 > $out/ncu9.log 2>&1
N="ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ --csv"
timeout 600 $N -c 150 --log-file $out/r2_launches_ncu_1024img.csv $B > $out/ncu1.log 2>&1
