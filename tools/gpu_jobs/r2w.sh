out=gpurun_out/r2w; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python tools/bench_configs.py > $out/r2_configs_1gpu.jsonl 2> $out/configs.err
timeout 300 python bench.py --no-cpu --no-e2e > $out/bench.json 2> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ --csv -c 120 --log-file $out/r2_launches_ncu_config5.csv python tools/bench_configs.py config5 > $out/ncu5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ --csv -c 200 --log-file $out/r2_launches_ncu_single.csv python tools/bench_configs.py single > $out/ncu1.log 2>&1
