out=gpurun_out/r3c; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python bench.py --no-cpu --no-e2e > $out/bench.json 2> $out/bench.err
B="python bench.py --images 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph"
N="ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ --csv"
timeout 600 $N -c 150 --log-file $out/launches_1024.csv $B > $out/ncu1.log 2>&1
timeout 600 $N -c 80 --log-file $out/launches_c5.csv python tools/bench_configs.py config5 > $out/ncu4.log 2>&1
