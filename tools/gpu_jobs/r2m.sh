out=gpurun_out/r2m; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 -k "no_write or dependent_launch or malformed" > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
for n in 1024 128; do
for st in 3 4; do
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e --streams $st > $out/bench_${n}_s$st.json 2> $out/bench_${n}_s$st.err
done; done
B="python bench.py --images 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 150 --csv --log-file $out/launches_1024.csv $B > $out/ncu_1024.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 150 --csv --log-file $out/launches_128.csv python bench.py --images 128 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $out/ncu_128.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jb_fwd_fast -s 4 -c 1 -o $out/fwd_fast_1024 $B > $out/ncu_full_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jb_inv_fast -s 4 -c 1 -o $out/inv_fast_1024 $B > $out/ncu_full_inv.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ -c 80 --csv --log-file $out/launches_c5.csv python tools/bench_configs.py config5 > $out/ncu_c5.log 2>&1
ls -la $out
