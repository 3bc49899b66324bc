out=gpurun_out/r2l; mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 300 python tools/bench_configs.py > $out/configs.json 2> $out/configs.err
for n in 1024 128; do
for st in 2 3; do
timeout 300 python bench.py --images $n --steps 20 --warmup 3 --no-cpu --no-e2e --streams $st > $out/bench_${n}_s$st.json 2> $out/bench_${n}_s$st.err
done; done
