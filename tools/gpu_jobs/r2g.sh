out=gpurun_out/r2g; mkdir -p $out
for args in "0 240 256 both" "0 2160 3840 both" "0 1000 3000 both" "2 1000 3000 both"; do
  echo "== $args" >> $out/dbg.log
  timeout 120 python tools/gpu_jobs/dbg_large.py $args >> $out/dbg.log 2>&1 || echo "FAILED rc=$?" >> $out/dbg.log
done
grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^  File\|^    " $out/dbg.log | tail -30
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -15 $out/pytest.log
timeout 300 python tools/bench_configs.py > $out/configs.json 2> $out/configs.err
