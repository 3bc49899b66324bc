out=gpurun_out/r2f; mkdir -p $out
for args in "0 240 256 both" "0 120 128 both" "2 2160 3840 fwd" "4 2160 3840 fwd" "0 2160 3840 coef" "0 2160 3840 fwd" "0 360 3840 fwd" "0 120 3840 fwd" "0 120 1024 fwd" "0 2160 3840 both"; do
  echo "== $args" >> $out/dbg.log
  timeout 120 python tools/gpu_jobs/dbg_large.py $args >> $out/dbg.log 2>&1 || echo "FAILED rc=$?" >> $out/dbg.log
done
grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^  File\|^    " $out/dbg.log | tail -60
