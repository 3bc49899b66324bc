out=gpurun_out/r3j; mkdir -p $out
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "rc=$?"
tail -c 300 $out/bench_default.json
