out=gpurun_out/r3e; mkdir -p $out
for s in 32 64 128 48; do
timeout 300 python bench.py --no-cpu --e2e-sub $s > $out/bench_sub$s.json 2> $out/bench_sub$s.err
done
