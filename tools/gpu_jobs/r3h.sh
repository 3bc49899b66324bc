out=gpurun_out/r3h; mkdir -p $out
for s in 3 4 2 3 4; do
timeout 200 python bench.py --images 128 --no-cpu --no-e2e --streams $s --steps 40 >> $out/b128_s$s.json 2>> $out/b128.err
done
for s in 3 4; do
timeout 200 python bench.py --no-cpu --no-e2e --streams $s --steps 20 >> $out/b1024_s$s.json 2>> $out/b1024.err
done
