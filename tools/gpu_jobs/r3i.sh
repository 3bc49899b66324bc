out=gpurun_out/r3i; mkdir -p $out
for n in 256 512; do for s in 3 4 3 4; do
timeout 200 python bench.py --images $n --no-cpu --no-e2e --streams $s --steps 30 >> $out/b${n}_s$s.json 2>> $out/b.err
done; done
timeout 200 python bench.py --images 128 --no-cpu --no-e2e --steps 30 > $out/b128_default.json 2>> $out/b.err
