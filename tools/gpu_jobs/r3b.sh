out=gpurun_out/r3b; mkdir -p $out
B="python bench.py --images 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph"
F="ncu --set full --clock-control none --import-source on"
timeout 900 $F -k regex:jb_frame_stitch -s 4 -c 1 -o $out/stitch_1024 $B > $out/ncu1.log 2>&1
timeout 900 $F -k regex:jb_gather_chunks -s 4 -c 1 -o $out/gather_1024 $B > $out/ncu2.log 2>&1
timeout 900 $F -k regex:jb_frame_prep -s 4 -c 1 -o $out/prep_1024 $B > $out/ncu3.log 2>&1
ls -la $out
