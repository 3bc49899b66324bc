# final single-GPU evidence of the round: bench lines, config table, launch lists, ncu --set full captures
out=gpurun_out/r3d; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q --maxfail=20 --timeout=150 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 600 python bench.py > $out/r2_bench_1gpu.json 2> $out/bench_1gpu.err
timeout 300 python bench.py --images 128 --no-cpu > $out/r2_bench_128img_1gpu.json 2> $out/bench_128.err
timeout 300 python bench.py --images 128 --no-cpu --no-e2e --streams 1 --call-graphs --no-pdl > $out/r2_bench_128img_1gpu_one_stream_call_graphs.json 2> $out/bench_128b.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/r2_bench_reference_arm.json 2> $out/bench_ref.err
timeout 300 python tools/bench_configs.py > $out/r2_configs_1gpu.jsonl 2> $out/configs.err
timeout 300 python tools/link_probe.py > $out/r2_link_probe_1gpu.json 2> $out/link.err
B="python bench.py --images 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph"
N="ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jb_ --csv"
timeout 600 $N -c 150 --log-file $out/r2_launches_ncu_1024img.csv $B > $out/ncu1.log 2>&1
timeout 600 $N -c 150 --log-file $out/r2_launches_ncu_128img.csv python bench.py --images 128 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $out/ncu2.log 2>&1
timeout 600 $N -c 80 --log-file $out/r2_launches_ncu_config3.csv python tools/bench_configs.py config3 > $out/ncu3.log 2>&1
timeout 600 $N -c 80 --log-file $out/r2_launches_ncu_config5.csv python tools/bench_configs.py config5 > $out/ncu4.log 2>&1
F="ncu --set full --clock-control none --import-source on"
timeout 900 $F -k regex:jb_fwd_fast -s 4 -c 1 -o $out/fwd_fast_1024 $B > $out/ncu5.log 2>&1
timeout 900 $F -k regex:jb_inv_fast -s 4 -c 1 -o $out/inv_fast_1024 $B > $out/ncu6.log 2>&1
timeout 900 $F -k regex:jb_fwd_large -s 2 -c 1 -o $out/fwd_large_c3 python tools/bench_configs.py config3 > $out/ncu7.log 2>&1
timeout 900 $F -k regex:jb_inv_large -s 2 -c 1 -o $out/inv_large_c3 python tools/bench_configs.py config3 > $out/ncu8.log 2>&1
timeout 900 $F -k regex:jb_fwd_fast -s 2 -c 1 -o $out/fwd_fast_dft_c5 python tools/bench_configs.py config5 > $out/ncu9.log 2>&1
timeout 900 $F -k regex:jb_inv_fast -s 2 -c 1 -o $out/inv_fast_dft_c5 python tools/bench_configs.py config5 > $out/ncu11.log 2>&1
timeout 600 $N -c 200 --log-file $out/r2_launches_ncu_single.csv python tools/bench_configs.py single > $out/ncu12.log 2>&1
timeout 900 $F -k regex:jb_frame_walk -s 4 -c 1 -o $out/walk_1024 $B > $out/ncu10.log 2>&1
ls -la $out
