set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench4.json 2> gpurun_out/f_bench4.err; tail -2 gpurun_out/f_bench4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_ref4.json 2> gpurun_out/f_ref4.err; tail -2 gpurun_out/f_ref4.err
timeout 600 python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/f_bench2.json 2> gpurun_out/f_bench2.err; tail -2 gpurun_out/f_bench2.err
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/f_bench3.json 2> gpurun_out/f_bench3.err; tail -2 gpurun_out/f_bench3.err
for c in 2 3; do timeout 120 python tools/layer_profile.py --config $c > gpurun_out/f_lp$c.txt 2>&1; done
timeout 120 python tools/layer_profile.py --config 4 --dense > gpurun_out/f_lp4_dense.txt 2>&1
timeout 120 python tools/layer_profile.py --config 4 > gpurun_out/f_lp4_gated.txt 2>&1
head -1 gpurun_out/f_lp*.txt
python tools/profile_net.py face_landmark 1024 3 > gpurun_out/pn_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tcb_dwpw -s 38 -c 1 -o gpurun_out/r2_tcb_final5 -f python tools/profile_net.py face_landmark 1024 3 > gpurun_out/pn_ncu.log 2>&1
ls -la gpurun_out/r2_tcb_final5.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py 1024 > gpurun_out/ps_ncu2.log 2>&1
timeout 300 python bench.py --impl reference --config 2 --steps 3 --warmup 1 > gpurun_out/f_ref2.json 2> gpurun_out/f_ref2.err
timeout 300 python bench.py --impl reference --config 3 --steps 3 --warmup 1 > gpurun_out/f_ref3.json 2> gpurun_out/f_ref3.err
