# Round-2 profile captures (run under gpurun; everything it leaves in gpurun_out/ stays small: the .ncu-rep files are exported to
# CSV on the box and never copied back).  Each ncu run is preceded by the same command without ncu.
#   1. launch lists (gpu__time_duration) of the decoder train step / greedy / beam workloads  -> tools/launch_summary.py
#   2. ncu --set full of the key kernels of every bench workload (C2 train, C3 train, greedy C4, beam C5)
#      -> tools/ncu_traffic.py -> profiles/r02_kernel_traffic.json (bench.py's roofline.traffic) + profiles/r02_ncu_full_summary.txt
set -x
D=gpurun_out
C3="--dims 2048,128,256,1024,6400,20,196 --batch 512"
C5="--dims 2048,128,256,512,10000,20,256 --decode 5"
KEY='attention_step_fwd_pipe|attention_step_bwd_pipe|persistent|EpiLstm'
python tools/decoder_step.py --iters 3 > $D/plain_train.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/r02_train_launches.csv python tools/decoder_step.py --iters 3 > $D/ncu_a.log 2>&1
python tools/decoder_step.py --iters 3 --decode 1 --batch 1024 > $D/plain_greedy.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/r02_greedy_launches.csv python tools/decoder_step.py --iters 3 --decode 1 --batch 1024 > $D/ncu_b.log 2>&1
python tools/decoder_step.py --iters 3 $C5 > $D/plain_beam.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/r02_beam_launches.csv python tools/decoder_step.py --iters 3 $C5 > $D/ncu_c.log 2>&1
# --set full: one whole iteration's worth of the key kernels (the first iteration is skipped)
python tools/decoder_step.py --iters 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$KEY" -s 62 -c 62 -o /tmp/full_train python tools/decoder_step.py --iters 2 > $D/ncu_d.log 2>&1
ncu -i /tmp/full_train.ncu-rep --page raw --csv > $D/r02_full_train_raw.csv 2> $D/ncu_e.log
python tools/decoder_step.py --iters 2 $C3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$KEY" -s 62 -c 62 -o /tmp/full_c3 python tools/decoder_step.py --iters 2 $C3 > $D/ncu_f.log 2>&1
ncu -i /tmp/full_c3.ncu-rep --page raw --csv > $D/r02_full_c3_raw.csv 2>> $D/ncu_e.log
python tools/decoder_step.py --iters 2 --decode 1 --batch 1024 > /dev/null 2>&1 && \
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$KEY" -s 124 -c 36 -o /tmp/full_greedy python tools/decoder_step.py --iters 2 --decode 1 --batch 1024 > $D/ncu_g.log 2>&1
ncu -i /tmp/full_greedy.ncu-rep --page raw --csv > $D/r02_full_greedy_raw.csv 2>> $D/ncu_e.log
python tools/decoder_step.py --iters 2 $C5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"attention_step_fwd_group|EpiLstm|gemm_tn_tc_kernel<128|row_topk" -s 125 -c 40 -o /tmp/full_beam python tools/decoder_step.py --iters 2 $C5 > $D/ncu_h.log 2>&1
ncu -i /tmp/full_beam.ncu-rep --page raw --csv > $D/r02_full_beam_raw.csv 2>> $D/ncu_e.log
ls -la /tmp/*.ncu-rep $D/
