# Round-2 profile captures (run under gpurun; everything it leaves in gpurun_out/ stays small: the .ncu-rep is exported to CSV
# on the box and deleted).  Each ncu run is preceded by the same command without ncu.
set -x
D=gpurun_out
python tools/decoder_step.py --iters 3 > $D/plain_train.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/r02_train_launches.csv python tools/decoder_step.py --iters 3 > $D/ncu_a.log 2>&1
python tools/decoder_step.py --iters 3 --decode 1 --batch 1024 > $D/plain_greedy.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/r02_greedy_launches.csv python tools/decoder_step.py --iters 3 --decode 1 --batch 1024 > $D/ncu_b.log 2>&1
python tools/decoder_step.py --iters 3 --dims 2048,128,256,512,10000,20,256 --decode 5 > $D/plain_beam.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/r02_beam_launches.csv python tools/decoder_step.py --iters 3 --dims 2048,128,256,512,10000,20,256 --decode 5 > $D/ncu_c.log 2>&1
python tools/decoder_step.py --iters 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:"attention_step_fwd_pipe|attention_step_bwd_pipe|persistent|gemm_nt_tc|EpiLstm|embed_grad|param_grads_finalize|dP_deferred|dann_alpha" -s 60 -c 24 -o /tmp/r02_train_full python tools/decoder_step.py --iters 2 > $D/ncu_d.log 2>&1
ncu -i /tmp/r02_train_full.ncu-rep --page raw --csv > $D/r02_train_full_raw.csv 2> $D/ncu_e.log
ls -la /tmp/r02_train_full.ncu-rep $D/
