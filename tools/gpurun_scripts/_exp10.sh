cd $GRAFT_REPO_ROOT
python -m pytest tests/test_sched_prepost_gpu.py tests/test_frontend_gpu.py -x -q 2>&1 | tail -4
python tools/ncu_targets_r02.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_igemm|gn_apply|conv_in_direct|attention_hd8|guidance_stream' -s 19 -c 6 -o gpurun_out/r02_full python tools/ncu_targets_r02.py > gpurun_out/ncu_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
