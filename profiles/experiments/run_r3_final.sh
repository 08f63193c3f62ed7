# final validation of the round: GPU tests, default bench line, launch list, per-layer profile
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_r3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r3.log; tail -2 gpurun_out/t_r3.log
timeout 600 python bench.py > gpurun_out/b_r3.json 2> gpurun_out/b_r3.err; echo "bench rc=$?"
python profiles/one_step.py 4 > gpurun_out/plain_r3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 480 --csv --log-file gpurun_out/launches_r3.csv \
    python profiles/one_step.py 4 > gpurun_out/ncu_list_r3.log 2>&1
ARGUS_PROFILE_DETAIL=1 python profiles/profile_detail.py > gpurun_out/detail_r3.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
