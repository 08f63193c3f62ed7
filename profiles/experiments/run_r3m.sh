# in-kernel split-K for the small-batch eval forward: correctness first (short timeouts: a wrong barrier would hang)
timeout 240 python -m pytest tests/test_inference_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_oracle_model.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python profiles/inference_latency.py 2>&1 | tail -1 | cut -c1-330
ARGUS_EVAL_SPLITK=0 timeout 120 python profiles/inference_latency.py 2>&1 | tail -1 | cut -c1-330
timeout 120 python profiles/profile_eval_detail.py 1 2>&1 | head -12
