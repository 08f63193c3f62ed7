# one-launch small-batch inference tail: correctness, then batch-1 latency
timeout 300 python -m pytest tests/test_inference_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_oracle_model.py tests/test_sync_weights_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python profiles/inference_latency.py 2>&1 | tail -1 | cut -c1-330
timeout 120 python profiles/profile_eval_detail.py 1 2>&1 | grep -E "total|head|avgpool|M2_"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
