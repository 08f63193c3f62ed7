python profiles/profile_eval_detail.py 1 2>&1 | head -44
